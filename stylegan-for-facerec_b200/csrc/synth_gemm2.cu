// synth_gemm2.cu -- the same implicit GEMM as synth_gemm.cu, issued as tcgen05.mma.cta_group::2:
// a pair of CTAs (one thread-block cluster = one TPC) computes TWO pixel tiles against one weight
// tile; each CTA keeps its own activation tile and only HALF of the weight tile in shared memory,
// the leader CTA issues M=256 MMAs for both, each CTA's TMEM receives its own 128 rows.
// Why: a single-CTA 128x128x16 UMMA reads 8 KB of operands per 64 clk = the whole 128 B/clk shared-
// memory bandwidth of the SM (measured: 35 % tensor-pipe utilisation on the Cout = 128 layers vs
// 71-75 % at N = 256); with cta_group::2 each SM reads A (4 KB) + half of B (2 KB).
// Protocol (differences from synth_gemm.cu):
//   * both producers issue their TMA loads with .cta_group::2 and signal the LEADER's full barrier
//     (barrier address with the CTA-rank bit cleared); the leader arms it with the bytes of both CTAs;
//   * the leader's MMA thread frees a stage / publishes an accumulator with a MULTICAST commit that
//     arrives on the same barrier in both CTAs;
//   * the epilogue warps of both CTAs arrive on the leader's tmem_empty barrier (count 16).
// Everything else (tile geometry, epilogue math) is identical.  Original header follows.
//
// synth_gemm.cu -- ModulatedConv2d as ONE tcgen05/TMEM implicit-GEMM kernel fed by TMA (sm_100a).
//
// Replaces the ~12 PyTorch ops + cuDNN grouped conv of ModulatedConv2d.forward / StyledConv.forward
// (model.py:232-273, 331-337 of the reference).  Formulation:
//     y[b,p,co] = demod[b,co] * sum_{tap,ci} Wp[tap,co,ci] * Xm[b, p+tap, ci]
// with Xm = x * style already applied by the PRODUCER of x (previous layer's epilogue), Wp the
// batch-shared bf16 weights (conv_scale folded in), demod a per-(b,co) epilogue scale.  So the
// whole batch is one GEMM  [M = B*H*W pixels] x [N = Cout] x [K = taps*Cin]  -- no per-sample
// weights, no groups=B.
//
//   A (activations): NHWC bf16, TMA 4-D tiled loads {64 ch, TW, TH, NB}; the filter tap is a
//       coordinate offset and TMA's out-of-bounds zero fill is the convolution's zero padding.
//   B (weights): [tap][Cout][Cin] bf16, TMA 3-D loads {64, BLOCK_N, 1}.  Both land in shared memory
//       as K-major SWIZZLE_128B tiles, which is the canonical UMMA operand layout.
//   D: fp32 accumulators in TMEM, 128 lanes x BLOCK_N columns, double buffered (2 x 256 columns)
//       so the epilogue of tile i overlaps the MMAs of tile i+1.
//   Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM allocator,
//       warps 2..9 = epilogue, two warps per TMEM lane quarter (tcgen05.ld -> demod, +noise, +bias,
//       lrelu, ToRGB partial dot, x next layer's style, bf16 pack, store).  Eight warps because the
//       128-channel layers have only 4.6k cycles of MMA per tile to hide the epilogue under.
//   The transposed stride-2 convolution of the up-sampling layers runs as four polyphase
//   sub-problems (4/2/2/1 taps) of the same kernel: no multiplies by inserted zeros.
//   Persistent: grid = #SMs; every CTA owns a contiguous chunk of each sub-problem's tiles.
#include "common.cuh"
#include "synth_gemm.cuh"
#include "tc_ptx.cuh"

namespace sg2 {

using namespace tc;

// 16 epilogue warps, four per TMEM lane quarter: the epilogue is a chain of dependent latencies (tcgen05.ld -> table
// loads -> math -> staging -> store) and with 8 warps it, not the tensor pipe, paced every layer with N <= 128 and the
// fused up-sampling convs (knock-out runs, round 2: -10 .. -45 % with the epilogue removed)
constexpr int kEpiWarps2 = 4 * kGemm2EpiGroups;
constexpr int kEpiThreads2 = 32 * kEpiWarps2;
constexpr int kGemmThreads2 = 64 + kEpiThreads2;      // TMA warp + MMA warp + epilogue warps
constexpr int kABytes2 = kBlockM * kBlockK * 2;          // 16 KiB
constexpr int kBBytesMax2 = kMaxBlockN * kBlockK * 2;    // 32 KiB
constexpr int kEpiCap2 = 512;                            // NB * BLOCK_N entries of per-sample epilogue params
constexpr float kSlope2 = 0.2f;
// The operand ring is one byte array cut into stages of (16 KiB A + BLOCK_N*128 B of B): narrow
// BLOCK_N means short MMAs per stage, so more stages are needed to cover the TMA latency
// (4 stages at N=256, 6 at N=128, 8 at N<=64).
constexpr int kRingBytes2 = kGemm2RingBytes;
constexpr int kMaxStages2 = 8;

struct __align__(1024) Gemm2Smem {
    uint8_t ring[kRingBytes2];
    uint8_t stg[kEpiWarps2][2048];          // per-warp transposition buffers of the coalesced epilogue store
    float e_demod[kEpiCap2];
    float e_next[kEpiCap2];
    float e_wrgb[3][kEpiCap2];
    float e_bias[kMaxBlockN];
    uint64_t full[kMaxStages2], empty[kMaxStages2];
    uint64_t tmem_full[2], tmem_empty[2];
    uint64_t b_full;              // resident-weights mode: both halves of the weight tensor have landed (leader's barrier)
    uint32_t tmem_base;
};

struct TileCoord2 {
    int nt, x0, y0, b0, dummy;
};

// Tile order inside a sub-problem: x fastest, then y, then the N tile, then the sample block --
// consecutive tiles of a CTA share (sample, N tile), so the staged epilogue parameters are reused.
__device__ __forceinline__ int pair_groups(const GemmSub &g) { return (g.tiles_x * g.tiles_y * g.tiles_b + 1) / 2; }
__device__ __forceinline__ TileCoord2 decode_tile2(const GemmParams &p, const GemmSub &g, int local, int rank) {
    TileCoord2 t;
    const int pg = pair_groups(g), sxyb = g.tiles_x * g.tiles_y * g.tiles_b;
    const int q = local % pg;
    t.nt = local / pg;                 // N tile slowest: both CTAs of the pair always share it
    int sp = 2 * q + rank;
    t.dummy = sp >= sxyb;             // odd tile count: the surplus CTA recomputes the last tile, stores nothing
    if (t.dummy) sp = sxyb - 1;
    const int bx = sp % g.tiles_x;
    sp /= g.tiles_x;
    const int by = sp % g.tiles_y;
    const int bb = sp / g.tiles_y;
    t.x0 = bx * g.TW;
    t.y0 = by * g.TH;
    t.b0 = bb * g.NB;
    return t;
}

// Every CTA takes one contiguous chunk of EVERY sub-problem (the polyphase sub-problems of the
// transposed conv cost 4/2/2/1 taps per tile, so chunking them separately keeps CTAs balanced).
struct TileRange2 { int lo, hi; };
__device__ __forceinline__ TileRange2 cta_range2(const GemmParams &p, const GemmSub &g, int seg) {
    const int count = pair_groups(g) * p.n_tiles_n;          // work items of a CLUSTER
    const int ncl = (int)gridDim.x / 2, cid = (int)blockIdx.x / 2;
    const int per = (count + ncl - 1) / ncl;
    TileRange2 r;
    r.lo = min(count, cid * per);
    r.hi = min(count, r.lo + per);
    if (p.seg_tiles > 0) {             // segment `seg` of the chunk (see GemmParams::seg_tiles)
        r.lo = min(r.hi, r.lo + seg * p.seg_tiles);
        r.hi = min(r.hi, r.lo + p.seg_tiles);
    }
    return r;
}

__device__ __forceinline__ uint32_t pack_bf162(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}

__global__ void __launch_bounds__(kGemmThreads2, 1)
modconv_gemm2_kernel(const __grid_constant__ GemmParams p, const __grid_constant__ CUtensorMap tmA0,
                    const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                    const __grid_constant__ CUtensorMap tmA3, const __grid_constant__ CUtensorMap tmB) {
    extern __shared__ uint8_t smem_raw[];
    Gemm2Smem &sm = *reinterpret_cast<Gemm2Smem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < kMaxStages2; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], 1); }
        // tmem_empty lives in the leader: the epilogue warps of BOTH CTAs arrive on it
        for (int i = 0; i < 2; ++i) { mbar_init(&sm.tmem_full[i], 1); mbar_init(&sm.tmem_empty[i], 2 * kEpiWarps2); }
        mbar_init(&sm.b_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(&sm.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync();                              // peer barriers initialised, peer TMEM allocated
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;
    // K chunk per stage: 64 channels (128-byte rows, SWIZZLE_128B) or, for the 32-channel 1024^2 tail,
    // 32 channels (64-byte rows, SWIZZLE_64B)
    const uint32_t bk = (uint32_t)p.block_k, row_bytes = bk * 2;
    const uint32_t a_stage = kBlockM * row_bytes;
    const uint32_t b_bytes = (uint32_t)(p.block_n / 2) * row_bytes;   // this CTA's half of the weight tile
    const uint32_t kpk = (uint32_t)p.kpack;                  // K chunks per pipeline stage (see synth_gemm.cu)
    // Resident weights (p.resident2): this CTA's half of EVERY weight tile ([tap][K chunk], b_bytes each) is loaded once
    // and stays in front of the ring; the ring then streams activation tiles only.  The layer's operand traffic from L2
    // drops from (A + B/2) to A per tap, which is what bounds the narrow layers.
    const bool resident = p.resident2 != 0;
    const uint32_t ring_off = resident ? (uint32_t)(kGemmMaxTaps * p.kchunks) * b_bytes : 0u;
    const uint32_t stage_bytes = resident ? kpk * a_stage : kpk * (a_stage + b_bytes);   // multiple of 1 KiB: swizzle atoms stay aligned
    const uint32_t nstages = min((uint32_t)kMaxStages2, ((uint32_t)kRingBytes2 - ring_off) / stage_bytes);

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform control flow, one elected lane issues) =====
        uint32_t stage = 0, phase = 0;
        if (resident) {
            if (elect_one()) {
                if (leader) mbar_arrive_expect_tx(&sm.b_full, 2u * ring_off);
                for (int tap = 0; tap < kGemmMaxTaps; ++tap)
                    for (int kc = 0; kc < p.kchunks; ++kc)
                        tma_load_3d_2sm(smem_u32(sm.ring) + (uint32_t)(tap * p.kchunks + kc) * b_bytes, &tmB, &sm.b_full, kc * (int)bk,
                                        rank * (p.block_n / 2), tap);
            }
            __syncwarp();
        }
        for (int seg = 0; seg < p.nseg; ++seg)
        for (int s = 0; s < p.nsub; ++s) {
            const GemmSub &g = p.sub[s];
            const CUtensorMap *tmA = s == 0 ? &tmA0 : (s == 1 ? &tmA1 : (s == 2 ? &tmA2 : &tmA3));
            const uint32_t a_bytes = (uint32_t)(g.TH * g.TW * g.NB) * row_bytes;
            const TileRange2 tr = cta_range2(p, g, seg);
            for (int local = tr.lo; local < tr.hi; ++local) {
                const TileCoord2 t = decode_tile2(p, g, local, rank);
                const int wrow = t.nt * p.block_n + rank * (p.block_n / 2);
                for (int kc = 0; kc < p.kchunks; kc += (int)kpk) {
                    for (int tap = 0; tap < g.ntaps; ++tap) {
                        mbar_wait(&sm.empty[stage], phase ^ 1);
                        const uint32_t slot = smem_u32(sm.ring) + ring_off + stage * stage_bytes;
                        const int ax = t.x0 + g.dx[tap], ay = t.y0 + g.dy[tap], wt = g.wtap[tap];
                        if (p.multi_map) {                    // this tap's own activation tensor
                            const int mi = g.tap_map[tap];
                            tmA = mi == 0 ? &tmA0 : (mi == 1 ? &tmA1 : (mi == 2 ? &tmA2 : &tmA3));
                        }
                        if (SG2_DBG(p) & 6) {                 // bottleneck analysis only (variant build)
                            if (elect_one()) {
                                const bool la = !(SG2_DBG(p) & 2), lb = !(SG2_DBG(p) & 4) && !resident;
                                if (leader) {
                                    if (!la && !lb) mbar_arrive(&sm.full[stage]);
                                    else mbar_arrive_expect_tx(&sm.full[stage], 2 * kpk * ((la ? a_bytes : 0u) + (lb ? b_bytes : 0u)));
                                }
                                for (uint32_t u = 0; u < kpk; ++u) {
                                    if (la) tma_load_4d_2sm(slot + u * a_stage, tmA, &sm.full[stage], (kc + (int)u) * (int)bk, ax, ay, t.b0);
                                    if (lb) tma_load_3d_2sm(slot + kpk * a_stage + u * b_bytes, &tmB, &sm.full[stage], (kc + (int)u) * (int)bk, wrow, wt);
                                }
                            }
                        } else if (resident) {
                            if (elect_one()) {
                                if (leader) mbar_arrive_expect_tx(&sm.full[stage], 2 * kpk * a_bytes);
                                tma_load_4d_2sm(slot, tmA, &sm.full[stage], kc * (int)bk, ax, ay, t.b0);
                                if (kpk == 2) tma_load_4d_2sm(slot + a_stage, tmA, &sm.full[stage], (kc + 1) * (int)bk, ax, ay, t.b0);
                            }
                        } else if (elect_one()) {
                            // the leader's barrier collects the bytes of both CTAs
                            if (leader) mbar_arrive_expect_tx(&sm.full[stage], 2 * kpk * (a_bytes + b_bytes));
                            tma_load_4d_2sm(slot, tmA, &sm.full[stage], kc * (int)bk, ax, ay, t.b0);
                            tma_load_3d_2sm(slot + kpk * a_stage, &tmB, &sm.full[stage], kc * (int)bk, wrow, wt);
                            if (kpk == 2) {
                                tma_load_4d_2sm(slot + a_stage, tmA, &sm.full[stage], (kc + 1) * (int)bk, ax, ay, t.b0);
                                tma_load_3d_2sm(slot + 2 * a_stage + b_bytes, &tmB, &sm.full[stage], (kc + 1) * (int)bk,
                                                wrow, wt);
                            }
                        }
                        __syncwarp();
                        if (++stage == nstages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: the leader CTA's warp, one elected lane issues ============
        if (leader) {
            const uint32_t idesc = make_idesc_bf16(2 * kBlockM, (uint32_t)p.block_n);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            if (resident) {
                mbar_wait(&sm.b_full, 0);
                tc_fence_after();
            }
            for (int seg = 0; seg < p.nseg; ++seg)
            for (int s = 0; s < p.nsub; ++s) {
                const GemmSub &g = p.sub[s];
                const int nstage = p.kchunks / (int)kpk * g.ntaps;
                const TileRange2 tr = cta_range2(p, g, seg);
                for (int local = tr.lo; local < tr.hi; ++local) {
                    mbar_wait(&sm.tmem_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * kMaxBlockN;
                    int tap = 0, kc = 0;                      // the producer's order: K chunk outer, tap inner
                    for (int k0 = 0; k0 < nstage; ++k0) {
                        mbar_wait(&sm.full[stage], phase);
                        tc_fence_after();
                        const uint32_t slot = smem_u32(sm.ring) + ring_off + stage * stage_bytes;
                        const uint64_t adesc = make_smem_desc(slot, row_bytes);
                        const uint64_t bdesc = resident
                            ? make_smem_desc(smem_u32(sm.ring) + (uint32_t)(g.wtap[tap] * p.kchunks + kc) * b_bytes, row_bytes)
                            : make_smem_desc(slot + kpk * a_stage, row_bytes);
                        if (++tap == g.ntaps) { tap = 0; kc += (int)kpk; }
                        if (elect_one()) {
                            // advance 32 bytes (>>4 = 2) inside the swizzle row per K = 16 step
                            umma_bf16_2sm(d_tmem, adesc, bdesc, idesc, k0 != 0);
                            if (!(SG2_DBG(p) & 8)) {
                            umma_bf16_2sm(d_tmem, adesc + 2, bdesc + 2, idesc, 1);
                            if (bk == 64) {
                                umma_bf16_2sm(d_tmem, adesc + 4, bdesc + 4, idesc, 1);
                                umma_bf16_2sm(d_tmem, adesc + 6, bdesc + 6, idesc, 1);
                            }
                            if (kpk == 2) {       // the next K chunk: b_bytes further in the resident layout too
                                const uint64_t a2 = adesc + (a_stage >> 4), b2 = bdesc + (b_bytes >> 4);
                                umma_bf16_2sm(d_tmem, a2, b2, idesc, 1);
                                umma_bf16_2sm(d_tmem, a2 + 2, b2 + 2, idesc, 1);
                                umma_bf16_2sm(d_tmem, a2 + 4, b2 + 4, idesc, 1);
                                umma_bf16_2sm(d_tmem, a2 + 6, b2 + 6, idesc, 1);
                            }
                            }
                            umma_commit_2sm_mc(&sm.empty[stage], 3);    // frees the smem slot in BOTH CTAs
                        }
                        __syncwarp();
                        if (++stage == nstages) { stage = 0; phase ^= 1; }
                    }
                    if (elect_one()) umma_commit_2sm_mc(&sm.tmem_full[acc], 3);   // -> epilogue of both CTAs
                    __syncwarp();
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue: 8 warps, two per TMEM lane quarter =====================
        // warp (2 + 4h + i) may access TMEM lanes 32*((2+i)&3) ..; the two warps of a quarter split the
        // accumulator columns in alternating 32-column chunks (h = 0: even chunks, h = 1: odd chunks).
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;            // column group 0..3: this warp takes the 32-column chunks half, half + 4, ...
        const int m = q * 32 + lane;                 // accumulator row = pixel of the tile
        const int et = threadIdx.x - 64;             // 0..511 within the epilogue group
        const float nw = (p.mode == 0 && p.noise) ? __ldg(p.noise_weight) : 0.f;
        const int N = p.block_n;
        const bool up4 = p.up4 != 0;
        const int cr = p.cout_real;                  // channels of the output tensor (N = 4 * cr column blocks when up4)
        uint32_t acc = 0, acc_phase = 0;
        int staged_key = -1;
        for (int seg = 0; seg < p.nseg; ++seg)
        for (int s = 0; s < p.nsub; ++s) {
            const GemmSub &g = p.sub[s];
            const int per = g.TH * g.TW;
            const int nb = m / per;
            const int rem = m - nb * per;
            const int ty = rem / g.TW, tx = rem - ty * g.TW;
            const int pb = (nb < g.NB ? nb : 0) * N;    // row of the staged per-sample params
            const TileRange2 tr = cta_range2(p, g, seg);
            for (int local = tr.lo; local < tr.hi; ++local) {
                const TileCoord2 t = decode_tile2(p, g, local, rank);
                const int n0 = t.nt * N;
                const int y = t.y0 + ty, x = t.x0 + tx, b = t.b0 + nb;
                const bool valid = !t.dummy && nb < g.NB && y < g.PH && x < g.PW && b < p.B;
                float nz = 0.f, nz1 = 0.f, nz2 = 0.f, nz3 = 0.f;      // up4: one noise value per output phase
                if (valid && p.mode == 0 && p.noise) {   // issued early: overlaps the staging below
                    if (up4) {
                        const float2 *np = reinterpret_cast<const float2 *>(p.noise + (long long)b * p.noise_bstride +
                                                                           (long long)(2 * y) * (2 * g.PW) + 2 * x);
                        const float2 r0 = __ldg(np), r1 = __ldg(np + g.PW);
                        nz = r0.x; nz1 = r0.y; nz2 = r1.x; nz3 = r1.y;
                    } else {
                        nz = __ldg(p.noise + (long long)b * p.noise_bstride + (long long)y * g.PW + x);
                    }
                }
                // ---- stage the per-(sample, channel) epilogue parameters when they change ----
                const int key = (t.b0 * kMaxBlockN + t.nt) * 16 + g.NB;
                if (key != staged_key) {
                    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads2) : "memory");
                    for (int i = et; i < g.NB * N; i += kEpiThreads2) {
                        const int sb = i / N, col = i - sb * N;
                        const int bb = t.b0 + sb < p.B ? t.b0 + sb : p.B - 1;
                        // up4: the four phase blocks of a tile share the per-channel tables
                        const long long o = up4 ? (long long)bb * cr + (n0 + col) % cr : (long long)bb * p.Cout + n0 + col;
                        sm.e_demod[i] = __ldg(p.demod + o);
                        if (p.mode == 0) {
                            // lrelu gain sqrt(2) (fused_bias_act_kernel.cu:47) folded into both consumers
                            sm.e_next[i] = p.next_style ? 1.41421356237f * __ldg(p.next_style + o) : 0.f;
                            if (p.rgb_w) {
                                const float rs = 1.41421356237f * __ldg(p.rgb_style + o);
#pragma unroll
                                for (int c = 0; c < 3; ++c)
                                    sm.e_wrgb[c][i] = rs * __ldg(p.rgb_w + (long long)c * p.Cout + n0 + col);
                            }
                        }
                    }
                    if (p.mode == 0)
                        for (int i = et; i < N; i += kEpiThreads2) sm.e_bias[i] = __ldg(p.bias + (up4 ? (n0 + i) % cr : n0 + i));
                    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads2) : "memory");
                    staged_key = key;
                }
                nz *= nw; nz1 *= nw; nz2 *= nw; nz3 *= nw;
                uint32_t off16 = 0xffffffffu;     // this pixel's row in 16-byte units from p.out (0xFFFFFFFF: not stored)
                if (valid && p.out && !(SG2_DBG(p) & 1))
                    off16 = up4 ? (uint32_t)(((((long long)b * g.out_H + 2 * y) * g.out_W + 2 * x) * cr) >> 3)   // pixel (2y, 2x)
                                : (uint32_t)((g.out_off + (((long long)b * g.out_H + y) * g.out_W + x) * p.Cout + n0) >> 3);

                mbar_wait(&sm.tmem_full[acc], acc_phase);
                tc_fence_after();
                float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kMaxBlockN;
                bool handed_back = false;
                for (int c0 = 32 * half; c0 < N; c0 += 32 * kGemm2EpiGroups) {
                    if (SG2_DBG(p) & 16) break;            // knock-out: no epilogue work at all
                    uint32_t r[32];
                    tmem_ld32(t_row + c0, r);
                    tmem_ld_wait();
                    if (c0 + 32 * kGemm2EpiGroups >= N) {   // this warp's last TMEM read of the tile: the accumulator goes back to
                        tc_fence_before();                  // the MMA thread BEFORE the math and the stores, not after them
                        __syncwarp();
                        if (lane == 0) mbar_arrive_leader(&sm.tmem_empty[acc]);   // the leader's MMA thread waits for both CTAs
                        handed_back = true;
                    }
                    uint32_t packed[16];
                    uint32_t o16 = off16 == 0xffffffffu ? off16 : off16 + (uint32_t)(c0 >> 3);
                    float nzc = nz;
                    if (up4) {            // column block (py, px) of input pixel (y, x) -> output pixel (2y+py, 2x+px)
                        const int ph = (n0 + c0) / cr, chb = (n0 + c0) - ph * cr;
                        nzc = ph == 0 ? nz : (ph == 1 ? nz1 : (ph == 2 ? nz2 : nz3));
                        if (off16 != 0xffffffffu)
                            o16 = off16 + (uint32_t)((((ph >> 1) * g.out_W + (ph & 1)) * cr + chb) >> 3);
                    }
                    if (p.mode == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 d4 = lds128f(smem_u32(&sm.e_demod[pb + c0 + j]));
                            const float4 b4 = lds128f(smem_u32(&sm.e_bias[c0 + j]));
                            const float4 s4 = lds128f(smem_u32(&sm.e_next[pb + c0 + j]));
                            float v[4];
                            v[0] = fmaf(__uint_as_float(r[j + 0]), d4.x, nzc) + b4.x;
                            v[1] = fmaf(__uint_as_float(r[j + 1]), d4.y, nzc) + b4.y;
                            v[2] = fmaf(__uint_as_float(r[j + 2]), d4.z, nzc) + b4.z;
                            v[3] = fmaf(__uint_as_float(r[j + 3]), d4.w, nzc) + b4.w;
#pragma unroll
                            for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], kSlope2 * v[e]);   // lrelu (gain folded downstream)
                            if (p.clamp > 0.f) {
#pragma unroll
                                for (int e = 0; e < 4; ++e) v[e] = fminf(fmaxf(v[e], -p.clamp), p.clamp);
                            }
                            if (p.rgb_w) {
                                const float4 w0 = lds128f(smem_u32(&sm.e_wrgb[0][pb + c0 + j]));
                                const float4 w1 = lds128f(smem_u32(&sm.e_wrgb[1][pb + c0 + j]));
                                const float4 w2 = lds128f(smem_u32(&sm.e_wrgb[2][pb + c0 + j]));
                                rgb0 += v[0] * w0.x + v[1] * w0.y + v[2] * w0.z + v[3] * w0.w;
                                rgb1 += v[0] * w1.x + v[1] * w1.y + v[2] * w1.z + v[3] * w1.w;
                                rgb2 += v[0] * w2.x + v[1] * w2.y + v[2] * w2.z + v[3] * w2.w;
                            }
                            packed[j / 2 + 0] = pack_bf162(v[0] * s4.x, v[1] * s4.y);
                            packed[j / 2 + 1] = pack_bf162(v[2] * s4.z, v[3] * s4.w);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 d4 = lds128f(smem_u32(&sm.e_demod[pb + c0 + j]));
                            packed[j / 2 + 0] = pack_bf162(__uint_as_float(r[j + 0]) * d4.x, __uint_as_float(r[j + 1]) * d4.y);
                            packed[j / 2 + 1] = pack_bf162(__uint_as_float(r[j + 2]) * d4.z, __uint_as_float(r[j + 3]) * d4.w);
                        }
                    }
                    if (p.out)
                        store_rows64_coalesced(smem_u32(sm.stg[warp - 2]), packed, o16, reinterpret_cast<uint8_t *>(p.out), lane);
                }
                if (!handed_back) {                         // a warp without a chunk of this tile (N < 128)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&sm.tmem_empty[acc]);
                }
                const int parts = N >= 32 * kGemm2EpiGroups ? kGemm2EpiGroups : N / 32;   // column groups that own a chunk
                if (valid && p.mode == 0 && p.rgb_w && half < parts) {   // each column group writes its own partial plane
                    const long long plane = (long long)g.PH * g.PW;
                    float *rp = p.rgb_part + ((((long long)t.nt * parts + half) * p.B + b) * 3) * plane + (long long)y * g.PW + x;
                    rp[0] = rgb0;
                    rp[plane] = rgb1;
                    rp[2 * plane] = rgb2;
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync();                              // neither CTA exits (or frees TMEM) while the pair is still working
    if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
}

int launch_modconv_gemm2(const GemmParams &p, const CUtensorMap *tmA, const CUtensorMap &tmB, int sms,
                        cudaStream_t st) {
    static_assert(sizeof(Gemm2Smem) + 1024 <= 227 * 1024, "GemmSmem exceeds the 227 KiB CTA limit");
    static_assert(sizeof(GemmParams) + 5 * sizeof(CUtensorMap) <= 4000, "kernel parameter space");
    const size_t smem = sizeof(Gemm2Smem) + 1024;
    static std::atomic<int> configured{0};
    if (!configured.load(std::memory_order_acquire)) {
        SG2_CUDA_OK(cudaFuncSetAttribute(modconv_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.store(1, std::memory_order_release);
    }
    SG2_REQUIRE(p.block_n % 32 == 0 && p.block_n >= 32 && p.block_n <= kMaxBlockN, SG2_ERR_BAD_ARG, "gemm2: bad BLOCK_N %d", p.block_n);
    SG2_REQUIRE((p.block_k == 64 || p.block_k == 32) && p.Cin % p.block_k == 0 && p.Cout % p.block_n == 0, SG2_ERR_UNSUPPORTED,
                "gemm: Cin %d must be a multiple of BLOCK_K %d and Cout %d of BLOCK_N %d", p.Cin, p.block_k, p.Cout, p.block_n);
    for (int s = 0; s < p.nsub; ++s)
        SG2_REQUIRE(p.sub[s].NB * p.block_n <= kEpiCap2 && p.sub[s].TH * p.sub[s].TW * p.sub[s].NB <= kBlockM,
                    SG2_ERR_BAD_ARG, "gemm: tile of sub-problem %d too large", s);
    if (p.up4)
        SG2_REQUIRE(p.mode == 0 && p.nsub == 1 && p.cout_real % 32 == 0 && p.Cout == 4 * p.cout_real && p.sub[0].NB == 1 &&
                        p.sub[0].out_W == 2 * p.sub[0].PW && !p.rgb_w,
                    SG2_ERR_BAD_ARG, "gemm2: bad fused up-sampling plan (Cout %d, channels %d)", p.Cout, p.cout_real);
    if (p.resident2)
        SG2_REQUIRE(p.n_tiles_n == 1 && kGemmMaxTaps * p.kchunks * (p.block_n / 2) * p.block_k * 2 + 2 * p.kpack * kBlockM * p.block_k * 2 <= kRingBytes2,
                    SG2_ERR_BAD_ARG, "gemm2: resident weights do not fit (BLOCK_N %d, %d K chunks)", p.block_n, p.kchunks);
    long most = 0;
    for (int s2 = 0; s2 < p.nsub; ++s2) {
        const GemmSub &g = p.sub[s2];
        most = std::max(most, ((long)g.tiles_x * g.tiles_y * g.tiles_b + 1) / 2 * p.n_tiles_n);
    }
    const int n_clusters = (int)std::min<long>(most, sms / 2);
    if (n_clusters <= 0) return SG2_OK;
    GemmParams q = p;                      // segments of the polyphase interleave (GemmParams::seg_tiles)
    q.nseg = 1;
    if (q.seg_tiles > 0) {
        long max_per = 1;
        for (int s2 = 0; s2 < q.nsub; ++s2) {
            const GemmSub &g = q.sub[s2];
            const long count = ((long)g.tiles_x * g.tiles_y * g.tiles_b + 1) / 2 * q.n_tiles_n;
            max_per = std::max(max_per, (count + n_clusters - 1) / n_clusters);
        }
        q.nseg = (int)((max_per + q.seg_tiles - 1) / q.seg_tiles);
    }
    SG2_REQUIRE(!p.multi_map || p.nsub == 1, SG2_ERR_BAD_ARG, "gemm2: a multi-tensor tap list needs one sub-problem");
    const int nmaps = p.multi_map ? 4 : p.nsub;
    const CUtensorMap &a0 = tmA[0];
    const CUtensorMap &a1 = tmA[nmaps > 1 ? 1 : 0];
    const CUtensorMap &a2 = tmA[nmaps > 2 ? 2 : 0];
    const CUtensorMap &a3 = tmA[nmaps > 3 ? 3 : 0];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * n_clusters);
    cfg.blockDim = dim3(kGemmThreads2);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SG2_CUDA_OK(cudaLaunchKernelEx(&cfg, modconv_gemm2_kernel, q, a0, a1, a2, a3, tmB));
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

}  // namespace sg2
