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
#include <algorithm>

#include "common.cuh"
#include "synth_gemm.cuh"
#include "tc_ptx.cuh"

namespace sg2 {

using namespace tc;

constexpr int kEpiWarps = 8;
constexpr int kMma2Warp = 2 + kEpiWarps;                // second MMA-issuing warp (resident-weights layers)
constexpr int kGemmThreads = 64 + 32 * kEpiWarps + 32;   // TMA warp + MMA warp + epilogue warps + second MMA warp
constexpr int kABytes = kBlockM * kBlockK * 2;          // 16 KiB
constexpr int kBBytesMax = kMaxBlockN * kBlockK * 2;    // 32 KiB
constexpr int kEpiCap = 512;                            // NB * BLOCK_N entries of per-sample epilogue params
constexpr float kSlope = 0.2f;
// The operand ring is one byte array cut into stages of (16 KiB A + BLOCK_N*128 B of B): narrow
// BLOCK_N means short MMAs per stage, so more stages are needed to cover the TMA latency
// (4 stages at N=256, 6 at N=128, 8 at N<=64).
constexpr int kRingBytes = kGemmRingBytes;
constexpr int kMaxStages = 8;

struct __align__(1024) GemmSmem {
    uint8_t ring[kRingBytes];
    uint8_t stg[kEpiWarps][2048];           // per-warp transposition buffers of the coalesced epilogue store
    float e_demod[kEpiCap];
    float e_next[kEpiCap];
    float e_wrgb[3][kEpiCap];
    float e_bias[kMaxBlockN];
    uint64_t full[kMaxStages], empty[kMaxStages];
    uint64_t tmem_full[4], tmem_empty[4];   // 4 accumulators of 128 columns at BLOCK_N <= 128, else 2 of 256
    uint64_t b_full;              // resident-weights mode: the weight tensor has landed
    uint32_t tmem_base;
};

struct TileCoord {
    int nt, x0, y0, b0;
};

// Tile order inside a sub-problem: x fastest, then y, then the N tile, then the sample block --
// consecutive tiles of a CTA share (sample, N tile), so the staged epilogue parameters are reused.
__device__ __forceinline__ TileCoord decode_tile(const GemmParams &p, const GemmSub &g, int local) {
    TileCoord t;
    const int bx = local % g.tiles_x;
    local /= g.tiles_x;
    const int by = local % g.tiles_y;
    local /= g.tiles_y;
    t.nt = local % p.n_tiles_n;
    const int bb = local / p.n_tiles_n;
    t.x0 = bx * g.TW;
    t.y0 = by * g.TH;
    t.b0 = bb * g.NB;
    return t;
}

// A CTA walks consecutive tiles, so only the first one is decoded with divisions.
__device__ __forceinline__ void next_tile(const GemmParams &p, const GemmSub &g, TileCoord &t) {
    t.x0 += g.TW;
    if (t.x0 >= g.tiles_x * g.TW) {
        t.x0 = 0; t.y0 += g.TH;
        if (t.y0 >= g.tiles_y * g.TH) {
            t.y0 = 0;
            if (++t.nt == p.n_tiles_n) { t.nt = 0; t.b0 += g.NB; }
        }
    }
}

// Every CTA takes one contiguous chunk of EVERY sub-problem (the polyphase sub-problems of the
// transposed conv cost 4/2/2/1 taps per tile, so chunking them separately keeps CTAs balanced).
struct TileRange { int lo, hi; };
__device__ __forceinline__ TileRange cta_range(const GemmParams &p, const GemmSub &g, int seg) {
    const int count = g.tiles_x * g.tiles_y * g.tiles_b * p.n_tiles_n;
    const int per = (count + (int)gridDim.x - 1) / (int)gridDim.x;
    TileRange r;
    r.lo = min(count, (int)blockIdx.x * per);
    r.hi = min(count, r.lo + per);
    if (p.seg_tiles > 0) {             // segment `seg` of the chunk
        r.lo = min(r.hi, r.lo + seg * p.seg_tiles);
        r.hi = min(r.hi, r.lo + p.seg_tiles);
    }
    return r;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}
// 16-byte shared-memory load through the shared window (the staged tables are reached through plain
// pointers, which would otherwise compile to generic loads)
__device__ __forceinline__ float4 lds4(const float *p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}

// staged epilogue parameters of this thread's sample (shared memory), indexed by tile column
struct EpiTables { const float *demod, *next, *w0, *w1, *w2, *bias; };

// W accumulator columns [c0, c0 + W) of one pixel row: TMEM -> registers -> epilogue math -> bf16 -> global.
// off16: destination of this pixel's accumulator row in 16-byte units from p.out (0xFFFFFFFF: not stored)
template <int W>
__device__ __forceinline__ void epi_cols(const GemmParams &p, const EpiTables &e, uint32_t taddr, int c0, float nz,
                                         uint32_t off16, uint32_t stg, int lane, float &rgb0, float &rgb1, float &rgb2) {
    uint32_t r[W];
    tmem_ld_cols(taddr, r);
    tmem_ld_wait();
    uint32_t packed[W / 2];
    if (p.mode == 0) {
#pragma unroll
        for (int j = 0; j < W; j += 4) {
            const float4 d4 = lds4(e.demod + c0 + j), b4 = lds4(e.bias + c0 + j), s4 = lds4(e.next + c0 + j);
            float v[4];
            v[0] = fmaf(__uint_as_float(r[j + 0]), d4.x, nz) + b4.x;
            v[1] = fmaf(__uint_as_float(r[j + 1]), d4.y, nz) + b4.y;
            v[2] = fmaf(__uint_as_float(r[j + 2]), d4.z, nz) + b4.z;
            v[3] = fmaf(__uint_as_float(r[j + 3]), d4.w, nz) + b4.w;
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = fmaxf(v[k], kSlope * v[k]);   // lrelu (gain folded downstream)
            if (p.clamp > 0.f) {
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = fminf(fmaxf(v[k], -p.clamp), p.clamp);
            }
            if (p.rgb_w) {
                const float4 w0 = lds4(e.w0 + c0 + j), w1 = lds4(e.w1 + c0 + j), w2 = lds4(e.w2 + c0 + j);
                rgb0 += v[0] * w0.x + v[1] * w0.y + v[2] * w0.z + v[3] * w0.w;
                rgb1 += v[0] * w1.x + v[1] * w1.y + v[2] * w1.z + v[3] * w1.w;
                rgb2 += v[0] * w2.x + v[1] * w2.y + v[2] * w2.z + v[3] * w2.w;
            }
            packed[j / 2 + 0] = pack_bf16(v[0] * s4.x, v[1] * s4.y);
            packed[j / 2 + 1] = pack_bf16(v[2] * s4.z, v[3] * s4.w);
        }
    } else {
#pragma unroll
        for (int j = 0; j < W; j += 4) {
            const float4 d4 = lds4(e.demod + c0 + j);
            packed[j / 2 + 0] = pack_bf16(__uint_as_float(r[j + 0]) * d4.x, __uint_as_float(r[j + 1]) * d4.y);
            packed[j / 2 + 1] = pack_bf16(__uint_as_float(r[j + 2]) * d4.z, __uint_as_float(r[j + 3]) * d4.w);
        }
    }
    if (p.out == nullptr) return;             // last layer: only the ToRGB partial sums leave the kernel
    const uint32_t o = off16 == 0xffffffffu ? off16 : off16 + (uint32_t)(c0 >> 3);
    if constexpr (W == 32) {
        store_rows64_coalesced(stg, packed, o, reinterpret_cast<uint8_t *>(p.out), lane);
    } else {
        if (o != 0xffffffffu) {
            uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(p.out) + ((size_t)o << 4));
#pragma unroll
            for (int v4 = 0; v4 < W / 8; ++v4)
                dst[v4] = make_uint4(packed[4 * v4], packed[4 * v4 + 1], packed[4 * v4 + 2], packed[4 * v4 + 3]);
        }
    }
}

__global__ void __launch_bounds__(kGemmThreads, 1)
modconv_gemm_kernel(const __grid_constant__ GemmParams p, const __grid_constant__ CUtensorMap tmA0,
                    const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                    const __grid_constant__ CUtensorMap tmA3, const __grid_constant__ CUtensorMap tmB) {
    extern __shared__ uint8_t smem_raw[];
    GemmSmem &sm = *reinterpret_cast<GemmSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < kMaxStages; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&sm.tmem_full[i], 1); mbar_init(&sm.tmem_empty[i], p.epi_alt ? kEpiWarps / 2 : kEpiWarps); }
        mbar_init(&sm.b_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&sm.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;
    // K chunk per stage: 64 channels (128-byte rows, SWIZZLE_128B) or, for the 32-channel 1024^2 tail,
    // 32 channels (64-byte rows, SWIZZLE_64B)
    const uint32_t bk = (uint32_t)p.block_k, row_bytes = bk * 2;
    const uint32_t a_stage = kBlockM * row_bytes;
    const uint32_t b_bytes = (uint32_t)p.block_n * row_bytes;
    // kpack K chunks per pipeline stage: narrow N means short MMAs, so two chunks (8 MMAs) per barrier
    // round trip keep the tensor pipe fed while the issuing lane does its per-stage bookkeeping
    const uint32_t kpk = (uint32_t)p.kpack;
    const bool resident = p.resident != 0;
    // accumulator ring in TMEM: short-K tiles (polyphase sub-problems, the narrow tail) need the MMA warp
    // to run several tiles ahead of the epilogue
    const uint32_t nacc = p.block_n <= 128 ? 4u : 2u, acc_cols = 512u / nacc;
    // multiple of 1 KiB: swizzle atoms stay aligned
    const uint32_t stage_bytes = resident ? (uint32_t)p.stage_bytes : kpk * (a_stage + b_bytes);
    const uint32_t ring_off = resident ? (uint32_t)p.resb_bytes : 0u;      // resident weights sit in front of the ring
    const uint32_t nstages = min((uint32_t)kMaxStages, ((uint32_t)kRingBytes - ring_off) / stage_bytes);

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform control flow, one elected lane issues) =====
        if (resident) {
            // weights once: [tap][K chunk] tiles of BLOCK_N rows, K-major swizzled like the streamed ones
            if (elect_one()) {
                mbar_arrive_expect_tx(&sm.b_full, (uint32_t)p.resb_bytes);
                for (int tap = 0; tap < kGemmMaxTaps; ++tap)
                    for (int kc = 0; kc < p.kchunks; ++kc)
                        tma_load_3d(sm.ring + (uint32_t)(tap * p.kchunks + kc) * b_bytes, &tmB, &sm.b_full, kc * (int)bk, 0, tap);
            }
            __syncwarp();
            uint32_t stage = 0, phase = 0;
            for (int seg = 0; seg < p.nseg; ++seg)
            for (int s = 0; s < p.nsub; ++s) {
                const GemmSub &g = p.sub[s];
                const CUtensorMap *tmA = s == 0 ? &tmA0 : (s == 1 ? &tmA1 : (s == 2 ? &tmA2 : &tmA3));
                const uint32_t slab_bytes = (uint32_t)(g.slab_rows * g.TW) * row_bytes;
                const int nslab = g.nslab, sdx[3] = {g.slab_dx[0], g.slab_dx[1], g.slab_dx[2]};
                const TileRange tr = cta_range(p, g, seg);
                TileCoord t = decode_tile(p, g, tr.lo);
                for (int local = tr.lo; local < tr.hi; ++local, next_tile(p, g, t)) {
                    const int ay = t.y0 + g.slab_dy0;
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        mbar_wait(&sm.empty[stage], phase ^ 1);
                        uint8_t *slot = sm.ring + ring_off + stage * stage_bytes;
                        if (SG2_DBG(p) & 2) {                 // bottleneck analysis only (variant build): no activation loads
                            if (elect_one()) mbar_arrive(&sm.full[stage]);
                        } else if (elect_one()) {
                            mbar_arrive_expect_tx(&sm.full[stage], (uint32_t)g.nslab * slab_bytes);
#pragma unroll
                            for (int sl = 0; sl < 3; ++sl)
                                if (sl < nslab)
                                    tma_load_4d(slot + sl * slab_bytes, tmA, &sm.full[stage], kc * (int)bk, t.x0 + sdx[sl], ay, t.b0);
                        }
                        __syncwarp();
                        if (++stage == nstages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        } else {
            uint32_t stage = 0, phase = 0;
            for (int seg = 0; seg < p.nseg; ++seg)
            for (int s = 0; s < p.nsub; ++s) {
                const GemmSub &g = p.sub[s];
                const CUtensorMap *tmA = s == 0 ? &tmA0 : (s == 1 ? &tmA1 : (s == 2 ? &tmA2 : &tmA3));
                const uint32_t a_bytes = (uint32_t)(g.TH * g.TW * g.NB) * row_bytes;
                const TileRange tr = cta_range(p, g, seg);
                TileCoord t = decode_tile(p, g, tr.lo);
                for (int local = tr.lo; local < tr.hi; ++local, next_tile(p, g, t)) {
                    for (int kc = 0; kc < p.kchunks; kc += (int)kpk) {
                        for (int tap = 0; tap < g.ntaps; ++tap) {
                            mbar_wait(&sm.empty[stage], phase ^ 1);
                            uint8_t *slot = sm.ring + stage * stage_bytes;
                            const int ax = t.x0 + g.dx[tap], ay = t.y0 + g.dy[tap], wt = g.wtap[tap];
                            if (p.multi_map) {           // this tap's own activation tensor
                                const int mi = g.tap_map[tap];
                                tmA = mi == 0 ? &tmA0 : (mi == 1 ? &tmA1 : (mi == 2 ? &tmA2 : &tmA3));
                            }
                            if (SG2_DBG(p) & 6) {       // bottleneck analysis only
                                if (elect_one()) {
                                    const bool la = !(SG2_DBG(p) & 2), lb = !(SG2_DBG(p) & 4);
                                    if (!la && !lb) mbar_arrive(&sm.full[stage]);
                                    else mbar_arrive_expect_tx(&sm.full[stage], kpk * ((la ? a_bytes : 0u) + (lb ? b_bytes : 0u)));
                                    for (uint32_t u = 0; u < kpk; ++u) {
                                        if (la) tma_load_4d(slot + u * a_stage, tmA, &sm.full[stage], (kc + (int)u) * (int)bk, ax, ay, t.b0);
                                        if (lb) tma_load_3d(slot + kpk * a_stage + u * b_bytes, &tmB, &sm.full[stage], (kc + (int)u) * (int)bk,
                                                            t.nt * p.block_n, wt);
                                    }
                                }
                            } else if (elect_one()) {
                                mbar_arrive_expect_tx(&sm.full[stage], kpk * (a_bytes + b_bytes));
                                tma_load_4d(slot, tmA, &sm.full[stage], kc * (int)bk, ax, ay, t.b0);
                                tma_load_3d(slot + kpk * a_stage, &tmB, &sm.full[stage], kc * (int)bk, t.nt * p.block_n, wt);
                                if (kpk == 2) {
                                    tma_load_4d(slot + a_stage, tmA, &sm.full[stage], (kc + 1) * (int)bk, ax, ay, t.b0);
                                    tma_load_3d(slot + 2 * a_stage + b_bytes, &tmB, &sm.full[stage], (kc + 1) * (int)bk,
                                                t.nt * p.block_n, wt);
                                }
                            }
                            __syncwarp();
                            if (++stage == nstages) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1 || warp == kMma2Warp) {
        // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =======
        // Resident-weights layers have short tiles (18 MMAs of 16 clk at Cin = Cout = 32): the ISSUING warp, not the
        // tensor pipe, paced them (ncu: ~280 instructions and 1660 clk per tile).  With p.mma2 two warps issue
        // alternate tiles (different accumulators, their own stages); each walks the same tile sequence and only
        // advances the ring / accumulator counters over the other warp's tiles.
        const bool second = warp == kMma2Warp;
        if (second && !(resident && p.mma2)) {
        } else if (resident) {
            const uint32_t idesc = make_idesc_bf16(kBlockM, (uint32_t)p.block_n);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, it = 0;
            const uint32_t my_parity = second ? 1u : 0u;
            const bool split = p.mma2 != 0;
            const uint32_t ring = smem_u32(sm.ring);
            const uint32_t ksteps = bk / 16;
            mbar_wait(&sm.b_full, 0);
            for (int seg = 0; seg < p.nseg; ++seg)
#pragma unroll
            for (int s = 0; s < kGemmMaxSub; ++s) {          // unrolled: p.sub[s] fields become static constant-bank reads
                if (s >= p.nsub) break;
                const GemmSub &g = p.sub[s];
                const TileRange tr = cta_range(p, g, seg);
                // per-tap descriptor offsets are read from the (constant-bank) parameter block right where they are
                // used: they stay in UNIFORM registers.  Hoisting them into a local array put them in vector registers
                // and cost an R2UR + IADD chain per MMA -- on the narrow 1024^2 layers (18 MMAs of 16 clk per tile)
                // the issuing warp, not the tensor pipe, paced the kernel (ncu: ~280 instructions per tile).
                const uint32_t wb16 = ((uint32_t)p.kchunks * b_bytes) >> 4;   // 16-byte units between the weights of two taps
                const int ntaps = g.ntaps;
                for (int local = tr.lo; local < tr.hi; ++local, ++it) {
                    if (split && (it & 1u) != my_parity) {      // the other issuing warp's tile
                        for (int kc = 0; kc < p.kchunks; ++kc)
                            if (++stage == nstages) { stage = 0; phase ^= 1; }
                        if (++acc == nacc) { acc = 0; acc_phase ^= 1; }
                        continue;
                    }
                    mbar_wait(&sm.tmem_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * acc_cols;
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        mbar_wait(&sm.full[stage], phase);
                        tc_fence_after();
                        const uint64_t a0 = make_smem_desc(ring + ring_off + stage * stage_bytes, row_bytes);
                        const uint64_t b0 = make_smem_desc(ring + (uint32_t)kc * b_bytes, row_bytes);
                        if (elect_one()) {
#pragma unroll
                            for (int tap = 0; tap < kGemmMaxTaps; ++tap) {
                                if (tap < ntaps) {
                                    const uint64_t adesc = a0 + ((uint32_t)g.tap_aoff[tap] >> 4);
                                    const uint64_t bdesc = b0 + (uint32_t)g.wtap[tap] * wb16;
                                    // 32 bytes (>>4 = 2) inside the swizzle row per K = 16 step
                                    umma_bf16(d_tmem, adesc, bdesc, idesc, (kc | tap) != 0);
                                    if (SG2_DBG(p) & 8) continue;          // knock-out: one MMA per tap
                                    umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1);
                                    if (ksteps == 4) {
                                        umma_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1);
                                        umma_bf16(d_tmem, adesc + 6, bdesc + 6, idesc, 1);
                                    }
                                }
                            }
                            umma_commit(&sm.empty[stage]);
                        }
                        __syncwarp();
                        if (++stage == nstages) { stage = 0; phase ^= 1; }
                    }
                    if (elect_one()) umma_commit(&sm.tmem_full[acc]);
                    __syncwarp();
                    if (++acc == nacc) { acc = 0; acc_phase ^= 1; }
                }
            }
        } else {
            const uint32_t idesc = make_idesc_bf16(kBlockM, (uint32_t)p.block_n);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (int seg = 0; seg < p.nseg; ++seg)
            for (int s = 0; s < p.nsub; ++s) {
                const GemmSub &g = p.sub[s];
                const int nstage = p.kchunks / (int)kpk * g.ntaps;
                const TileRange tr = cta_range(p, g, seg);
                for (int local = tr.lo; local < tr.hi; ++local) {
                    mbar_wait(&sm.tmem_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * acc_cols;
                    for (int k0 = 0; k0 < nstage; ++k0) {
                        mbar_wait(&sm.full[stage], phase);
                        tc_fence_after();
                        const uint32_t slot = smem_u32(sm.ring) + stage * stage_bytes;
                        const uint64_t adesc = make_smem_desc(slot, row_bytes);
                        const uint64_t bdesc = make_smem_desc(slot + kpk * a_stage, row_bytes);
                        if (elect_one()) {
                            // advance 32 bytes (>>4 = 2) inside the swizzle row per K = 16 step
                            umma_bf16(d_tmem, adesc, bdesc, idesc, k0 != 0);
                            if (!(SG2_DBG(p) & 8)) {
                            umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1);
                            if (bk == 64) {
                                umma_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1);
                                umma_bf16(d_tmem, adesc + 6, bdesc + 6, idesc, 1);
                            }
                            if (kpk == 2) {      // second K chunk of the stage (bk == 64 by construction)
                                const uint64_t a2 = adesc + (a_stage >> 4), b2 = bdesc + (b_bytes >> 4);
                                umma_bf16(d_tmem, a2, b2, idesc, 1);
                                umma_bf16(d_tmem, a2 + 2, b2 + 2, idesc, 1);
                                umma_bf16(d_tmem, a2 + 4, b2 + 4, idesc, 1);
                                umma_bf16(d_tmem, a2 + 6, b2 + 6, idesc, 1);
                            }
                            }
                            umma_commit(&sm.empty[stage]);       // frees the smem slot when these MMAs retire
                        }
                        __syncwarp();
                        if (++stage == nstages) { stage = 0; phase ^= 1; }
                    }
                    if (elect_one()) umma_commit(&sm.tmem_full[acc]);   // accumulator complete -> epilogue
                    __syncwarp();
                    if (++acc == nacc) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else if (warp < kMma2Warp) {
        // ===================== epilogue: 8 warps, two per TMEM lane quarter =====================
        // warp (2 + 4h + i) may access TMEM lanes 32*((2+i)&3) ..  Two schedules:
        //  * column split (default): the two warps of a quarter alternate 32-column chunks of every tile;
        //  * tile split (p.epi_alt, narrow BLOCK_N): warps 2-5 take the even tiles and warps 6-9 the odd ones, so
        //    the per-pixel work (coordinates, noise, addresses, ToRGB store) is done once per pixel.
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2;
        const bool alt = p.epi_alt != 0;
        const int half = alt ? 0 : grp;
        const int m = q * 32 + lane;                 // accumulator row = pixel of the tile
        const uint32_t stg = smem_u32(sm.stg[warp - 2]);   // this warp's store staging buffer
        const int et = alt ? ((threadIdx.x - 64) & 127) : threadIdx.x - 64;   // thread index inside the staging group
        const int ethreads = alt ? 128 : 256, bar_id = alt ? 1 + grp : 1, step = alt ? 2 : 1;
        // staged per-(sample, channel) parameters: one copy per group in tile-split mode
        const int po = alt ? grp * (kEpiCap / 2) : 0;
        float *e_demod = sm.e_demod + po, *e_next = sm.e_next + po, *e_bias = sm.e_bias + (alt ? grp * (kMaxBlockN / 2) : 0);
        float *e_w0 = sm.e_wrgb[0] + po, *e_w1 = sm.e_wrgb[1] + po, *e_w2 = sm.e_wrgb[2] + po;
        const float nw = (p.mode == 0 && p.noise) ? __ldg(p.noise_weight) : 0.f;
        const bool has_noise = p.mode == 0 && p.noise != nullptr;
        const int N = p.block_n;
        uint32_t it0 = 0;                            // tiles of the earlier sub-problems (= the MMA warp's count)
        int staged_key = -1;
        for (int seg = 0; seg < p.nseg; ++seg)
        for (int s = 0; s < p.nsub; ++s) {
            const GemmSub &g = p.sub[s];
            const int per = g.TH * g.TW;
            const int nb = m / per;
            const int rem = m - nb * per;
            const int ty = rem / g.TW, tx = rem - ty * g.TW;
            const int pb = (nb < g.NB ? nb : 0) * N;    // row of the staged per-sample params
            const int PH = g.PH, PW = g.PW, NB = g.NB;
            const long long plane = (long long)PH * PW;
            const TileRange tr = cta_range(p, g, seg);
            int local = tr.lo + (alt ? (int)((grp - it0) & 1u) : 0);
            TileCoord t = decode_tile(p, g, local);
            auto noise_at = [&](const TileCoord &tc) -> float {
                const int y = tc.y0 + ty, x = tc.x0 + tx, b = tc.b0 + nb;
                if (has_noise && nb < NB && y < PH && x < PW && b < p.B)
                    return __ldg(p.noise + (long long)b * p.noise_bstride + (long long)y * PW + x);
                return 0.f;
            };
            float nz_cur = local < tr.hi ? noise_at(t) : 0.f;
            for (; local < tr.hi; local += step) {
                // the noise of this group's NEXT tile is fetched now and consumed one iteration later
                TileCoord tn = t;
                next_tile(p, g, tn);
                if (alt) next_tile(p, g, tn);
                const float nz_nxt = local + step < tr.hi ? noise_at(tn) : 0.f;
                const uint32_t it = it0 + (uint32_t)(local - tr.lo);
                const uint32_t acc = it & (nacc - 1), acc_phase = (it / nacc) & 1u;
                const int n0 = t.nt * N;
                const int y = t.y0 + ty, x = t.x0 + tx, b = t.b0 + nb;
                const bool valid = nb < NB && y < PH && x < PW && b < p.B;
                // ---- stage the per-(sample, channel) epilogue parameters when they change ----
                const int key = (t.b0 * kMaxBlockN + t.nt) * 16 + NB;
                if (key != staged_key) {
                    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(ethreads) : "memory");
                    for (int i = et; i < NB * N; i += ethreads) {
                        const int sb = i / N, col = i - sb * N;
                        const int bb = t.b0 + sb < p.B ? t.b0 + sb : p.B - 1;
                        const long long o = (long long)bb * p.Cout + n0 + col;
                        e_demod[i] = __ldg(p.demod + o);
                        if (p.mode == 0) {
                            // lrelu gain sqrt(2) (fused_bias_act_kernel.cu:47) folded into both consumers
                            e_next[i] = p.next_style ? 1.41421356237f * __ldg(p.next_style + o) : 0.f;
                            if (p.rgb_w) {
                                const float rs = 1.41421356237f * __ldg(p.rgb_style + o);
                                e_w0[i] = rs * __ldg(p.rgb_w + n0 + col);
                                e_w1[i] = rs * __ldg(p.rgb_w + (long long)p.Cout + n0 + col);
                                e_w2[i] = rs * __ldg(p.rgb_w + 2LL * p.Cout + n0 + col);
                            }
                        }
                    }
                    if (p.mode == 0)
                        for (int i = et; i < N; i += ethreads) e_bias[i] = __ldg(p.bias + n0 + i);
                    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(ethreads) : "memory");
                    staged_key = key;
                }
                const float nz = nz_cur * nw;
                uint32_t off16 = 0xffffffffu;
                if (valid && p.out && !(SG2_DBG(p) & 1))
                    off16 = (uint32_t)((g.out_off + (((long long)b * g.out_H + y) * g.out_W + x) * p.Cout + n0) >> 3);

                mbar_wait(&sm.tmem_full[acc], acc_phase);
                tc_fence_after();
                float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * acc_cols;
                const EpiTables tab{e_demod + pb, e_next + pb, e_w0 + pb, e_w1 + pb, e_w2 + pb, e_bias};
                if (SG2_DBG(p) & 16) {   // knock-out: no epilogue work at all
                } else if (alt) {            // this warp owns every column of the tile
                    if (N == 16) epi_cols<16>(p, tab, t_row, 0, nz, off16, stg, lane, rgb0, rgb1, rgb2);
                    else
                        for (int c0 = 0; c0 < N; c0 += 32) epi_cols<32>(p, tab, t_row + c0, c0, nz, off16, stg, lane, rgb0, rgb1, rgb2);
                } else if (N >= 64) {   // the two warps of a lane quarter alternate 32-column chunks
                    for (int c0 = 32 * half; c0 < N; c0 += 64) epi_cols<32>(p, tab, t_row + c0, c0, nz, off16, stg, lane, rgb0, rgb1, rgb2);
                } else if (16 * half < N) {   // N = 32 / 16 with NB * N > 256: 16 columns per warp
                    epi_cols<16>(p, tab, t_row + 16 * half, 16 * half, nz, off16, stg, lane, rgb0, rgb1, rgb2);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.tmem_empty[acc]);
                if (valid && p.mode == 0 && p.rgb_w) {   // each column half writes its own partial plane
                    float *rp = p.rgb_part + ((((long long)t.nt * (alt ? 1 : 2) + half) * p.B + b) * 3) * plane + (long long)y * PW + x;
                    rp[0] = rgb0;
                    rp[plane] = rgb1;
                    rp[2 * plane] = rgb2;
                }
                t = tn;
                nz_cur = nz_nxt;
            }
            it0 += (uint32_t)(tr.hi - tr.lo);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int launch_modconv_gemm(const GemmParams &p, const CUtensorMap *tmA, const CUtensorMap &tmB, int sms,
                        cudaStream_t st) {
    static_assert(sizeof(GemmSmem) + 1024 <= 227 * 1024, "GemmSmem exceeds the 227 KiB CTA limit");
    static_assert(sizeof(GemmParams) + 5 * sizeof(CUtensorMap) <= 4000, "kernel parameter space");
    const size_t smem = sizeof(GemmSmem) + 1024;
    static std::atomic<int> configured{0};
    if (!configured.load(std::memory_order_acquire)) {
        SG2_CUDA_OK(cudaFuncSetAttribute(modconv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.store(1, std::memory_order_release);
    }
    SG2_REQUIRE(p.block_n % 16 == 0 && p.block_n >= 16 && p.block_n <= kMaxBlockN, SG2_ERR_BAD_ARG, "gemm: bad BLOCK_N %d", p.block_n);
    SG2_REQUIRE(!p.up4 && !p.resident2, SG2_ERR_UNSUPPORTED, "gemm: the fused up-sampling conv runs on the cta_group::2 kernel only");
    if (p.resident) {
        SG2_REQUIRE(p.n_tiles_n == 1 && p.kpack == 1 && p.resb_bytes == 9 * p.Cin * p.block_n * 2 && p.stage_bytes % 1024 == 0 &&
                        p.resb_bytes % 1024 == 0 && p.resb_bytes + 2 * p.stage_bytes <= kRingBytes,
                    SG2_ERR_BAD_ARG, "gemm: bad resident-weights plan (weights %d B, stage %d B)", p.resb_bytes, p.stage_bytes);
        for (int s = 0; s < p.nsub; ++s)
            SG2_REQUIRE(p.sub[s].NB == 1 && p.sub[s].TW % 8 == 0 && p.sub[s].nslab >= 1 && p.sub[s].nslab <= 3 &&
                            p.sub[s].nslab * p.sub[s].slab_rows * p.sub[s].TW * p.block_k * 2 <= p.stage_bytes,
                        SG2_ERR_BAD_ARG, "gemm: bad slab plan of sub-problem %d", s);
    }
    SG2_REQUIRE(p.kpack == 1 || (p.kpack == 2 && p.block_k == 64 && p.kchunks % 2 == 0), SG2_ERR_BAD_ARG, "gemm: bad kpack %d", p.kpack);
    SG2_REQUIRE((p.block_k == 64 || p.block_k == 32) && p.Cin % p.block_k == 0 && p.Cout % p.block_n == 0, SG2_ERR_UNSUPPORTED,
                "gemm: Cin %d must be a multiple of BLOCK_K %d and Cout %d of BLOCK_N %d", p.Cin, p.block_k, p.Cout, p.block_n);
    for (int s = 0; s < p.nsub; ++s)
        SG2_REQUIRE(p.sub[s].NB * p.block_n <= kEpiCap && p.sub[s].TH * p.sub[s].TW * p.sub[s].NB <= kBlockM,
                    SG2_ERR_BAD_ARG, "gemm: tile of sub-problem %d too large", s);
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    if (grid <= 0) return SG2_OK;
    GemmParams q = p;                      // segments of the polyphase interleave (GemmParams::seg_tiles)
    q.nseg = 1;
    if (q.seg_tiles > 0) {
        int max_per = 1;
        for (int s = 0; s < q.nsub; ++s) {
            const GemmSub &g = q.sub[s];
            const int count = g.tiles_x * g.tiles_y * g.tiles_b * q.n_tiles_n;
            max_per = std::max(max_per, (count + grid - 1) / grid);
        }
        q.nseg = (max_per + q.seg_tiles - 1) / q.seg_tiles;
    }
    SG2_REQUIRE(!p.multi_map || (p.nsub == 1 && !p.resident), SG2_ERR_BAD_ARG, "gemm: a multi-tensor tap list needs one streamed sub-problem");
    const int nmaps = p.multi_map ? 4 : p.nsub;
    const CUtensorMap &a0 = tmA[0];
    const CUtensorMap &a1 = tmA[nmaps > 1 ? 1 : 0];
    const CUtensorMap &a2 = tmA[nmaps > 2 ? 2 : 0];
    const CUtensorMap &a3 = tmA[nmaps > 3 ? 3 : 0];
    modconv_gemm_kernel<<<grid, kGemmThreads, smem, st>>>(q, a0, a1, a2, a3, tmB);
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

}  // namespace sg2
