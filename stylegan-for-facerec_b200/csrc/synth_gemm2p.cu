// synth_gemm2p.cu -- the transposed (up-sampling) convolution of ModulatedConv2d (model.py:246-252 of the reference) as
// a MERGED POLYPHASE walk on cta_group::2 tcgen05.
//
// conv_transpose2d(x, w, stride 2) = four polyphase planes T[2y+py, 2x+px], plane (py, px) being a convolution of the
// INPUT with 4 / 2 / 2 / 1 of the nine 3x3 taps, all of them reading x at offsets dy, dx in {0, -1}.  synth_gemm.cu /
// synth_gemm2.cu run the planes as four sub-problems one after the other: every plane loads its own activation tile per
// tap, so the input is fetched nine times (ncu, round 1: DRAM traffic 2.0x algorithmic, and with 32 KiB of operand fill
// per 4 MMAs the shared-memory port -- fill writes + UMMA operand reads -- holds the 256 -> 128 layer at 50 % tensor pipe).
// Here ONE tile walk serves all four planes:
//   * per K chunk the activation window of a 16 x 8 pixel tile is loaded ONCE as two slabs (dx = 0 and dx = -1) of
//     17 rows (the dy = -1 halo inside the slab); a tap's A operand is a slab + (dy + 1) * 1 KiB -- a descriptor offset;
//   * the nine weight tiles of the K chunk stream through their own ring, each CTA of the pair holding half of the rows
//     (cta_group::2: the leader issues M = 256 MMAs, each SM reads A (4 KiB) + B/2 per MMA instead of A + B);
//   * four accumulators of BLOCK_N <= 128 columns live in TMEM at once, one per plane; within a K chunk the taps are
//     issued plane by plane (1, 2, 2, 4 taps), so plane p's accumulator completes -- and its epilogue starts -- while
//     the remaining planes' MMAs of the last K chunk still run, and plane p of the NEXT tile waits only for that epilogue.
// Operand fill per tile and K chunk: 34 KiB of activations + 9 weight half-tiles instead of 9 x (16 KiB + half-tile).
// Epilogue: plain scaled store (x demodulation, bf16) of each plane into the (r+1)^2 polyphase buffer the FIR kernel reads.
#include "common.cuh"
#include "synth_gemm.cuh"
#include "tc_ptx.cuh"

namespace sg2 {

using namespace tc;

constexpr int kP4EpiWarps = 16;                         // four per TMEM lane quarter (see synth_gemm2.cu)
constexpr int kP4Threads = 64 + 32 * kP4EpiWarps;      // TMA warp + MMA warp + epilogue warps
constexpr int kP4TH = 16, kP4TW = 8;                    // tile: 16 rows x 8 pixels (a dy shift = 8 rows = one swizzle atom)
constexpr int kP4SlabBytes = (kP4TH + 1) * kP4TW * 128;  // 17 KiB: one dx slab of a 64-channel K chunk
constexpr int kP4AStage = 2 * kP4SlabBytes;             // 34 KiB
constexpr int kP4AStages = 2;
constexpr int kP4BStages = 14;
constexpr int kP4BBytesMax = 64 * 128;                  // half of a 128-row weight tile
constexpr int kP4Cap = 128;                             // BLOCK_N entries of per-sample epilogue params

struct __align__(1024) Poly4Smem {
    uint8_t a_ring[kP4AStages * kP4AStage];
    uint8_t b_ring[kP4BStages * kP4BBytesMax];
    uint8_t stg[kP4EpiWarps][2048];          // per-warp transposition buffers of the coalesced epilogue store
    float e_demod[kP4Cap];
    uint64_t a_full[kP4AStages], a_empty[kP4AStages];
    uint64_t b_full[kP4BStages], b_empty[kP4BStages];
    uint64_t tmem_full[4], tmem_empty[4];    // one accumulator per polyphase plane
    uint32_t tmem_base;
};

struct P4Tile { int nt, x0, y0, b, dummy; };

__device__ __forceinline__ int p4_pair_groups(const GemmSub &g) { return (g.tiles_x * g.tiles_y * g.tiles_b + 1) / 2; }
// the walk uses plane (0, 0)'s tile grid (the largest: (r+1) x (r+1)); the N tile is the slowest index
__device__ __forceinline__ P4Tile p4_decode(const GemmParams &p, int item, int rank) {
    const GemmSub &g = p.sub[0];
    P4Tile t;
    const int pg = p4_pair_groups(g), sxyb = g.tiles_x * g.tiles_y * g.tiles_b;
    const int q = item % pg;
    t.nt = item / pg;
    int sp = 2 * q + rank;
    t.dummy = sp >= sxyb;             // odd tile count: the surplus CTA recomputes the last tile, stores nothing
    if (t.dummy) sp = sxyb - 1;
    const int bx = sp % g.tiles_x;
    sp /= g.tiles_x;
    const int by = sp % g.tiles_y;
    t.b = sp / g.tiles_y;
    t.x0 = bx * kP4TW;
    t.y0 = by * kP4TH;
    return t;
}
struct P4Range { int lo, hi; };
__device__ __forceinline__ P4Range p4_range(const GemmParams &p) {
    const int count = p4_pair_groups(p.sub[0]) * p.n_tiles_n;          // work items of a CLUSTER
    const int ncl = (int)gridDim.x / 2, cid = (int)blockIdx.x / 2;
    const int per = (count + ncl - 1) / ncl;
    P4Range r;
    r.lo = min(count, cid * per);
    r.hi = min(count, r.lo + per);
    return r;
}

__device__ __forceinline__ uint32_t p4_pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}

__global__ void __launch_bounds__(kP4Threads, 1)
modconv_gemm2_poly4_kernel(const __grid_constant__ GemmParams p, const __grid_constant__ CUtensorMap tmA,
                           const __grid_constant__ CUtensorMap tmB) {
    extern __shared__ uint8_t smem_raw[];
    Poly4Smem &sm = *reinterpret_cast<Poly4Smem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < kP4AStages; ++i) { mbar_init(&sm.a_full[i], 1); mbar_init(&sm.a_empty[i], 1); }
        for (int i = 0; i < kP4BStages; ++i) { mbar_init(&sm.b_full[i], 1); mbar_init(&sm.b_empty[i], 1); }
        // tmem_empty lives in the leader: the epilogue warps of BOTH CTAs arrive on it
        for (int i = 0; i < 4; ++i) { mbar_init(&sm.tmem_full[i], 1); mbar_init(&sm.tmem_empty[i], 2 * kP4EpiWarps); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(&sm.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync();                              // peer barriers initialised, peer TMEM allocated
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;
    const uint32_t b_bytes = (uint32_t)(p.block_n / 2) * 128u;     // this CTA's half of a weight tile (64-channel K chunk)
    const P4Range range = p4_range(p);

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform control flow, one elected lane issues) =====
        uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
        for (int item = range.lo; item < range.hi; ++item) {
            const P4Tile t = p4_decode(p, item, rank);
            const int wrow = t.nt * p.block_n + rank * (p.block_n / 2);
            for (int kc = 0; kc < p.kchunks; ++kc) {
                mbar_wait(&sm.a_empty[sa], pa ^ 1);
                if (elect_one()) {
                    // the leader's barrier collects the bytes of both CTAs
                    if (leader) mbar_arrive_expect_tx(&sm.a_full[sa], 2u * kP4AStage);
                    const uint32_t dst = smem_u32(sm.a_ring) + sa * kP4AStage;
                    tma_load_4d_2sm(dst, &tmA, &sm.a_full[sa], kc * 64, t.x0, t.y0 - 1, t.b);                    // dx =  0
                    tma_load_4d_2sm(dst + kP4SlabBytes, &tmA, &sm.a_full[sa], kc * 64, t.x0 - 1, t.y0 - 1, t.b);  // dx = -1
                }
                __syncwarp();
                if (++sa == kP4AStages) { sa = 0; pa ^= 1; }
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    mbar_wait(&sm.b_empty[sb], pb ^ 1);
                    if (elect_one()) {
                        if (leader) mbar_arrive_expect_tx(&sm.b_full[sb], 2u * b_bytes);
                        tma_load_3d_2sm(smem_u32(sm.b_ring) + sb * b_bytes, &tmB, &sm.b_full[sb], kc * 64, wrow, p.m_wtap[j]);
                    }
                    __syncwarp();
                    if (++sb == kP4BStages) { sb = 0; pb ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: the leader CTA's warp, one elected lane issues ============
        if (leader) {
            const uint32_t idesc = make_idesc_bf16(2 * kBlockM, (uint32_t)p.block_n);
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
            for (int item = range.lo; item < range.hi; ++item, ++it) {
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    mbar_wait(&sm.a_full[sa], pa);
                    tc_fence_after();
                    const uint32_t a_base = smem_u32(sm.a_ring) + sa * kP4AStage;
                    const bool first_kc = kc == 0, last_kc = kc == p.kchunks - 1;
#pragma unroll
                    for (int j = 0; j < 9; ++j) {
                        const uint32_t ph = (uint32_t)p.m_ph[j];
                        if (first_kc && p.m_first[j]) {          // plane ph of the previous tile has been drained
                            mbar_wait(&sm.tmem_empty[ph], (it & 1u) ^ 1u);
                            tc_fence_after();
                        }
                        mbar_wait(&sm.b_full[sb], pb);
                        tc_fence_after();
                        const uint64_t adesc = make_smem_desc(a_base + (uint32_t)p.m_aoff[j], 128);
                        const uint64_t bdesc = make_smem_desc(smem_u32(sm.b_ring) + sb * b_bytes, 128);
                        const uint32_t d_tmem = tmem_base + ph * 128u;
                        if (elect_one()) {
                            // advance 32 bytes (>>4 = 2) inside the swizzle row per K = 16 step
                            umma_bf16_2sm(d_tmem, adesc, bdesc, idesc, (first_kc && p.m_first[j]) ? 0u : 1u);
                            umma_bf16_2sm(d_tmem, adesc + 2, bdesc + 2, idesc, 1);
                            umma_bf16_2sm(d_tmem, adesc + 4, bdesc + 4, idesc, 1);
                            umma_bf16_2sm(d_tmem, adesc + 6, bdesc + 6, idesc, 1);
                            umma_commit_2sm_mc(&sm.b_empty[sb], 3);                  // frees the weight slot in BOTH CTAs
                            if (last_kc && p.m_last[j]) umma_commit_2sm_mc(&sm.tmem_full[ph], 3);   // plane complete -> epilogues
                        }
                        __syncwarp();
                        if (++sb == kP4BStages) { sb = 0; pb ^= 1; }
                    }
                    if (elect_one()) umma_commit_2sm_mc(&sm.a_empty[sa], 3);         // frees the activation slabs in BOTH CTAs
                    __syncwarp();
                    if (++sa == kP4AStages) { sa = 0; pa ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue: 16 warps, four per TMEM lane quarter =====================
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;            // column group 0..3
        const int m = q * 32 + lane;                 // accumulator row = pixel of the tile
        const int ty = m / kP4TW, tx = m - ty * kP4TW;
        const int et = threadIdx.x - 64;             // 0..511 within the epilogue group
        const int N = p.block_n;
        uint32_t it = 0;
        int staged_key = -1;
        for (int item = range.lo; item < range.hi; ++item, ++it) {
            const P4Tile t = p4_decode(p, item, rank);
            const int n0 = t.nt * N;
            const int y = t.y0 + ty, x = t.x0 + tx;
            const int key = t.b * 16 + t.nt;
            if (key != staged_key) {                  // demodulation row of this (sample, N tile)
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kP4EpiWarps) : "memory");
                if (et < N) sm.e_demod[et] = __ldg(p.demod + (long long)t.b * p.Cout + n0 + et);
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kP4EpiWarps) : "memory");
                staged_key = key;
            }
#pragma unroll
            for (int ph = 3; ph >= 0; --ph) {
                const GemmSub &g = p.sub[ph];
                const bool valid = !t.dummy && y < g.PH && x < g.PW && !(SG2_DBG(p) & 1);
                uint32_t off16 = 0xffffffffu;     // this pixel's row in 16-byte units from p.out (0xFFFFFFFF: not stored)
                if (valid)
                    off16 = (uint32_t)((g.out_off + (((long long)t.b * g.out_H + y) * g.out_W + x) * p.Cout + n0) >> 3);
                mbar_wait(&sm.tmem_full[ph], it & 1u);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)ph * 128u;
                bool handed_back = false;
                for (int c0 = 32 * half; c0 < N; c0 += 128) {
                    if (SG2_DBG(p) & 16) break;
                    uint32_t r[32];
                    tmem_ld32(t_row + c0, r);
                    tmem_ld_wait();
                    if (c0 + 128 >= N) {                    // last TMEM read of this plane: hand the accumulator back before the stores
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_leader(&sm.tmem_empty[ph]);   // the leader's MMA thread waits for both CTAs
                        handed_back = true;
                    }
                    uint32_t packed[16];
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 d4 = lds128f(smem_u32(&sm.e_demod[c0 + j]));
                        packed[j / 2 + 0] = p4_pack(__uint_as_float(r[j + 0]) * d4.x, __uint_as_float(r[j + 1]) * d4.y);
                        packed[j / 2 + 1] = p4_pack(__uint_as_float(r[j + 2]) * d4.z, __uint_as_float(r[j + 3]) * d4.w);
                    }
                    store_rows64_coalesced(smem_u32(sm.stg[warp - 2]), packed, off16 == 0xffffffffu ? off16 : off16 + (uint32_t)(c0 >> 3),
                                           reinterpret_cast<uint8_t *>(p.out), lane);
                }
                if (!handed_back) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&sm.tmem_empty[ph]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync();                              // neither CTA exits (or frees TMEM) while the pair is still working
    if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
}

int launch_modconv_gemm2_poly4(const GemmParams &p, const CUtensorMap &tmA, const CUtensorMap &tmB, int sms, cudaStream_t st) {
    static_assert(sizeof(Poly4Smem) + 1024 <= 227 * 1024, "Poly4Smem exceeds the 227 KiB CTA limit");
    const size_t smem = sizeof(Poly4Smem) + 1024;
    static std::atomic<int> configured{0};
    if (!configured.load(std::memory_order_acquire)) {
        SG2_CUDA_OK(cudaFuncSetAttribute(modconv_gemm2_poly4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.store(1, std::memory_order_release);
    }
    SG2_REQUIRE(p.poly4 && p.nsub == 4 && p.mode == 1 && p.block_k == 64 && p.Cin % 64 == 0 &&
                    (p.block_n == 64 || p.block_n == 128) && p.Cout % p.block_n == 0 && p.out && p.demod,
                SG2_ERR_BAD_ARG, "gemm2p: bad merged-polyphase plan (Cin %d, Cout %d, BLOCK_N %d)", p.Cin, p.Cout, p.block_n);
    for (int s = 0; s < 4; ++s)
        SG2_REQUIRE(p.sub[s].TH == kP4TH && p.sub[s].TW == kP4TW && p.sub[s].NB == 1, SG2_ERR_BAD_ARG,
                    "gemm2p: plane %d must use 16 x 8 tiles of one sample", s);
    const GemmSub &g = p.sub[0];
    const long items = ((long)g.tiles_x * g.tiles_y * g.tiles_b + 1) / 2 * p.n_tiles_n;
    const int n_clusters = (int)std::min<long>(items, sms / 2);
    if (n_clusters <= 0) return SG2_OK;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * n_clusters);
    cfg.blockDim = dim3(kP4Threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SG2_CUDA_OK(cudaLaunchKernelEx(&cfg, modconv_gemm2_poly4_kernel, p, tmA, tmB));
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

}  // namespace sg2
