// synth_fir.cu -- the blur of the up-sampling layers on the tensor core.
//
// Blur(pad (1,1), 4x4 taps) over the four polyphase planes T of the transposed conv, fused with
// noise + bias + leaky-relu + next-layer modulation (model.py:254-257, 282-287, 335 of the
// reference, three full passes there).  A SIMT stencil needs ~16 FMA + conversions per output
// element, which at bf16 storage is ~1.1x the instruction budget of an HBM-bound kernel on B200
// (measured: 26 % of HBM peak, profiles/kernels_r01_v1.json).  So the stencil is recast as a small
// dense product per tile and given to tcgen05:
//
//     out[128 pixels, N ch] = K[128, 256] . Twin[256, N ch]
//
//   * Twin = the T window that the 16x8 output tile needs: 4 planes x (10 x 6) pixels, each plane
//     padded to 64 rows (pad rows are zeroed once and multiply zero columns of K).  It is loaded by
//     TMA straight from the planes (out-of-range rows/cols zero-filled) and used AS IT LANDS as the
//     MN-major B operand (pixels = K dimension, 64 channels = one 128-byte swizzle row).
//   * K = the banded Toeplitz matrix of the 4x4 taps (values {1,3,9}/16: exact in bf16), built on
//     the host per plan, resident in shared memory for the whole kernel (A operand, K-major).
//   * fp32 accumulation in TMEM; the epilogue warps only do +noise, +bias, lrelu, x style, pack.
// The multiplications by structural zeros cost 16x more MACs than the stencil (still < 30 % of
// the tile's HBM time on the tensor pipe) and buy an instruction stream 4x shorter.
#include "common.cuh"
#include "synth_kernels.cuh"
#include "tc_ptx.cuh"

namespace sg2 {

using namespace tc;

constexpr int FT_TH = 16, FT_TW = 8;             // output tile (pixels)
constexpr int FT_RA = 10, FT_RB = 6;             // window rows / cols per plane
constexpr int FT_PLANE_ROWS = 64;                // 60 used + 4 zero pad rows
constexpr int FT_K = 4 * FT_PLANE_ROWS;          // 256
constexpr int FT_N = 128;                        // channels per tile (two 64-channel column blocks)
constexpr int FT_CB_BYTES = FT_K * 128;          // one column block of the window: 32 KiB
constexpr int FT_RING_CBS = 4;                   // window ring: 4 column blocks = 2 stages at N=128, 4 stages at N=64
constexpr int FT_OUT_CB_BYTES = 128 * 128;       // staged output tile, one column block: 128 pixels x 128 B
constexpr int FT_A_BYTES = 128 * FT_K * 2;       // 64 KiB Toeplitz
constexpr int FT_EPI_WARPS = 8;
constexpr int FT_THREADS = 64 + 32 * FT_EPI_WARPS;

struct __align__(1024) FirSmem {
    uint8_t a[FT_A_BYTES];
    uint8_t b[FT_RING_CBS * FT_CB_BYTES];
    uint8_t o[2 * FT_OUT_CB_BYTES];                // bf16 output tile in the TMA store layout (SWIZZLE_128B)
    float e_bias[FT_N];
    float e_next[FT_N];
    uint64_t a_full;
    uint64_t full[FT_RING_CBS], empty[FT_RING_CBS];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

struct FirTile { int n, ct, y0, x0; };

__device__ __forceinline__ FirTile fir_decode(const UpfirTcParams &p, int tile) {
    FirTile t;
    const int bx = tile % p.tiles_x;
    tile /= p.tiles_x;
    const int by = tile % p.tiles_y;
    tile /= p.tiles_y;
    t.ct = tile % p.tiles_c;
    t.n = tile / p.tiles_c;
    t.x0 = bx * FT_TW;
    t.y0 = by * FT_TH;
    return t;
}

// MN-major SWIZZLE_128B operand: 64-element (128 B) rows along N, 8-row groups along K 1024 B apart
// (SBO), 64-element N blocks `lbo_bytes` apart (LBO).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(FT_THREADS, 1)
upfir_tc_kernel(const __grid_constant__ UpfirTcParams p, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmT0, const __grid_constant__ CUtensorMap tmT1,
                const __grid_constant__ CUtensorMap tmT2, const __grid_constant__ CUtensorMap tmT3,
                const __grid_constant__ CUtensorMap tmO) {
    extern __shared__ uint8_t smem_raw[];
    FirSmem &sm = *reinterpret_cast<FirSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // zero the pad rows (60..63 of every plane region) once: TMA never writes them
    for (int i = threadIdx.x; i < FT_RING_CBS * 4 * 4 * 8; i += FT_THREADS) {
        const int v = i & 7, row = (i >> 3) & 3, plane = (i >> 5) & 3, cb = i >> 7;
        *reinterpret_cast<uint4 *>(&sm.b[cb * FT_CB_BYTES + (plane * FT_PLANE_ROWS + 60 + row) * 128 + v * 16]) =
            make_uint4(0, 0, 0, 0);
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmT0);
        mbar_init(&sm.a_full, 1);
        for (int i = 0; i < FT_RING_CBS; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&sm.tmem_full[i], 1); mbar_init(&sm.tmem_empty[i], FT_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&sm.tmem_base, 256);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // pad-row zeros visible to the MMA (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    const int per = (p.total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    // round-robin: at any moment the 148 CTAs work on ~5 adjacent tile rows, so the halo rows a tile
    // shares with its vertical neighbours are still in L2 when the neighbour loads them
    const int tile_lo = (int)blockIdx.x, tile_step = (int)gridDim.x;
    (void)per;
    const int tile_hi = p.total_tiles;
    const int ncb = p.block_n / 64;                 // column blocks per tile (1 or 2)
    const uint32_t nstages = FT_RING_CBS / ncb, stage_bytes = ncb * FT_CB_BYTES;

    if (warp == 0) {
        // ===================== TMA producer =====================
        // warp-uniform loop, one elected lane issues (no divergent single-thread region)
        if (tile_lo < tile_hi) {
            // Toeplitz matrix: 4 K-atoms of [128 rows][64 k] each
            if (elect_one()) {
                mbar_arrive_expect_tx(&sm.a_full, FT_A_BYTES);
                for (int ka = 0; ka < 4; ++ka)
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                        ::"r"(smem_u32(sm.a + ka * 16384)), "l"(reinterpret_cast<uint64_t>(&tmK)), "r"(smem_u32(&sm.a_full)),
                        "r"(ka * 64), "r"(0)
                        : "memory");
            }
            __syncwarp();
            uint32_t stage = 0, phase = 0;
            for (int tile = tile_lo; tile < tile_hi; tile += tile_step) {
                const FirTile t = fir_decode(p, tile);
                const int a0 = t.y0 / 2, b0 = t.x0 / 2;     // first cell of the tile
                mbar_wait(&sm.empty[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&sm.full[stage], (uint32_t)(ncb * 4 * FT_RA * FT_RB * 128));
                    for (int cb = 0; cb < ncb; ++cb) {
                        const int c = t.ct * p.block_n + cb * 64;
                        uint8_t *dst = sm.b + stage * stage_bytes + cb * FT_CB_BYTES;
                        tma_load_4d(dst + 0 * FT_PLANE_ROWS * 128, &tmT0, &sm.full[stage], c, b0 - 1, a0 - 1, t.n);
                        tma_load_4d(dst + 1 * FT_PLANE_ROWS * 128, &tmT1, &sm.full[stage], c, b0 - 1, a0 - 1, t.n);
                        tma_load_4d(dst + 2 * FT_PLANE_ROWS * 128, &tmT2, &sm.full[stage], c, b0 - 1, a0 - 1, t.n);
                        tma_load_4d(dst + 3 * FT_PLANE_ROWS * 128, &tmT3, &sm.full[stage], c, b0 - 1, a0 - 1, t.n);
                    }
                }
                __syncwarp();
                if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (tile_lo < tile_hi) {
            // kind::f16, D=f32, A=B=bf16, A K-major, B MN-major (bit 16), M=128, N=block_n
            const uint32_t idesc = make_idesc_bf16(128, (uint32_t)p.block_n) | (1u << 16);
            mbar_wait(&sm.a_full, 0);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            const uint32_t a_base = smem_u32(sm.a);
            for (int tile = tile_lo; tile < tile_hi; tile += tile_step) {
                mbar_wait(&sm.tmem_empty[acc], acc_phase ^ 1);
                mbar_wait(&sm.full[stage], phase);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * FT_N;
                const uint32_t b_base = smem_u32(sm.b) + stage * stage_bytes;
                const uint64_t adesc0 = make_smem_desc(a_base, 128);
                const uint64_t bdesc0 = make_smem_desc_mn(b_base, FT_CB_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < FT_K / 16; ++kk) {
                        const uint64_t adesc = adesc0 + (uint64_t)(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4);
                        const uint64_t bdesc = bdesc0 + (uint64_t)((kk * 16 * 128) >> 4);
                        umma_bf16(d_tmem, adesc, bdesc, idesc, kk != 0);
                    }
                    umma_commit(&sm.empty[stage]);
                    umma_commit(&sm.tmem_full[acc]);
                }
                __syncwarp();
                if (++stage == nstages) { stage = 0; phase ^= 1; }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue: 8 warps, two per TMEM lane quarter =====================
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int m = q * 32 + lane;                 // output pixel of the tile: (m / 8, m % 8)
        const int oy = m >> 3, ox = m & 7;
        const int et = threadIdx.x - 64;
        const int R = 2 * p.r, N = p.block_n;
        const float nw = p.noise ? __ldg(p.noise_weight) : 0.f;
        uint32_t acc = 0, acc_phase = 0;
        int staged_key = -1;
        for (int tile = tile_lo; tile < tile_hi; tile += tile_step) {
            const FirTile t = fir_decode(p, tile);
            const int Y = t.y0 + oy, X = t.x0 + ox, c0 = t.ct * N;
            const bool valid = Y < R && X < R;
            float nz = 0.f;
            if (valid && p.noise) nz = __ldg(p.noise + (long long)t.n * p.noise_bstride + (long long)Y * R + X);
            const int key = t.n * 64 + t.ct;
            if (key != staged_key) {
                asm volatile("bar.sync 1, 256;" ::: "memory");
                for (int i = et; i < N; i += 256) {
                    sm.e_bias[i] = __ldg(p.bias + c0 + i);
                    sm.e_next[i] = 1.41421356237f * __ldg(p.next_style + (long long)t.n * p.C + c0 + i);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                staged_key = key;
            }
            nz *= nw;
            // the previous tile's TMA store must have finished reading the staging buffer
            if (et == 0) tma_store_wait_read();
            asm volatile("bar.sync 2, 256;" ::: "memory");
            mbar_wait(&sm.tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * FT_N;
            for (int cc = 32 * half; cc < N; cc += 64) {
                uint32_t r[32];
                tmem_ld32(t_row + cc, r);
                tmem_ld_wait();
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = *reinterpret_cast<const float4 *>(&sm.e_bias[cc + j]);
                    const float4 s4 = *reinterpret_cast<const float4 *>(&sm.e_next[cc + j]);
                    float v0 = __uint_as_float(r[j + 0]) + nz + b4.x, v1 = __uint_as_float(r[j + 1]) + nz + b4.y;
                    float v2 = __uint_as_float(r[j + 2]) + nz + b4.z, v3 = __uint_as_float(r[j + 3]) + nz + b4.w;
                    v0 = fmaxf(v0, 0.2f * v0) * s4.x; v1 = fmaxf(v1, 0.2f * v1) * s4.y;
                    v2 = fmaxf(v2, 0.2f * v2) * s4.z; v3 = fmaxf(v3, 0.2f * v3) * s4.w;
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1 = __floats2bfloat162_rn(v2, v3);
                    packed[j / 2] = *reinterpret_cast<uint32_t *>(&h0);
                    packed[j / 2 + 1] = *reinterpret_cast<uint32_t *>(&h1);
                }
                // stage the 64 bytes of this pixel in the TMA layout: 16-byte chunk index XOR (row & 7)
                uint8_t *orow = sm.o + (cc >> 6) * FT_OUT_CB_BYTES + m * 128;
                const int ch0 = (cc & 63) >> 3;
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4)
                    *reinterpret_cast<uint4 *>(orow + (((ch0 + v4) ^ (m & 7)) << 4)) =
                        make_uint4(packed[4 * v4], packed[4 * v4 + 1], packed[4 * v4 + 2], packed[4 * v4 + 3]);
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.tmem_empty[acc]);
            asm volatile("bar.sync 3, 256;" ::: "memory");
            if (et == 0) {      // rows/cols beyond the image are clipped by the TMA unit
                for (int cb = 0; cb < ncb; ++cb)
                    tma_store_4d(&tmO, sm.o + cb * FT_OUT_CB_BYTES, c0 + cb * 64, t.x0, t.y0, t.n);
                tma_store_commit();
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    if (threadIdx.x == 64) tma_store_wait_all();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// host: Toeplitz matrix of the flipped taps kf[4][4] for the 16x8 tile, bf16 [128][256] row-major
void build_fir_toeplitz(uint16_t *out, const float *kf) {
    for (int m = 0; m < 128; ++m) {
        const int oy = m >> 3, ox = m & 7;
        for (int k = 0; k < FT_K; ++k) {
            const int plane = k / FT_PLANE_ROWS, idx = k % FT_PLANE_ROWS;
            float v = 0.f;
            if (idx < FT_RA * FT_RB) {
                const int ra = idx / FT_RB, rb = idx % FT_RB;
                const int py = plane >> 1, px = plane & 1;
                const int j = 2 * ra + py - oy - 1, i = 2 * rb + px - ox - 1;
                if (j >= 0 && j < 4 && i >= 0 && i < 4) v = kf[j * 4 + i];
            }
            __nv_bfloat16 h = __float2bfloat16_rn(v);
            out[m * FT_K + k] = *reinterpret_cast<uint16_t *>(&h);
        }
    }
}

int launch_upfir_tc(const UpfirTcParams &p, const CUtensorMap &tmK, const CUtensorMap *tmT, const CUtensorMap &tmO,
                    int sms, cudaStream_t st) {
    static_assert(sizeof(FirSmem) + 1024 <= 227 * 1024, "FirSmem exceeds the 227 KiB CTA limit");
    const size_t smem = sizeof(FirSmem) + 1024;
    static std::atomic<int> configured{0};
    if (!configured.load(std::memory_order_acquire)) {
        SG2_CUDA_OK(cudaFuncSetAttribute(upfir_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.store(1, std::memory_order_release);
    }
    SG2_REQUIRE(p.block_n == 64 || p.block_n == 128, SG2_ERR_BAD_ARG, "upfir_tc: block_n must be 64 or 128");
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    if (grid <= 0) return SG2_OK;
    upfir_tc_kernel<<<grid, FT_THREADS, smem, st>>>(p, tmK, tmT[0], tmT[1], tmT[2], tmT[3], tmO);
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

}  // namespace sg2
