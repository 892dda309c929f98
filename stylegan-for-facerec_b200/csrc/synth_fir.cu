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
//     the host per plan, resident in TENSOR MEMORY for the whole kernel (A operand of tcgen05.mma).
//   * fp32 accumulation in TMEM; the epilogue warps only do +noise, +bias, lrelu, x style, pack -- and
//     read the noise tile / bias / style from shared memory, where the TMA producer put them (FirAux).
// The multiplications by structural zeros cost 16x more MACs than the stencil (still < 30 % of
// the tile's HBM time on the tensor pipe) and buy an instruction stream 4x shorter.
//
// A tile is always 128 accumulator columns wide, cut into column blocks of `cbw` channels that each
// map to one (sample, channel offset): C >= 128 -> two 64-channel blocks of one sample; C = 64 -> two
// samples; C = 32 (the 1024^2 tail) -> four samples with 64-byte rows (SWIZZLE_64B operand atoms).
// Every epilogue warp owns 32 pixels x 32 columns and stages them in shared memory in the TMA store layout
// (double buffered); the tile leaves through TMA stores (direct 16-byte register stores measured 20 % slower: profiles/experiments).
#include "common.cuh"
#include "synth_kernels.cuh"
#include "tc_ptx.cuh"

namespace sg2 {

using namespace tc;

constexpr int FT_TH = 16, FT_TW = 8;             // output tile (pixels)
constexpr int FT_RA = 10, FT_RB = 6;             // window rows / cols per plane
constexpr int FT_PLANE_ROWS = 64;                // 60 used + 4 zero pad rows
constexpr int FT_K = 4 * FT_PLANE_ROWS;          // 256
constexpr int FT_N = 128;                        // accumulator columns per tile
constexpr int FT_STAGE_BYTES = FT_K * FT_N * 2;  // the window of one tile: 64 KiB whatever the column block width
// The Toeplitz matrix (A operand) lives in TENSOR MEMORY (128 lanes x 128 columns of packed bf16 pairs, written once
// per CTA with tcgen05.st): the MMAs read only the window from shared memory, and the 64 KiB it would take there hold
// the second output staging buffer and the ring of per-tile epilogue inputs instead.
constexpr int FT_STAGES = 2;
constexpr int FT_OBUF = 2;                       // output staging buffers: tile i is staged while tile i - 1 is still being stored
constexpr int FT_AUX = FT_STAGES + 2;            // ring of per-tile epilogue inputs (see FirAux)
constexpr int FT_TMEM_COLS = 512;                // 2 accumulators (2 x 128 columns) + 128 columns of A
#ifndef SG2_FIR_EPI_WARPS
#define SG2_FIR_EPI_WARPS 16
#endif
constexpr int FT_EPI_WARPS = SG2_FIR_EPI_WARPS;  // 4 warps per TMEM lane quarter: a 32-column chunk each (8: two chunks each)
constexpr int FT_EPI_THREADS = 32 * FT_EPI_WARPS;
constexpr int FT_CPW = 16 / FT_EPI_WARPS;        // 32-column chunks per epilogue warp
constexpr int FT_THREADS = 64 + FT_EPI_THREADS;
static_assert(FT_EPI_WARPS == 8 || FT_EPI_WARPS == 16, "epilogue: 2 or 4 warps per TMEM lane quarter");

// Everything the epilogue needs for one tile besides the accumulator: the tile's noise (16 x 8 pixels per sample), the bias and
// the consumer's style of its 128 columns.  The TMA producer fetches it together with the tile's window (one tensor load + bulk
// copies on their own mbarrier), FT_AUX tiles deep.  The epilogue warps issue NO global loads: an ordinary load queues behind the
// ~120 KiB of tensor loads this SM keeps in flight (Little's law: 148 SMs x 120 KiB / 3.7 TB/s ~ 5 us), and with the noise / the
// parameter table fetched by the epilogue threads themselves -- even two tiles ahead -- that wait was 22 % of the kernel's time
// (knock-out: 0.555 -> 0.435 ms on the 256^2 layer with those loads removed).
struct __align__(128) FirAux {
    float noise[4][FT_TH * FT_TW];                 // [sample of the tile][pixel]; one map when the noise is broadcast over the batch
    float bias[FT_N];
    float style[FT_N];                             // style of the consumer (without the sqrt(2) gain)
};

struct __align__(1024) FirSmem {
    uint8_t b[FT_STAGES * FT_STAGE_BYTES];
    uint8_t o[FT_OBUF][128 * FT_N * 2];            // bf16 output tile [column block][128 px][cbw] in the TMA store layout
    FirAux aux[FT_AUX];
    uint64_t full[FT_STAGES], empty[FT_STAGES], aux_full[FT_AUX];
    uint64_t tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
};

struct FirTile { int n0, ct, y0, x0; };

// Tiles are dealt round-robin (tile = blockIdx.x + i * gridDim.x); the walker keeps the mixed-radix digits
// (x, y, channel group, sample group) and adds the decomposed stride with carries: no division per tile.
struct FirWalk {
    int bx, by, ct, ng;          // current digits
    int sx, sy, sc, sn;          // stride digits
    __device__ __forceinline__ void init(const UpfirTcParams &p, int tile, int step) {
        bx = tile % p.tiles_x; tile /= p.tiles_x;
        by = tile % p.tiles_y; tile /= p.tiles_y;
        ct = tile % p.tiles_c; ng = tile / p.tiles_c;
        sx = step % p.tiles_x; step /= p.tiles_x;
        sy = step % p.tiles_y; step /= p.tiles_y;
        sc = step % p.tiles_c; sn = step / p.tiles_c;
    }
    __device__ __forceinline__ void next(const UpfirTcParams &p) {
        bx += sx;
        int c = bx >= p.tiles_x;
        bx -= c ? p.tiles_x : 0;
        by += sy + c;
        c = by >= p.tiles_y;
        by -= c ? p.tiles_y : 0;
        ct += sc + c;
        c = ct >= p.tiles_c;
        ct -= c ? p.tiles_c : 0;
        ng += sn + c;
    }
    __device__ __forceinline__ FirTile tile(const UpfirTcParams &p) const {
        FirTile t;
        t.n0 = ng * p.nsamp; t.ct = ct; t.y0 = by * FT_TH; t.x0 = bx * FT_TW;
        return t;
    }
};

// MN-major swizzled operand (mma_traits_sm100.hpp canonical layouts): rows of `row_bytes` (128: SWIZZLE_128B,
// 64: SWIZZLE_64B) along N, 8-row groups along K 8*row_bytes apart (SBO), N atoms `lbo_bytes` apart (LBO).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2 : 4;
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)((8 * row_bytes) >> 4) << 32) | (1ull << 46) | (layout << 61);
}

// plain (non-tensor) bulk copy global -> shared, completing on an mbarrier; 16-byte aligned addresses and size
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(FT_THREADS, 1)
upfir_tc_kernel(const __grid_constant__ UpfirTcParams p, const __grid_constant__ CUtensorMap tmN,
                const __grid_constant__ CUtensorMap tmT0, const __grid_constant__ CUtensorMap tmT1,
                const __grid_constant__ CUtensorMap tmT2, const __grid_constant__ CUtensorMap tmT3,
                const __grid_constant__ CUtensorMap tmO) {
    // the kernel has no static shared memory: the dynamic window starts at the CTA's shared base, which satisfies the
    // declared 1 KiB alignment (swizzle atoms)
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    FirSmem &sm = *reinterpret_cast<FirSmem *>(smem_raw);
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cbw = p.cbw, ncb = FT_N / cbw;                 // column block width (channels) and count
    const uint32_t row_bytes = (uint32_t)cbw * 2, cb_bytes = FT_K * row_bytes;
    const int cps = ncb / p.nsamp;                           // column blocks per sample

    // zero the pad rows (60..63 of every plane region of every column block) once: TMA never writes them
    {
        const int v16 = (int)row_bytes / 16;                 // 16-byte pieces per row
        const int total = FT_STAGES * ncb * 4 * 4 * v16;
        for (int i = threadIdx.x; i < total; i += FT_THREADS) {
            const int v = i % v16, row = (i / v16) & 3, plane = (i / (v16 * 4)) & 3, cb = i / (v16 * 16);
            *reinterpret_cast<uint4 *>(&sm.b[cb * cb_bytes + (plane * FT_PLANE_ROWS + 60 + row) * row_bytes + v * 16]) =
                make_uint4(0, 0, 0, 0);
        }
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmT0);
        tma_prefetch_desc(&tmO);
        for (int i = 0; i < FT_STAGES; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], 1); }
        for (int i = 0; i < FT_AUX; ++i) mbar_init(&sm.aux_full[i], 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&sm.tmem_full[i], 1); mbar_init(&sm.tmem_empty[i], FT_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&sm.tmem_base, FT_TMEM_COLS);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // pad-row zeros visible to the MMA (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;
    // Toeplitz matrix -> TMEM columns [256, 384): lane = output pixel m, column j = taps (2j, 2j+1) of its row.
    // The first four epilogue warps cover the four lane quarters.
    if (warp >= 2 && warp < 6) {
        const int mrow = (warp & 3) * 32 + lane;
        const uint4 *src = reinterpret_cast<const uint4 *>(p.toeplitz + (size_t)mrow * FT_K);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            uint32_t r[32];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const uint4 q4 = __ldg(src + ch * 8 + v);
                r[4 * v] = q4.x; r[4 * v + 1] = q4.y; r[4 * v + 2] = q4.z; r[4 * v + 3] = q4.w;
            }
            tmem_st32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256u + 32u * ch, r);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // round-robin: at any moment the 148 CTAs work on ~5 adjacent tile rows, so the halo rows a tile
    // shares with its vertical neighbours are still in L2 when the neighbour loads them
    const int tile_lo = (int)blockIdx.x, tile_step = (int)gridDim.x, tile_hi = p.total_tiles;
    const int noise_maps = p.noise ? (p.noise_bstride ? p.nsamp : 1) : 0;    // noise maps per tile

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
        if (tile_lo < tile_hi) {
            const uint32_t plane_bytes = FT_PLANE_ROWS * row_bytes;
            // per column block: channel offset inside the tile's channel group and sample offset (tile-invariant)
            int coff[4], noff[4];
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) { coff[cb] = (cb % cps) * cbw; noff[cb] = cb / cps; }
            const int cgroup = cps * cbw;
            const uint32_t aux_bytes = (uint32_t)(2 * FT_N * 4 + noise_maps * FT_TH * FT_TW * 4);
            uint32_t stage = 0, phase = 0, slot = 0;
            FirWalk w;
            w.init(p, tile_lo, tile_step);
            for (int tile = tile_lo; tile < tile_hi; tile += tile_step, w.next(p)) {
                const FirTile t = w.tile(p);
                const int a0 = t.y0 / 2 - 1, b0 = t.x0 / 2 - 1;     // first window cell of the tile
                // stage free <=> the MMAs of tile i - 2 are done <=> (accumulator hand-back) the epilogue of tile i - 4 has read
                // its aux slot, which is the one this tile overwrites (FT_AUX = FT_STAGES + 2)
                mbar_wait(&sm.empty[stage], phase ^ 1);
                if (elect_one()) {
                    if (SG2_DBG(p) & 1) {
                        mbar_arrive(&sm.full[stage]);
                    } else {
                        mbar_arrive_expect_tx(&sm.full[stage], (uint32_t)(ncb * 4 * FT_RA * FT_RB) * row_bytes);
#pragma unroll
                        for (int cb = 0; cb < 4; ++cb) {
                            if (cb < ncb) {
                                const int c = t.ct * cgroup + coff[cb], n = t.n0 + noff[cb];   // samples >= B read as zero
                                uint8_t *dst = sm.b + stage * FT_STAGE_BYTES + cb * cb_bytes;
                                tma_load_4d(dst + 0 * plane_bytes, &tmT0, &sm.full[stage], c, b0, a0, n);
                                tma_load_4d(dst + 1 * plane_bytes, &tmT1, &sm.full[stage], c, b0, a0, n);
                                tma_load_4d(dst + 2 * plane_bytes, &tmT2, &sm.full[stage], c, b0, a0, n);
                                tma_load_4d(dst + 3 * plane_bytes, &tmT3, &sm.full[stage], c, b0, a0, n);
                            }
                        }
                    }
                    FirAux &ax = sm.aux[slot];
                    mbar_arrive_expect_tx(&sm.aux_full[slot], aux_bytes);
                    // pixels / samples beyond the tensor arrive as zeros; they are clipped by the store anyway
                    if (noise_maps) tma_load_3d(ax.noise, &tmN, &sm.aux_full[slot], t.x0, t.y0, p.noise_bstride ? t.n0 : 0);
#pragma unroll
                    for (int cb = 0; cb < 4; ++cb) {
                        if (cb < ncb) {
                            const int c = t.ct * cgroup + coff[cb], n = min(t.n0 + noff[cb], p.B - 1);
                            bulk_load(ax.bias + cb * cbw, p.bias + c, (uint32_t)cbw * 4u, &sm.aux_full[slot]);
                            bulk_load(ax.style + cb * cbw, p.next_style + (long long)n * p.C + c, (uint32_t)cbw * 4u, &sm.aux_full[slot]);
                        }
                    }
                }
                __syncwarp();
                if (++stage == FT_STAGES) { stage = 0; phase ^= 1; }
                slot = (slot + 1) & (FT_AUX - 1);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (tile_lo < tile_hi) {
            // kind::f16, D=f32, A=B=bf16, A from tensor memory (K-major), B MN-major (bit 16), M=128, N=128
            const uint32_t idesc = make_idesc_bf16(128, FT_N) | (1u << 16);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            const uint32_t kstep = (16 * row_bytes) >> 4;        // 16 K rows per MMA, in 16-byte units
            for (int tile = tile_lo; tile < tile_hi; tile += tile_step) {
                mbar_wait(&sm.tmem_empty[acc], acc_phase ^ 1);
                mbar_wait(&sm.full[stage], phase);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * FT_N;
                const uint64_t bdesc0 = make_smem_desc_mn(smem_u32(sm.b) + stage * FT_STAGE_BYTES, cb_bytes, row_bytes);
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < ((SG2_DBG(p) & 8) ? 1 : FT_K / 16); ++kk)      // 16 taps = 8 TMEM columns per K step
                        umma_bf16_ts(d_tmem, tmem_base + 256u + 8u * kk, bdesc0 + (uint64_t)(kk * kstep), idesc, kk != 0);
                    umma_commit(&sm.empty[stage]);
                    umma_commit(&sm.tmem_full[acc]);
                }
                __syncwarp();
                if (++stage == FT_STAGES) { stage = 0; phase ^= 1; }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue: FT_EPI_WARPS warps, FT_EPI_WARPS / 4 per TMEM lane quarter =====
        // The knock-out analysis (profiles/experiments) showed this role, not the loads or the MMAs, paces the
        // kernel: 16 warps (one 32-column chunk each, 4 per scheduler to hide the latencies), packed f32x2 math, every
        // input of the tile already in shared memory when the accumulator completes, one block-wide barrier per tile.
        // warp (q, kq) owns pixels 32q..32q+31 x chunks kq*FT_CPW .. of every tile
        const int q = warp & 3, kq = (warp - 2) >> 2;
        const int m = q * 32 + lane;                 // output pixel of the tile: (m / 8, m % 8)
        const int et = threadIdx.x - 64;
        const float nw = p.noise ? __ldg(p.noise_weight) : 0.f;
        // the noise map behind each of this warp's chunks (the sample of the chunk's column block, or the one broadcast map)
        int ds[FT_CPW];
#pragma unroll
        for (int ci = 0; ci < FT_CPW; ++ci) ds[ci] = (p.noise && p.noise_bstride) ? ((32 * (kq * FT_CPW + ci)) / cbw) / cps : 0;
        const float2 gain2 = make_float2(1.41421356237f, 1.41421356237f), slope2 = make_float2(0.2f, 0.2f);
        uint32_t acc = 0, acc_phase = 0, it = 0;
        FirWalk w;
        w.init(p, tile_lo, tile_step);
        for (int tile = tile_lo; tile < tile_hi; tile += tile_step, ++it) {
            const FirTile t = w.tile(p);
            w.next(p);
            const uint32_t slot = it & (FT_AUX - 1);
            const FirAux &ax = sm.aux[slot];
            const uint32_t bias_s = smem_u32(ax.bias), style_s = smem_u32(ax.style);
            const uint32_t o_s = smem_u32(sm.o[it & 1u]);
            mbar_wait(&sm.aux_full[slot], (it / FT_AUX) & 1u);
            mbar_wait(&sm.tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * FT_N;
            uint32_t packed[FT_CPW][16];
#pragma unroll
            for (int ci = 0; ci < FT_CPW; ++ci) {
                if (SG2_DBG(p) & 4) break;
                const int cc = 32 * (kq * FT_CPW + ci);          // first accumulator column of the chunk
                const float nz = noise_maps ? nw * ax.noise[ds[ci]][m] : 0.f;
                const float2 nz2 = make_float2(nz, nz);
                uint32_t r[32];
                tmem_ld32(t_row + cc, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = lds128f(bias_s + (uint32_t)(cc + j) * 4u), s4 = lds128f(style_s + (uint32_t)(cc + j) * 4u);
                    // (acc + noise + bias) -> lrelu -> x sqrt(2) * style of the consumer, two columns per instruction
                    float2 a = __fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(r[j + 0]), __uint_as_float(r[j + 1])), nz2),
                                          make_float2(b4.x, b4.y));
                    float2 b = __fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), nz2),
                                          make_float2(b4.z, b4.w));
                    const float2 al = __fmul2_rn(a, slope2), bl = __fmul2_rn(b, slope2);
                    const float2 sa = __fmul2_rn(make_float2(s4.x, s4.y), gain2), sb = __fmul2_rn(make_float2(s4.z, s4.w), gain2);
                    a = __fmul2_rn(make_float2(fmaxf(a.x, al.x), fmaxf(a.y, al.y)), sa);
                    b = __fmul2_rn(make_float2(fmaxf(b.x, bl.x), fmaxf(b.y, bl.y)), sb);
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(b.x, b.y);
                    packed[ci][j / 2] = *reinterpret_cast<uint32_t *>(&h0);
                    packed[ci][j / 2 + 1] = *reinterpret_cast<uint32_t *>(&h1);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.tmem_empty[acc]);     // accumulator drained, aux slot read: the MMA warp may go on
            // staging buffer it & 1 is free: before the barrier of the previous tile, thread 0 waited for the store of tile
            // it - 2 to finish reading it
#pragma unroll
            for (int ci = 0; ci < FT_CPW; ++ci) {
                if (SG2_DBG(p) & 4) break;
                const int k = kq * FT_CPW + ci;                  // 32-column chunk of the tile
                // stage this pixel's 64 bytes in the TMA layout of the store box
                if (cbw == 64) {      // [128 px][128 B] per column block, SWIZZLE_128B: 16-byte chunk index XOR (row & 7)
                    const uint32_t orow = o_s + (uint32_t)((k >> 1) * (128 * 128) + m * 128);
#pragma unroll
                    for (int v4 = 0; v4 < 4; ++v4)
                        sts128(orow + (uint32_t)(((4 * (k & 1) + v4) ^ (m & 7)) << 4), packed[ci][4 * v4], packed[ci][4 * v4 + 1],
                               packed[ci][4 * v4 + 2], packed[ci][4 * v4 + 3]);
                } else {              // [128 px][64 B] per column block, SWIZZLE_64B: chunk index XOR ((row >> 1) & 3)
                    const uint32_t orow = o_s + (uint32_t)(k * (128 * 64) + m * 64);
#pragma unroll
                    for (int v4 = 0; v4 < 4; ++v4)
                        sts128(orow + (uint32_t)((v4 ^ ((m >> 1) & 3)) << 4), packed[ci][4 * v4], packed[ci][4 * v4 + 1],
                               packed[ci][4 * v4 + 2], packed[ci][4 * v4 + 3]);
                }
            }
            if (!(SG2_DBG(p) & 2)) {
                fence_proxy_async();
                // the store of the previous tile must have read ITS buffer before the next tile is staged there
                if (et == 0) tma_store_wait_read();
                asm volatile("bar.sync 1, %0;" ::"n"(FT_EPI_THREADS) : "memory");
                if (et == 0) {          // rows / columns / samples beyond the tensor are clipped by the TMA unit
                    for (int cb = 0; cb < ncb; ++cb)
                        tma_store_4d(&tmO, sm.o[it & 1u] + cb * (128 * (int)row_bytes), t.ct * (cps * cbw) + (cb % cps) * cbw, t.x0,
                                     t.y0, t.n0 + cb / cps);
                    tma_store_commit();
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (et == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, FT_TMEM_COLS);
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

int launch_upfir_tc(const UpfirTcParams &p, const CUtensorMap &tmN, const CUtensorMap *tmT, const CUtensorMap &tmO,
                    int sms, cudaStream_t st) {
    static_assert(sizeof(FirSmem) <= 227 * 1024, "FirSmem exceeds the 227 KiB CTA limit");
    const size_t smem = sizeof(FirSmem);
    static std::atomic<int> configured{0};
    if (!configured.load(std::memory_order_acquire)) {
        SG2_CUDA_OK(cudaFuncSetAttribute(upfir_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.store(1, std::memory_order_release);
    }
    SG2_REQUIRE((p.cbw == 64 || p.cbw == 32) && p.nsamp >= 1 && (FT_N / p.cbw) % p.nsamp == 0, SG2_ERR_BAD_ARG,
                "upfir_tc: bad column blocking (cbw %d, %d samples per tile)", p.cbw, p.nsamp);
    SG2_REQUIRE(p.nsamp <= 4 && p.B >= 1, SG2_ERR_BAD_ARG, "upfir_tc: at most 4 samples per tile");
    SG2_REQUIRE((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.next_style) & 15) == 0 && p.C % 4 == 0,
                SG2_ERR_BAD_ARG, "upfir_tc: the bias / style rows are fetched with 16-byte bulk copies and must be 16-byte aligned");
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    if (grid <= 0) return SG2_OK;
    upfir_tc_kernel<<<grid, FT_THREADS, smem, st>>>(p, tmN, tmT[0], tmT[1], tmT[2], tmT[3], tmO);
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

}  // namespace sg2
