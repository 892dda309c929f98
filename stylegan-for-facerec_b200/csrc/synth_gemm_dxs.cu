// synth_gemm_dxs.cu -- the narrow 3x3 styled convolutions of the 512^2 / 1024^2 octaves (64 -> 64, 32 -> 32; ModulatedConv2d
// + NoiseInjection + FusedLeakyReLU + ToRGB, model.py:232-273,282-287,331-359 of the reference) with the HORIZONTAL taps
// stacked along N.
//
// Why another formulation: a tcgen05 MMA whose operands both come from shared memory reads them at ~64 B/clk per SM
// (round-2 measurements: every layer's time matches (A + B bytes per MMA) / 64 B/clk when that exceeds the tensor floor).
// A 128-row A slice is 4 KiB per K = 16 step whatever N is, so at N = Cout = 32 the tensor pipe gets 16 clk of work for
// 80 clk of operand reads, and the tap-by-tap kernel (nine A views per tile) ran the 32 -> 32 layer at 18 % tensor pipe.
// Here one A view serves three taps:
//     F_dx[y, x'] = sum_{dy, ci} X[y + dy, x', ci] * W[dy, dx][co, ci]          (N = 3 * Cout columns (dx, co); K = 3 * Cin)
//     out[y, x]   = F_{-1}[y, x - 1] + F_0[y, x] + F_{+1}[y, x + 1]
// so a tile needs THREE A views (the dy shifts of one activation slab) instead of nine, each feeding an MMA three times as
// wide.  The tile is 4 rows x 32 pixels: a TMEM lane quarter is one image row, the +-1 pixel shift of the final sum is a
// warp shuffle, and the first / last column of a tile is halo (tiles advance by 30 pixels, 94 % of the rows are outputs).
//   A: one TMA box {Cin, 32 px, 6 rows} per tile (the dy halo inside the slab); the view of tap row dy starts (dy+1)*32 rows
//      into the slab (a whole number of swizzle atoms).   B: all 3 x [3*Cout, Cin] weight tiles resident in shared memory.
//   D: 4 accumulators of 128 columns (N = 96) or 2 of 256 (N = 192) in TMEM.
//   Warps: 0 = TMA producer, 1 = MMA issuer, 2..17 = epilogue (four per lane quarter; tile groups take alternate tiles so
//      that per-pixel work -- noise, coordinates, the ToRGB partial sums -- is done once per pixel).
// Epilogue = the styled-conv epilogue of synth_gemm.cu: x demod, + noise, + bias, lrelu, ToRGB partial dot products,
// x sqrt(2) * next layer's style, bf16, coalesced store (the last layer of the network stores only the ToRGB planes).
#include "common.cuh"
#include "synth_gemm.cuh"
#include "tc_ptx.cuh"

namespace sg2 {

using namespace tc;

constexpr int kDxEpiWarps = 16;
constexpr int kDxThreads = 64 + 32 * kDxEpiWarps;
constexpr int kDxTH = 4, kDxTW = 32, kDxStep = 30;       // tile 4 x 32 input columns, 30 of them outputs
constexpr int kDxRing = 176 * 1024;                      // resident weights + activation slab stages
constexpr int kDxMaxStages = 8;
constexpr float kDxSlope = 0.2f;

struct __align__(1024) DxsSmem {
    uint8_t ring[kDxRing];
    uint8_t stg[kDxEpiWarps][2048];          // per-warp transposition buffers of the coalesced epilogue store
    float e_demod[4][64], e_bias[4][64], e_next[4][64], e_w[4][3][64];   // per tile group
    uint64_t full[kDxMaxStages], empty[kDxMaxStages];
    uint64_t tmem_full[4], tmem_empty[4];
    uint64_t b_full;
    uint32_t tmem_base;
};

struct DxTile { int b, y0, x0; };
__device__ __forceinline__ DxTile dx_decode(const DxsParams &p, int tile) {
    DxTile t;
    const int bx = tile % p.tiles_x;
    tile /= p.tiles_x;
    const int by = tile % p.tiles_y;
    t.b = tile / p.tiles_y;
    t.x0 = bx * kDxStep - 1;          // input column of lane 0 (halo)
    t.y0 = by * kDxTH;
    return t;
}

__device__ __forceinline__ uint32_t dx_pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
}

__global__ void __launch_bounds__(kDxThreads, 1)
modconv_dxs_kernel(const __grid_constant__ DxsParams p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
    extern __shared__ uint8_t smem_raw[];
    DxsSmem &sm = *reinterpret_cast<DxsSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int N = 3 * p.Cout;
    const uint32_t nacc = N <= 128 ? 4u : 2u, acc_cols = 512u / nacc;
    const int col_groups = 4 / (int)nacc;                     // epilogue column split (1 at N = 96, 2 at N = 192)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < kDxMaxStages; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(&sm.tmem_full[i], 1); mbar_init(&sm.tmem_empty[i], 4 * col_groups); }
        mbar_init(&sm.b_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&sm.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;
    const uint32_t row_bytes = (uint32_t)p.Cin * 2u;                          // 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
    const uint32_t b_tile = (uint32_t)N * row_bytes;                          // one dy row of taps: [3*Cout, Cin]
    const uint32_t resb = 3u * b_tile;
    const uint32_t slab = (uint32_t)((kDxTH + 2) * kDxTW) * row_bytes;        // 192 pixel rows
    const uint32_t nstages = min((uint32_t)kDxMaxStages, ((uint32_t)kDxRing - resb) / slab);
    const int per = (p.total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int lo = min(p.total_tiles, (int)blockIdx.x * per), hi = min(p.total_tiles, lo + per);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            mbar_arrive_expect_tx(&sm.b_full, resb);
            for (int dy = 0; dy < 3; ++dy) tma_load_3d(sm.ring + dy * b_tile, &tmB, &sm.b_full, 0, 0, dy);
        }
        __syncwarp();
        uint32_t stage = 0, phase = 0;
        for (int tile = lo; tile < hi; ++tile) {
            const DxTile t = dx_decode(p, tile);
            mbar_wait(&sm.empty[stage], phase ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&sm.full[stage], slab);
                tma_load_4d(sm.ring + resb + stage * slab, &tmA, &sm.full[stage], 0, t.x0, t.y0 - 1, t.b);
            }
            __syncwarp();
            if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_bf16(kBlockM, (uint32_t)N);
        const uint32_t ring = smem_u32(sm.ring);
        const uint32_t view = (uint32_t)kDxTW * row_bytes;                   // a dy shift = 32 rows
        const uint32_t ksteps = (uint32_t)p.Cin / 16u;
        uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
        mbar_wait(&sm.b_full, 0);
        for (int tile = lo; tile < hi; ++tile) {
            mbar_wait(&sm.tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            mbar_wait(&sm.full[stage], phase);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * acc_cols;
            const uint32_t a0 = ring + resb + stage * slab;
            if (elect_one()) {
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    const uint64_t adesc = make_smem_desc(a0 + dy * view, row_bytes);
                    const uint64_t bdesc = make_smem_desc(ring + dy * b_tile, row_bytes);
                    // 32 bytes (>>4 = 2) inside the swizzle row per K = 16 step
                    umma_bf16(d_tmem, adesc, bdesc, idesc, dy != 0);
                    umma_bf16(d_tmem, adesc + 2, bdesc + 2, idesc, 1);
                    if (ksteps == 4) {
                        umma_bf16(d_tmem, adesc + 4, bdesc + 4, idesc, 1);
                        umma_bf16(d_tmem, adesc + 6, bdesc + 6, idesc, 1);
                    }
                }
                umma_commit(&sm.empty[stage]);
                umma_commit(&sm.tmem_full[acc]);
            }
            __syncwarp();
            if (++stage == nstages) { stage = 0; phase ^= 1; }
            if (++acc == nacc) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===================== epilogue: 16 warps =====================
        const int q = warp & 3;                              // TMEM lane quarter = tile row
        const int grp = (warp - 2) >> 2;
        const int tg = grp % (int)nacc, cg = grp / (int)nacc;   // tile group (takes tiles tg, tg + nacc, ...), column group
        const int gthreads = 128 * col_groups;               // threads that share this tile group's staged tables
        const int gt = cg * 128 + q * 32 + lane;
        const int bar_id = 1 + tg;
        const int Cout = p.Cout;
        const int cbase = cg * 32;                           // this warp's 32 output channels
        const float nw = p.noise ? __ldg(p.noise_weight) : 0.f;
        float *e_demod = sm.e_demod[tg], *e_bias = sm.e_bias[tg], *e_next = sm.e_next[tg];
        float *e_w0 = sm.e_w[tg][0], *e_w1 = sm.e_w[tg][1], *e_w2 = sm.e_w[tg][2];
        const uint32_t stg = smem_u32(sm.stg[warp - 2]);
        const long long plane = (long long)p.R * p.R;
        int staged_b = -1;
        for (int tile = lo + tg; tile < hi; tile += (int)nacc) {
            const uint32_t it = (uint32_t)(tile - lo);
            const uint32_t acc = it % nacc, acc_phase = (it / nacc) & 1u;
            const DxTile t = dx_decode(p, tile);
            const int y = t.y0 + q, x = t.x0 + lane;
            const bool valid = lane >= 1 && lane <= kDxStep && x < p.R;
            float nz = 0.f;
            if (valid && p.noise) nz = nw * __ldg(p.noise + (long long)t.b * p.noise_bstride + (long long)y * p.R + x);
            // last layer: bias + Upsample(skip) of this pixel (up 2, pad (2,1), 4x4 taps: a 2 x 2 neighbourhood of the skip image);
            // the loads are issued here so that they fly while the warp waits for its accumulator
            float up0 = 0.f, up1 = 0.f, up2 = 0.f;
            if (valid && p.image) {
                up0 = __ldg(p.rgb_bias); up1 = __ldg(p.rgb_bias + 1); up2 = __ldg(p.rgb_bias + 2);
                if (p.prev) {
                    const int S = p.R >> 1;
                    const int iy0 = (y - 1) >> 1, ky0 = 2 * iy0 + 2 - y, ix0 = (x - 1) >> 1, kx0 = 2 * ix0 + 2 - x;
                    const float *sp = p.prev + (long long)t.b * 3 * S * S;
#pragma unroll
                    for (int a2 = 0; a2 < 2; ++a2) {
                        const int iy = iy0 + a2;
                        if (iy < 0 || iy >= S) continue;
#pragma unroll
                        for (int b2 = 0; b2 < 2; ++b2) {
                            const int ix = ix0 + b2;
                            if (ix < 0 || ix >= S) continue;
                            const float w = p.kf[(ky0 + 2 * a2) * 4 + kx0 + 2 * b2];
                            const float *q = sp + (long long)iy * S + ix;
                            up0 = fmaf(w, __ldg(q), up0);
                            up1 = fmaf(w, __ldg(q + (long long)S * S), up1);
                            up2 = fmaf(w, __ldg(q + 2LL * S * S), up2);
                        }
                    }
                }
            }
            if (t.b != staged_b) {          // per-(sample, channel) tables of this tile group
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(gthreads) : "memory");
                for (int i = gt; i < Cout; i += gthreads) {
                    const long long o = (long long)t.b * Cout + i;
                    e_demod[i] = __ldg(p.demod + o);
                    e_bias[i] = __ldg(p.bias + i);
                    // lrelu gain sqrt(2) (fused_bias_act_kernel.cu:47) folded into both consumers
                    e_next[i] = p.next_style ? 1.41421356237f * __ldg(p.next_style + o) : 0.f;
                    if (p.rgb_w) {
                        const float rs = 1.41421356237f * __ldg(p.rgb_style + o);
                        e_w0[i] = rs * __ldg(p.rgb_w + i);
                        e_w1[i] = rs * __ldg(p.rgb_w + Cout + i);
                        e_w2[i] = rs * __ldg(p.rgb_w + 2 * Cout + i);
                    }
                }
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(gthreads) : "memory");
                staged_b = t.b;
            }
            mbar_wait(&sm.tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * acc_cols;
            float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
            uint32_t packed[16];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c0 = cbase + 16 * h;
                uint32_t fm[16], f0[16], fp[16];
                tmem_ld16(t_row + (uint32_t)c0, fm);                    // dx = -1 block
                tmem_ld16(t_row + (uint32_t)(Cout + c0), f0);           // dx =  0
                tmem_ld16(t_row + (uint32_t)(2 * Cout + c0), fp);       // dx = +1
                tmem_ld_wait();
                if (h == 1) {                // every TMEM read of this warp is done: hand the accumulator back early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.tmem_empty[acc]);
                }
                // (scalar fp32 math: a packed f32x2 version of this loop -- FFMA2 / FADD2 / FMUL2 over channel pairs -- spills
                //  at the 96 registers 576 threads leave and measured 1.52 -> 1.87 ms; kept out)
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 d4 = lds128f(smem_u32(e_demod + c0 + j)), b4 = lds128f(smem_u32(e_bias + c0 + j));
                    const float4 s4 = lds128f(smem_u32(e_next + c0 + j));
                    const float dd[4] = {d4.x, d4.y, d4.z, d4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        // out[x] = F_{-1}[x-1] + F_0[x] + F_{+1}[x+1]: the neighbours' blocks come by shuffle
                        const float a = __shfl_up_sync(0xffffffffu, __uint_as_float(fm[j + e]), 1);
                        const float c = __shfl_down_sync(0xffffffffu, __uint_as_float(fp[j + e]), 1);
                        const float sum = a + __uint_as_float(f0[j + e]) + c;
                        v[e] = fmaf(sum, dd[e], nz) + bb[e];
                        v[e] = fmaxf(v[e], kDxSlope * v[e]);             // lrelu (gain folded downstream)
                        if (p.clamp > 0.f) v[e] = fminf(fmaxf(v[e], -p.clamp), p.clamp);
                    }
                    if (p.rgb_w) {
                        const float4 w0 = lds128f(smem_u32(e_w0 + c0 + j)), w1 = lds128f(smem_u32(e_w1 + c0 + j));
                        const float4 w2 = lds128f(smem_u32(e_w2 + c0 + j));
                        rgb0 += v[0] * w0.x + v[1] * w0.y + v[2] * w0.z + v[3] * w0.w;
                        rgb1 += v[0] * w1.x + v[1] * w1.y + v[2] * w1.z + v[3] * w1.w;
                        rgb2 += v[0] * w2.x + v[1] * w2.y + v[2] * w2.z + v[3] * w2.w;
                    }
                    packed[8 * h + j / 2 + 0] = dx_pack(v[0] * ss[0], v[1] * ss[1]);
                    packed[8 * h + j / 2 + 1] = dx_pack(v[2] * ss[2], v[3] * ss[3]);
                }
            }
            if (p.out) {
                uint32_t off16 = 0xffffffffu;     // this pixel's 64 bytes in 16-byte units from p.out (0xFFFFFFFF: not stored)
                if (valid) off16 = (uint32_t)(((((long long)t.b * p.R + y) * p.R + x) * Cout + cbase) >> 3);
                store_rows64_coalesced(stg, packed, off16, reinterpret_cast<uint8_t *>(p.out), lane);
            }
            if (valid && p.rgb_w && p.image) {    // final image: ToRGB sum + (bias + 2x up-sampled skip, gathered before the wait)
                float *ip = p.image + ((long long)t.b * 3) * plane + (long long)y * p.R + x;
                ip[0] = rgb0 + up0;
                ip[plane] = rgb1 + up1;
                ip[2 * plane] = rgb2 + up2;
            } else if (valid && p.rgb_w) {        // each column group writes its own partial plane
                float *rp = p.rgb_part + (((long long)cg * p.B + t.b) * 3) * plane + (long long)y * p.R + x;
                rp[0] = rgb0;
                rp[plane] = rgb1;
                rp[2 * plane] = rgb2;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int launch_modconv_dxs(const DxsParams &p, const CUtensorMap &tmA, const CUtensorMap &tmB, int sms, cudaStream_t st) {
    static_assert(sizeof(DxsSmem) + 1024 <= 227 * 1024, "DxsSmem exceeds the 227 KiB CTA limit");
    const size_t smem = sizeof(DxsSmem) + 1024;
    static std::atomic<int> configured{0};
    if (!configured.load(std::memory_order_acquire)) {
        SG2_CUDA_OK(cudaFuncSetAttribute(modconv_dxs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.store(1, std::memory_order_release);
    }
    SG2_REQUIRE((p.Cin == 32 || p.Cin == 64) && (p.Cout == 32 || p.Cout == 64) && p.R >= 32 && p.R % kDxTH == 0 && p.demod && p.bias,
                SG2_ERR_BAD_ARG, "gemm(dx-stacked): bad plan (Cin %d, Cout %d, R %d)", p.Cin, p.Cout, p.R);
    SG2_REQUIRE(!p.image || (p.Cout == 32 && p.rgb_w && p.rgb_bias), SG2_ERR_BAD_ARG, "gemm(dx-stacked): the fused final image needs Cout = 32 and a ToRGB");
    SG2_REQUIRE(p.tiles_x == (p.R + kDxStep - 1) / kDxStep && p.tiles_y == p.R / kDxTH && p.total_tiles == p.tiles_x * p.tiles_y * p.B,
                SG2_ERR_BAD_ARG, "gemm(dx-stacked): bad tile grid");
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    if (grid <= 0) return SG2_OK;
    modconv_dxs_kernel<<<grid, kDxThreads, smem, st>>>(p, tmA, tmB);
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

}  // namespace sg2
