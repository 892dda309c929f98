// upfirdn2d_planes.cu -- the op-API upfirdn2d on SMALL planes (<= 40 x 40 inputs: the 4^2 .. 32^2 octaves of the model, where
// a tensor is thousands of tiny NCHW planes), the three model geometries (blur 1/1, up-sampling 2/1, down-sampling 1/2), <= 4x4
// taps.  Same definition as upfirdn2d_kernel of the reference (op/upfirdn2d_kernel.cu:52-137): zero insertion, pad / crop,
// correlation with the flipped taps, decimation.
//
// The row-streaming kernel (upfirdn2d_stream.cu) gives every plane at least four lanes and a prefetch ring of whole rows: at
// 4 x 4 .. 32 x 32 pixels most lanes idle and the ring never fills (7-28 % of the HBM roofline, profiles/opbench_r02.jsonl).
// Here a CTA takes a BATCH of whole planes at a time -- they are contiguous in memory, so the batch is one contiguous run of
// elements, read with fully coalesced loads -- and keeps it in shared memory as fp32 tiles with a 4-pixel ZERO BORDER: the
// padding of the op is that border, the stencil needs no bounds checks.  Every thread then produces runs of four consecutive
// outputs of one row (taps and window offsets resolved at compile time per geometry and pad parity) and stores them as one
// vector where the row length allows.  Index decompositions (element -> plane, row, column) use an exact float-reciprocal
// division (indices < 2^15), not integer division.  fp32 accumulation for every storage dtype, taps applied per output in
// the reference's order (rows, then columns).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace sg2 {

constexpr int PL_BR = 4;              // widest zero border (pixels) around a staged plane; the launch uses what the geometry needs
constexpr int PL_THREADS = 256;
constexpr int PL_MAX_IN = 40;         // largest input extent handled here
constexpr int PL_TILE_FLOATS = 8192;  // shared-memory budget of one batch (32 KiB): 5 CTAs per SM

struct UfdPlanesParams {
    int in_h, in_w, out_h, out_w;
    int pad_x0, pad_y0, kh, kw;
    long long planes;
    int G;                 // planes per batch
    int sp, tile;          // pitch and size (floats) of one bordered plane tile
    int bl, bt;            // border to the left of / above the plane (the right / bottom border is in sp / tile)
    int quads;             // ceil(out_w / 4)
    float r_in_plane, r_in_w, r_quads, r_items;   // reciprocals for the exact float division below
    int vec_store;
};

// floor(i / d) for 0 <= i < 2^15 given r = 1 / d: (i + 0.5) * r is at least 0.5 / d away from an integer while its rounding
// error is below i * 2^-22 / d -- exact
__device__ __forceinline__ int fdiv(int i, float r) { return (int)(((float)i + 0.5f) * r); }

// PHX: parity of pad_x0 for UP == 2 (decides at compile time which tap columns meet the inserted zeros), 0 otherwise
template <typename T, int UP, int DOWN, int PHX>
__global__ void __launch_bounds__(PL_THREADS)
upfirdn2d_planes_kernel(T *__restrict__ out, const T *__restrict__ x, const float *__restrict__ taps, const UfdPlanesParams p) {
    extern __shared__ __align__(16) float pl_smem[];
    // flipped taps, zero padded to 4 x 4 (upfirdn2d_kernel.cu:71-81)
    float kf[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
            kf[a][b] = (a < p.kh && b < p.kw) ? __ldg(taps + (p.kh - 1 - a) * p.kw + (p.kw - 1 - b)) : 0.f;
    // borders are written once; the interiors are overwritten by every batch
    for (int i = threadIdx.x; i < p.G * p.tile; i += PL_THREADS) pl_smem[i] = 0.f;
    __syncthreads();

    const int in_plane = p.in_h * p.in_w, out_plane = p.out_h * p.out_w;
    const int items_per_plane = p.out_h * p.quads;
    const long long batches = (p.planes + p.G - 1) / p.G;
    for (long long batch = blockIdx.x; batch < batches; batch += gridDim.x) {
        const long long g0 = batch * p.G;
        const int g_here = (int)min((long long)p.G, p.planes - g0);
        // ---- load: one contiguous run of g_here * in_plane elements ----
        const T *src = x + g0 * in_plane;
        const int n_in = g_here * in_plane;
        for (int i = threadIdx.x; i < n_in; i += PL_THREADS) {
            const int g = fdiv(i, p.r_in_plane), rem = i - g * in_plane;
            const int iy = fdiv(rem, p.r_in_w), ix = rem - iy * p.in_w;
            pl_smem[g * p.tile + (iy + p.bt) * p.sp + ix + p.bl] = Cvt<T>::to_f(__ldg(src + i));
        }
        __syncthreads();
        // ---- compute: runs of four outputs ----
        T *dst = out + g0 * out_plane;
        const int n_items = g_here * items_per_plane;
        for (int it = threadIdx.x; it < n_items; it += PL_THREADS) {
            const int g = fdiv(it, p.r_items);
            const int rem = it - g * items_per_plane;
            const int oy = fdiv(rem, p.r_quads), ox0 = (rem - oy * p.quads) * 4;
            const float *tile = pl_smem + g * p.tile + p.bt * p.sp + p.bl;      // pixel (0, 0) of plane g
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            if constexpr (UP == 2) {
                // row Y = oy + a - pad_y0 of the up-sampled signal is an inserted zero row unless it is even: tap rows a = par, par + 2
                // (par from the output row: a run-time select between two tap rows, no divergent code).  Columns: ox0 is a multiple
                // of four, so X = ox0 + e + b - pad_x0 is even iff e + b - PHX is: resolved at compile time.
                const int par = (oy - p.pad_y0) & 1;
                const float *base = tile + ((ox0 - (p.pad_x0 - PHX)) >> 1);       // (even, possibly negative: inside the border)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int Y = oy + 2 * h + par - p.pad_y0;                    // even
                    const float *row = base + (Y >> 1) * p.sp;                    // arithmetic shift: exact, floor for border rows
                    float ka[4];
#pragma unroll
                    for (int b = 0; b < 4; ++b) ka[b] = par ? kf[2 * h + 1][b] : kf[2 * h][b];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const int Xc = e + b - PHX;
                            if (Xc & 1) continue;
                            acc[e] = fmaf(row[Xc >> 1], ka[b], acc[e]);
                        }
                }
            } else {
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    // tap row a meets input row oy * DOWN + a - pad_y0, tap column b of output ox0 + e column (ox0 + e) * DOWN + b - pad_x0
                    const float *row = tile + (oy * DOWN + a - p.pad_y0) * p.sp + ox0 * DOWN - p.pad_x0;
#pragma unroll
                    for (int e = 0; e < 4; ++e)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[e] = fmaf(row[e * DOWN + b], kf[a][b], acc[e]);
                }
            }
            T *o = dst + (long long)g * out_plane + oy * p.out_w + ox0;
            const int n_ok = min(4, p.out_w - ox0);
            if (p.vec_store && n_ok == 4) {
                T pk[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) pk[e] = Cvt<T>::from_f(acc[e]);
                if constexpr (sizeof(T) == 4) *reinterpret_cast<uint4 *>(o) = *reinterpret_cast<const uint4 *>(pk);
                else *reinterpret_cast<uint2 *>(o) = *reinterpret_cast<const uint2 *>(pk);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (e < n_ok) o[e] = Cvt<T>::from_f(acc[e]);
            }
        }
        __syncthreads();      // the next batch overwrites the tiles
    }
}

// Returns SG2_OK when the launch was made, 1 when this path does not apply (the caller goes on to the streaming / tiled
// kernels), another status on errors.
template <typename T>
int launch_upfirdn2d_planes(void *out, const void *x, const float *taps, int64_t planes, int in_h, int in_w, int out_h,
                            int out_w, int kh, int kw, int up, int down, int pad_x0, int pad_y0, cudaStream_t st) {
    if (!((up == 1 && (down == 1 || down == 2)) || (up == 2 && down == 1)) || kh > 4 || kw > 4) return 1;
    if (in_h > PL_MAX_IN || in_w > PL_MAX_IN || planes < 16) return 1;
    // where it measured faster than the streaming kernel (profiles/opbench_r02_planes.jsonl); SG2_UPFIRDN_PLANES=2 lifts the limits
    {
        const char *env = getenv("SG2_UPFIRDN_PLANES");
        const int big = std::max(in_h, in_w);
        const bool all = env && atoi(env) == 2;
        const int lim = up == 2 ? 8 : (down == 2 ? 24 : 20);
        if (!all && (big > lim || (down == 2 && big < 8))) return 1;
    }
    // every tap of every output (taps zero padded to 4 x 4, rows of four outputs) must fall inside the bordered tile: the border
    // each side needs, from the first / last row of the up-sampled signal that is touched
    auto border = [&](int pad0, int n_out, int n_in, int &lo_b, int &hi_b) {
        const int lo = -pad0, hi = (n_out - 1) * down + 3 - pad0;
        const int ilo = up == 2 ? (lo >> 1) : lo, ihi = up == 2 ? (hi >> 1) : hi;
        lo_b = std::max(0, -ilo);
        hi_b = std::max(0, ihi - (n_in - 1));
        return lo_b <= PL_BR && hi_b <= PL_BR + 4;
    };
    const int quads = (out_w + 3) / 4;
    int bt, bb, bl, br;
    if (!border(pad_y0, out_h, in_h, bt, bb) || !border(pad_x0, quads * 4, in_w, bl, br)) return 1;
    UfdPlanesParams p;
    p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w;
    p.pad_x0 = pad_x0; p.pad_y0 = pad_y0; p.kh = kh; p.kw = kw; p.planes = planes;
    p.bl = bl; p.bt = bt;
    p.sp = in_w + bl + br;
    p.tile = (in_h + bt + bb) * p.sp;
    const int sms = sm_count();
    int G = std::max(1, PL_TILE_FLOATS / p.tile);
    // enough batches for every SM to hold several CTAs' worth
    while (G > 1 && (planes + G - 1) / G < (int64_t)sms * 8) G = (G + 1) / 2;
    p.G = G;
    p.quads = quads;
    if ((int64_t)G * in_h * in_w >= 32768 || (int64_t)G * out_h * quads >= 32768) return 1;     // range of the float division
    p.r_in_plane = 1.0f / (float)(in_h * in_w); p.r_in_w = 1.0f / (float)in_w;
    p.r_quads = 1.0f / (float)quads; p.r_items = 1.0f / (float)(out_h * quads);
    const int es = (int)sizeof(T);
    p.vec_store = (out_w % 4 == 0 && reinterpret_cast<uintptr_t>(out) % (4 * es) == 0) ? 1 : 0;
    const size_t smem = (size_t)G * p.tile * sizeof(float);
    const int64_t batches = (planes + G - 1) / G;
    const int grid = (int)std::min<int64_t>(batches, (int64_t)sms * 5);
    T *o = (T *)out;
    const T *xi = (const T *)x;
    const int phx = up == 2 ? (pad_x0 & 1) : 0;
#define SG2_PL_LAUNCH(UP_, DOWN_, PX_) \
    upfirdn2d_planes_kernel<T, UP_, DOWN_, PX_><<<grid, PL_THREADS, smem, st>>>(o, xi, taps, p)
    if (up == 1 && down == 1) SG2_PL_LAUNCH(1, 1, 0);
    else if (up == 1 && down == 2) SG2_PL_LAUNCH(1, 2, 0);
    else if (phx == 0) SG2_PL_LAUNCH(2, 1, 0);
    else SG2_PL_LAUNCH(2, 1, 1);
#undef SG2_PL_LAUNCH
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

template int launch_upfirdn2d_planes<float>(void *, const void *, const float *, int64_t, int, int, int, int, int, int, int, int,
                                            int, int, cudaStream_t);
template int launch_upfirdn2d_planes<__half>(void *, const void *, const float *, int64_t, int, int, int, int, int, int, int, int,
                                             int, int, cudaStream_t);
template int launch_upfirdn2d_planes<__nv_bfloat16>(void *, const void *, const float *, int64_t, int, int, int, int, int, int,
                                                    int, int, int, int, cudaStream_t);

}  // namespace sg2
