// upfirdn2d.cu -- shared-memory-staged FIR up/down/blur stencils (HBM-bound).
//
// Replaces upfirdn2d_kernel (op/upfirdn2d_kernel.cu:52-137 of the reference).  Same definition:
// zero-insert upsample by `up`, pad (negative pads crop), true 2-D convolution with the taps
// (correlation with the flipped taps, :77), keep every `down`-th sample; output size as :167-168.
//
// Design (differs from the reference, which computes one output per thread with 16 smem tap
// reads each and supports only six hard-coded modes):
//   * the three geometries the model uses -- blur (1,1), skip upsample (2,1), downsample (1,2),
//     each with <= 4x4 taps -- are compile-time polyphase stencils: the tile grid is shifted so
//     that the polyphase phase of every output of a thread is a compile-time constant, the
//     thread keeps its input window, the 16 taps and a TYxTX block of outputs in registers and
//     zero taps of the polyphase decomposition are never multiplied;
//   * accumulation is always fp32, whatever the storage dtype (the reference accumulates fp16
//     in fp16);
//   * everything else (up/down <= 4, taps <= 16x16, minor > 1) goes through a generic kernel
//     with the same staging; configurations outside that return SG2_ERR_UNSUPPORTED instead of
//     uninitialised memory (:172-268 of the reference has no default case).
#include <stdlib.h>

#include "common.cuh"

namespace sg2 {

__host__ __device__ __forceinline__ int floordiv(int a, int b) {
    int q = a / b;
    return (q * b > a) ? q - 1 : q;
}

constexpr int GEN_TILE_W = 32, GEN_TILE_H = 8, GEN_MAXK = 16, GEN_MAXF = 4;
constexpr int GEN_IN_W = ((GEN_TILE_W - 1) * GEN_MAXF + GEN_MAXK - 1) + 1;  // worst case, up = 1
constexpr int GEN_IN_H = ((GEN_TILE_H - 1) * GEN_MAXF + GEN_MAXK - 1) + 1;

struct UfdGenParams {
    int in_h, in_w, out_h, out_w, minor;
    int up_x, up_y, down_x, down_y, pad_x0, pad_y0, kh, kw;
    int64_t major;
};

struct UfdParams {
    int in_h, in_w, out_h, out_w;
    int pad_x0, pad_y0;
    int kh, kw;
    int shift_x, shift_y;   // tile-grid shift so that phases are compile-time
    int vec_store;          // rows of a thread's TX outputs are aligned for one vector store
    int64_t planes;
};

// ---- specialised polyphase kernel ---------------------------------------------------------------
// One block = 32 x 8 threads; thread (tx,ty) produces TX x TY outputs; taps padded to K x K.
template <int UP, int DOWN, int K, int TX, int TY>
struct Geo {
    // offset of output j (relative to an aligned thread origin) into the input window and tap phase
    __host__ __device__ static constexpr int in_off(int j) { return (j * DOWN + UP - 1) / UP; }
    __host__ __device__ static constexpr int k0(int j) { return (in_off(j) + 1) * UP - (j * DOWN + UP - 1) - 1; }
    __host__ __device__ static constexpr int ntap(int j) { return (K - k0(j) + UP - 1) / UP; }
    static constexpr int WIN_X = in_off(TX - 1) + ntap(TX - 1);   // window width per thread
    static constexpr int WIN_Y = in_off(TY - 1) + ntap(TY - 1);
    static constexpr int TILE_W = 32 * TX, TILE_H = 8 * TY;
    // input tile staged per block
    static constexpr int IN_W = in_off(TILE_W - 1) + ntap(TX - 1);
    static constexpr int IN_H = in_off(TILE_H - 1) + ntap(TY - 1);
    // a thread's window starts at WSTRIDE * tx floats: read it with WSTRIDE-wide vector loads so a
    // warp touches consecutive 8/16-byte chunks (no bank conflicts)
    static constexpr int WSTRIDE = TX * DOWN / UP;
    static constexpr int VEC = WSTRIDE >= 4 ? 4 : 2;
    static constexpr int NV = (WIN_X + VEC - 1) / VEC;
    static constexpr int IN_NEED = 31 * WSTRIDE + NV * VEC;
    static constexpr int IN_WP = ((IN_W > IN_NEED ? IN_W : IN_NEED) + 3) & ~3;
    static_assert(WSTRIDE * UP == TX * DOWN && (WSTRIDE == 2 || WSTRIDE == 4), "window stride");
};

template <typename T, int UP, int DOWN, int K, int TX, int TY>
__global__ void __launch_bounds__(256)
upfirdn2d_poly_kernel(T *__restrict__ out, const T *__restrict__ x, const float *__restrict__ taps,
                      UfdParams p) {
    using G = Geo<UP, DOWN, K, TX, TY>;
    __shared__ __align__(16) float s_in[G::IN_H * G::IN_WP];
    __shared__ float s_k[K * K];

    const int tid = threadIdx.x;
    if (tid < K * K) {   // flipped taps, zero padded to K x K  (upfirdn2d_kernel.cu:71-81)
        const int ky = tid / K, kx = tid % K;
        s_k[tid] = (ky < p.kh && kx < p.kw) ? taps[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)] : 0.f;
    }
    // tile origin in output coordinates (may be negative by the phase shift)
    const int ox0 = blockIdx.x * G::TILE_W - p.shift_x;
    const int oy0 = blockIdx.y * G::TILE_H - p.shift_y;
    // first input sample the tile touches:  (o*DOWN - pad) is a multiple of UP by construction
    const int ix0 = (ox0 * DOWN - p.pad_x0) / UP;
    const int iy0 = (oy0 * DOWN - p.pad_y0) / UP;

    for (int64_t plane = blockIdx.z; plane < p.planes; plane += gridDim.z) {
        const T *xp = x + plane * p.in_h * p.in_w;
        __syncthreads();
        // one warp per tile row: coalesced, no per-element division, row validity hoisted
        for (int ry = tid >> 5; ry < G::IN_H; ry += 8) {
            const int iy = iy0 + ry;
            const bool row_ok = iy >= 0 && iy < p.in_h;
            const T *xr = xp + (int64_t)iy * p.in_w;
            float *sr = s_in + ry * G::IN_WP;
#pragma unroll 2
            for (int rx = tid & 31; rx < G::IN_W; rx += 32) {
                const int ix = ix0 + rx;
                sr[rx] = (row_ok && ix >= 0 && ix < p.in_w) ? Cvt<T>::to_f(xr[ix]) : 0.f;
            }
        }
        __syncthreads();

        const int tx = tid & 31, ty = tid >> 5;
        const int wx0 = G::in_off(tx * TX) , wy0 = G::in_off(ty * TY);   // thread window origin
        float win[G::WIN_Y][G::NV * G::VEC];
#pragma unroll
        for (int r = 0; r < G::WIN_Y; ++r)
#pragma unroll
            for (int c = 0; c < G::NV; ++c) {
                const float *src = &s_in[(wy0 + r) * G::IN_WP + wx0 + c * G::VEC];
                if (G::VEC == 4) {
                    const float4 q = *reinterpret_cast<const float4 *>(src);
                    win[r][c * 4 + 0] = q.x; win[r][c * 4 + 1] = q.y;
                    win[r][c * 4 + 2] = q.z; win[r][c * 4 + 3] = q.w;
                } else {
                    const float2 q = *reinterpret_cast<const float2 *>(src);
                    win[r][c * 2 + 0] = q.x; win[r][c * 2 + 1] = q.y;
                }
            }
        float kreg[K][K];
#pragma unroll
        for (int r = 0; r < K; ++r)
#pragma unroll
            for (int c = 0; c < K; ++c) kreg[r][c] = s_k[r * K + c];

        T *op = out + plane * p.out_h * p.out_w;
#pragma unroll
        for (int j = 0; j < TY; ++j) {
            const int oy = oy0 + ty * TY + j;
            float row[TX];
#pragma unroll
            for (int i = 0; i < TX; ++i) {
                float acc = 0.f;
#pragma unroll
                for (int a = 0; a < G::ntap(j); ++a)
#pragma unroll
                    for (int b = 0; b < G::ntap(i); ++b)
                        acc += win[G::in_off(j) + a][G::in_off(i) + b] *
                               kreg[G::k0(j) + a * UP][G::k0(i) + b * UP];
                row[i] = acc;
            }
            if (oy < 0 || oy >= p.out_h) continue;
            const int oxs = ox0 + tx * TX;
            T *dst = op + (int64_t)oy * p.out_w + oxs;
            if (p.vec_store && oxs >= 0 && oxs + TX <= p.out_w) {       // TX contiguous outputs in one store
                T pk[TX];
#pragma unroll
                for (int i = 0; i < TX; ++i) pk[i] = Cvt<T>::from_f(row[i]);
                if (sizeof(T) * TX == 16) *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(pk);
                else if (sizeof(T) * TX == 8) *reinterpret_cast<uint2 *>(dst) = *reinterpret_cast<const uint2 *>(pk);
                else *reinterpret_cast<uint32_t *>(dst) = *reinterpret_cast<const uint32_t *>(pk);
            } else {
#pragma unroll
                for (int i = 0; i < TX; ++i)
                    if (oxs + i >= 0 && oxs + i < p.out_w) dst[i] = Cvt<T>::from_f(row[i]);
            }
        }
    }
}

// ---- generic kernel (runtime up/down/taps, minor >= 1) ----------------------------------------------

template <typename T>
__global__ void __launch_bounds__(256)
upfirdn2d_generic_kernel(T *__restrict__ out, const T *__restrict__ x,
                         const float *__restrict__ taps, UfdGenParams p) {
    __shared__ float s_in[GEN_IN_H * GEN_IN_W];
    __shared__ float s_k[GEN_MAXK * GEN_MAXK];
    const int tid = threadIdx.x;
    for (int i = tid; i < p.kh * p.kw; i += 256) {
        const int ky = i / p.kw, kx = i - ky * p.kw;
        s_k[i] = taps[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];
    }
    const int ox0 = blockIdx.x * GEN_TILE_W, oy0 = blockIdx.y * GEN_TILE_H;
    // input footprint of the tile (upfirdn2d_kernel.cu:85-88 index math)
    const int mid_x0 = ox0 * p.down_x + p.up_x - 1 - p.pad_x0;
    const int mid_y0 = oy0 * p.down_y + p.up_y - 1 - p.pad_y0;
    const int ix0 = floordiv(mid_x0, p.up_x), iy0 = floordiv(mid_y0, p.up_y);
    const int in_w_t = ((GEN_TILE_W - 1) * p.down_x + p.kw - 1) / p.up_x + 1;
    const int in_h_t = ((GEN_TILE_H - 1) * p.down_y + p.kh - 1) / p.up_y + 1;
    const int tx = tid & 31, ty = tid >> 5;
    const int ox = ox0 + tx, oy = oy0 + ty;
    const int mid_x = mid_x0 + tx * p.down_x, mid_y = mid_y0 + ty * p.down_y;
    const int in_x = floordiv(mid_x, p.up_x), in_y = floordiv(mid_y, p.up_y);
    const int kx0 = (in_x + 1) * p.up_x - mid_x - 1, ky0 = (in_y + 1) * p.up_y - mid_y - 1;
    const int64_t slices = p.major * p.minor;
    for (int64_t s = blockIdx.z; s < slices; s += gridDim.z) {
        const int64_t mj = s / p.minor;
        const int mn = (int)(s - mj * p.minor);
        __syncthreads();
        for (int i = tid; i < in_h_t * in_w_t; i += 256) {
            const int ry = i / in_w_t, rx = i - ry * in_w_t;
            const int iy = iy0 + ry, ix = ix0 + rx;
            float v = 0.f;
            if (iy >= 0 && ix >= 0 && iy < p.in_h && ix < p.in_w)
                v = Cvt<T>::to_f(x[((mj * p.in_h + iy) * p.in_w + ix) * p.minor + mn]);
            s_in[ry * GEN_IN_W + rx] = v;
        }
        __syncthreads();
        if (ox < p.out_w && oy < p.out_h) {
            float acc = 0.f;
            for (int a = 0, ky = ky0; ky < p.kh; ++a, ky += p.up_y)
                for (int b = 0, kx = kx0; kx < p.kw; ++b, kx += p.up_x)
                    acc += s_in[(in_y - iy0 + a) * GEN_IN_W + (in_x - ix0 + b)] * s_k[ky * p.kw + kx];
            out[((mj * p.out_h + oy) * p.out_w + ox) * p.minor + mn] = Cvt<T>::from_f(acc);
        }
    }
}

// ---- small images (<= 16 x 16 outputs): one thread per output, planes packed into the grid ------------
// At 4x4 ... 16x16 a 128 x 32 tile per block is >= 94 % idle; everything is cache resident, so plain
// gather loads with the same polyphase index math (upfirdn2d_kernel.cu:112-129) win.
template <typename T>
__global__ void __launch_bounds__(256)
upfirdn2d_small_kernel(T *__restrict__ out, const T *__restrict__ x, const float *__restrict__ taps,
                       UfdGenParams p) {
    __shared__ float s_k[GEN_MAXK * GEN_MAXK];
    for (int i = threadIdx.x; i < p.kh * p.kw; i += 256) {
        const int ky = i / p.kw, kx = i - ky * p.kw;
        s_k[i] = taps[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];
    }
    __syncthreads();
    const int opix = p.out_h * p.out_w;
    const int64_t total = p.major * opix;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t plane = i / opix;
        const int r = (int)(i - plane * opix);
        const int oy = r / p.out_w, ox = r - oy * p.out_w;
        const int mid_x = ox * p.down_x + p.up_x - 1 - p.pad_x0, mid_y = oy * p.down_y + p.up_y - 1 - p.pad_y0;
        const int in_x = floordiv(mid_x, p.up_x), in_y = floordiv(mid_y, p.up_y);
        const int kx0 = (in_x + 1) * p.up_x - mid_x - 1, ky0 = (in_y + 1) * p.up_y - mid_y - 1;
        const T *xp = x + plane * p.in_h * p.in_w;
        float acc = 0.f;
        for (int a = 0, ky = ky0; ky < p.kh; ++a, ky += p.up_y) {
            const int iy = in_y + a;
            if (iy < 0 || iy >= p.in_h) continue;
            for (int b = 0, kx = kx0; kx < p.kw; ++b, kx += p.up_x) {
                const int ix = in_x + b;
                if (ix < 0 || ix >= p.in_w) continue;
                acc += Cvt<T>::to_f(xp[iy * p.in_w + ix]) * s_k[ky * p.kw + kx];
            }
        }
        out[i] = Cvt<T>::from_f(acc);
    }
}

template <typename T, int UP, int DOWN, int TX, int TY>
static int launch_poly(void *out, const void *x, const float *taps, int64_t planes, int in_h,
                       int in_w, int out_h, int out_w, int kh, int kw, int pad_x0, int pad_y0,
                       cudaStream_t st) {
    using G = Geo<UP, DOWN, 4, TX, TY>;
    UfdParams p;
    p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w;
    p.pad_x0 = pad_x0; p.pad_y0 = pad_y0; p.kh = kh; p.kw = kw; p.planes = planes;
    // shift s in [0, UP) with (-s*DOWN - pad) % UP == 0  (DOWN == 1 whenever UP > 1 here)
    auto shift = [](int pad) { return UP == 1 ? 0 : ((-pad) % UP + UP) % UP; };
    p.shift_x = shift(pad_x0); p.shift_y = shift(pad_y0);
    p.vec_store = p.shift_x == 0 && out_w % TX == 0 && (reinterpret_cast<uintptr_t>(out) % (sizeof(T) * TX)) == 0 &&
                  ((int64_t)out_h * out_w * sizeof(T)) % (sizeof(T) * TX) == 0;
    dim3 grid((out_w + p.shift_x + G::TILE_W - 1) / G::TILE_W,
              (out_h + p.shift_y + G::TILE_H - 1) / G::TILE_H,
              (unsigned)std::min<int64_t>(planes, 32768));
    upfirdn2d_poly_kernel<T, UP, DOWN, 4, TX, TY><<<grid, 256, 0, st>>>((T *)out, (const T *)x, taps, p);
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

template <typename T>
int launch_upfirdn2d_stream(void *out, const void *x, const float *taps, int64_t planes, int in_h, int in_w, int out_h,
                            int out_w, int kh, int kw, int up, int down, int pad_x0, int pad_y0, cudaStream_t st);
template <typename T>
int launch_upfirdn2d_pk(void *out, const void *x, const float *taps, int64_t planes, int in_h, int in_w, int out_h, int out_w,
                        int kh, int kw, int up, int down, int pad_x0, int pad_y0, cudaStream_t st);
template <typename T>
int launch_upfirdn2d_planes(void *out, const void *x, const float *taps, int64_t planes, int in_h, int in_w, int out_h,
                            int out_w, int kh, int kw, int up, int down, int pad_x0, int pad_y0, cudaStream_t st);

}  // namespace sg2

using namespace sg2;

extern "C" int sg2_upfirdn2d(void *out, const void *x, const float *kernel, int64_t major, int in_h,
                             int in_w, int minor, int kh, int kw, int up_x, int up_y, int down_x,
                             int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1, int dtype,
                             sg2_stream_t stream) {
    SG2_REQUIRE(major >= 0 && in_h >= 1 && in_w >= 1 && minor >= 1, SG2_ERR_BAD_ARG,
                "upfirdn2d: bad input shape");
    SG2_REQUIRE(kh >= 1 && kw >= 1 && up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1,
                SG2_ERR_BAD_ARG, "upfirdn2d: kernel size, up and down must be >= 1");
    SG2_REQUIRE(kh <= GEN_MAXK && kw <= GEN_MAXK && up_x <= GEN_MAXF && up_y <= GEN_MAXF &&
                    down_x <= GEN_MAXF && down_y <= GEN_MAXF,
                SG2_ERR_UNSUPPORTED,
                "upfirdn2d: unsupported configuration (up=%d,%d down=%d,%d taps=%dx%d); supported: "
                "up, down <= %d and taps <= %dx%d",
                up_x, up_y, down_x, down_y, kh, kw, GEN_MAXF, GEN_MAXK, GEN_MAXK);
    const int out_h = (in_h * up_y + pad_y0 + pad_y1 - kh) / down_y + 1;
    const int out_w = (in_w * up_x + pad_x0 + pad_x1 - kw) / down_x + 1;
    SG2_REQUIRE(in_h * up_y + pad_y0 + pad_y1 >= kh && in_w * up_x + pad_x0 + pad_x1 >= kw &&
                    out_h >= 1 && out_w >= 1,
                SG2_ERR_BAD_ARG, "upfirdn2d: empty output (%d x %d)", out_h, out_w);
    if (major == 0) return SG2_OK;
    SG2_REQUIRE(out && x && kernel, SG2_ERR_BAD_ARG, "upfirdn2d: null tensor pointer");
    cudaStream_t st = as_stream(stream);
    const bool sym = up_x == up_y && down_x == down_y && minor == 1 && kh <= 4 && kw <= 4;
    SG2_DISPATCH_DTYPE(dtype, {
        if (sym && sizeof(T) == 2) {   // 2-byte storage, blur / down-2 (planes from 16 outputs wide when there are thousands): packed-math kernel (upfirdn2d_pk.cu); 1 = not applicable
            const char *env_pk = getenv("SG2_UPFIRDN_PK");                // A/B switch (read per call): 0 = off
            if (!env_pk || atoi(env_pk) != 0) {
                const int rc = launch_upfirdn2d_pk<T>(out, x, kernel, major, in_h, in_w, out_h, out_w, kh, kw, up_x, down_x,
                                                      pad_x0, pad_y0, st);
                if (rc <= 0) return rc;
            }
        }
        if (sym) {   // many small planes (the 4^2 .. 32^2 octaves): batches of whole planes per CTA (upfirdn2d_planes.cu); 1 = not applicable
            const char *env_pl = getenv("SG2_UPFIRDN_PLANES");            // A/B switch (read per call): 0 = off
            if (!env_pl || atoi(env_pl) != 0) {
                const int rc = launch_upfirdn2d_planes<T>(out, x, kernel, major, in_h, in_w, out_h, out_w, kh, kw, up_x, down_x,
                                                          pad_x0, pad_y0, st);
                if (rc <= 0) return rc;
            }
        }
        if (minor == 1 && out_h * out_w <= 256 && !(sym && out_w > 8 && out_h > 8)) {   // tiny planes (<= 8 x 8 ... 16 x 16 non-model geometries)
            UfdGenParams p;
            p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w; p.minor = 1;
            p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y;
            p.pad_x0 = pad_x0; p.pad_y0 = pad_y0; p.kh = kh; p.kw = kw; p.major = major;
            const int64_t total = major * out_h * out_w;
            unsigned blocks = (unsigned)std::min<int64_t>(ceil_div64(total, 256), (int64_t)sm_count() * 32);
            upfirdn2d_small_kernel<T><<<blocks, 256, 0, st>>>((T *)out, (const T *)x, kernel, p);
            SG2_LAUNCH_CHECK();
            return SG2_OK;
        }
        if (sym) {   // the three model geometries: row-streaming kernel (upfirdn2d_stream.cu); 1 = not applicable
            static const char *env_tiled = getenv("SG2_UPFIRDN_TILED");     // A/B switch: force the tiled kernels
            if (!env_tiled || atoi(env_tiled) == 0) {
                const int rc = launch_upfirdn2d_stream<T>(out, x, kernel, major, in_h, in_w, out_h, out_w, kh, kw, up_x, down_x,
                                                          pad_x0, pad_y0, st);
                if (rc <= 0) return rc;
            }
        }
        if (sym && up_x == 1 && down_x == 1)
            return launch_poly<T, 1, 1, 4, 4>(out, x, kernel, major, in_h, in_w, out_h, out_w, kh, kw, pad_x0, pad_y0, st);
        if (sym && up_x == 2 && down_x == 1)
            return launch_poly<T, 2, 1, 4, 4>(out, x, kernel, major, in_h, in_w, out_h, out_w, kh, kw, pad_x0, pad_y0, st);
        if (sym && up_x == 1 && down_x == 2)
            return launch_poly<T, 1, 2, 2, 2>(out, x, kernel, major, in_h, in_w, out_h, out_w, kh, kw, pad_x0, pad_y0, st);
        UfdGenParams p;
        p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w; p.minor = minor;
        p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y;
        p.pad_x0 = pad_x0; p.pad_y0 = pad_y0; p.kh = kh; p.kw = kw; p.major = major;
        dim3 grid((out_w + GEN_TILE_W - 1) / GEN_TILE_W, (out_h + GEN_TILE_H - 1) / GEN_TILE_H,
                  (unsigned)std::min<int64_t>(major * minor, 32768));
        upfirdn2d_generic_kernel<T><<<grid, 256, 0, st>>>((T *)out, (const T *)x, kernel, p);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}
