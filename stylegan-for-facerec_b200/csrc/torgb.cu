// torgb.cu -- ToRGB tail (model.py:350-359): out = conv1x1 + bias[c] + upfirdn2d(skip, k, up=2, pad)
// fused into one pass over the 3-channel image instead of three (bias add, upsample, add).
#include "common.cuh"

namespace sg2 {

__host__ __device__ __forceinline__ int fdiv(int a, int b) { int q = a / b; return (q * b > a) ? q - 1 : q; }

template <typename T>
__global__ void __launch_bounds__(256)
torgb_combine_kernel(T *__restrict__ out, const T *__restrict__ conv, const T *__restrict__ bias,
                     const T *__restrict__ skip, const float *__restrict__ taps, int kh, int kw,
                     int pad0, int64_t planes, int C, int H, int W) {
    __shared__ float s_k[16 * 16];
    for (int i = threadIdx.x; i < kh * kw; i += 256) {
        const int ky = i / kw, kx = i - ky * kw;
        s_k[i] = taps[(kh - 1 - ky) * kw + (kw - 1 - kx)];      // flipped: true convolution
    }
    __syncthreads();
    const int SH = H / 2, SW = W / 2;
    const int64_t total = planes * H * W;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t plane = i / ((int64_t)H * W);
        const int r = (int)(i - plane * H * W);
        const int oy = r / W, ox = r - oy * W;
        float v = Cvt<T>::to_f(conv[i]);
        if (bias) v += Cvt<T>::to_f(__ldg(bias + (int)(plane % C)));
        if (skip) {
            // polyphase up-by-2 (same index math as upfirdn2d_kernel.cu:112-129 with up=2, down=1)
            const int mid_y = oy + 1 - pad0, mid_x = ox + 1 - pad0;
            const int iy0 = fdiv(mid_y, 2), ix0 = fdiv(mid_x, 2);
            const int ky0 = (iy0 + 1) * 2 - mid_y - 1, kx0 = (ix0 + 1) * 2 - mid_x - 1;
            const T *sp = skip + plane * SH * SW;
            float acc = 0.f;
            for (int a = 0, ky = ky0; ky < kh; ++a, ky += 2) {
                const int iy = iy0 + a;
                if (iy < 0 || iy >= SH) continue;
                for (int b = 0, kx = kx0; kx < kw; ++b, kx += 2) {
                    const int ix = ix0 + b;
                    if (ix < 0 || ix >= SW) continue;
                    acc += Cvt<T>::to_f(sp[iy * SW + ix]) * s_k[ky * kw + kx];
                }
            }
            v = v + acc;
        }
        out[i] = Cvt<T>::from_f(v);
    }
}

}  // namespace sg2

using namespace sg2;

extern "C" int sg2_torgb_combine(void *out, const void *conv, const void *bias, const void *skip,
                                 const float *kernel, int kh, int kw, int pad0, int pad1, int64_t B,
                                 int C, int H, int W, int dtype, sg2_stream_t stream) {
    SG2_REQUIRE(B >= 0 && C >= 1 && H >= 1 && W >= 1, SG2_ERR_BAD_ARG, "torgb_combine: bad shape");
    if (B == 0) return SG2_OK;
    SG2_REQUIRE(out && conv, SG2_ERR_BAD_ARG, "torgb_combine: null tensor pointer");
    if (skip) {
        SG2_REQUIRE(kernel && kh >= 1 && kw >= 1 && kh <= 16 && kw <= 16, SG2_ERR_UNSUPPORTED,
                    "torgb_combine: taps must be 1..16 per axis");
        SG2_REQUIRE(H % 2 == 0 && W % 2 == 0 && (H / 2) * 2 + pad0 + pad1 - kh + 1 == H &&
                        (W / 2) * 2 + pad0 + pad1 - kw + 1 == W,
                    SG2_ERR_BAD_ARG, "torgb_combine: skip upsample does not produce %dx%d", H, W);
    }
    const int64_t total = B * C * (int64_t)H * W;
    unsigned blocks = (unsigned)std::min<int64_t>(ceil_div64(total, 256), (int64_t)sm_count() * 16);
    SG2_DISPATCH_DTYPE(dtype, {
        torgb_combine_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>(
            (T *)out, (const T *)conv, (const T *)bias, (const T *)skip, kernel, kh, kw, pad0, B * C, C, H, W);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// The modulated 1x1 convolution of ToRGB (model.py:350-355: ModulatedConv2d(in, 3, 1, demodulate=False)) and its adjoint
// for the differentiable path -- an HBM-bound pass over the activation each way (fp32 math, exact):
//   forward : y[b,k,p]  = sum_c w[k,c] * s[b,c] * x[b,c,p]                                (k < K <= 4)
//   backward: gx[b,c,p] = s[b,c] * t,  gs[b,c] = sum_p x[b,c,p] * t,  t = sum_k w[k,c] * gy[b,k,p]
namespace sg2 {

constexpr int RGB_MAXK = 4, RGB_MAXC = 1024;

template <typename T>
__global__ void __launch_bounds__(256)
rgb_mod_fwd_kernel(T *__restrict__ y, const T *__restrict__ x, const float *__restrict__ w, const float *__restrict__ s,
                   int C, int K, long long HW) {
    __shared__ float ws[RGB_MAXK][RGB_MAXC];
    const long long b = blockIdx.y;
    for (int i = threadIdx.x; i < K * C; i += 256) {
        const int k = i / C, c = i - k * C;
        ws[k][c] = w[i] * s[b * C + c];
    }
    __syncthreads();
    const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
    if (p >= HW) return;
    const T *xp = x + b * C * HW + p;
    float acc[RGB_MAXK] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int c = 0; c < C; ++c) {
        const float v = Cvt<T>::to_f(xp[(long long)c * HW]);
#pragma unroll
        for (int k = 0; k < RGB_MAXK; ++k)
            if (k < K) acc[k] = fmaf(ws[k][c], v, acc[k]);
    }
#pragma unroll
    for (int k = 0; k < RGB_MAXK; ++k)
        if (k < K) y[(b * K + k) * HW + p] = Cvt<T>::from_f(acc[k]);
}

constexpr int RGB_CT = 8;      // channels per block of the adjoint

template <typename T>
__global__ void __launch_bounds__(256)
rgb_mod_bwd_kernel(T *__restrict__ gx, float *__restrict__ gs, const T *__restrict__ gy, const T *__restrict__ x,
                   const float *__restrict__ w, const float *__restrict__ s, int C, int K, long long HW, long long chunk) {
    __shared__ float part[8][RGB_CT];
    const long long b = blockIdx.z;
    const int c0 = blockIdx.y * RGB_CT;
    float wk[RGB_MAXK][RGB_CT], sc[RGB_CT], acc[RGB_CT];
#pragma unroll
    for (int j = 0; j < RGB_CT; ++j) {
        const int c = min(c0 + j, C - 1);
        sc[j] = s[b * C + c];
        acc[j] = 0.f;
#pragma unroll
        for (int k = 0; k < RGB_MAXK; ++k) wk[k][j] = k < K ? w[k * C + c] : 0.f;
    }
    const long long p_lo = (long long)blockIdx.x * chunk, p_hi = min(p_lo + chunk, HW);
    for (long long p = p_lo + threadIdx.x; p < p_hi; p += 256) {
        float g[RGB_MAXK];
#pragma unroll
        for (int k = 0; k < RGB_MAXK; ++k) g[k] = k < K ? Cvt<T>::to_f(gy[(b * K + k) * HW + p]) : 0.f;
#pragma unroll
        for (int j = 0; j < RGB_CT; ++j) {
            if (c0 + j >= C) break;
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < RGB_MAXK; ++k) t = fmaf(wk[k][j], g[k], t);
            const long long o = (b * C + c0 + j) * HW + p;
            acc[j] = fmaf(Cvt<T>::to_f(x[o]), t, acc[j]);
            gx[o] = Cvt<T>::from_f(sc[j] * t);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < RGB_CT; ++j) {
        float r = acc[j];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) r += __shfl_xor_sync(0xffffffffu, r, d);
        if (lane == 0) part[warp][j] = r;
    }
    __syncthreads();
    if (threadIdx.x < RGB_CT && c0 + threadIdx.x < C) {
        float r = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) r += part[wv][threadIdx.x];
        atomicAdd(gs + b * C + c0 + threadIdx.x, r);
    }
}

}  // namespace sg2

static int check_rgb_mod(const char *what, int64_t B, int C, int K, int64_t HW) {
    SG2_REQUIRE(B >= 0 && B <= 65535 && C >= 1 && C <= RGB_MAXC && K >= 1 && K <= RGB_MAXK && HW >= 1, SG2_ERR_UNSUPPORTED,
                "%s: bad shape (B %lld, C %d <= %d, K %d <= %d, HW %lld)", what, (long long)B, C, RGB_MAXC, K, RGB_MAXK, (long long)HW);
    return SG2_OK;
}

// y [B,K,HW] = sum_c w[k,c] * s[b,c] * x[b,c,p]; w [K,C] fp32 (conv scale folded), s [B,C] fp32; x / y of `dtype`
extern "C" int sg2_rgb_modconv_fwd(void *y, const void *x, const float *w, const float *s, int64_t B, int C, int K, int64_t HW,
                                   int dtype, sg2_stream_t stream) {
    int rc = check_rgb_mod("rgb_modconv_fwd", B, C, K, HW);
    if (rc || B == 0) return rc;
    SG2_REQUIRE(y && x && w && s, SG2_ERR_BAD_ARG, "rgb_modconv_fwd: null pointer");
    dim3 grid((unsigned)((HW + 255) / 256), (unsigned)B);
    SG2_DISPATCH_DTYPE(dtype, {
        rgb_mod_fwd_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((T *)y, (const T *)x, w, s, C, K, (long long)HW);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}

// adjoint: gx [B,C,HW] (`dtype`) and gs [B,C] fp32 (accumulated with atomics: zero it first) from gy [B,K,HW] and x
extern "C" int sg2_rgb_modconv_bwd(void *gx, float *gs, const void *gy, const void *x, const float *w, const float *s, int64_t B,
                                   int C, int K, int64_t HW, int dtype, sg2_stream_t stream) {
    int rc = check_rgb_mod("rgb_modconv_bwd", B, C, K, HW);
    if (rc || B == 0) return rc;
    SG2_REQUIRE(gx && gs && gy && x && w && s, SG2_ERR_BAD_ARG, "rgb_modconv_bwd: null pointer");
    const long long chunk = 4096;
    dim3 grid((unsigned)((HW + chunk - 1) / chunk), (unsigned)((C + RGB_CT - 1) / RGB_CT), (unsigned)B);
    SG2_DISPATCH_DTYPE(dtype, {
        rgb_mod_bwd_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((T *)gx, gs, (const T *)gy, (const T *)x, w, s, C, K,
                                                                    (long long)HW, chunk);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}
