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
