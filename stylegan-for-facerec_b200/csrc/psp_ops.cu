// psp_ops.cu -- the step right after the decoder in ReStyle / pSp:
//   * face_pool = AdaptiveAvgPool2d((256, 256)) on the decoder output   (restyle-encoder/models/psp.py:33,113-114;
//     training/coach_restyle_psp.py:143-156; utils/inference_utils.py:36-38) -- for the integer ratios that
//     occur (1024 -> 256: 4x4 means; 256 -> 256: identity) a box filter;
//   * F.interpolate(y_hat, 112, mode='bilinear') before the identity / L2 losses (coach_restyle_psp.py:156),
//     align_corners=False, no antialiasing (the PyTorch defaults the reference relies on).
// NCHW, fp32 / fp16 / bf16 storage, fp32 math; a few MB per batch: launch-latency territory, one pass each.
#include <algorithm>

#include "common.cuh"

namespace sg2 {

template <typename T>
__global__ void __launch_bounds__(256)
avg_pool_int_kernel(T *__restrict__ out, const T *__restrict__ x, long long total, int OH, int OW, int f) {
    const float inv = 1.f / (float)(f * f);
    const int IW = OW * f;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const int ox = (int)(idx % OW);
        const long long t = idx / OW;
        const int oy = (int)(t % OH);
        const long long plane = t / OH;
        const T *p = x + (plane * OH * f + (long long)oy * f) * IW + (long long)ox * f;
        float acc = 0.f;
        for (int a = 0; a < f; ++a)
            for (int b = 0; b < f; ++b) acc += Cvt<T>::to_f(p[(long long)a * IW + b]);
        out[idx] = Cvt<T>::from_f(acc * inv);
    }
}

// same index math as ATen's area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=false)
__device__ __forceinline__ void bilinear_src(int dst, float scale, int in_size, int &i0, int &i1, float &l0, float &l1) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = (int)src;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = src - (float)i0;
    l0 = 1.f - l1;
}

template <typename T>
__global__ void __launch_bounds__(256)
resize_bilinear_kernel(T *__restrict__ out, const T *__restrict__ x, long long total, int IH, int IW, int OH, int OW,
                       float sh, float sw) {
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const int ox = (int)(idx % OW);
        const long long t = idx / OW;
        const int oy = (int)(t % OH);
        const long long plane = t / OH;
        int y0, y1, x0, x1;
        float ly0, ly1, lx0, lx1;
        bilinear_src(oy, sh, IH, y0, y1, ly0, ly1);
        bilinear_src(ox, sw, IW, x0, x1, lx0, lx1);
        const T *p = x + plane * IH * IW;
        const float v00 = Cvt<T>::to_f(p[(long long)y0 * IW + x0]), v01 = Cvt<T>::to_f(p[(long long)y0 * IW + x1]);
        const float v10 = Cvt<T>::to_f(p[(long long)y1 * IW + x0]), v11 = Cvt<T>::to_f(p[(long long)y1 * IW + x1]);
        out[idx] = Cvt<T>::from_f(ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11));
    }
}

// ---- adjoints (the ReStyle coaches back-propagate the image losses through both steps) ------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
avg_pool_int_bwd_kernel(T *__restrict__ gx, const T *__restrict__ gy, long long total, int OH, int OW, int f) {
    const float inv = 1.f / (float)(f * f);
    const int IW = OW * f, IH = OH * f;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const int ix = (int)(idx % IW);
        const long long t = idx / IW;
        const int iy = (int)(t % IH);
        const long long plane = t / IH;
        gx[idx] = Cvt<T>::from_f(Cvt<T>::to_f(gy[(plane * OH + iy / f) * OW + ix / f]) * inv);
    }
}

// weight with which output index `dst` reads input index `i` (0 when it does not)
__device__ __forceinline__ float bilinear_weight(int dst, float scale, int in_size, int i) {
    int i0, i1;
    float l0, l1;
    bilinear_src(dst, scale, in_size, i0, i1, l0, l1);
    return (i0 == i ? l0 : 0.f) + (i1 == i ? l1 : 0.f);
}

// output indices that can read input index i: src(dst) in (i - 1, i + 1), one extra on each side for rounding
__device__ __forceinline__ void bilinear_dst_range(int i, float scale, int out_size, int &lo, int &hi) {
    const float inv = 1.f / scale;
    lo = i == 0 ? 0 : max(0, (int)floorf(((float)i - 0.5f) * inv - 0.5f) - 1);
    hi = min(out_size - 1, (int)ceilf(((float)i + 1.5f) * inv - 0.5f) + 1);
}

// gather form (one thread per input cell, no atomics: deterministic)
template <typename T>
__global__ void __launch_bounds__(256)
resize_bilinear_bwd_kernel(T *__restrict__ gx, const T *__restrict__ gy, long long total, int IH, int IW, int OH, int OW,
                           float sh, float sw) {
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const int ix = (int)(idx % IW);
        const long long t = idx / IW;
        const int iy = (int)(t % IH);
        const long long plane = t / IH;
        int ylo, yhi, xlo, xhi;
        bilinear_dst_range(iy, sh, OH, ylo, yhi);
        bilinear_dst_range(ix, sw, OW, xlo, xhi);
        const T *g = gy + plane * OH * OW;
        float acc = 0.f;
        for (int oy = ylo; oy <= yhi; ++oy) {
            const float wy = bilinear_weight(oy, sh, IH, iy);
            if (wy == 0.f) continue;
            float row = 0.f;
            for (int ox = xlo; ox <= xhi; ++ox) {
                const float wx = bilinear_weight(ox, sw, IW, ix);
                if (wx != 0.f) row += wx * Cvt<T>::to_f(g[(long long)oy * OW + ox]);
            }
            acc += wy * row;
        }
        gx[idx] = Cvt<T>::from_f(acc);
    }
}

// images in [-1, 1] -> uint8 (tensor2im of the reference's inference scripts, restyle-encoder/utils/common.py:5-11: ((x + 1) / 2).clip(0, 1) * 255),
// done on the device so that only a quarter of the bytes crosses PCIe
__device__ __forceinline__ uint32_t to_u8(float x) {
    float v = (x + 1.f) * 0.5f;
    v = fminf(fmaxf(v, 0.f), 1.f) * 255.f;
    return (uint32_t)v;                                       // truncation, like numpy's astype('uint8')
}

// 4 elements per thread: one 16-byte (fp32) / 8-byte (fp16, bf16) load, one 4-byte store
template <typename T>
__global__ void __launch_bounds__(256)
image_to_uint8_kernel(uint8_t *__restrict__ out, const T *__restrict__ x, long long total, int vec_ok) {
    const long long n4 = vec_ok ? total / 4 : 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        float v[4];
        if constexpr (sizeof(T) == 4) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(x) + i);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
            const uint2 q = __ldg(reinterpret_cast<const uint2 *>(x) + i);
            const T *h = reinterpret_cast<const T *>(&q);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = Cvt<T>::to_f(h[e]);
        }
        reinterpret_cast<uint32_t *>(out)[i] = to_u8(v[0]) | (to_u8(v[1]) << 8) | (to_u8(v[2]) << 16) | (to_u8(v[3]) << 24);
    }
    // tail (or everything, when the pointers are not aligned for the vector path)
    for (long long i = n4 * 4 + (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256)
        out[i] = (uint8_t)to_u8(Cvt<T>::to_f(x[i]));
}

}  // namespace sg2

using namespace sg2;

extern "C" int sg2_image_to_uint8(void *out, const void *x, int64_t total, int dtype, sg2_stream_t stream) {
    SG2_REQUIRE(total >= 0, SG2_ERR_BAD_ARG, "image_to_uint8: bad size");
    if (total == 0) return SG2_OK;
    SG2_REQUIRE(out && x, SG2_ERR_BAD_ARG, "image_to_uint8: null pointer");
    const int vec_ok = reinterpret_cast<uintptr_t>(out) % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0;
    const unsigned blocks = (unsigned)std::min<long long>((total / 4 + 255) / 256 + 1, (long long)sm_count() * 64);
    SG2_DISPATCH_DTYPE(dtype, {
        image_to_uint8_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>((uint8_t *)out, (const T *)x, (long long)total, vec_ok);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}

extern "C" int sg2_avg_pool_int(void *out, const void *x, int64_t planes, int out_h, int out_w, int factor, int dtype,
                                sg2_stream_t stream) {
    SG2_REQUIRE(planes >= 0 && out_h >= 1 && out_w >= 1 && factor >= 1 && factor <= 64, SG2_ERR_BAD_ARG, "avg_pool_int: bad shape");
    if (planes == 0) return SG2_OK;
    SG2_REQUIRE(out && x, SG2_ERR_BAD_ARG, "avg_pool_int: null pointer");
    const long long total = (long long)planes * out_h * out_w;
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, (long long)sm_count() * 32);
    SG2_DISPATCH_DTYPE(dtype, {
        avg_pool_int_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>((T *)out, (const T *)x, total, out_h, out_w, factor);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}

extern "C" int sg2_resize_bilinear(void *out, const void *x, int64_t planes, int in_h, int in_w, int out_h, int out_w, int dtype,
                                   sg2_stream_t stream) {
    SG2_REQUIRE(planes >= 0 && in_h >= 1 && in_w >= 1 && out_h >= 1 && out_w >= 1, SG2_ERR_BAD_ARG, "resize_bilinear: bad shape");
    if (planes == 0) return SG2_OK;
    SG2_REQUIRE(out && x, SG2_ERR_BAD_ARG, "resize_bilinear: null pointer");
    const long long total = (long long)planes * out_h * out_w;
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, (long long)sm_count() * 32);
    const float sh = (float)in_h / (float)out_h, sw = (float)in_w / (float)out_w;
    SG2_DISPATCH_DTYPE(dtype, {
        resize_bilinear_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>((T *)out, (const T *)x, total, in_h, in_w, out_h, out_w, sh, sw);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}

// adjoints of the two ops above: grad_x from grad_out (same shapes / factor as the forward call)
extern "C" int sg2_avg_pool_int_bwd(void *grad_x, const void *grad_out, int64_t planes, int out_h, int out_w, int factor,
                                    int dtype, sg2_stream_t stream) {
    SG2_REQUIRE(planes >= 0 && out_h >= 1 && out_w >= 1 && factor >= 1 && factor <= 64, SG2_ERR_BAD_ARG, "avg_pool_int_bwd: bad shape");
    if (planes == 0) return SG2_OK;
    SG2_REQUIRE(grad_x && grad_out, SG2_ERR_BAD_ARG, "avg_pool_int_bwd: null pointer");
    const long long total = (long long)planes * out_h * out_w * factor * factor;
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, (long long)sm_count() * 32);
    SG2_DISPATCH_DTYPE(dtype, {
        avg_pool_int_bwd_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>((T *)grad_x, (const T *)grad_out, total, out_h, out_w, factor);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}

extern "C" int sg2_resize_bilinear_bwd(void *grad_x, const void *grad_out, int64_t planes, int in_h, int in_w, int out_h,
                                       int out_w, int dtype, sg2_stream_t stream) {
    SG2_REQUIRE(planes >= 0 && in_h >= 1 && in_w >= 1 && out_h >= 1 && out_w >= 1, SG2_ERR_BAD_ARG, "resize_bilinear_bwd: bad shape");
    if (planes == 0) return SG2_OK;
    SG2_REQUIRE(grad_x && grad_out, SG2_ERR_BAD_ARG, "resize_bilinear_bwd: null pointer");
    const long long total = (long long)planes * in_h * in_w;
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, (long long)sm_count() * 32);
    const float sh = (float)in_h / (float)out_h, sw = (float)in_w / (float)out_w;
    SG2_DISPATCH_DTYPE(dtype, {
        resize_bilinear_bwd_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>((T *)grad_x, (const T *)grad_out, total, in_h, in_w, out_h, out_w, sh, sw);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}
