// layout_ops.cu -- the two passes either side of a tensor-core convolution in the differentiable path
// (stylegan2/tc_route.py): NCHW (fp32 / fp16 / bf16) <-> NHWC bf16 with the per-(sample, channel) factor of the
// modulated convolution folded in, and -- for the backward direction -- the per-(sample, channel) reduction that the
// adjoint of that factor needs, in the same pass:
//     modulate  : xh[b,p,c] = bf16(x[b,c,p] * s[b,c])                     (ModulatedConv2d, model.py:236-237, factored)
//     its adjoint: gx[b,c,p] = s[b,c] * gh[b,p,c],  gs[b,c] = sum_p x[b,c,p] * gh[b,p,c]
//     demodulate: y[b,c,p]  = d[b,c] * yh[b,p,c]                          (model.py:239-240, factored)
//     its adjoint: gh[b,p,c] = bf16(d[b,c] * gy[b,c,p]), gd[b,c] = sum_p gy[b,c,p] * yh[b,p,c]
// For the layers without a blur between convolution and activation the second pass also applies the StyledConv tail
// (NoiseInjection + FusedLeakyReLU, model.py:282-287,331-337): y = lrelu(d * yh + w * noise + bias) * sqrt(2), and the
// first one its adjoint: g = gy * (y > 0 ? 1 : alpha) * sqrt(2) (fused_bias_act_kernel.cu:28-47, grad = 1) before
// everything else, with grad_bias reduced in the same pass.
// Both kernels move a 64-channel x 32-pixel tile through shared memory so that each side is read / written in 128-byte
// rows; HBM-bound, (4 + 2) bytes per element (+ 2 or 4 for the reduction operand).
#include <algorithm>

#include "common.cuh"

namespace sg2 {

constexpr int LT_C = 64, LT_P = 32;

// element offset of pixel p of sample b on the NHWC side.  poly_w == 0: plain [B][HW][C].  poly_w = W > 0: the image is
// W x W and stored as the four polyphase planes of the transposed convolution, [(py,px)][B][P][P][C] with P = (W+1)/2,
// pixel (y, x) living in plane (y&1, x&1) at (y>>1, x>>1) -- what sg2_conv_transpose3x3_tc writes and what the
// polyphase form of its input gradient reads, so no interleaving copy is needed on either side.
__device__ __forceinline__ long long nhwc_offset(long long b, long long p, int C, long long HW, int poly_w, long long B) {
    if (poly_w == 0) return (b * HW + p) * C;
    const int pi = (int)p;                                        // W <= 32767: 32-bit division
    const int y = pi / poly_w, x = pi - y * poly_w;
    const int P = (poly_w + 1) >> 1, s = (y & 1) * 2 + (x & 1);
    return ((((long long)s * B + b) * P + (y >> 1)) * P + (x >> 1)) * C;
}

// NCHW -> NHWC bf16 (x scale); optional red[b,c] += sum_p x[b,c,p] * other[b,p,c]
template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(__nv_bfloat16 *__restrict__ out, const T *__restrict__ x, const float *__restrict__ scale,
                    const __nv_bfloat16 *__restrict__ other, float *__restrict__ red, int C, long long HW,
                    const T *__restrict__ act_ref, float alpha, float gain, float *__restrict__ red_sum, int poly_w) {
    __shared__ float tile[LT_C][LT_P + 1];
    __shared__ float part[8][LT_C];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long p0 = (long long)blockIdx.x * LT_P;
    const int c0 = blockIdx.y * LT_C;
    const long long b = blockIdx.z;
#pragma unroll
    for (int j = 0; j < LT_C / 8; ++j) {
        const int c = c0 + warp + 8 * j;
        const long long p = p0 + lane;
        float v = 0.f;
        if (c < C && p < HW) {
            v = Cvt<T>::to_f(x[(b * C + c) * HW + p]);
            if (act_ref) v *= (Cvt<T>::to_f(act_ref[(b * C + c) * HW + p]) > 0.f ? 1.f : alpha) * gain;   // lrelu'(y) * gain
        }
        tile[warp + 8 * j][lane] = v;
    }
    __syncthreads();
    const int c = c0 + 2 * lane;                       // this lane's channel pair
    const bool cok = c < C;                            // C is even
    const float s0 = (cok && scale) ? __ldg(scale + b * C + c) : 1.f, s1 = (cok && scale) ? __ldg(scale + b * C + c + 1) : 1.f;
    float r0 = 0.f, r1 = 0.f, t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int j = 0; j < LT_P / 8; ++j) {
        const int pl = warp + 8 * j;
        const long long p = p0 + pl;
        if (cok && p < HW) {
            const float v0 = tile[2 * lane][pl], v1 = tile[2 * lane + 1][pl];
            const long long o = nhwc_offset(b, p, C, HW, poly_w, gridDim.z) + c;
            *reinterpret_cast<__nv_bfloat162 *>(out + o) = __floats2bfloat162_rn(v0 * s0, v1 * s1);
            t0 += v0;
            t1 += v1;
            if (other) {
                const float2 q = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(other + o));
                r0 = fmaf(v0, q.x, r0);
                r1 = fmaf(v1, q.y, r1);
            }
        }
    }
    if (other) {                                       // block-uniform
        part[warp][2 * lane] = r0;
        part[warp][2 * lane + 1] = r1;
        __syncthreads();
        if (threadIdx.x < LT_C && c0 + threadIdx.x < C) {
            float acc = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) acc += part[w][threadIdx.x];
            atomicAdd(red + b * C + c0 + threadIdx.x, acc);
        }
    }
    if (red_sum) {                                     // block-uniform: sum_p of the (activation-gradient) tile
        __syncthreads();
        part[warp][2 * lane] = t0;
        part[warp][2 * lane + 1] = t1;
        __syncthreads();
        if (threadIdx.x < LT_C && c0 + threadIdx.x < C) {
            float acc = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) acc += part[w][threadIdx.x];
            atomicAdd(red_sum + b * C + c0 + threadIdx.x, acc);
        }
    }
}

// NHWC bf16 -> NCHW (x scale); optional red[b,c] += sum_p other[b,c,p] * h[b,p,c]
template <typename T>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(T *__restrict__ out, const __nv_bfloat16 *__restrict__ h, const float *__restrict__ scale,
                    const T *__restrict__ other, float *__restrict__ red, int C, long long HW, int act,
                    const T *__restrict__ noise, long long noise_bstride, const T *__restrict__ noise_weight,
                    const T *__restrict__ bias, float alpha, float gain, int poly_w, int n_parts, int part_pitch) {
    __shared__ float tile[LT_C][LT_P + 1];
    __shared__ float csum[LT_C];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long p0 = (long long)blockIdx.x * LT_P;
    const int c0 = blockIdx.y * LT_C;
    const long long b = blockIdx.z;
    {
        const int c = c0 + 2 * lane;
#pragma unroll
        for (int j = 0; j < LT_P / 8; ++j) {
            const int pl = warp + 8 * j;
            const long long p = p0 + pl;
            float2 q = make_float2(0.f, 0.f);
            if (c < C && p < HW) {
                if (n_parts == 0) {
                    q = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(h + nhwc_offset(b, p, C, HW, poly_w, gridDim.z) + c));
                } else {
                    // sum of n_parts tensors [B][pitch][pitch][C], of which the top-left W x W corner (W = poly_w) is read:
                    // the polyphase components of the transposed conv's input gradient, accumulated in fp32
                    const int pi = (int)p;
                    const int y = pi / poly_w, x = pi - y * poly_w;
                    const long long base = ((b * part_pitch + y) * part_pitch + x) * C + c;
                    const long long stride = (long long)gridDim.z * part_pitch * part_pitch * C;
                    for (int k = 0; k < n_parts; ++k) {
                        const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(h + k * stride + base));
                        q.x += t.x;
                        q.y += t.y;
                    }
                }
            }
            tile[2 * lane][pl] = q.x;
            tile[2 * lane + 1][pl] = q.y;
        }
    }
    __syncthreads();
    float nz = 0.f;                                    // this lane's pixel: noise_weight * noise[b or 0, p]
    if (act && noise && p0 + lane < HW) nz = Cvt<T>::to_f(noise_weight[0]) * Cvt<T>::to_f(noise[b * noise_bstride + p0 + lane]);
#pragma unroll
    for (int j = 0; j < LT_C / 8; ++j) {
        const int c = c0 + warp + 8 * j;               // warp-uniform
        if (c >= C) continue;
        const long long p = p0 + lane;
        const float v = tile[warp + 8 * j][lane];
        const float s = scale ? __ldg(scale + b * C + c) : 1.f;
        float r = 0.f;
        if (p < HW) {
            const long long o = (b * C + c) * HW + p;
            float y = v * s;
            if (act) {
                y += nz + (bias ? Cvt<T>::to_f(bias[c]) : 0.f);
                y = (y > 0.f ? y : y * alpha) * gain;
            }
            out[o] = Cvt<T>::from_f(y);
            if (other) r = Cvt<T>::to_f(other[o]) * v;
        }
        if (other) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) r += __shfl_xor_sync(0xffffffffu, r, d);
            if (lane == 0) csum[warp + 8 * j] = r;
        }
    }
    if (other) {                                       // block-uniform: one coalesced atomic per channel of the tile
        __syncthreads();
        if (threadIdx.x < LT_C && c0 + threadIdx.x < C) atomicAdd(red + b * C + c0 + threadIdx.x, csum[threadIdx.x]);
    }
}

}  // namespace sg2

using namespace sg2;

static int check_layout_args(const char *what, const void *a, const void *b, const void *other, const float *red, int64_t B,
                             int C, int64_t HW) {
    SG2_REQUIRE(B >= 0 && B <= 65535 && C >= 2 && C % 2 == 0 && C <= 65535 * LT_C && HW >= 1, SG2_ERR_BAD_ARG,
                "%s: bad shape (B %lld, C %d, HW %lld; C must be even)", what, (long long)B, C, (long long)HW);
    if (B == 0) return SG2_OK;
    SG2_REQUIRE(a && b, SG2_ERR_BAD_ARG, "%s: null pointer", what);
    SG2_REQUIRE((other == nullptr) == (red == nullptr), SG2_ERR_BAD_ARG, "%s: the reduction needs both its operand and its output", what);
    return SG2_OK;
}

static int run_nchw_to_nhwc(void *out, const void *x, const float *scale, const void *other, float *red, int64_t B, int C,
                            int64_t HW, int dtype, sg2_stream_t stream, int poly_w) {
    int rc = check_layout_args("nchw_to_nhwc_bf16", out, x, other, red, B, C, HW);
    if (rc || B == 0) return rc;
    SG2_REQUIRE((reinterpret_cast<uintptr_t>(out) & 3) == 0 && (reinterpret_cast<uintptr_t>(other) & 3) == 0, SG2_ERR_BAD_ARG,
                "nchw_to_nhwc_bf16: NHWC tensors must be 4-byte aligned");
    dim3 grid((unsigned)((HW + LT_P - 1) / LT_P), (unsigned)((C + LT_C - 1) / LT_C), (unsigned)B);
    SG2_DISPATCH_DTYPE(dtype, {
        nchw_to_nhwc_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((__nv_bfloat16 *)out, (const T *)x, scale,
                                                                     (const __nv_bfloat16 *)other, red, C, (long long)HW,
                                                                     nullptr, 0.f, 1.f, nullptr, poly_w);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}

static int run_nhwc_to_nchw(void *out, const void *h, const float *scale, const void *other, float *red, int64_t B, int C,
                            int64_t HW, int dtype, sg2_stream_t stream, int poly_w, int n_parts = 0, int part_pitch = 0) {
    int rc = check_layout_args("nhwc_bf16_to_nchw", out, h, other, red, B, C, HW);
    if (rc || B == 0) return rc;
    SG2_REQUIRE((reinterpret_cast<uintptr_t>(h) & 3) == 0, SG2_ERR_BAD_ARG, "nhwc_bf16_to_nchw: NHWC tensors must be 4-byte aligned");
    dim3 grid((unsigned)((HW + LT_P - 1) / LT_P), (unsigned)((C + LT_C - 1) / LT_C), (unsigned)B);
    SG2_DISPATCH_DTYPE(dtype, {
        nhwc_to_nchw_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((T *)out, (const __nv_bfloat16 *)h, scale, (const T *)other,
                                                                     red, C, (long long)HW, 0, nullptr, 0, nullptr, nullptr, 0.f, 1.f, poly_w, n_parts, part_pitch);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}

// out[b,c,p] = lrelu(scale[b,c] * h[b,p,c] + noise_weight[0] * noise[b or 0, p] + bias[c], alpha) * gain: the demodulation,
// NoiseInjection and FusedLeakyReLU (model.py:239-240,282-287,335) in the NHWC -> NCHW pass.  noise / noise_weight / bias
// are tensors of `dtype` (noise and bias may be NULL); noise_bstride = HW for per-sample noise, 0 for one shared map.
extern "C" int sg2_nhwc_bf16_to_nchw_act(void *out, const void *h, const float *scale, const void *noise, int64_t noise_bstride,
                                         const void *noise_weight, const void *bias, float alpha, float gain, int64_t B, int C,
                                         int64_t HW, int dtype, sg2_stream_t stream) {
    int rc = check_layout_args("nhwc_bf16_to_nchw_act", out, h, nullptr, nullptr, B, C, HW);
    if (rc || B == 0) return rc;
    SG2_REQUIRE((reinterpret_cast<uintptr_t>(h) & 3) == 0, SG2_ERR_BAD_ARG, "nhwc_bf16_to_nchw_act: NHWC tensors must be 4-byte aligned");
    SG2_REQUIRE(!noise || noise_weight, SG2_ERR_BAD_ARG, "nhwc_bf16_to_nchw_act: noise without its weight");
    SG2_REQUIRE(noise_bstride == 0 || noise_bstride == HW, SG2_ERR_BAD_ARG, "nhwc_bf16_to_nchw_act: noise stride must be 0 or HW");
    dim3 grid((unsigned)((HW + LT_P - 1) / LT_P), (unsigned)((C + LT_C - 1) / LT_C), (unsigned)B);
    SG2_DISPATCH_DTYPE(dtype, {
        nhwc_to_nchw_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((T *)out, (const __nv_bfloat16 *)h, scale, nullptr, nullptr, C,
                                                                     (long long)HW, 1, (const T *)noise, (long long)noise_bstride,
                                                                     (const T *)noise_weight, (const T *)bias, alpha, gain, 0, 0, 0);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}

// The adjoint of the pass above: g = gy * (y > 0 ? 1 : alpha) * gain (y = its output), then as sg2_nchw_to_nhwc_bf16 on g:
// out[b,p,c] = bf16(g * scale[b,c]), red[b,c] += sum_p g * other[b,p,c] (grad of the demodulation factor; other/red may be
// NULL), and red_sum[b,c] += sum_p g (summed over b by the caller: grad of the bias; may be NULL).
extern "C" int sg2_nchw_to_nhwc_bf16_actgrad(void *out, const void *gy, const void *y, float alpha, float gain, const float *scale,
                                             const void *other, float *red, float *red_sum, int64_t B, int C, int64_t HW,
                                             int dtype, sg2_stream_t stream) {
    int rc = check_layout_args("nchw_to_nhwc_bf16_actgrad", out, gy, other, red, B, C, HW);
    if (rc || B == 0) return rc;
    SG2_REQUIRE(y, SG2_ERR_BAD_ARG, "nchw_to_nhwc_bf16_actgrad: null activation output");
    SG2_REQUIRE((reinterpret_cast<uintptr_t>(out) & 3) == 0 && (reinterpret_cast<uintptr_t>(other) & 3) == 0, SG2_ERR_BAD_ARG,
                "nchw_to_nhwc_bf16_actgrad: NHWC tensors must be 4-byte aligned");
    dim3 grid((unsigned)((HW + LT_P - 1) / LT_P), (unsigned)((C + LT_C - 1) / LT_C), (unsigned)B);
    SG2_DISPATCH_DTYPE(dtype, {
        nchw_to_nhwc_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((__nv_bfloat16 *)out, (const T *)gy, scale,
                                                                     (const __nv_bfloat16 *)other, red, C, (long long)HW,
                                                                     (const T *)y, alpha, gain, red_sum, 0);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}

extern "C" int sg2_nchw_to_nhwc_bf16(void *out, const void *x, const float *scale, const void *other, float *red, int64_t B,
                                     int C, int64_t HW, int dtype, sg2_stream_t stream) {
    return run_nchw_to_nhwc(out, x, scale, other, red, B, C, HW, dtype, stream, 0);
}
extern "C" int sg2_nhwc_bf16_to_nchw(void *out, const void *h, const float *scale, const void *other, float *red, int64_t B,
                                     int C, int64_t HW, int dtype, sg2_stream_t stream) {
    return run_nhwc_to_nchw(out, h, scale, other, red, B, C, HW, dtype, stream, 0);
}

// The same two passes with the NHWC side stored as the four polyphase planes of a W x W image (W odd):
// planes[(y&1)*2 + (x&1)][b][y>>1][x>>1][c], each plane (W+1)/2 squared -- the layout sg2_conv_transpose3x3_tc writes and
// the polyphase input gradient reads.  sg2_nchw_to_polyphase_bf16 writes only the valid cells: zero the planes first.
extern "C" int sg2_nchw_to_polyphase_bf16(void *planes, const void *x, const float *scale, const void *other_planes, float *red,
                                          int64_t B, int C, int W, int dtype, sg2_stream_t stream) {
    SG2_REQUIRE(W >= 1 && W % 2 == 1 && W <= 32767, SG2_ERR_BAD_ARG, "nchw_to_polyphase_bf16: W must be odd, got %d", W);
    return run_nchw_to_nhwc(planes, x, scale, other_planes, red, B, C, (int64_t)W * W, dtype, stream, W);
}
extern "C" int sg2_polyphase_bf16_to_nchw(void *out, const void *planes, const float *scale, const void *other, float *red,
                                          int64_t B, int C, int W, int dtype, sg2_stream_t stream) {
    SG2_REQUIRE(W >= 1 && W % 2 == 1 && W <= 32767, SG2_ERR_BAD_ARG, "polyphase_bf16_to_nchw: W must be odd, got %d", W);
    return run_nhwc_to_nchw(out, planes, scale, other, red, B, C, (int64_t)W * W, dtype, stream, W);
}

// out[b,c,y,x] = scale[b,c] * sum_k parts[k][b][y][x][c] for the top-left W x W corner of n_parts (1..4) tensors
// [B][pitch][pitch][C] bf16, summed in fp32; other / red as in sg2_nhwc_bf16_to_nchw.  The four polyphase components of the
// transposed conv's input gradient (sg2_conv_taps_tc outputs) become grad_x and grad_s in this one pass.
extern "C" int sg2_sum_parts_bf16_to_nchw(void *out, const void *parts, int n_parts, int pitch, const float *scale, const void *other,
                                          float *red, int64_t B, int C, int W, int dtype, sg2_stream_t stream) {
    SG2_REQUIRE(n_parts >= 1 && n_parts <= 4 && W >= 1 && pitch >= W && pitch <= 32767, SG2_ERR_BAD_ARG,
                "sum_parts_bf16_to_nchw: bad geometry (%d parts, W %d, pitch %d)", n_parts, W, pitch);
    return run_nhwc_to_nchw(out, parts, scale, other, red, B, C, (int64_t)W * W, dtype, stream, W, n_parts, pitch);
}
