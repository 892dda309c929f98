// linear.cu -- the small dense pieces around the convolutions (all warp-shuffle kernels):
//   * EqualLinear (model.py:147-157)                      -> equal_linear_kernel
//   * mapping network PixelNorm + n_mlp x EqualLinear     -> mapping_layer_kernel, one launch per
//     layer with PixelNorm fused into the first (model.py:10-15, 378-387; 36 launches in the reference)
//   * modulation affine + demodulation coefficients       -> modulation kernels
//     (model.py:235-240), computed WITHOUT materialising per-sample weights:
//         demod[b,co] = rsqrt( sum_ci style[b,ci]^2 * wsq[ci,co] + 1e-8 ),
//         wsq[ci,co]  = conv_scale^2 * sum_k weight[co,ci,k]^2
//     which is algebraically the reference's rsqrt(sum (scale*W*s)^2 + 1e-8).
#include <stdlib.h>

#include "common.cuh"

namespace sg2 {

constexpr float kLreluSlope = 0.2f;
constexpr float kLreluGain = 1.41421356237309515f;   // 2 ** 0.5 as the reference's Python float

// out[b,o] = act((x[b,:] . w[o,:]) * w_scale + bias[o]*lr_mul); one warp per output feature o,
// the weight row lives in registers and is reused for every sample of the batch chunk.
template <typename T, int MAXJ>   // in_dim <= 32*MAXJ
__global__ void __launch_bounds__(256)
equal_linear_kernel(T *__restrict__ out, const T *__restrict__ x, const T *__restrict__ w,
                    const T *__restrict__ bias, int64_t B, int in_dim, int out_dim, float w_scale,
                    float lr_mul, int act, int64_t b_chunk) {
    const int lane = threadIdx.x & 31;
    const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (o >= out_dim) return;
    float wr[MAXJ];
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
        const int idx = lane + 32 * j;
        wr[j] = idx < in_dim ? Cvt<T>::to_f(w[(int64_t)o * in_dim + idx]) : 0.f;
    }
    const float bv = bias ? Cvt<T>::to_f(bias[o]) * lr_mul : 0.f;
    const int64_t b0 = (int64_t)blockIdx.y * b_chunk;
    const int64_t b1 = b0 + b_chunk < B ? b0 + b_chunk : B;
    for (int64_t b = b0; b < b1; ++b) {
        const T *xr = x + b * in_dim;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            const int idx = lane + 32 * j;
            if (idx < in_dim) acc += Cvt<T>::to_f(xr[idx]) * wr[j];
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            float v = acc * w_scale + bv;
            if (act) v = (v > 0.f ? v : v * kLreluSlope) * kLreluGain;
            out[b * out_dim + o] = Cvt<T>::from_f(v);
        }
    }
}

// same contract for wide inputs (in_dim > 2048, e.g. the discriminator's 8192 -> 512 layer):
// the weight row is re-read through L1/L2 instead of being held in registers
template <typename T>
__global__ void __launch_bounds__(256)
equal_linear_wide_kernel(T *__restrict__ out, const T *__restrict__ x, const T *__restrict__ w,
                         const T *__restrict__ bias, int64_t B, int in_dim, int out_dim,
                         float w_scale, float lr_mul, int act, int64_t b_chunk) {
    const int lane = threadIdx.x & 31;
    const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (o >= out_dim) return;
    const float bv = bias ? Cvt<T>::to_f(bias[o]) * lr_mul : 0.f;
    const int64_t b0 = (int64_t)blockIdx.y * b_chunk;
    const int64_t b1 = b0 + b_chunk < B ? b0 + b_chunk : B;
    const T *wr = w + (int64_t)o * in_dim;
    for (int64_t b = b0; b < b1; ++b) {
        const T *xr = x + b * in_dim;
        float acc = 0.f;
        for (int idx = lane; idx < in_dim; idx += 32) acc += Cvt<T>::to_f(xr[idx]) * Cvt<T>::to_f(wr[idx]);
        acc = warp_sum(acc);
        if (lane == 0) {
            float v = acc * w_scale + bv;
            if (act) v = (v > 0.f ? v : v * kLreluSlope) * kLreluGain;
            out[b * out_dim + o] = Cvt<T>::from_f(v);
        }
    }
}

// weight [Cout,Cin,k,k] -> wt[(ci*kk + t)*Cout + co] = conv_scale*w ; wsq[ci*Cout+co] = conv_scale^2*sum_t w^2
template <typename T>
__global__ void __launch_bounds__(256)
modconv_prep_kernel(float *__restrict__ wt, float *__restrict__ wsq, const T *__restrict__ weight,
                    int Cin, int Cout, int kk, float conv_scale) {
    // tile transpose through smem: block handles 32 co x 32 ci
    __shared__ float tile[32][9][33];
    const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {           // r: co within tile, tx: ci within tile
        const int co = co0 + r, ci = ci0 + tx;
        for (int t = 0; t < kk; ++t)
            tile[r][t][tx] = (co < Cout && ci < Cin)
                                 ? Cvt<T>::to_f(weight[((int64_t)co * Cin + ci) * kk + t]) * conv_scale : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {           // r: ci within tile, tx: co within tile
        const int ci = ci0 + r, co = co0 + tx;
        if (ci < Cin && co < Cout) {
            float ss = 0.f;
            for (int t = 0; t < kk; ++t) {
                const float v = tile[tx][t][r];
                wt[((int64_t)ci * kk + t) * Cout + co] = v;
                ss += v * v;
            }
            if (wsq) wsq[(int64_t)ci * Cout + co] = ss;
        }
    }
}

// style[b,ci] = (latent[b,:] . mod_w[ci,:]) * mod_scale + mod_b[ci]*lr_mul  (fp32 out); warp per ci
template <typename T, int MAXJ>
__global__ void __launch_bounds__(256)
modulation_style_kernel(float *__restrict__ style, const T *__restrict__ latent,
                        int64_t latent_stride, const T *__restrict__ mod_w,
                        const T *__restrict__ mod_b, int64_t B, int style_dim, int Cin,
                        float mod_scale, float lr_mul) {
    const int lane = threadIdx.x & 31;
    const int ci = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (ci >= Cin) return;
    float wr[MAXJ];
#pragma unroll
    for (int j = 0; j < MAXJ; ++j) {
        const int idx = lane + 32 * j;
        wr[j] = idx < style_dim ? Cvt<T>::to_f(mod_w[(int64_t)ci * style_dim + idx]) : 0.f;
    }
    const float bv = Cvt<T>::to_f(mod_b[ci]) * lr_mul;
    for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
        const T *lr = latent + b * latent_stride;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
            const int idx = lane + 32 * j;
            if (idx < style_dim) acc += Cvt<T>::to_f(lr[idx]) * wr[j];
        }
        acc = warp_sum(acc);
        if (lane == 0) style[b * Cin + ci] = acc * mod_scale + bv;
    }
}

// demod[b,co] = rsqrt(sum_ci style[b,ci]^2 * wsq[ci,co] + 1e-8); thread per (b,co), coalesced on co
__global__ void __launch_bounds__(256)
modulation_demod_kernel(float *__restrict__ demod, const float *__restrict__ style,
                        const float *__restrict__ wsq, int Cin, int Cout) {
    extern __shared__ float s_s2[];   // [Cin]
    const int64_t b = blockIdx.y;
    for (int i = threadIdx.x; i < Cin; i += 256) {
        const float s = style[b * Cin + i];
        s_s2[i] = s * s;
    }
    __syncthreads();
    const int co = blockIdx.x * 256 + threadIdx.x;
    if (co >= Cout) return;
    float acc = 0.f;
    for (int ci = 0; ci < Cin; ++ci) acc += s_s2[ci] * wsq[(int64_t)ci * Cout + co];
    demod[b * Cout + co] = rsqrtf(acc + 1e-8f);
}

}  // namespace sg2

using namespace sg2;

extern "C" int sg2_equal_linear_fwd(void *out, const void *x, const void *w, const void *bias,
                                    int64_t B, int in_dim, int out_dim, float w_scale, float lr_mul,
                                    int act, int dtype, sg2_stream_t stream) {
    SG2_REQUIRE(B >= 0 && in_dim >= 1 && out_dim >= 1, SG2_ERR_BAD_ARG, "equal_linear: bad shape");
    if (B == 0) return SG2_OK;
    SG2_REQUIRE(out && x && w, SG2_ERR_BAD_ARG, "equal_linear: null tensor pointer");
    SG2_REQUIRE(in_dim <= 8192, SG2_ERR_UNSUPPORTED, "equal_linear: in_dim %d > 8192", in_dim);
    cudaStream_t st = as_stream(stream);
    const int sms = sm_count();
    const unsigned gx = (out_dim + 7) / 8;
    // split the batch so the grid covers the chip about twice
    int64_t splits = std::max<int64_t>(1, std::min<int64_t>(B, (2 * sms + gx - 1) / gx));
    splits = std::min<int64_t>(splits, 65535);
    const int64_t chunk = ceil_div64(B, splits);
    dim3 grid(gx, (unsigned)ceil_div64(B, chunk));
#define SG2_EL(MJ) equal_linear_kernel<T, MJ><<<grid, 256, 0, st>>>((T *)out, (const T *)x, (const T *)w, (const T *)bias, B, in_dim, out_dim, w_scale, lr_mul, act, chunk)
    SG2_DISPATCH_DTYPE(dtype, {
        if (in_dim <= 512) SG2_EL(16);
        else if (in_dim <= 1024) SG2_EL(32);
        else if (in_dim <= 2048) SG2_EL(64);
        else
            equal_linear_wide_kernel<T><<<grid, 256, 0, st>>>((T *)out, (const T *)x, (const T *)w, (const T *)bias, B, in_dim, out_dim, w_scale, lr_mul, act, chunk);
        SG2_LAUNCH_CHECK();
    });
#undef SG2_EL
    return SG2_OK;
}

namespace sg2 {
// One mapping layer: out[b,o] = lrelu((xn[b,:] . w[o,:]) * w_scale + bias[o]*lr_mul) * sqrt(2), where
// xn = pixel_norm(x) for the first layer (model.py:14-15 fused into the load) and x otherwise.
// Block = 4 output features (their weight rows staged in shared memory) x up to 64 samples
// (8 warps x 8 samples); a lane keeps its slice of the sample row in registers, so every global
// load is a coalesced 128-byte line and the reduction is a warp shuffle.  The whole chip works on
// every layer (dim/4 x ceil(B/64) blocks) instead of the previous one-block-per-two-samples chain.
constexpr int MAP_ROWS = 4, MAP_SPB = 64;

template <typename T, int NJ>   // dim == 32 * NJ
__global__ void __launch_bounds__(256)
mapping_layer_kernel(T *__restrict__ out, const T *__restrict__ x, const T *__restrict__ w,
                     const T *__restrict__ bias, int64_t B, float w_scale, float lr_mul, int pixel_norm) {
    constexpr int dim = 32 * NJ;
    __shared__ float s_w[MAP_ROWS][dim];
    __shared__ float s_b[MAP_ROWS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int o0 = blockIdx.x * MAP_ROWS;
    for (int i = threadIdx.x; i < MAP_ROWS * dim; i += 256) {
        const int r = i / dim, j = i - r * dim;
        s_w[r][j] = Cvt<T>::to_f(w[(int64_t)(o0 + r) * dim + j]);
    }
    if (threadIdx.x < MAP_ROWS) s_b[threadIdx.x] = Cvt<T>::to_f(bias[o0 + threadIdx.x]) * lr_mul;
    __syncthreads();
    const int64_t b0 = (int64_t)blockIdx.y * MAP_SPB;
    for (int s = warp; s < MAP_SPB; s += 8) {
        const int64_t b = b0 + s;
        if (b >= B) break;
        float xv[NJ];
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            xv[j] = Cvt<T>::to_f(x[b * dim + lane + 32 * j]);
            ss += xv[j] * xv[j];
        }
        if (pixel_norm) {
            ss = warp_sum(ss);
            const float rn = rsqrtf(ss / (float)dim + 1e-8f);
#pragma unroll
            for (int j = 0; j < NJ; ++j) xv[j] *= rn;
        }
        float acc[MAP_ROWS];
#pragma unroll
        for (int r = 0; r < MAP_ROWS; ++r) {
            acc[r] = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) acc[r] += xv[j] * s_w[r][lane + 32 * j];
        }
#pragma unroll
        for (int r = 0; r < MAP_ROWS; ++r) acc[r] = warp_sum(acc[r]);
        if (lane < MAP_ROWS) {
            float v = (lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3]) * w_scale + s_b[lane];
            v = (v > 0.f ? v : v * kLreluSlope) * kLreluGain;
            out[b * dim + o0 + lane] = Cvt<T>::from_f(v);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
pixel_norm_kernel(T *__restrict__ out, const T *__restrict__ x, int64_t B, int dim, int pixel_norm) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    float ss = 0.f;
    for (int j = lane; j < dim; j += 32) { const float v = Cvt<T>::to_f(x[b * dim + j]); ss += v * v; }
    ss = warp_sum(ss);
    const float rn = pixel_norm ? rsqrtf(ss / (float)dim + 1e-8f) : 1.f;
    for (int j = lane; j < dim; j += 32) out[b * dim + j] = Cvt<T>::from_f(Cvt<T>::to_f(x[b * dim + j]) * rn);
}

// The same layer as a small tiled fp32 GEMM (an experiment against the shuffle-reduction kernel above, which is latency
// bound at ~12 us per layer for B = 64, 8 layers per forward; this one measured slower and is opt-in): block = MG_F output features x MG_S samples, K in chunks of MG_K through
// double-buffered shared memory with the next chunk prefetched into registers, a thread owns one feature x two
// samples.  PixelNorm of the first layer is a per-sample scale of the result (the warp that stages a sample's row
// also accumulates its sum of squares).
constexpr int MG_F = 8, MG_S = 64, MG_K = 32, MG_PITCH = MG_K + 4;
template <typename T>
__global__ void __launch_bounds__(256)
mapping_gemm_kernel(T *__restrict__ out, const T *__restrict__ x, const T *__restrict__ w, const T *__restrict__ bias,
                    int64_t B, int dim, float w_scale, float lr_mul, int pixel_norm) {
    __shared__ __align__(16) float s_x[2][MG_S][MG_PITCH];
    __shared__ __align__(16) float s_w[2][MG_F][MG_PITCH];
    __shared__ float s_rn[MG_S];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int o0 = blockIdx.x * MG_F;
    const int64_t b0 = (int64_t)blockIdx.y * MG_S;
    // staging roles: warp `warp` stages sample rows warp, warp + 8, ... (lane = k) and weight row `warp`
    float xr[8], wr, ss[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) ss[i] = 0.f;
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t b = b0 + warp + 8 * i;
            xr[i] = b < B ? Cvt<T>::to_f(x[b * dim + k0 + lane]) : 0.f;
        }
        wr = Cvt<T>::to_f(w[(int64_t)(o0 + warp) * dim + k0 + lane]);
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s_x[buf][warp + 8 * i][lane] = xr[i];
            ss[i] = fmaf(xr[i], xr[i], ss[i]);
        }
        s_w[buf][warp][lane] = wr;
    };
    const int f = tid & 7, sg = tid >> 3;          // compute role: feature f, samples sg and sg + 32
    float acc0 = 0.f, acc1 = 0.f;
    fetch(0);
    stash(0);
    __syncthreads();
    const int nchunks = dim / MG_K;
    for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        if (c + 1 < nchunks) fetch((c + 1) * MG_K);
#pragma unroll
        for (int k = 0; k < MG_K; k += 4) {
            const float4 w4 = *reinterpret_cast<const float4 *>(&s_w[buf][f][k]);
            const float4 xa = *reinterpret_cast<const float4 *>(&s_x[buf][sg][k]);
            const float4 xb = *reinterpret_cast<const float4 *>(&s_x[buf][sg + 32][k]);
            acc0 = fmaf(xa.x, w4.x, fmaf(xa.y, w4.y, fmaf(xa.z, w4.z, fmaf(xa.w, w4.w, acc0))));
            acc1 = fmaf(xb.x, w4.x, fmaf(xb.y, w4.y, fmaf(xb.z, w4.z, fmaf(xb.w, w4.w, acc1))));
        }
        if (c + 1 < nchunks) stash(buf ^ 1);       // the other buffer was last read before the previous barrier
        __syncthreads();
    }
    if (pixel_norm) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float t = warp_sum(ss[i]);
            if (lane == 0) s_rn[warp + 8 * i] = rsqrtf(t / (float)dim + 1e-8f);   // model.py:15
        }
        __syncthreads();
    }
    const float bv = Cvt<T>::to_f(bias[o0 + f]) * lr_mul;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int sidx = sg + 32 * h;
        const int64_t b = b0 + sidx;
        if (b >= B) continue;
        float v = (h == 0 ? acc0 : acc1) * (pixel_norm ? s_rn[sidx] : 1.f) * w_scale + bv;
        v = (v > 0.f ? v : v * kLreluSlope) * kLreluGain;
        out[b * dim + o0 + f] = Cvt<T>::from_f(v);
    }
}

template <typename T>
static int launch_mapping_layer(T *out, const T *x, const T *w, const T *bias, int64_t B, int dim,
                                float w_scale, float lr_mul, int pixel_norm, cudaStream_t st) {
    // A/B switch: the tiled-GEMM kernel.  Measured in the step (tools/timeline.py): 19 us per layer vs 11.6 us for the
    // shuffle-reduction kernel at B = 64 (64 blocks x 16 synchronised K chunks is too little parallelism) -> off.
    static const char *env_gemm = getenv("SG2_MAPPING_GEMM");
    if (env_gemm && atoi(env_gemm)) {
        dim3 g2(dim / MG_F, (unsigned)ceil_div64(B, MG_S));
        mapping_gemm_kernel<T><<<g2, 256, 0, st>>>(out, x, w, bias, B, dim, w_scale, lr_mul, pixel_norm);
        SG2_LAUNCH_CHECK();
        return SG2_OK;
    }
    dim3 grid(dim / MAP_ROWS, (unsigned)ceil_div64(B, MAP_SPB));
#define SG2_ML(NJ) case NJ: mapping_layer_kernel<T, NJ><<<grid, 256, 0, st>>>(out, x, w, bias, B, w_scale, lr_mul, pixel_norm); break
    switch (dim / 32) {
        SG2_ML(1); SG2_ML(2); SG2_ML(4); SG2_ML(8); SG2_ML(16);
        default: set_error("mapping: style_dim %d not in {32,64,128,256,512}", dim); return SG2_ERR_UNSUPPORTED;
    }
#undef SG2_ML
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}
}  // namespace sg2

extern "C" int sg2_mapping_fwd(void *w_out, const void *z, const void *const *weights,
                               const void *const *biases, int n_mlp, int64_t B, int dim,
                               float lr_mul, int pixel_norm, void *scratch, int dtype,
                               sg2_stream_t stream) {
    SG2_REQUIRE(B >= 0 && n_mlp >= 0 && n_mlp <= 32, SG2_ERR_BAD_ARG, "mapping: bad n_mlp/B");
    SG2_REQUIRE(dim == 32 || dim == 64 || dim == 128 || dim == 256 || dim == 512, SG2_ERR_UNSUPPORTED,
                "mapping: style_dim must be one of 32, 64, 128, 256, 512, got %d", dim);
    SG2_REQUIRE(B <= 65535ll * MAP_SPB, SG2_ERR_UNSUPPORTED, "mapping: batch too large");
    if (B == 0) return SG2_OK;
    SG2_REQUIRE(w_out && z && (n_mlp == 0 || (weights && biases)), SG2_ERR_BAD_ARG, "mapping: null pointer");
    SG2_REQUIRE(n_mlp <= 1 || scratch, SG2_ERR_BAD_ARG, "mapping: n_mlp > 1 needs a [B, dim] scratch buffer");
    for (int i = 0; i < n_mlp; ++i)
        SG2_REQUIRE(weights[i] && biases[i], SG2_ERR_BAD_ARG, "mapping: null layer pointer");
    const float w_scale = lr_mul / sqrtf((float)dim);   // model.py:144
    cudaStream_t st = as_stream(stream);
    SG2_DISPATCH_DTYPE(dtype, {
        if (n_mlp == 0) {
            pixel_norm_kernel<T><<<(unsigned)ceil_div64(B, 8), 256, 0, st>>>((T *)w_out, (const T *)z, B, dim, pixel_norm);
            SG2_LAUNCH_CHECK();
            return SG2_OK;
        }
        const T *src = (const T *)z;
        for (int l = 0; l < n_mlp; ++l) {
            // ping-pong so that the last layer lands in w_out
            T *dst = ((n_mlp - 1 - l) & 1) ? (T *)scratch : (T *)w_out;
            int rc = launch_mapping_layer<T>(dst, src, (const T *)weights[l], (const T *)biases[l], B, dim,
                                             w_scale, lr_mul, l == 0 ? pixel_norm : 0, st);
            if (rc) return rc;
            src = dst;
        }
    });
    return SG2_OK;
}

extern "C" int sg2_modconv2d_prep(float *wt, float *wsq, const void *weight, int Cin, int Cout,
                                  int k, float conv_scale, int dtype, sg2_stream_t stream) {
    SG2_REQUIRE(Cin >= 1 && Cout >= 1, SG2_ERR_BAD_ARG, "modconv2d_prep: bad channel counts");
    SG2_REQUIRE(k == 1 || k == 3, SG2_ERR_UNSUPPORTED, "modconv2d: kernel size must be 1 or 3, got %d", k);
    SG2_REQUIRE(wt && weight, SG2_ERR_BAD_ARG, "modconv2d_prep: null pointer");
    dim3 grid((Cout + 31) / 32, (Cin + 31) / 32);
    SG2_DISPATCH_DTYPE(dtype, {
        modconv_prep_kernel<T><<<grid, 256, 0, as_stream(stream)>>>(wt, wsq, (const T *)weight, Cin, Cout, k * k, conv_scale);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}

extern "C" int sg2_modulation_fwd(float *style, float *demod, const void *latent,
                                  int64_t latent_stride, const void *mod_w, const void *mod_b,
                                  const float *wsq, int64_t B, int style_dim, int Cin, int Cout,
                                  float mod_scale, float lr_mul, int dtype, sg2_stream_t stream) {
    SG2_REQUIRE(B >= 0 && style_dim >= 1 && Cin >= 1, SG2_ERR_BAD_ARG, "modulation: bad shape");
    if (B == 0) return SG2_OK;
    SG2_REQUIRE(style && latent && mod_w && mod_b, SG2_ERR_BAD_ARG, "modulation: null pointer");
    SG2_REQUIRE(style_dim <= 2048, SG2_ERR_UNSUPPORTED, "modulation: style_dim %d > 2048", style_dim);
    SG2_REQUIRE(!demod || (wsq && Cout >= 1), SG2_ERR_BAD_ARG, "modulation: demod needs wsq and Cout");
    SG2_REQUIRE(B <= 65535 * 64ll, SG2_ERR_UNSUPPORTED, "modulation: batch too large");
    cudaStream_t st = as_stream(stream);
    dim3 grid((Cin + 7) / 8, (unsigned)std::min<int64_t>(B, 4096));
    SG2_DISPATCH_DTYPE(dtype, {
        if (style_dim <= 512)
            modulation_style_kernel<T, 16><<<grid, 256, 0, st>>>(style, (const T *)latent, latent_stride, (const T *)mod_w, (const T *)mod_b, B, style_dim, Cin, mod_scale, lr_mul);
        else
            modulation_style_kernel<T, 64><<<grid, 256, 0, st>>>(style, (const T *)latent, latent_stride, (const T *)mod_w, (const T *)mod_b, B, style_dim, Cin, mod_scale, lr_mul);
        SG2_LAUNCH_CHECK();
    });
    if (demod) {
        SG2_REQUIRE(B <= 65535, SG2_ERR_UNSUPPORTED, "modulation: batch > 65535 with demodulation");
        SG2_REQUIRE(Cin <= 12000, SG2_ERR_UNSUPPORTED, "modulation: Cin too large");
        dim3 g2((Cout + 255) / 256, (unsigned)B);
        modulation_demod_kernel<<<g2, 256, sizeof(float) * Cin, st>>>(demod, style, wsq, Cin, Cout);
        SG2_LAUNCH_CHECK();
    }
    return SG2_OK;
}
