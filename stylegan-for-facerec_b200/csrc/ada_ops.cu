// ada_ops.cu -- the two ops the StyleGAN2-ADA variant of the decoder adds to the kernel set
// (restyle-encoder/models/stylegan2_ada of the reference, selected by psp.py:24-30):
//
//   * SmoothUpsample (stylegan2_ada/utils.py:76-95): nearest x2 -> ReplicationPad2d((2,1,2,1)) -> 4x4 FIR
//     (a correlation, conv2d does not flip), fused here with what always follows it: + noise * strength,
//     + bias, leaky-relu, x gain, clamp (SynthesisLayer2.forward, generator.py:198-204) or + the ToRGB
//     output (SynthesisBlock.forward, generator.py:134-137).  The nearest up-sampling is never
//     materialised: output (2i+py, 2j+px) is a 16-tap sum over the clamped 3x3 input neighbourhood of
//     (i, j); a thread owns one input cell = a 2x2 output block, 9 loads for 4 outputs.
//   * bias + activation + gain + clamp (clamp_gain, utils.py:6-7) with the optional noise term, for the
//     layers that do not up-sample and for ToRGBLayer2 (generator.py:148-151).
//
// NCHW, fp32 / fp16 / bf16 storage, fp32 math.  HBM-bound elementwise work.
#include <algorithm>

#include "common.cuh"

namespace sg2 {

struct AdaEpilogue {
    const void *noise;          // [B or 1, 1, OH, OW] or null
    long long noise_bstride;    // OH*OW, or 0 for one map broadcast over the batch
    const void *noise_strength; // [1]
    const void *bias;           // [C] or null
    const void *addend;         // [B, C, OH, OW] or null
    int act;                    // 1 linear, 3 leaky-relu(alpha)
    float alpha, gain, clamp;   // clamp <= 0: none
};

template <typename T>
__device__ __forceinline__ float ada_finish(float v, const AdaEpilogue &e, float nstr, float bias, long long plane_px,
                                            long long b, long long idx) {
    if (e.noise) v += Cvt<T>::to_f(static_cast<const T *>(e.noise)[b * e.noise_bstride + plane_px]) * nstr;
    v += bias;
    if (e.addend) v += Cvt<T>::to_f(static_cast<const T *>(e.addend)[idx]);
    if (e.act == 3) v = v > 0.f ? v : v * e.alpha;
    v *= e.gain;
    if (e.clamp > 0.f) v = fminf(fmaxf(v, -e.clamp), e.clamp);
    return v;
}

// one thread = one input cell (i, j) of one plane = output block (2i..2i+1, 2j..2j+1)
template <typename T>
__global__ void __launch_bounds__(256)
smooth_up2x_kernel(T *__restrict__ out, const T *__restrict__ x, const float *__restrict__ taps, long long planes, int C, int H,
                   int W, AdaEpilogue e) {
    __shared__ float s_k[16];
    if (threadIdx.x < 16) s_k[threadIdx.x] = taps[threadIdx.x];
    __syncthreads();
    const int j = blockIdx.x * 32 + (threadIdx.x & 31), i = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (i >= H || j >= W) return;
    const float nstr = e.noise ? Cvt<T>::to_f(static_cast<const T *>(e.noise_strength)[0]) : 0.f;
    const int im = max(i - 1, 0), ip = min(i + 1, H - 1), jm = max(j - 1, 0), jp = min(j + 1, W - 1);
    const int OW = 2 * W;
    const long long opx = 4LL * H * W;
    // window row / column of tap a (resp. b) for output parity 0 and 1: rows 2i+p+a-2 of the nearest-up-sampled,
    // edge-replicated image are input rows (i-1, i-1, i, i) for p = 0 and (i-1, i, i, i+1) for p = 1
    const int sel[2][4] = {{0, 0, 1, 1}, {0, 1, 1, 2}};
    for (long long plane = blockIdx.z; plane < planes; plane += gridDim.z) {
        const T *xp = x + plane * (long long)H * W;
        float w[3][3];
        const int rr[3] = {im, i, ip}, cc[3] = {jm, j, jp};
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) w[r][c] = Cvt<T>::to_f(xp[(long long)rr[r] * W + cc[c]]);
        const long long b = plane / C;
        const int ch = (int)(plane - b * C);
        const float bias = e.bias ? Cvt<T>::to_f(static_cast<const T *>(e.bias)[ch]) : 0.f;
        T *op = out + plane * opx;
#pragma unroll
        for (int py = 0; py < 2; ++py) {
            float v[2];
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                float acc = 0.f;
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb) acc = fmaf(s_k[a * 4 + bb], w[sel[py][a]][sel[px][bb]], acc);
                const long long ppx = (long long)(2 * i + py) * OW + 2 * j + px;
                v[px] = ada_finish<T>(acc, e, nstr, bias, ppx, b, plane * opx + ppx);
            }
            T *dst = op + (long long)(2 * i + py) * OW + 2 * j;
            dst[0] = Cvt<T>::from_f(v[0]);
            dst[1] = Cvt<T>::from_f(v[1]);
        }
    }
}

// adjoint of the plain SmoothUpsample (no epilogue): grad_x[i][j] = sum over the outputs (Y, X) and taps (a, b) whose
// clamped source cell is (i, j) of taps[a][b] * grad_out[Y][X].  One thread per input cell; the (few) matching
// (output row, tap row) and (output column, tap column) pairs are enumerated with the forward's own index rule, so
// the edge replication is transposed exactly.
template <typename T>
__global__ void __launch_bounds__(256)
smooth_up2x_bwd_kernel(T *__restrict__ gx, const T *__restrict__ gy, const float *__restrict__ taps, long long planes, int H,
                       int W) {
    __shared__ float s_k[16];
    if (threadIdx.x < 16) s_k[threadIdx.x] = taps[threadIdx.x];
    __syncthreads();
    const int j = blockIdx.x * 32 + (threadIdx.x & 31), i = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (i >= H || j >= W) return;
    const int off[2][4] = {{-1, -1, 0, 0}, {-1, 0, 0, 1}};      // source cell of tap a relative to i' for output parity p
    // matching (output index, tap) pairs along one axis: at most 3 cells x 2 parities x 4 taps, typically 6
    int ry[12], ra[12], nr = 0, cx[12], cb[12], nc = 0;
    for (int ii = max(i - 1, 0); ii <= min(i + 1, H - 1); ++ii)
        for (int py = 0; py < 2; ++py)
            for (int a = 0; a < 4; ++a)
                if (min(max(ii + off[py][a], 0), H - 1) == i && nr < 12) { ry[nr] = 2 * ii + py; ra[nr] = a; ++nr; }
    for (int jj = max(j - 1, 0); jj <= min(j + 1, W - 1); ++jj)
        for (int px = 0; px < 2; ++px)
            for (int b = 0; b < 4; ++b)
                if (min(max(jj + off[px][b], 0), W - 1) == j && nc < 12) { cx[nc] = 2 * jj + px; cb[nc] = b; ++nc; }
    const int OW = 2 * W;
    for (long long plane = blockIdx.z; plane < planes; plane += gridDim.z) {
        const T *gp = gy + plane * 4LL * H * W;
        float acc = 0.f;
        for (int r = 0; r < nr; ++r)
            for (int c = 0; c < nc; ++c) acc = fmaf(s_k[ra[r] * 4 + cb[c]], Cvt<T>::to_f(gp[(long long)ry[r] * OW + cx[c]]), acc);
        gx[plane * (long long)H * W + (long long)i * W + j] = Cvt<T>::from_f(acc);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
ada_bias_act_kernel(T *__restrict__ out, const T *__restrict__ x, long long total, int C, long long HW, AdaEpilogue e) {
    const float nstr = e.noise ? Cvt<T>::to_f(static_cast<const T *>(e.noise_strength)[0]) : 0.f;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const long long plane = idx / HW, ppx = idx - plane * HW;
        const long long b = plane / C;
        const int ch = (int)(plane - b * C);
        const float bias = e.bias ? Cvt<T>::to_f(static_cast<const T *>(e.bias)[ch]) : 0.f;
        out[idx] = Cvt<T>::from_f(ada_finish<T>(Cvt<T>::to_f(x[idx]), e, nstr, bias, ppx, b, idx));
    }
}

}  // namespace sg2

using namespace sg2;

extern "C" int sg2_smooth_upsample2x(void *out, const void *x, const float *taps, int64_t B, int C, int H, int W,
                                     const void *noise, int64_t noise_bstride, const void *noise_strength, const void *bias,
                                     const void *addend, int act, float alpha, float gain, float clamp, int dtype,
                                     sg2_stream_t stream) {
    SG2_REQUIRE(B >= 0 && C >= 1 && H >= 1 && W >= 1, SG2_ERR_BAD_ARG, "smooth_upsample2x: bad shape");
    SG2_REQUIRE(act == 1 || act == 3, SG2_ERR_BAD_ARG, "smooth_upsample2x: act must be 1 (linear) or 3 (leaky-relu)");
    if (B == 0) return SG2_OK;
    SG2_REQUIRE(out && x && taps && (!noise || noise_strength), SG2_ERR_BAD_ARG, "smooth_upsample2x: null pointer");
    AdaEpilogue e{noise, (long long)noise_bstride, noise_strength, bias, addend, act, alpha, gain, clamp};
    const long long planes = (long long)B * C;
    dim3 grid((W + 31) / 32, (H + 7) / 8, (unsigned)std::min<long long>(planes, 16384));
    SG2_DISPATCH_DTYPE(dtype, {
        smooth_up2x_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((T *)out, (const T *)x, taps, planes, C, H, W, e);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}

extern "C" int sg2_ada_bias_act(void *out, const void *x, const void *noise, int64_t noise_bstride, const void *noise_strength,
                                const void *bias, int64_t B, int C, int64_t HW, int act, float alpha, float gain, float clamp,
                                int dtype, sg2_stream_t stream) {
    SG2_REQUIRE(B >= 0 && C >= 1 && HW >= 1, SG2_ERR_BAD_ARG, "ada_bias_act: bad shape");
    SG2_REQUIRE(act == 1 || act == 3, SG2_ERR_BAD_ARG, "ada_bias_act: act must be 1 (linear) or 3 (leaky-relu)");
    if (B == 0) return SG2_OK;
    SG2_REQUIRE(out && x && (!noise || noise_strength), SG2_ERR_BAD_ARG, "ada_bias_act: null pointer");
    AdaEpilogue e{noise, (long long)noise_bstride, noise_strength, bias, nullptr, act, alpha, gain, clamp};
    const long long total = (long long)B * C * HW;
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, (long long)sm_count() * 32);
    SG2_DISPATCH_DTYPE(dtype, {
        ada_bias_act_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>((T *)out, (const T *)x, total, C, (long long)HW, e);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}

extern "C" int sg2_smooth_upsample2x_bwd(void *grad_x, const void *grad_out, const float *taps, int64_t planes, int H, int W,
                                         int dtype, sg2_stream_t stream) {
    SG2_REQUIRE(planes >= 0 && H >= 1 && W >= 1, SG2_ERR_BAD_ARG, "smooth_upsample2x_bwd: bad shape");
    if (planes == 0) return SG2_OK;
    SG2_REQUIRE(grad_x && grad_out && taps, SG2_ERR_BAD_ARG, "smooth_upsample2x_bwd: null pointer");
    dim3 grid((W + 31) / 32, (H + 7) / 8, (unsigned)std::min<long long>(planes, 16384));
    SG2_DISPATCH_DTYPE(dtype, {
        smooth_up2x_bwd_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((T *)grad_x, (const T *)grad_out, taps, (long long)planes, H, W);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}
