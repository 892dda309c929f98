// bias_act.cu -- fused bias + activation family (HBM-bound, 16-byte vectorised, 64-bit safe).
//
// Replaces fused_bias_act_kernel (op/fused_bias_act_kernel.cu:18-49 of the reference): same
// arithmetic order  y = act(x + b) * scale  so fp32 results are bit-identical, but
//   * 128-bit loads/stores, 4 vectors in flight per thread, grid sized to the SM count;
//   * the bias row is found with one multiply-shift per 16-byte vector instead of an integer
//     div + mod per element;
//   * 64-bit element counts (the reference's `int size_x` overflows at 2^31 elements, :65).
// Also here: the noise-injection + bias + activation fusion (model.py:282-287,335) and the
// grad_bias reduction (op/fused_act.py:31-36).
#include "common.cuh"

namespace sg2 {

// q = n / d for 0 <= n < 2^31, d >= 1, by multiply-high + shift
struct FastDiv {
    uint32_t d, mul, shr;
    FastDiv() : d(1), mul(0), shr(0) {}
    explicit FastDiv(uint32_t dd) : d(dd) {
        if (dd == 1) { mul = 0; shr = 0; return; }
        uint32_t lg = 0;
        while ((1ull << lg) < dd) ++lg;            // ceil(log2 d)
        uint32_t p = 31 + lg;
        mul = (uint32_t)(((1ull << p) + dd - 1) / dd);
        shr = p - 32;
    }
    __host__ __device__ __forceinline__ uint32_t div(uint32_t n) const {
#ifdef __CUDA_ARCH__
        return d == 1 ? n : (__umulhi(n, mul) >> shr);
#else
        return d == 1 ? n : (uint32_t)(((uint64_t)n * mul) >> 32) >> shr;
#endif
    }
};

template <int ACT, int GRAD>
__device__ __forceinline__ float act_apply(float x, float ref, float alpha) {
    if (ACT == 3) {
        if (GRAD == 0) return x > 0.f ? x : x * alpha;
        if (GRAD == 1) return ref > 0.f ? x : x * alpha;
        return 0.f;
    }
    return GRAD == 2 ? 0.f : x;
}

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

// Vector kernel.  Element i belongs to "row" i / step_b and bias channel row % size_b.
// BIAS_MODE 0: none; 1: one bias per vector (step_b % N == 0); 2: bias varies inside the vector
// with step_b == 1 and size_b % N == 0 (2-D inputs such as the mapping MLP).
template <typename T, int ACT, int GRAD, int BIAS_MODE, bool USE_REF, bool SMALL>
__global__ void __launch_bounds__(kThreads)
bias_act_vec_kernel(T *__restrict__ out, const T *__restrict__ x, const T *__restrict__ bias,
                    const T *__restrict__ ref, int64_t n_vec, FastDiv row_div, FastDiv ch_div,
                    int64_t step_v, int64_t size_b, float alpha, float scale) {
    constexpr int N = Vec16<T>::N;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t base = (int64_t)blockIdx.x * kThreads * kUnroll + threadIdx.x; base < n_vec;
         base += stride * kUnroll) {
        Vec16<T> xv[kUnroll], rv[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t v = base + (int64_t)u * kThreads;
            if (v < n_vec) {
                xv[u] = ld16_stream(x + v * N);
                if (USE_REF) rv[u] = ld16_stream(ref + v * N);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t v = base + (int64_t)u * kThreads;
            if (v >= n_vec) continue;
            float bval = 0.f;
            int64_t c0 = 0;
            if (BIAS_MODE == 1) {
                int64_t row, ch;
                if (SMALL) {
                    uint32_t r32 = row_div.div((uint32_t)v);
                    ch = r32 - ch_div.div(r32) * ch_div.d;
                } else {
                    row = v / step_v;
                    ch = row % size_b;
                }
                bval = Cvt<T>::to_f(__ldg(bias + ch));
            } else if (BIAS_MODE == 2) {
                c0 = (v * N) % size_b;
            }
            Vec16<T> o;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                float xf = Cvt<T>::to_f(xv[u].v[j]);
                if (BIAS_MODE == 1) xf += bval;
                if (BIAS_MODE == 2) xf += Cvt<T>::to_f(__ldg(bias + c0 + j));
                const float rf = USE_REF ? Cvt<T>::to_f(rv[u].v[j]) : 0.f;
                o.v[j] = Cvt<T>::from_f(act_apply<ACT, GRAD>(xf, rf, alpha) * scale);
            }
            st16(out + v * N, o);
        }
    }
}

// Scalar fallback for shapes the vector path cannot take (odd row lengths, unaligned pointers).
template <typename T, int ACT, int GRAD>
__global__ void __launch_bounds__(kThreads)
bias_act_scalar_kernel(T *__restrict__ out, const T *__restrict__ x, const T *__restrict__ bias,
                       const T *__restrict__ ref, int64_t n, int64_t step_b, int64_t size_b,
                       float alpha, float scale) {
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        float xf = Cvt<T>::to_f(x[i]);
        if (bias) xf += Cvt<T>::to_f(__ldg(bias + (i / step_b) % size_b));
        const float rf = ref ? Cvt<T>::to_f(ref[i]) : 0.f;
        out[i] = Cvt<T>::from_f(act_apply<ACT, GRAD>(xf, rf, alpha) * scale);
    }
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename T, int ACT, int GRAD>
static int launch_bias_act(void *out, const void *x, const void *bias, const void *ref, int64_t n,
                           int64_t step_b, int64_t size_b, float alpha, float scale,
                           cudaStream_t st) {
    constexpr int N = Vec16<T>::N;
    T *o = (T *)out;
    const T *xi = (const T *)x, *bi = (const T *)bias, *ri = (const T *)ref;
    const int sms = sm_count();
    bool vec_ok = (n % N == 0) && aligned16(out) && aligned16(x) && (!ref || aligned16(ref));
    int mode = 0;
    if (bias) {
        if (step_b % N == 0) mode = 1;
        else if (step_b == 1 && size_b % N == 0) mode = 2;
        else vec_ok = false;
    }
    if (!vec_ok) {
        int64_t blocks = ceil_div64(n, kThreads);
        if (blocks > (int64_t)sms * 32) blocks = (int64_t)sms * 32;
        bias_act_scalar_kernel<T, ACT, GRAD><<<(unsigned)blocks, kThreads, 0, st>>>(
            o, xi, bi, ri, n, step_b, size_b, alpha, scale);
        SG2_LAUNCH_CHECK();
        return SG2_OK;
    }
    const int64_t n_vec = n / N;
    const int64_t step_v = mode == 1 ? step_b / N : 1;
    const bool small = n_vec < (1ll << 31) && step_v < (1ll << 31) && size_b < (1ll << 31);
    FastDiv rd((uint32_t)(small ? step_v : 1)), cd((uint32_t)(small ? size_b : 1));
    int64_t blocks = ceil_div64(n_vec, (int64_t)kThreads * kUnroll);
    if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;  // grid-stride beyond 16 CTAs/SM
#define SG2_BA_LAUNCH(MODE, REF, SMALL)                                                        \
    bias_act_vec_kernel<T, ACT, GRAD, MODE, REF, SMALL><<<(unsigned)blocks, kThreads, 0, st>>>( \
        o, xi, bi, ri, n_vec, rd, cd, step_v, size_b, alpha, scale)
    const bool use_ref = ref != nullptr && GRAD == 1;
    if (mode == 0) { if (use_ref) SG2_BA_LAUNCH(0, true, true); else SG2_BA_LAUNCH(0, false, true); }
    else if (mode == 1) {
        if (small) { if (use_ref) SG2_BA_LAUNCH(1, true, true); else SG2_BA_LAUNCH(1, false, true); }
        else { if (use_ref) SG2_BA_LAUNCH(1, true, false); else SG2_BA_LAUNCH(1, false, false); }
    } else { if (use_ref) SG2_BA_LAUNCH(2, true, true); else SG2_BA_LAUNCH(2, false, true); }
#undef SG2_BA_LAUNCH
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

// ---- noise injection + bias + activation (StyledConv tail) --------------------------------------
// x [B,C,HW]; noise [B or 1, HW] (same dtype); one thread per 16-byte vector of one (b,c) row.
template <typename T, int ACT>
__global__ void __launch_bounds__(kThreads)
noise_bias_act_kernel(T *__restrict__ out, const T *__restrict__ x, const T *__restrict__ noise,
                      int64_t noise_bstride, const T *__restrict__ noise_weight,
                      const T *__restrict__ bias, int C, int64_t hw_v, int64_t rows, float alpha,
                      float scale) {
    constexpr int N = Vec16<T>::N;
    const float nw = noise ? Cvt<T>::to_f(__ldg(noise_weight)) : 0.f;
    // grid: x over vectors of a row, y over rows (grid-stride on both)
    for (int64_t row = blockIdx.y; row < rows; row += gridDim.y) {
        const int64_t b = row / C;
        const int c = (int)(row - b * C);
        const float bval = bias ? Cvt<T>::to_f(__ldg(bias + c)) : 0.f;
        const T *xr = x + row * hw_v * N;
        T *orow = out + row * hw_v * N;
        const T *nr = noise ? noise + b * noise_bstride : nullptr;
        for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < hw_v;
             v += (int64_t)gridDim.x * kThreads) {
            Vec16<T> xv = ld16_stream(xr + v * N), nv, o;
            if (noise) nv = ld16(nr + v * N);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                float f = Cvt<T>::to_f(xv.v[j]);
                if (noise) f = f + nw * Cvt<T>::to_f(nv.v[j]);
                f += bval;
                o.v[j] = Cvt<T>::from_f(act_apply<ACT, 0>(f, 0.f, alpha) * scale);
            }
            st16(orow + v * N, o);
        }
    }
}

template <typename T, int ACT>
__global__ void __launch_bounds__(kThreads)
noise_bias_act_scalar_kernel(T *__restrict__ out, const T *__restrict__ x,
                             const T *__restrict__ noise, int64_t noise_bstride,
                             const T *__restrict__ noise_weight, const T *__restrict__ bias, int C,
                             int64_t HW, int64_t n, float alpha, float scale) {
    const float nw = noise ? Cvt<T>::to_f(__ldg(noise_weight)) : 0.f;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kThreads) {
        const int64_t row = i / HW, p = i - row * HW, b = row / C;
        const int c = (int)(row - b * C);
        float f = Cvt<T>::to_f(x[i]);
        if (noise) f = f + nw * Cvt<T>::to_f(noise[b * noise_bstride + p]);
        if (bias) f += Cvt<T>::to_f(__ldg(bias + c));
        out[i] = Cvt<T>::from_f(act_apply<ACT, 0>(f, 0.f, alpha) * scale);
    }
}

// ---- grad_bias[c] = sum_{b,hw} g[b,c,hw] -----------------------------------------------------------
// grid (C, S): block (c, s) reduces its slice of the B*HW elements of channel c and adds the
// partial to the fp32 output with one atomic (S == 1 -> plain store, bit-deterministic).
template <typename T>
__global__ void __launch_bounds__(kThreads)
grad_bias_kernel(float *__restrict__ gb, const T *__restrict__ g, int64_t B, int64_t C, int64_t HW,
                 int S) {
    const int c = blockIdx.x;
    const int64_t per = B * HW;
    const int64_t chunk = (per + S - 1) / S;
    const int64_t lo = (int64_t)blockIdx.y * chunk;
    const int64_t hi = lo + chunk < per ? lo + chunk : per;
    float acc = 0.f;
    for (int64_t i = lo + threadIdx.x; i < hi; i += kThreads) {
        const int64_t b = i / HW, p = i - b * HW;
        acc += Cvt<T>::to_f(g[(b * C + c) * HW + p]);
    }
    __shared__ float part[kThreads / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < kThreads / 32 ? part[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) {
            if (S == 1) gb[c] = v; else atomicAdd(gb + c, v);
        }
    }
}

}  // namespace sg2

using namespace sg2;

extern "C" int sg2_fused_bias_act(void *out, const void *x, const void *bias, const void *ref,
                                  int64_t n, int64_t step_b, int64_t size_b, int act, int grad,
                                  float alpha, float scale, int dtype, sg2_stream_t stream) {
    SG2_REQUIRE(n >= 0, SG2_ERR_BAD_ARG, "fused_bias_act: negative element count");
    if (n == 0) return SG2_OK;
    SG2_REQUIRE(out && x, SG2_ERR_BAD_ARG, "fused_bias_act: null tensor pointer");
    SG2_REQUIRE(!bias || (step_b >= 1 && size_b >= 1), SG2_ERR_BAD_ARG,
                "fused_bias_act: step_b/size_b must be >= 1 when a bias is given");
    SG2_REQUIRE(act == 1 || act == 3, SG2_ERR_UNSUPPORTED,
                "fused_bias_act: act must be 1 (linear) or 3 (lrelu), got %d", act);
    SG2_REQUIRE(grad >= 0 && grad <= 2, SG2_ERR_BAD_ARG, "fused_bias_act: grad must be 0..2");
    SG2_REQUIRE(grad != 1 || act != 3 || ref, SG2_ERR_BAD_ARG,
                "fused_bias_act: grad=1 needs the saved output as ref");
    cudaStream_t st = as_stream(stream);
    SG2_DISPATCH_DTYPE(dtype, {
        const int code = act * 10 + grad;
        switch (code) {
            case 10: case 11: return launch_bias_act<T, 1, 0>(out, x, bias, ref, n, step_b, size_b, alpha, scale, st);
            case 12: return launch_bias_act<T, 1, 2>(out, x, bias, ref, n, step_b, size_b, alpha, scale, st);
            case 30: return launch_bias_act<T, 3, 0>(out, x, bias, ref, n, step_b, size_b, alpha, scale, st);
            case 31: return launch_bias_act<T, 3, 1>(out, x, bias, ref, n, step_b, size_b, alpha, scale, st);
            default: return launch_bias_act<T, 3, 2>(out, x, bias, ref, n, step_b, size_b, alpha, scale, st);
        }
    });
    return SG2_OK;
}

extern "C" int sg2_noise_bias_act(void *out, const void *x, const void *noise,
                                  int64_t noise_bstride, const void *noise_weight, const void *bias,
                                  int64_t B, int C, int64_t HW, int act, float alpha,
                                  float act_scale, int dtype, sg2_stream_t stream) {
    SG2_REQUIRE(B >= 0 && C >= 1 && HW >= 1, SG2_ERR_BAD_ARG, "noise_bias_act: bad shape");
    if (B == 0) return SG2_OK;
    SG2_REQUIRE(out && x, SG2_ERR_BAD_ARG, "noise_bias_act: null tensor pointer");
    SG2_REQUIRE(!noise || noise_weight, SG2_ERR_BAD_ARG, "noise_bias_act: noise without weight");
    SG2_REQUIRE(act == 1 || act == 3, SG2_ERR_UNSUPPORTED, "noise_bias_act: act must be 1 or 3");
    cudaStream_t st = as_stream(stream);
    const int sms = sm_count();
    SG2_DISPATCH_DTYPE(dtype, {
        constexpr int N = Vec16<T>::N;
        const bool vec_ok = HW % N == 0 && aligned16(out) && aligned16(x) &&
                            (!noise || (aligned16(noise) && noise_bstride % N == 0));
        if (vec_ok) {
            const int64_t hw_v = HW / N, rows = B * C;
            unsigned gx = (unsigned)std::min<int64_t>(ceil_div64(hw_v, kThreads), 1024);
            unsigned gy = (unsigned)std::min<int64_t>(rows, std::max<int64_t>(1, (int64_t)sms * 16 / gx));
            dim3 grid(gx, gy);
            if (act == 3)
                noise_bias_act_kernel<T, 3><<<grid, kThreads, 0, st>>>((T *)out, (const T *)x, (const T *)noise, noise_bstride, (const T *)noise_weight, (const T *)bias, C, hw_v, rows, alpha, act_scale);
            else
                noise_bias_act_kernel<T, 1><<<grid, kThreads, 0, st>>>((T *)out, (const T *)x, (const T *)noise, noise_bstride, (const T *)noise_weight, (const T *)bias, C, hw_v, rows, alpha, act_scale);
        } else {
            const int64_t n = B * C * HW;
            unsigned blocks = (unsigned)std::min<int64_t>(ceil_div64(n, kThreads), (int64_t)sms * 32);
            if (act == 3)
                noise_bias_act_scalar_kernel<T, 3><<<blocks, kThreads, 0, st>>>((T *)out, (const T *)x, (const T *)noise, noise_bstride, (const T *)noise_weight, (const T *)bias, C, HW, n, alpha, act_scale);
            else
                noise_bias_act_scalar_kernel<T, 1><<<blocks, kThreads, 0, st>>>((T *)out, (const T *)x, (const T *)noise, noise_bstride, (const T *)noise_weight, (const T *)bias, C, HW, n, alpha, act_scale);
        }
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}

extern "C" int sg2_bias_act_grad_bias(void *grad_bias, const void *grad_input, int64_t B, int64_t C,
                                      int64_t HW, int dtype, sg2_stream_t stream) {
    SG2_REQUIRE(B >= 1 && C >= 1 && HW >= 1, SG2_ERR_BAD_ARG, "grad_bias: bad shape");
    SG2_REQUIRE(grad_bias && grad_input, SG2_ERR_BAD_ARG, "grad_bias: null tensor pointer");
    SG2_REQUIRE(C <= 0x7fffffff, SG2_ERR_UNSUPPORTED, "grad_bias: too many channels");
    cudaStream_t st = as_stream(stream);
    const int sms = sm_count();
    // enough CTAs to fill the chip twice, but never slices shorter than 4 passes of the block
    int64_t S = 1;
    if (C < 2 * sms) S = std::min<int64_t>(ceil_div64(2 * sms, C), std::max<int64_t>(1, B * HW / (kThreads * 4)));
    if (S > 65535) S = 65535;
    if (S > 1) SG2_CUDA_OK(cudaMemsetAsync(grad_bias, 0, sizeof(float) * C, st));
    SG2_DISPATCH_DTYPE(dtype, {
        grad_bias_kernel<T><<<dim3((unsigned)C, (unsigned)S), kThreads, 0, st>>>(
            (float *)grad_bias, (const T *)grad_input, B, C, HW, (int)S);
        SG2_LAUNCH_CHECK();
    });
    return SG2_OK;
}

// host-side self check of the multiply-shift division used by the vector kernels (CPU test hook)
extern "C" int sg2_selftest_fastdiv(uint32_t d, uint32_t n) {
    sg2::FastDiv f(d);
    return f.div(n) == n / d ? 0 : 1;
}
