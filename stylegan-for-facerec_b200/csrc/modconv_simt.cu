// modconv_simt.cu -- exact (fp32 FFMA) modulated convolution, NCHW, any storage dtype.
//
// This is the fp32-parity path of ModulatedConv2d.forward (model.py:232-273): tcgen05 has no
// fp32 input kind, so the <=1e-3 max-abs gate against the reference's fp32 output is met with
// CUDA-core FFMA; the bf16 tensor-core engine lives in synth_*.cu.  Formulation (shared with
// the engine): the weight operand is shared across the batch,
//     out[b,co,p] = demod[b,co] * sum_{ci,t} wt[ci,t,co] * (style[b,ci] * x[b,ci,p+t])
// i.e. activations are modulated while they are staged into shared memory and demodulation is
// a per-(b,co) scale in the epilogue -- no [B,Cout,Cin,k,k] weights, no groups=B.
//
// Implicit GEMM: M = Cout, N = (sample, output pixel) flattened over the whole batch so that the
// 4x4 / 8x8 layers still fill tiles, K = Cin x taps.  The three geometries (stride-1 "same" conv,
// stride-2 transposed conv as four polyphase sub-convolutions that never multiply inserted zeros,
// stride-2 conv) differ only in a small tap table.
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace sg2 {

constexpr int BM = 64, BN = 64, KC = 8, MAXT = 9;

struct ConvGeo {
    int B, Cin, Cout, H, W, OH, OW, kk;   // kk = k*k taps in the weight tensor
    int PH, PW;                            // phase-local output extent
    int in_sy, in_sx;                      // input step per phase-local output step
    int out_sy, out_sx, out_oy, out_ox;    // output pixel = local * out_s + out_o
    int ntaps;
    int dy[MAXT], dx[MAXT], widx[MAXT];    // input offset and weight tap index per active tap
};

template <typename T>
__global__ void __launch_bounds__(256)
modconv_simt_kernel(T *__restrict__ out, const T *__restrict__ x, const float *__restrict__ wt,
                    const float *__restrict__ style, const float *__restrict__ demod, ConvGeo g) {
    __shared__ __align__(16) float Ws[KC * MAXT][BM];
    __shared__ __align__(16) float Xs[KC * MAXT][BN];

    __shared__ int s_dy[MAXT], s_dx[MAXT], s_wi[MAXT];   // tap table (dynamic indexing of params spills)

    const int tid = threadIdx.x;
    if (tid < g.ntaps) { s_dy[tid] = g.dy[tid]; s_dx[tid] = g.dx[tid]; s_wi[tid] = g.widx[tid]; }
    const int tx = tid & 15, ty = tid >> 4;
    const int co0 = blockIdx.y * BM;
    const int64_t n0 = (int64_t)blockIdx.x * BN;
    const int P = g.PH * g.PW;
    const int64_t Ntot = (int64_t)g.B * P;

    // staging role: this thread always fills column ln of Xs / Ws, rows lk, lk+4, ...
    const int ln = tid & 63, lk = tid >> 6;
    const int64_t n_st = n0 + ln;
    const bool n_ok = n_st < Ntot;
    int sb = 0, soy = 0, sox = 0;
    if (n_ok) {
        sb = (int)(n_st / P);
        const int r = (int)(n_st - (int64_t)sb * P);
        soy = r / g.PW;
        sox = r - soy * g.PW;
    }
    const int iy_base = soy * g.in_sy, ix_base = sox * g.in_sx;
    const T *xb = x + (int64_t)sb * g.Cin * g.H * g.W;
    const float *sty = style ? style + (int64_t)sb * g.Cin : nullptr;
    const int co_st = co0 + ln;
    const bool co_ok = co_st < g.Cout;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int ci0 = 0; ci0 < g.Cin; ci0 += KC) {
        const int nk = KC * g.ntaps;
        __syncthreads();
        for (int kk = lk; kk < nk; kk += 4) {
            const int cil = kk / g.ntaps, t = kk - cil * g.ntaps;
            const int ci = ci0 + cil;
            float xv = 0.f, wv = 0.f;
            if (ci < g.Cin) {
                if (co_ok) wv = __ldg(wt + ((int64_t)ci * g.kk + s_wi[t]) * g.Cout + co_st);
                const int iy = iy_base + s_dy[t], ix = ix_base + s_dx[t];
                if (n_ok && iy >= 0 && ix >= 0 && iy < g.H && ix < g.W)
                    xv = Cvt<T>::to_f(xb[((int64_t)ci * g.H + iy) * g.W + ix]) * (sty ? __ldg(sty + ci) : 1.f);
            }
            Ws[kk][ln] = wv;
            Xs[kk][ln] = xv;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < nk; ++kk) {
            const float4 a = *reinterpret_cast<const float4 *>(&Ws[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Xs[kk][tx * 4]);
            acc[0][0] += a.x * b.x; acc[0][1] += a.x * b.y; acc[0][2] += a.x * b.z; acc[0][3] += a.x * b.w;
            acc[1][0] += a.y * b.x; acc[1][1] += a.y * b.y; acc[1][2] += a.y * b.z; acc[1][3] += a.y * b.w;
            acc[2][0] += a.z * b.x; acc[2][1] += a.z * b.y; acc[2][2] += a.z * b.z; acc[2][3] += a.z * b.w;
            acc[3][0] += a.w * b.x; acc[3][1] += a.w * b.y; acc[3][2] += a.w * b.z; acc[3][3] += a.w * b.w;
        }
    }

#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t n = n0 + tx * 4 + j;
        if (n >= Ntot) continue;
        const int b = (int)(n / P);
        const int r = (int)(n - (int64_t)b * P);
        const int ly = r / g.PW, lx = r - ly * g.PW;
        const int oy = ly * g.out_sy + g.out_oy, ox = lx * g.out_sx + g.out_ox;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int co = co0 + ty * 4 + i;
            if (co >= g.Cout) continue;
            float v = acc[i][j];
            if (demod) v *= __ldg(demod + (int64_t)b * g.Cout + co);
            out[(((int64_t)b * g.Cout + co) * g.OH + oy) * g.OW + ox] = Cvt<T>::from_f(v);
        }
    }
}


// ---- the same implicit GEMM with a 128-wide register-blocked tile (every layer with Cout a multiple of 32) -----------------
// 256 threads as (BM / 8) x (BN / 8), an 8 x 8 block of outputs per thread (two 4-wide halves BM/2 resp. BN/2 apart: 16-byte
// conflict-free operand reads), BM x BN = 128 x 128, 64 x 256 or 32 x 512 so that narrow layers keep every thread busy.
// Per K chunk (KC input channels x taps) a thread runs 64 FMAs per 4 LDS.128, against 16 per 2 in the 64 x 64 kernel above, and
// the staging gathers one (sample, pixel) column per thread with its tap geometry resolved once per tap, not once per
// element.  Same summation order inside a chunk (channel-major, then taps) as the small kernel: results agree to rounding.
template <int BM_, int BN_, int KC_>
struct BigTile {
    static constexpr int BM = BM_, BN = BN_, KC = KC_;
    static constexpr int TM = BM / 8, TN = BN / 8;                  // threads along M / N
    static_assert(TM * TN == 256, "256 threads");
    static constexpr int CPT = BN > 256 ? BN / 256 : 1;             // columns staged per thread
    static constexpr int CS = BN < 256 ? 256 / BN : 1;              // threads that share a column (they split the chunk's channels)
    static constexpr size_t smem = (size_t)KC * MAXT * (BM + BN) * sizeof(float);
};

template <typename T, typename CFG>
__global__ void __launch_bounds__(256, 2)
modconv_simt_big_kernel(T *__restrict__ out, const T *__restrict__ x, const float *__restrict__ wt,
                        const float *__restrict__ style, const float *__restrict__ demod, ConvGeo g) {
    constexpr int BM = CFG::BM, BN = CFG::BN, KC = CFG::KC, CPT = CFG::CPT, CS = CFG::CS;
    extern __shared__ __align__(16) float big_smem[];
    float (*Ws)[BM] = reinterpret_cast<float (*)[BM]>(big_smem);                       // [KC * ntaps][BM]
    float (*Xs)[BN] = reinterpret_cast<float (*)[BN]>(big_smem + KC * MAXT * BM);      // [KC * ntaps][BN]
    __shared__ int s_dy[MAXT], s_dx[MAXT], s_wi[MAXT];
    __shared__ int s_woff[KC * MAXT];     // weight offset of k-row kk = cil * ntaps + t inside a chunk: (cil * kk_total + widx[t]) * Cout

    const int tid = threadIdx.x;
    if (tid < g.ntaps) { s_dy[tid] = g.dy[tid]; s_dx[tid] = g.dx[tid]; s_wi[tid] = g.widx[tid]; }
    if (tid < KC * g.ntaps) {
        const int cil = tid / g.ntaps, t = tid - cil * g.ntaps;
        s_woff[tid] = (cil * g.kk + g.widx[t]) * g.Cout;
    }
    const int tx = tid % CFG::TN, ty = tid / CFG::TN;
    const int co0 = blockIdx.y * BM;
    const int64_t n0 = (int64_t)blockIdx.x * BN;
    const int P = g.PH * g.PW;
    const int64_t Ntot = (int64_t)g.B * P;
    const int ntaps = g.ntaps, nk = KC * ntaps;
    const int64_t HW = (int64_t)g.H * g.W;

    // staging role: CPT columns of Xs per thread, channels cs, cs + CS, ... of the chunk
    const int cs = CS > 1 ? tid / BN : 0;
    int iy_base[CPT], ix_base[CPT];
    const T *xb[CPT];
    const float *sty[CPT];
    bool n_ok[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int64_t n_st = n0 + (tid % BN) + j * 256;
        n_ok[j] = n_st < Ntot;
        int sb = 0, soy = 0, sox = 0;
        if (n_ok[j]) {
            sb = (int)(n_st / P);
            const int r = (int)(n_st - (int64_t)sb * P);
            soy = r / g.PW;
            sox = r - soy * g.PW;
        }
        iy_base[j] = soy * g.in_sy; ix_base[j] = sox * g.in_sx;
        xb[j] = x + (int64_t)sb * g.Cin * HW;
        sty[j] = style ? style + (int64_t)sb * g.Cin : nullptr;
    }

    // accumulators as pairs of adjacent columns: one fma.rn.f32x2 (weight value as the broadcast scalar, a pair of
    // activations straight out of the 16-byte shared load) does the work of two FFMAs -- half the issue slots for the same
    // fused multiply-adds in the same order (the kernel issued 72 instructions per 64 FMAs: issue-bound before the FP32 pipe)
    float2 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);
    __syncthreads();                       // tap table

    for (int ci0 = 0; ci0 < g.Cin; ci0 += KC) {
        __syncthreads();                   // the previous chunk has been consumed
        // weights: rows kk = cil * ntaps + t, BM consecutive output channels each (coalesced)
        {
            // (the launcher guarantees Cin % KC == 0 and Cout % BM == 0: no bounds checks; 32-bit offsets: the tensor is < 2^31 elements)
            const float *wc = wt + (int64_t)ci0 * g.kk * g.Cout + co0 + (tid % BM);
#pragma unroll 4
            for (int kk = tid / BM; kk < nk; kk += 256 / BM) Ws[kk][tid % BM] = __ldg(wc + s_woff[kk]);
        }
        // activations: one (sample, pixel) column per thread; the tap's input position is resolved once, then the chunk's channels
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int col = (tid % BN) + j * 256;
            float sv[KC];
#pragma unroll
            for (int c = 0; c < KC; ++c) sv[c] = 1.f;
            if (sty[j]) {
#pragma unroll
                for (int c = cs; c < KC; c += CS) sv[c] = ci0 + c < g.Cin ? __ldg(sty[j] + ci0 + c) : 0.f;
            }
            for (int t = 0; t < ntaps; ++t) {
                const int iy = iy_base[j] + s_dy[t], ix = ix_base[j] + s_dx[t];
                const bool ok = n_ok[j] && iy >= 0 && ix >= 0 && iy < g.H && ix < g.W;
                const T *src = xb[j] + ((int64_t)ci0 * g.H + iy) * g.W + ix;
#pragma unroll
                for (int c = cs; c < KC; c += CS) {
                    float v = 0.f;
                    if (ok && ci0 + c < g.Cin) v = Cvt<T>::to_f(src[c * HW]) * sv[c];
                    Xs[c * ntaps + t][col] = v;
                }
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < nk; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&Ws[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&Ws[kk][BM / 2 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Xs[kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Xs[kk][BN / 2 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(make_float2(a[i], a[i]), make_float2(b[2 * j], b[2 * j + 1]), acc[i][j]);
        }
    }

#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int64_t n = n0 + (j < 4 ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4));
        if (n >= Ntot) continue;
        const int b = (int)(n / P);
        const int r = (int)(n - (int64_t)b * P);
        const int ly = r / g.PW, lx = r - ly * g.PW;
        const int oy = ly * g.out_sy + g.out_oy, ox = lx * g.out_sx + g.out_ox;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int co = co0 + (i < 4 ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4));
            if (co >= g.Cout) continue;
            float v = (j & 1) ? acc[i][j >> 1].y : acc[i][j >> 1].x;
            if (demod) v *= __ldg(demod + (int64_t)b * g.Cout + co);
            out[(((int64_t)b * g.Cout + co) * g.OH + oy) * g.OW + ox] = Cvt<T>::from_f(v);
        }
    }
}

// ---- the big tile, software-pipelined (fp32 storage; tiles that lie inside one sample) -------------------------------------
// ncu of the kernel above: 35 % of the stall samples sit on the first use of the staged loads (every warp loads, waits, then
// computes; 16 warps per SM do not cover it).  Here chunk c + 1 travels by cp.async (no registers held by loads in flight)
// into the other half of a double buffer while chunk c is multiplied: weights as 16-byte copies of four consecutive output
// channels, activations as 4-byte copies whose zero-fill form IS the convolution's zero padding.  cp.async cannot modulate
// on the way, so the style is multiplied into the WEIGHT tile once it has landed (one pass over Ws per chunk) -- legal because
// all BN columns of a tile belong to one sample (tiles are dealt per sample; planes of at least BN pixels: every layer from
// 16 x 16 up, including the (r+1)^2 polyphase planes of the transposed conv).
__device__ __forceinline__ void cpa16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa4_zfill(void *dst, const void *src, bool valid) {
    const int n = valid ? 4 : 0;           // src-size 0: four zero bytes are written, the source is not read
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(n) : "memory");
}

template <typename CFG>
__global__ void __launch_bounds__(256, 2)
modconv_simt_pipe_kernel(float *__restrict__ out, const float *__restrict__ x, const float *__restrict__ wt,
                         const float *__restrict__ style, const float *__restrict__ demod, ConvGeo g) {
    constexpr int BM = CFG::BM, BN = CFG::BN, KC = CFG::KC, CPT = CFG::CPT, CS = CFG::CS;
    constexpr int BUF = KC * MAXT * (BM + BN);                                      // floats per stage
    extern __shared__ __align__(16) float big_smem[];
    __shared__ int s_dy[MAXT], s_dx[MAXT];
    __shared__ int s_woff[KC * MAXT];
    __shared__ int s_kcil[KC * MAXT];      // channel (inside the chunk) of k-row kk

    const int tid = threadIdx.x;
    const int ntaps = g.ntaps, nk = KC * ntaps;
    if (tid < ntaps) { s_dy[tid] = g.dy[tid]; s_dx[tid] = g.dx[tid]; }
    if (tid < nk) {
        const int cil = tid / ntaps, t = tid - cil * ntaps;
        s_woff[tid] = (cil * g.kk + g.widx[t]) * g.Cout;
        s_kcil[tid] = cil;
    }
    const int tx = tid % CFG::TN, ty = tid / CFG::TN;
    const int co0 = blockIdx.y * BM;
    const int P = g.PH * g.PW;
    const int64_t HW = (int64_t)g.H * g.W;
    // tiles never straddle samples: ceil(P / BN) tiles per sample, the last one partly empty (columns >= P are zero-filled
    // and not stored) -- this is what lets the modulation sit on the weight tile
    const int tps = (P + BN - 1) / BN;
    const int sb = (int)blockIdx.x / tps;                  // the tile's sample
    const int p0 = ((int)blockIdx.x - sb * tps) * BN;
    const float *xs = x + (int64_t)sb * g.Cin * HW;
    const float *sty = style ? style + (int64_t)sb * g.Cin : nullptr;

    // staging role: CPT columns of Xs per thread, channels cs, cs + CS, ... of the chunk
    const int cs = CS > 1 ? tid / BN : 0;
    int iy_base[CPT], ix_base[CPT];
    bool col_ok[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int r = p0 + (tid % BN) + j * 256;
        col_ok[j] = r < P;
        const int soy = r / g.PW, sox = r - soy * g.PW;
        iy_base[j] = soy * g.in_sy; ix_base[j] = sox * g.in_sx;
    }
    float *s_sty = big_smem + 2 * BUF;     // the tile's style vector (one sample per tile), staged once
    if (sty)
        for (int i = tid; i < g.Cin; i += 256) s_sty[i] = __ldg(sty + i);
    __syncthreads();                       // tables

    auto issue = [&](int ci0, int buf) {
        float (*Ws)[BM] = reinterpret_cast<float (*)[BM]>(big_smem + buf * BUF);
        float (*Xs)[BN] = reinterpret_cast<float (*)[BN]>(big_smem + buf * BUF + KC * MAXT * BM);
        // weights: four consecutive output channels per copy
        const float *wc = wt + (int64_t)ci0 * g.kk * g.Cout + co0;
        for (int i = tid; i < nk * (BM / 4); i += 256) {
            const int kk = i / (BM / 4), c4 = (i - kk * (BM / 4)) * 4;
            cpa16(&Ws[kk][c4], wc + s_woff[kk] + c4);
        }
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int col = (tid % BN) + j * 256;
            for (int t = 0; t < ntaps; ++t) {
                const int iy = iy_base[j] + s_dy[t], ix = ix_base[j] + s_dx[t];
                const bool ok = col_ok[j] && iy >= 0 && ix >= 0 && iy < g.H && ix < g.W;
                const float *src = xs + ((int64_t)ci0 * g.H + (ok ? iy : 0)) * g.W + (ok ? ix : 0);
#pragma unroll
                for (int c = cs; c < KC; c += CS) cpa4_zfill(&Xs[c * ntaps + t][col], src + c * HW, ok);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // accumulators as pairs of adjacent columns: one fma.rn.f32x2 (weight value as the broadcast scalar, a pair of
    // activations straight out of the 16-byte shared load) does the work of two FFMAs -- half the issue slots for the same
    // fused multiply-adds in the same order (the kernel issued 72 instructions per 64 FMAs: issue-bound before the FP32 pipe)
    float2 acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);

    const int nchunks = g.Cin / KC;
    issue(0, 0);
    for (int c = 0; c < nchunks; ++c) {
        const int buf = c & 1;
        if (c + 1 < nchunks) {
            issue((c + 1) * KC, buf ^ 1);                 // (its previous contents were consumed before the barrier that ended chunk c - 1)
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();                                   // every thread's copies of chunk c are visible
        float (*Ws)[BM] = reinterpret_cast<float (*)[BM]>(big_smem + buf * BUF);
        float (*Xs)[BN] = reinterpret_cast<float (*)[BN]>(big_smem + buf * BUF + KC * MAXT * BM);
        if (sty) {                                         // modulation, applied to the weight rows of this chunk
            for (int i = tid; i < nk * (BM / 4); i += 256) {
                const int kk = i / (BM / 4), c4 = (i - kk * (BM / 4)) * 4;
                const float sv = s_sty[c * KC + s_kcil[kk]];
                float4 w4 = *reinterpret_cast<float4 *>(&Ws[kk][c4]);
                w4.x *= sv; w4.y *= sv; w4.z *= sv; w4.w *= sv;
                *reinterpret_cast<float4 *>(&Ws[kk][c4]) = w4;
            }
            __syncthreads();
        }
#pragma unroll 4
        for (int kk = 0; kk < nk; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&Ws[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&Ws[kk][BM / 2 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Xs[kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Xs[kk][BN / 2 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(make_float2(a[i], a[i]), make_float2(b[2 * j], b[2 * j + 1]), acc[i][j]);
        }
        __syncthreads();                                   // chunk c consumed: its buffer may be refilled
    }

#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int r = p0 + (j < 4 ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4));
        if (r >= P) continue;
        const int ly = r / g.PW, lx = r - ly * g.PW;
        const int oy = ly * g.out_sy + g.out_oy, ox = lx * g.out_sx + g.out_ox;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int co = co0 + (i < 4 ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4));
            float v = (j & 1) ? acc[i][j >> 1].y : acc[i][j >> 1].x;
            if (demod) v *= __ldg(demod + (int64_t)sb * g.Cout + co);
            out[(((int64_t)sb * g.Cout + co) * g.OH + oy) * g.OW + ox] = v;
        }
    }
}

template <typename CFG>
static int launch_pipe(float *out, const float *x, const float *wt, const float *style, const float *demod, const ConvGeo &g,
                       int64_t Ntot, cudaStream_t st) {
    constexpr int kMaxCin = 2048;                                    // style vector staged in shared memory
    if (g.Cin > kMaxCin) return 1;
    const size_t smem = 2 * CFG::smem + (size_t)g.Cin * sizeof(float);
    static std::atomic<int> configured{0};
    if (!configured.load(std::memory_order_acquire)) {
        SG2_CUDA_OK(cudaFuncSetAttribute(modconv_simt_pipe_kernel<CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(2 * CFG::smem + kMaxCin * sizeof(float))));
        configured.store(1, std::memory_order_release);
    }
    const int P = g.PH * g.PW;
    dim3 grid((unsigned)(g.B * ((P + CFG::BN - 1) / CFG::BN)), g.Cout / CFG::BM);
    modconv_simt_pipe_kernel<CFG><<<grid, 256, smem, st>>>(out, x, wt, style, demod, g);
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

template <typename T, typename CFG>
static int launch_big(T *out, const T *x, const float *wt, const float *style, const float *demod, const ConvGeo &g, int64_t Ntot,
                      cudaStream_t st) {
    static std::atomic<int> configured{0};
    if (!configured.load(std::memory_order_acquire)) {
        SG2_CUDA_OK(cudaFuncSetAttribute(modconv_simt_big_kernel<T, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CFG::smem));
        configured.store(1, std::memory_order_release);
    }
    dim3 grid((unsigned)ceil_div64(Ntot, CFG::BN), (g.Cout + CFG::BM - 1) / CFG::BM);
    modconv_simt_big_kernel<T, CFG><<<grid, 256, CFG::smem, st>>>(out, x, wt, style, demod, g);
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

}  // namespace sg2

using namespace sg2;

extern "C" int sg2_modconv2d_fwd(void *out, const void *x, const float *wt, const float *style,
                                 const float *demod, int64_t B, int Cin, int Cout, int H, int W,
                                 int k, int mode, int dtype, sg2_stream_t stream) {
    SG2_REQUIRE(B >= 0 && Cin >= 1 && Cout >= 1 && H >= 1 && W >= 1, SG2_ERR_BAD_ARG,
                "modconv2d: bad shape");
    SG2_REQUIRE(k == 1 || k == 3, SG2_ERR_UNSUPPORTED, "modconv2d: kernel size must be 1 or 3, got %d", k);
    SG2_REQUIRE(mode >= 0 && mode <= 2, SG2_ERR_BAD_ARG, "modconv2d: mode must be 0, 1 or 2");
    SG2_REQUIRE(mode != 2 || (H >= k && W >= k), SG2_ERR_BAD_ARG, "modconv2d: input smaller than kernel");
    if (B == 0) return SG2_OK;
    SG2_REQUIRE(out && x && wt, SG2_ERR_BAD_ARG, "modconv2d: null tensor pointer");
    SG2_REQUIRE(B <= (1 << 20), SG2_ERR_UNSUPPORTED, "modconv2d: batch too large");
    cudaStream_t st = as_stream(stream);

    ConvGeo g;
    g.B = (int)B; g.Cin = Cin; g.Cout = Cout; g.H = H; g.W = W; g.kk = k * k;
    const int nphase = mode == 1 ? 4 : 1;
    for (int ph = 0; ph < nphase; ++ph) {
        g.ntaps = 0;
        if (mode == 0) {          // out[y,x] = sum_t w[t] * in[y + ty - k/2, x + tx - k/2]
            g.OH = H; g.OW = W; g.PH = H; g.PW = W;
            g.in_sy = g.in_sx = 1; g.out_sy = g.out_sx = 1; g.out_oy = g.out_ox = 0;
            for (int a = 0; a < k; ++a)
                for (int b = 0; b < k; ++b) {
                    g.dy[g.ntaps] = a - k / 2; g.dx[g.ntaps] = b - k / 2; g.widx[g.ntaps] = a * k + b;
                    ++g.ntaps;
                }
        } else if (mode == 2) {   // out[y,x] = sum_t w[t] * in[2y + ty, 2x + tx]
            g.OH = (H - k) / 2 + 1; g.OW = (W - k) / 2 + 1; g.PH = g.OH; g.PW = g.OW;
            g.in_sy = g.in_sx = 2; g.out_sy = g.out_sx = 1; g.out_oy = g.out_ox = 0;
            for (int a = 0; a < k; ++a)
                for (int b = 0; b < k; ++b) {
                    g.dy[g.ntaps] = a; g.dx[g.ntaps] = b; g.widx[g.ntaps] = a * k + b;
                    ++g.ntaps;
                }
        } else {                  // transposed: out[2i+a, 2j+b] += in[i,j] * w[a,b]; phase = parity
            const int py = ph >> 1, px = ph & 1;
            g.OH = (H - 1) * 2 + k; g.OW = (W - 1) * 2 + k;
            g.PH = (g.OH - py + 1) / 2; g.PW = (g.OW - px + 1) / 2;
            if (g.PH <= 0 || g.PW <= 0) continue;
            g.in_sy = g.in_sx = 1; g.out_sy = g.out_sx = 2; g.out_oy = py; g.out_ox = px;
            for (int a = py; a < k; a += 2)
                for (int b = px; b < k; b += 2) {   // out row 2*ly+py takes tap a from input row ly-(a-py)/2
                    g.dy[g.ntaps] = -(a - py) / 2; g.dx[g.ntaps] = -(b - px) / 2; g.widx[g.ntaps] = a * k + b;
                    ++g.ntaps;
                }
        }
        const int64_t Ntot = B * (int64_t)g.PH * g.PW;
        SG2_REQUIRE(ceil_div64(Ntot, BN) <= 0x7fffffffll, SG2_ERR_UNSUPPORTED, "modconv2d: too many pixels");
        dim3 grid((unsigned)ceil_div64(Ntot, BN), (Cout + BM - 1) / BM);
        // register-blocked big tiles where they fill: Cout a multiple of 32 and enough (sample, pixel) columns for a few waves
        static const char *env_big = getenv("SG2_MODCONV_BIG");       // A/B switch: 0 = the 64 x 64 kernel everywhere
        const bool big_ok = (!env_big || atoi(env_big) != 0) && Cout % 32 == 0 && Cin % 8 == 0 && k == 3 &&
                            (int64_t)Cin * 9 * Cout < (1ll << 31);
        // fp32 storage, tiles inside one sample, 16-byte aligned weight rows: the pipelined form (KC = 4 per stage)
        static const char *env_pipe = getenv("SG2_MODCONV_PIPE");     // A/B switch: 0 = off
        const int P = g.PH * g.PW;
        if (big_ok && (!env_pipe || atoi(env_pipe) != 0) && dtype == SG2_F32 && (reinterpret_cast<uintptr_t>(wt) & 15) == 0 &&
            Cin % 4 == 0) {
            int rc = 1;
            if (Cout % 128 == 0 && P >= 128) rc = launch_pipe<BigTile<128, 128, 4>>((float *)out, (const float *)x, wt, style, demod, g, Ntot, st);
            else if (Cout % 64 == 0 && Cout < 128 && P >= 256) rc = launch_pipe<BigTile<64, 256, 4>>((float *)out, (const float *)x, wt, style, demod, g, Ntot, st);
            if (rc < 0 || rc > 1) return rc;
            if (rc == 0) continue;
        }
        SG2_DISPATCH_DTYPE(dtype, {
            int rc = 1;
            if (big_ok && Cout % 128 == 0 && Ntot >= 128 * 8) rc = launch_big<T, BigTile<128, 128, 8>>((T *)out, (const T *)x, wt, style, demod, g, Ntot, st);
            else if (big_ok && Cout % 64 == 0 && Cout < 128 && Ntot >= 256 * 8) rc = launch_big<T, BigTile<64, 256, 8>>((T *)out, (const T *)x, wt, style, demod, g, Ntot, st);
            else if (big_ok && Cout == 32 && Ntot >= 512 * 8) rc = launch_big<T, BigTile<32, 512, 4>>((T *)out, (const T *)x, wt, style, demod, g, Ntot, st);
            if (rc < 0 || rc > 1) return rc;
            if (rc == 1) {
                modconv_simt_kernel<T><<<grid, 256, 0, st>>>((T *)out, (const T *)x, wt, style, demod, g);
                SG2_LAUNCH_CHECK();
            }
        });
    }
    return SG2_OK;
}
