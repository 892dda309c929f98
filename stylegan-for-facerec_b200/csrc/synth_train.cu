// synth_train.cu -- the whole-network engine walked BACKWARD: dL/d(styles) of a frozen decoder.
//
// The ReStyle / pSp coaches back-propagate an image loss through the frozen StyleGAN2 decoder into the encoder
// (coach_restyle_psp.py:138-168); what they need from the decoder is dL/d(latent).  In training mode the forward walk
// (synth.cu) keeps every styled conv's stored output  st_l = lrelu(y_l) * sqrt(2) * s_{l+1}  (NHWC bf16, the same tensor
// the next conv consumes), and this file derives everything else from it:
//
//   per styled conv l, last to first
//     E-pass (one HBM-bound kernel):  a = st / (sqrt(2) s_next);  g_a = g_st * sqrt(2) s_next + sum_k g_rgb[k] * rho[k]
//         reductions  dL/ds_next[b,c] = sum_p g_st * a * sqrt(2),  dL/ds_rgb[b,c] = sum_p a * sqrt(2) * sum_k g_rgb[k] w_rgb[k,c],
//                     dL/dd[b,c]     = sum_p g_y * (y - noise - bias) / d          (y = d * conv + noise + bias)
//         g_y = g_a * lrelu'(a),   out: g_conv = g_y * d  (bf16 NHWC)                [fused_act.py:20-38, model.py:236-240,287]
//     input gradient on the tcgen05 kernel:  plain conv -> 3x3 conv with flipped taps and swapped channels;
//         up-sampling layer -> FIR adjoint into the four polyphase planes, then ONE 9-tap GEMM whose taps read the four
//         planes (GemmParams::multi_map)                                                [autograd of model.py:246-257]
//   the ToRGB skip chain backwards: dL/d(skip_j-1) = Upsample^T(dL/d(skip_j))                      [model.py:350-359]
//   demodulation: dL/ds[b,ci] -= s[b,ci] * sum_co dL/dd[b,co] d^3 wsq[ci,co]                       [model.py:238-240]
// The caller (engine.py) turns dL/ds of every layer into dL/d(latent) with the modulation weights.  Noise maps and weights
// receive no gradient here: this is the frozen-decoder direction; anything else takes the autograd path of tc_route.py.
#include "synth_plan.cuh"

using namespace sg2;
using namespace sg2plan;

namespace {

constexpr float kSqrt2 = 1.41421356237f;
constexpr float kSlopeT = 0.2f;

// wadj[t'][ci][co] = scale * w[co][ci][t],  t = flip ? 8 - t' : t'     (w fp32 [Cout, Cin, 3, 3])
__global__ void __launch_bounds__(256)
pack_adjoint_weight_kernel(__nv_bfloat16 *__restrict__ wadj, const float *__restrict__ w, int Cin, int Cout, float scale, int flip) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;     // over (t', ci, co), co fastest
    if (i >= (int64_t)9 * Cin * Cout) return;
    const int co = (int)(i % Cout);
    const int64_t rest = i / Cout;
    const int ci = (int)(rest % Cin), tp = (int)(rest / Cin);
    const int t = flip ? 8 - tp : tp;
    wadj[i] = __float2bfloat16_rn(scale * w[((int64_t)co * Cin + ci) * 9 + t]);
}

struct EpassParams {
    const __nv_bfloat16 *st;      // [B, HW, C] kept forward output
    const __nv_bfloat16 *g_st;    // [B, HW, C] gradient w.r.t. it, or null (last layer)
    __nv_bfloat16 *g_conv;        // [B, HW, C] out
    int B, HW, C;
    const float *s_next;          // [B, C] style of the consumer conv, or null (1)
    const float *d;               // [B, C] demodulation
    const float *bias;            // [C]
    const float *noise;           // [B or 1, HW] or null
    long long noise_bstride;
    const float *noise_weight;    // [1]
    const float *g_rgb;           // [B, 3, HW] fp32 gradient of the skip image at this resolution, or null
    const float *rgb_w;           // [3, C] ToRGB weights, equalised-lr scale folded
    const float *s_rgb;           // [B, C]
    float *g_s_next;              // [B, C] += , or null
    float *g_s_rgb;               // [B, C] +=
    float *g_d;                   // [B, C] +=
};

__device__ __forceinline__ void unpack8(const uint4 &v, float (&f)[8]) {
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 v;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return v;
}

// grid (blocks per sample, B); 256 threads; a thread owns 8 consecutive channels and walks pixels.  HBM-bound by design
// (2 x 16-byte loads + one 16-byte store per 8 elements): the loads of the NEXT pixel are issued before the math of the
// current one; two blocks per SM (128 registers: no spills).
struct EpassIn { uint4 st, gs; float nz, g0, g1, g2; };
template <bool RGB>
__device__ __forceinline__ EpassIn epass_load(const EpassParams &p, int b, int px, int c0, float nw) {
    EpassIn in;
    const long long e = ((long long)b * p.HW + px) * p.C + c0;
    in.st = __ldg(reinterpret_cast<const uint4 *>(p.st + e));
    in.gs = p.g_st ? __ldg(reinterpret_cast<const uint4 *>(p.g_st + e)) : make_uint4(0, 0, 0, 0);
    in.nz = p.noise ? nw * __ldg(p.noise + (long long)b * p.noise_bstride + px) : 0.f;
    in.g0 = in.g1 = in.g2 = 0.f;
    if (RGB) {
        const float *gr = p.g_rgb + (long long)b * 3 * p.HW + px;
        in.g0 = __ldg(gr); in.g1 = __ldg(gr + p.HW); in.g2 = __ldg(gr + 2 * (long long)p.HW);
    }
    return in;
}

template <bool RGB>      // RGB: this conv feeds a ToRGB (every non-up-sampling layer); the up-sampling layers carry 32 registers less
__global__ void __launch_bounds__(256, 2)
train_epass_kernel(EpassParams p) {
    __shared__ float red[3][512];
    const int b = blockIdx.y;
    const int cpt = p.C >> 3;                   // threads per pixel
    const int ppi = 256 / cpt;                  // pixels per block iteration
    const int cg = threadIdx.x % cpt, pl = threadIdx.x / cpt;
    const int c0 = cg * 8;
    for (int i = threadIdx.x; i < 3 * 512; i += 256) (&red[0][0])[i] = 0.f;
    __syncthreads();
    float sig[8], isig[8], dd[8], bb[8], w0[8], w1[8], w2[8];     // w_k pre-multiplied by sqrt(2) * s_rgb
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        const float sn = p.s_next ? __ldg(p.s_next + (long long)b * p.C + c) : 1.f;
        sig[j] = kSqrt2 * sn;
        isig[j] = fabsf(sig[j]) > 1e-20f ? 1.f / sig[j] : 0.f;
        dd[j] = __ldg(p.d + (long long)b * p.C + c);
        bb[j] = __ldg(p.bias + c);
        w0[j] = w1[j] = w2[j] = 0.f;
        if (RGB) {
            const float sr = kSqrt2 * __ldg(p.s_rgb + (long long)b * p.C + c);
            w0[j] = sr * __ldg(p.rgb_w + c); w1[j] = sr * __ldg(p.rgb_w + p.C + c); w2[j] = sr * __ldg(p.rgb_w + 2 * p.C + c);
        }
    }
    // reductions: rsn = sum g_st * a,  rsr = sum (sum_k g_rgb[k] w'_k) * a,  rd = sum g_y * (y - noise - bias); the constant
    // factors (sqrt(2), 1 / (sqrt(2) s_rgb), 1 / d) are applied once at the end
    float rsn[8], rsr[8], rd[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) rsn[j] = rsr[j] = rd[j] = 0.f;
    const float nw = p.noise ? __ldg(p.noise_weight) : 0.f;
    const int stride = gridDim.x * ppi;
    int px = blockIdx.x * ppi + pl;
    EpassIn cur;
    if (px < p.HW) cur = epass_load<RGB>(p, b, px, c0, nw);
    while (px < p.HW) {
        const int pxn = px + stride;
        EpassIn nxt;
        if (pxn < p.HW) nxt = epass_load<RGB>(p, b, pxn, c0, nw);
        float st[8], gs[8], out[8];
        unpack8(cur.st, st);
        unpack8(cur.gs, gs);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float a = st[j] * isig[j];
            const float grw = RGB ? cur.g0 * w0[j] + cur.g1 * w1[j] + cur.g2 * w2[j] : 0.f;      // sum_k g_rgb[k] * rho[k, c]
            const float ga = fmaf(gs[j], sig[j], grw);
            rsn[j] = fmaf(gs[j], a, rsn[j]);
            if (RGB) rsr[j] = fmaf(grw, a, rsr[j]);
            const bool pos = a > 0.f;
            const float gy = pos ? ga : kSlopeT * ga;
            const float y = pos ? a : a * (1.f / kSlopeT);
            rd[j] = fmaf(gy, y - cur.nz - bb[j], rd[j]);
            out[j] = gy * dd[j];
        }
        *reinterpret_cast<uint4 *>(p.g_conv + ((long long)b * p.HW + px) * p.C + c0) = pack8(out);
        cur = nxt;
        px = pxn;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (p.g_s_next) atomicAdd(&red[0][c0 + j], rsn[j] * kSqrt2);
        if (RGB) {
            // rho = sqrt(2) s_rgb w: dL/ds_rgb = sum a * sqrt(2) * sum_k g_k w_k = rsr / s_rgb (a channel with s_rgb == 0 carries no rgb term)
            const float sr = __ldg(p.s_rgb + (long long)b * p.C + c0 + j);
            atomicAdd(&red[1][c0 + j], fabsf(sr) > 1e-20f ? rsr[j] / sr : 0.f);
        }
        atomicAdd(&red[2][c0 + j], rd[j] / dd[j]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < p.C; c += 256) {
        const long long o = (long long)b * p.C + c;
        if (p.g_s_next) atomicAdd(p.g_s_next + o, red[0][c]);
        if (RGB) atomicAdd(p.g_s_rgb + o, red[1][c]);
        atomicAdd(p.g_d + o, red[2][c]);
    }
}

// adjoint of the 4x4 blur after the transposed conv (pad (1,1)): g [B, 2r, 2r, C] -> four polyphase planes
// gT[(py,px)][B][P][P][C], P = r + 1, gT[u, v] = sum_{a,b} kf[a][b] * g[u - a + 1, v - b + 1]   (u = 2y + py, v = 2x + px).
// A thread owns 8 channels of the 2 x 2 patch (u, v) in {2y, 2y+1} x {2x, 2x+1} -- one pixel of each plane: the patch reads a
// 5 x 5 window of g (25 16-byte loads for 4 stores instead of 64).
struct FirTParams { const __nv_bfloat16 *g; __nv_bfloat16 *planes; int B, r, C; float kf[16]; };
__global__ void __launch_bounds__(256)
train_firT_kernel(FirTParams p) {
    const int P = p.r + 1, R = 2 * p.r, c8n = p.C >> 3;
    const long long plane_elems = (long long)p.B * P * P * p.C;
    const long long total = (long long)p.B * P * P * c8n;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int c8 = (int)(i % c8n);
        long long rest = i / c8n;
        const int x = (int)(rest % P);
        rest /= P;
        const int y = (int)(rest % P), b = (int)(rest / P);
        float acc[4][8];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[q][j] = 0.f;
        // window rows gy = 2y - 2 + wy (wy = 0..4); output row u = 2y + py takes tap a = u + 1 - gy = py + 3 - wy
#pragma unroll
        for (int wy = 0; wy < 5; ++wy) {
            const int gy = 2 * y - 2 + wy;
            if (gy < 0 || gy >= R) continue;
#pragma unroll
            for (int wx = 0; wx < 5; ++wx) {
                const int gx = 2 * x - 2 + wx;
                if (gx < 0 || gx >= R) continue;
                float f[8];
                unpack8(__ldg(reinterpret_cast<const uint4 *>(p.g + (((long long)b * R + gy) * R + gx) * p.C + c8 * 8)), f);
#pragma unroll
                for (int py = 0; py < 2; ++py) {
                    const int a = py + 3 - wy;
                    if (a < 0 || a > 3) continue;
#pragma unroll
                    for (int px = 0; px < 2; ++px) {
                        const int bq = px + 3 - wx;
                        if (bq < 0 || bq > 3) continue;
                        const float k = p.kf[a * 4 + bq];
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[py * 2 + px][j] = fmaf(k, f[j], acc[py * 2 + px][j]);
                    }
                }
            }
        }
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
            const int py = ph >> 1, px = ph & 1;
            if (y >= P - py || x >= P - px) continue;       // the odd planes are one row / column shorter
            *reinterpret_cast<uint4 *>(p.planes + ph * plane_elems + (((long long)b * P + y) * P + x) * p.C + c8 * 8) = pack8(acc[ph]);
        }
    }
}

// adjoint of the skip Upsample (up 2, pad (2,1), 4x4 taps): g_prev[i, j] = sum_{Y, X} kf[(2i+2-Y)*4 + (2j+2-X)] * g[Y, X]
struct RgbUpTParams { const float *g; float *g_prev; int planes, S; float kf[16]; };
__global__ void __launch_bounds__(256)
rgb_up_adjoint_kernel(RgbUpTParams p) {
    const int S = p.S, R = 2 * S;
    const long long total = (long long)p.planes * S * S;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int j = (int)(i % S), ii = (int)((i / S) % S);
        const long long pl = i / ((long long)S * S);
        const float *g = p.g + pl * R * R;
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int Y = 2 * ii + 2 - a;
            if (Y < 0 || Y >= R) continue;
#pragma unroll
            for (int bq = 0; bq < 4; ++bq) {
                const int X = 2 * j + 2 - bq;
                if (X < 0 || X >= R) continue;
                acc = fmaf(p.kf[a * 4 + bq], __ldg(g + (long long)Y * R + X), acc);
            }
        }
        p.g_prev[i] = acc;
    }
}

// conv1's input is const[c, p] * s0[b, c]:  dL/ds0[b, c] += sum_p g_xm0[b, p, c] * const[c, p]   (16 pixels)
__global__ void __launch_bounds__(256)
const_grad_kernel(float *__restrict__ g_s0, const __nv_bfloat16 *__restrict__ g_xm0, const float *__restrict__ cst, int B, int C) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= B * C) return;
    const int b = i / C, c = i - b * C;
    float acc = 0.f;
    for (int px = 0; px < 16; ++px) acc += __bfloat162float(g_xm0[((long long)b * 16 + px) * C + c]) * __ldg(cst + c * 16 + px);
    g_s0[i] += acc;
}

// d = rsqrt(sum_ci s^2 wsq + eps):  dL/ds[b, ci] -= s[b, ci] * sum_co dL/dd[b, co] * d[b, co]^3 * wsq[ci, co]
// grid (ceil(cin / 8), ceil(B / 8), layers); a warp owns one ci, its lanes walk co (coalesced wsq rows), 8 samples at a time
struct DemodGradJob { const float *style, *wsq, *demod, *g_d; float *g_s; int cin, cout; };
struct DemodGradJobs { DemodGradJob job[kMaxJobs]; int n; };
__global__ void __launch_bounds__(256)
demod_grad_kernel(DemodGradJobs jobs, int B) {
    __shared__ float t[8][512];
    const DemodGradJob &j = jobs.job[blockIdx.z];
    const int b0 = blockIdx.y * 8, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8 * j.cout; i += 256) {
        const int bb = i / j.cout, co = i - bb * j.cout, b = b0 + bb;
        float v = 0.f;
        if (b < B) {
            const float dv = __ldg(j.demod + (long long)b * j.cout + co);
            v = __ldg(j.g_d + (long long)b * j.cout + co) * dv * dv * dv;
        }
        t[bb][co] = v;
    }
    __syncthreads();
    const int ci = blockIdx.x * 8 + warp;
    if (ci >= j.cin) return;
    float acc[8];
#pragma unroll
    for (int bb = 0; bb < 8; ++bb) acc[bb] = 0.f;
    const float *wr = j.wsq + (long long)ci * j.cout;
    for (int co = lane; co < j.cout; co += 32) {
        const float w = __ldg(wr + co);
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) acc[bb] = fmaf(w, t[bb][co], acc[bb]);
    }
#pragma unroll
    for (int bb = 0; bb < 8; ++bb) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[bb] += __shfl_xor_sync(0xffffffffu, acc[bb], o);
    }
    if (lane < 8 && b0 + lane < B) {
        float a = acc[0];
#pragma unroll
        for (int bb = 1; bb < 8; ++bb) a = lane == bb ? acc[bb] : a;
        const long long o = (long long)(b0 + lane) * j.cin + ci;
        j.g_s[o] -= __ldg(j.style + o) * a;
    }
}

__global__ void fill_kernel(float *p, float v, long long n) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) p[i] = v;
}

// the four polyphase planes of the output gradient as the A operand of the transposed conv's input-gradient GEMM
int encode_plane_maps(Layer &BL, const __nv_bfloat16 *planes, int B, int r) {
    EncodeTiledFn enc = get_encode();
    SG2_REQUIRE(enc, SG2_ERR_CUDA, "engine: cuTensorMapEncodeTiled entry point not available");
    const int C = BL.p.cin, P = r + 1;
    const GemmSub &q = BL.gp.sub[0];
    const CUtensorMapSwizzle swz = BL.gp.block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const size_t plane = (size_t)B * P * P * C;
    for (int s = 0; s < 4; ++s) {
        const int py = s >> 1, px = s & 1;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)(P - px), (cuuint64_t)(P - py), (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)P * C * 2, (cuuint64_t)P * P * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)BL.gp.block_k, (cuuint32_t)q.TW, (cuuint32_t)q.TH, (cuuint32_t)q.NB};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult rc = enc(&BL.tmA[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)(planes + plane * s), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SG2_REQUIRE(rc == CUDA_SUCCESS, SG2_ERR_CUDA, "engine: cuTensorMapEncodeTiled(gradient plane) failed with %d", (int)rc);
    }
    return SG2_OK;
}

}  // namespace

namespace sg2plan {

int train_pack(sg2_synth *S, uint8_t *ws, cudaStream_t st) {
    for (Layer &L : S->layers) {
        if (L.rgb) continue;
        const int64_t n = (int64_t)9 * L.p.cin * L.p.cout;
        pack_adjoint_weight_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>((__nv_bfloat16 *)(ws + L.wadj), L.p.weight, L.p.cin, L.p.cout,
                                                                                 1.0f / sqrtf((float)L.p.cin * 9), L.p.upsample ? 0 : 1);
        SG2_LAUNCH_CHECK();
    }
    const long long n1 = (long long)S->max_batch * 512;
    fill_kernel<<<64, 256, 0, st>>>((float *)(ws + S->off_ones), 1.0f, n1);
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

}  // namespace sg2plan

// Switch the plan to training mode.  Call right after sg2_synth_create (before sg2_synth_workspace_bytes / sg2_synth_pack):
// the workspace grows by one kept activation per styled conv, the adjoint weight packs and the reduction scratch.
extern "C" int sg2_synth_enable_training(sg2_synth *S) {
    SG2_REQUIRE(S, SG2_ERR_BAD_ARG, "synth_enable_training: null plan");
    if (S->train) return SG2_OK;
    SG2_REQUIRE(!S->ada, SG2_ERR_UNSUPPORTED, "synth_enable_training: the stylegan2_ada plan has no backward walk (use the autograd route)");
    const int B = S->max_batch;
    size_t off = S->ws_bytes;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
    size_t max_T = 0;
    int ci = 0;
    S->off_act_in = take(sizeof(__nv_bfloat16) * (size_t)B * 16 * S->layers[0].p.cin);
    for (Layer &L : S->layers) {
        if (L.rgb) continue;
        SG2_REQUIRE(L.p.cout % 8 == 0 && 256 % (L.p.cout / 8) == 0 && L.p.cout <= 512 && L.p.cin <= 512, SG2_ERR_UNSUPPORTED,
                    "engine (training): channel counts must be powers of two in [8, 512], got %d -> %d", L.p.cin, L.p.cout);
        L.conv_index = ci++;
        L.keep = take(sizeof(__nv_bfloat16) * (size_t)B * L.res_out * L.res_out * L.p.cout);
        L.wadj = take(sizeof(__nv_bfloat16) * 9 * (size_t)L.p.cin * L.p.cout);
        L.gd = take(sizeof(float) * (size_t)B * L.p.cout);
        if (L.p.upsample)
            max_T = std::max(max_T, sizeof(__nv_bfloat16) * 4 * (size_t)B * (L.res_in + 1) * (L.res_in + 1) * L.p.cout);
        // the input-gradient GEMM of this layer: channel roles swapped, always a stride-1 walk over the layer's INPUT grid
        Layer BL;
        memset(&BL.p, 0, sizeof(BL.p));
        BL.p.cin = L.p.cout; BL.p.cout = L.p.cin; BL.p.ksize = 3; BL.p.upsample = 0; BL.p.resolution = L.res_in;
        BL.rgb = false;
        BL.res_in = BL.res_out = L.res_in;
        int rc;
        if (L.p.upsample) {
            int taps[27], planes[9], n = 0;
            for (int ph = 0; ph < 4; ++ph) {
                const int py = ph >> 1, px = ph & 1;
                for (int a = py; a < 3; a += 2)
                    for (int b = px; b < 3; b += 2) {
                        taps[3 * n] = (a - py) / 2; taps[3 * n + 1] = (b - px) / 2; taps[3 * n + 2] = a * 3 + b;
                        planes[n++] = ph;
                    }
            }
            rc = plan_gemm(S, BL, taps, n, planes);
        } else {
            rc = plan_gemm(S, BL);
        }
        if (rc) return rc;
        finalize_tiles(BL.gp, B);
        S->blayers.push_back(BL);
    }
    S->off_gT = take(max_T);
    S->off_ones = take(sizeof(float) * (size_t)B * 512);
    // the two inference ping-pong buffers are free in training mode: gradient w.r.t. the conv output / its input
    S->off_gc = S->off_act[0];
    S->off_gx = S->off_act[1];
    // dL/d(skip image) of the lower resolutions live in the (now idle) skip buffers of the forward walk
    size_t o = S->off_rgb[0];
    S->off_grgb.clear();
    for (Layer &L : S->layers)
        if (L.rgb) { S->off_grgb.push_back(o); o += align_up(sizeof(float) * (size_t)B * 3 * L.p.resolution * L.p.resolution); }
    S->ws_bytes = off;
    S->train = true;
    S->cached_ws = nullptr;
    return SG2_OK;
}

// dL/d(styles) of every plan row from dL/d(image).  Must follow a sg2_synth_forward of the SAME batch on the same
// workspace (it reads the activations that call kept).  grad_styles: fp32, row r (the order of sg2_synth_create's layer
// table) at offset B * sum_{rows before r} cin, [B, cin_r]; overwritten.
extern "C" int sg2_synth_backward(sg2_synth *S, void *workspace, int64_t B64, const float *const *noise, const int64_t *noise_bstride,
                                  const float *grad_image, float *grad_styles, sg2_stream_t stream) {
    SG2_REQUIRE(S && S->train, SG2_ERR_BAD_ARG, "synth_backward: the plan is not in training mode (sg2_synth_enable_training)");
    SG2_REQUIRE(B64 >= 0 && B64 <= S->max_batch, SG2_ERR_BAD_ARG, "synth_backward: batch %lld exceeds the plan's max_batch %d",
                (long long)B64, S->max_batch);
    if (B64 == 0) return SG2_OK;
    SG2_REQUIRE(workspace && grad_image && grad_styles && noise && noise_bstride, SG2_ERR_BAD_ARG, "synth_backward: null pointer");
    SG2_REQUIRE(S->cached_ws == workspace && S->cached_B == (int)B64, SG2_ERR_BAD_ARG,
                "synth_backward: no forward pass of this batch on this workspace precedes it");
    const int B = (int)B64;
    cudaStream_t st = as_stream(stream);
    uint8_t *ws = static_cast<uint8_t *>(workspace);
    __nv_bfloat16 *Gc = (__nv_bfloat16 *)(ws + S->off_gc), *Gx = (__nv_bfloat16 *)(ws + S->off_gx), *GT = (__nv_bfloat16 *)(ws + S->off_gT);
    const float *ones = (const float *)(ws + S->off_ones);
    const size_t nL = S->layers.size();

    // row offsets of grad_styles
    std::vector<long long> goff(nL);
    long long gtot = 0;
    for (size_t i = 0; i < nL; ++i) { goff[i] = gtot; gtot += (long long)B * S->layers[i].p.cin; }
    SG2_CUDA_OK(cudaMemsetAsync(grad_styles, 0, sizeof(float) * gtot, st));
    for (Layer &L : S->layers)
        if (!L.rgb) SG2_CUDA_OK(cudaMemsetAsync(ws + L.gd, 0, sizeof(float) * (size_t)B * L.p.cout, st));

    // descriptors of the input-gradient GEMMs
    if (S->bcached_ws != workspace || S->bcached_B != B) {
        for (Layer &L : S->layers) {
            if (L.rgb) continue;
            Layer &BL = S->blayers[L.conv_index];
            finalize_tiles(BL.gp, B);
            int rc = encode_maps(S, BL, Gc, (const __nv_bfloat16 *)(ws + L.wadj), B);
            if (rc) return rc;
            if (L.p.upsample) {
                rc = encode_plane_maps(BL, GT, B, L.res_in);
                if (rc) return rc;
            }
        }
        S->bcached_ws = workspace;
        S->bcached_B = B;
    }

    // 1. the skip chain backwards: dL/d(skip_j) for every ToRGB
    std::vector<const float *> grgb;
    {
        std::vector<int> res;
        for (Layer &L : S->layers) if (L.rgb) res.push_back(L.p.resolution);
        const int n = (int)res.size();
        grgb.assign(n, nullptr);
        grgb[n - 1] = grad_image;
        for (int j = n - 1; j >= 1; --j) {
            RgbUpTParams up;
            up.g = grgb[j]; up.g_prev = (float *)(ws + S->off_grgb[j - 1]); up.planes = B * 3; up.S = res[j - 1];
            memcpy(up.kf, S->kf, sizeof(up.kf));
            const long long total = (long long)up.planes * up.S * up.S;
            rgb_up_adjoint_kernel<<<(unsigned)std::min<long long>(ceil_div64(total, 256), 4096), 256, 0, st>>>(up);
            SG2_LAUNCH_CHECK();
            grgb[j - 1] = up.g_prev;
        }
    }

    // 2. the styled convs, last to first
    int noise_idx = 0, rgb_idx = 0;
    std::vector<int> noise_of(nL, -1), rgb_of(nL, -1);
    for (size_t i = 0; i < nL; ++i) {
        if (S->layers[i].rgb) { rgb_of[i] = rgb_idx++; continue; }
        noise_of[i] = noise_idx++;
    }
    const __nv_bfloat16 *g_cur = nullptr;          // gradient w.r.t. the stored output of the layer being processed
    for (int i = (int)nL - 1; i >= 0; --i) {
        Layer &L = S->layers[i];
        if (L.rgb) continue;
        Layer *next_conv = nullptr, *rgb = nullptr;
        int next_row = -1, rgb_row = -1;
        for (size_t j = i + 1; j < nL; ++j) {
            if (S->layers[j].rgb) { if (j == (size_t)i + 1) { rgb = &S->layers[j]; rgb_row = (int)j; } }
            else { next_conv = &S->layers[j]; next_row = (int)j; break; }
        }
        EpassParams e;
        memset(&e, 0, sizeof(e));
        e.st = (const __nv_bfloat16 *)(ws + L.keep);
        e.g_st = g_cur;
        e.g_conv = Gc;
        e.B = B; e.HW = L.res_out * L.res_out; e.C = L.p.cout;
        e.s_next = next_conv ? (const float *)(ws + next_conv->style) : nullptr;
        e.d = (const float *)(ws + L.demod);
        e.bias = L.p.act_bias;
        e.noise = noise[noise_of[i]]; e.noise_bstride = noise_bstride[noise_of[i]]; e.noise_weight = L.p.noise_weight;
        if (rgb) {
            e.g_rgb = grgb[rgb_of[rgb_row]];
            e.rgb_w = (const float *)(ws + rgb->rgbw);
            e.s_rgb = (const float *)(ws + rgb->style);
            e.g_s_rgb = grad_styles + goff[rgb_row];
        }
        e.g_s_next = next_conv ? grad_styles + goff[next_row] : nullptr;
        e.g_d = (float *)(ws + L.gd);
        SG2_REQUIRE(g_cur || rgb, SG2_ERR_BAD_ARG, "synth_backward: layer %d receives no gradient", i);
        {
            const int ppi = 256 / (e.C / 8);
            // >= 32 pixels per thread where the plane allows it (per-block set-up: 8 x 7 constants, 24 shared-memory atomics),
            // while keeping at least ~4 blocks per SM over the batch
            int per = 32;
            while (per > 4 && (long long)((e.HW + ppi * per - 1) / (ppi * per)) * B < 4 * 148) per >>= 1;
            const int nblk = std::max(1, std::min((e.HW + ppi * per - 1) / (ppi * per), 2048));
            if (e.g_rgb) train_epass_kernel<true><<<dim3(nblk, B), 256, 0, st>>>(e);
            else train_epass_kernel<false><<<dim3(nblk, B), 256, 0, st>>>(e);
            SG2_LAUNCH_CHECK();
        }
        Layer &BL = S->blayers[L.conv_index];
        if (L.p.upsample) {
            FirTParams f;
            f.g = Gc; f.planes = GT; f.B = B; f.r = L.res_in; f.C = L.p.cout;
            memcpy(f.kf, S->kf, sizeof(f.kf));
            // (row-major thread order: a 2-D thread tile for L1 reuse of the windows measured slower, 1.14 -> 1.48 ms at B = 32)
            const long long total = (long long)B * (f.r + 1) * (f.r + 1) * (f.C / 8);
            train_firT_kernel<<<(unsigned)std::min<long long>(ceil_div64(total, 256), 148 * 32), 256, 0, st>>>(f);
            SG2_LAUNCH_CHECK();
        }
        GemmParams g = BL.gp;
        g.mode = 1;
        g.demod = ones;
        g.out = Gx;
        int rc = BL.two_sm ? launch_modconv_gemm2(g, BL.tmA, BL.tmB, S->sms, st) : launch_modconv_gemm(g, BL.tmA, BL.tmB, S->sms, st);
        if (rc) return rc;
        g_cur = Gx;
    }
    // 3. conv1's input modulation (the learned constant) and the demodulation path of every styled conv
    {
        Layer &L0 = S->layers[0];
        const int n = B * L0.p.cin;
        const_grad_kernel<<<(n + 255) / 256, 256, 0, st>>>(grad_styles + goff[0], g_cur, S->const_input, B, L0.p.cin);
        SG2_LAUNCH_CHECK();
        DemodGradJobs dj;
        dj.n = 0;
        for (size_t i = 0; i < nL; ++i) {
            Layer &L = S->layers[i];
            if (L.rgb) continue;
            DemodGradJob &j = dj.job[dj.n++];
            j.style = (const float *)(ws + L.style); j.wsq = (const float *)(ws + L.wsq); j.demod = (const float *)(ws + L.demod);
            j.g_d = (const float *)(ws + L.gd); j.g_s = grad_styles + goff[i]; j.cin = L.p.cin; j.cout = L.p.cout;
        }
        int max_cin = 0;
        for (int q = 0; q < dj.n; ++q) max_cin = std::max(max_cin, dj.job[q].cin);
        demod_grad_kernel<<<dim3((max_cin + 7) / 8, (B + 7) / 8, dj.n), 256, 0, st>>>(dj, B);
        SG2_LAUNCH_CHECK();
    }
    return SG2_OK;
}
