// synth_kernels.cu -- the small kernels around the tcgen05 GEMM in the bf16 synthesis engine:
// weight packing (once per weight update), style / demodulation tables (once per forward, all
// layers in one launch each), constant input, the polyphase FIR + noise/bias/lrelu/modulate pass of
// the up-sampling layers and the RGB skip chain.  All HBM-bound or tiny; NHWC bf16 activations.
#include "common.cuh"
#include "synth_kernels.cuh"

namespace sg2 {

// ---- pack: W fp32 [Cout,Cin,k,k] -> Wp bf16 [k*k][Cout][Cin] (conv_scale folded), wsq fp32 [Cin][Cout]
__global__ void __launch_bounds__(256)
pack_conv_weight_kernel(__nv_bfloat16 *__restrict__ wp, float *__restrict__ wsq, const float *__restrict__ w,
                        int Cin, int Cout, int kk, float scale) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;     // over (co, ci), ci fastest
    if (i >= (int64_t)Cin * Cout) return;
    const int co = (int)(i / Cin), ci = (int)(i - (int64_t)co * Cin);
    float ss = 0.f;
    for (int t = 0; t < kk; ++t) {
        const float v = w[((int64_t)co * Cin + ci) * kk + t] * scale;
        wp[((int64_t)t * Cout + co) * Cin + ci] = __float2bfloat16_rn(v);
        ss += v * v;
    }
    if (wsq) wsq[(int64_t)ci * Cout + co] = ss;
}

// ---- pack for the fused up-sampling conv: conv_transpose2d(stride 2, 3x3) followed by the 4x4 blur (model.py:246-257) is
// one stride-2 transposed conv with the 6x6 kernel (w * blur); output phase (py, px) of it is a dense 3x3 conv of the
// input:   out[2y+py, 2x+px] = sum_{dy,dx in -1..1} x[y+dy, x+dx] . Wc[(dy,dx)][(py,px)]
//          Wc[(dy,dx)][(py,px)][co][ci] = sum_{a,b in 0..3} kf[a][b] * w[co][ci][py+a-1-2dy][px+b-1-2dx]   (indices in 0..2)
// kf = the flipped blur taps (x4 gain included) as upfirdn2d applies them with pad (1, 1).
// Layout: bf16 [tap = (dy+1)*3 + (dx+1)][n = (py*2+px)*Cout + co][ci], conv_scale folded in.
struct Kf16 { float v[16]; };
__global__ void __launch_bounds__(256)
pack_upfused_weight_kernel(__nv_bfloat16 *__restrict__ wp, const float *__restrict__ w, int Cin, int Cout, float scale, Kf16 kf) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;     // over (tap, phase, co, ci), ci fastest
    if (i >= (int64_t)36 * Cin * Cout) return;
    const int ci = (int)(i % Cin);
    int64_t rest = i / Cin;
    const int co = (int)(rest % Cout);
    rest /= Cout;
    const int ph = (int)(rest & 3), tap = (int)(rest >> 2);
    const int py = ph >> 1, px = ph & 1, dy = tap / 3 - 1, dx = tap % 3 - 1;
    const float *wk = w + ((int64_t)co * Cin + ci) * 9;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int a2 = py + a - 1 - 2 * dy;
        if (a2 < 0 || a2 > 2) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int b2 = px + b - 1 - 2 * dx;
            if (b2 < 0 || b2 > 2) continue;
            acc += kf.v[a * 4 + b] * wk[a2 * 3 + b2];
        }
    }
    wp[((int64_t)tap * 4 * Cout + (int64_t)ph * Cout + co) * Cin + ci] = __float2bfloat16_rn(acc * scale);
}

// ToRGB weight fp32 [3,Cin] -> scaled fp32 [3][Cin]
__global__ void pack_rgb_weight_kernel(float *__restrict__ out, const float *__restrict__ w, int n, float scale) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[i] = w[i] * scale;
}

// ---- styles of every layer in one launch: s_l[b,ci] = latent[b, idx_l, :] . mod_w_l[ci,:] * scale + mod_b_l[ci]
// A small fp32 SGEMM per layer, [B, style_dim] x [style_dim, cin]: block = STY_CI input channels of one
// layer x STY_SB samples, K staged through shared memory in chunks of STY_K; a thread owns one
// channel x 8 samples, so the latent values are warp-wide broadcasts and the weight rows are read
// conflict-free (row pitch STY_K + 4 floats).  Every weight is read from HBM once per sample tile.
constexpr int STY_SB = 64, STY_CI = kStyleBlockCi, STY_K = 32, STY_PITCH = STY_K + 4;
static_assert(STY_CI == 32 && STY_SB == 64, "styles_kernel: a warp spans the channel tile, 8 warps x 8 the sample tile");
__global__ void __launch_bounds__(256)
styles_kernel(StyleJobs jobs, const float *__restrict__ latent, int B, int n_latent, int style_dim) {
    __shared__ __align__(16) float s_lat[STY_SB][STY_PITCH];
    __shared__ __align__(16) float s_w[STY_CI][STY_PITCH];
    int j = 0;
    while (j + 1 < jobs.n && (int)blockIdx.x >= jobs.job[j + 1].block_begin) ++j;
    const StyleJob &job = jobs.job[j];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ci0 = ((int)blockIdx.x - job.block_begin) * STY_CI, b0 = (int)blockIdx.y * STY_SB;
    const int ci = ci0 + lane;
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.f;
    for (int k0 = 0; k0 < style_dim; k0 += STY_K) {
        __syncthreads();
        // stage: 64 latent rows x 32 k (8 floats per thread) and 32 weight rows x 32 k (4 per thread)
#pragma unroll
        for (int i = 0; i < (STY_SB * STY_K) / 256; ++i) {
            const int e = tid + 256 * i, row = e / STY_K, k = e - row * STY_K;
            const int b = b0 + row;
            s_lat[row][k] = b < B ? __ldg(latent + ((int64_t)b * n_latent + job.latent_index) * style_dim + k0 + k) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < (STY_CI * STY_K) / 256; ++i) {
            const int e = tid + 256 * i, row = e / STY_K, k = e - row * STY_K;
            s_w[row][k] = ci0 + row < job.cin ? __ldg(job.mod_w + (int64_t)(ci0 + row) * style_dim + k0 + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < STY_K; k += 4) {
            const float4 w4 = *reinterpret_cast<const float4 *>(&s_w[lane][k]);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float4 l4 = *reinterpret_cast<const float4 *>(&s_lat[warp + 8 * u][k]);
                acc[u] = fmaf(l4.x, w4.x, fmaf(l4.y, w4.y, fmaf(l4.z, w4.z, fmaf(l4.w, w4.w, acc[u]))));
            }
        }
    }
    if (ci >= job.cin) return;
    const float bv = __ldg(job.mod_b + ci);
    const float scale = rsqrtf((float)style_dim);            // EqualLinear scale, lr_mul = 1 (model.py:144)
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int b = b0 + warp + 8 * u;
        if (b < B) job.out[(int64_t)b * job.cin + ci] = acc[u] * scale + bv;
    }
}

// ---- demod of every styled conv in one launch: d[b,co] = rsqrt(sum_ci s^2 * wsq[ci,co] + 1e-8)
// block = 64 output channels x 4 slices of the input channels (reduced through shared memory) x
// DEMOD_SB samples: each wsq element is loaded once for DEMOD_SB samples, 4x more loads in flight.
constexpr int DEMOD_SB = 8, DEMOD_CO = 64;
__global__ void __launch_bounds__(256)
demod_kernel(DemodJobs jobs, int B) {
    extern __shared__ float s_dm[];                          // [DEMOD_SB][cin] s^2, then [4][DEMOD_SB][64] partials
    const DemodJob &job = jobs.job[blockIdx.z];
    const int b0 = blockIdx.y * DEMOD_SB;
    if ((int)blockIdx.x * DEMOD_CO >= job.cout) return;
    const int nb = min(DEMOD_SB, B - b0);
    float *s_s2 = s_dm, *s_part = s_dm + DEMOD_SB * job.cin;
    for (int i = threadIdx.x; i < DEMOD_SB * job.cin; i += 256) {
        const int sb = i / job.cin, ci = i - sb * job.cin;
        const float s = sb < nb ? job.style[(int64_t)(b0 + sb) * job.cin + ci] : 0.f;
        s_s2[i] = s * s;
    }
    __syncthreads();
    const int col = threadIdx.x & 63, slice = threadIdx.x >> 6;
    const int co = blockIdx.x * DEMOD_CO + col;
    float acc[DEMOD_SB];
#pragma unroll
    for (int sb = 0; sb < DEMOD_SB; ++sb) acc[sb] = 0.f;
    if (co < job.cout) {
        const int per = (job.cin + 3) / 4, lo = slice * per, hi = min(job.cin, lo + per);
#pragma unroll 8
        for (int ci = lo; ci < hi; ++ci) {
            const float w = __ldg(job.wsq + (int64_t)ci * job.cout + co);
#pragma unroll
            for (int sb = 0; sb < DEMOD_SB; ++sb) acc[sb] += s_s2[sb * job.cin + ci] * w;
        }
    }
#pragma unroll
    for (int sb = 0; sb < DEMOD_SB; ++sb) s_part[(slice * DEMOD_SB + sb) * DEMOD_CO + col] = acc[sb];
    __syncthreads();
    for (int o = threadIdx.x; o < DEMOD_SB * DEMOD_CO; o += 256) {
        const int sb = o / DEMOD_CO, c2 = o - sb * DEMOD_CO;
        const int co2 = blockIdx.x * DEMOD_CO + c2;
        if (sb < nb && co2 < job.cout) {
            float t = 0.f;
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) t += s_part[(sl * DEMOD_SB + sb) * DEMOD_CO + c2];
            job.demod[(int64_t)(b0 + sb) * job.cout + co2] = rsqrtf(t + 1e-8f);
        }
    }
}

// ---- constant input, pre-modulated by conv1's style: X0[b,y,x,c] = const[c,y,x] * s[b,c]   (NHWC bf16)
__global__ void __launch_bounds__(256)
const_input_kernel(__nv_bfloat16 *__restrict__ out, const float *__restrict__ cst, const float *__restrict__ style,
                   int B, int C, int HW) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (int64_t)B * HW * C) return;
    const int c = (int)(i % C);
    const int p = (int)((i / C) % HW);
    const int b = (int)(i / ((int64_t)C * HW));
    out[i] = __float2bfloat16_rn(cst[(int64_t)c * HW + p] * style[(int64_t)b * C + c]);
}

// ---- up-sampling layer tail, SIMT stencil (used for the 32-channel 1024^2 layer; the 64-channel
// multiples go through the tensor-core kernel in synth_fir.cu): blur (4x4, pad (1,1)) over the 4
// polyphase planes of the transposed conv, + noise, + bias, lrelu, x sqrt(2) * next style, NHWC bf16.
// Block: 256 threads = (8 x CB) cells (each a 2x2 output block) x (CH/8) channel vectors.
constexpr int FIR_CA = 8;
template <int CH> struct FirGeo {
    static constexpr int VEC = CH / 8, CELLS = 256 / VEC, CB = CELLS / FIR_CA;
    static constexpr int ROWS = 2 * FIR_CA + 3, COLS = 2 * CB + 3;
};

template <int CH>
__global__ void __launch_bounds__(256)
upfir_kernel(UpfirParams p) {
    using G = FirGeo<CH>;
    __shared__ __align__(16) __nv_bfloat16 s_t[G::ROWS][G::COLS][CH];
    const int tid = threadIdx.x;
    const int tiles_b = (p.r + G::CB - 1) / G::CB;
    const int a0 = (blockIdx.x / tiles_b) * FIR_CA, b0 = (blockIdx.x % tiles_b) * G::CB;
    const int c0 = blockIdx.y * CH;
    const int n = blockIdx.z;
    const int P = p.r + 1;                                   // allocated plane extent
    // stage the T window: rows t = 2*a0-1 .. 2*a0+2*CA+1, cols likewise
    for (int i = tid; i < G::ROWS * G::COLS * G::VEC; i += 256) {
        const int v = i % G::VEC, pix = i / G::VEC;
        const int row = pix / G::COLS, col = pix - row * G::COLS;
        const int t = 2 * a0 - 1 + row, u = 2 * b0 - 1 + col;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (t >= 0 && u >= 0 && t <= 2 * p.r && u <= 2 * p.r) {
            const int ph = ((t & 1) << 1) | (u & 1);
            const int a = t >> 1, b = u >> 1;
            const __nv_bfloat16 *src = p.T + (int64_t)ph * p.plane_stride +
                                       (((int64_t)n * P + a) * P + b) * p.C + c0 + v * 8;
            val = __ldg(reinterpret_cast<const uint4 *>(src));
        }
        *reinterpret_cast<uint4 *>(&s_t[row][col][v * 8]) = val;
    }
    __syncthreads();
    const int v = tid % G::VEC, cell = tid / G::VEC;
    const int ca = cell / G::CB, cb = cell - ca * G::CB;
    const int a = a0 + ca, b = b0 + cb;
    if (a >= p.r || b >= p.r) return;
    float acc[2][2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[i][j][e] = 0.f;
    // window rows 2*ca .. 2*ca+4 ; output (oy, ox) uses window rows oy..oy+3, cols ox..ox+3
#pragma unroll
    for (int wr = 0; wr < 5; ++wr) {
#pragma unroll
        for (int wc = 0; wc < 5; ++wc) {
            const uint4 raw = *reinterpret_cast<const uint4 *>(&s_t[2 * ca + wr][2 * cb + wc][v * 8]);
            const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&raw);
            float x[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h[e]);
                x[2 * e] = f.x; x[2 * e + 1] = f.y;
            }
#pragma unroll
            for (int oy = 0; oy < 2; ++oy) {
                const int j = wr - oy;
                if (j < 0 || j > 3) continue;
#pragma unroll
                for (int ox = 0; ox < 2; ++ox) {
                    const int i = wc - ox;
                    if (i < 0 || i > 3) continue;
                    const float k = p.kf[j * 4 + i];
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[oy][ox][e] += k * x[e];
                }
            }
        }
    }
    const int R = 2 * p.r;
    const float nw = p.noise ? __ldg(p.noise_weight) : 0.f;
    float bias[8], sn[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        bias[e] = __ldg(p.bias + c0 + v * 8 + e);
        sn[e] = 1.41421356237f * __ldg(p.next_style + (int64_t)n * p.C + c0 + v * 8 + e);
    }
#pragma unroll
    for (int oy = 0; oy < 2; ++oy)
#pragma unroll
        for (int ox = 0; ox < 2; ++ox) {
            const int Y = 2 * a + oy, X = 2 * b + ox;
            const float nz = p.noise ? nw * __ldg(p.noise + (int64_t)n * p.noise_bstride + (int64_t)Y * R + X) : 0.f;
            uint32_t packed[4];
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                float v0 = acc[oy][ox][e] + nz + bias[e], v1 = acc[oy][ox][e + 1] + nz + bias[e + 1];
                v0 = fmaxf(v0, 0.2f * v0) * sn[e];
                v1 = fmaxf(v1, 0.2f * v1) * sn[e + 1];
                __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
                packed[e / 2] = *reinterpret_cast<uint32_t *>(&h);
            }
            __nv_bfloat16 *dst = p.out + (((int64_t)n * R + Y) * R + X) * p.C + c0 + v * 8;
            *reinterpret_cast<uint4 *>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
}

// ---- RGB skip chain: rgb[n,c,Y,X] = sum_nt part[nt,n,c,Y,X] + bias[c] + up2(prev)[n,c,Y,X]   (NCHW fp32)
__host__ __device__ __forceinline__ int fdiv2(int a) { return a >= 0 ? a / 2 : -((-a + 1) / 2); }

// one quad (four consecutive pixels of row Y starting at X0, a multiple of 4) of plane `plane`: sum of the partial ToRGB planes
// + bias (+ clamp) + the up-sampled skip image
__device__ __forceinline__ float4 rgb_quad(const RgbParams &p, const float *sp, float bias, int64_t total, int64_t i, int Y, int X0,
                                           int R, int S) {
    float4 v = make_float4(bias, bias, bias, bias);
    for (int nt = 0; nt < p.n_parts; ++nt) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p.part + (int64_t)nt * total + i));
        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    if (p.clamp > 0.f) {             // stylegan2_ada: clamp(y + bias) before the skip is added
        v.x = fminf(fmaxf(v.x, -p.clamp), p.clamp); v.y = fminf(fmaxf(v.y, -p.clamp), p.clamp);
        v.z = fminf(fmaxf(v.z, -p.clamp), p.clamp); v.w = fminf(fmaxf(v.w, -p.clamp), p.clamp);
    }
    if (sp && p.smooth) {
        // SmoothUpsample (stylegan2_ada/utils.py:76-95): output (2i+py, 2j+px) = sum_{a,b} kf[a][b] *
        // skip[clamp(i + off[py][a])][clamp(j + off[px][b])], off = {{-1,-1,0,0},{-1,0,0,1}}
        const int i0 = Y >> 1, py = Y & 1;
        float up[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int oy = py ? (a == 0 ? -1 : (a == 3 ? 1 : 0)) : (a < 2 ? -1 : 0);
            const float *row = sp + min(max(i0 + oy, 0), S - 1) * S;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int X = X0 + e, j0 = X >> 1, px = X & 1;
#pragma unroll
                for (int bq = 0; bq < 4; ++bq) {
                    const int ox = px ? (bq == 0 ? -1 : (bq == 3 ? 1 : 0)) : (bq < 2 ? -1 : 0);
                    up[e] = fmaf(p.kf[a * 4 + bq], __ldg(row + min(max(j0 + ox, 0), S - 1)), up[e]);
                }
            }
        }
        v.x += up[0]; v.y += up[1]; v.z += up[2]; v.w += up[3];
    } else if (sp) {
        // up=2, pad (2,1), 4x4 taps (index math of upfirdn2d_kernel.cu:112-129, specialised to a quad that
        // starts at a multiple of 4): output row Y reads skip rows iy0, iy0+1 with tap rows ky0, ky0+2;
        // the quad reads skip columns h-1 .. h+2 (h = X0/2) with tap columns (0,2) (1,3) (0,2) (1,3)
        const int iy0 = fdiv2(Y - 1), ky0 = 2 * iy0 + 2 - Y, h = X0 >> 1;
        float up0 = 0.f, up1 = 0.f, up2 = 0.f, up3 = 0.f;
#pragma unroll
        for (int a2 = 0; a2 < 2; ++a2) {
            const int iy = iy0 + a2;
            if (iy < 0 || iy >= S) continue;
            const float *row = sp + iy * S;
            float k[4];             // tap row ky0 + 2*a2 (static indices: the taps stay in the constant bank)
#pragma unroll
            for (int j = 0; j < 4; ++j) k[j] = ky0 ? p.kf[(1 + 2 * a2) * 4 + j] : p.kf[(2 * a2) * 4 + j];
            const float c0 = h >= 1 ? __ldg(row + h - 1) : 0.f, c1 = __ldg(row + h);
            const float c2 = h + 1 < S ? __ldg(row + h + 1) : 0.f, c3 = h + 2 < S ? __ldg(row + h + 2) : 0.f;
            up0 += k[0] * c0 + k[2] * c1;
            up1 += k[1] * c1 + k[3] * c2;
            up2 += k[0] * c1 + k[2] * c2;
            up3 += k[1] * c2 + k[3] * c3;
        }
        v.x += up0; v.y += up1; v.z += up2; v.w += up3;
    }
    return v;
}

template <int POOL>      // 0: the plain launch; 2 / 4: the last launch of a forward with the face-pooled image written on the way
__global__ void __launch_bounds__(256)
rgb_combine_kernel(RgbParams p) {
    // grid.y = plane (n*3 + c); a thread owns 4 consecutive pixels of a row: float4 traffic, 32-bit index math.
    // POOL = 2 / 4: the thread walks the POOL rows of its quad column and also emits their average --
    // AdaptiveAvgPool2d((R / POOL, R / POOL)), pSp's face_pool (psp.py:33,113-114) -- so the pooled image costs no second pass
    // over the full-resolution one; p.out may then be null (pooled image only).
    constexpr int PROW = POOL > 1 ? POOL : 1;
    const int R = p.R, S = R / 2, quads = R / 4, nrows = R / PROW;
    const int64_t plane_px = (int64_t)R * R, total = (int64_t)p.B * 3 * plane_px;
    for (int plane = blockIdx.y; plane < p.B * 3; plane += gridDim.y) {
        const float bias = __ldg(p.bias + plane % 3);
        const float *sp = p.prev ? p.prev + (int64_t)plane * S * S : nullptr;
        for (int q = blockIdx.x * 256 + threadIdx.x; q < nrows * quads; q += gridDim.x * 256) {
            const int Yp = q / quads, X0 = (q - Yp * quads) * 4;
            float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < PROW; ++r) {
                const int Y = Yp * PROW + r;
                const int64_t i = (int64_t)plane * plane_px + (int64_t)Y * R + X0;
                const float4 v = rgb_quad(p, sp, bias, total, i, Y, X0, R, S);
                if (POOL == 0 || p.out) *reinterpret_cast<float4 *>(p.out + i) = v;
                if (POOL) { sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w; }
            }
            if (POOL == 4) {
                p.pool_out[(int64_t)plane * (plane_px / 16) + (int64_t)Yp * (R / 4) + (X0 >> 2)] = (sum.x + sum.y + sum.z + sum.w) * (1.f / 16.f);
            } else if (POOL == 2) {
                *reinterpret_cast<float2 *>(p.pool_out + (int64_t)plane * (plane_px / 4) + (int64_t)Yp * (R / 2) + (X0 >> 1)) =
                    make_float2((sum.x + sum.y) * 0.25f, (sum.z + sum.w) * 0.25f);
            }
        }
    }
}

// ---- stylegan2_ada up-sampling layer: SmoothUpsample + noise + bias + lrelu + clamp + next style, NHWC bf16 -------------
// a thread owns 8 channels of one INPUT cell (i, j) = the 2x2 output block (2i.., 2j..): 9 clamped 16-byte loads, 4 stores
__global__ void __launch_bounds__(256)
smooth_up_nhwc_kernel(SmoothUpParams p) {
    const int r = p.r, R = 2 * r, c8n = p.C >> 3;
    const long long total = (long long)p.B * r * r * c8n;
    const float nw = p.noise ? __ldg(p.noise_weight) : 0.f;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const int c8 = (int)(idx % c8n);
        long long rest = idx / c8n;
        const int j = (int)(rest % r);
        rest /= r;
        const int i = (int)(rest % r), b = (int)(rest / r);
        const int rr[3] = {max(i - 1, 0), i, min(i + 1, r - 1)}, cc[3] = {max(j - 1, 0), j, min(j + 1, r - 1)};
        float w[9][8];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p.T + (((long long)b * r + rr[a]) * r + cc[q]) * p.C + c8 * 8));
                const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&v);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 t = __bfloat1622float2(h[e]);
                    w[a * 3 + q][2 * e] = t.x; w[a * 3 + q][2 * e + 1] = t.y;
                }
            }
        float bias[8], sn[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            bias[e] = __ldg(p.bias + c8 * 8 + e);
            sn[e] = 1.41421356237f * __ldg(p.next_style + (long long)b * p.C + c8 * 8 + e);
        }
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
            const int Y = 2 * i + (ph >> 1), X = 2 * j + (ph & 1);
            const float nz = p.noise ? nw * __ldg(p.noise + (long long)b * p.noise_bstride + (long long)Y * R + X) : 0.f;
            uint4 o;
            __nv_bfloat162 *oh = reinterpret_cast<__nv_bfloat162 *>(&o);
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float acc = 0.f;
#pragma unroll
                for (int t = 0; t < 9; ++t) acc = fmaf(p.wp[ph][t], w[t][e], acc);
                acc += nz + bias[e];
                acc = fmaxf(acc, 0.2f * acc);
                if (p.clamp > 0.f) acc = fminf(fmaxf(acc, -p.clamp), p.clamp);
                v[e] = acc * sn[e];
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) oh[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
            *reinterpret_cast<uint4 *>(p.out + (((long long)b * R + Y) * R + X) * p.C + c8 * 8) = o;
        }
    }
}

int launch_smooth_up(const SmoothUpParams &p, cudaStream_t st) {
    SG2_REQUIRE(p.C % 8 == 0 && p.r >= 1, SG2_ERR_BAD_ARG, "smooth_up: bad shape");
    const long long total = (long long)p.B * p.r * p.r * (p.C / 8);
    smooth_up_nhwc_kernel<<<(unsigned)std::min<long long>(ceil_div64(total, 256), 148 * 32), 256, 0, st>>>(p);
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

// ---- launch helpers --------------------------------------------------------------------------------
int launch_pack_conv_weight(__nv_bfloat16 *wp, float *wsq, const float *w, int Cin, int Cout, int kk, float scale,
                            cudaStream_t st) {
    const int64_t n = (int64_t)Cin * Cout;
    pack_conv_weight_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(wp, wsq, w, Cin, Cout, kk, scale);
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}
int launch_pack_upfused_weight(__nv_bfloat16 *wp, const float *w, int Cin, int Cout, float scale, const float *kf16_host,
                               cudaStream_t st) {
    Kf16 kf;
    for (int i = 0; i < 16; ++i) kf.v[i] = kf16_host[i];
    const int64_t n = (int64_t)36 * Cin * Cout;
    pack_upfused_weight_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(wp, w, Cin, Cout, scale, kf);
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}
int launch_pack_rgb_weight(float *out, const float *w, int n, float scale, cudaStream_t st) {
    pack_rgb_weight_kernel<<<(n + 255) / 256, 256, 0, st>>>(out, w, n, scale);
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}
int launch_styles(const StyleJobs &jobs, int total_blocks, const float *latent, int B, int n_latent, int style_dim,
                  cudaStream_t st) {
    dim3 grid(total_blocks, (B + STY_SB - 1) / STY_SB);
    SG2_REQUIRE(style_dim % STY_K == 0, SG2_ERR_UNSUPPORTED, "styles: style_dim %d is not a multiple of %d", style_dim, STY_K);
    styles_kernel<<<grid, 256, 0, st>>>(jobs, latent, B, n_latent, style_dim);
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}
int launch_demod(const DemodJobs &jobs, int max_cin, int max_cout, int B, cudaStream_t st) {
    dim3 grid((max_cout + DEMOD_CO - 1) / DEMOD_CO, (B + DEMOD_SB - 1) / DEMOD_SB, jobs.n);
    demod_kernel<<<grid, 256, sizeof(float) * (DEMOD_SB * max_cin + 4 * DEMOD_SB * DEMOD_CO), st>>>(jobs, B);
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}
int launch_const_input(__nv_bfloat16 *out, const float *cst, const float *style, int B, int C, int HW, cudaStream_t st) {
    const int64_t n = (int64_t)B * C * HW;
    const_input_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(out, cst, style, B, C, HW);
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}
int launch_upfir(const UpfirParams &p, int B, cudaStream_t st) {
    const int ta = (p.r + FIR_CA - 1) / FIR_CA;
    if (p.C % 64 == 0) {
        const int tb = (p.r + FirGeo<64>::CB - 1) / FirGeo<64>::CB;
        upfir_kernel<64><<<dim3(ta * tb, p.C / 64, B), 256, 0, st>>>(p);
    } else {
        SG2_REQUIRE(p.C % 32 == 0, SG2_ERR_UNSUPPORTED, "upfir: channel count %d is not a multiple of 32", p.C);
        const int tb = (p.r + FirGeo<32>::CB - 1) / FirGeo<32>::CB;
        upfir_kernel<32><<<dim3(ta * tb, p.C / 32, B), 256, 0, st>>>(p);
    }
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}
int launch_rgb_combine(const RgbParams &p, int sms, cudaStream_t st) {
    SG2_REQUIRE(p.pool == 0 || ((p.pool == 2 || p.pool == 4) && p.pool_out && p.R % (4 * p.pool / (p.pool == 4 ? 4 : 2)) == 0 && p.R >= 4 * p.pool),
                SG2_ERR_BAD_ARG, "rgb_combine: pooled output needs factor 2 or 4 and a buffer (R %d, pool %d)", p.R, p.pool);
    SG2_REQUIRE(p.out || p.pool, SG2_ERR_BAD_ARG, "rgb_combine: no output");
    const int quads = (p.R / (p.pool > 1 ? p.pool : 1)) * (p.R / 4);
    // enough blocks to keep the loads of a whole plane in flight (a cap of 64 left every thread 16 dependent-looking
    // iterations at 1024^2: 2.7 TB/s)
    dim3 grid((unsigned)std::min(512, (quads + 255) / 256), (unsigned)std::min(p.B * 3, 65535));
    (void)sms;
    if (p.pool == 4) rgb_combine_kernel<4><<<grid, 256, 0, st>>>(p);
    else if (p.pool == 2) rgb_combine_kernel<2><<<grid, 256, 0, st>>>(p);
    else rgb_combine_kernel<0><<<grid, 256, 0, st>>>(p);
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

}  // namespace sg2
