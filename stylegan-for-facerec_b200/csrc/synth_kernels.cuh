// synth_kernels.cuh -- parameter blocks + launchers of the engine's auxiliary kernels.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sg2 {

constexpr int kMaxJobs = 40;   // 17 styled convs + 9 ToRGB at 1024^2
constexpr int kStyleBlockCi = 32;   // input channels per block of styles_kernel

struct StyleJob {
    const float *mod_w, *mod_b;   // [cin, style_dim], [cin]
    float *out;                   // [B, cin]
    int cin, latent_index, block_begin;
};
struct StyleJobs { int n; StyleJob job[kMaxJobs]; };

struct DemodJob {
    const float *style, *wsq;     // [B, cin], [cin, cout]
    float *demod;                 // [B, cout]
    int cin, cout;
};
struct DemodJobs { int n; DemodJob job[kMaxJobs]; };

struct UpfirParams {
    const __nv_bfloat16 *T;       // 4 planes [(py,px)][B][r+1][r+1][C], demodulated transposed-conv output
    long long plane_stride;       // elements between planes
    __nv_bfloat16 *out;           // [B][2r][2r][C]
    int r, C;
    const float *noise; long long noise_bstride; const float *noise_weight;   // noise itself is read through the tensor map tmN
    const float *bias;            // [C], 16-byte aligned
    const float *next_style;      // [B][C], 16-byte aligned
    float kf[16];                 // flipped 4x4 taps
};

// tensor-core variant (synth_fir.cu): T is read through TMA descriptors
struct UpfirTcParams {
    __nv_bfloat16 *out;           // [B][2r][2r][C]
    int r, C, B;
    int cbw, nsamp;               // a tile = 128 columns = (128 / cbw) column blocks of cbw channels over nsamp samples
    int tiles_x, tiles_y, tiles_c, total_tiles;   // tiles_c = channel groups per sample (C / 128, or 1)
    int store_mode;               // 0: one thread stores whole column blocks (box cbw x 8 x 16)
    int dbg;                      // SG2_FIR_DBG knock-outs for bottleneck analysis (results are WRONG when set): 1 no loads, 2 no stores, 4 no epilogue math, 8 no MMA
    const float *noise; long long noise_bstride; const float *noise_weight;   // noise itself is read through the tensor map tmN
    const float *bias;            // [C], 16-byte aligned
    const float *next_style;      // [B][C], 16-byte aligned
    const __nv_bfloat16 *toeplitz; // [128][256] row-major Toeplitz matrix of the taps (copied into tensor memory)
};

struct RgbParams {
    float *out;                   // [B,3,R,R]
    const float *part;            // [n_parts][B,3,R,R]
    int n_parts;
    const float *bias;            // [3]
    const float *prev;            // [B,3,R/2,R/2] or null
    int B, R;
    float kf[16];
    // stylegan2_ada (generator.py:134-137,148-151): the ToRGB output is clamped to +-clamp before the skip is added, and the
    // running image is up-sampled by SmoothUpsample (nearest x2, edge replication, 4x4 correlation with kf as given)
    float clamp;                  // 0: none
    int smooth;                   // 0: zero-insertion up-sampling, pad (2,1), flipped taps (rosinality Upsample)
    // last launch of a forward only: also write the image average-pooled by `pool` (2 or 4; 0 = off) to pool_out
    // [B,3,R/pool,R/pool]; `out` may then be null
    int pool;
    float *pool_out;
};

// stylegan2_ada up-sampling layer after its convolution (generator.py:198-204, utils.py:76-95): SmoothUpsample of the
// demodulated conv output + noise + bias + leaky ReLU + clamp, times the consumer's style; NHWC bf16
struct SmoothUpParams {
    const __nv_bfloat16 *T;       // [B, r, r, C]
    __nv_bfloat16 *out;           // [B, 2r, 2r, C]
    int B, r, C;
    const float *noise; long long noise_bstride; const float *noise_weight;   // [B or 1, 2r*2r]
    const float *bias;            // [C]
    const float *next_style;      // [B, C]
    float clamp;                  // on the activation before the sqrt(2) gain; 0: none
    float wp[4][9];               // per output phase (py*2+px): the 16 taps folded onto the clamped 3x3 input neighbourhood
};
int launch_smooth_up(const SmoothUpParams &p, cudaStream_t st);

int launch_pack_conv_weight(__nv_bfloat16 *wp, float *wsq, const float *w, int Cin, int Cout, int kk, float scale, cudaStream_t st);
// fused up-sampling conv: composite (3x3 weights * 4x4 blur) polyphase weights, bf16 [9][4*Cout][Cin]; kf16_host = flipped taps
int launch_pack_upfused_weight(__nv_bfloat16 *wp, const float *w, int Cin, int Cout, float scale, const float *kf16_host,
                               cudaStream_t st);
int launch_pack_rgb_weight(float *out, const float *w, int n, float scale, cudaStream_t st);
int launch_styles(const StyleJobs &jobs, int total_blocks, const float *latent, int B, int n_latent, int style_dim, cudaStream_t st);
int launch_demod(const DemodJobs &jobs, int max_cin, int max_cout, int B, cudaStream_t st);
int launch_const_input(__nv_bfloat16 *out, const float *cst, const float *style, int B, int C, int HW, cudaStream_t st);
int launch_upfir(const UpfirParams &p, int B, cudaStream_t st);
int launch_rgb_combine(const RgbParams &p, int sms, cudaStream_t st);
void build_fir_toeplitz(uint16_t *out /*[128][256] bf16 bits*/, const float *kf /*flipped 4x4 taps*/);
int launch_upfir_tc(const UpfirTcParams &p, const CUtensorMap &tmN /* noise [B or 1][2r][2r] fp32, box 8 x 16 x maps per tile */, const CUtensorMap *tmT /*[4]*/, const CUtensorMap &tmO,
                    int sms, cudaStream_t st);

}  // namespace sg2
