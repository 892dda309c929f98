// synth_gemm.cuh -- parameter blocks of the tcgen05 implicit-GEMM modulated-convolution kernel.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace sg2 {

constexpr int kGemmMaxSub = 4;    // polyphase sub-problems per launch (1 plain conv, 4 transposed conv)
constexpr int kGemmMaxTaps = 9;
constexpr int kBlockM = 128;      // pixels per tile = TMEM lanes
constexpr int kBlockK = 64;       // bf16 channels per stage = one 128-byte swizzle row
constexpr int kStages = 4;
constexpr int kMaxBlockN = 256;
constexpr int kGemm2RingBytes = 176 * 1024;  // operand ring of the cta_group::2 kernel (5 stages of 16 + 16 KiB at N = 256)
constexpr int kGemm2EpiGroups = 4;           // cta_group::2 kernel: column groups of the epilogue (16 warps = 4 per TMEM lane quarter);
                                             // a tile writes min(4, BLOCK_N / 32) partial ToRGB planes
constexpr int kGemmRingBytes = 196 * 1024;   // operand ring of the single-CTA kernel (4 stages of 16 + 32 KiB at N = 256; the
                                             // 128->64 up-conv of the 1024^2 tail: 144 KiB resident weights + 3 x 17 KiB stages)

// One polyphase sub-problem: an output plane [B, PH, PW, Cout] whose pixel (y, x) is
// sum over taps t of  W[wtap[t]] . X[y + dy[t], x + dx[t], :]   (zero outside X).
struct GemmSub {
    int PH, PW;                  // valid output extent of this plane
    int TH, TW, NB;              // tile extent: TH x TW pixels of NB samples (TH*TW*NB <= 128)
    int tiles_x, tiles_y, tiles_b;
    int tile_begin;              // first global tile index of this sub-problem
    int ntaps;
    int dy[kGemmMaxTaps], dx[kGemmMaxTaps], wtap[kGemmMaxTaps];
    int tap_map[kGemmMaxTaps];   // GemmParams::multi_map: which of the four activation tensor maps tap t reads (0..3)
    long long out_off;           // element offset of the plane inside `out`
    int out_H, out_W;            // allocated plane extent (row pitch = out_W * Cout)
    // resident-weights mode: the activation window of a tile is loaded ONCE per K chunk as `nslab`
    // slabs (one per distinct dx) of slab_rows x TW pixels starting at row y0 + slab_dy0; tap t reads
    // the 128 pixel rows that start tap_aoff[t] bytes into the stage (a dy shift = TW rows, TW % 8 == 0)
    int nslab, slab_dx[3], slab_dy0, slab_rows;
    int tap_aoff[kGemmMaxTaps];
};

struct GemmParams {
    int nsub;
    GemmSub sub[kGemmMaxSub];
    int B, Cin, Cout;
    int block_n, n_tiles_n, total_tiles, kchunks;   // kchunks = Cin / block_k
    int block_k;                    // channels per K chunk: 64 (SWIZZLE_128B rows) or 32 (SWIZZLE_64B)
    int kpack;                      // K chunks per pipeline stage: 1, or 2 for narrow BLOCK_N
    // resident-weights mode (narrow layers whose 9*Cin*Cout bf16 weights fit in shared memory): the whole
    // weight tensor is loaded once per CTA, the ring holds only activation slabs (stage_bytes each)
    int resident, resb_bytes, stage_bytes;
    int resident2;                  // cta_group::2 kernel: each CTA keeps its half of every weight tile resident, the ring streams A only
    int mma2;                       // resident mode: two warps issue the MMAs of alternate tiles
    // polyphase interleave: a CTA's chunk of every sub-problem is cut into segments of seg_tiles tiles and the
    // sub-problems are walked segment by segment (seg outer, sub-problem inner), so the four phases of the transposed
    // conv re-read their common input while it is still in L2 (ncu: the phases read it from DRAM 4x otherwise).
    // seg_tiles == 0: one segment = the whole chunk (plain convolutions).
    int seg_tiles, nseg;
    // epilogue schedule of the single-CTA kernel: 1 = the two warp groups take alternate tiles (narrow BLOCK_N,
    // ONE ToRGB partial plane per N tile), 0 = they split the columns of every tile (two partial planes)
    int epi_alt;
    // Fused up-sampling conv (the narrow, HBM-bound octaves): conv_transpose2d(stride 2) followed by the 4x4 blur is ONE
    // stride-2 transposed convolution with a 6x6 kernel (3x3 weights convolved with the blur taps); each of its four
    // output phases (py, px) is a dense 3x3 convolution of the INPUT with its own weights.  The four phases are
    // concatenated along N (column = (py*2 + px) * cout_real + co), so the layer is a plain 3x3 conv with N = 4*Cout
    // whose epilogue writes column block (py, px) of input pixel (y, x) to output pixel (2y+py, 2x+px): no (2r+1)^2
    // intermediate, no FIR pass.  4x the tensor FLOPs of the polyphase form -- used where the layer is bandwidth-bound.
    // multi_map (nsub == 1): the taps of the one sub-problem read DIFFERENT tensors (tmA0..tmA3 chosen by GemmSub::tap_map) --
    // the input gradient of the transposed conv, whose nine taps read the four polyphase planes of the output gradient
    int multi_map;
    int up4;
    int cout_real;
    // Merged polyphase walk of the transposed conv (synth_gemm2p.cu): one tile walk serves the four planes; the nine taps
    // in issue order (plane 3, 2, 1, 0): plane, byte offset of the tap's A operand inside the two-slab activation stage,
    // weight tap, first / last tap of its plane
    int poly4;
    int m_ph[kGemmMaxTaps], m_aoff[kGemmMaxTaps], m_wtap[kGemmMaxTaps], m_first[kGemmMaxTaps], m_last[kGemmMaxTaps];                  // channels of the output tensor (= Cout / 4 when up4, else Cout)
    int dbg;                        // SG2_GEMM_DBG knock-outs for bottleneck analysis (results WRONG when set): 1 no stores, 2 no A loads, 4 no B loads, 8 one MMA per stage
    // epilogue
    int mode;                       // 0: styled conv (noise, bias, lrelu, next-style, ToRGB); 1: plain scaled store
    float clamp;                    // mode 0: clamp the activation to +-clamp after the leaky ReLU (before the folded sqrt(2) gain); 0 = none
                                    // (stylegan2_ada clamp_gain(x, sqrt(2), 256), utils.py:6-7: clamp = 256 / sqrt(2))
    const float *demod;             // [B, Cout]
    const float *noise;             // [B or 1, PH*PW] fp32 or null
    long long noise_bstride;
    const float *noise_weight;      // [1]
    const float *bias;              // [Cout]
    const float *next_style;        // [B, Cout] style of the consumer conv (sqrt(2) gain applied here), or null (no store)
    const float *rgb_w;             // [3, Cout] ToRGB weights (1x1, scale folded), or null
    const float *rgb_style;         // [B, Cout] ToRGB modulation
    float *rgb_part;                // [n_tiles_n, B, 3, PH, PW] fp32 partial ToRGB sums
    __nv_bfloat16 *out;             // NHWC bf16
};

int launch_modconv_gemm(const GemmParams &p, const CUtensorMap *tmA /*[nsub]*/, const CUtensorMap &tmB,
                        int sm_count, cudaStream_t st);
// cta_group::2 variant (synth_gemm2.cu): tmB must have box rows = block_n / 2
int launch_modconv_gemm2(const GemmParams &p, const CUtensorMap *tmA /*[nsub]*/, const CUtensorMap &tmB,
                         int sm_count, cudaStream_t st);

// dx-stacked narrow styled conv (synth_gemm_dxs.cu): Cin, Cout in {32, 64}, one K chunk, N = 3 * Cout
struct DxsParams {
    int B, R, Cin, Cout;
    int tiles_x, tiles_y, total_tiles;       // tiles of 4 rows x 30 output columns (32 input columns)
    const float *demod;                      // [B, Cout]
    const float *noise;                      // [B or 1, R*R] fp32 or null
    long long noise_bstride;
    const float *noise_weight;               // [1]
    const float *bias;                       // [Cout]
    const float *next_style;                 // [B, Cout] or null (no activation store)
    const float *rgb_w;                      // [3, Cout] or null
    const float *rgb_style;                  // [B, Cout]
    float clamp;                             // as GemmParams::clamp
    float *rgb_part;                         // [col_groups][B, 3, R, R], col_groups = 1 (Cout = 32) or 2 (Cout = 64)
    __nv_bfloat16 *out;                      // NHWC bf16 or null
    // last layer of the network (Cout = 32: one column group holds a pixel's whole ToRGB sum): write the final image
    // directly -- rgb + bias + Upsample(skip) (model.py:350-359) -- instead of a partial plane for rgb_combine_kernel
    float *image;                            // [B, 3, R, R] fp32 or null
    const float *rgb_bias;                   // [3]
    const float *prev;                       // [B, 3, R/2, R/2] running skip image, or null
    float kf[16];                            // flipped 4x4 taps (x4) of the skip Upsample
};
// tmA: NHWC activations, box {Cin, 32, 6, 1}; tmB: weights [3][3*Cout][Cin] (= the packed [tap][Cout][Cin] layout), box {Cin, 3*Cout, 1}
int launch_modconv_dxs(const DxsParams &p, const CUtensorMap &tmA, const CUtensorMap &tmB, int sm_count, cudaStream_t st);

// merged polyphase walk on cta_group::2 (synth_gemm2p.cu): tmA box = {64, 8, 17, 1}, tmB box rows = block_n / 2
int launch_modconv_gemm2_poly4(const GemmParams &p, const CUtensorMap &tmA, const CUtensorMap &tmB, int sm_count, cudaStream_t st);

}  // namespace sg2
