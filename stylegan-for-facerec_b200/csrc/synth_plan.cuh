// synth_plan.cuh -- the launch plan of the whole-network engine (private to csrc/): per-layer tiling, workspace offsets and
// cached TMA descriptors.  synth.cu builds and walks it forward; synth_train.cu walks it backward.
#pragma once
#include <string>
#include <vector>

#include "common.cuh"
#include "synth_gemm.cuh"
#include "synth_kernels.cuh"
#include "tc_ptx.cuh"

namespace sg2plan {
using namespace sg2;

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

inline size_t align_up(size_t v, size_t a = 1024) { return (v + a - 1) / a * a; }

struct Layer {
    sg2_conv_params p;
    bool rgb;
    int res_in, res_out;
    // workspace offsets (bytes)
    size_t wp = 0, wsq = 0, rgbw = 0, style = 0, demod = 0;
    // GEMM tiling (styled convs)
    int block_n = 0;
    bool two_sm = false;  // run this layer on the cta_group::2 kernel (synth_gemm2.cu)
    bool dxs = false;        // narrow plain conv on the dx-stacked kernel (synth_gemm_dxs.cu)
    DxsParams dp;
    CUtensorMap tmDA, tmDB;
    size_t keep = 0;         // training: this layer's stored (activated, next-style-modulated) output, kept for the backward pass
    size_t wadj = 0;         // training: adjoint weights bf16 [9][Cin][Cout] (taps flipped for the plain conv)
    size_t gd = 0;           // training: dL/d demod [B, Cout] fp32
    int conv_index = -1;     // ordinal among the styled convs
    bool ada_up = false;     // stylegan2_ada up-sampling layer: 3x3 conv at the INPUT resolution, then SmoothUpsample + epilogue
    bool fused_up = false;   // up-sampling layer run as ONE 3x3 conv with N = 4*Cout composite (weights * blur) columns, no FIR pass
    int n_gemm = 0;          // N of the GEMM: Cout, or 4*Cout for a fused up-sampling layer
    GemmParams gp;        // static part, pointers filled per forward
    CUtensorMap tmA[kGemmMaxSub], tmB;
    CUtensorMap tmT[4];   // up-sampling layers: the 4 polyphase planes as the FIR kernel reads them
    CUtensorMap tmO;      // ... and its output [B][2r][2r][C] as the TMA store writes it (box 64 ch x 8 x 16)
};

}  // namespace sg2plan

struct sg2_synth {
    typedef sg2plan::Layer Layer;
    int size, style_dim, max_batch, log_size, n_latent, num_layers;
    std::vector<sg2plan::Layer> layers;
    const float *const_input;
    float kf[16];                 // flipped blur taps (x4)
    size_t off_act[2], off_T, off_rgb[2], off_part, off_toeplitz, ws_bytes;
    std::vector<uint16_t> toeplitz;   // host copy of the FIR Toeplitz matrix (bf16 bits), uploaded by pack
    bool fir_simt = false;
    // descriptor cache
    void *cached_ws = nullptr;
    int cached_B = -1;
    // stylegan2_ada decoder variant (generator.py:55-204 of restyle-encoder/models/stylegan2_ada): no equalised-lr scale on
    // the conv weights, conv -> SmoothUpsample ordering in the up-sampling layers, clamp_gain(..., 256), ToRGB clamp
    bool ada = false;
    float kf_raw[16];             // the resampling taps as given (SmoothUpsample correlates, it does not flip)
    float wp_ada[4][9];           // per output phase: the 16 taps folded onto the clamped 3x3 input neighbourhood
    // training mode (synth_train.cu): forward keeps every layer's output, sg2_synth_backward walks the plan in reverse
    bool train = false;
    std::vector<sg2plan::Layer> blayers;      // one input-gradient GEMM per styled conv (channel roles swapped)
    size_t off_gc = 0, off_gx = 0, off_gT = 0, off_ones = 0, off_red = 0, red_bytes = 0, off_act_in = 0;
    std::vector<size_t> off_grgb;             // dL/d(skip image) per ToRGB, fp32 NCHW
    void *bcached_ws = nullptr;
    int bcached_B = -1;
    // profiling hooks
    // pooled output of the last rgb_combine launch (sg2_synth_set_pooled_output): factor 0 = off
    int pool = 0;
    bool pool_keep_full = true;
    float *pool_out = nullptr;
    cudaEvent_t *events = nullptr;
    int n_events = 0, events_used = 0;
    std::string description;
    int sms = 148;
};


namespace sg2plan {
// synth.cu
// custom_taps: n_custom rows {dy, dx, weight tap}; tap_planes (with custom_taps): the activation tensor map (0..3) each tap reads
int plan_gemm(sg2_synth *S, Layer &L, const int *custom_taps = nullptr, int n_custom = 0, const int *tap_planes = nullptr);
void finalize_tiles(GemmParams &g, int B);
int encode_maps(sg2_synth *S, Layer &L, const __nv_bfloat16 *x, const __nv_bfloat16 *wp, int B);
int rec(sg2_synth *S, cudaStream_t st, const char *what = "");
// synth_train.cu
int train_pack(sg2_synth *S, uint8_t *ws, cudaStream_t st);
}  // namespace sg2plan
