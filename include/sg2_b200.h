/* sg2_b200.h -- C ABI of libsg2_b200.so: the B200 (sm_100a) StyleGAN2 synthesis path.
 *
 * Drop-in boundary for the hot path of seva100/stylegan-for-facerec.
 * Every entry point is `extern "C"`, takes raw device pointers, sizes, a dtype enum and a
 * cudaStream_t (as void*), returns an int status (0 = ok) and never allocates, frees or
 * synchronises: the caller (PyTorch host code) owns all memory and the stream, so every call is
 * CUDA-graph capturable.  On failure the message is available from sg2_last_error().
 *
 * Reference interfaces replaced (paths relative to /root/reference/backbone/stylegan2, identical
 * under restyle-encoder/models/stylegan2):
 *   sg2_fused_bias_act        <- pybind `fused.fused_bias_act`      op/fused_bias_act.cpp:11-21
 *                                 (kernel op/fused_bias_act_kernel.cu:18-99)
 *   sg2_bias_act_grad_bias    <- `grad_input.sum(dim)`              op/fused_act.py:31-36
 *   sg2_upfirdn2d             <- pybind `upfirdn2d.upfirdn2d`       op/upfirdn2d.cpp:12-23
 *                                 (kernel op/upfirdn2d_kernel.cu:52-272)
 *   sg2_equal_linear_fwd      <- EqualLinear.forward                model.py:147-157
 *   sg2_mapping_fwd           <- Generator.style (PixelNorm + MLP)  model.py:10-15,378-387
 *   sg2_modulation_fwd        <- modulation affine + demod coeffs   model.py:235-240
 *   sg2_modconv2d_*           <- ModulatedConv2d.forward            model.py:232-273
 *   sg2_noise_bias_act        <- NoiseInjection + FusedLeakyReLU    model.py:282-287,335
 *   sg2_torgb_combine         <- ToRGB bias + Upsample(skip) + add  model.py:350-359
 *   sg2_synth_*               <- Generator.forward synthesis loop   model.py:520-542
 */
#ifndef SG2_B200_H
#define SG2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG2_ABI_VERSION 1

/* status codes */
#define SG2_OK 0
#define SG2_ERR_BAD_ARG 1      /* null pointer, non-positive size, unknown enum */
#define SG2_ERR_UNSUPPORTED 2  /* configuration outside what the kernels implement */
#define SG2_ERR_CUDA 3         /* a CUDA runtime / driver call failed (see sg2_last_error) */
#define SG2_ERR_NO_DEVICE 4    /* no sm_100 device visible */

/* storage dtypes (arithmetic is always fp32 accumulate) */
#define SG2_F32 0
#define SG2_F16 1
#define SG2_BF16 2

typedef void *sg2_stream_t; /* cudaStream_t */

int sg2_abi_version(void);
/* thread-local, valid until the next failing call on this thread */
const char *sg2_last_error(void);
/* number of kernel launches issued through this library by the calling process (all threads) */
int64_t sg2_launch_count(void);
/* a CUDA-graph replay re-launches kernels without passing through this library: the host adds them */
void sg2_note_launches(int64_t n);
/* CPU-only self check of the multiply-shift division used by the vector kernels: 0 if n/d matches */
int sg2_selftest_fastdiv(uint32_t d, uint32_t n);

/* ---------------------------------------------------------------------------------------------
 * fused bias + activation.   out[i] = act(x[i] + bias[(i / step_b) % size_b]) * scale
 * act: 1 linear, 3 leaky-relu(alpha).  grad: 0 value, 1 first derivative gated on ref[i] > 0,
 * 2 second derivative (zero).  bias / ref may be NULL.  In-place (out == x) is allowed.          */
int sg2_fused_bias_act(void *out, const void *x, const void *bias, const void *ref, int64_t n,
                       int64_t step_b, int64_t size_b, int act, int grad, float alpha, float scale,
                       int dtype, sg2_stream_t stream);

/* grad_bias[c] = sum over b, hw of grad_input[b, c, hw]   (grad_input laid out [B, C, HW]);
 * grad_bias is always fp32 [C] (bit-deterministic unless C < 2*SMs, where slices are combined
 * with fp32 atomics). */
int sg2_bias_act_grad_bias(void *grad_bias, const void *grad_input, int64_t B, int64_t C,
                           int64_t HW, int dtype, sg2_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * upfirdn2d on x[major, in_h, in_w, minor] -> out[major, out_h, out_w, minor],
 * out_h = (in_h*up_y + pad_y0 + pad_y1 - kh) / down_y + 1 (likewise out_w).
 * kernel: kh*kw fp32 taps on the device, as the caller holds them (the op applies them flipped,
 * i.e. a true convolution).  Supported: 1 <= up, down <= 4 per axis, kh, kw <= 16, any pads
 * (negative pads crop).  Anything else returns SG2_ERR_UNSUPPORTED (the reference returns
 * uninitialised memory there, upfirdn2d_kernel.cu:172-268).                                     */
int sg2_upfirdn2d(void *out, const void *x, const float *kernel, int64_t major, int in_h, int in_w,
                  int minor, int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                  int pad_x0, int pad_x1, int pad_y0, int pad_y1, int dtype, sg2_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * EqualLinear: out[b,o] = act( (sum_j x[b,j] * w[o,j]) * w_scale + bias[o] * lr_mul )
 * act = 0: linear; act = 1: leaky-relu(0.2) * sqrt(2).  bias may be NULL.  x/w/bias/out in dtype. */
int sg2_equal_linear_fwd(void *out, const void *x, const void *w, const void *bias, int64_t B,
                         int in_dim, int out_dim, float w_scale, float lr_mul, int act, int dtype,
                         sg2_stream_t stream);

/* Mapping network: w = MLP(pixel_norm(z)); weights[n_mlp][dim][dim], biases[n_mlp][dim] given as
 * arrays of n_mlp device pointers (host arrays).  dim in {32, 64, 128, 256, 512}.  scratch: a
 * [B, dim] buffer of the same dtype (layers ping-pong between it and w_out; may be NULL when
 * n_mlp <= 1).  One launch per layer, PixelNorm fused into the first.                           */
int sg2_mapping_fwd(void *w_out, const void *z, const void *const *weights,
                    const void *const *biases, int n_mlp, int64_t B, int dim, float lr_mul,
                    int pixel_norm, void *scratch, int dtype, sg2_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Modulated convolution, exact fp32-accumulate path (NCHW, SIMT).  Three calls:
 *  1. sg2_modconv2d_prep: weight[Cout,Cin,k,k] (dtype) -> wt fp32 [Cin*k*k, Cout] (scaled by
 *     conv_scale, the caller passes 1/sqrt(Cin*k*k), model.py:214-215) and, if wsq != NULL,
 *     wsq fp32 [Cin, Cout] = conv_scale^2 * sum_k w^2.
 *  2. sg2_modulation_fwd: style[b,ci] = latent[b,:] . mod_w[ci,:] * mod_scale + mod_b[ci]*lr_mul
 *     (fp32) and, if demod != NULL, demod[b,co] = rsqrt(sum_ci style^2 * wsq[ci,co] + 1e-8).
 *  3. sg2_modconv2d_fwd: out[b,co] = demod[b,co] * conv(x[b] * style[b], wt)   with
 *     mode 0: stride 1, zero pad k/2            out H x W
 *     mode 1: transposed, stride 2, pad 0       out (2H+k-2) x (2W+k-2)   (blur is a separate
 *             sg2_upfirdn2d call, model.py:254-257)
 *     mode 2: stride 2, pad 0                   out ((H-k)/2+1) x ((W-k)/2+1)
 *     demod may be NULL (ToRGB); style may be NULL (plain shared-weight conv, used by the
 *     backward pass).  k in {1, 3}.                                                   */
int sg2_modconv2d_prep(float *wt, float *wsq, const void *weight, int Cin, int Cout, int k,
                       float conv_scale, int dtype, sg2_stream_t stream);
int sg2_modulation_fwd(float *style, float *demod, const void *latent, int64_t latent_stride,
                       const void *mod_w, const void *mod_b, const float *wsq, int64_t B,
                       int style_dim, int Cin, int Cout, float mod_scale, float lr_mul, int dtype,
                       sg2_stream_t stream);
int sg2_modconv2d_fwd(void *out, const void *x, const float *wt, const float *style,
                      const float *demod, int64_t B, int Cin, int Cout, int H, int W, int k,
                      int mode, int dtype, sg2_stream_t stream);

/* out = act(x + noise_weight[0] * noise[b or 0, 0, h, w] + bias[c]) * act_scale    (NCHW)
 * noise_bstride = H*W for per-sample noise, 0 for one noise map broadcast over the batch.
 * noise or bias may be NULL; act: 1 linear, 3 leaky-relu(alpha).                                */
int sg2_noise_bias_act(void *out, const void *x, const void *noise, int64_t noise_bstride,
                       const void *noise_weight, const void *bias, int64_t B, int C, int64_t HW,
                       int act, float alpha, float act_scale, int dtype, sg2_stream_t stream);

/* ToRGB tail: out[b,c,y,x] = conv[b,c,y,x] + bias[c] + upfirdn2d(skip, kernel, up=2, pad)[b,c,y,x]
 * skip [B,C,H/2,W/2] may be NULL (first ToRGB).  kernel: kh*kw fp32 taps (already x4).          */
int sg2_torgb_combine(void *out, const void *conv, const void *bias, const void *skip,
                      const float *kernel, int kh, int kw, int pad0, int pad1, int64_t B, int C,
                      int H, int W, int dtype, sg2_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * StyleGAN2-ADA variant of the decoder (restyle-encoder/models/stylegan2_ada, selected by psp.py:24-30).
 *
 * sg2_smooth_upsample2x  <- SmoothUpsample.forward (stylegan2_ada/utils.py:76-95): nearest x2, replication pad
 *   (2,1,2,1), 4x4 correlation with `taps` (fp32, 16 values, row-major, NOT flipped), fused with what follows it
 *   in SynthesisLayer2.forward (generator.py:198-204) / SynthesisBlock.forward (generator.py:134-137):
 *     out = clamp( act( up(x) + noise * noise_strength[0] + bias[c] + addend ) * gain, -clamp, clamp )
 *   x [B,C,H,W] -> out [B,C,2H,2W]; noise [B or 1,1,2H,2W] (noise_bstride = 4*H*W or 0), bias [C], addend
 *   [B,C,2H,2W] may each be NULL; act: 1 linear, 3 leaky-relu(alpha); clamp <= 0 disables the clamp.
 * sg2_ada_bias_act  <- clamp_gain(act(x + noise + bias), gain, clamp) (utils.py:6-7; generator.py:148-151,204)
 *   for the layers without up-sampling and for ToRGBLayer2.                                          */
int sg2_smooth_upsample2x(void *out, const void *x, const float *taps, int64_t B, int C, int H, int W,
                          const void *noise, int64_t noise_bstride, const void *noise_strength,
                          const void *bias, const void *addend, int act, float alpha, float gain,
                          float clamp, int dtype, sg2_stream_t stream);
/* adjoint of the plain SmoothUpsample: grad_out [planes, 2H, 2W] -> grad_x [planes, H, W] (autograd of utils.py:89-95) */
int sg2_smooth_upsample2x_bwd(void *grad_x, const void *grad_out, const float *taps, int64_t planes, int H, int W,
                              int dtype, sg2_stream_t stream);
int sg2_ada_bias_act(void *out, const void *x, const void *noise, int64_t noise_bstride,
                     const void *noise_strength, const void *bias, int64_t B, int C, int64_t HW, int act,
                     float alpha, float gain, float clamp, int dtype, sg2_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The step after the decoder in ReStyle / pSp (restyle-encoder/models/psp.py:33,113-114;
 * training/coach_restyle_psp.py:143-156): face_pool = AdaptiveAvgPool2d((256,256)) for integer ratios
 * (x [planes, out_h*factor, out_w*factor] -> out [planes, out_h, out_w]) and
 * F.interpolate(..., mode='bilinear', align_corners=False) (x [planes, in_h, in_w] -> [planes, out_h, out_w]). */
int sg2_avg_pool_int(void *out, const void *x, int64_t planes, int out_h, int out_w, int factor, int dtype,
                     sg2_stream_t stream);
int sg2_resize_bilinear(void *out, const void *x, int64_t planes, int in_h, int in_w, int out_h, int out_w,
                        int dtype, sg2_stream_t stream);
/* Their adjoints (what autograd derives for the two calls above when the coaches back-propagate the image losses,
 * coach_restyle_psp.py:86-101,143-156): grad_x from grad_out, same shape arguments as the forward call. */
int sg2_avg_pool_int_bwd(void *grad_x, const void *grad_out, int64_t planes, int out_h, int out_w, int factor,
                         int dtype, sg2_stream_t stream);
int sg2_resize_bilinear_bwd(void *grad_x, const void *grad_out, int64_t planes, int in_h, int in_w, int out_h,
                            int out_w, int dtype, sg2_stream_t stream);
/* tensor2im on the device (restyle-encoder/utils/common.py:5-11): out[i] = uint8(clip((x[i] + 1) / 2, 0, 1) * 255),
 * same element order as x; `out` 4-byte aligned. */
int sg2_image_to_uint8(void *out, const void *x, int64_t total, int dtype, sg2_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Whole-network bf16 synthesis engine (NHWC activations, tcgen05/TMEM implicit GEMM fed by TMA,
 * fused epilogues).  Replaces the loop of Generator.forward, model.py:520-533.
 *
 * Per-layer parameter table handed over by the host (device pointers to the fp32 master
 * parameters exactly as the module's state_dict holds them).                                    */
typedef struct sg2_conv_params {
    const float *weight;      /* [1,Cout,Cin,k,k] */
    const float *mod_weight;  /* [Cin, style_dim] */
    const float *mod_bias;    /* [Cin] */
    const float *noise_weight;/* [1]        (styled convs; NULL for ToRGB) */
    const float *act_bias;    /* [Cout]     (styled convs) or ToRGB bias [1,3,1,1] */
    int32_t cin, cout, ksize; /* ksize 3 (styled conv) or 1 (ToRGB) */
    int32_t upsample;         /* 1: transposed stride-2 conv + blur (model.py:246-257) */
    int32_t latent_index;     /* which row of latent[b, :, :] modulates this layer */
    int32_t resolution;       /* output resolution of the layer */
} sg2_conv_params;

typedef struct sg2_synth sg2_synth; /* opaque launch plan */

/* Builds the static launch plan for Generator(size, style_dim, channel_multiplier) at up to
 * max_batch samples.  layers: conv1, to_rgb1, then per octave (up conv, conv, to_rgb) in network
 * order; n_layers = 2 + 3*(log2(size)-2).  const_input: [1,C,4,4] fp32.  blur_taps: 4x4 fp32 HOST
 * array (make_kernel([1,3,3,1])*4, model.py:18-26,77-78).  Does not touch the GPU.              */
int sg2_synth_create(sg2_synth **plan, int size, int style_dim, int max_batch,
                     const sg2_conv_params *layers, int n_layers, const float *const_input,
                     const float *blur_taps_host);
/* The same plan for the stylegan2_ada decoder variant (restyle-encoder/models/stylegan2_ada/generator.py:55-204, the
 * decoder psp.py:24-30 builds with opts.generator_ada): same layer table order (first_block.conv1, first_block.torgb, then
 * per block conv0 with upsample = 1, conv1, torgb; mod_weight / mod_bias = the affine FullyConnectedLayer, noise_weight =
 * noise_strength), w_dim in place of style_dim, resample_taps_host = the 4x4 SmoothUpsample kernel (utils.py:76-83).
 * Differences handled by the plan: no equalised-lr conv scale, conv -> SmoothUpsample ordering in the up-sampling layers,
 * clamp_gain(x, sqrt(2), 256) after every activation, ToRGB clamp, SmoothUpsample of the running image.              */
int sg2_synth_create_ada(sg2_synth **plan, int size, int w_dim, int max_batch,
                         const sg2_conv_params *layers, int n_layers, const float *const_input,
                         const float *resample_taps_host);
void sg2_synth_destroy(sg2_synth *plan);
/* bytes of device workspace the caller must provide (packed weights + activations + scratch) */
int64_t sg2_synth_workspace_bytes(const sg2_synth *plan);
/* human-readable launch plan (one line per kernel), for tests and DESIGN.md; returns chars written */
int sg2_synth_describe(const sg2_synth *plan, char *buf, int buflen);
/* (re)pack the master weights into the engine layouts inside `workspace`; call once after
 * create and again whenever the parameters change.                                              */
int sg2_synth_pack(sg2_synth *plan, void *workspace, sg2_stream_t stream);
/* latent [B, n_latent, style_dim] fp32; noise[l] fp32 [B or 1, 1, r, r] for each styled conv
 * (n_noise = 2*log2(size)-3 pointers, host array), noise_bstride[l] = r*r or 0; every noise map densely packed and
 * 16-byte aligned (the up-sampling layers read their tile of it through a TMA tensor map);
 * image [B,3,size,size] fp32 (NCHW, as the reference returns it).                               */
int sg2_synth_forward(sg2_synth *plan, void *workspace, const float *latent, int64_t B,
                      const float *const *noise, const int64_t *noise_bstride, float *image,
                      sg2_stream_t stream);

/* face_pool folded into the forward (psp.py:33,113-114 AdaptiveAvgPool2d((256,256)) after the decoder):
 * with factor 2 or 4 every following sg2_synth_forward also writes the image average-pooled by `factor` to
 * pooled [B,3,size/factor,size/factor] fp32 from inside its last launch; keep_full = 0 skips the full-resolution
 * image altogether (sg2_synth_forward's `image` may then be null).  factor 0 switches it off.  Inference plans only. */
int sg2_synth_set_pooled_output(sg2_synth *plan, float *pooled, int factor, int keep_full);

/* Training mode of the plan (the frozen decoder inside the ReStyle / pSp fine-tuning step,
 * coach_restyle_psp.py:138-168; autograd of model.py:232-359 w.r.t. the styles).
 * sg2_synth_enable_training: call right after sg2_synth_create, before sg2_synth_workspace_bytes / sg2_synth_pack; the
 *   workspace grows by one kept NHWC bf16 activation per styled conv, the adjoint weight packs and reduction scratch, and
 *   sg2_synth_forward then keeps every layer's output instead of recycling two buffers.
 * sg2_synth_backward: dL/d(style) of every plan row from grad_image [B,3,size,size] fp32.  Must follow the
 *   sg2_synth_forward of the same batch on the same workspace.  grad_styles (fp32, overwritten): row r of the layer table at
 *   offset B * sum_{rows before r} cin, shape [B, cin_r]; the caller maps it to dL/d(latent) through the modulation
 *   weights (EqualLinear, model.py:152-155).  noise / noise_bstride as given to the forward call.  Noise maps and
 *   parameters receive no gradient (frozen-decoder direction).                                                      */
int sg2_synth_enable_training(sg2_synth *plan);
int sg2_synth_backward(sg2_synth *plan, void *workspace, int64_t B, const float *const *noise, const int64_t *noise_bstride,
                       const float *grad_image, float *grad_styles, sg2_stream_t stream);

/* Optional per-kernel timing: the caller passes CUDA events it created (cudaEvent_t handles);
 * every forward records events[0] before the first launch and events[k] after the k-th launch, in
 * the order sg2_synth_describe lists them.  Pass NULL / 0 to switch it off.                     */
int sg2_synth_set_profile_events(sg2_synth *plan, void **events, int n_events);
int sg2_synth_profile_events_used(const sg2_synth *plan);

/* The modulated 1x1 convolution of ToRGB (model.py:350-355: ModulatedConv2d(in, 3, 1, demodulate=False)) and what
 * autograd derives for it, one HBM-bound pass over the activation each way, fp32 math:
 *   fwd: y[b,k,p] = sum_c w[k,c] * s[b,c] * x[b,c,p]          w [K,C] fp32 (conv scale folded), s [B,C] fp32, K <= 4
 *   bwd: gx[b,c,p] = s[b,c] * t, gs[b,c] += sum_p x[b,c,p] * t, t = sum_k w[k,c] * gy[b,k,p]   (gs fp32, zero it first) */
int sg2_rgb_modconv_fwd(void *y, const void *x, const float *w, const float *s, int64_t B, int C, int K, int64_t HW,
                        int dtype, sg2_stream_t stream);
int sg2_rgb_modconv_bwd(void *gx, float *gs, const void *gy, const void *x, const float *w, const float *s, int64_t B,
                        int C, int K, int64_t HW, int dtype, sg2_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * One 3x3 stride-1 'same' convolution on the tensor-core kernel (tcgen05/TMEM, operands by TMA), outside the
 * whole-network plan: the contraction inside ModulatedConv2d.forward (model.py:254-273: F.conv2d with padding 1)
 * once modulation and demodulation are factored out, and -- called with flipped taps and swapped channel roles --
 * what autograd derives as its input gradient.  NHWC bf16 operands, fp32 accumulation:
 *   out[b,y,x,co] = scale[b,co] * sum_{a,c,ci} x[b, y+a-1, x+c-1, ci] * wp[a*3+c][co][ci]
 * x [B,r,r,Cin], out [B,r,r,Cout] bf16 (16-byte aligned), scale [B,Cout] fp32, Cin % 32 == 0, Cout % 16 == 0, r >= 4.
 * sg2_conv3x3_tc_pack: weight fp32 [Cout,Cin,3,3] -> wp bf16 [9][Cout][Cin], `scale` folded in.            */
int sg2_conv3x3_tc_pack(void *wp, const float *weight, int cin, int cout, float scale, sg2_stream_t stream);
int sg2_conv3x3_tc(void *out, const void *x, const void *wp, const float *scale, int64_t B, int r, int cin,
                   int cout, sg2_stream_t stream);
/* The same kernel over a subset of the 3x3 window: taps = ntaps rows {dy, dx, w} (HOST array, dy/dx in -1..1, w in 0..8),
 *   out[b,y,x,co] = scale[b,co] * sum_t sum_ci x[b, y+dy_t, x+dx_t, ci] * wp[w_t][co][ci]
 * (the polyphase components of the stride-2 convolution autograd derives for the transposed conv are of this form). */
int sg2_conv_taps_tc(void *out, const void *x, const void *wp, const float *scale, int64_t B, int r, int cin,
                     int cout, const int *taps, int ntaps, sg2_stream_t stream);
/* F.conv_transpose2d(x, w, stride=2, padding=0) (model.py:246-252, the up-sampling ModulatedConv2d before its blur) as
 * four polyphase planes: planes[(py*2+px)][b][y][x][co] = scale[b,co] * T[b, 2y+py, 2x+px, co]; each plane is allocated
 * (r+1) x (r+1), its valid extent is (r+1-py) x (r+1-px) and the rest is not written.  wp as above, Cout % 32 == 0. */
int sg2_conv_transpose3x3_tc(void *planes, const void *x, const void *wp, const float *scale, int64_t B, int r, int cin,
                    int cout, sg2_stream_t stream);

/* The two passes either side of a tensor-core convolution in the differentiable path, with the per-(sample, channel)
 * factor of the factored ModulatedConv2d (model.py:236-240: style on the way in, demodulation on the way out) folded in,
 * and the reduction its adjoint needs in the same pass.  `dtype` is that of the NCHW tensors; NHWC tensors are bf16.
 *   sg2_nchw_to_nhwc_bf16: out[b,p,c] = bf16(x[b,c,p] * scale[b,c]);  if other: red[b,c] += sum_p x[b,c,p] * other[b,p,c]
 *   sg2_nhwc_bf16_to_nchw: out[b,c,p] = scale[b,c] * h[b,p,c];        if other: red[b,c] += sum_p other[b,c,p] * h[b,p,c]
 * scale may be NULL (1); other/red both NULL or both set (red fp32 [B,C], accumulated with atomics: zero it first). */
int sg2_nchw_to_nhwc_bf16(void *out, const void *x, const float *scale, const void *other, float *red, int64_t B,
                          int C, int64_t HW, int dtype, sg2_stream_t stream);
int sg2_nhwc_bf16_to_nchw(void *out, const void *h, const float *scale, const void *other, float *red, int64_t B,
                          int C, int64_t HW, int dtype, sg2_stream_t stream);
/* The same two passes with the NHWC side stored as the four polyphase planes of a W x W image (W odd):
 * planes[(y&1)*2 + (x&1)][b][y>>1][x>>1][c], each plane (W+1)/2 squared -- the layout sg2_conv_transpose3x3_tc writes and
 * the polyphase form of its input gradient reads (no interleaving copies).  sg2_nchw_to_polyphase_bf16 writes only
 * the valid cells of the planes: zero them first. */
int sg2_nchw_to_polyphase_bf16(void *planes, const void *x, const float *scale, const void *other_planes, float *red,
                               int64_t B, int C, int W, int dtype, sg2_stream_t stream);
int sg2_polyphase_bf16_to_nchw(void *out, const void *planes, const float *scale, const void *other, float *red,
                               int64_t B, int C, int W, int dtype, sg2_stream_t stream);
/* out[b,c,y,x] = scale[b,c] * sum_k parts[k][b][y][x][c] over the top-left W x W corner of n_parts (1..4) bf16 tensors
 * [B][pitch][pitch][C], summed in fp32; other / red as above.  Turns the four polyphase components of the transposed
 * conv's input gradient (sg2_conv_taps_tc outputs) into grad_x and grad_s in one pass. */
int sg2_sum_parts_bf16_to_nchw(void *out, const void *parts, int n_parts, int pitch, const float *scale, const void *other,
                               float *red, int64_t B, int C, int W, int dtype, sg2_stream_t stream);
/* The second pass with the StyledConv tail (NoiseInjection + FusedLeakyReLU, model.py:282-287,331-337) applied:
 *   out[b,c,p] = lrelu(scale[b,c] * h[b,p,c] + noise_weight[0] * noise[b or 0, p] + bias[c], alpha) * gain
 * (noise, noise_weight, bias: tensors of `dtype`; noise / bias may be NULL; noise_bstride = HW or 0), and its adjoint:
 *   g = gy * (y > 0 ? 1 : alpha) * gain   (fused_bias_act_kernel.cu:28-47 with grad = 1, y = the output above)
 *   out[b,p,c] = bf16(g * scale[b,c]);  red[b,c] += sum_p g * other[b,p,c];  red_sum[b,c] += sum_p g
 * (other/red both NULL or both set; red_sum may be NULL; both fp32 [B,C], accumulated with atomics: zero them first). */
int sg2_nhwc_bf16_to_nchw_act(void *out, const void *h, const float *scale, const void *noise, int64_t noise_bstride,
                              const void *noise_weight, const void *bias, float alpha, float gain, int64_t B, int C,
                              int64_t HW, int dtype, sg2_stream_t stream);
int sg2_nchw_to_nhwc_bf16_actgrad(void *out, const void *gy, const void *y, float alpha, float gain, const float *scale,
                                  const void *other, float *red, float *red_sum, int64_t B, int C, int64_t HW,
                                  int dtype, sg2_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SG2_B200_H */
