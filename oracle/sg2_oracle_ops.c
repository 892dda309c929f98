/* CPU oracle (plain C) for the two native kernels of the reference and the modulated convolution.
 * TEST INFRASTRUCTURE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this library; the product path never does.
 *
 * Each function restates, with scalar loops, what the reference computes:
 *   oracle_fused_bias_act_f32  <- op/fused_bias_act_kernel.cu:18-49 (index math :21-31, switch :36-47)
 *   oracle_upfirdn2d_f32       <- op/upfirdn2d_kernel.cu:52-137 (tap flip :77, zero fill :99-103,
 *                                 polyphase index math :112-121, accumulation :123-129) and the
 *                                 output size of :167-168
 *   oracle_modconv2d_f32       <- model.py:232-273 (modulate :236, demodulate :238-240, grouped conv /
 *                                 conv_transpose stride 2 :246-271); the blur that follows the
 *                                 transposed conv is applied by the caller with oracle_upfirdn2d_f32
 * Accumulation is in double so the result is a ground truth rather than one more fp32 ordering.
 * Parity pin: tests/golden/ops.npz + layers.npz, produced by the reference itself
 * (tests/golden/make_golden.py); checked in tests/test_oracle_golden.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static int floordiv(int a, int b) { int q = a / b; if (q * b > a) q--; return q; }

/* y = act(x + b[(i / step_b) % size_b]) * scale ; act/grad codes as in the reference switch */
int oracle_fused_bias_act_f32(float *out, const float *x, const float *bias, const float *ref,
                              int64_t n, int64_t step_b, int64_t size_b, int act, int grad,
                              float alpha, float scale) {
    for (int64_t i = 0; i < n; ++i) {
        float v = x[i];
        if (bias) v += bias[(i / step_b) % size_b];
        float r = ref ? ref[i] : 0.0f;
        float y;
        switch (act * 10 + grad) {
            default:
            case 10: case 11: y = v; break;
            case 12: case 32: y = 0.0f; break;
            case 30: y = (v > 0.0f) ? v : v * alpha; break;
            case 31: y = (r > 0.0f) ? v : v * alpha; break;
        }
        out[i] = y * scale;
    }
    return 0;
}

int oracle_upfirdn2d_out(int in, int up, int down, int p0, int p1, int k) {
    return (in * up + p0 + p1 - k + down) / down;          /* upfirdn2d_kernel.cu:167 */
}

/* x [major,in_h,in_w,minor] -> out [major,out_h,out_w,minor]; kernel [kh,kw] (un-flipped, as passed) */
int oracle_upfirdn2d_f32(float *out, const float *x, const float *kernel, int64_t major, int in_h,
                         int in_w, int minor, int kh, int kw, int up_x, int up_y, int down_x,
                         int down_y, int px0, int px1, int py0, int py1) {
    const int out_h = oracle_upfirdn2d_out(in_h, up_y, down_y, py0, py1, kh);
    const int out_w = oracle_upfirdn2d_out(in_w, up_x, down_x, px0, px1, kw);
    if (out_h <= 0 || out_w <= 0) return 1;
    for (int64_t m = 0; m < major; ++m)
        for (int oy = 0; oy < out_h; ++oy)
            for (int ox = 0; ox < out_w; ++ox) {
                /* position of the output sample on the zero-inserted, padded grid */
                const int mid_y = oy * down_y + up_y - 1 - py0;
                const int mid_x = ox * down_x + up_x - 1 - px0;
                const int iy0 = floordiv(mid_y, up_y), ix0 = floordiv(mid_x, up_x);
                const int ky0 = (iy0 + 1) * up_y - mid_y - 1, kx0 = (ix0 + 1) * up_x - mid_x - 1;
                for (int c = 0; c < minor; ++c) {
                    double acc = 0.0;
                    for (int ty = 0; ky0 + ty * up_y < kh; ++ty)
                        for (int tx = 0; kx0 + tx * up_x < kw; ++tx) {
                            const int iy = iy0 + ty, ix = ix0 + tx;
                            if (iy < 0 || ix < 0 || iy >= in_h || ix >= in_w) continue;
                            const int fy = kh - 1 - (ky0 + ty * up_y), fx = kw - 1 - (kx0 + tx * up_x);
                            acc += (double)x[((m * in_h + iy) * in_w + ix) * minor + c] *
                                   (double)kernel[fy * kw + fx];
                        }
                    out[((m * out_h + oy) * out_w + ox) * minor + c] = (float)acc;
                }
            }
    return 0;
}

/* mode 0: conv stride 1 pad k/2 ; mode 1: conv_transpose stride 2 pad 0 (out = 2*in + k - 2) ;
 * mode 2: conv stride 2 pad 0.  x [B,Cin,H,W], weight [Cout,Cin,k,k], style [B,Cin] (already the
 * output of the modulation affine), out [B,Cout,OH,OW]. */
int oracle_modconv2d_f32(float *out, const float *x, const float *weight, const float *style,
                         int B, int Cin, int Cout, int H, int W, int k, int demodulate, int mode) {
    const double scale = 1.0 / sqrt((double)Cin * k * k);
    const int OH = mode == 0 ? H : mode == 1 ? (H - 1) * 2 + k : (H - k) / 2 + 1;
    const int OW = mode == 0 ? W : mode == 1 ? (W - 1) * 2 + k : (W - k) / 2 + 1;
    double *wm = (double *)malloc(sizeof(double) * Cout * Cin * k * k);
    double *acc = (double *)malloc(sizeof(double) * OH * OW);
    if (!wm || !acc) return 2;
    for (int b = 0; b < B; ++b) {
        for (int co = 0; co < Cout; ++co) {
            double ss = 0.0;
            for (int ci = 0; ci < Cin; ++ci)
                for (int t = 0; t < k * k; ++t) {
                    double v = scale * weight[(co * Cin + ci) * k * k + t] * style[b * Cin + ci];
                    wm[(co * Cin + ci) * k * k + t] = v;
                    ss += v * v;
                }
            if (demodulate) {
                double d = 1.0 / sqrt(ss + 1e-8);
                for (int i = 0; i < Cin * k * k; ++i) wm[co * Cin * k * k + i] *= d;
            }
        }
        for (int co = 0; co < Cout; ++co) {
            for (int i = 0; i < OH * OW; ++i) acc[i] = 0.0;
            for (int ci = 0; ci < Cin; ++ci) {
                const float *xp = x + ((int64_t)(b * Cin + ci) * H) * W;
                const double *wp = wm + (co * Cin + ci) * k * k;
                if (mode == 1) {
                    for (int iy = 0; iy < H; ++iy)
                        for (int ix = 0; ix < W; ++ix)
                            for (int ky = 0; ky < k; ++ky)
                                for (int kx = 0; kx < k; ++kx)
                                    acc[(iy * 2 + ky) * OW + ix * 2 + kx] += xp[iy * W + ix] * wp[ky * k + kx];
                } else {
                    const int st = mode == 2 ? 2 : 1, pd = mode == 0 ? k / 2 : 0;
                    for (int oy = 0; oy < OH; ++oy)
                        for (int ox = 0; ox < OW; ++ox)
                            for (int ky = 0; ky < k; ++ky)
                                for (int kx = 0; kx < k; ++kx) {
                                    const int iy = oy * st + ky - pd, ix = ox * st + kx - pd;
                                    if (iy < 0 || ix < 0 || iy >= H || ix >= W) continue;
                                    acc[oy * OW + ox] += xp[iy * W + ix] * wp[ky * k + kx];
                                }
                }
            }
            float *op = out + ((int64_t)(b * Cout + co) * OH) * OW;
            for (int i = 0; i < OH * OW; ++i) op[i] = (float)acc[i];
        }
    }
    free(wm); free(acc);
    return 0;
}
