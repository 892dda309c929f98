// synth.cu -- whole-network bf16 synthesis engine: static launch plan + C ABI (sg2_synth_*).
//
// Replaces the synthesis loop of Generator.forward (model.py:520-533 of the reference, ~250
// kernel launches) by a fixed sequence of ~3 launches per octave:
//     styles (all layers) -> demod (all layers) -> const input
//     conv1 GEMM [+ToRGB in its epilogue] -> rgb combine
//     per octave: up-conv GEMM (4 polyphase sub-problems) -> FIR+noise+bias+lrelu+modulate
//                 -> conv GEMM [+ToRGB] -> rgb combine (+ 2x up-sampled skip)
// Activations are NHWC bf16 and are stored ALREADY MODULATED by the style of the conv that will
// consume them, so every conv is a batch-shared-weight GEMM.  The host owns all memory: the plan
// only holds offsets into the caller's workspace and (cached) TMA descriptors.
#include <string>
#include <vector>

#include "synth_plan.cuh"

using namespace sg2;
using namespace sg2plan;

namespace sg2plan {

// choose the spatial tile of a PH x PW plane: TH*TW <= 128 with the fewest tiles, wide tiles preferred
void choose_tile(int PH, int PW, int maxB, int &TH, int &TW, int &NB) {
    if (PH * PW <= kBlockM) {
        TH = PH; TW = PW;
        NB = std::max(1, std::min(maxB, kBlockM / (PH * PW)));
        return;
    }
    NB = 1;
    long best = -1;
    for (int tw = 4; tw <= std::min(PW, kBlockM); ++tw) {
        const int th = std::min(PH, kBlockM / tw);
        if (th < 1) continue;
        const long tiles = (long)((PH + th - 1) / th) * ((PW + tw - 1) / tw);
        if (best < 0 || tiles < best || (tiles == best && tw > TW)) { best = tiles; TH = th; TW = tw; }
    }
}

// custom_taps (plain convolutions only): n_custom rows of {dy, dx, weight tap index} instead of the 3x3 window
int plan_gemm(sg2_synth *S, Layer &L, const int *custom_taps, int n_custom, const int *tap_planes) {
    const int B = S->max_batch;
    GemmParams &g = L.gp;
    memset(&g, 0, sizeof(g));
    const bool fused = L.fused_up;
    const int cin = L.p.cin, cout = fused ? 4 * L.p.cout : L.p.cout;
    const bool up = L.p.upsample != 0 && !fused && !L.ada_up;
    const int r = L.res_in;
    L.n_gemm = cout;
    g.block_k = cin % 64 == 0 ? 64 : 32;
    g.Cin = cin; g.Cout = cout; g.kchunks = cin / g.block_k;
    g.cout_real = L.p.cout; g.up4 = fused ? 1 : 0;
    g.mode = up ? 1 : 0;
    g.nsub = up ? 4 : 1;
    // sub-problems
    for (int s = 0; s < g.nsub; ++s) {
        GemmSub &q = g.sub[s];
        if (!up) {
            q.PH = q.PW = r; q.out_H = q.out_W = fused ? 2 * r : r; q.out_off = 0;
            q.ntaps = 0;
            if (custom_taps) {
                for (int t = 0; t < n_custom; ++t) {
                    q.dy[t] = custom_taps[3 * t]; q.dx[t] = custom_taps[3 * t + 1]; q.wtap[t] = custom_taps[3 * t + 2];
                    q.tap_map[t] = tap_planes ? tap_planes[t] : 0;
                }
                q.ntaps = n_custom;
                g.multi_map = tap_planes ? 1 : 0;
            } else {
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) { q.dy[q.ntaps] = a - 1; q.dx[q.ntaps] = b - 1; q.wtap[q.ntaps] = a * 3 + b; ++q.ntaps; }
            }
        } else {
            // transposed stride-2 conv: T[2i+a, 2j+b] += x[i,j] * w[a,b]; plane (py,px) holds T[2y+py, 2x+px]
            const int py = s >> 1, px = s & 1;
            q.PH = r + 1 - py; q.PW = r + 1 - px; q.out_H = q.out_W = r + 1;
            q.ntaps = 0;
            for (int a = py; a < 3; a += 2)
                for (int b = px; b < 3; b += 2) { q.dy[q.ntaps] = -(a - py) / 2; q.dx[q.ntaps] = -(b - px) / 2; q.wtap[q.ntaps] = a * 3 + b; ++q.ntaps; }
        }
        choose_tile(q.PH, q.PW, B, q.TH, q.TW, q.NB);
    }
    // BLOCK_N: as wide as possible while the launch still has >= 2 waves of tiles, and the staged
    // per-sample epilogue parameters fit (NB * BLOCK_N <= 512)
    int best_n = 0;
    long best_tiles = 0;
    for (int bn : {256, 128, 64, 32, 16}) {
        if (cout % bn) continue;
        bool fits = true;
        long tiles = 0;
        for (int s = 0; s < g.nsub; ++s) {
            const GemmSub &q = g.sub[s];
            if (q.NB * bn > 512) fits = false;
            tiles += (long)((q.PH + q.TH - 1) / q.TH) * ((q.PW + q.TW - 1) / q.TW) * ((B + q.NB - 1) / q.NB) * (cout / bn);
        }
        if (!fits) continue;
        best_n = bn;
        best_tiles = tiles;
        if (tiles >= 2L * S->sms || bn <= 64) break;
    }
    SG2_REQUIRE(best_n > 0, SG2_ERR_UNSUPPORTED, "engine: no BLOCK_N for Cout=%d", cout);
    g.block_n = L.block_n = best_n;
    g.n_tiles_n = cout / best_n;
    static const char *envk = getenv("SG2_GEMM_KPACK");
    g.kpack = (g.block_k == 64 && g.kchunks % 2 == 0 && (envk ? atoi(envk) == 2 : best_n <= 128)) ? 2 : 1;
    // cta_group::2 (CTA pairs share the weight tile, synth_gemm2.cu).  SG2_GEMM_2SM=0 off, 2 = every layer,
    // 1 / unset = where it measured faster on B200 (profiles/experiments/README.md): plain convs with at
    // least a wave of tiles and BLOCK_N >= 128 or a deep K (-7..-11 %), up-sampling convs only at
    // BLOCK_N = 256 (short-K tiles lose more to the pair hand-shake than they gain: the 64-channel
    // 512^2 layer of the 1024^2 network is 23 % slower as a pair).
    static const char *env2 = getenv("SG2_GEMM_2SM");
    const int mode2 = env2 ? atoi(env2) : 1;
    const bool auto2 = up ? (best_n > 128 && best_tiles >= 4L * S->sms)
                          : (best_tiles >= (long)S->sms && (best_n >= 128 || cin >= 256));
    L.two_sm = best_n >= 32 && g.block_k == 64 && (mode2 == 2 || (mode2 == 1 && auto2) || fused);
    // cta_group::2 with resident weights: each CTA of the pair keeps its half of all 9 * kchunks weight tiles in shared
    // memory next to >= 4 activation stages.  For the fused up-sampling conv of the last octave (64 -> 4 x 32) this removes
    // the weight stream, which is 1/3 of the layer's L2 -> SM traffic.  SG2_GEMM_RES2=0 switches it off.
    static const char *envr2 = getenv("SG2_GEMM_RES2");
    g.resident2 = 0;
    if (L.two_sm && !up && !custom_taps && g.n_tiles_n == 1 && (envr2 ? atoi(envr2) != 0 : fused)) {     // SG2_GEMM_RES2=1: every pair layer that fits
        const int half_b = 9 * g.kchunks * (best_n / 2) * g.block_k * 2;
        const int a_stage = kBlockM * g.block_k * 2 * g.kpack;
        if (half_b + 4 * a_stage <= kGemm2RingBytes) g.resident2 = 1;
    }

    // Resident weights (the narrow, high-resolution tail of the 512^2 / 1024^2 networks): when all 9*Cin*Cout
    // bf16 weights fit in shared memory next to >= 2 activation stages, load them once per CTA and stream
    // only activations -- one slab per distinct dx with the dy halo inside it, so a 3x3 tile costs 3 TMA
    // loads and ONE barrier round trip per K chunk instead of 9.  SG2_GEMM_RESIDENT=0 switches it off.
    static const char *envr = getenv("SG2_GEMM_RESIDENT");
    bool nb1 = true;
    for (int s = 0; s < g.nsub; ++s) nb1 = nb1 && g.sub[s].NB == 1;
    if ((!envr || atoi(envr) != 0) && !fused && !tap_planes && g.n_tiles_n == 1 && nb1 && r >= 16 && cin % 32 == 0) {
        const int resb = 18 * cin * best_n;
        int pick_bk = 0, pick_stage = 0;
        for (int bk : {64, 32}) {
            if (cin % bk) continue;
            int stage = 0;
            for (int s = 0; s < g.nsub; ++s) {
                const GemmSub &q = g.sub[s];
                int dxs[3], ndx = 0, dymin = 0, dymax = 0;
                for (int t = 0; t < q.ntaps; ++t) {
                    bool seen = false;
                    for (int i = 0; i < ndx; ++i) seen = seen || dxs[i] == q.dx[t];
                    if (!seen) dxs[ndx++] = q.dx[t];
                    dymin = std::min(dymin, q.dy[t]); dymax = std::max(dymax, q.dy[t]);
                }
                stage = std::max(stage, ndx * (16 + dymax - dymin) * 8 * bk * 2);
            }
            stage = (stage + 1023) & ~1023;
            const int nst = (kGemmRingBytes - resb) / stage;
            if (resb < kGemmRingBytes && (nst >= 3 || (nst >= 2 && bk == 32))) { pick_bk = bk; pick_stage = stage; break; }
        }
        if (pick_bk) {
            static const char *envm = getenv("SG2_GEMM_MMA2");
            // opt-in (SG2_GEMM_MMA2=1): correct on the <= 256^2 networks, faults on the 32/64-channel 1024^2 tail
            // (tile-split epilogue + SWIZZLE_64B operands) -- kept for the next round, see profiles/experiments
            g.mma2 = (envm && atoi(envm) != 0) ? 1 : 0;
            g.resident = 1; g.resb_bytes = resb; g.stage_bytes = pick_stage;
            g.resident2 = 0;
            g.block_k = pick_bk; g.kchunks = cin / pick_bk; g.kpack = 1;
            L.two_sm = false;
            for (int s = 0; s < g.nsub; ++s) {
                GemmSub &q = g.sub[s];
                q.TH = 16; q.TW = 8; q.NB = 1;
                int dymin = 0, dymax = 0;
                q.nslab = 0;
                for (int t = 0; t < q.ntaps; ++t) {
                    bool seen = false;
                    for (int i = 0; i < q.nslab; ++i) seen = seen || q.slab_dx[i] == q.dx[t];
                    if (!seen) q.slab_dx[q.nslab++] = q.dx[t];
                    dymin = std::min(dymin, q.dy[t]); dymax = std::max(dymax, q.dy[t]);
                }
                q.slab_dy0 = dymin;
                q.slab_rows = q.TH + dymax - dymin;
                const int row_bytes = pick_bk * 2, slab_bytes = q.slab_rows * q.TW * row_bytes;
                for (int t = 0; t < q.ntaps; ++t) {
                    int sl = 0;
                    while (q.slab_dx[sl] != q.dx[t]) ++sl;
                    q.tap_aoff[t] = sl * slab_bytes + (q.dy[t] - dymin) * q.TW * row_bytes;
                }
            }
        }
    }
    // Merged polyphase walk (synth_gemm2p.cu): the four planes of the transposed conv share one tile walk, one activation
    // load per K chunk and a 2-CTA weight split.  Default: the Cout = 128 layer (256 -> 128 at 128^2 -> 256^2), which the
    // separate-planes plan holds at 50 % tensor pipe with 2x the algorithmic DRAM traffic; SG2_POLY4=0 never, 2 = every
    // transposed conv whose shape qualifies.
    static const char *envp4 = getenv("SG2_POLY4");
    const int p4mode = envp4 ? atoi(envp4) : 1;
    g.poly4 = 0;
    if (up && p4mode != 0 && cin % 64 == 0 && r >= 16 && (cout % 128 == 0 || cout == 64) && (p4mode == 2 || cout == 128)) {
        g.poly4 = 1;
        g.block_k = 64; g.kchunks = cin / 64; g.kpack = 1; g.resident = 0; g.mma2 = 0; g.resident2 = 0;
        g.block_n = L.block_n = cout % 128 == 0 ? 128 : 64;
        g.n_tiles_n = cout / g.block_n;
        L.two_sm = true;
        int j = 0;
        for (int s = 3; s >= 0; --s) {
            GemmSub &q = g.sub[s];
            q.TH = 16; q.TW = 8; q.NB = 1;
            for (int t = 0; t < q.ntaps; ++t, ++j) {
                g.m_ph[j] = s;
                g.m_aoff[j] = (q.dx[t] == 0 ? 0 : (16 + 1) * 8 * 128) + (q.dy[t] + 1) * 8 * 128;
                g.m_wtap[j] = q.wtap[t];
                g.m_first[j] = t == 0;
                g.m_last[j] = t == q.ntaps - 1;
            }
        }
    }
    // transposed conv: optionally walk the four polyphase sub-problems segment by segment so that their common input
    // stays in L2 (ncu: the phases re-read it from DRAM, 4x the algorithmic input bytes).  Measured: no gain at 256^2
    // (the layer is bound by L2->SM operand delivery, not DRAM) and a loss on the 1024^2 tail -> opt-in, SG2_GEMM_SEG=n.
    static const char *envs = getenv("SG2_GEMM_SEG");
    g.seg_tiles = up ? (envs ? atoi(envs) : 0) : 0;
    g.nseg = 1;
    static const char *envd = getenv("SG2_GEMM_DBG");
    g.dbg = envd ? atoi(envd) : 0;
    static const char *enva = getenv("SG2_GEMM_EPI_ALT");
    bool small = best_n <= 64;
    for (int s = 0; s < g.nsub; ++s) small = small && g.sub[s].NB * best_n <= 256;
    g.epi_alt = (!L.two_sm && small && (!enva || atoi(enva) != 0)) ? 1 : 0;
    return SG2_OK;
}

void finalize_tiles(GemmParams &g, int B) {
    int t = 0;
    for (int s = 0; s < g.nsub; ++s) {
        GemmSub &q = g.sub[s];
        q.tiles_x = (q.PW + q.TW - 1) / q.TW;
        q.tiles_y = (q.PH + q.TH - 1) / q.TH;
        q.tiles_b = (B + q.NB - 1) / q.NB;
        q.tile_begin = t;
        t += q.tiles_x * q.tiles_y * q.tiles_b * g.n_tiles_n;
    }
    g.total_tiles = t;
    g.B = B;
}

int encode_maps(sg2_synth *S, Layer &L, const __nv_bfloat16 *x, const __nv_bfloat16 *wp, int B) {
    EncodeTiledFn enc = get_encode();
    SG2_REQUIRE(enc, SG2_ERR_CUDA, "engine: cuTensorMapEncodeTiled entry point not available");
    const int C = L.p.cin, r = L.res_in;
    const CUtensorMapSwizzle swz = L.gp.block_k == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    for (int s = 0; s < L.gp.nsub; ++s) {
        const GemmSub &q = L.gp.sub[s];
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)r, (cuuint64_t)r, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)r * C * 2, (cuuint64_t)r * r * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)L.gp.block_k, (cuuint32_t)q.TW,
                             (cuuint32_t)(L.gp.poly4 ? q.TH + 1 : (L.gp.resident ? q.slab_rows : q.TH)), (cuuint32_t)q.NB};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult rc = enc(&L.tmA[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)x, dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SG2_REQUIRE(rc == CUDA_SUCCESS, SG2_ERR_CUDA, "engine: cuTensorMapEncodeTiled(A) failed with %d (C=%d r=%d box %d,%d,%d)",
                    (int)rc, C, r, q.TW, q.TH, q.NB);
    }
    {
        const int n_gemm = L.n_gemm ? L.n_gemm : L.p.cout;
        cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)n_gemm, 9};
        cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)C * n_gemm * 2};
        cuuint32_t box[3] = {(cuuint32_t)L.gp.block_k, (cuuint32_t)(L.two_sm ? L.block_n / 2 : L.block_n), 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult rc = enc(&L.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void *)wp, dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SG2_REQUIRE(rc == CUDA_SUCCESS, SG2_ERR_CUDA, "engine: cuTensorMapEncodeTiled(B) failed with %d", (int)rc);
    }
    return SG2_OK;
}

int encode_dxs_maps(Layer &L, const __nv_bfloat16 *x, const __nv_bfloat16 *wp, int B) {
    EncodeTiledFn enc = get_encode();
    SG2_REQUIRE(enc, SG2_ERR_CUDA, "engine: cuTensorMapEncodeTiled entry point not available");
    const int C = L.p.cin, r = L.res_out, N = 3 * L.p.cout;
    const CUtensorMapSwizzle swz = C == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    cuuint32_t es[4] = {1, 1, 1, 1};
    {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)r, (cuuint64_t)r, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)r * C * 2, (cuuint64_t)r * r * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)C, 32, 6, 1};
        CUresult rc = enc(&L.tmDA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SG2_REQUIRE(rc == CUDA_SUCCESS, SG2_ERR_CUDA, "engine: cuTensorMapEncodeTiled(dx-stacked A) failed with %d", (int)rc);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)N, 3};
        cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)C * N * 2};
        cuuint32_t box[3] = {(cuuint32_t)C, (cuuint32_t)N, 1};
        CUresult rc = enc(&L.tmDB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void *)wp, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SG2_REQUIRE(rc == CUDA_SUCCESS, SG2_ERR_CUDA, "engine: cuTensorMapEncodeTiled(dx-stacked B) failed with %d", (int)rc);
    }
    return SG2_OK;
}

inline int fir_store_mode() {
    static const char *e = getenv("SG2_FIR_STORE");
    return e ? atoi(e) : 0;
}

// TMA views of the 4 polyphase planes [(py,px)][B][r+1][r+1][C]: valid extent (r+1-py) x (r+1-px), so the
// never-written last row/column of the odd planes reads as zero; box = 10 x 6 pixels x 64 channels
int encode_fir_maps(Layer &L, const __nv_bfloat16 *T, const __nv_bfloat16 *out, int B) {
    EncodeTiledFn enc = get_encode();
    SG2_REQUIRE(enc, SG2_ERR_CUDA, "engine: cuTensorMapEncodeTiled entry point not available");
    const int C = L.p.cout, r = L.res_in, P = r + 1;
    const int cbw = C % 64 == 0 ? 64 : 32;            // column block width of the FIR kernel (synth_fir.cu)
    const CUtensorMapSwizzle swz = cbw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const size_t plane = (size_t)B * P * P * C;
    for (int s = 0; s < 4; ++s) {
        const int py = s >> 1, px = s & 1;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)(P - px), (cuuint64_t)(P - py), (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)P * C * 2, (cuuint64_t)P * P * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)cbw, 6, 10, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult rc = enc(&L.tmT[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)(T + plane * s), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SG2_REQUIRE(rc == CUDA_SUCCESS, SG2_ERR_CUDA, "engine: cuTensorMapEncodeTiled(T plane) failed with %d", (int)rc);
    }
    {
        const int R = 2 * r;
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)R, (cuuint64_t)R, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)R * C * 2, (cuuint64_t)R * R * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)cbw, 8, (cuuint32_t)(fir_store_mode() ? 4 : 16), 1};   // a column block of the tile / of one warp
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult rc = enc(&L.tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void *)out, dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SG2_REQUIRE(rc == CUDA_SUCCESS, SG2_ERR_CUDA, "engine: cuTensorMapEncodeTiled(FIR out) failed with %d", (int)rc);
    }
    return SG2_OK;
}

// after every launch: optional timing event; with SG2_SYNTH_DEBUG=1 also a sync that names the
// first failing kernel (debug only -- never set while capturing a CUDA graph)
int rec(sg2_synth *S, cudaStream_t st, const char *what) {
    if (S->events && S->events_used < S->n_events) cudaEventRecord(S->events[S->events_used++], st);
    static const bool debug = getenv("SG2_SYNTH_DEBUG") != nullptr;
    if (debug) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { set_error("engine: kernel '%s' failed: %s", what, cudaGetErrorString(e)); return SG2_ERR_CUDA; }
    }
    return SG2_OK;
}

}  // namespace sg2plan

static int synth_create_impl(sg2_synth **plan, int size, int style_dim, int max_batch,
                             const sg2_conv_params *layers, int n_layers, const float *const_input,
                             const float *blur_taps_host, bool ada) {
    SG2_REQUIRE(plan, SG2_ERR_BAD_ARG, "synth_create: null plan pointer");
    *plan = nullptr;
    int log_size = 0;
    while ((1 << log_size) < size) ++log_size;
    SG2_REQUIRE((1 << log_size) == size && size >= 8 && size <= 1024, SG2_ERR_UNSUPPORTED,
                "synth_create: size must be a power of two in [8, 1024], got %d", size);
    SG2_REQUIRE(style_dim >= 32 && style_dim <= 512 && style_dim % 32 == 0, SG2_ERR_UNSUPPORTED,
                "synth_create: style_dim must be a multiple of 32 in [32, 512], got %d", style_dim);
    SG2_REQUIRE(max_batch >= 1 && max_batch <= 4096, SG2_ERR_BAD_ARG, "synth_create: bad max_batch %d", max_batch);
    SG2_REQUIRE(layers && n_layers == 2 + 3 * (log_size - 2) && n_layers <= kMaxJobs, SG2_ERR_BAD_ARG,
                "synth_create: expected %d layer rows, got %d", 2 + 3 * (log_size - 2), n_layers);
    SG2_REQUIRE(const_input && blur_taps_host, SG2_ERR_BAD_ARG, "synth_create: null pointer");
    sg2_synth *S = new sg2_synth();
    S->size = size; S->style_dim = style_dim; S->max_batch = max_batch; S->log_size = log_size;
    S->n_latent = 2 * log_size - 2; S->num_layers = 2 * (log_size - 2) + 1;
    S->const_input = const_input;
    S->sms = 148;
    S->ada = ada;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) { S->kf[a * 4 + b] = blur_taps_host[(3 - a) * 4 + (3 - b)]; S->kf_raw[a * 4 + b] = blur_taps_host[a * 4 + b]; }
    {   // SmoothUpsample: tap a of output parity p reads input cell i + off[p][a] (clamped): fold the 16 taps onto the 3x3 cells
        const int sel[2][4] = {{0, 0, 1, 1}, {0, 1, 1, 2}};
        for (int ph = 0; ph < 4; ++ph) {
            for (int t = 0; t < 9; ++t) S->wp_ada[ph][t] = 0.f;
            for (int a = 0; a < 4; ++a)
                for (int b = 0; b < 4; ++b) S->wp_ada[ph][sel[ph >> 1][a] * 3 + sel[ph & 1][b]] += blur_taps_host[a * 4 + b];
        }
    }
    const int B = max_batch;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
    size_t max_act = 0, max_T = 0, max_part = 0;
    int prev_cout = 0;
    for (int i = 0; i < n_layers; ++i) {
        Layer L;
        L.p = layers[i];
        L.rgb = (i % 3 == 1) || (i >= 2 && (i - 2) % 3 == 2);
        const bool ok_ptrs = L.p.weight && L.p.mod_weight && L.p.mod_bias && L.p.act_bias && (L.rgb || L.p.noise_weight);
        if (!ok_ptrs) { delete S; set_error("synth_create: layer %d has a null parameter pointer", i); return SG2_ERR_BAD_ARG; }
        const int expect_k = L.rgb ? 1 : 3;
        if (L.p.ksize != expect_k || (L.rgb && L.p.cout != 3) || L.p.cin % 8 || (!L.rgb && L.p.cout % 8) ||
            L.p.latent_index < 0 || L.p.latent_index >= S->n_latent || (i > 0 && !L.rgb && L.p.cin != prev_cout)) {
            delete S;
            set_error("synth_create: layer %d (cin %d cout %d k %d up %d) does not fit the StyleGAN2 synthesis topology",
                      i, L.p.cin, L.p.cout, L.p.ksize, L.p.upsample);
            return SG2_ERR_UNSUPPORTED;
        }
        L.res_out = L.p.resolution;
        L.res_in = L.p.upsample ? L.res_out / 2 : L.res_out;
        L.style = take(sizeof(float) * B * L.p.cin);
        if (L.rgb) {
            L.rgbw = take(sizeof(float) * 3 * L.p.cin);
        } else {
            if (L.p.cin % 32 || L.p.cout % 16 || (L.p.upsample && !ada && L.p.cout % 32)) {
                delete S;
                set_error("engine: styled conv %d with Cin %d / Cout %d needs Cin %% 32 == 0 and Cout %% 16 == 0 "
                          "(Cout %% 32 == 0 for up-sampling layers)", i, L.p.cin, L.p.cout);
                return SG2_ERR_UNSUPPORTED;
            }
            // Fused up-sampling conv (GemmParams::up4): where the polyphase transposed conv + FIR pass is bandwidth-bound --
            // the narrow octaves, Cout <= 64: FLOP/byte 96..192 -- the blur is folded into the weights and
            // the (2r+1)^2 intermediate never exists.  SG2_UPFUSED=0: never, 2: every up-sampling layer that qualifies.
            static const char *envf = getenv("SG2_UPFUSED");
            const int fmode = envf ? atoi(envf) : 1;
            L.fused_up = !ada && L.p.upsample && fmode != 0 && L.p.cin % 64 == 0 && L.p.cout % 32 == 0 && L.res_out >= 32 &&
                         (L.p.cout <= 64 || fmode == 2);
            L.ada_up = ada && L.p.upsample;
            L.wp = take(sizeof(__nv_bfloat16) * 9 * L.p.cout * L.p.cin * (L.fused_up ? 4 : 1));
            L.wsq = take(sizeof(float) * L.p.cin * L.p.cout);
            L.demod = take(sizeof(float) * B * L.p.cout);
            int rc = plan_gemm(S, L);
            if (rc) { delete S; return rc; }
            finalize_tiles(L.gp, B);
            // dx-stacked kernel for the narrowest plain conv (32 -> 32 at 1024^2): N = 3 * Cout, three activation views per
            // tile instead of nine.  Measured (round 2, B = 32): 1.83 -> 1.45 ms at Cout = 32; at Cout = 64 the three-fold
            // accumulator read-back costs what the saved operand reads gain (0.94 -> 0.98 ms), so that layer stays on the
            // tap-by-tap kernel.  SG2_DXS=0: never, 2: also Cout = 64.
            static const char *envx = getenv("SG2_DXS");
            const int xmode = envx ? atoi(envx) : 1;
            L.dxs = !L.p.upsample && xmode != 0 && (L.p.cin == 32 || L.p.cin == 64) &&
                    (L.p.cout == 32 || (L.p.cout == 64 && xmode == 2)) && L.res_out >= 64;
            if (L.dxs) {
                memset(&L.dp, 0, sizeof(L.dp));
                L.dp.R = L.res_out; L.dp.Cin = L.p.cin; L.dp.Cout = L.p.cout;
                L.dp.tiles_x = (L.res_out + 29) / 30; L.dp.tiles_y = L.res_out / 4;
                L.block_n = 3 * L.p.cout;
            }
            prev_cout = L.p.cout;
            max_act = std::max(max_act, sizeof(__nv_bfloat16) * (size_t)B * L.res_in * L.res_in * L.p.cin);
            max_act = std::max(max_act, sizeof(__nv_bfloat16) * (size_t)B * L.res_out * L.res_out * L.p.cout);
            if (L.ada_up)
                max_T = std::max(max_T, sizeof(__nv_bfloat16) * (size_t)B * L.res_in * L.res_in * L.p.cout);
            else if (L.p.upsample && !L.fused_up)
                max_T = std::max(max_T, sizeof(__nv_bfloat16) * 4 * (size_t)B * (L.res_in + 1) * (L.res_in + 1) * L.p.cout);
            else if (!L.p.upsample)
                max_part = std::max(max_part, sizeof(float) * (L.two_sm ? kGemm2EpiGroups : 2) * (size_t)L.gp.n_tiles_n * B * 3 * L.res_out * L.res_out);
        }
        S->layers.push_back(L);
    }
    S->off_act[0] = take(max_act); S->off_act[1] = take(max_act);
    S->off_T = take(max_T);
    S->off_rgb[0] = take(sizeof(float) * (size_t)B * 3 * size * size);
    S->off_rgb[1] = take(sizeof(float) * (size_t)B * 3 * size * size);
    S->off_part = take(max_part);
    S->off_toeplitz = take(128 * 256 * 2);
    S->toeplitz.resize(128 * 256);
    build_fir_toeplitz(S->toeplitz.data(), S->kf);
    S->fir_simt = getenv("SG2_FIR_SIMT") != nullptr;
    S->ws_bytes = off;
    // description (one line per launch of a forward pass, in launch order)
    char line[512];
    std::string d;
    int k = 0;
    auto add = [&](const char *kind, const char *what, double flops, double bytes, int tiles, int bn) {
        snprintf(line, sizeof(line), "%d %s %s flops=%.6g bytes=%.6g tiles=%d block_n=%d\n", k++, kind, what, flops, bytes, tiles, bn);
        d += line;
    };
    add("styles", "all", 0, 0, 0, 0);
    add("demod", "all", 0, 0, 0, 0);
    add("const_input", "input", 0, 0, 0, 0);
    for (size_t i = 0; i < S->layers.size(); ++i) {
        const Layer &L = S->layers[i];
        char name[64];
        snprintf(name, sizeof(name), "L%zu_%dx%d_%d->%d%s", i, L.res_out, L.res_out, L.p.cin, L.p.cout,
                 L.p.upsample ? (L.fused_up ? "_upfused" : "_up") : (L.dxs ? "_dxs" : ""));
        const double px_in = (double)L.res_in * L.res_in, px_out = (double)L.res_out * L.res_out;
        if (L.rgb) {
            add("rgb_combine", name, 2.0 * 3 * L.p.cin * px_out, 0, 0, 0);
        } else {
            // algorithmic FLOPs per image in the reference's formulation
            const double fl = 2.0 * 9 * L.p.cin * L.p.cout * (L.p.upsample ? px_in : px_out);
            // algorithmic bytes: the input read once + what the kernel writes: the (2r+1)^2 intermediate of an up-sampling
            // layer, the output plane of a plain one -- except the LAST conv, which stores no activation at all, only
            // the three fp32 ToRGB planes
            const bool last = i + 2 >= S->layers.size();
            const double out_b = L.ada_up ? 2.0 * px_in * L.p.cout
                                 : (L.p.upsample && !L.fused_up) ? 2.0 * (2 * L.res_in + 1.0) * (2 * L.res_in + 1.0) * L.p.cout
                                                                 : (last ? 4.0 * 3 * px_out : 2.0 * px_out * L.p.cout);
            add("gemm", name, fl, 2.0 * px_in * L.p.cin + out_b, L.dxs ? L.dp.tiles_x * L.dp.tiles_y * B : L.gp.total_tiles, L.block_n);
            if (L.ada_up) add("smoothup", name, 0, 2.0 * (px_in + px_out) * L.p.cout, 0, 0);
            else if (L.p.upsample && !L.fused_up) add("upfir", name, 0, 2.0 * ((2 * L.res_in + 1.0) * (2 * L.res_in + 1.0) + px_out) * L.p.cout, 0, 0);
        }
    }
    S->description = d;
    *plan = S;
    return SG2_OK;
}

extern "C" int sg2_synth_create(sg2_synth **plan, int size, int style_dim, int max_batch,
                                const sg2_conv_params *layers, int n_layers, const float *const_input,
                                const float *blur_taps_host) {
    return synth_create_impl(plan, size, style_dim, max_batch, layers, n_layers, const_input, blur_taps_host, false);
}
// The stylegan2_ada decoder (restyle-encoder/models/stylegan2_ada/generator.py:55-204) on the same plan: same layer table
// (first_block.conv1, first_block.torgb, then per block conv0 (upsample = 1), conv1, torgb), resample_taps_host = the 4x4
// SmoothUpsample kernel (utils.py:76-83, sum 1).
extern "C" int sg2_synth_create_ada(sg2_synth **plan, int size, int w_dim, int max_batch,
                                    const sg2_conv_params *layers, int n_layers, const float *const_input,
                                    const float *resample_taps_host) {
    return synth_create_impl(plan, size, w_dim, max_batch, layers, n_layers, const_input, resample_taps_host, true);
}

extern "C" void sg2_synth_destroy(sg2_synth *p) { delete p; }
extern "C" int64_t sg2_synth_workspace_bytes(const sg2_synth *p) { return p ? (int64_t)p->ws_bytes : 0; }
extern "C" int sg2_synth_describe(const sg2_synth *p, char *buf, int buflen) {
    if (!p || !buf || buflen <= 0) return 0;
    const int n = std::min<int>((int)p->description.size(), buflen - 1);
    memcpy(buf, p->description.data(), n);
    buf[n] = 0;
    return n;
}

extern "C" int sg2_synth_set_pooled_output(sg2_synth *p, float *pooled, int factor, int keep_full) {
    SG2_REQUIRE(p, SG2_ERR_BAD_ARG, "synth_set_pooled_output: null plan");
    SG2_REQUIRE(factor == 0 || ((factor == 2 || factor == 4) && pooled && p->size % (4 * factor) == 0), SG2_ERR_BAD_ARG,
                "synth_set_pooled_output: factor must be 0 (off), 2 or 4 with a buffer, and divide the image size (%d)", p->size);
    SG2_REQUIRE(!p->train || factor == 0, SG2_ERR_UNSUPPORTED, "synth_set_pooled_output: not in training mode (the backward expects the full image gradient)");
    p->pool = factor;
    p->pool_out = factor ? pooled : nullptr;
    p->pool_keep_full = factor == 0 || keep_full != 0;
    return SG2_OK;
}

extern "C" int sg2_synth_set_profile_events(sg2_synth *p, void **events, int n_events) {
    SG2_REQUIRE(p, SG2_ERR_BAD_ARG, "synth_set_profile_events: null plan");
    p->events = reinterpret_cast<cudaEvent_t *>(events);
    p->n_events = events ? n_events : 0;
    p->events_used = 0;
    return SG2_OK;
}
extern "C" int sg2_synth_profile_events_used(const sg2_synth *p) { return p ? p->events_used : 0; }

extern "C" int sg2_synth_pack(sg2_synth *S, void *workspace, sg2_stream_t stream) {
    SG2_REQUIRE(S && workspace, SG2_ERR_BAD_ARG, "synth_pack: null pointer");
    SG2_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, SG2_ERR_BAD_ARG, "synth_pack: workspace must be 1 KiB aligned");
    int dev = 0, major = 0;
    SG2_CUDA_OK(cudaGetDevice(&dev));
    SG2_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    SG2_REQUIRE(major == 10, SG2_ERR_NO_DEVICE, "engine: needs an sm_100 device (tcgen05/TMEM), found compute capability %d.x", major);
    S->sms = sm_count();
    cudaStream_t st = as_stream(stream);
    uint8_t *ws = static_cast<uint8_t *>(workspace);
    for (Layer &L : S->layers) {
        if (L.rgb) {
            int rc = launch_pack_rgb_weight((float *)(ws + L.rgbw), L.p.weight, 3 * L.p.cin, 1.0f / sqrtf((float)L.p.cin), st);
            if (rc) return rc;
        } else {
            // (a fused up-sampling layer needs wsq from the plain weights and then its own composite pack over wp)
            // (stylegan2_ada: no equalised-lr scale on the conv weights -- the demodulation normalises them, utils.py:120-137)
            int rc = launch_pack_conv_weight((__nv_bfloat16 *)(ws + L.wp), (float *)(ws + L.wsq), L.p.weight, L.p.cin,
                                             L.p.cout, 9, S->ada ? 1.0f : 1.0f / sqrtf((float)L.p.cin * 9), st);
            if (rc) return rc;
            if (L.fused_up) {
                rc = launch_pack_upfused_weight((__nv_bfloat16 *)(ws + L.wp), L.p.weight, L.p.cin, L.p.cout,
                                                1.0f / sqrtf((float)L.p.cin * 9), S->kf, st);
                if (rc) return rc;
            }
        }
    }
    SG2_CUDA_OK(cudaMemcpyAsync(ws + S->off_toeplitz, S->toeplitz.data(), S->toeplitz.size() * 2, cudaMemcpyHostToDevice, st));
    if (S->train) {
        int rc = train_pack(S, ws, st);
        if (rc) return rc;
    }
    S->cached_ws = nullptr;   // descriptors are rebuilt by the next forward
    S->bcached_ws = nullptr;
    return SG2_OK;
}

extern "C" int sg2_synth_forward(sg2_synth *S, void *workspace, const float *latent, int64_t B64,
                                 const float *const *noise, const int64_t *noise_bstride, float *image,
                                 sg2_stream_t stream) {
    SG2_REQUIRE(S, SG2_ERR_BAD_ARG, "synth_forward: null plan");
    SG2_REQUIRE(B64 >= 0 && B64 <= S->max_batch, SG2_ERR_BAD_ARG, "synth_forward: batch %lld exceeds the plan's max_batch %d",
                (long long)B64, S->max_batch);
    if (B64 == 0) return SG2_OK;                 // empty batch: nothing to launch (tensor pointers may be null)
    SG2_REQUIRE(workspace && latent && noise && noise_bstride, SG2_ERR_BAD_ARG, "synth_forward: null pointer");
    SG2_REQUIRE(image || (S->pool && !S->pool_keep_full), SG2_ERR_BAD_ARG, "synth_forward: null image (allowed only with a pooled-only output)");
    const int B = (int)B64;
    cudaStream_t st = as_stream(stream);
    uint8_t *ws = static_cast<uint8_t *>(workspace);
    __nv_bfloat16 *act[2] = {(__nv_bfloat16 *)(ws + S->off_act[0]), (__nv_bfloat16 *)(ws + S->off_act[1])};
    __nv_bfloat16 *Tbuf = (__nv_bfloat16 *)(ws + S->off_T);
    float *rgbbuf[2] = {(float *)(ws + S->off_rgb[0]), (float *)(ws + S->off_rgb[1])};
    float *part = (float *)(ws + S->off_part);
    S->events_used = 0;

    // where every styled conv reads and writes.  Inference: two ping-pong buffers (plain conv act[0] -> act[1], up-sampling
    // layer act[1] -> act[0]).  Training: every layer writes its own kept buffer, which the next layer reads.
    const size_t nL = S->layers.size();
    std::vector<__nv_bfloat16 *> inp(nL, nullptr), outp(nL, nullptr);
    {
        __nv_bfloat16 *prev = S->train ? (__nv_bfloat16 *)(ws + S->off_act_in) : act[0];
        for (size_t i = 0; i < nL; ++i) {
            Layer &L = S->layers[i];
            if (L.rgb) continue;
            if (S->train) { inp[i] = prev; outp[i] = (__nv_bfloat16 *)(ws + L.keep); prev = outp[i]; }
            else { inp[i] = L.p.upsample ? act[1] : act[0]; outp[i] = L.p.upsample ? act[0] : act[1]; }
        }
    }
    const float *ones = S->train ? (const float *)(ws + S->off_ones) : nullptr;
    // (re)build tile tables + TMA descriptors when the workspace or the batch changed
    if (S->cached_ws != workspace || S->cached_B != B) {
        for (size_t i = 0; i < nL; ++i) {
            Layer &L = S->layers[i];
            if (L.rgb) continue;
            finalize_tiles(L.gp, B);
            const __nv_bfloat16 *x = inp[i];
            int rc = L.dxs ? encode_dxs_maps(L, x, (const __nv_bfloat16 *)(ws + L.wp), B)
                           : encode_maps(S, L, x, (const __nv_bfloat16 *)(ws + L.wp), B);
            if (rc) return rc;
            if (L.p.upsample && !L.fused_up) {
                rc = encode_fir_maps(L, Tbuf, outp[i], B);
                if (rc) return rc;
            }
        }
        S->cached_ws = workspace;
        S->cached_B = B;
    }

    // 1. styles of all layers, demod of all styled convs
    StyleJobs sj; sj.n = 0;
    DemodJobs dj; dj.n = 0;
    int blocks = 0, max_cin = 0, max_cout = 0;
    for (Layer &L : S->layers) {
        StyleJob &j = sj.job[sj.n++];
        j.mod_w = L.p.mod_weight; j.mod_b = L.p.mod_bias; j.out = (float *)(ws + L.style);
        j.cin = L.p.cin; j.latent_index = L.p.latent_index; j.block_begin = blocks;
        blocks += (L.p.cin + kStyleBlockCi - 1) / kStyleBlockCi;
        if (!L.rgb) {
            DemodJob &d = dj.job[dj.n++];
            d.style = (const float *)(ws + L.style); d.wsq = (const float *)(ws + L.wsq); d.demod = (float *)(ws + L.demod);
            d.cin = L.p.cin; d.cout = L.p.cout;
            max_cin = std::max(max_cin, L.p.cin); max_cout = std::max(max_cout, L.p.cout);
        }
    }
    int rc = rec(S, st, "begin");
    if (rc) return rc;
    rc = launch_styles(sj, blocks, latent, B, S->n_latent, S->style_dim, st);
    if (rc) return rc;
    if ((rc = rec(S, st, "styles"))) return rc;
    rc = launch_demod(dj, max_cin, max_cout, B, st);
    if (rc) return rc;
    if ((rc = rec(S, st, "demod"))) return rc;
    // 2. constant input pre-modulated by conv1's style
    Layer &L0 = S->layers[0];
    rc = launch_const_input(inp[0], S->const_input, (const float *)(ws + L0.style), B, L0.p.cin, 16, st);
    if (rc) return rc;
    if ((rc = rec(S, st, "const_input"))) return rc;

    // 3. the layers
    int noise_idx = 0;
    int rgb_cur = -1;      // which rgb buffer holds the running skip image (-1: none yet)
    for (size_t i = 0; i < nL; ++i) {
        Layer &L = S->layers[i];
        if (L.rgb) continue;   // handled together with the conv that feeds it
        // consumer of this conv's output: the next styled conv (if any); ToRGB if the next row is rgb
        Layer *next_conv = nullptr, *rgb = nullptr;
        for (size_t j = i + 1; j < nL; ++j) {
            if (S->layers[j].rgb) { if (j == i + 1) rgb = &S->layers[j]; }
            else { next_conv = &S->layers[j]; break; }
        }
        GemmParams g = L.gp;
        g.demod = (const float *)(ws + L.demod);
        const float act_clamp = S->ada ? 256.0f / 1.41421356237f : 0.f;     // clamp_gain(x, sqrt(2), 256) before the folded gain
        g.clamp = act_clamp;
        const float *nz = noise[noise_idx];
        const int64_t nzs = noise_bstride[noise_idx];
        ++noise_idx;
        if (L.ada_up) {
            // stylegan2_ada: 3x3 conv at the input resolution (x demod) -> Tbuf, then SmoothUpsample + noise + bias + lrelu +
            // clamp + next layer's modulation in one pass (generator.py:198-204)
            SG2_REQUIRE(next_conv, SG2_ERR_BAD_ARG, "engine: up-sampling conv without a consumer");
            g.mode = 1;
            g.out = Tbuf;
            rc = L.two_sm ? launch_modconv_gemm2(g, L.tmA, L.tmB, S->sms, st) : launch_modconv_gemm(g, L.tmA, L.tmB, S->sms, st);
            if (rc) return rc;
            if ((rc = rec(S, st, "gemm(ada up)"))) return rc;
            SmoothUpParams sp;
            sp.T = Tbuf; sp.out = outp[i]; sp.B = B; sp.r = L.res_in; sp.C = L.p.cout;
            sp.noise = nz; sp.noise_bstride = nzs; sp.noise_weight = L.p.noise_weight;
            sp.bias = L.p.act_bias; sp.next_style = (const float *)(ws + next_conv->style);
            sp.clamp = act_clamp;
            memcpy(sp.wp, S->wp_ada, sizeof(sp.wp));
            rc = launch_smooth_up(sp, st);
            if (rc) return rc;
            if ((rc = rec(S, st, "smoothup"))) return rc;
        } else if (L.fused_up) {
            // conv_transpose + blur + noise + bias + lrelu + next layer's modulation in ONE launch: act[1] -> act[0]
            SG2_REQUIRE(next_conv, SG2_ERR_BAD_ARG, "engine: up-sampling conv without a consumer");
            g.noise = nz; g.noise_bstride = nzs; g.noise_weight = L.p.noise_weight;
            g.bias = L.p.act_bias;
            g.next_style = (const float *)(ws + next_conv->style);
            g.out = outp[i];
            rc = L.two_sm ? launch_modconv_gemm2(g, L.tmA, L.tmB, S->sms, st) : launch_modconv_gemm(g, L.tmA, L.tmB, S->sms, st);
            if (rc) return rc;
            if ((rc = rec(S, st, "gemm(upfused)"))) return rc;
        } else if (L.dxs) {
            DxsParams d = L.dp;
            d.B = B; d.total_tiles = d.tiles_x * d.tiles_y * B;
            d.clamp = act_clamp;
            d.demod = (const float *)(ws + L.demod);
            d.noise = nz; d.noise_bstride = nzs; d.noise_weight = L.p.noise_weight;
            d.bias = L.p.act_bias;
            // (training keeps the last layer's activation too: stored with a unit "next style")
            d.next_style = next_conv ? (const float *)(ws + next_conv->style) : ones;
            d.out = (next_conv || S->train) ? outp[i] : nullptr;
            if (rgb) {
                d.rgb_w = (const float *)(ws + rgb->rgbw);
                d.rgb_style = (const float *)(ws + rgb->style);
                d.rgb_part = part;
            }
            // the network's last layer can write the final image from its epilogue (no partial plane, no rgb_combine pass).
            // Measured at 1024^2, B = 32: the conv grows by 0.34 ms (its epilogue is the paced role: every instruction added per
            // pixel shows), rgb_combine saves 0.26 ms -> opt-in only, SG2_DXS_IMAGE=1.
            static const char *envfi = getenv("SG2_DXS_IMAGE");
            const bool fuse_image = rgb && !next_conv && !S->ada && !S->pool && L.p.cout == 32 && envfi && atoi(envfi) != 0;
            if (fuse_image) {
                d.image = image; d.rgb_bias = rgb->p.act_bias; d.prev = rgb_cur >= 0 ? rgbbuf[rgb_cur] : nullptr;
                memcpy(d.kf, S->kf, sizeof(d.kf));
            }
            rc = launch_modconv_dxs(d, L.tmDA, L.tmDB, S->sms, st);
            if (rc) return rc;
            if ((rc = rec(S, st, "gemm(dx-stacked)"))) return rc;
            if (fuse_image) {
                if ((rc = rec(S, st, "rgb_combine (fused into the conv)"))) return rc;
            } else if (rgb) {
                RgbParams rp;
                const bool last = next_conv == nullptr;
                const int dst = rgb_cur == 0 ? 1 : 0;
                rp.out = last ? image : rgbbuf[dst];
                rp.pool = 0; rp.pool_out = nullptr;
                if (last && S->pool) {          // face_pool folded into the last launch (sg2_synth_set_pooled_output)
                    rp.pool = S->pool; rp.pool_out = S->pool_out;
                    if (!S->pool_keep_full) rp.out = nullptr;
                }
                rp.part = part; rp.n_parts = 3 * L.p.cout <= 128 ? 1 : 2; rp.bias = rgb->p.act_bias;
                rp.prev = rgb_cur >= 0 ? rgbbuf[rgb_cur] : nullptr;
                rp.B = B; rp.R = L.res_out;
                rp.clamp = S->ada ? 256.0f : 0.f; rp.smooth = S->ada ? 1 : 0;
                memcpy(rp.kf, S->ada ? S->kf_raw : S->kf, sizeof(rp.kf));
                rc = launch_rgb_combine(rp, S->sms, st);
                if (rc) return rc;
                if ((rc = rec(S, st, "rgb_combine"))) return rc;
                rgb_cur = dst;
            }
        } else if (!L.p.upsample) {
            g.noise = nz; g.noise_bstride = nzs; g.noise_weight = L.p.noise_weight;
            g.bias = L.p.act_bias;
            g.next_style = next_conv ? (const float *)(ws + next_conv->style) : ones;
            g.out = (next_conv || S->train) ? outp[i] : nullptr;
            if (rgb) {
                g.rgb_w = (const float *)(ws + rgb->rgbw);
                g.rgb_style = (const float *)(ws + rgb->style);
                g.rgb_part = part;
            }
            rc = L.two_sm ? launch_modconv_gemm2(g, L.tmA, L.tmB, S->sms, st) : launch_modconv_gemm(g, L.tmA, L.tmB, S->sms, st);
            if (rc) return rc;
            if ((rc = rec(S, st, "gemm(conv)"))) return rc;
            if (rgb) {
                RgbParams rp;
                const bool last = next_conv == nullptr;
                const int dst = rgb_cur == 0 ? 1 : 0;
                rp.out = last ? image : rgbbuf[dst];
                rp.pool = 0; rp.pool_out = nullptr;
                if (last && S->pool) {          // face_pool folded into the last launch (sg2_synth_set_pooled_output)
                    rp.pool = S->pool; rp.pool_out = S->pool_out;
                    if (!S->pool_keep_full) rp.out = nullptr;
                }
                // partial ToRGB planes per N tile: the cta_group::2 kernel has 4 column groups, the single-CTA one 2 (or 1)
                const int parts_per_tile = L.two_sm ? std::min(kGemm2EpiGroups, g.block_n / 32) : (g.epi_alt ? 1 : 2);
                rp.part = part; rp.n_parts = parts_per_tile * g.n_tiles_n; rp.bias = rgb->p.act_bias;
                rp.prev = rgb_cur >= 0 ? rgbbuf[rgb_cur] : nullptr;
                rp.B = B; rp.R = L.res_out;
                rp.clamp = S->ada ? 256.0f : 0.f; rp.smooth = S->ada ? 1 : 0;
                memcpy(rp.kf, S->ada ? S->kf_raw : S->kf, sizeof(rp.kf));
                rc = launch_rgb_combine(rp, S->sms, st);
                if (rc) return rc;
                if ((rc = rec(S, st, "rgb_combine"))) return rc;
                rgb_cur = dst;
            }
        } else {
            // transposed conv -> 4 polyphase planes in Tbuf (demodulated), then FIR + noise + bias + lrelu + modulate
            const long long plane = (long long)B * (L.res_in + 1) * (L.res_in + 1) * L.p.cout;
            for (int s = 0; s < g.nsub; ++s) g.sub[s].out_off = plane * s;
            g.out = Tbuf;
            rc = g.poly4 ? launch_modconv_gemm2_poly4(g, L.tmA[0], L.tmB, S->sms, st)
                         : (L.two_sm ? launch_modconv_gemm2(g, L.tmA, L.tmB, S->sms, st) : launch_modconv_gemm(g, L.tmA, L.tmB, S->sms, st));
            if (rc) return rc;
            if ((rc = rec(S, st, "gemm(up)"))) return rc;
            SG2_REQUIRE(next_conv, SG2_ERR_BAD_ARG, "engine: up-sampling conv without a consumer");
            UpfirParams up;
            up.T = Tbuf; up.plane_stride = plane; up.out = outp[i]; up.r = L.res_in; up.C = L.p.cout;
            up.noise = nz; up.noise_bstride = nzs; up.noise_weight = L.p.noise_weight;
            up.bias = L.p.act_bias; up.next_style = (const float *)(ws + next_conv->style);
            memcpy(up.kf, S->kf, sizeof(up.kf));
            if (S->fir_simt || L.p.cout % 32 != 0) {   // odd widths: SIMT stencil
                rc = launch_upfir(up, B, st);
            } else {
                UpfirTcParams tp;
                tp.out = outp[i]; tp.r = L.res_in; tp.C = L.p.cout; tp.B = B;
                tp.cbw = L.p.cout % 64 == 0 ? 64 : 32;
                tp.nsamp = L.p.cout % 128 == 0 ? 1 : 128 / L.p.cout;      // 64 channels: 2 samples per tile, 32: 4
                tp.tiles_c = L.p.cout % 128 == 0 ? L.p.cout / 128 : 1;
                tp.tiles_x = (2 * L.res_in + 7) / 8; tp.tiles_y = (2 * L.res_in + 15) / 16;
                tp.total_tiles = tp.tiles_x * tp.tiles_y * tp.tiles_c * ((B + tp.nsamp - 1) / tp.nsamp);
                static const char *envd = getenv("SG2_FIR_DBG");
                tp.store_mode = fir_store_mode();
                tp.dbg = envd ? atoi(envd) : 0;
                tp.noise = nz; tp.noise_bstride = nzs; tp.noise_weight = L.p.noise_weight;
                tp.bias = L.p.act_bias; tp.next_style = up.next_style;
                tp.toeplitz = (const __nv_bfloat16 *)(ws + S->off_toeplitz);
                // the noise maps are the caller's tensors: their tensor map (tile = 8 x 16 pixels x the samples of a tile) is
                // encoded per launch (host only, ~1 us) and travels as a kernel parameter
                CUtensorMap tmN;
                memset(&tmN, 0, sizeof(tmN));
                if (nz) {
                    EncodeTiledFn enc = get_encode();
                    const int R = 2 * L.res_in;
                    SG2_REQUIRE((reinterpret_cast<uintptr_t>(nz) & 15) == 0 && (nzs == 0 || nzs == (long long)R * R), SG2_ERR_BAD_ARG,
                                "engine: noise maps must be 16-byte aligned and densely packed per sample (layer %d)", i);
                    const cuuint64_t nmaps = nzs ? (cuuint64_t)B : 1;
                    cuuint64_t dims[3] = {(cuuint64_t)R, (cuuint64_t)R, nmaps};
                    cuuint64_t strides[2] = {(cuuint64_t)R * 4, (cuuint64_t)R * R * 4};
                    cuuint32_t box[3] = {8, 16, (cuuint32_t)(nzs ? tp.nsamp : 1)};
                    cuuint32_t es[3] = {1, 1, 1};
                    CUresult crc = enc(&tmN, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)nz, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    SG2_REQUIRE(crc == CUDA_SUCCESS, SG2_ERR_CUDA, "engine: cuTensorMapEncodeTiled(noise) failed with %d", (int)crc);
                }
                rc = launch_upfir_tc(tp, tmN, L.tmT, L.tmO, S->sms, st);
            }
            if (rc) return rc;
            if ((rc = rec(S, st, "upfir"))) return rc;
        }
    }
    return SG2_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// One 3x3 stride-1 'same' convolution on the tensor-core kernel, outside the whole-network plan: the contraction of
// ModulatedConv2d (model.py:232-273 of the reference) once modulation / demodulation are factored out, and -- with the
// taps flipped and the channel roles swapped by the caller -- its input gradient.  NHWC bf16 in and out, fp32
// accumulation, out[b,y,x,co] = scale[b,co] * sum_{a,b',ci} x[b, y+a-1, x+b'-1, ci] * wp[a*3+b'][co][ci].
// Used by the differentiable path (stylegan2/tc_route.py) when bf16 operands are allowed.
extern "C" int sg2_conv3x3_tc_pack(void *wp, const float *weight, int cin, int cout, float scale, sg2_stream_t stream) {
    SG2_REQUIRE(wp && weight && cin >= 1 && cout >= 1, SG2_ERR_BAD_ARG, "conv3x3_tc_pack: bad argument");
    return launch_pack_conv_weight((__nv_bfloat16 *)wp, nullptr, weight, cin, cout, 9, scale, as_stream(stream));
}

namespace {

// shared body of the stand-alone entries: plan one layer, encode its descriptors, launch with the plain scaled store
int run_single_conv(void *out, const void *x, const void *wp, const float *scale, int64_t B64, int r, int cin, int cout,
                    int upsample, const int *taps, int ntaps, sg2_stream_t stream, const char *what) {
    SG2_REQUIRE(B64 >= 0 && B64 <= 65535 && r >= 4 && r <= 4096, SG2_ERR_BAD_ARG, "%s: bad shape (B %lld, r %d)", what,
                (long long)B64, r);
    SG2_REQUIRE(cin >= 32 && cin % 32 == 0 && cout >= 16 && cout % 16 == 0 && (!upsample || cout % 32 == 0) && cout <= 4096 &&
                    cin <= 4096,
                SG2_ERR_UNSUPPORTED, "%s: needs Cin %% 32 == 0 and Cout %% 16 == 0 (%% 32 for the transposed conv), got %d -> %d",
                what, cin, cout);
    if (taps) {
        SG2_REQUIRE(ntaps >= 1 && ntaps <= kGemmMaxTaps, SG2_ERR_BAD_ARG, "%s: 1..%d taps, got %d", what, kGemmMaxTaps, ntaps);
        for (int t = 0; t < ntaps; ++t)
            SG2_REQUIRE(taps[3 * t] >= -1 && taps[3 * t] <= 1 && taps[3 * t + 1] >= -1 && taps[3 * t + 1] <= 1 && taps[3 * t + 2] >= 0 &&
                            taps[3 * t + 2] < 9,
                        SG2_ERR_BAD_ARG, "%s: tap %d (dy %d, dx %d, weight %d) outside the 3x3 window", what, t, taps[3 * t],
                        taps[3 * t + 1], taps[3 * t + 2]);
    }
    if (B64 == 0) return SG2_OK;
    SG2_REQUIRE(out && x && wp && scale, SG2_ERR_BAD_ARG, "%s: null pointer", what);
    SG2_REQUIRE(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wp)) & 15) == 0,
                SG2_ERR_BAD_ARG, "%s: tensors must be 16-byte aligned", what);
    SG2_REQUIRE((long long)B64 * (r + 1) * (r + 1) * cout * (upsample ? 4 : 1) < (1LL << 34), SG2_ERR_UNSUPPORTED,
                "%s: output too large for 32-bit row offsets", what);
    int dev = 0, major = 0;
    SG2_CUDA_OK(cudaGetDevice(&dev));
    SG2_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    SG2_REQUIRE(major == 10, SG2_ERR_NO_DEVICE, "%s: needs an sm_100 device (tcgen05/TMEM), found compute capability %d.x", what, major);
    const int B = (int)B64;
    sg2_synth S;                                   // only max_batch and the SM count are read by the tile planner
    S.max_batch = B;
    S.sms = sm_count();
    Layer L;
    memset(&L.p, 0, sizeof(L.p));
    L.p.cin = cin; L.p.cout = cout; L.p.ksize = 3; L.p.upsample = upsample; L.p.resolution = upsample ? 2 * r : r;
    L.rgb = false;
    L.res_in = r; L.res_out = L.p.resolution;
    int rc = plan_gemm(&S, L, taps, ntaps);
    if (rc) return rc;
    finalize_tiles(L.gp, B);
    rc = encode_maps(&S, L, (const __nv_bfloat16 *)x, (const __nv_bfloat16 *)wp, B);
    if (rc) return rc;
    GemmParams g = L.gp;
    g.mode = 1;                                    // plain scaled store
    g.demod = scale;
    g.out = (__nv_bfloat16 *)out;
    if (upsample) {
        const long long plane = (long long)B * (r + 1) * (r + 1) * cout;
        for (int s = 0; s < g.nsub; ++s) g.sub[s].out_off = plane * s;
    }
    if (g.poly4) return launch_modconv_gemm2_poly4(g, L.tmA[0], L.tmB, S.sms, as_stream(stream));
    return L.two_sm ? launch_modconv_gemm2(g, L.tmA, L.tmB, S.sms, as_stream(stream))
                    : launch_modconv_gemm(g, L.tmA, L.tmB, S.sms, as_stream(stream));
}

}  // namespace

extern "C" int sg2_conv3x3_tc(void *out, const void *x, const void *wp, const float *scale, int64_t B, int r, int cin,
                              int cout, sg2_stream_t stream) {
    return run_single_conv(out, x, wp, scale, B, r, cin, cout, 0, nullptr, 0, stream, "conv3x3_tc");
}

// the same kernel with an arbitrary subset of the 3x3 window: out[b,y,x,co] = scale[b,co] * sum_t sum_ci
// x[b, y+dy_t, x+dx_t, ci] * wp[w_t][co][ci]; taps = ntaps rows of {dy, dx, w} (host array).  The polyphase
// components of the stride-2 convolution that is the input gradient of the transposed conv are of this form.
extern "C" int sg2_conv_taps_tc(void *out, const void *x, const void *wp, const float *scale, int64_t B, int r, int cin,
                                int cout, const int *taps, int ntaps, sg2_stream_t stream) {
    SG2_REQUIRE(taps, SG2_ERR_BAD_ARG, "conv_taps_tc: null tap list");
    return run_single_conv(out, x, wp, scale, B, r, cin, cout, 0, taps, ntaps, stream, "conv_taps_tc");
}

// stride-2 transposed 3x3 convolution (F.conv_transpose2d(x, w, stride=2, padding=0), model.py:246-252) as its four
// polyphase planes: planes[(py,px)][b][y][x][co] = scale[b,co] * T[b, 2y+py, 2x+px, co], allocated extent (r+1) x (r+1)
// per plane, valid extent (r+1-py) x (r+1-px) (the rest is not written).
extern "C" int sg2_conv_transpose3x3_tc(void *planes, const void *x, const void *wp, const float *scale, int64_t B, int r, int cin,
                               int cout, sg2_stream_t stream) {
    return run_single_conv(planes, x, wp, scale, B, r, cin, cout, 1, nullptr, 0, stream, "conv_transpose3x3_tc");
}
