// upfirdn2d_stream.cu -- the op-API upfirdn2d (NCHW planes, minor == 1, <= 4x4 taps, the three
// geometries the model uses: blur 1/1, skip up-sample 2/1, down-sample 1/2) as a ROW-STREAMING stencil.
//
// Replaces upfirdn2d_kernel (op/upfirdn2d_kernel.cu:52-137 of the reference; same definition: zero
// insertion, pad / crop, true convolution = correlation with the flipped taps, decimation) for the
// shapes that carry the bytes.  The first version of this op staged 2-D tiles through shared memory
// behind block-wide barriers and sat at 12-55 % of the HBM roofline (profiles/opbench_r01.jsonl);
// the bound is HBM, so the kernel is organised around keeping bytes in flight, not around tiles:
//
//   * a warp owns a vertical strip (4 output columns per lane) of one plane and walks DOWN it;
//   * input rows are fetched with plain coalesced loads US_PF rows ahead (register prefetch ring), any
//     row pitch / base alignment (the 257-wide blur input has neither a 16-byte pitch nor aligned rows;
//     TMA tiles need 16-byte aligned box starts -- tools/probes/tma_probe.cu -- so they cannot express
//     this access), converted to fp32 and staged in a per-warp line buffer from which every lane reads
//     its window with 16-byte loads;
//   * every input row is read from HBM once per band; it updates the <= 4 output rows it contributes
//     to (accumulators in registers, polyphase taps resolved at compile time: no multiplies by the
//     inserted zeros) and the finished output row leaves with one 8/16-byte store per lane;
//   * no block-wide barrier anywhere: warps run independently (__syncwarp only);
//   * round 2: rows as 16-byte (fp32 down-sampling) / 8-byte (2-byte up-sampling) vectors when pitch and base allow (template
//     parameter VD), the up-sampling accumulators as row pairs on packed fma.rn.f32x2.  Blur and down-sampling in 2-byte
//     storage run on upfirdn2d_pk.cu instead (this kernel is their fall-back).
//
// fp32 accumulation for every storage dtype; per output the taps are applied in the same order as
// the reference kernel (rows, then columns), so fp32 results match it to the last bits.
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace sg2 {

using namespace tc;

#ifndef SG2_US_PACKED
#define SG2_US_PACKED 0      // blur: packed f32x2 FMAs over output pairs (measured slower: the pair-building moves eat the saved issue slots)
#endif
constexpr int US_PF = 4;         // input rows in flight per warp (registers)
constexpr int US_WARPS = 8;
constexpr int US_THREADS = 32 * US_WARPS;

struct UfdStreamParams {
    int in_h, in_w, out_h, out_w;
    int pad_x0, pad_y0, kh, kw;
    int wl_log2;          // lanes per group (3..5): a group of 1 << wl_log2 lanes owns one strip
    long long planes;
    int rh;               // output rows per band (even)
    int n_strips, n_bands;
    long long items;      // planes * n_strips * n_bands
    int line_floats;      // shared-memory pitch of one staged row (multiple of 4)
    int vec_store;        // output rows are aligned for one vector store per lane
};

// outputs per lane (US_TX_WIDE for blur / up-sampling: halves the per-output share of loads, stores and loop
// control -- the kernel is issue-bound; the down-sampling window is already 8 inputs per lane at 4 outputs)
#ifndef SG2_US_TX_WIDE
#define SG2_US_TX_WIDE 8
#endif
template <int UP, int DOWN>
struct SGeo {
    static constexpr int TX = DOWN == 2 ? 4 : SG2_US_TX_WIDE;       // output columns per lane
    static constexpr int LS = TX * DOWN / UP;                       // input elements between the windows of adjacent lanes
    static constexpr int WU = UP == 2 ? TX / 2 + 2 : DOWN * (TX - 1) + 4;   // window elements used
    static constexpr int WR = (WU + 3) & ~3;                        // window elements read (16-byte loads)
    static constexpr int R = DOWN == 2 ? 2 : 4;                     // output rows in flight
    static_assert(TX == 4 || TX == 8, "4 or 8 outputs per lane");
};

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}

// WR elements of T starting at shared address addr (aligned to ALIGN bytes) -> fp32
template <typename T, int WR, int ALIGN>
__device__ __forceinline__ void load_window(uint32_t addr, float (&w)[WR]) {
    constexpr int NW = WR * (int)sizeof(T) / 4;                   // 32-bit words
    constexpr int V = ALIGN % 16 == 0 ? 4 : (ALIGN % 8 == 0 ? 2 : 1);   // words per load
    constexpr int FULL = NW / V, TAIL = NW % V;
    static_assert(TAIL == 0 || TAIL == 2, "window tail");
    uint32_t u[NW];
#pragma unroll
    for (int i = 0; i < FULL; ++i) {
        if constexpr (V == 4) {
            const uint4 q = lds128(addr + 16 * i);
            u[4 * i] = q.x; u[4 * i + 1] = q.y; u[4 * i + 2] = q.z; u[4 * i + 3] = q.w;
        } else if constexpr (V == 2) {
            const uint2 q = lds64(addr + 8 * i);
            u[2 * i] = q.x; u[2 * i + 1] = q.y;
        } else {
            u[i] = lds32(addr + 4 * i);
        }
    }
    if constexpr (TAIL == 2) {
        const uint2 q = lds64(addr + 4 * V * FULL);
        u[V * FULL] = q.x; u[V * FULL + 1] = q.y;
    }
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int j = 0; j < WR; ++j) w[j] = __uint_as_float(u[j]);
    } else if constexpr (std::is_same<T, __nv_bfloat16>::value) {
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            w[2 * j] = __uint_as_float(u[j] << 16);
            w[2 * j + 1] = __uint_as_float(u[j] & 0xffff0000u);
        }
    } else {
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&u[j]));
            w[2 * j] = f.x; w[2 * j + 1] = f.y;
        }
    }
}

template <typename T, int N>
__device__ __forceinline__ void store_row(T *dst, const float (&v)[N], int n_ok, bool vec) {
    if (vec && n_ok == N) {
        T pk[N];
#pragma unroll
        for (int i = 0; i < N; ++i) pk[i] = Cvt<T>::from_f(v[i]);
        constexpr int BYTES = N * (int)sizeof(T);
        if constexpr (BYTES == 32) {
            reinterpret_cast<uint4 *>(dst)[0] = reinterpret_cast<const uint4 *>(pk)[0];
            reinterpret_cast<uint4 *>(dst)[1] = reinterpret_cast<const uint4 *>(pk)[1];
        } else if constexpr (BYTES == 16) {
            *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(pk);
        } else {
            *reinterpret_cast<uint2 *>(dst) = *reinterpret_cast<const uint2 *>(pk);
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (i < n_ok) dst[i] = Cvt<T>::from_f(v[i]);
    }
}

// PHX / PHY: parity of (pad_x0, pad_y0) for UP == 2 (which polyphase pattern a 4-aligned output column /
// an even output row starts with); unused (0) otherwise.
// WLOG2: lanes per strip (log2) -- a template parameter so that the strided row loads take immediate offsets.
// VD: -1 = rows fetched element by element (any pitch / alignment); 0..3 = fp32 rows whose pitch and base are multiples of 16
// bytes (the 2^k-wide planes of the down- and up-sampling passes): the staged line starts VD elements before the strip's first
// column, on a 16-byte boundary, and rows travel as 16-byte loads / shared stores (the scalar form keeps the L1/LSU pipe 91 %
// busy on the fp32 down-sampling pass: 21 memory instructions per row and lane, 8 here); a vector lies entirely inside or
// entirely outside the plane, so the zero padding is a predicate on the load and nothing is ever read outside the tensor.
template <typename T, int UP, int DOWN, int PHX, int PHY, int WLOG2, int VD = -1>
__global__ void __launch_bounds__(US_THREADS)
upfirdn2d_stream_kernel(T *__restrict__ out, const T *__restrict__ x, const float *__restrict__ taps, const UfdStreamParams p) {
    using G = SGeo<UP, DOWN>;
    constexpr bool VEC = VD >= 0;
    // vector rows: 4 elements per load (16 bytes of fp32, 8 bytes of bf16 / fp16); the 2-byte up-sampling pass, whose rows
    // are short (4 + 1 elements per lane) and whose time went to waiting for them (ncu: long_scoreboard + wait), also keeps
    // 6 rows in flight instead of 4 -- a vector per row and lane costs fewer registers than 5 scalars did
    using VecT = typename std::conditional<sizeof(T) == 4, float4, uint2>::type;
    constexpr int PF = (VEC && UP == 2 && sizeof(T) == 2) ? 6 : US_PF;
    constexpr int D = VEC ? VD : 0;                                // window elements before the strip's first column
    constexpr int WRD = VEC ? (D + G::WU + 3) & ~3 : G::WR;        // window elements read
    constexpr int NI = VEC ? 1 : G::LS + 1;                        // loads per lane per row: ceil(line / lanes), lanes >= 4
    constexpr int NVF = G::LS / 4, NVX = (WRD - G::LS) / 4;        // vector rows: full vectors per lane, extra vectors (first lanes)
    extern __shared__ __align__(16) float us_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WL = 1 << WLOG2, NS = 32 >> WLOG2;
    const int g = lane >> WLOG2, t = lane & (WL - 1);              // group of the lane, lane inside the group
    constexpr int line_len = (WL - 1) * G::LS + WRD;               // staged elements of one row
    // positions t + WL*i, i < LS, always exist (WL*LS - 1 < line_len); the last one only for the first lanes
    static_assert(G::LS - 1 < G::WR, "line layout");
    const bool has_last = t < (VEC ? NVX : G::WR - G::LS);
    // per warp: two line buffers (row s is staged while the windows of row s - 1 may still be read) x NS groups
    float *wbase = us_smem + (size_t)warp * (2 * NS * p.line_floats);
    const uint32_t ring_s = smem_u32(wbase);

    // flipped taps, zero padded to 4 x 4 (upfirdn2d_kernel.cu:71-81): kf[a][b] multiplies the sample a rows / b
    // columns after the first one of the (up-sampled, padded) window
    float kf[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
            kf[a][b] = (a < p.kh && b < p.kw) ? __ldg(taps + (p.kh - 1 - a) * p.kw + (p.kw - 1 - b)) : 0.f;

#if SG2_US_PACKED && SG2_US_TX_WIDE == 4
    float2 kf2[4][4];                                              // (tap, tap): the multiplier pair of the packed FMAs
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) kf2[a][b] = make_float2(kf[a][b], kf[a][b]);
#endif
    const long long plane_in = (long long)p.in_h * p.in_w, plane_out = (long long)p.out_h * p.out_w;
    const long long total_groups = (long long)gridDim.x * US_WARPS * NS;
    const long long gid0 = ((long long)blockIdx.x * US_WARPS + warp) * NS + g;

    for (long long item0 = gid0 - g; item0 < p.items; item0 += total_groups) {   // warp-uniform trip count
        const long long item = item0 + g;
        const bool active = item < p.items;
        // item -> (plane, band, strip), strip fastest
        long long rest = active ? item : 0;
        const int strip = (int)(rest % p.n_strips); rest /= p.n_strips;
        const int band = (int)(rest % p.n_bands);
        const long long plane = rest / p.n_bands;
        const int xs0 = strip * (WL * G::TX), y0 = band * p.rh, y1 = min(p.out_h, y0 + p.rh);
        const int nrows = y1 - y0;
        // first input row / column of the strip and number of input rows to walk
        int iy_first, cx0, nsteps;
        if (UP == 1) {
            iy_first = DOWN * y0 - p.pad_y0;
            cx0 = DOWN * xs0 - p.pad_x0;
            nsteps = DOWN * (nrows - 1) + 4;
        } else {
            iy_first = (y0 - p.pad_y0 + PHY) / 2;                  // exact: y0 - pad_y0 + PHY is even
            cx0 = (xs0 - p.pad_x0 + PHX) / 2;
            nsteps = (nrows + 2 - PHY) / 2 + 1;                    // rows Y = y0 + PHY + 2s <= y1 + 2
        }
        cx0 -= D;                                                  // (vector rows: a multiple of 4 now)
        if (!active) nsteps = 0;
        int nsteps_max = nsteps;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nsteps_max = max(nsteps_max, __shfl_xor_sync(0xffffffffu, nsteps_max, o));

        // this lane stages line positions t + WL*i: which of them exist but lie OUTSIDE the plane (row invariant).
        // Rows are fetched without per-column predicates -- a column outside the plane reads a neighbouring row's
        // element, which a fix-up then overwrites with zero in the staged line -- except for the rows whose
        // over-read would leave the tensor (start of the first plane, end of the last plane).
        uint32_t outside = 0;
#pragma unroll
        for (int i = 0; i < (VEC ? 0 : NI); ++i) {
            const int pos = t + WL * i, col = cx0 + pos;
            if ((i < G::LS || has_last) && (col < 0 || col >= p.in_w)) outside |= 1u << i;
        }
        const long long item_first = plane * plane_in + cx0, total_in = p.planes * plane_in;   // element index of (row 0, position 0)
        const T *xp = x + plane * plane_in + cx0 + t;             // (row 0, position t)

        // ---- register prefetch ring: US_PF rows in flight ----
        T pre[VEC ? 1 : PF][NI];
        VecT prev[VEC ? PF : 1][VEC ? NVF + 1 : 1];
        // vector rows: this lane's vectors sit at line positions 4 (t + WL i); which of them lie inside the plane (row invariant)
        uint32_t vin = 0;
        if constexpr (VEC) {
#pragma unroll
            for (int i = 0; i <= NVF; ++i) {
                const int col = cx0 + 4 * (t + WL * i);
                if ((i < NVF || has_last) && col >= 0 && col + 4 <= p.in_w) vin |= 1u << i;
            }
        }
        auto fetch_vec = [&](int s, VecT (&r)[VEC ? NVF + 1 : 1]) {
            const int iy = iy_first + s;
            const bool row_ok = s < nsteps && iy >= 0 && iy < p.in_h;
            const VecT *rp = reinterpret_cast<const VecT *>(x + plane * plane_in + (long long)iy * p.in_w + cx0) + t;
#pragma unroll
            for (int i = 0; i <= NVF; ++i) {
                if (row_ok && ((vin >> i) & 1u)) r[i] = __ldg(rp + WL * i);
                else if constexpr (sizeof(T) == 4) r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                else r[i] = make_uint2(0u, 0u);
            }
        };
        auto vec_to_f4 = [](const VecT &v) {
            if constexpr (sizeof(T) == 4) {
                return v;
            } else if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
                                   __uint_as_float(v.y & 0xffff0000u));
            } else {
                const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2 *>(&v.y));
                return make_float4(a.x, a.y, b.x, b.y);
            }
        };
        auto fetch = [&](int s, T (&r)[NI]) {
            if constexpr (VEC) return;
            else {
            const int iy = iy_first + s;
            const bool row_ok = s < nsteps && iy >= 0 && iy < p.in_h;
            const T *rp = xp + (long long)iy * p.in_w;
            if (!row_ok) {
#pragma unroll
                for (int i = 0; i < NI; ++i) r[i] = Cvt<T>::from_f(0.f);
            } else if (item_first + (long long)iy * p.in_w < 0 || item_first + (long long)iy * p.in_w + line_len > total_in) {
                // the over-read of this row would leave the tensor (first rows of the first plane, last rows of the
                // last one; several rows when the plane is narrower than the staged line): per-column predicates
#pragma unroll
                for (int i = 0; i < NI; ++i)
                    r[i] = ((i < G::LS || has_last) && !((outside >> i) & 1u)) ? __ldg(rp + WL * i) : Cvt<T>::from_f(0.f);
            } else {
#pragma unroll
                for (int i = 0; i < G::LS; ++i) r[i] = __ldg(rp + WL * i);
                r[G::LS] = has_last ? __ldg(rp + WL * G::LS) : Cvt<T>::from_f(0.f);
            }
            }
        };
#pragma unroll
        for (int d = 0; d < PF; ++d) {
            if constexpr (VEC) fetch_vec(d, prev[d]); else fetch(d, pre[d]);
        }

        float acc[UP == 2 ? 1 : G::R][G::TX];                      // (up-sampling keeps its rows as pairs, below)
#pragma unroll
        for (int r = 0; r < (UP == 2 ? 1 : G::R); ++r)
#pragma unroll
            for (int i = 0; i < G::TX; ++i) acc[r][i] = 0.f;
        float2 acc2[2][G::TX];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < G::TX; ++i) acc2[r][i] = make_float2(0.f, 0.f);
        T *oplane = out + plane * plane_out;
        const int x0 = xs0 + G::TX * t;
        const int n_ok = active ? max(0, min(G::TX, p.out_w - x0)) : 0;

        for (int sb = 0; sb < nsteps_max; sb += PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int s = sb + u;
                if (s >= nsteps_max) break;
                // stage row s (fp32) in line buffer s & 1, refill its register slot with row s + US_PF
                const uint32_t line = ring_s + (uint32_t)(((u & 1) * NS + g) * p.line_floats) * 4u;
                float *lp = wbase + ((u & 1) * NS + g) * p.line_floats + t;
                if constexpr (VEC) {
                    float4 *lv = reinterpret_cast<float4 *>(wbase + ((u & 1) * NS + g) * p.line_floats) + t;
#pragma unroll
                    for (int i = 0; i < NVF; ++i) lv[WL * i] = vec_to_f4(prev[u][i]);
                    if (has_last) lv[WL * NVF] = vec_to_f4(prev[u][NVF]);
                    __syncwarp();
                    fetch_vec(s + PF, prev[u]);
                } else {
#pragma unroll
                    for (int i = 0; i < G::LS; ++i) lp[WL * i] = Cvt<T>::to_f(pre[u][i]);
                    if (has_last) lp[WL * G::LS] = Cvt<T>::to_f(pre[u][G::LS]);
                    if (outside) {                                 // zero padding: few lanes, border strips only
#pragma unroll
                        for (int i = 0; i < NI; ++i)
                            if ((outside >> i) & 1u) lp[WL * i] = 0.f;
                    }
                    __syncwarp();
                    fetch(s + PF, pre[u]);
                }
                float wfull[WRD];
                load_window<float, WRD, 16>(line + (uint32_t)(t * G::LS * 4), wfull);
                const float *w = wfull + D;                        // w[j]: the window element j columns after the strip's first

                if constexpr (UP == 1 && DOWN == 1) {
                    // input row s feeds tap row a of output row (s - a); ring slot (s - a) & 3 = (u - a) & 3
                    // packed f32x2 FMAs over output pairs (0,1) and (2,3): half the FMA issue slots (the kernel is issue
                    // bound: 16 FMAs per output are 73 % of the FP32 pipe at the bf16 HBM roofline); per output the taps
                    // are still applied in ascending order, each lane of the pair is an ordinary fused multiply-add
#if SG2_US_PACKED && SG2_US_TX_WIDE == 4
                    float2 wp[6];
#pragma unroll
                    for (int j = 0; j < 6; ++j) wp[j] = make_float2(w[j], w[j + 1]);
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int r = (u - a + 4) & 3;
                        float2 v01 = a == 0 ? make_float2(0.f, 0.f) : make_float2(acc[r][0], acc[r][1]);
                        float2 v23 = a == 0 ? make_float2(0.f, 0.f) : make_float2(acc[r][2], acc[r][3]);
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            v01 = __ffma2_rn(wp[b], kf2[a][b], v01);
                            v23 = __ffma2_rn(wp[b + 2], kf2[a][b], v23);
                        }
                        acc[r][0] = v01.x; acc[r][1] = v01.y; acc[r][2] = v23.x; acc[r][3] = v23.y;
                    }
#else
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int r = (u - a + 4) & 3;
#pragma unroll
                        for (int i = 0; i < G::TX; ++i) {
                            float v = a == 0 ? 0.f : acc[r][i];
#pragma unroll
                            for (int b = 0; b < 4; ++b) v = fmaf(w[i + b], kf[a][b], v);
                            acc[r][i] = v;
                        }
                    }
#endif
                    const int ol = s - 3;                          // finished: output row y0 + s - 3
                    if (ol >= 0 && ol < nrows && n_ok > 0)
                        store_row<T, G::TX>(oplane + (long long)(y0 + ol) * p.out_w + x0, acc[(u + 1) & 3], n_ok, p.vec_store != 0);
                } else if constexpr (UP == 1 && DOWN == 2) {
                    // input row s feeds tap rows a = (s & 1), (s & 1) + 2 of output rows (s - a) / 2
                    const int e = u & 1;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int a = e + 2 * h;
                        const int r = (((u - a) / 2) + 2) & 1;     // (s - a) / 2 mod 2 with s = sb + u, sb % 4 == 0
#pragma unroll
                        for (int i = 0; i < G::TX; ++i) {
                            float v = a == 0 ? 0.f : acc[r][i];
#pragma unroll
                            for (int b = 0; b < 4; ++b) v = fmaf(w[2 * i + b], kf[a][b], v);
                            acc[r][i] = v;
                        }
                    }
                    if (e == 1) {                                  // tap row 3 done: output row (s - 3) / 2 is finished
                        const int ol = (s - 3) / 2;
                        if (s >= 3 && ol < nrows && n_ok > 0)
                            store_row<T, G::TX>(oplane + (long long)(y0 + ol) * p.out_w + x0, acc[(((u - 3) / 2) + 2) & 1], n_ok,
                                      p.vec_store != 0);
                    }
                } else {
                    // UP == 2: input row s sits at up-sampled row Y = y0 + PHY + 2s and feeds tap row a of output row
                    // Y - a: the rows (Y - 1, Y) start here with tap rows (1, 0), the rows (Y - 3, Y - 2) finish with tap rows
                    // (3, 2) -- and the next input row sees the same two pairs one step older.  The accumulators are kept as
                    // those ROW PAIRS, so that one packed fma.rn.f32x2 (window value as the broadcast scalar, a pair of taps)
                    // does the work of two FFMAs: 32 instead of 64 per input row and lane.  Per output the products arrive
                    // in the order of the scalar form (window columns ascending, taps inside): bit-identical results.
                    // Window element c sits at up-sampled column x0 + PHX + 2c and feeds tap b of output i = 2c + PHX - b.
                    const int pn = u & 1;                          // pair that starts with this row; pn ^ 1 finishes
                    float2 vn[G::TX], vo[G::TX];
#pragma unroll
                    for (int i = 0; i < G::TX; ++i) { vn[i] = make_float2(0.f, 0.f); vo[i] = acc2[pn ^ 1][i]; }
#pragma unroll
                    for (int c = 0; c < G::WU; ++c)
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const int i = 2 * c + PHX - b;
                            if (i >= 0 && i < G::TX) {
                                vn[i] = __ffma2_rn(make_float2(w[c], w[c]), make_float2(kf[1][b], kf[0][b]), vn[i]);
                                vo[i] = __ffma2_rn(make_float2(w[c], w[c]), make_float2(kf[3][b], kf[2][b]), vo[i]);
                            }
                        }
#pragma unroll
                    for (int i = 0; i < G::TX; ++i) acc2[pn][i] = vn[i];
                    // finished: tap rows 3 and 2 -> output rows Y - 3 and Y - 2
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int oy = y0 + PHY + 2 * s - 3 + h;
                        float row[G::TX];
#pragma unroll
                        for (int i = 0; i < G::TX; ++i) row[i] = h == 0 ? vo[i].x : vo[i].y;
                        if (oy >= y0 && oy < y1 && n_ok > 0)
                            store_row<T, G::TX>(oplane + (long long)oy * p.out_w + x0, row, n_ok, p.vec_store != 0);
                    }
                }
            }
        }
    }
}

// ---- host ------------------------------------------------------------------------------------------
template <int VD>
static int launch_stream_vec(void *out, const void *x, const float *taps, const UfdStreamParams &p, int grid, size_t smem, cudaStream_t st) {
    float *o = (float *)out;
    const float *xi = (const float *)x;
    switch (p.wl_log2) {
        case 2: upfirdn2d_stream_kernel<float, 1, 2, 0, 0, 2, VD><<<grid, US_THREADS, smem, st>>>(o, xi, taps, p); break;
        case 3: upfirdn2d_stream_kernel<float, 1, 2, 0, 0, 3, VD><<<grid, US_THREADS, smem, st>>>(o, xi, taps, p); break;
        default: upfirdn2d_stream_kernel<float, 1, 2, 0, 0, 4, VD><<<grid, US_THREADS, smem, st>>>(o, xi, taps, p); break;
    }
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

// 2-byte up-sampling with 8-byte rows, the model's geometry (pad_x0 = 2: the staged line starts 3 elements early)
template <typename T, int PHY>
static int launch_stream_vec_up(void *out, const void *x, const float *taps, const UfdStreamParams &p, int grid, size_t smem, cudaStream_t st) {
    T *o = (T *)out;
    const T *xi = (const T *)x;
    switch (p.wl_log2) {
        case 2: upfirdn2d_stream_kernel<T, 2, 1, 0, PHY, 2, 3><<<grid, US_THREADS, smem, st>>>(o, xi, taps, p); break;
        case 3: upfirdn2d_stream_kernel<T, 2, 1, 0, PHY, 3, 3><<<grid, US_THREADS, smem, st>>>(o, xi, taps, p); break;
        case 4: upfirdn2d_stream_kernel<T, 2, 1, 0, PHY, 4, 3><<<grid, US_THREADS, smem, st>>>(o, xi, taps, p); break;
        default: upfirdn2d_stream_kernel<T, 2, 1, 0, PHY, 5, 3><<<grid, US_THREADS, smem, st>>>(o, xi, taps, p); break;
    }
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

template <typename T, int UP, int DOWN, int PHX, int PHY>
static int launch_stream_t(void *out, const void *x, const float *taps, const UfdStreamParams &p, int grid, size_t smem,
                           cudaStream_t st) {
    T *o = (T *)out;
    const T *xi = (const T *)x;
    switch (p.wl_log2) {
        case 2: upfirdn2d_stream_kernel<T, UP, DOWN, PHX, PHY, 2><<<grid, US_THREADS, smem, st>>>(o, xi, taps, p); break;
        case 3: upfirdn2d_stream_kernel<T, UP, DOWN, PHX, PHY, 3><<<grid, US_THREADS, smem, st>>>(o, xi, taps, p); break;
        case 4: upfirdn2d_stream_kernel<T, UP, DOWN, PHX, PHY, 4><<<grid, US_THREADS, smem, st>>>(o, xi, taps, p); break;
        default:
            if constexpr (DOWN == 2) { set_error("upfirdn2d_stream: bad strip width"); return SG2_ERR_BAD_ARG; }
            else upfirdn2d_stream_kernel<T, UP, DOWN, PHX, PHY, 5><<<grid, US_THREADS, smem, st>>>(o, xi, taps, p);
    }
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

static bool us_debug() {
    static const char *e = getenv("SG2_US_DEBUG");
    return e && atoi(e);
}

// Returns SG2_OK when the launch was made, 1 when this path does not apply (the caller falls back to the
// tiled kernels), another status on errors.
template <typename T>
int launch_upfirdn2d_stream(void *out, const void *x, const float *taps, int64_t planes, int in_h, int in_w, int out_h,
                            int out_w, int kh, int kw, int up, int down, int pad_x0, int pad_y0, cudaStream_t st) {
    if (!((up == 1 && (down == 1 || down == 2)) || (up == 2 && down == 1)) || kh > 4 || kw > 4) return 1;
    UfdStreamParams p;
    p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w;
    p.pad_x0 = pad_x0; p.pad_y0 = pad_y0; p.kh = kh; p.kw = kw;
    // lanes per strip: 4 output columns per lane (the 8-inputs-per-lane down-sampling window stays at 16 lanes: 9
    // prefetch registers per row)
    const int TX = down == 2 ? 4 : SG2_US_TX_WIDE;                   // SGeo<UP, DOWN>::TX
    int wl = 2;                                                      // >= 4 lanes per plane, up to 8 planes per warp
    while ((TX << wl) < out_w && wl < (down == 2 ? 4 : 5)) ++wl;
    p.wl_log2 = wl;
    p.planes = planes;
    const int WL = 1 << wl, NS = 32 >> wl;
    const int LS = TX * down / up, WU = up == 2 ? TX / 2 + 2 : down * (TX - 1) + 4, WR = (WU + 3) & ~3;
    // fp32 down-sampling of planes whose rows start on 16-byte boundaries (the 2^k-wide planes): rows as 16-byte vectors,
    // the staged line starts vd elements before the strip's first column (see the kernel)
    int vd = -1;
    if (std::is_same<T, float>::value && up == 1 && down == 2 && in_w % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0) {
        static const char *env_vec = getenv("SG2_UPFIRDN_VEC");         // A/B switch: 0 = element-wise rows
        if (!env_vec || atoi(env_vec) != 0) vd = ((-pad_x0) % 4 + 4) % 4;
    }
    if (sizeof(T) == 2 && up == 2 && down == 1 && pad_x0 == 2 && in_w % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0) {
        static const char *env_vec = getenv("SG2_UPFIRDN_VEC");
        if (!env_vec || atoi(env_vec) != 0) vd = 3;                     // cx0 = (0 - 2) / 2 = -1: 3 elements after a multiple of 4
    }
    p.line_floats = ((WL - 1) * LS + (vd >= 0 ? ((vd + WU + 3) & ~3) : WR) + 3) & ~3;
    p.n_strips = (out_w + TX * WL - 1) / (TX * WL);
    // band height: tall enough to amortise the vertical halo, short enough for >= 4 items per resident group
    const int sms = sm_count();
    const int64_t groups = (int64_t)sms * 4 * US_WARPS * NS;
    int rh = 128;                                                   // tallest band that still gives every resident group two items
    while (rh > 16 && planes * p.n_strips * ((out_h + rh - 1) / rh) < 2 * groups) rh /= 2;
    p.rh = rh;
    p.n_bands = (out_h + rh - 1) / rh;
    p.items = planes * p.n_strips * p.n_bands;
    const int es = (int)sizeof(T);
    p.vec_store = (out_w % TX == 0 && reinterpret_cast<uintptr_t>(out) % std::min(16, TX * es) == 0) ? 1 : 0;
    const size_t smem = (size_t)US_WARPS * 2 * NS * p.line_floats * sizeof(float);
    const int64_t want = (p.items + (int64_t)US_WARPS * NS - 1) / ((int64_t)US_WARPS * NS);
    const int grid = (int)std::min<int64_t>(want, (int64_t)sms * 4);
    if (grid <= 0) return SG2_OK;
    if (us_debug())
        fprintf(stderr, "[sg2 upfirdn2d_stream] up %d down %d, %lld planes %dx%d -> %dx%d, %d lanes per strip, bands of %d rows, "
                        "%lld items, grid %d\n", up, down, (long long)planes, in_h, in_w, out_h, out_w, WL, p.rh, p.items, grid);
    const int phx = up == 2 ? (pad_x0 & 1) : 0, phy = up == 2 ? (pad_y0 & 1) : 0;
    if constexpr (std::is_same<T, float>::value) {
        switch (vd) {
            case 0: return launch_stream_vec<0>(out, x, taps, p, grid, smem, st);
            case 1: return launch_stream_vec<1>(out, x, taps, p, grid, smem, st);
            case 2: return launch_stream_vec<2>(out, x, taps, p, grid, smem, st);
            case 3: return launch_stream_vec<3>(out, x, taps, p, grid, smem, st);
            default: break;
        }
    }
    if constexpr (sizeof(T) == 2) {
        if (vd == 3) return phy == 0 ? launch_stream_vec_up<T, 0>(out, x, taps, p, grid, smem, st) : launch_stream_vec_up<T, 1>(out, x, taps, p, grid, smem, st);
    }
    if (up == 1 && down == 1) return launch_stream_t<T, 1, 1, 0, 0>(out, x, taps, p, grid, smem, st);
    if (up == 1 && down == 2) return launch_stream_t<T, 1, 2, 0, 0>(out, x, taps, p, grid, smem, st);
    if (phx == 0 && phy == 0) return launch_stream_t<T, 2, 1, 0, 0>(out, x, taps, p, grid, smem, st);
    if (phx == 1 && phy == 0) return launch_stream_t<T, 2, 1, 1, 0>(out, x, taps, p, grid, smem, st);
    if (phx == 0 && phy == 1) return launch_stream_t<T, 2, 1, 0, 1>(out, x, taps, p, grid, smem, st);
    return launch_stream_t<T, 2, 1, 1, 1>(out, x, taps, p, grid, smem, st);
}

template int launch_upfirdn2d_stream<float>(void *, const void *, const float *, int64_t, int, int, int, int, int, int, int, int,
                                            int, int, cudaStream_t);
template int launch_upfirdn2d_stream<__half>(void *, const void *, const float *, int64_t, int, int, int, int, int, int, int, int,
                                             int, int, cudaStream_t);
template int launch_upfirdn2d_stream<__nv_bfloat16>(void *, const void *, const float *, int64_t, int, int, int, int, int, int,
                                                    int, int, int, int, cudaStream_t);

}  // namespace sg2
