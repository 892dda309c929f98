// upfirdn2d_pk.cu -- upfirdn2d for 2-BYTE storage (bf16 / fp16, NCHW planes, minor == 1, <= 4x4 taps), blur (up 1 / down 1)
// and down-sampling (up 1 / down 2), on packed fp32 math.
//
// Same definition as upfirdn2d_kernel (op/upfirdn2d_kernel.cu:52-137 of the reference: pad / crop, correlation with the
// flipped taps, decimation), fp32 accumulation, taps applied per output in the reference's order (tap rows ascending,
// columns inside), every multiply-add an IEEE fused multiply-add: results are those of upfirdn2d_stream.cu bit for bit.
//
// Why a second kernel.  In 2-byte storage the blur moves 4 bytes per output and spends 16 multiply-adds on it: at the
// HBM roofline (6.55 TB/s measured) that is 26e12 FMA/s, 85 % of what the FP32 pipe delivers (tools/probes/ffma_probe.cu:
// 30-35e12 FMA/s through FFMA and through fma.rn.f32x2 alike).  The row-streaming kernel issued 34 instructions per
// output (scalar FFMA + scalar 2-byte loads + fp32 staging) and stopped at 0.46 (blur) / 0.34 (down-2) of the roofline.  Here
//   * the multiply-adds are fma.rn.f32x2 over pairs of adjacent outputs (SASS FFMA2, the tap as a broadcast scalar operand
//     out of a uniform register): 8 issue slots per output instead of 16; outer-product taps (what make_kernel builds; decided
//     on the device) take a separable blur -- 4 issue slots per output, within one ulp of the storage type of the 16-tap order;
//   * rows travel HBM -> shared memory as aligned 16-byte cp.async chunks, whatever the row pitch (a 257-wide bf16 row is
//     514 bytes: no TMA tile, no vector load can start on it): the chunk grid is anchored at the 16-byte boundary below
//     the row segment, so the segment starts q = 0..7 elements into its first chunk; q is warp-uniform and advances by
//     in_w mod 8 per row.  A lane reads its window with two 16-byte shared loads and unpacks it with one conversion per
//     pair half; which of the eight unpack sequences applies is known statically where the geometry allows (template
//     parameter QS below), else an eight-way switch per row;
//   * copies, the zero-padding fix-up and the warp barrier happen once per BATCH of 4 or 8 rows, a ring of 2-3 batches per
//     warp; the prefetch cursor walks across work items, so a warp that finishes a band already has the first rows of its
//     next one in flight;
//   * zero padding: rows outside the plane are zero-filled by the copy itself (cp.async ignore-src), columns outside the
//     plane are zeroed in shared memory by the lanes of the group (at most 15 staged positions); over-reads never leave
//     the tensor (chunks that would are clamped / zero-filled -- the last 2..14 bytes of a tensor included).
//
// Work split: a group of WL = 2..32 lanes owns a strip of WL * TX output columns (TX = 8 blur, 4 down-2: 8 input elements
// = one chunk per lane and row either way) and walks down a band of rh rows; the 32 / WL groups of a warp take consecutive
// bands of the SAME plane and strip (rh a multiple of 8, so that every group sees the same q) or, with thousands of narrow
// planes, the same band of the planes P, P + 8, P + 16, ... (eight planes apart the alignment is the same again).
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace sg2 {

using namespace tc;

#ifndef PK_KO
#define PK_KO 0                  // knock-out bits for bottleneck analysis (variant builds only; results are WRONG when set):
#endif                           // 1 no FMAs, 2 no global stores, 4 no global -> shared copies, 8 unpack case 0 only
constexpr int PK_WARPS = 4;      // small CTAs: five of them fit the register file (<= 96 registers per thread) = 20 warps per SM
constexpr int PK_THREADS = 32 * PK_WARPS;
#ifndef PK_RB_N
#define PK_RB_N 3                // batches in the ring (modes with 4-row batches, up to 4 groups per warp)
#endif
#ifndef PK_BULK
#define PK_BULK 1               // rows by cp.async.bulk + mbarrier where that is faster (0: 16-byte cp.async per lane everywhere)
#endif
#ifndef PK_CTAS_N
#define PK_CTAS_N 5
#endif
constexpr int PK_CTAS = PK_CTAS_N;

struct UfdPkParams {
    int in_h, in_w, out_h, out_w;
    int pad_x0, pad_y0, kh, kw;
    int rh;                   // output rows per band (a multiple of 8 when the groups of a warp are bands of one plane)
    int n_strips, n_sbands;   // super-band = the 32 / WL bands one warp walks side by side
    long long planes, items;  // items = planes * n_sbands * n_strips (strip fastest)
    unsigned long long xb, xe;   // first byte of the input tensor (16-byte aligned) / one past its last byte
    int vec_store;
    int by_planes;            // NS > 1: the groups of a warp are planes 8 apart (else consecutive bands of one plane)
    int no_sep;               // never take the separable blur (A/B switch, tests of the 16-tap path)
};

// One conversion per pair half, written straight into the half of the 64-bit register pair the packed FMA reads.  An
// element that sits in two pairs is converted twice, by two DIFFERENT instructions (V = 0 / 1): ptxas merges identical
// conversions and then MOVES the value into its second pair, and both its moves and its shift-as-IMAD issue on the FMA
// pipe, the one this kernel saturates; PRMT / LOP3 run on the ALU pipe.  (fp16: cvt either way, HADD2.F32.)
template <typename T, int V> __device__ __forceinline__ float pk_lo(uint32_t w) {
    float f;
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        if constexpr (V == 0) asm volatile("prmt.b32 %0, %1, 0, 0x1044;" : "=f"(f) : "r"(w));    // bytes {0, 0, w.b0, w.b1} = w << 16
        else asm volatile("prmt.b32 %0, 0, %1, 0x5400;" : "=f"(f) : "r"(w));
    } else {
        asm volatile("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, l;\n\t}" : "=f"(f) : "r"(w));
    }
    return f;
}
template <typename T, int V> __device__ __forceinline__ float pk_hi(uint32_t w) {
    float f;
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        if constexpr (V == 0) asm volatile("and.b32 %0, %1, 0xffff0000;" : "=f"(f) : "r"(w));
        else asm volatile("prmt.b32 %0, %1, 0, 0x3244;" : "=f"(f) : "r"(w));                      // bytes {0, 0, w.b2, w.b3}
    } else {
        asm volatile("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, h;\n\t}" : "=f"(f) : "r"(w));
    }
    return f;
}

// element e (static after unrolling) of the lane's three chunks
template <typename T, int V>
__device__ __forceinline__ float pk_elem(const uint32_t (&c)[9], int e) {
    return (e & 1) ? pk_hi<T, V>(c[e >> 1]) : pk_lo<T, V>(c[e >> 1]);
}

// NP pairs (window[j], window[j + STEP]) for a window that starts Q elements into the chunks; EVEN / ODD starts
// blur: E[n] = (w[2n], w[2n+1]), O[n] = (w[2n+1], w[2n+2]); down-2: D[j] = (w[j], w[j+2])
template <typename T, int DOWN, int Q>
__device__ __forceinline__ void pk_unpack(const uint32_t (&c)[9], float2 (&P)[DOWN == 1 ? 10 : 8]) {
    if constexpr (DOWN == 1) {
#pragma unroll
        for (int n = 0; n < 5; ++n) {
            P[n] = make_float2(pk_elem<T, 0>(c, Q + 2 * n), pk_elem<T, 0>(c, Q + 2 * n + 1));
            P[5 + n] = make_float2(pk_elem<T, 1>(c, Q + 2 * n + 1), pk_elem<T, 1>(c, Q + 2 * n + 2));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) P[j] = make_float2(pk_elem<T, 0>(c, Q + j), pk_elem<T, 1>(c, Q + j + 2));
    }
}

__device__ __forceinline__ void pk_cp16(uint32_t dst, unsigned long long src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pk_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void pk_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t pk_lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void pk_sts16_zero(uint32_t addr) {
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"((unsigned short)0) : "memory");
}

// predicated forms (no branch around a single instruction)
__device__ __forceinline__ void pk_cp16_if(uint32_t dst, unsigned long long src, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}" ::"r"(dst), "l"(src), "r"((int)on) : "memory");
}
__device__ __forceinline__ void pk_sts128_zero_if(uint32_t addr, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %1, 0;\n\t@p st.shared.v4.b32 [%0], {%2, %2, %2, %2};\n\t}" ::"r"(addr), "r"((int)on), "r"(0) : "memory");
}
__device__ __forceinline__ void pk_sts16_zero_if(uint32_t addr, bool on) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %1, 0;\n\t@p st.shared.b16 [%0], %2;\n\t}" ::"r"(addr), "r"((int)on), "h"((unsigned short)0) : "memory");
}

// 16-byte copy, or 16 zero bytes when `zero` is set (cp.async's ignore-src operand: one LDGSTS.ZFILL, no branch)
__device__ __forceinline__ void pk_cp16_or_zero(uint32_t dst, unsigned long long src, bool zero) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\tcp.async.cg.shared.global [%0], [%1], 16, p;\n\t}" ::"r"(dst), "l"(src), "r"((int)zero) : "memory");
}

template <typename T>
__device__ __forceinline__ uint32_t pk_pack2(float2 v) {
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
        return *reinterpret_cast<const uint32_t *>(&h);
    } else {
        const __half2 h = __floats2half2_rn(v.x, v.y);
        return *reinterpret_cast<const uint32_t *>(&h);
    }
}
// one finished row of N outputs per lane: a predicated vector store when the whole row of the lane is valid and aligned
// (vec), element by element otherwise (n_ok of them: right edge of the plane, unaligned rows -- a rarely taken branch)
template <typename T, int N>
__device__ __forceinline__ void pk_store_row(T *dst, const float2 (&v)[N / 2], bool vec, int n_ok) {
    uint32_t pk[N / 2];
#pragma unroll
    for (int i = 0; i < N / 2; ++i) pk[i] = pk_pack2<T>(v[i]);
    if constexpr (N == 8)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p st.global.v4.b32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(dst), "r"(pk[0]), "r"(pk[1]),
                     "r"(pk[2]), "r"(pk[3]), "r"((int)vec) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p st.global.v2.b32 [%0], {%1, %2};\n\t}" ::"l"(dst), "r"(pk[0]), "r"(pk[1]), "r"((int)vec)
                     : "memory");
    if (n_ok > 0) {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            if (2 * i < n_ok) dst[2 * i] = Cvt<T>::from_f(v[i].x);
            if (2 * i + 1 < n_ok) dst[2 * i + 1] = Cvt<T>::from_f(v[i].y);
        }
    }
}

// QS: how the alignment q of a row is known.
//   -1      read per row, eight-way switch (any geometry); batches of 4 rows
//   0 .. 7  the same constant for every row of the launch (in_w a multiple of 8: the 2^k-wide down-sampling inputs)
//   8       q = position of the row in a batch of 8 (in_w = 1 mod 8: the 2^k + 1 wide blur inputs; the first batch of a
//           band starts q0 slots in, so that the row with alignment q sits at position q)
// In the static modes the unpack code is straight-line and the compiler interleaves it with the FMAs.
template <int QS, int NS> struct PkMode { static constexpr int BATCH = QS == 8 ? 8 : 4, RB = (QS == 8 || NS >= 16) ? 2 : (NS <= 4 ? PK_RB_N : 3); };   // (16 groups per warp: a shorter ring fits 48 KiB)

// WLOG2: lanes per group (log2, >= 1).  TX outputs per lane; every lane consumes 8 input elements (one chunk) per row.
template <typename T, int DOWN, int WLOG2, int QS>
__global__ void __launch_bounds__(PK_THREADS, PK_CTAS)
upfirdn2d_pk_kernel(T *__restrict__ out, const float *__restrict__ taps, const UfdPkParams p) {
    constexpr int TX = DOWN == 1 ? 8 : 4;
    constexpr int WU = DOWN == 1 ? 11 : 10;                        // window elements a lane uses
    constexpr int WL = 1 << WLOG2, NS = 32 >> WLOG2;
    constexpr int GB = (WL + 2) * 16;                              // bytes of one group's staged row: WL + 2 chunks
    constexpr int SLOT = NS * GB;
    constexpr int NP = DOWN == 1 ? 10 : 8;
    constexpr int R = DOWN == 1 ? 4 : 2;                           // output rows in flight
    constexpr int BATCH = PkMode<QS, NS>::BATCH, RB = PkMode<QS, NS>::RB;  // rows per barrier / prefetch round, rounds in the ring
    constexpr int RING = RB * BATCH * SLOT;
    extern __shared__ __align__(16) unsigned char pk_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> WLOG2, t = lane & (WL - 1);
    // rows by cp.async.bulk: only where it measured faster -- one group per warp and batches of 8 rows (the 2^k + 1 wide blur
    // planes from 129 outputs up: 0.78 -> 0.81); with several groups per warp (many small copies) or 4-row batches (an
    // mbarrier round trip per 4 rows) it measured slower (down-2 0.69 -> 0.62, 32^2 planes 0.66 -> 0.45)
    constexpr bool BULK = PK_BULK && QS == 8 && NS == 1;
    constexpr int WSTRIDE = RING + 512 + 64;                       // per warp: the ring, 16 bytes of scratch per lane, RB mbarriers
    uint32_t ring = smem_u32(pk_smem) + (uint32_t)warp * WSTRIDE + (uint32_t)g * GB;   // this group's part of slot 0
    const uint32_t scratch = smem_u32(pk_smem) + (uint32_t)warp * WSTRIDE + RING + 16u * lane;   // see pf_batch
    uint64_t *full = reinterpret_cast<uint64_t *>(pk_smem + (size_t)warp * WSTRIDE + RING + 512);   // full[b]: batch slot b has landed
    if (BULK && lane == 0) {
#pragma unroll
        for (int b = 0; b < RB; ++b) mbar_init(full + b, NS);      // one arrival per group leader and batch
        fence_barrier_init();
    }
    __syncwarp();
    asm volatile("" : "+r"(ring));                                 // keep it in a register (ptxas re-derives it from S2R otherwise)

    // flipped taps, zero padded to 4 x 4: kf[a][b] multiplies the sample a rows / b columns after the window's first
    float kf[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
            kf[a][b] = (a < p.kh && b < p.kw) ? __ldg(taps + (p.kh - 1 - a) * p.kw + (p.kw - 1 - b)) : 0.f;

    // Outer-product taps (what make_kernel builds, model.py:39-48 of the reference -- every caller of the op): the blur
    // runs as a horizontal 4-tap pass on the new row + 4 vertical accumulations, 32 packed FMAs per row instead of 64.
    // Decided here from the taps themselves: kf[a][b] == ca[a] * rb[b] to 4 ulp of the largest tap (ca = its column,
    // rb = its row / the tap); fp32 rounding differs from the 16-tap order by ~1e-7 relative, far below the storage type's.
    bool sep = false;
    float ca[4] = {0.f, 0.f, 0.f, 0.f}, rb[4] = {0.f, 0.f, 0.f, 0.f};
    if constexpr (DOWN == 1) {
        float pv = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (fabsf(kf[a][b]) > fabsf(pv)) {
                    pv = kf[a][b];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { ca[i] = kf[i][b]; rb[i] = kf[a][i]; }
                }
        if (pv != 0.f) {
            sep = true;
#pragma unroll
            for (int i = 0; i < 4; ++i) rb[i] = rb[i] / pv;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) sep = sep && fabsf(kf[a][b] - ca[a] * rb[b]) <= 4.8e-7f * fabsf(pv);
        }
        if (p.no_sep) sep = false;
    }

    auto body = [&](auto sep_tag) {
    constexpr bool SEP = decltype(sep_tag)::value;
    (void)SEP;
    const long long stride = (long long)gridDim.x * PK_WARPS;
    const long long item0 = (long long)blockIdx.x * PK_WARPS + warp;
    const unsigned long long pitch = 2ull * (unsigned long long)p.in_w;

    // item -> first output row of this lane's band, strip origin, first input row; need = input rows the bands of the
    // warp walk (warp-uniform: group 0 has the tallest band); the walk covers steps s = -u0 .. in whole batches, steps
    // outside [0, need) are skipped (u0 = alignment of the band's first row in mode 8, else 0)
    // The groups of a warp (NS > 1: planes narrower than 32 lanes) take either consecutive bands of one plane (rh a
    // multiple of 8) or, when there are planes enough (p.by_planes), the same band of the planes P, P + 8, P + 16, ...:
    // eight planes apart the alignment is the same whatever the plane size, and whole-plane bands have the smallest halo.
    struct Item { long long plane; int y0, xs0, iy0, cx0, need, u0; bool live; };
    auto decode = [&](long long item) {
        Item it;
        const int strip = (int)(item % p.n_strips);
        const long long rest = item / p.n_strips;
        const int sband = (int)(rest % p.n_sbands);
        it.plane = rest / p.n_sbands;
        int y00;                                                   // group 0
        if (p.by_planes) {
            it.plane = (it.plane >> 3) * (8 * NS) + (it.plane & 7) + 8 * g;
            y00 = sband * p.rh;
            it.y0 = y00;
        } else {
            y00 = sband * NS * p.rh;
            it.y0 = y00 + g * p.rh;
        }
        it.live = it.plane < p.planes;
        it.xs0 = strip * (WL * TX);
        it.cx0 = DOWN * it.xs0 - p.pad_x0;
        it.iy0 = DOWN * it.y0 - p.pad_y0;
        it.need = DOWN * (min(p.rh, p.out_h - y00) - 1) + 4;
        it.u0 = 0;
        if constexpr (QS == 8)   // alignment of the first row (the same for every group of the warp)
            it.u0 = (int)(((p.xb >> 1) + (unsigned long long)((it.plane * p.in_h + DOWN * y00 - p.pad_y0) * (long long)p.in_w + it.cx0)) & 7);
        return it;
    };

    // ---- prefetch cursor: whole batches, RB - 1 of them ahead of the consumer, across items ----
    long long pf_item = item0;
    bool pf_live = pf_item < p.items, pf_safe = false;
    int pf_s = 0, pf_need = 0, pf_iy = 0;                          // step / input row of the batch's first slot, steps of the item
    uint32_t pf_off = 0;                                           // ring offset of the slot the next row goes to
    unsigned long long pf_a = 0;                                   // byte address of (row pf_iy, column cx0) of this group, + 16 t
    auto pf_open = [&]() {
        const Item it = decode(pf_item);
        pf_s = -it.u0; pf_need = it.need; pf_iy = it.iy0 - it.u0;
        const unsigned long long a0 = p.xb + 2ull * (unsigned long long)((it.plane * p.in_h + it.iy0) * (long long)p.in_w + it.cx0);
        const unsigned long long first = a0 & ~15ull, last = (a0 + (unsigned long long)(it.need - 1) * pitch) & ~15ull;
        pf_safe = first >= p.xb && first <= last && last + GB <= p.xe;   // every chunk of every row of the band lies inside the tensor
        if (!it.live) { pf_safe = true; pf_iy = -0x40000000; }      // a group beyond the last plane stages zeros
        pf_a = a0 - it.u0 * pitch + 16ull * t;
    };
    if (pf_live) pf_open();
    int pf_b = 0;                                                  // batch slot (0 .. RB - 1) the next batch goes to
    auto pf_batch = [&]() {                                        // one batch of rows (never straddles items)
        if (pf_live) {
            const uint32_t dst0 = ring + pf_off + 16u * t;
            if constexpr (BULK) {
                // Rows as ONE cp.async.bulk each (the TMA engine, 16-byte aligned start, (WL + 2) * 16 bytes), issued by the
                // group's first lane and counted on the batch's mbarrier: nothing of the row passes the LSU / L1 data pipe on
                // its way in (tools/probes/rowload_probe.cu: +3 .. 9 % on the memory-side ceiling of this walk).  Rows outside the plane are zeroed by the lanes; a band that touches either end of the
                // tensor is fetched by every lane with clamped 16-byte cp.async instead and waited for on the spot.
                const bool unsafe = !pf_safe;
                const bool any_unsafe = __any_sync(0xffffffffu, unsafe);
                int rows_ok = 0;
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    const bool in_item = (unsigned)(pf_s + j) < (unsigned)pf_need;
                    const bool row_ok = in_item && (unsigned)(pf_iy + j) < (unsigned)p.in_h && !(PK_KO & 4);
                    const uint32_t dst = dst0 + j * SLOT;
                    if (!unsafe) {
                        pk_sts128_zero_if(dst, in_item && !row_ok);
                        pk_sts128_zero_if(dst + 16u * WL, in_item && !row_ok && t < 2);
                        rows_ok += row_ok ? 1 : 0;
                    } else {
                        const unsigned long long src = (pf_a + j * pitch) & ~15ull;
                        auto copy = [&](uint32_t d, unsigned long long ca) {
                            const long long rem = (long long)(p.xe - ca);
                            const int bytes = (!row_ok || ca < p.xb || rem <= 0) ? 0 : (rem < 16 ? (int)rem : 16);
                            pk_cp16(d, bytes ? ca : p.xb, bytes);
                        };
                        copy(dst, src);
                        if (t < 2) copy(dst + 16u * WL, src + 16ull * WL);
                    }
                }
                if (any_unsafe) {                                  // (warp-uniform) the clamped copies land before the arrival below
                    pk_commit();
                    pk_wait<0>();
                    __syncwarp();
                }
                if (t == 0) {
                    uint64_t *bar = full + pf_b;
                    if (unsafe || rows_ok == 0) {
                        mbar_arrive(bar);
                    } else {
                        fence_proxy_async();                       // the slots were read / patched through the generic proxy
                        mbar_arrive_expect_tx(bar, (uint32_t)(rows_ok * GB));
#pragma unroll
                        for (int j = 0; j < BATCH; ++j) {
                            const bool row_ok = (unsigned)(pf_s + j) < (unsigned)pf_need && (unsigned)(pf_iy + j) < (unsigned)p.in_h && !(PK_KO & 4);
                            if (row_ok)
                                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst0 + j * SLOT),
                                             "l"((pf_a + j * pitch) & ~15ull), "r"(GB), "r"(smem_u32(bar)) : "memory");
                        }
                    }
                }
            } else if (pf_safe) {
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    const bool in_item = (unsigned)(pf_s + j) < (unsigned)pf_need;
                    const bool row_ok = in_item && (unsigned)(pf_iy + j) < (unsigned)p.in_h;
                    const unsigned long long src = (pf_a + j * pitch) & ~15ull;   // chunk t of the row (16 t is folded into pf_a)
                    const uint32_t dst = dst0 + j * SLOT;
                    // rows outside the plane: zero-filled by the copy itself.  Chunks WL and WL + 1 of the row are copied by
                    // lanes 0 and 1; the other lanes zero-fill a 16-byte scratch slot of their own (no global traffic) so
                    // that the instruction needs neither a predicate nor the branch ptxas builds around one
                    const bool zero = !row_ok || (PK_KO & 4);
                    pk_cp16_or_zero(dst, zero ? p.xb : src, zero);
                    pk_cp16_or_zero(t < 2 ? dst + 16u * WL : scratch, zero || t >= 2 ? p.xb : src + 16ull * WL, zero || t >= 2);
                }
            } else {                                               // first / last band of the tensor: clamp every chunk
#pragma unroll 1
                for (int j = 0; j < BATCH; ++j) {
                    const bool row_ok = (unsigned)(pf_s + j) < (unsigned)pf_need && (unsigned)(pf_iy + j) < (unsigned)p.in_h;
                    const unsigned long long src = (pf_a + j * pitch) & ~15ull;
                    auto copy = [&](uint32_t dst, unsigned long long ca) {
                        const long long rem = (long long)(p.xe - ca);
                        const int bytes = (!row_ok || ca < p.xb || rem <= 0) ? 0 : (rem < 16 ? (int)rem : 16);
                        pk_cp16(dst, bytes ? ca : p.xb, bytes);
                    };
                    copy(dst0 + j * SLOT, src);
                    if (t < 2) copy(dst0 + j * SLOT + 16u * WL, src + 16ull * WL);
                }
            }
            pf_a += BATCH * pitch; pf_iy += BATCH; pf_s += BATCH;
            if (pf_s >= pf_need) {
                pf_item += stride;
                pf_live = pf_item < p.items;
                if (pf_live) pf_open();
            }
        }
        if constexpr (!BULK) pk_commit();
        pf_off += BATCH * SLOT;
        if (pf_off == RING) pf_off = 0;
        pf_b = pf_b + 1 == RB ? 0 : pf_b + 1;
    };
#pragma unroll 1
    for (int d = 0; d < RB - 1; ++d) pf_batch();

    uint32_t cs_off = 0;                                           // ring offset of the batch the consumer reads next
    int cs_b = 0;                                                  // ... its batch slot and the phase of that slot's mbarrier
    uint32_t cs_par = 0;
    for (long long item = item0; item < p.items; item += stride) {
        const Item it = decode(item);
        const int nrows = it.live ? max(0, min(p.rh, p.out_h - it.y0)) : 0;   // output rows of this lane's band
        const int x0 = it.xs0 + TX * t;
        const int n_out = nrows > 0 ? max(0, min(TX, p.out_w - x0)) : 0;
        const bool vec_ok = p.vec_store != 0 && n_out == TX;
        const int n_part = vec_ok ? 0 : n_out;                     // outputs of a row stored element by element
        // row the first store step points at (blur: s = 0 -> row -3, down-2: s = 1 -> row -1; never dereferenced)
        T *orow = out + (it.plane * p.out_h + it.y0 - (DOWN == 1 ? 3 : 1)) * (long long)p.out_w + x0;
        const int need = it.need;
        // alignment of the row segment inside its first chunk, in elements (the same for every group of the warp)
        int q = (int)((p.xb + 2ull * (unsigned long long)((it.plane * p.in_h + it.iy0) * (long long)p.in_w + it.cx0)) >> 1) & 7;
        // zero padding left / right of the plane: the staged positions that lie outside the plane AND inside the window of a
        // lane that has outputs -- at most 3 on the left, 12 on the right (host check) -- are zeroed in shared memory once
        // the batch has landed, position j of that list by lane j % WL of the group
        const int t_last = (min(WL * TX, p.out_w - it.xs0) - 1) / TX;              // last lane with outputs
        const int n_left = max(0, -it.cx0), pos_r = max(n_left, p.in_w - it.cx0);
        const int n_right = max(0, 8 * t_last + WU - pos_r);
        auto fixpos = [&](int j) { return j < n_left ? 2 * j : (j - n_left < n_right ? 2 * (pos_r + j - n_left) : -1); };
        const int fix0 = fixpos(t), fix1 = fixpos(t + WL);         // byte offsets of this lane's positions; 2 WL >= n_left + n_right (host check)
        const bool any_fix = n_left + n_right > 0;                 // warp-uniform

        float2 acc[R][TX / 2];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int i = 0; i < TX / 2; ++i) acc[r][i] = make_float2(0.f, 0.f);

        for (int sb = -it.u0; sb < need; sb += BATCH) {
            if constexpr (BULK) {
                mbar_wait(full + cs_b, cs_par);                    // the batch has landed (every group's rows)
                if (++cs_b == RB) { cs_b = 0; cs_par ^= 1u; }
            } else {
                pk_wait<RB - 2>();                                 // this lane's copies of the batch have landed
            }
            __syncwarp();                                          // ... and everybody's; every lane is done with the previous batch
            const uint32_t line0 = ring + cs_off;
            if (any_fix) {
#pragma unroll
                for (int j = 0; j < BATCH; ++j) {
                    const int qj = QS == 8 ? j : (QS >= 0 ? QS : ((q + j * p.in_w) & 7));
                    pk_sts16_zero_if(line0 + j * SLOT + (uint32_t)(2 * qj + fix0), fix0 >= 0);
                    pk_sts16_zero_if(line0 + j * SLOT + (uint32_t)(2 * qj + fix1), fix1 >= 0);
                }
                __syncwarp();
            }
            pf_batch();                                            // refill the slots of the previous batch
            cs_off += BATCH * SLOT;
            if (cs_off == RING) cs_off = 0;
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
                const int s = sb + u;
                if ((unsigned)s >= (unsigned)need) continue;       // slots before the band's first row (mode 8) / after its last
                const uint32_t line = line0 + u * SLOT + 16u * t;
                const int qs = QS == 8 ? u : (QS >= 0 ? QS : q);   // static in the modes 0 .. 8
                uint32_t c[9];
                {
                    const uint4 c0 = lds128(line), c1 = lds128(line + 16u);
                    c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w;
                    c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
                    c[8] = qs + WU > 16 ? pk_lds32(line + 32u) : 0u;
                }
                float2 P[NP];
                switch ((PK_KO & 8) ? 0 : qs) {
                    case 0: pk_unpack<T, DOWN, 0>(c, P); break;
                    case 1: pk_unpack<T, DOWN, 1>(c, P); break;
                    case 2: pk_unpack<T, DOWN, 2>(c, P); break;
                    case 3: pk_unpack<T, DOWN, 3>(c, P); break;
                    case 4: pk_unpack<T, DOWN, 4>(c, P); break;
                    case 5: pk_unpack<T, DOWN, 5>(c, P); break;
                    case 6: pk_unpack<T, DOWN, 6>(c, P); break;
                    default: pk_unpack<T, DOWN, 7>(c, P); break;
                }
                if constexpr (QS < 0) q = (q + p.in_w) & 7;

                if constexpr (DOWN == 1) {
                    // input row s is tap row a of output row s - a (ring slot (u - a) & 3); output pair ip = columns
                    // (2 ip, 2 ip + 1) takes the window pair that starts at 2 ip + b: E[ip + b/2] or O[ip + (b-1)/2]
                    if constexpr (SEP) {
                        float2 h[4];                               // the row filtered horizontally, per output pair
#pragma unroll
                        for (int ip = 0; ip < 4; ++ip) {
                            float2 v = make_float2(0.f, 0.f);
#pragma unroll
                            for (int b = 0; b < 4; ++b)
                                v = __ffma2_rn((b & 1) ? P[5 + ip + (b >> 1)] : P[ip + (b >> 1)], make_float2(rb[b], rb[b]), v);
                            h[ip] = v;
                        }
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            const int r = (u - a + 8) & 3;
#pragma unroll
                            for (int ip = 0; ip < 4; ++ip)
                                acc[r][ip] = __ffma2_rn(h[ip], make_float2(ca[a], ca[a]), a == 0 ? make_float2(0.f, 0.f) : acc[r][ip]);
                        }
                    } else {
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        const int r = (u - a + 8) & 3;
#pragma unroll
                        for (int ip = 0; ip < 4; ++ip) {
                            float2 v = a == 0 ? make_float2(0.f, 0.f) : acc[r][ip];
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                if ((PK_KO & 1) && (a | b)) continue;
                                v = __ffma2_rn((b & 1) ? P[5 + ip + (b >> 1)] : P[ip + (b >> 1)], make_float2(kf[a][b], kf[a][b]), v);
                            }
                            acc[r][ip] = v;
                        }
                    }
                    }
                    // finished: output row y0 + s - 3
                    const bool in_band = (unsigned)(s - 3) < (unsigned)nrows;
                    pk_store_row<T, 8>(orow, acc[(u + 1) & 3], in_band && vec_ok && !((PK_KO & 2) && p.out_h > 0), in_band && !((PK_KO & 2) && p.out_h > 0) ? n_part : 0);
                    orow += p.out_w;
                } else {
                    // input row s is tap row a = (s & 1), (s & 1) + 2 of output row (s - a) / 2; output pair ip = columns
                    // (2 ip, 2 ip + 1) takes (w[4 ip + b], w[4 ip + b + 2]) = D[4 ip + b]
                    const int e = u & 1;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int a = e + 2 * h;
                        const int r = (((u - a + 8) / 2)) & 1;     // (s - a) / 2 mod 2, sb a multiple of 4
#pragma unroll
                        for (int ip = 0; ip < 2; ++ip) {
                            float2 v = a == 0 ? make_float2(0.f, 0.f) : acc[r][ip];
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                if ((PK_KO & 1) && (a | b)) continue;
                                v = __ffma2_rn(P[4 * ip + b], make_float2(kf[a][b], kf[a][b]), v);
                            }
                            acc[r][ip] = v;
                        }
                    }
                    if (e == 1) {                                  // tap row 3 done: output row (s - 3) / 2 is finished
                        const bool in_band = (unsigned)((s - 3) >> 1) < (unsigned)nrows;   // s = 1: -1 -> out of range
                        pk_store_row<T, 4>(orow, acc[((u - 3 + 8) / 2) & 1], in_band && vec_ok && !((PK_KO & 2) && p.out_h > 0), in_band && !((PK_KO & 2) && p.out_h > 0) ? n_part : 0);
                        orow += p.out_w;
                    }
                }
            }
        }
    }
    pk_wait<0>();
    };
    if constexpr (DOWN == 1) {
        if (sep) body(std::true_type{}); else body(std::false_type{});
    } else {
        body(std::false_type{});
    }
}

// ---- host ------------------------------------------------------------------------------------------
static bool pk_debug() {
    static const char *e = getenv("SG2_US_DEBUG");
    return e && atoi(e);
}

template <typename T, int DOWN, int WLOG2, int QS>
static int pk_launch_t(void *out, const float *taps, const UfdPkParams &p, int grid, cudaStream_t st) {
    constexpr int WL = 1 << WLOG2, NS = 32 >> WLOG2;
    constexpr size_t smem = (size_t)PK_WARPS * (PkMode<QS, NS>::RB * PkMode<QS, NS>::BATCH * NS * (WL + 2) * 16 + 512 + 64);
    static_assert(smem <= 48 * 1024, "dynamic shared memory without the opt-in attribute");
    upfirdn2d_pk_kernel<T, DOWN, WLOG2, QS><<<grid, PK_THREADS, smem, st>>>((T *)out, taps, p);
    SG2_LAUNCH_CHECK();
    return SG2_OK;
}

template <typename T, int DOWN, int QS>
static int pk_launch_w(int wl, void *out, const float *taps, const UfdPkParams &p, int grid, cudaStream_t st) {
    switch (wl) {
        case 1:
            if constexpr (QS != 8) return pk_launch_t<T, DOWN, 1, QS>(out, taps, p, grid, st);
            else return 1;
        case 2:
            if constexpr (QS != 8) return pk_launch_t<T, DOWN, 2, QS>(out, taps, p, grid, st);
            else return 1;
        case 3: return pk_launch_t<T, DOWN, 3, QS>(out, taps, p, grid, st);
        case 4: return pk_launch_t<T, DOWN, 4, QS>(out, taps, p, grid, st);
        default: return pk_launch_t<T, DOWN, 5, QS>(out, taps, p, grid, st);
    }
}

// Returns SG2_OK when the launch was made, 1 when this path does not apply (the caller goes on to the row-streaming
// kernel), another status on errors.
template <typename T>
int launch_upfirdn2d_pk(void *out, const void *x, const float *taps, int64_t planes, int in_h, int in_w, int out_h, int out_w,
                        int kh, int kw, int up, int down, int pad_x0, int pad_y0, cudaStream_t st) {
    if constexpr (sizeof(T) != 2) {
        return 1;
    } else {
        if (up != 1 || (down != 1 && down != 2) || kh > 4 || kw > 4) return 1;
        if (reinterpret_cast<uintptr_t>(x) % 16 != 0) return 1;
        const int TX = down == 1 ? 8 : 4, WU = down == 1 ? 11 : 10;
        if (out_w <= TX || out_h < 8 || pad_x0 > 3) return 1;         // small planes: upfirdn2d_planes.cu / the streaming kernel
        int wl = 1;
        while ((TX << wl) < out_w && wl < 5) ++wl;
        const int WL = 1 << wl, NS = 32 >> wl;
        // 2 / 4 lanes per strip (planes up to 16 / 32 outputs wide; half that for down-2): only as 16 / 8 planes per warp, many planes
        if (wl <= 2 && planes < (int64_t)8 * NS * sm_count()) return 1;
        // the kernel zeroes at most 3 + 12 staged positions per row and group (see there): every strip must fit
        for (int xs0 = 0; xs0 < out_w; xs0 += WL * TX) {
            const int cx0 = down * xs0 - pad_x0, t_last = (std::min(WL * TX, out_w - xs0) - 1) / TX;
            const int n_left = std::max(0, -cx0), pos_r = std::max(n_left, in_w - cx0);
            if (n_left > 3 || 8 * t_last + WU - pos_r > 12 || n_left + std::max(0, 8 * t_last + WU - pos_r) > 2 * WL) return 1;
        }
        // how the kernel knows the alignment q of a row (see PkMode): a constant of the launch, the row's position in a
        // batch of 8, or read per row
        int qs = -1;
        if (in_w % 8 == 0 && pad_x0 >= 0) qs = (-pad_x0) & 7;        // 0, 7, 6, 5
        else if (down == 1 && in_w % 8 == 1 && wl > 2) qs = 8;      // (8 groups per warp: the ring of mode 8 would not fit)
        {
            const char *e = getenv("SG2_UPFIRDN_PK_QS");              // A/B switch (read per call): -1 = always the per-row switch
            if (e && atoi(e) < 0) qs = -1;
        }
        UfdPkParams p;
        p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w;
        p.pad_x0 = pad_x0; p.pad_y0 = pad_y0; p.kh = kh; p.kw = kw;
        p.planes = planes;
        p.n_strips = (out_w + TX * WL - 1) / (TX * WL);
        // band height: one super-band per plane if that still gives every resident warp ~3 items, otherwise shorter bands
        // (a multiple of 8 when a warp walks several bands side by side: its groups must see the same alignment)
        const int sms = sm_count();
        int per_warp = 3;
        if (const char *e = getenv("SG2_PK_ITEMS")) per_warp = std::max(1, atoi(e));   // experiment knob (read per call)
        const int64_t want_items = (int64_t)sms * PK_CTAS * PK_WARPS * per_warp;
        // narrow planes, many of them: the groups of a warp take planes 8 apart (see the kernel); the item count then
        // runs over blocks of 8 NS planes
        p.by_planes = NS > 1 && planes >= (int64_t)8 * NS * sms;
        {
            const char *e = getenv("SG2_UPFIRDN_PK_BYPLANES");        // A/B switch (read per call)
            if (e) p.by_planes = NS > 1 && (atoi(e) != 0 || wl <= 2);
        }
        const int64_t plane_slots = p.by_planes ? (planes + 8 * NS - 1) / (8 * NS) * 8 : planes;
        const int gb = p.by_planes ? 1 : NS;                         // bands of one plane walked side by side
        auto fit = [&](int v) { return gb > 1 ? (v + 7) & ~7 : v; };
        auto sbands = [&](int rh) { return (out_h + gb * rh - 1) / (gb * rh); };
        int rh = fit((out_h + gb - 1) / gb);
        const int rh_min = gb > 1 ? 8 : 16;
        while (rh > rh_min && plane_slots * p.n_strips * sbands(rh) < want_items) rh = std::max(rh_min, fit((rh + 1) / 2));
        p.rh = rh;
        p.n_sbands = sbands(rh);
        p.items = plane_slots * p.n_strips * p.n_sbands;
        p.xb = reinterpret_cast<uintptr_t>(x);
        p.xe = p.xb + 2ull * (unsigned long long)planes * in_h * in_w;
        {
            const char *e = getenv("SG2_UPFIRDN_PK_SEP");             // A/B switch (read per call): 0 = always the 16-tap blur
            p.no_sep = e && atoi(e) == 0;
        }
        p.vec_store = (out_w % TX == 0 && reinterpret_cast<uintptr_t>(out) % (TX * 2) == 0) ? 1 : 0;
        const int64_t want = (p.items + PK_WARPS - 1) / PK_WARPS;
        const int grid = (int)std::min<int64_t>(want, (int64_t)sms * PK_CTAS);
        if (grid <= 0) return SG2_OK;
        if (pk_debug())
            fprintf(stderr, "[sg2 upfirdn2d_pk] down %d, %lld planes %dx%d -> %dx%d, %d lanes per strip, bands of %d rows, %d super-bands, "
                            "%lld items, grid %d, q mode %d\n", down, (long long)planes, in_h, in_w, out_h, out_w, WL, p.rh, p.n_sbands,
                    p.items, grid, qs);
#define SG2_PK_Q(D, Q) case Q: return pk_launch_w<T, D, Q>(wl, out, taps, p, grid, st);
        if (down == 1) {
            switch (qs) { SG2_PK_Q(1, 0) SG2_PK_Q(1, 5) SG2_PK_Q(1, 6) SG2_PK_Q(1, 7) SG2_PK_Q(1, 8) default: return pk_launch_w<T, 1, -1>(wl, out, taps, p, grid, st); }
        } else {
            switch (qs) { SG2_PK_Q(2, 0) SG2_PK_Q(2, 5) SG2_PK_Q(2, 6) SG2_PK_Q(2, 7) default: return pk_launch_w<T, 2, -1>(wl, out, taps, p, grid, st); }
        }
#undef SG2_PK_Q
    }
}

template int launch_upfirdn2d_pk<float>(void *, const void *, const float *, int64_t, int, int, int, int, int, int, int, int, int, int,
                                        cudaStream_t);
template int launch_upfirdn2d_pk<__half>(void *, const void *, const float *, int64_t, int, int, int, int, int, int, int, int, int, int,
                                         cudaStream_t);
template int launch_upfirdn2d_pk<__nv_bfloat16>(void *, const void *, const float *, int64_t, int, int, int, int, int, int, int, int,
                                                int, int, cudaStream_t);

}  // namespace sg2
