// common.cuh -- shared helpers for libsg2_b200 (sm_100a only).
#pragma once
// Knock-out branches for bottleneck analysis (SG2_GEMM_DBG / SG2_FIR_DBG: loads, MMAs, math or stores removed one at a
// time; results are WRONG when set) exist only in a variant build: `python build.py --tag ko -DSG2_KNOCKOUT=1`
// (tools/knockout.sh).  The shipped library compiles them out.
#ifndef SG2_KNOCKOUT
#define SG2_KNOCKOUT 0
#endif
#define SG2_DBG(p) (SG2_KNOCKOUT ? (p).dbg : 0)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/sg2_b200.h"

namespace sg2 {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define SG2_REQUIRE(cond, code, ...)      \
    do {                                  \
        if (!(cond)) {                    \
            sg2::set_error(__VA_ARGS__);  \
            return (code);                \
        }                                 \
    } while (0)

#define SG2_CUDA_OK(expr)                                                                   \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            sg2::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                           __LINE__);                                                       \
            return SG2_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

// after a kernel launch: catches bad launch configs without synchronising
#define SG2_LAUNCH_CHECK()                 \
    do {                                   \
        sg2::count_launch();               \
        SG2_CUDA_OK(cudaGetLastError());   \
    } while (0)

inline cudaStream_t as_stream(sg2_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // SMs of the current device (cached per device), 148 on B200
// driver entry point of cuTensorMapEncodeTiled (cast to the driver's signature by the caller), or null
void *tensor_map_encode_fn();

// ---- dtype dispatch -----------------------------------------------------------------------------
template <typename T> struct Cvt;
template <> struct Cvt<float> {
    __device__ __forceinline__ static float to_f(float v) { return v; }
    __device__ __forceinline__ static float from_f(float v) { return v; }
};
template <> struct Cvt<__half> {
    __device__ __forceinline__ static float to_f(__half v) { return __half2float(v); }
    __device__ __forceinline__ static __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct Cvt<__nv_bfloat16> {
    __device__ __forceinline__ static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ __forceinline__ static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

#define SG2_DISPATCH_DTYPE(dtype, ...)                                              \
    switch (dtype) {                                                                \
        case SG2_F32: { using T = float; __VA_ARGS__; break; }                      \
        case SG2_F16: { using T = __half; __VA_ARGS__; break; }                     \
        case SG2_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }             \
        default: sg2::set_error("unknown dtype %d", (int)(dtype)); return SG2_ERR_BAD_ARG; \
    }

// 16-byte vector of T
template <typename T> struct Vec16 { static constexpr int N = 16 / sizeof(T); T v[N]; };

template <typename T>
__device__ __forceinline__ Vec16<T> ld16(const T *p) {
    Vec16<T> r;
    *reinterpret_cast<uint4 *>(&r) = *reinterpret_cast<const uint4 *>(p);
    return r;
}
// streaming (read-once) 16-byte load: do not pollute L1
template <typename T>
__device__ __forceinline__ Vec16<T> ld16_stream(const T *p) {
    Vec16<T> r;
    uint4 u;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
    *reinterpret_cast<uint4 *>(&r) = u;
    return r;
}
template <typename T>
__device__ __forceinline__ void st16(T *p, const Vec16<T> &r) {
    *reinterpret_cast<uint4 *>(p) = *reinterpret_cast<const uint4 *>(&r);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace sg2
