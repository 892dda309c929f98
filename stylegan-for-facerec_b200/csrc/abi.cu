// abi.cu -- library-wide plumbing of the C ABI: version, thread-local error text, launch counter,
// per-device attribute cache.  No kernels here.
#include <stdarg.h>

#include <mutex>

#include "common.cuh"

namespace sg2 {

static thread_local char t_error[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static std::atomic<int> cache[64];   // immutable once filled; 0 = unknown
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cache[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

void *tensor_map_encode_fn() {
    static std::atomic<void *> fn{nullptr};
    void *p = fn.load(std::memory_order_acquire);
    if (!p) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        fn.store(p, std::memory_order_release);
    }
    return p;
}

}  // namespace sg2

extern "C" int sg2_abi_version(void) { return SG2_ABI_VERSION; }
extern "C" const char *sg2_last_error(void) { return sg2::t_error; }
extern "C" int64_t sg2_launch_count(void) { return sg2::g_launches.load(std::memory_order_relaxed); }
extern "C" void sg2_note_launches(int64_t n) { sg2::g_launches.fetch_add(n, std::memory_order_relaxed); }
