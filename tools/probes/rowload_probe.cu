// rowload_probe.cu -- how fast can a warp-per-plane row walk pull 2-byte rows (514-byte pitch: 257 bf16) from HBM into a
// shared-memory ring, with nothing else to do?  The ceiling of upfirdn2d_pk.cu's loader, by mechanism:
//   mode 0: 16-byte cp.async (LDGSTS.128) per lane, one commit group per batch of 8 rows      (what the kernel does)
//   mode 1: one cp.async.bulk (UBLKCP) per row issued by lane 0, one mbarrier per batch
//   mode 2: plain 16-byte ld.global per lane into registers (no shared memory), summed
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a rowload_probe.cu -o rowload_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int WARPS = 4, BATCH = 8, RB = 2, ROWB = 34 * 16;   // 34 chunks cover a 257-element row segment at any alignment

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE, int WR = 0>      // WR: the store pattern beside the loads -- 0 none, 1 down-2 (8 bytes per lane every second row), 2 blur (16 bytes per lane and row)
__global__ void __launch_bounds__(32 * WARPS, 5) k_rows(const uint16_t *x, long long planes, int in_h, int in_w, unsigned *sink, uint16_t *y = nullptr) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *ring = smem + warp * (RB * BATCH * ROWB + 64);
    uint64_t *bar = reinterpret_cast<uint64_t *>(ring + RB * BATCH * ROWB);
    if (MODE == 1 && lane == 0) {
        for (int i = 0; i < RB; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const unsigned long long xb = (unsigned long long)x, xe = xb + 2ull * planes * in_h * in_w;
    unsigned acc = 0;
    const long long stride = (long long)gridDim.x * WARPS;
    const int nb = in_h / BATCH;                                  // whole batches only (the probe skips the last row)
    int slot = 0, phase = 0;
    for (long long plane = (long long)blockIdx.x * WARPS + warp; plane < planes; plane += stride) {
        const unsigned long long a0 = xb + 2ull * plane * in_h * in_w;
        auto issue = [&](int b, int sl) {
            for (int j = 0; j < BATCH; ++j) {
                unsigned long long src = (a0 + 2ull * (unsigned long long)(b * BATCH + j) * in_w) & ~15ull;
                if (src + ROWB > xe) src = (xe - ROWB) & ~15ull;
                const uint32_t dst = s32(ring + (sl * BATCH + j) * ROWB);
                if (MODE == 0) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * lane), "l"(src + 16ull * lane) : "memory");
                    if (lane < 2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (32 + lane)), "l"(src + 16ull * (32 + lane)) : "memory");
                } else if (lane == 0) {
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                                 "r"(ROWB), "r"(s32(bar + sl)) : "memory");
                }
            }
            if (MODE == 0) asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (MODE == 2) {
            for (int r = 0; r < in_h; ++r) {
                const unsigned long long src = (a0 + 2ull * (unsigned long long)r * in_w) & ~15ull;
                uint4 v;
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + 16ull * lane));
                acc += v.x ^ v.y ^ v.z ^ v.w;
            }
            continue;
        }
        if (MODE == 1 && lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar + slot)), "r"(BATCH * ROWB) : "memory");
        issue(0, slot);
        for (int b = 0; b < nb; ++b) {
            const int nslot = slot ^ 1;
            if (b + 1 < nb) {
                if (MODE == 1 && lane == 0)
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar + nslot)), "r"(BATCH * ROWB) : "memory");
                issue(b + 1, nslot);
            }
            if (MODE == 0) {
                if (b + 1 < nb) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
            } else {
                uint32_t done = 0;
                while (!done)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done)
                                 : "r"(s32(bar + slot)), "r"((phase >> slot) & 1) : "memory");
                phase ^= 1 << slot;
            }
            acc += *reinterpret_cast<const unsigned *>(ring + (slot * BATCH + (b & 7)) * ROWB + 16 * lane);   // touch the batch
            if (WR == 1) {          // output plane 128 x 128: rows 4 b .. 4 b + 3 of it from this batch of 8 input rows
                for (int j = 0; j < 4; ++j) {
                    const int oy = 4 * b + j;
                    if (oy < 128) *reinterpret_cast<uint2 *>(y + (plane * 128 + oy) * 128 + 4 * lane) = make_uint2(acc, acc + j);
                }
            } else if (WR == 2) {   // output plane 256 x 256: rows 8 b .. 8 b + 7
                for (int j = 0; j < 8; ++j) {
                    const int oy = 8 * b + j;
                    if (oy < 256) *reinterpret_cast<uint4 *>(y + (plane * 256 + oy) * 256 + 8 * lane) = make_uint4(acc, acc + j, acc, acc);
                }
            }
            __syncwarp();
            slot = nslot;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int MODE, int WR = 0>
static void run(const char *name, const uint16_t *x, long long planes, int h, int w, unsigned *sink, uint16_t *y = nullptr) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t smem = WARPS * (RB * BATCH * ROWB + 64);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) k_rows<MODE, WR><<<sms * 5, 32 * WARPS, smem>>>(x, planes, h, w, sink, y);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k_rows<MODE, WR><<<sms * 5, 32 * WARPS, smem>>>(x, planes, h, w, sink, y);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (2.0 * planes * h * w + (WR == 1 ? 2.0 * planes * 128 * 128 : WR == 2 ? 2.0 * planes * 256 * 256 : 0.0)) * 5;
    printf("%-28s %.3f ms per pass, %.2f TB/s (%s)\n", name, ms / 5, bytes / (ms * 1e-3) * 1e-12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const long long planes = 8192;
    const int h = 257, w = 257;
    uint16_t *x;
    unsigned *sink;
    cudaMalloc(&x, 2ull * planes * h * w + 4096);
    cudaMalloc(&sink, 4);
    cudaMemset(x, 1, 2ull * planes * h * w + 4096);
    run<0>("cp.async 16 B per lane", x, planes, h, w, sink);
    run<1>("cp.async.bulk per row", x, planes, h, w, sink);
    run<2>("ld.global 16 B per lane", x, planes, h, w, sink);
    run<0>("cp.async 16 B per lane", x, planes, h, w, sink);
    run<1>("cp.async.bulk per row", x, planes, h, w, sink);
    // the same walk with the store pattern of the two kernels beside it (no math): the memory-side ceiling of each read / write mix
    uint16_t *y;
    cudaMalloc(&y, 2ull * planes * 256 * 256);
    run<0, 1>("cp.async + down-2 stores", x, planes, h, w, sink, y);
    run<0, 2>("cp.async + blur stores", x, planes, h, w, sink, y);
    run<1, 1>("bulk + down-2 stores", x, planes, h, w, sink, y);
    run<1, 2>("bulk + blur stores", x, planes, h, w, sink, y);
    return 0;
}
