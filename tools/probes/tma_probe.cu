// tma_probe.cu -- minimal flat-view TMA row load, one variant per process (an illegal instruction kills the context).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu -lcuda
//   ./tma_probe <es 2|4> <box_elems> <l2promo 0..3> <coord> <n_elems> <grid_constant 0|1> <dst_off>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ void body(const CUtensorMap *tm, float *out, int coord, int box_bytes, int dst_off) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 8192);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(box_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(s32(smem) + dst_off), "l"(reinterpret_cast<uint64_t>(tm)), "r"(s32(bar)), "r"(coord), "r"(0) : "memory");
    }
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(s32(bar)), "r"(0) : "memory");
    if (threadIdx.x < 8) out[threadIdx.x] = reinterpret_cast<float *>(smem + dst_off)[threadIdx.x];
}
__global__ void k_gc(const __grid_constant__ CUtensorMap tm, float *out, int coord, int box_bytes, int dst_off) {
    body(&tm, out, coord, box_bytes, dst_off);
}
__global__ void k_ptr(const CUtensorMap *tm, float *out, int coord, int box_bytes, int dst_off) {
    body(tm, out, coord, box_bytes, dst_off);
}

int main(int argc, char **argv) {
    const int es = atoi(argv[1]), box = atoi(argv[2]), promo = atoi(argv[3]), coord = atoi(argv[4]);
    const long long n = atoll(argv[5]);
    const int gc = atoi(argv[6]), dst_off = atoi(argv[7]);
    float *x, *out;
    cudaMalloc(&x, n * es + 4096);
    cudaMalloc(&out, 64);
    float *h = (float *)malloc(n * es);
    for (long long i = 0; i < n * es / 4; ++i) h[i] = (float)i;
    cudaMemcpy(x, h, n * es, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)n, 1};
    cuuint64_t strides[1] = {(cuuint64_t)((n * es + 15) / 16 * 16)};
    cuuint32_t bx[2] = {(cuuint32_t)box, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = cuTensorMapEncodeTiled(&tm, es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, dims,
                                         strides, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                         (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("es %d box %d promo %d coord %d n %lld gc %d dst_off %d: encode rc %d; ", es, box, promo, coord, n, gc, dst_off, (int)rc);
    if (rc) { printf("\n"); return 0; }
    if (gc) k_gc<<<1, 32, 16384>>>(tm, out, coord, box * es, dst_off);
    else {
        CUtensorMap *dtm;
        cudaMalloc(&dtm, sizeof(tm));
        cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
        k_ptr<<<1, 32, 16384>>>(dtm, out, coord, box * es, dst_off);
    }
    cudaError_t e = cudaDeviceSynchronize();
    float r[8] = {0};
    if (e == cudaSuccess) cudaMemcpy(r, out, 32, cudaMemcpyDeviceToHost);
    printf("run: %s; first words %g %g %g\n", cudaGetErrorString(e), r[0], r[1], r[2]);
    return 0;
}
