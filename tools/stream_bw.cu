// stream_bw.cu -- micro-benchmark: how fast can one persistent CTA per SM pull a contiguous byte range out of HBM?
//   mode 0: cp.async.bulk ring (like b200_gemv_stream.cu) with consumers that only wait + release
//   mode 1: plain 128-bit LDG with N loads in flight per thread
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_bw stream_bw.cu ; run: ./stream_bw
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(288) bulk_kernel(const uint8_t *src, size_t total, int stage_bytes, int stages, unsigned *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)stages * stage_bytes);
    uint64_t *empty = full + 16;
    const size_t per = total / gridDim.x / stage_bytes * stage_bytes;
    const uint8_t *base = src + (size_t)blockIdx.x * per;
    const int iters = (int)(per / stage_bytes);
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 8) {
        if (lane == 0)
            for (int it = 0; it < iters; it++) {
                const int s = it % stages;
                mbar_wait(&empty[s], ((it / stages) & 1) ^ 1);
                mbar_expect_tx(&full[s], stage_bytes);
                bulk_g2s(smem + (size_t)s * stage_bytes, base + (size_t)it * stage_bytes, stage_bytes, &full[s]);
            }
        return;
    }
    unsigned acc = 0;
    for (int it = 0; it < iters; it++) {
        const int s = it % stages;
        mbar_wait(&full[s], (it / stages) & 1);
        acc += *reinterpret_cast<const unsigned *>(smem + (size_t)s * stage_bytes + threadIdx.x * 16);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (acc == 0x12345678) sink[0] = acc;
}

template <int U>
__global__ void __launch_bounds__(256) ldg_kernel(const uint4 *src, size_t total_vec, unsigned *sink) {
    const size_t per = total_vec / gridDim.x;
    const uint4 *base = src + (size_t)blockIdx.x * per;
    unsigned acc = 0;
    for (size_t i = threadIdx.x; i + (U - 1) * 256 < per; i += 256 * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(base + i + u * 256));
#pragma unroll
        for (int u = 0; u < U; u++) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678) sink[0] = acc;
}

int main() {
    const size_t bufsz = (size_t)1 << 30;  // 1 GiB, rotate to defeat L2
    uint8_t *buf;
    unsigned *sink;
    cudaMalloc(&buf, bufsz);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bufsz);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const size_t sizes[] = {9437184, 37748736, 132120576, 536870912};
    for (size_t total : sizes) {
        printf("== %.1f MB per launch\n", total / 1e6);
        const int reps = total > 100e6 ? 4 : 20;
        for (int stage_kb : {4, 8, 16, 32}) {
            for (int stages : {2, 4, 6, 12}) {
                if (stage_kb * stages > 192) continue;
                const int sb = stage_kb * 1024;
                const size_t smem = (size_t)sb * stages + 256;
                float best = 1e9;
                for (int trial = 0; trial < 3; trial++) {
                    cudaEventRecord(e0);
                    for (int r = 0; r < reps; r++) bulk_kernel<<<148, 288, smem>>>(buf + ((size_t)(r + trial * reps) * total) % (bufsz - total), total, sb, stages, sink);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (ms < best) best = ms;
                }
                printf("bulk stage %2d KB x %2d: %7.2f us/launch  %7.1f GB/s\n", stage_kb, stages, best * 1e3 / reps, total / (best / reps * 1e-3) / 1e9);
            }
        }
        for (int grid_mult : {1, 2, 4, 8}) {
            float best = 1e9;
            for (int trial = 0; trial < 3; trial++) {
                cudaEventRecord(e0);
                for (int r = 0; r < reps; r++) ldg_kernel<8><<<148 * grid_mult, 256>>>((const uint4 *)(buf + ((size_t)(r + trial * reps) * total) % (bufsz - total)), total / 16, sink);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("ldg U=8 grid %4d: %7.2f us/launch  %7.1f GB/s\n", 148 * grid_mult, best * 1e3 / reps, total / (best / reps * 1e-3) / 1e9);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
