// tools/mma_peak.cu -- what the 5th-generation tensor cores of THIS GPU deliver for the two instruction kinds the prefill
// GEMM can use: tcgen05.mma kind::i8 (int8 x int8 -> int32, K = 32 per instruction) and kind::f16 (fp16 x fp16 -> fp32,
// K = 16), issued back to back from one thread per CTA (cta_group::1, M = 128, N = 256) or per CTA pair (cta_group::2,
// M = 256, N = 256) on operands that already sit in shared memory (SWIZZLE_128B, K-major).  No loads, no epilogue: the
// roofline denominators for bench.py's c2_* fractions (BASELINE.md asked for a measured int8 peak instead of 2 x bf16).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/mma_peak tools/mma_peak.cu && tools/_build/mma_peak
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// KIND 0 = i8, 1 = f16; CG = cta_group
template <int KIND, int CG>
__global__ void __launch_bounds__(128, 1) peak_kernel(int iters, unsigned long long *cycles) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 4 * 49152);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = threadIdx.x; i < 4 * 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        __syncthreads();
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // instruction descriptor: i8: D = S32 (2 << 4), A = B = S8 (1 << 7, 1 << 10); f16: D = F32 (1 << 4), A = B = F16 (0); N = 256, M = 128 * CG
    constexpr uint32_t M = 128 * CG, N = 256;
    constexpr uint32_t idesc = (KIND == 0 ? ((2u << 4) | (1u << 7) | (1u << 10)) : (1u << 4)) | ((N >> 3) << 17) | ((M >> 4) << 24);
    if (warp == 1 && lane == 0 && rank == 0) {
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
            const uint32_t sa = smem_u32(smem + (it & 3) * 49152);
            const uint64_t da = make_desc_sw128(sa), db = make_desc_sw128(sa + 16384);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t acc = (it | j) != 0 ? 1u : 0u;
                const uint32_t td = tmem_base + (uint32_t)((it & 1) * 256);
                if (KIND == 0 && CG == 1)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(td), "l"(da + (uint64_t)(j * 2)), "l"(db + (uint64_t)(j * 2)), "r"(idesc), "r"(acc) : "memory");
                if (KIND == 1 && CG == 1)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(td), "l"(da + (uint64_t)(j * 2)), "l"(db + (uint64_t)(j * 2)), "r"(idesc), "r"(acc) : "memory");
                if (KIND == 0 && CG == 2)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(td), "l"(da + (uint64_t)(j * 2)), "l"(db + (uint64_t)(j * 2)), "r"(idesc), "r"(acc) : "memory");
                if (KIND == 1 && CG == 2)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(td), "l"(da + (uint64_t)(j * 2)), "l"(db + (uint64_t)(j * 2)), "r"(idesc), "r"(acc) : "memory");
            }
        }
        if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        mbar_wait(smem_u32(bar), 0);
        cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        __syncthreads();
    }
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

template <int KIND, int CG>
static void run(const char *name, int sms, int iters) {
    const int smem = 4 * 49152 + 1024 + 64;
    CK(cudaFuncSetAttribute(peak_kernel<KIND, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    unsigned long long *cyc;
    CK(cudaMalloc(&cyc, sizeof(unsigned long long) * sms));
    CK(cudaMemset(cyc, 0, sizeof(unsigned long long) * sms));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms / CG * CG);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&cfg, peak_kernel<KIND, CG>, iters, cyc));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double K = KIND == 0 ? 32 : 16;
        const double ops = 2.0 * 128 * CG * 256 * K * 4.0 * iters * (cfg.gridDim.x / CG);
        unsigned long long c0 = 0;
        CK(cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost));
        printf("%-28s grid %3u  %8.3f ms  %8.1f T%s/s   CTA0: %.1f cycles per MMA instruction\n", name, cfg.gridDim.x, ms, ops / (ms * 1e-3) / 1e12,
               KIND == 0 ? "OP" : "FLOP", (double)c0 / (4.0 * iters));
    }
    CK(cudaFree(cyc));
}

int main(int argc, char **argv) {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    printf("tcgen05.mma issue-only peak, %d SMs, %d x 4 instructions per CTA (pair), N = 256, operands resident in shared memory\n", sms, iters);
    run<1, 1>("kind::f16 cta_group::1 M128", sms, iters);
    run<1, 2>("kind::f16 cta_group::2 M256", sms, iters);
    run<0, 1>("kind::i8  cta_group::1 M128", sms, iters);
    run<0, 2>("kind::i8  cta_group::2 M256", sms, iters);
    return 0;
}
