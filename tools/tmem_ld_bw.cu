// tmem_ld_bw.cu -- how fast can tcgen05.ld move accumulators from TMEM to registers?  (The prefill GEMM reads every 32-wide
// k-block's int32 partials back: 4 bytes per 32 MACs, so this number bounds it.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/tmem_ld_bw tools/tmem_ld_bw.cu && tools/_build/tmem_ld_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD_X32(regs, addr)                                                                                                     \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                   \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
                 : "=r"(regs[0]), "=r"(regs[1]), "=r"(regs[2]), "=r"(regs[3]), "=r"(regs[4]), "=r"(regs[5]), "=r"(regs[6]), "=r"(regs[7]),  \
                   "=r"(regs[8]), "=r"(regs[9]), "=r"(regs[10]), "=r"(regs[11]), "=r"(regs[12]), "=r"(regs[13]), "=r"(regs[14]),            \
                   "=r"(regs[15]), "=r"(regs[16]), "=r"(regs[17]), "=r"(regs[18]), "=r"(regs[19]), "=r"(regs[20]), "=r"(regs[21]),          \
                   "=r"(regs[22]), "=r"(regs[23]), "=r"(regs[24]), "=r"(regs[25]), "=r"(regs[26]), "=r"(regs[27]), "=r"(regs[28]),          \
                   "=r"(regs[29]), "=r"(regs[30]), "=r"(regs[31])                                                                          \
                 : "r"(addr))
#define LD_X64_PACK16(regs, addr)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "                                                         \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
                 : "=r"(regs[0]), "=r"(regs[1]), "=r"(regs[2]), "=r"(regs[3]), "=r"(regs[4]), "=r"(regs[5]), "=r"(regs[6]), "=r"(regs[7]),  \
                   "=r"(regs[8]), "=r"(regs[9]), "=r"(regs[10]), "=r"(regs[11]), "=r"(regs[12]), "=r"(regs[13]), "=r"(regs[14]),            \
                   "=r"(regs[15]), "=r"(regs[16]), "=r"(regs[17]), "=r"(regs[18]), "=r"(regs[19]), "=r"(regs[20]), "=r"(regs[21]),          \
                   "=r"(regs[22]), "=r"(regs[23]), "=r"(regs[24]), "=r"(regs[25]), "=r"(regs[26]), "=r"(regs[27]), "=r"(regs[28]),          \
                   "=r"(regs[29]), "=r"(regs[30]), "=r"(regs[31])                                                                          \
                 : "r"(addr))

// MODE 0: 32x32b.x32, wait after every load   1: two x32 loads in flight   2: x64.pack::16b (64 columns -> 32 registers)
template <int MODE>
__global__ void __launch_bounds__(512, 1) ld_kernel(int iters, unsigned long long *cycles, uint32_t *sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t a[32], b[32], acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        const uint32_t col = (uint32_t)(((warp >> 2) * 128 + (i & 1) * 64) & 511);
        if (MODE == 0) {
            LD_X32(a, base + col);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += a[0] ^ a[31];
        } else if (MODE == 1) {
            LD_X32(a, base + col);
            LD_X32(b, base + col + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += a[0] ^ b[31];
        } else {
            LD_X64_PACK16(a, base + col);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += a[0] ^ a[31];
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "n"(512) : "memory");
}

int main() {
    unsigned long long *cyc;
    uint32_t *sink;
    cudaMalloc(&cyc, 1024 * 8);
    cudaMalloc(&sink, 4);
    const int iters = 4096;
    const char *names[3] = {"32x32b.x32, one in flight", "32x32b.x32, two in flight", "32x32b.x32.pack::16b"};
    for (int mode = 0; mode < 3; mode++)
        for (int warps = 4; warps <= 16; warps *= 2) {
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0) ld_kernel<0><<<148, warps * 32>>>(iters, cyc, sink);
                if (mode == 1) ld_kernel<1><<<148, warps * 32>>>(iters, cyc, sink);
                if (mode == 2) ld_kernel<2><<<148, warps * 32>>>(iters, cyc, sink);
            }
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d warps %d: %s\n", mode, warps, cudaGetErrorString(e)); return 1; }
            unsigned long long h[148];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double c = 0;
            for (int i = 0; i < 148; i++) c += (double)h[i];
            c /= 148;
            // TMEM cells read per warp-load: 32 lanes x (32 | 64 | 64) columns x 4 bytes
            const double cells = (mode == 0 ? 32.0 : 64.0) * 32 * 4;
            const double regs = (mode == 1 ? 64.0 : 32.0) * 32 * 4;
            printf("%-28s %2d warps: %7.1f cycles per iteration per warp; per SM: %6.1f B/clk of TMEM cells, %6.1f B/clk into registers\n",
                   names[mode], warps, c / iters, cells * warps * iters / c, regs * warps * iters / c);
        }
    return 0;
}
