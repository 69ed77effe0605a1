// p2p_latency.cu -- NVLink small-message latency between two GPUs of one box, single process (cudaDeviceEnablePeerAccess).
// GPU0 kernel writes a tagged word into GPU1 memory, GPU1 kernel echoes it back into GPU0 memory; GPU0 times the round
// trip on its own clock.  Variants: how the word is stored (plain / volatile / +fence.sys) and how many bytes accompany it.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o p2p_latency p2p_latency.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t ldv(const uint32_t *p) { uint32_t v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }

template <int MODE>
__device__ __forceinline__ void put(uint32_t *p, uint32_t v) {
    if (MODE == 0) { *p = v; }
    else if (MODE == 1) { asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
    else if (MODE == 2) { *p = v; __threadfence_system(); }
    else { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
}

// ping side: for each iteration write payload words (lanes) + tag to the peer, wait for the echo in local memory
template <int MODE>
__global__ void ping(uint32_t *peer, const uint32_t *local, int iters, int payload_lanes, unsigned long long *out_ns) {
    unsigned long long t0 = 0, t1 = 0;
    for (int i = 1; i <= iters; i++) {
        if (i == 2 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        if ((int)threadIdx.x < payload_lanes) put<MODE>(peer + threadIdx.x, (uint32_t)i);
        if (threadIdx.x == 0) while (ldv(local) != (uint32_t)i) {}
        __syncthreads();
    }
    if (threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); *out_ns = (t1 - t0) / (iters - 1); }
}

// pong side: wait until ALL payload words carry i, then echo i back
template <int MODE>
__global__ void pong(uint32_t *peer, const uint32_t *local, int iters, int payload_lanes) {
    for (int i = 1; i <= iters; i++) {
        if ((int)threadIdx.x < payload_lanes) while (ldv(local + threadIdx.x) != (uint32_t)i) {}
        __syncthreads();
        if (threadIdx.x == 0) put<MODE>(peer, (uint32_t)i);
    }
}

int main() {
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("need 2 GPUs\n"); return 0; }
    uint32_t *b0, *b1;
    unsigned long long *ns;
    CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0)); CK(cudaMalloc(&b0, 4096)); CK(cudaMemset(b0, 0, 4096)); CK(cudaMallocManaged(&ns, 8));
    CK(cudaSetDevice(1)); CK(cudaDeviceEnablePeerAccess(0, 0)); CK(cudaMalloc(&b1, 4096)); CK(cudaMemset(b1, 0, 4096));
    cudaStream_t s0, s1;
    CK(cudaSetDevice(0)); CK(cudaStreamCreate(&s0));
    CK(cudaSetDevice(1)); CK(cudaStreamCreate(&s1));
    const int iters = 2000;
    const char *names[4] = {"plain st", "st.volatile", "st + fence.sys", "st.release.sys"};
    for (int lanes : {1, 32}) {
        for (int mode = 0; mode < 4; mode++) {
            CK(cudaSetDevice(0)); CK(cudaMemset(b0, 0, 4096));
            CK(cudaSetDevice(1)); CK(cudaMemset(b1, 0, 4096));
            CK(cudaDeviceSynchronize());
            CK(cudaSetDevice(1));
            switch (mode) {
                case 0: pong<0><<<1, 32, 0, s1>>>(b0, b1, iters, lanes); break;
                case 1: pong<1><<<1, 32, 0, s1>>>(b0, b1, iters, lanes); break;
                case 2: pong<2><<<1, 32, 0, s1>>>(b0, b1, iters, lanes); break;
                default: pong<3><<<1, 32, 0, s1>>>(b0, b1, iters, lanes); break;
            }
            CK(cudaSetDevice(0));
            switch (mode) {
                case 0: ping<0><<<1, 32, 0, s0>>>(b1, b0, iters, lanes, ns); break;
                case 1: ping<1><<<1, 32, 0, s0>>>(b1, b0, iters, lanes, ns); break;
                case 2: ping<2><<<1, 32, 0, s0>>>(b1, b0, iters, lanes, ns); break;
                default: ping<3><<<1, 32, 0, s0>>>(b1, b0, iters, lanes, ns); break;
            }
            CK(cudaStreamSynchronize(s0));
            CK(cudaSetDevice(1)); CK(cudaStreamSynchronize(s1));
            printf("%-16s payload %2d words: round trip %6llu ns  (one way ~%llu ns)\n", names[mode], lanes, *ns, *ns / 2);
        }
    }
    return 0;
}
