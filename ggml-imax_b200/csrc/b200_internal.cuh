// b200_internal.cuh -- shared internals of the sm_100a kernels behind include/ggml_b200.h
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "ggml_b200.h"

struct b200_ctx {
    int          device;
    cudaStream_t stream;
    bool         owns_stream;
    int          sm_count;
    int          cc_major, cc_minor;
    // scratch for quantized activations (qs plane, d plane, per-block sums) -- grown on demand
    void        *ws;
    size_t       ws_size;
    // staging for repack at set_tensor / un-repack at get_tensor
    void        *stage;
    size_t       stage_size;
    // options
    int          opt_pdl;
    int          opt_gemm;
    int          opt_gemv_max_n;
    int          opt_gemv_stream;
    int          opt_gemm_exact;        // 1: prefill GEMM = exact int8 block dots + fp32 scaling (slow); 0: fp16 tcgen05 path
    // decode plans (b200_plan.cu)
    int          opt_plan_pub_min_k;    // shortest in-plan src1 that is quantized once per GPU (0 = never)
    int          opt_plan_pub_dist;     // ... when its producer lies at least this many ops back
    int          opt_plan_l2_window;    // ring slots the L2 prefetcher runs ahead of the weight stream (0 = off)
    int          opt_plan_slots;        // ring slots (0 = as many as fit)
    int          opt_plan_evict_first;  // ring copies carry the L2 evict-first policy
    int          opt_plan_trace;        // plans created from now on record a device-side timeline
    int          opt_plan_timeout_ms;   // bound of every wait on global memory inside a plan (0 = default)
    // abort word: a kernel whose bounded wait expired leaves a code in abort_dev (device memory, polled by the other waits)
    // and in abort_host (pinned host memory, abort_host_dev = its device alias), where b200_synchronize finds it for free
    uint32_t    *abort_host, *abort_host_dev, *abort_dev;
    // optional device-side timeline: 8 x u64 %globaltimer stamps per (launch, CTA), see b200_ctx_set_trace
    unsigned long long *trace;
    int64_t      trace_capacity;   // in launches
    int64_t      trace_next;       // next launch slot
    int64_t      launches;
    int64_t      launches_mark;    // value at b200_graph_begin: recording launches nothing
    char         err[512];
};

void b200_set_error(b200_ctx *ctx, const char *fmt, ...);

#define B200_TRACE_STAMPS 8
#define B200_TRACE_MAX_CTAS 160

#define B200_CUDA_TRY(ctx, call)                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            b200_set_error((ctx), "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,             \
                           cudaGetErrorString(e__));                                               \
            (void)cudaGetLastError();                                                              \
            return B200_ERR_CUDA;                                                                  \
        }                                                                                          \
    } while (0)

#define B200_REQUIRE(ctx, cond, code)                                                              \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            b200_set_error((ctx), "%s:%d: requirement failed: %s", __FILE__, __LINE__, #cond);     \
            return (code);                                                                         \
        }                                                                                          \
    } while (0)

// cudaSetDevice only when the calling thread is on another device (the query is a thread-local read; a redundant set is not free
// on a launch-bound path such as a GPT-2 decode step)
static inline cudaError_t b200_use_device(int device) {
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur == device) return cudaSuccess;
    return cudaSetDevice(device);
}
// a kernel's dynamic shared-memory limit has to be raised once per device, not once per launch
#define B200_SMEM_LIMIT_ONCE(ctx, kern, bytes)                                                                         \
    do {                                                                                                               \
        static unsigned long long done__ = 0;                                                                          \
        static int bytes__ = 0;                                                                                        \
        const unsigned long long bit__ = 1ull << ((ctx)->device & 63);                                                 \
        if (!(done__ & bit__) || bytes__ < (int)(bytes)) {                                                             \
            B200_CUDA_TRY((ctx), cudaFuncSetAttribute((kern), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            done__ |= bit__;                                                                                           \
            if (bytes__ < (int)(bytes)) bytes__ = (int)(bytes);                                                        \
        }                                                                                                              \
    } while (0)

int b200_ws_reserve(b200_ctx *ctx, size_t bytes);     // ctx->ws >= bytes
int b200_stage_reserve(b200_ctx *ctx, size_t bytes);  // ctx->stage >= bytes

__host__ __device__ static inline size_t b200_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- storage geometry of a repacked quantized tensor -----------------------------------------
static inline int b200_qs_bytes(int type) { return type == B200_TYPE_Q4_0 ? 16 : 32; }
static inline int b200_wire_bytes(int type) { return type == B200_TYPE_Q4_0 ? B200_Q4_0_BYTES : B200_Q8_0_BYTES; }

// ---- launchers implemented in the .cu files --------------------------------------------------
int b200_launch_repack(b200_ctx *ctx, int type, void *tensor_dev, int64_t nblocks_total, const void *wire_dev,
                       int64_t block_off, int64_t nblocks);
int b200_launch_unrepack(b200_ctx *ctx, int type, const void *tensor_dev, int64_t nblocks_total, void *wire_dev,
                         int64_t block_off, int64_t nblocks);

struct b200_gemv_params {
    int            type;
    const uint8_t *qs;         // qs plane of src0 (already offset to its first block)
    const __half  *d;          // d plane of src0 (already offset)
    int64_t        k, m, ne02, ne03;
    const float   *x;          // src1
    int64_t        n, ne12, ne13;
    size_t         nb11, nb12, nb13;
    float         *dst;        // dense [ne13][ne12][dst_n][m], already offset to this launch's first column
    int64_t        dst_n;      // columns of the whole dst (>= n when the caller chunks columns)
    int32_t       *dots;       // non-null: dump per-block int32 partials [n][m][k/32] instead of dst
    unsigned long long *trace; // non-null: this launch's [gridDim.x][8] timestamp slots
    const b200_gather *gather; // host pointer (launcher copies it into a kernel parameter); null: plain local dst
    // optional epilogue of a plain 2-D launch (b200_mul_mat_fused): dst = act(W x + bias) + residual
    const float   *bias;       // [m] or null
    const float   *residual;   // dense like dst ([n][m]) or null; may be dst itself
    const float   *residual2;  // a second one, added after the first
    int            act;        // B200_EPI_NONE / B200_EPI_GELU
    // bound of the waits on peers' tagged stores (fused all-gather path); filled in by the launcher from the context
    uint32_t      *abort_dev, *abort_host;
    unsigned long long wait_timeout_ns;
};

// the epilogue the GEMV kernels apply to a finished dst element (row = weight row, idx = its index in the dense dst)
__device__ __forceinline__ float b200_gemv_epilogue(const b200_gemv_params &p, float v, int64_t row, int64_t idx) {
    if (p.bias) v += p.bias[row];
    if (p.act == B200_EPI_GELU) v = 0.5f * v * (1.0f + tanhf(0.79788456080286535587989211986876f * v * (1.0f + 0.044715f * v * v)));   // src/ggml.c:1966
    if (p.residual) v += p.residual[idx];
    if (p.residual2) v += p.residual2[idx];
    return v;
}
int b200_launch_gemv(b200_ctx *ctx, const b200_gemv_params &p);
// Q5_0 / IQ4_NL in wire format (b200_wire_formats.cu)
int b200_launch_gemv_wire(b200_ctx *ctx, const b200_mul_mat_args *a);
int b200_launch_get_rows_wire(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *rows, const b200_tensor *dst);
bool b200_try_launch_gemv_stream(b200_ctx *ctx, const b200_gemv_params &p, int *rc);
bool b200_try_launch_gemv_stream_batch(b200_ctx *ctx, const b200_gemv_params *ps, int count, int *rc);
int b200_launch_gather_finish(b200_ctx *ctx, const b200_gather &gd, const void *ll_src, float *out, int64_t count);

struct b200_gemm_params {
    int            type;
    const uint8_t *qs;
    const __half  *d;
    int64_t        k, m;       // one 2-D weight matrix
    const int8_t  *aq;         // quantized activations, planar [n][k]
    const __half  *ad;         // [n][k/32]
    int64_t        n;
    float         *dst;        // [n][m]
    int32_t       *dots;       // non-null: dump per-block int32 partials
    void          *scratch;    // b200_gemm_scratch_bytes(type, k, m, n) bytes
};
int b200_launch_gemm(b200_ctx *ctx, const b200_gemm_params &p);
bool b200_gemm_available(void);
size_t b200_gemm_scratch_bytes(int type, int64_t k, int64_t m, int64_t n);
// the fp16 prefill path (b200_gemm_f16.cu): quantize_row_q8_0 -> fp16 X' (one launch), then the persistent pair kernel
size_t b200_gemm_f16_scratch_bytes(int64_t k, int64_t m, int64_t n, int sm_count);
int b200_launch_gemm_f16(b200_ctx *ctx, int type, const uint8_t *qs, const __half *d, int64_t k, int64_t m, const float *x, int64_t n,
                         size_t x_row_stride, float *dst, void *scratch);
// activation scratch layout inside ctx->ws for one prefill mul_mat: [int8 plane n*k][fp16 scales n*k/32][gemm scratch]
size_t b200_prefill_ws_bytes(int type, int64_t k, int64_t m, int64_t n);
