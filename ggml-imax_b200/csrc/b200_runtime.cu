// b200_runtime.cu -- contexts, device buffers and transfers behind include/ggml_b200.h.
// Stands in for the buffer/stream plumbing of a ggml backend (src/ggml-backend-impl.h:18-117);
// no kernels of the hot path live here.
#include "b200_internal.cuh"

#include <stdarg.h>
#include <stdlib.h>

static char g_last_error[512] = "";

void b200_set_error(b200_ctx *ctx, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    snprintf(g_last_error, sizeof(g_last_error), "%s", buf);
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s", buf);
    if (getenv("B200_VERBOSE")) fprintf(stderr, "[ggml_b200] %s\n", buf);
}

extern "C" {

const char *b200_last_error(const b200_ctx *ctx) { return ctx ? ctx->err : g_last_error; }

int b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

int b200_device_info(int device, char *name, size_t name_len, size_t *free_bytes, size_t *total_bytes,
                     int *sm_count, int *cc_major, int *cc_minor) {
    cudaDeviceProp prop;
    B200_CUDA_TRY(NULL, cudaGetDeviceProperties(&prop, device));
    if (name && name_len) snprintf(name, name_len, "%s", prop.name);
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (free_bytes || total_bytes) {
        int cur = 0;
        B200_CUDA_TRY(NULL, cudaGetDevice(&cur));
        B200_CUDA_TRY(NULL, cudaSetDevice(device));
        size_t f = 0, t = 0;
        B200_CUDA_TRY(NULL, cudaMemGetInfo(&f, &t));
        if (free_bytes) *free_bytes = f;
        if (total_bytes) *total_bytes = t;
        B200_CUDA_TRY(NULL, cudaSetDevice(cur));
    }
    return B200_OK;
}

static int ctx_create_common(int device, void *stream, bool own, b200_ctx **out) {
    if (!out) return B200_ERR_INVALID;
    *out = NULL;
    int n = b200_device_count();
    if (n <= 0) {
        b200_set_error(NULL, "no CUDA device visible: the B200 backend has no CPU fallback");
        return B200_ERR_CUDA;
    }
    if (device < 0 || device >= n) {
        b200_set_error(NULL, "device %d out of range (0..%d)", device, n - 1);
        return B200_ERR_INVALID;
    }
    cudaDeviceProp prop;
    B200_CUDA_TRY(NULL, cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        b200_set_error(NULL, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major,
                       prop.minor);
        return B200_ERR_UNSUPPORTED;
    }
    b200_ctx *ctx = (b200_ctx *)calloc(1, sizeof(b200_ctx));
    if (!ctx) return B200_ERR_ALLOC;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    ctx->opt_pdl = 1;
    ctx->opt_gemm = 1;
    ctx->opt_gemv_max_n = 8;
    ctx->opt_gemv_stream = 1;
    ctx->opt_gemm_exact = 0;
    ctx->opt_plan_pub_min_k = 4096;
    ctx->opt_plan_pub_dist = 2;
    ctx->opt_plan_l2_window = 8;
    ctx->opt_plan_evict_first = 1;
    {
        cudaError_t e0 = cudaSetDevice(device);
        if (e0 == cudaSuccess) e0 = cudaHostAlloc((void **)&ctx->abort_host, 64, cudaHostAllocMapped);
        if (e0 == cudaSuccess) {
            memset(ctx->abort_host, 0, 64);
            e0 = cudaHostGetDevicePointer((void **)&ctx->abort_host_dev, ctx->abort_host, 0);
        }
        if (e0 == cudaSuccess) e0 = cudaMalloc((void **)&ctx->abort_dev, 64);
        if (e0 == cudaSuccess) e0 = cudaMemset(ctx->abort_dev, 0, 64);
        if (e0 != cudaSuccess) {
            b200_set_error(NULL, "context setup on device %d failed: %s", device, cudaGetErrorString(e0));
            (void)cudaGetLastError();
            if (ctx->abort_host) cudaFreeHost(ctx->abort_host);
            if (ctx->abort_dev) cudaFree(ctx->abort_dev);
            (void)cudaGetLastError();
            free(ctx);
            return B200_ERR_CUDA;
        }
    }
    if (own) {
        cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            b200_set_error(NULL, "cudaStreamCreate: %s", cudaGetErrorString(e));
            (void)cudaGetLastError();
            cudaFreeHost(ctx->abort_host);
            cudaFree(ctx->abort_dev);
            free(ctx);
            return B200_ERR_CUDA;
        }
        ctx->owns_stream = true;
    } else {
        ctx->stream = (cudaStream_t)stream;
        ctx->owns_stream = false;
    }
    *out = ctx;
    return B200_OK;
}

int b200_ctx_create(int device, b200_ctx **out) { return ctx_create_common(device, NULL, true, out); }
int b200_ctx_create_on_stream(int device, void *cuda_stream, b200_ctx **out) {
    return ctx_create_common(device, cuda_stream, false, out);
}

void b200_ctx_destroy(b200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->ws && cudaFree(ctx->ws) != cudaSuccess) (void)cudaGetLastError();
    if (ctx->stage && cudaFree(ctx->stage) != cudaSuccess) (void)cudaGetLastError();
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->abort_host) cudaFreeHost(ctx->abort_host);
    if (ctx->abort_dev) cudaFree(ctx->abort_dev);
    (void)cudaGetLastError();
    free(ctx);
}

void *b200_ctx_stream(const b200_ctx *ctx) { return ctx ? (void *)ctx->stream : NULL; }
int b200_ctx_device(const b200_ctx *ctx) { return ctx ? ctx->device : -1; }
int64_t b200_ctx_launch_count(const b200_ctx *ctx) { return ctx ? ctx->launches : 0; }

int b200_ctx_set_option(b200_ctx *ctx, const char *key, int64_t value) {
    if (!ctx || !key) return B200_ERR_INVALID;
    if (!strcmp(key, "pdl")) { ctx->opt_pdl = value != 0; return B200_OK; }
    if (!strcmp(key, "gemm")) { ctx->opt_gemm = value != 0; return B200_OK; }
    if (!strcmp(key, "gemv_stream")) { ctx->opt_gemv_stream = value != 0; return B200_OK; }
    if (!strcmp(key, "gemv_max_n")) {
        if (value < 1 || value > 8) return B200_ERR_INVALID;
        ctx->opt_gemv_max_n = (int)value;
        return B200_OK;
    }
    if (!strcmp(key, "gemm_exact")) { ctx->opt_gemm_exact = value != 0; return B200_OK; }
    if (!strcmp(key, "plan_pub_min_k")) { ctx->opt_plan_pub_min_k = value > 0 ? (int)value : 0; return B200_OK; }
    if (!strcmp(key, "plan_pub_dist")) { ctx->opt_plan_pub_dist = value >= 1 ? (int)value : 1; return B200_OK; }
    if (!strcmp(key, "plan_l2_window")) { ctx->opt_plan_l2_window = value > 0 ? (int)(value > 4096 ? 4096 : value) : 0; return B200_OK; }
    if (!strcmp(key, "plan_evict_first")) { ctx->opt_plan_evict_first = value != 0; return B200_OK; }
    if (!strcmp(key, "plan_slots")) { ctx->opt_plan_slots = value > 0 ? (int)value : 0; return B200_OK; }
    if (!strcmp(key, "plan_trace")) { ctx->opt_plan_trace = value != 0; return B200_OK; }
    if (!strcmp(key, "plan_timeout_ms")) { ctx->opt_plan_timeout_ms = value > 0 ? (int)value : 0; return B200_OK; }
    b200_set_error(ctx, "unknown option '%s'", key);
    return B200_ERR_INVALID;
}

int b200_ctx_set_trace(b200_ctx *ctx, void *trace_dev, int64_t max_launches) {
    B200_REQUIRE(ctx, ctx != NULL, B200_ERR_INVALID);
    ctx->trace = (unsigned long long *)trace_dev;
    ctx->trace_capacity = trace_dev ? max_launches : 0;
    ctx->trace_next = 0;
    return B200_OK;
}

int b200_malloc(b200_ctx *ctx, void **dptr, size_t size) {
    B200_REQUIRE(ctx, ctx && dptr, B200_ERR_INVALID);
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(dptr, size ? size : 1);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        b200_set_error(ctx, "cudaMalloc(%zu) failed: %s", size, cudaGetErrorString(e));
        *dptr = NULL;
        return B200_ERR_ALLOC;
    }
    return B200_OK;
}

int b200_free(b200_ctx *ctx, void *dptr) {
    B200_REQUIRE(ctx, ctx != NULL, B200_ERR_INVALID);
    if (!dptr) return B200_OK;
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    B200_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    B200_CUDA_TRY(ctx, cudaFree(dptr));
    return B200_OK;
}

int b200_memset(b200_ctx *ctx, void *dptr, int value, size_t size) {
    B200_REQUIRE(ctx, ctx != NULL, B200_ERR_INVALID);
    if (!size) return B200_OK;
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    B200_CUDA_TRY(ctx, cudaMemsetAsync(dptr, value, size, ctx->stream));
    B200_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

int b200_upload_async(b200_ctx *ctx, void *dst_dev, const void *src_host, size_t size) {
    B200_REQUIRE(ctx, ctx != NULL, B200_ERR_INVALID);
    if (!size) return B200_OK;
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    B200_CUDA_TRY(ctx, cudaMemcpyAsync(dst_dev, src_host, size, cudaMemcpyHostToDevice, ctx->stream));
    return B200_OK;
}

int b200_download_async(b200_ctx *ctx, void *dst_host, const void *src_dev, size_t size) {
    B200_REQUIRE(ctx, ctx != NULL, B200_ERR_INVALID);
    if (!size) return B200_OK;
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    B200_CUDA_TRY(ctx, cudaMemcpyAsync(dst_host, src_dev, size, cudaMemcpyDeviceToHost, ctx->stream));
    return B200_OK;
}

struct b200_event {
    cudaEvent_t ev;
    int device;
};

int b200_event_create(b200_ctx *ctx, b200_event **out) {
    B200_REQUIRE(ctx, ctx && out, B200_ERR_INVALID);
    *out = NULL;
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    b200_event *e = (b200_event *)calloc(1, sizeof(b200_event));
    if (!e) return B200_ERR_ALLOC;
    e->device = ctx->device;
    const cudaError_t rc = cudaEventCreateWithFlags(&e->ev, cudaEventDisableTiming);
    if (rc != cudaSuccess) {
        b200_set_error(ctx, "cudaEventCreateWithFlags failed: %s", cudaGetErrorString(rc));
        (void)cudaGetLastError();
        free(e);
        return B200_ERR_CUDA;
    }
    *out = e;
    return B200_OK;
}

void b200_event_destroy(b200_event *ev) {
    if (!ev) return;
    if (cudaSetDevice(ev->device) == cudaSuccess) (void)cudaEventDestroy(ev->ev);
    (void)cudaGetLastError();
    free(ev);
}

int b200_event_record(b200_ctx *ctx, b200_event *ev) {
    B200_REQUIRE(ctx, ctx && ev && ev->device == ctx->device, B200_ERR_INVALID);
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    B200_CUDA_TRY(ctx, cudaEventRecord(ev->ev, ctx->stream));
    return B200_OK;
}

int b200_event_wait(b200_ctx *ctx, b200_event *ev) {
    B200_REQUIRE(ctx, ctx && ev, B200_ERR_INVALID);
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    B200_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ev->ev, 0));
    return B200_OK;
}

int b200_event_synchronize(b200_event *ev) {
    if (!ev) return B200_ERR_INVALID;
    if (cudaSetDevice(ev->device) != cudaSuccess || cudaEventSynchronize(ev->ev) != cudaSuccess) {
        b200_set_error(NULL, "cudaEventSynchronize failed: %s", cudaGetErrorString(cudaGetLastError()));
        return B200_ERR_CUDA;
    }
    return B200_OK;
}

// ---- a sequence of this context's launches recorded once and replayed as one CUDA graph ---------------------------------------
// (what the reference's CUDA backend does for a decode step, src/ggml-cuda.cu:2461-2709; here behind ggml_backend_graph_plan_*)
struct b200_graph {
    cudaGraphExec_t exec;
    int device;
    int64_t kernel_nodes;
};

int b200_graph_begin(b200_ctx *ctx) {
    B200_REQUIRE(ctx, ctx != NULL, B200_ERR_INVALID);
    B200_CUDA_TRY(ctx, b200_use_device(ctx->device));
    // relaxed: the recording thread may still allocate (a scratch area that has to grow invalidates nothing)
    B200_CUDA_TRY(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
    ctx->launches_mark = ctx->launches;
    return B200_OK;
}

int b200_graph_end(b200_ctx *ctx, b200_graph **out) {
    B200_REQUIRE(ctx, ctx != NULL, B200_ERR_INVALID);
    if (out) *out = NULL;
    cudaGraph_t graph = NULL;
    const cudaError_t rc = cudaStreamEndCapture(ctx->stream, &graph);
    ctx->launches = ctx->launches_mark;
    if (rc != cudaSuccess || graph == NULL) {
        b200_set_error(ctx, "the recorded sequence cannot be a CUDA graph: %s", cudaGetErrorString(rc));
        (void)cudaGetLastError();
        if (graph) (void)cudaGraphDestroy(graph);
        return B200_ERR_UNSUPPORTED;
    }
    if (!out) {
        (void)cudaGraphDestroy(graph);
        return B200_OK;
    }
    b200_graph *g = (b200_graph *)calloc(1, sizeof(b200_graph));
    if (!g) {
        (void)cudaGraphDestroy(graph);
        return B200_ERR_ALLOC;
    }
    g->device = ctx->device;
    size_t n = 0;
    if (cudaGraphGetNodes(graph, NULL, &n) == cudaSuccess) g->kernel_nodes = (int64_t)n;
    const cudaError_t ri = cudaGraphInstantiate(&g->exec, graph, 0);
    (void)cudaGraphDestroy(graph);
    if (ri != cudaSuccess) {
        b200_set_error(ctx, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ri));
        (void)cudaGetLastError();
        free(g);
        return B200_ERR_UNSUPPORTED;
    }
    *out = g;
    return B200_OK;
}

int b200_graph_launch(b200_ctx *ctx, b200_graph *g) {
    B200_REQUIRE(ctx, ctx && g && g->device == ctx->device, B200_ERR_INVALID);
    B200_CUDA_TRY(ctx, b200_use_device(ctx->device));
    B200_CUDA_TRY(ctx, cudaGraphLaunch(g->exec, ctx->stream));
    ctx->launches += g->kernel_nodes;
    return B200_OK;
}

int64_t b200_graph_node_count(const b200_graph *g) { return g ? g->kernel_nodes : 0; }

void b200_graph_destroy(b200_graph *g) {
    if (!g) return;
    if (cudaSetDevice(g->device) == cudaSuccess) (void)cudaGraphExecDestroy(g->exec);
    (void)cudaGetLastError();
    free(g);
}

int b200_upload(b200_ctx *ctx, void *dst_dev, const void *src_host, size_t size) {
    int rc = b200_upload_async(ctx, dst_dev, src_host, size);
    if (rc != B200_OK) return rc;
    B200_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

int b200_download(b200_ctx *ctx, void *dst_host, const void *src_dev, size_t size) {
    int rc = b200_download_async(ctx, dst_host, src_dev, size);
    if (rc != B200_OK) return rc;
    B200_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

int b200_copy_d2d(b200_ctx *ctx, void *dst_dev, const void *src_dev, size_t size) {
    B200_REQUIRE(ctx, ctx != NULL, B200_ERR_INVALID);
    if (!size) return B200_OK;
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    B200_CUDA_TRY(ctx, cudaMemcpyAsync(dst_dev, src_dev, size, cudaMemcpyDefault, ctx->stream));
    return B200_OK;
}

int b200_copy_2d(b200_ctx *ctx, void *dst_dev, size_t dpitch, const void *src_dev, size_t spitch, size_t width, size_t height) {
    B200_REQUIRE(ctx, ctx != NULL, B200_ERR_INVALID);
    if (!width || !height) return B200_OK;
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    B200_CUDA_TRY(ctx, cudaMemcpy2DAsync(dst_dev, dpitch, src_dev, spitch, width, height, cudaMemcpyDefault, ctx->stream));
    return B200_OK;
}

int b200_enable_peer_access(b200_ctx *ctx, int peer_device) {
    B200_REQUIRE(ctx, ctx != NULL, B200_ERR_INVALID);
    if (peer_device == ctx->device) return B200_OK;
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int can = 0;
    B200_CUDA_TRY(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
    if (!can) {
        b200_set_error(ctx, "device %d cannot access device %d directly", ctx->device, peer_device);
        return B200_ERR_UNSUPPORTED;
    }
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        b200_set_error(ctx, "cudaDeviceEnablePeerAccess(%d -> %d): %s", ctx->device, peer_device, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return B200_ERR_CUDA;
    }
    (void)cudaGetLastError();
    return B200_OK;
}

int b200_synchronize(b200_ctx *ctx) {
    B200_REQUIRE(ctx, ctx != NULL, B200_ERR_INVALID);
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    B200_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->abort_host && *(volatile uint32_t *)ctx->abort_host != 0u) {
        // a bounded wait inside a kernel expired: what 1 = tagged vector, 2 = publication count, 3 = export (b200_plan.cu), 4 = split-k slices
        // (b200_gemm_f16.cu), 5 = tagged src1 of the fused all-gather GEMV, 6 = gather_finish (b200_gemv_stream.cu)
        const uint32_t code = *(volatile uint32_t *)ctx->abort_host;
        *(volatile uint32_t *)ctx->abort_host = 0u;
        B200_CUDA_TRY(ctx, cudaMemsetAsync(ctx->abort_dev, 0, 64, ctx->stream));
        B200_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        b200_set_error(ctx, "a kernel gave up waiting (wait kind %u, detail 0x%x): a decode plan or fused all-gather whose peer rank died or launched a different sequence; results of this launch are invalid",
                       code & 0xffu, (code & 0x7fffffffu) >> 8);
        return B200_ERR_CUDA;
    }
    return B200_OK;
}

int b200_reserve_workspace(b200_ctx *ctx, int type, int64_t k, int64_t m, int64_t n) {
    B200_REQUIRE(ctx, ctx && k > 0 && n > 0 && m > 0 && k % 32 == 0, B200_ERR_INVALID);
    return b200_ws_reserve(ctx, b200_prefill_ws_bytes(type, k, m, n));
}

int b200_ipc_export(b200_ctx *ctx, void *dptr, void *handle64_out) {
    B200_REQUIRE(ctx, ctx && dptr && handle64_out, B200_ERR_INVALID);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    B200_CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle64_out, &h, sizeof(h));
    return B200_OK;
}

int b200_ipc_import(b200_ctx *ctx, const void *handle64, void **peer_ptr_out) {
    B200_REQUIRE(ctx, ctx && handle64 && peer_ptr_out, B200_ERR_INVALID);
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    B200_CUDA_TRY(ctx, cudaIpcOpenMemHandle(peer_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return B200_OK;
}

int b200_ipc_close(b200_ctx *ctx, void *peer_ptr) {
    B200_REQUIRE(ctx, ctx != NULL, B200_ERR_INVALID);
    if (!peer_ptr) return B200_OK;
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    B200_CUDA_TRY(ctx, cudaIpcCloseMemHandle(peer_ptr));
    return B200_OK;
}

int b200_host_malloc(void **hptr, size_t size) {
    if (!hptr) return B200_ERR_INVALID;
    cudaError_t e = cudaMallocHost(hptr, size ? size : 1);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        b200_set_error(NULL, "cudaMallocHost(%zu) failed: %s", size, cudaGetErrorString(e));
        *hptr = NULL;
        return B200_ERR_ALLOC;
    }
    return B200_OK;
}

int b200_host_free(void *hptr) {
    if (!hptr) return B200_OK;
    B200_CUDA_TRY(NULL, cudaFreeHost(hptr));
    return B200_OK;
}

}  // extern "C"

static int reserve(b200_ctx *ctx, void **p, size_t *cur, size_t bytes) {
    if (*cur >= bytes) return B200_OK;
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (*p) {
        B200_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        B200_CUDA_TRY(ctx, cudaFree(*p));
        *p = NULL;
        *cur = 0;
    }
    size_t want = b200_align_up(bytes + bytes / 4, 1 << 20);
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        b200_set_error(ctx, "workspace cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        return B200_ERR_ALLOC;
    }
    *cur = want;
    return B200_OK;
}

int b200_ws_reserve(b200_ctx *ctx, size_t bytes) { return reserve(ctx, &ctx->ws, &ctx->ws_size, bytes); }
int b200_stage_reserve(b200_ctx *ctx, size_t bytes) { return reserve(ctx, &ctx->stage, &ctx->stage_size, bytes); }
