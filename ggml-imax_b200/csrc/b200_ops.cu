// b200_ops.cu -- the operators either side of the quantized mul_mat in a GPT-2 / GPT-J decode graph, so that the
// reference's gpt-2-backend computes its WHOLE graph on this backend (examples/gpt-2/main-backend.cpp:442-717, :768).
//
// Each entry point stands in for one ggml_compute_forward_* of the reference CPU backend and is checked by the reference's own
// tests/test-backend-ops.cpp cases for that op (unmodified, against the reference CPU backend):
//   b200_op_get_rows       ggml_compute_forward_get_rows      src/ggml.c:13049 (_q :12874 -> dequantize_row_q4_0/_q8_0, _f16, _f32)
//   b200_op_binary         ggml_compute_forward_add_f32 / mul_f32 / div_f32     src/ggml.c:8568, :9687 (broadcast of src1 in every dim)
//   b200_op_unary          ggml_compute_forward_unary (gelu :10960, gelu_quick, silu, relu, tanh, ...)
//   b200_op_norm           ggml_compute_forward_norm_f32 :11353 / rms_norm_f32, optionally fused with the MUL (gain) and ADD (bias)
//                          that follow it in every transformer block
//   b200_op_scale          ggml_compute_forward_scale_f32 :12637
//   b200_op_diag_mask_inf  ggml_compute_forward_diag_mask_f32 :13301
//   b200_op_soft_max       ggml_compute_forward_soft_max_f32 :13393 (mask, scale, ALiBi max_bias), optionally fused with the SCALE
//                          and DIAG_MASK_INF in front of it
//   b200_op_copy           ggml_compute_forward_dup / cpy / cont :8535, :12818, :12826 (strided 4-D, F32/F16 conversions)
//   b200_op_mul_mat_dense  ggml_compute_forward_mul_mat :11808 with F32 / F16 src0 (K*Q and V*softmax(KQ) on permuted views)
//
// All of them are HBM-/latency-bound glue: coalesced 128-bit accesses where the layout allows, one CTA per row for the row
// reductions, no tensor cores.  Every kernel is safe for dst == src0 (ggml_gallocr computes these ops in place).
#include "b200_internal.cuh"

#include <math.h>

namespace {

// Every kernel of this file starts with pdl_enter(): it lets the next kernel of the stream begin launching at once (that kernel waits for
// THIS grid to complete before it touches memory -- a GEMV of the path fills its weight ring first, which depends on nothing) and then waits
// for the grid in front of it.  launch_k adds the launch attribute that makes the pair of instructions mean something (context option "pdl");
// without it both are no-ops.  A decode step of a whole model is a chain of ~2-5 us kernels: their launch latencies overlap instead of adding up.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
struct DynSmem { size_t bytes; };
template <typename... KArgs, typename... Args>
void launch_ks(b200_ctx *ctx, void (*kern)(KArgs...), dim3 grid, dim3 block, DynSmem smem, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem.bytes;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ctx->opt_pdl ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);      // errors surface in finish() through cudaGetLastError
}
template <typename... KArgs, typename... Args>
void launch_k(b200_ctx *ctx, void (*kern)(KArgs...), dim3 grid, dim3 block, Args... args) {
    launch_ks(ctx, kern, grid, block, DynSmem{0}, args...);
}

struct T4 {                      // a 4-D strided tensor as the kernels see it
    char *p;
    int64_t ne0, ne1, ne2, ne3;
    int64_t nb0, nb1, nb2, nb3;
};

T4 view(const b200_tensor *t) {
    T4 v;
    v.p = static_cast<char *>(t->data);
    v.ne0 = t->ne[0]; v.ne1 = t->ne[1]; v.ne2 = t->ne[2]; v.ne3 = t->ne[3];
    v.nb0 = t->nb[0]; v.nb1 = t->nb[1]; v.nb2 = t->nb[2]; v.nb3 = t->nb[3];
    return v;
}

int64_t nelements(const b200_tensor *t) { return t->ne[0] * t->ne[1] * t->ne[2] * t->ne[3]; }
int64_t nrows(const b200_tensor *t) { return t->ne[1] * t->ne[2] * t->ne[3]; }
int elt_size(int type) {
    switch (type) {
    case B200_TYPE_F32: case B200_TYPE_I32: return 4;
    case B200_TYPE_F16: case B200_TYPE_I16: return 2;
    default: return 0;
    }
}
bool contiguous(const b200_tensor *t) {
    const int64_t e = elt_size(t->type);
    return e && t->nb[0] == e && t->nb[1] == t->nb[0] * t->ne[0] && t->nb[2] == t->nb[1] * t->ne[1] && t->nb[3] == t->nb[2] * t->ne[2];
}
bool same_shape(const b200_tensor *a, const b200_tensor *b) {
    return a->ne[0] == b->ne[0] && a->ne[1] == b->ne[1] && a->ne[2] == b->ne[2] && a->ne[3] == b->ne[3];
}
bool aligned16(const b200_tensor *t) {
    return ((uintptr_t)t->data & 15) == 0 && t->nb[1] % 16 == 0 && t->nb[2] % 16 == 0 && t->nb[3] % 16 == 0;
}
int grid_for(int64_t work, int threads, int sm_count) {
    const int64_t g = (work + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide reductions over up to 32 warps; `red` = 33 floats of shared memory; every thread gets the result
template <bool MAX> __device__ __forceinline__ float block_reduce(float v, float *red) {
    v = MAX ? warp_max(v) : warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();                       // `red` may still be read from the previous reduction
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float w = lane < nw ? red[lane] : (MAX ? -INFINITY : 0.0f);
        w = MAX ? warp_max(w) : warp_sum(w);
        if (lane == 0) red[32] = w;
    }
    __syncthreads();
    return red[32];
}

// ---- GET_ROWS ---------------------------------------------------------------------------------------------------------------------
// dst[i10, i11, i12][0..nc) = row src1[i10, i11, i12] of src0[., i11, i12]; one thread per 4 consecutive output elements.
// Quantized rows come straight from the repacked planes: block b of the root tensor has its quants at qs + b * QSB and its scale at
// d[b]; Q4_0 element j < 16 is the low nibble of byte j, element j + 16 the high nibble (src/ggml-quants.c:980-998).
template <int TYPE>
__global__ void __launch_bounds__(256) get_rows_kernel(const char *__restrict__ src0, int64_t nb01, int64_t nb02, int64_t nb03, int64_t ne01,
                                                       const uint8_t *__restrict__ qs, const __half *__restrict__ qd, int64_t qoff,
                                                       const char *__restrict__ rows, int64_t nb10, int64_t nb11, int64_t nb12,
                                                       int64_t ne10, int64_t ne11, float *__restrict__ dst, int64_t nb1, int64_t nb2, int64_t nb3,
                                                       int64_t nc, int64_t nr, int *__restrict__ bad) {
    pdl_enter();
    const int64_t nc4 = nc >> 2;
    const int64_t total = nr * nc4;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / nc4, c = (t - r * nc4) << 2;
        const int64_t i12 = r / (ne11 * ne10), i11 = (r - i12 * ne11 * ne10) / ne10, i10 = r - i12 * ne11 * ne10 - i11 * ne10;
        const int64_t row = *reinterpret_cast<const int32_t *>(rows + i10 * nb10 + i11 * nb11 + i12 * nb12);
        float4 o;
        if (row < 0 || row >= ne01) {       // the reference would read out of bounds; report instead
            if (bad) *bad = 1;
            o = make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (TYPE == B200_TYPE_F32) {
            o = *reinterpret_cast<const float4 *>(src0 + row * nb01 + i11 * nb02 + i12 * nb03 + c * 4);
        } else if (TYPE == B200_TYPE_F16) {
            const uint2 h = *reinterpret_cast<const uint2 *>(src0 + row * nb01 + i11 * nb02 + i12 * nb03 + c * 2);
            const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&h.x)), b = __half22float2(*reinterpret_cast<const __half2 *>(&h.y));
            o = make_float4(a.x, a.y, b.x, b.y);
        } else {
            constexpr int WIRE = TYPE == B200_TYPE_Q4_0 ? B200_Q4_0_BYTES : B200_Q8_0_BYTES;
            const int64_t blk = qoff + (row * nb01 + i11 * nb02 + i12 * nb03) / WIRE + (c >> 5);
            const int j = (int)(c & 31);
            const float d = __half2float(qd[blk]);
            if (TYPE == B200_TYPE_Q4_0) {
                const uint32_t w = *reinterpret_cast<const uint32_t *>(qs + blk * 16 + (j & 15));
                const uint32_t n = j < 16 ? (w & 0x0F0F0F0Fu) : ((w >> 4) & 0x0F0F0F0Fu);
                o = make_float4(((int)(n & 0xff) - 8) * d, ((int)((n >> 8) & 0xff) - 8) * d, ((int)((n >> 16) & 0xff) - 8) * d, ((int)(n >> 24) - 8) * d);
            } else {
                const char4 q = *reinterpret_cast<const char4 *>(qs + blk * 32 + j);
                o = make_float4(q.x * d, q.y * d, q.z * d, q.w * d);
            }
        }
        *reinterpret_cast<float4 *>(reinterpret_cast<char *>(dst) + i10 * nb1 + i11 * nb2 + i12 * nb3 + c * 4) = o;
    }
}
// any nc / alignment (F32 and F16 sources only): one thread per element
template <int TYPE>
__global__ void __launch_bounds__(256) get_rows_scalar_kernel(const char *__restrict__ src0, int64_t nb01, int64_t nb02, int64_t nb03, int64_t ne01,
                                                              const char *__restrict__ rows, int64_t nb10, int64_t nb11, int64_t nb12, int64_t ne10,
                                                              int64_t ne11, char *__restrict__ dst, int64_t nb1, int64_t nb2, int64_t nb3, int64_t nc,
                                                              int64_t nr, int *__restrict__ bad) {
    pdl_enter();
    const int64_t total = nr * nc;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / nc, c = t - r * nc;
        const int64_t i12 = r / (ne11 * ne10), i11 = (r - i12 * ne11 * ne10) / ne10, i10 = r - i12 * ne11 * ne10 - i11 * ne10;
        const int64_t row = *reinterpret_cast<const int32_t *>(rows + i10 * nb10 + i11 * nb11 + i12 * nb12);
        float v = 0.f;
        if (row < 0 || row >= ne01) {
            if (bad) *bad = 1;
        } else if (TYPE == B200_TYPE_F32) {
            v = *reinterpret_cast<const float *>(src0 + row * nb01 + i11 * nb02 + i12 * nb03 + c * 4);
        } else {
            v = __half2float(*reinterpret_cast<const __half *>(src0 + row * nb01 + i11 * nb02 + i12 * nb03 + c * 2));
        }
        *reinterpret_cast<float *>(dst + i10 * nb1 + i11 * nb2 + i12 * nb3 + c * 4) = v;
    }
}

// ---- ADD / MUL / DIV with broadcast ------------------------------------------------------------------------------------------------
template <int OP> __device__ __forceinline__ float bin(float a, float b) { return OP == B200_OP_ADD ? a + b : (OP == B200_OP_MUL ? a * b : a / b); }

// rows are dense in all three tensors and src1 covers whole rows (ne10 == ne0): float4 per thread
template <int OP>
__global__ void __launch_bounds__(256) binary_rows_kernel(T4 a, T4 b, T4 d) {
    pdl_enter();
    const int64_t n4 = d.ne0 >> 2;
    const int64_t total = n4 * d.ne1 * d.ne2 * d.ne3;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = t / n4;
        const int64_t c = (t - r * n4) << 4;     // byte offset inside the row
        const int64_t i3 = r / (d.ne2 * d.ne1);
        r -= i3 * d.ne2 * d.ne1;
        const int64_t i2 = r / d.ne1, i1 = r - i2 * d.ne1;
        const float4 x = *reinterpret_cast<const float4 *>(a.p + i3 * a.nb3 + i2 * a.nb2 + i1 * a.nb1 + c);
        const float4 y = *reinterpret_cast<const float4 *>(b.p + (i3 % b.ne3) * b.nb3 + (i2 % b.ne2) * b.nb2 + (i1 % b.ne1) * b.nb1 + c);
        *reinterpret_cast<float4 *>(d.p + i3 * d.nb3 + i2 * d.nb2 + i1 * d.nb1 + c) =
            make_float4(bin<OP>(x.x, y.x), bin<OP>(x.y, y.y), bin<OP>(x.z, y.z), bin<OP>(x.w, y.w));
    }
}
// anything else: one thread per element, src1 indexed modulo its extent in every dimension (ggml_can_repeat(src1, src0))
template <int OP>
__global__ void __launch_bounds__(256) binary_generic_kernel(T4 a, T4 b, T4 d) {
    pdl_enter();
    const int64_t total = d.ne0 * d.ne1 * d.ne2 * d.ne3;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = t / d.ne0;
        const int64_t i0 = t - r * d.ne0;
        const int64_t i3 = r / (d.ne2 * d.ne1);
        r -= i3 * d.ne2 * d.ne1;
        const int64_t i2 = r / d.ne1, i1 = r - i2 * d.ne1;
        const float x = *reinterpret_cast<const float *>(a.p + i3 * a.nb3 + i2 * a.nb2 + i1 * a.nb1 + i0 * a.nb0);
        const float y = *reinterpret_cast<const float *>(b.p + (i3 % b.ne3) * b.nb3 + (i2 % b.ne2) * b.nb2 + (i1 % b.ne1) * b.nb1 + (i0 % b.ne0) * b.nb0);
        *reinterpret_cast<float *>(d.p + i3 * d.nb3 + i2 * d.nb2 + i1 * d.nb1 + i0 * d.nb0) = bin<OP>(x, y);
    }
}

// ---- unary / scale / diag_mask_inf: contiguous elementwise --------------------------------------------------------------------------
__device__ __forceinline__ float unary_apply(int op, float x) {
    switch (op) {
    case B200_UNARY_ABS: return fabsf(x);
    case B200_UNARY_SGN: return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f);
    case B200_UNARY_NEG: return -x;
    case B200_UNARY_STEP: return x > 0.f ? 1.f : 0.f;
    case B200_UNARY_TANH: return tanhf(x);
    case B200_UNARY_ELU: return x > 0.f ? x : expm1f(x);
    case B200_UNARY_RELU: return fmaxf(x, 0.f);
    case B200_UNARY_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case B200_UNARY_GELU: return 0.5f * x * (1.0f + tanhf(0.79788456080286535587989211986876f * x * (1.0f + 0.044715f * x * x)));   // src/ggml.c:1966
    case B200_UNARY_GELU_QUICK: return x * (1.0f / (1.0f + expf(-1.702f * x)));                                                  // :2000
    case B200_UNARY_SILU: return x / (1.0f + expf(-x));
    case B200_UNARY_HARDSWISH: return x * fminf(1.0f, fmaxf(0.0f, (x + 3.0f) / 6.0f));
    case B200_UNARY_HARDSIGMOID: return fminf(1.0f, fmaxf(0.0f, (x + 3.0f) / 6.0f));
    default: return x;
    }
}
// MODE 0: unary(op), 1: x * s, 2: diag_mask_inf(n_past = op; nc, nr)
template <int MODE>
__global__ void __launch_bounds__(256) elementwise_kernel(const float *__restrict__ x, float *__restrict__ y, int64_t n, int op, float s, int64_t nc, int64_t nr) {
    pdl_enter();
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const float v = x[t];
        float o;
        if (MODE == 0) o = unary_apply(op, v);
        else if (MODE == 1) o = v * s;
        else {
            const int64_t r = t / nc, i = t - r * nc, j = r % nr;
            o = i > op + j ? -INFINITY : v;
        }
        y[t] = o;
    }
}

// ---- NORM / RMS_NORM (+ gain, + bias) ---------------------------------------------------------------------------------------------------
// one CTA per row; the row is read three times (sum, squared deviations, output) from L1; writes only after the last reduction, so
// dst == src is fine.  gain / bias: optional row vectors of ne0 floats (the MUL and ADD that follow NORM in a transformer block).
template <bool RMS>
__global__ void __launch_bounds__(256) norm_kernel(T4 a, T4 d, const float *__restrict__ gain, const float *__restrict__ bias, float eps) {
    pdl_enter();
    __shared__ float red[33];
    int64_t r = blockIdx.x;
    const int64_t i3 = r / (a.ne2 * a.ne1);
    r -= i3 * a.ne2 * a.ne1;
    const int64_t i2 = r / a.ne1, i1 = r - i2 * a.ne1;
    const float *x = reinterpret_cast<const float *>(a.p + i3 * a.nb3 + i2 * a.nb2 + i1 * a.nb1);
    float *y = reinterpret_cast<float *>(d.p + i3 * d.nb3 + i2 * d.nb2 + i1 * d.nb1);
    const int64_t n = a.ne0;
    float mean = 0.f;
    if (!RMS) {
        float s = 0.f;
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
        mean = block_reduce<false>(s, red) / (float)n;
    }
    float s2 = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = x[i] - mean;
        s2 += v * v;
    }
    const float var = block_reduce<false>(s2, red) / (float)n;
    const float scale = 1.0f / sqrtf(var + eps);
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        float v = (x[i] - mean) * scale;
        if (gain) v *= gain[i];
        if (bias) v += bias[i];
        y[i] = v;
    }
}

// rows of up to 128 * VPL floats, 16-byte aligned: ONE warp per row keeps the row in registers -- every load of the row (and of gain /
// bias) is issued before the first reduction, so a decode step pays one memory latency instead of three, and no CTA barrier at all
template <bool RMS, int VPL>
__global__ void __launch_bounds__(128) norm_warp_kernel(T4 a, T4 d, const float *__restrict__ gain, const float *__restrict__ bias, float eps, int64_t nrows) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= nrows) return;
    const int64_t i3 = r / (a.ne2 * a.ne1);
    r -= i3 * a.ne2 * a.ne1;
    const int64_t i2 = r / a.ne1, i1 = r - i2 * a.ne1;
    const float4 *x = reinterpret_cast<const float4 *>(a.p + i3 * a.nb3 + i2 * a.nb2 + i1 * a.nb1);
    float4 *y = reinterpret_cast<float4 *>(d.p + i3 * d.nb3 + i2 * d.nb2 + i1 * d.nb1);
    const int n4 = (int)(a.ne0 >> 2);
    float4 v[VPL], g[VPL], b[VPL];
#pragma unroll
    for (int j = 0; j < VPL; j++) {
        const int i = lane + 32 * j;
        v[j] = i < n4 ? x[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (gain) g[j] = i < n4 ? reinterpret_cast<const float4 *>(gain)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) b[j] = i < n4 ? reinterpret_cast<const float4 *>(bias)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float inv_n = 1.0f / (float)a.ne0;
    float mean = 0.f;
    if (!RMS) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < VPL; j++) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
        mean = warp_sum(s) * inv_n;
    }
    float s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; j++) {
        if (lane + 32 * j < n4) {
            v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
            s2 += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
        }
    }
    const float scale = 1.0f / sqrtf(warp_sum(s2) * inv_n + eps);
#pragma unroll
    for (int j = 0; j < VPL; j++) {
        const int i = lane + 32 * j;
        if (i >= n4) continue;
        float4 o = make_float4(v[j].x * scale, v[j].y * scale, v[j].z * scale, v[j].w * scale);
        if (gain) { o.x *= g[j].x; o.y *= g[j].y; o.z *= g[j].z; o.w *= g[j].w; }
        if (bias) { o.x += b[j].x; o.y += b[j].y; o.z += b[j].z; o.w += b[j].w; }
        y[i] = o;
    }
}

// rows of 1025 .. 256 * 4 * VPT floats (GPT-J's 4096): one CTA per row, the row in registers, every load issued before the first reduction
template <bool RMS, int VPT>
__global__ void __launch_bounds__(256) norm_block_kernel(T4 a, T4 d, const float *__restrict__ gain, const float *__restrict__ bias, float eps) {
    pdl_enter();
    __shared__ float red[33];
    int64_t r = blockIdx.x;
    const int64_t i3 = r / (a.ne2 * a.ne1);
    r -= i3 * a.ne2 * a.ne1;
    const int64_t i2 = r / a.ne1, i1 = r - i2 * a.ne1;
    const float4 *x = reinterpret_cast<const float4 *>(a.p + i3 * a.nb3 + i2 * a.nb2 + i1 * a.nb1);
    float4 *y = reinterpret_cast<float4 *>(d.p + i3 * d.nb3 + i2 * d.nb2 + i1 * d.nb1);
    const int n4 = (int)(a.ne0 >> 2);
    float4 v[VPT], g[VPT], b[VPT];
#pragma unroll
    for (int j = 0; j < VPT; j++) {
        const int i = threadIdx.x + 256 * j;
        v[j] = i < n4 ? x[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (gain) g[j] = i < n4 ? reinterpret_cast<const float4 *>(gain)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) b[j] = i < n4 ? reinterpret_cast<const float4 *>(bias)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float inv_n = 1.0f / (float)a.ne0;
    float mean = 0.f;
    if (!RMS) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; j++) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
        mean = block_reduce<false>(s, red) * inv_n;
    }
    float s2 = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; j++) {
        if (threadIdx.x + 256 * j < n4) {
            v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
            s2 += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
        }
    }
    const float scale = 1.0f / sqrtf(block_reduce<false>(s2, red) * inv_n + eps);
#pragma unroll
    for (int j = 0; j < VPT; j++) {
        const int i = threadIdx.x + 256 * j;
        if (i >= n4) continue;
        float4 o = make_float4(v[j].x * scale, v[j].y * scale, v[j].z * scale, v[j].w * scale);
        if (gain) { o.x *= g[j].x; o.y *= g[j].y; o.z *= g[j].z; o.w *= g[j].w; }
        if (bias) { o.x += b[j].x; o.y += b[j].y; o.z += b[j].z; o.w += b[j].w; }
        y[i] = o;
    }
}

// ---- SOFT_MAX (+ scale, + mask with ALiBi slope, + causal mask) ------------------------------------------------------------------------------
// one CTA per row: w = x * scale + slope * mask[row % ne1] (and -inf where i > n_past + row % ne1 when the DIAG_MASK_INF in front was
// folded in), y = exp(w - max) / sum.  Three passes over the row, the last one writes: in-place safe.
template <typename MASK_T>
__global__ void __launch_bounds__(256) soft_max_kernel(const float *__restrict__ x, float *__restrict__ y, const MASK_T *__restrict__ mask, int64_t nc,
                                                       int64_t ne1, int64_t ne2, float scale, float max_bias, float m0, float m1, uint32_t n_head_log2,
                                                       int n_past) {
    pdl_enter();
    __shared__ float red[33];
    const int64_t row = blockIdx.x;
    const int64_t i1 = row % ne1;
    const uint32_t h = (uint32_t)((row / ne1) % ne2);
    const float slope = max_bias > 0.0f ? (h < n_head_log2 ? powf(m0, (float)(h + 1)) : powf(m1, (float)(2 * (h - n_head_log2) + 1))) : 1.0f;
    const float *xr = x + row * nc;
    float *yr = y + row * nc;
    const MASK_T *mr = mask ? mask + i1 * nc : nullptr;
    const int64_t lim = n_past >= 0 ? (int64_t)n_past + i1 : nc;     // columns i > lim are masked out
    auto val = [&](int64_t i) -> float {
        if (i > lim) return -INFINITY;
        float w = __fmul_rn(xr[i], scale);            // rounded like the SCALE kernel's result, never contracted into the subtraction below
        if (mr) w += slope * (float)mr[i];
        return w;
    };
    float mx = -INFINITY;
    for (int64_t i = threadIdx.x; i < nc; i += blockDim.x) mx = fmaxf(mx, val(i));
    mx = block_reduce<true>(mx, red);
    float sum = 0.f;
    for (int64_t i = threadIdx.x; i < nc; i += blockDim.x) {
        const float w = val(i);
        sum += w == -INFINITY ? 0.f : expf(w - mx);
    }
    sum = block_reduce<false>(sum, red);
    const float inv = 1.0f / sum;
    for (int64_t i = threadIdx.x; i < nc; i += blockDim.x) {
        const float w = val(i);
        yr[i] = w == -INFINITY ? 0.f : expf(w - mx) * inv;
    }
}

// ---- CPY / DUP / CONT ------------------------------------------------------------------------------------------------------------------
// element t of the flattened source goes to element t of the flattened destination (shapes may differ, ggml_cpy only needs the same
// element count); both sides strided.  Threads run along the DESTINATION's dim 0, so stores coalesce.
template <typename S, typename D> __device__ __forceinline__ D convert(S v);
template <> __device__ __forceinline__ float convert<float, float>(float v) { return v; }
template <> __device__ __forceinline__ __half convert<float, __half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float convert<__half, float>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ __half convert<__half, __half>(__half v) { return v; }
template <> __device__ __forceinline__ uint16_t convert<uint16_t, uint16_t>(uint16_t v) { return v; }
template <> __device__ __forceinline__ uint32_t convert<uint32_t, uint32_t>(uint32_t v) { return v; }

template <typename S, typename D>
__global__ void __launch_bounds__(256) copy_kernel(T4 a, T4 d, int64_t total) {
    pdl_enter();
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = t / d.ne0;
        const int64_t j0 = t - r * d.ne0;
        const int64_t j3 = r / (d.ne2 * d.ne1);
        r -= j3 * d.ne2 * d.ne1;
        const int64_t j2 = r / d.ne1, j1 = r - j2 * d.ne1;
        r = t / a.ne0;
        const int64_t i0 = t - r * a.ne0;
        const int64_t i3 = r / (a.ne2 * a.ne1);
        r -= i3 * a.ne2 * a.ne1;
        const int64_t i2 = r / a.ne1, i1 = r - i2 * a.ne1;
        const S v = *reinterpret_cast<const S *>(a.p + i3 * a.nb3 + i2 * a.nb2 + i1 * a.nb1 + i0 * a.nb0);
        *reinterpret_cast<D *>(d.p + j3 * d.nb3 + j2 * d.nb2 + j1 * d.nb1 + j0 * d.nb0) = convert<S, D>(v);
    }
}
// same shape, dense rows on both sides, 16-byte aligned: 128 bits per thread
__global__ void __launch_bounds__(256) copy_rows16_kernel(T4 a, T4 d, int64_t row_bytes) {
    pdl_enter();
    const int64_t n16 = row_bytes >> 4;
    const int64_t total = n16 * d.ne1 * d.ne2 * d.ne3;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = t / n16;
        const int64_t c = (t - r * n16) << 4;
        const int64_t i3 = r / (d.ne2 * d.ne1);
        r -= i3 * d.ne2 * d.ne1;
        const int64_t i2 = r / d.ne1, i1 = r - i2 * d.ne1;
        *reinterpret_cast<uint4 *>(d.p + i3 * d.nb3 + i2 * d.nb2 + i1 * d.nb1 + c) = *reinterpret_cast<const uint4 *>(a.p + i3 * a.nb3 + i2 * a.nb2 + i1 * a.nb1 + c);
    }
}

// ---- attention of a decode step as ONE kernel: K*Q -> SCALE -> DIAG_MASK_INF -> SOFT_MAX -> V*P -> merged heads ----------------------------
// (examples/gpt-j/main.cpp:490-530, examples/gpt-2/main-backend.cpp:567-610: six graph nodes, four launches otherwise).  One CTA per
// (head, token): scores of the T cached positions into shared memory (a warp per position, lanes over the head dimension), the
// softmax of soft_max_kernel (same rounding of the scale, same three steps), then a warp per output dimension over V's rows, which
// the caches keep contiguous along T.  K [hd][T][H] and V [T][hd][H] are F16 or F32 views with any row / head strides; Q is F32
// [hd][N][H]; the result goes straight to the merged layout dst[hd][H][N] the CPY / CONT behind the PERMUTE would produce.
template <typename KT, typename VT>
__global__ void __launch_bounds__(1024) attention_decode_kernel(T4 q, T4 k, T4 v, T4 d, float scale, int n_past) {
    pdl_enter();
    extern __shared__ float sc[];                      // T scores, then the probabilities
    __shared__ float red[33];
    const int h = blockIdx.x, i = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int hd = (int)q.ne0, T = (int)k.ne1;
    const int lim = n_past >= 0 ? min(T, n_past + i + 1) : T;              // positions t >= lim are masked out
    // 32 warps so that the chains of dependent memory latencies stay short: a decode step has T of a few hundred at most per launch here,
    // and a CTA that walks its positions (or its output rows) eight at a time spends its life waiting for L2
    const float *qr = reinterpret_cast<const float *>(q.p + (int64_t)i * q.nb1 + (int64_t)h * q.nb2);
    float qv[8];                                         // hd <= 256: this lane's slice of q stays in registers
#pragma unroll
    for (int j = 0; j < 8; j++) qv[j] = lane + 32 * j < hd ? qr[lane + 32 * j] : 0.f;
    for (int t = warp; t < lim; t += nw) {
        const KT *kr = reinterpret_cast<const KT *>(k.p + (int64_t)t * k.nb1 + (int64_t)h * k.nb2);
        float s = 0.f;
        if (hd <= 256) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (lane + 32 * j < hd) s += convert<KT, float>(kr[lane + 32 * j]) * qv[j];
        } else {
            for (int e = lane; e < hd; e += 32) s += convert<KT, float>(kr[e]) * qr[e];
        }
        s = warp_sum(s);
        if (lane == 0) sc[t] = __fmul_rn(s, scale);
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int t = threadIdx.x; t < lim; t += blockDim.x) mx = fmaxf(mx, sc[t]);
    mx = block_reduce<true>(mx, red);
    float sum = 0.f;
    for (int t = threadIdx.x; t < lim; t += blockDim.x) {
        const float e = expf(sc[t] - mx);
        sc[t] = e;
        sum += e;
    }
    sum = block_reduce<false>(sum, red);               // (its barriers also publish sc[])
    const float inv = 1.0f / sum;
    float *out = reinterpret_cast<float *>(d.p + (int64_t)h * d.nb1 + (int64_t)i * d.nb2);
    // four output rows per warp at a time: their loads are all in flight before the first multiply
    for (int e0 = warp * 4; e0 < hd; e0 += nw * 4) {
        const VT *vr[4];
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 4; r++) vr[r] = reinterpret_cast<const VT *>(v.p + (int64_t)min(e0 + r, hd - 1) * v.nb1 + (int64_t)h * v.nb2);
        for (int t = lane; t < lim; t += 32) {
            const float p = sc[t] * inv;
            float x[4];
#pragma unroll
            for (int r = 0; r < 4; r++) x[r] = convert<VT, float>(vr[r][t]);
#pragma unroll
            for (int r = 0; r < 4; r++) acc[r] += x[r] * p;
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const float a = warp_sum(acc[r]);
            if (lane == 0 && e0 + r < hd) out[e0 + r] = a;
        }
    }
}

// ---- ROPE (src/ggml.c:13775 f32, :13953 f16; forward only) -----------------------------------------------------------------------------
// One thread per rotated pair.  The angle of pair j is pos * theta_scale^j built by j successive fp32 multiplications, as the CPU
// loop builds it (ggml_rope_cache_init, :13750), so the argument of cosf / sinf is the CPU's bit for bit; YaRN mixing (rope_yarn,
// :13726) and the xPos factor follow the same expressions.  NeoX mode keeps two of the reference's quirks: freq_scale is applied
// twice (:13908 and inside rope_yarn) and the ramp index is the truncation of a value in (-1, 0], i.e. 0.  GLM mode is declined.
struct RopeK {
    int   n_dims, neox;
    float theta_scale, freq_scale, ext_factor, attn_factor, corr0, corr1, xpos_base;
    int   xpos_down;
};
__device__ __forceinline__ void rope_yarn_dev(float theta_extrap, const RopeK &r, int i0, float *c, float *s) {
    const float theta_interp = r.freq_scale * theta_extrap;
    float theta = theta_interp, mscale = r.attn_factor;
    if (r.ext_factor != 0.0f) {
        const float y = ((float)(i0 / 2) - r.corr0) / fmaxf(0.001f, r.corr1 - r.corr0);
        const float ramp_mix = (1.0f - fminf(1.0f, fmaxf(0.0f, y))) * r.ext_factor;
        theta = __fadd_rn(__fmul_rn(theta_interp, 1.0f - ramp_mix), __fmul_rn(theta_extrap, ramp_mix));   // no contraction: an ulp of a large angle is visible in its cosine
        mscale *= 1.0f + 0.1f * logf(1.0f / r.freq_scale);
    }
    *c = cosf(theta) * mscale;
    *s = sinf(theta) * mscale;
}
// TS -> TD: the CPY that stores a rotated k into an F16 cache folds into the store (same rounding, one conversion)
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) rope_kernel(T4 a, T4 d, const int32_t *__restrict__ pos, RopeK r) {
    pdl_enter();
    const int64_t half = a.ne0 >> 1;
    const int64_t total = half * a.ne1 * a.ne2 * a.ne3;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = t / half;
        const int j = (int)(t - row * half);                 // pair index inside the row
        const int64_t i3 = row / (a.ne2 * a.ne1);
        row -= i3 * a.ne2 * a.ne1;
        const int64_t i2 = row / a.ne1, i1 = row - i2 * a.ne1;
        const TS *src = reinterpret_cast<const TS *>(a.p + i3 * a.nb3 + i2 * a.nb2 + i1 * a.nb1);
        TD *dst = reinterpret_cast<TD *>(d.p + i3 * d.nb3 + i2 * d.nb2 + i1 * d.nb1);
        const int p = pos[i2];
        const int ic = 2 * j;
        if (r.neox && ic >= r.n_dims) {                      // beyond the rotated part: plain copy of the pair
            dst[ic] = convert<float, TD>(convert<TS, float>(src[ic]));
            dst[ic + 1] = convert<float, TD>(convert<TS, float>(src[ic + 1]));
            continue;
        }
        float theta = r.neox ? (float)p * r.freq_scale : (float)p;
        for (int q = 0; q < j; ++q) theta *= r.theta_scale;
        float c, s;
        rope_yarn_dev(theta, r, r.neox ? 0 : ic, &c, &s);
        const int i0 = r.neox ? j : ic, i1x = r.neox ? j + r.n_dims / 2 : ic + 1;
        const float x0 = convert<TS, float>(src[i0]), x1 = convert<TS, float>(src[i1x]);
        float zeta = 1.0f;
        if (sizeof(TS) == 4 && !r.neox && r.xpos_base != 0.0f) {
            zeta = powf(((float)ic + 0.4f * (float)a.ne0) / (1.4f * (float)a.ne0), (float)p / r.xpos_base);
            if (r.xpos_down) zeta = 1.0f / zeta;
        }
        dst[i0] = convert<float, TD>(x0 * c * zeta - x1 * s * zeta);
        dst[i1x] = convert<float, TD>(x0 * s * zeta + x1 * c * zeta);
    }
}

// ---- REPEAT (src/ggml.c:10323): dst[i] = src0[i mod src0 shape], 2- and 4-byte elements ------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) repeat_kernel(T4 a, T4 d, int64_t total) {
    pdl_enter();
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = t / d.ne0;
        const int64_t j0 = t - r * d.ne0;
        const int64_t j3 = r / (d.ne2 * d.ne1);
        r -= j3 * d.ne2 * d.ne1;
        const int64_t j2 = r / d.ne1, j1 = r - j2 * d.ne1;
        *reinterpret_cast<T *>(d.p + j3 * d.nb3 + j2 * d.nb2 + j1 * d.nb1 + j0 * d.nb0) =
            *reinterpret_cast<const T *>(a.p + (j3 % a.ne3) * a.nb3 + (j2 % a.ne2) * a.nb2 + (j1 % a.ne1) * a.nb1 + (j0 % a.ne0) * a.nb0);
    }
}

// ---- MUL_MAT with a dense (F32 / F16) src0 -----------------------------------------------------------------------------------------------
// dst[i3][i2][n][m] = sum_k src0[i3 / r3][i2 / r2][m][k] * src1[i3][i2][n][k]   (src/ggml.c:11808; both operands k-contiguous, any row /
// batch strides, so permuted views of the KV cache multiply in place).  CTA = TM x TN tile of one (i2, i3) slice, k in steps of 16
// through shared memory, RM x RN accumulators per thread, fp32 FMA.
template <typename A, int TM, int TN, int RM, int RN>
__global__ void __launch_bounds__((TM / RM) * (TN / RN)) mul_mat_dense_kernel(T4 a, T4 b, T4 d, int64_t r2, int64_t r3) {
    pdl_enter();
    constexpr int KT = 16;
    constexpr int NT = (TM / RM) * (TN / RN);
    __shared__ float As[KT][TM + 1];
    __shared__ float Bs[KT][TN + 1];
    const int64_t i2 = blockIdx.z % d.ne2, i3 = blockIdx.z / d.ne2;
    const int64_t m0 = (int64_t)blockIdx.x * TM, n0 = (int64_t)blockIdx.y * TN;
    const char *ap = a.p + (i3 / r3) * a.nb3 + (i2 / r2) * a.nb2;
    const char *bp = b.p + i3 * b.nb3 + i2 * b.nb2;
    const int tm = threadIdx.x % (TM / RM), tn = threadIdx.x / (TM / RM);
    const int64_t K = a.ne0, M = a.ne1, N = b.ne1;
    float acc[RM][RN];
#pragma unroll
    for (int i = 0; i < RM; i++)
#pragma unroll
        for (int j = 0; j < RN; j++) acc[i][j] = 0.f;
    for (int64_t k0 = 0; k0 < K; k0 += KT) {
        for (int e = threadIdx.x; e < TM * KT; e += NT) {
            const int kk = e % KT, mm = e / KT;
            float v = 0.f;
            if (m0 + mm < M && k0 + kk < K) {
                const A *row = reinterpret_cast<const A *>(ap + (m0 + mm) * a.nb1);
                v = (float)row[k0 + kk];
            }
            As[kk][mm] = v;
        }
        for (int e = threadIdx.x; e < TN * KT; e += NT) {
            const int kk = e % KT, nn = e / KT;
            float v = 0.f;
            if (n0 + nn < N && k0 + kk < K) v = reinterpret_cast<const float *>(bp + (n0 + nn) * b.nb1)[k0 + kk];
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < KT; kk++) {
            float av[RM], bv[RN];
#pragma unroll
            for (int i = 0; i < RM; i++) av[i] = As[kk][tm + i * (TM / RM)];
#pragma unroll
            for (int j = 0; j < RN; j++) bv[j] = Bs[kk][tn + j * (TN / RN)];
#pragma unroll
            for (int i = 0; i < RM; i++)
#pragma unroll
                for (int j = 0; j < RN; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    char *dp = d.p + i3 * d.nb3 + i2 * d.nb2;
#pragma unroll
    for (int j = 0; j < RN; j++) {
        const int64_t n = n0 + tn + j * (TN / RN);
        if (n >= N) continue;
#pragma unroll
        for (int i = 0; i < RM; i++) {
            const int64_t m = m0 + tm + i * (TM / RM);
            if (m < M) reinterpret_cast<float *>(dp + n * d.nb1)[m] = acc[i][j];
        }
    }
}

// decode shapes (N <= 8 columns): one warp per row of src0, lanes along k, every load of the row in flight at once; the row of src0 is
// read once for all N columns.  K*Q of a decode step is m = n_past + 1 rows of 64 floats per head, V*P is 64 rows of n_past + 1.
template <typename A, int NC>
__global__ void __launch_bounds__(256) mul_mat_dense_rows_kernel(T4 a, T4 b, T4 d, int64_t r2, int64_t r3) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= a.ne1) return;
    const int64_t i2 = blockIdx.y % d.ne2, i3 = blockIdx.y / d.ne2;
    const A *row = reinterpret_cast<const A *>(a.p + (i3 / r3) * a.nb3 + (i2 / r2) * a.nb2 + m * a.nb1);
    const char *bp = b.p + i3 * b.nb3 + i2 * b.nb2;
    const int64_t K = a.ne0;
    float acc[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) acc[c] = 0.f;
    for (int64_t k = lane; k < K; k += 32) {
        const float w = (float)row[k];
#pragma unroll
        for (int c = 0; c < NC; c++)
            if (c < b.ne1) acc[c] = fmaf(w, reinterpret_cast<const float *>(bp + c * b.nb1)[k], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < NC; c++) acc[c] = warp_sum(acc[c]);
    if (lane == 0) {
        char *dp = d.p + i3 * d.nb3 + i2 * d.nb2;
#pragma unroll
        for (int c = 0; c < NC; c++)
            if (c < b.ne1) reinterpret_cast<float *>(dp + c * d.nb1)[m] = acc[c];
    }
}

int finish(b200_ctx *ctx, const char *what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        b200_set_error(ctx, "%s launch failed: %s", what, cudaGetErrorString(e));
        return B200_ERR_CUDA;
    }
    ctx->launches++;
    return B200_OK;
}

}  // namespace

#define OPS_ENTER(ctx)                                                       \
    if (!(ctx)) return B200_ERR_INVALID;                                     \
    B200_CUDA_TRY((ctx), b200_use_device((ctx)->device))

extern "C" {

int b200_op_get_rows(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *rows, const b200_tensor *dst) {
    OPS_ENTER(ctx);
    B200_REQUIRE(ctx, src0 && rows && dst, B200_ERR_INVALID);
    B200_REQUIRE(ctx, rows->type == B200_TYPE_I32 && dst->type == B200_TYPE_F32, B200_ERR_UNSUPPORTED);
    B200_REQUIRE(ctx, dst->ne[0] == src0->ne[0] && dst->ne[1] == rows->ne[0] && dst->ne[2] == rows->ne[1] && dst->ne[3] == rows->ne[2], B200_ERR_INVALID);
    B200_REQUIRE(ctx, src0->ne[2] == rows->ne[1] && rows->ne[3] == 1 && dst->nb[0] == 4, B200_ERR_INVALID);
    const int64_t nc = src0->ne[0], nr = nelements(rows);
    if (nc == 0 || nr == 0) return B200_OK;
    const T4 r = view(rows), d = view(dst), a = view(src0);
    if (src0->type == B200_TYPE_Q5_0 || src0->type == B200_TYPE_IQ4_NL) return b200_launch_get_rows_wire(ctx, src0, rows, dst);
    const bool quant = src0->type == B200_TYPE_Q4_0 || src0->type == B200_TYPE_Q8_0;
    if (quant) {
        const int wire = b200_wire_bytes(src0->type);
        B200_REQUIRE(ctx, nc % B200_QK == 0 && src0->nb[0] == wire && src0->nb[1] % wire == 0 && src0->nb[2] % wire == 0 && src0->nb[3] % wire == 0, B200_ERR_UNSUPPORTED);
        B200_REQUIRE(ctx, aligned16(dst) && src0->q_total_blocks > 0, B200_ERR_UNSUPPORTED);
        const uint8_t *qs = static_cast<const uint8_t *>(src0->data);
        const __half *qd = reinterpret_cast<const __half *>(qs + src0->q_total_blocks * b200_qs_bytes(src0->type));
        const int grid = grid_for(nr * (nc / 4), 256, ctx->sm_count);
        if (src0->type == B200_TYPE_Q4_0)
            launch_k(ctx, get_rows_kernel<B200_TYPE_Q4_0>, grid, 256, nullptr, a.nb1, a.nb2, a.nb3, a.ne1, qs, qd, src0->q_block_off, r.p, r.nb0, r.nb1, r.nb2,
                                                                         r.ne0, r.ne1, reinterpret_cast<float *>(d.p), d.nb1, d.nb2, d.nb3, nc, nr, nullptr);
        else
            launch_k(ctx, get_rows_kernel<B200_TYPE_Q8_0>, grid, 256, nullptr, a.nb1, a.nb2, a.nb3, a.ne1, qs, qd, src0->q_block_off, r.p, r.nb0, r.nb1, r.nb2,
                                                                         r.ne0, r.ne1, reinterpret_cast<float *>(d.p), d.nb1, d.nb2, d.nb3, nc, nr, nullptr);
        return finish(ctx, "get_rows");
    }
    B200_REQUIRE(ctx, src0->type == B200_TYPE_F32 || src0->type == B200_TYPE_F16, B200_ERR_UNSUPPORTED);
    B200_REQUIRE(ctx, src0->nb[0] == elt_size(src0->type), B200_ERR_UNSUPPORTED);
    const bool vec = nc % 4 == 0 && aligned16(dst) && ((uintptr_t)src0->data & 15) == 0;
    if (vec && src0->type == B200_TYPE_F32 && aligned16(src0)) {
        launch_k(ctx, get_rows_kernel<B200_TYPE_F32>, grid_for(nr * (nc / 4), 256, ctx->sm_count), 256, 
            a.p, a.nb1, a.nb2, a.nb3, a.ne1, nullptr, nullptr, 0, r.p, r.nb0, r.nb1, r.nb2, r.ne0, r.ne1, reinterpret_cast<float *>(d.p), d.nb1, d.nb2, d.nb3, nc, nr, nullptr);
    } else if (vec && src0->type == B200_TYPE_F16 && src0->nb[1] % 8 == 0 && src0->nb[2] % 8 == 0 && src0->nb[3] % 8 == 0) {
        launch_k(ctx, get_rows_kernel<B200_TYPE_F16>, grid_for(nr * (nc / 4), 256, ctx->sm_count), 256, 
            a.p, a.nb1, a.nb2, a.nb3, a.ne1, nullptr, nullptr, 0, r.p, r.nb0, r.nb1, r.nb2, r.ne0, r.ne1, reinterpret_cast<float *>(d.p), d.nb1, d.nb2, d.nb3, nc, nr, nullptr);
    } else if (src0->type == B200_TYPE_F32) {
        launch_k(ctx, get_rows_scalar_kernel<B200_TYPE_F32>, grid_for(nr * nc, 256, ctx->sm_count), 256, a.p, a.nb1, a.nb2, a.nb3, a.ne1, r.p, r.nb0, r.nb1, r.nb2,
                                                                                                          r.ne0, r.ne1, d.p, d.nb1, d.nb2, d.nb3, nc, nr, nullptr);
    } else {
        launch_k(ctx, get_rows_scalar_kernel<B200_TYPE_F16>, grid_for(nr * nc, 256, ctx->sm_count), 256, a.p, a.nb1, a.nb2, a.nb3, a.ne1, r.p, r.nb0, r.nb1, r.nb2,
                                                                                                          r.ne0, r.ne1, d.p, d.nb1, d.nb2, d.nb3, nc, nr, nullptr);
    }
    return finish(ctx, "get_rows");
}

int b200_op_binary(b200_ctx *ctx, int op, const b200_tensor *src0, const b200_tensor *src1, const b200_tensor *dst) {
    OPS_ENTER(ctx);
    B200_REQUIRE(ctx, src0 && src1 && dst, B200_ERR_INVALID);
    B200_REQUIRE(ctx, op == B200_OP_ADD || op == B200_OP_MUL || op == B200_OP_DIV, B200_ERR_UNSUPPORTED);
    B200_REQUIRE(ctx, src0->type == B200_TYPE_F32 && src1->type == B200_TYPE_F32 && dst->type == B200_TYPE_F32, B200_ERR_UNSUPPORTED);
    B200_REQUIRE(ctx, same_shape(src0, dst), B200_ERR_INVALID);
    for (int i = 0; i < 4; i++) B200_REQUIRE(ctx, src1->ne[i] > 0 && dst->ne[i] % src1->ne[i] == 0, B200_ERR_INVALID);   // ggml_can_repeat(src1, src0)
    const int64_t n = nelements(dst);
    if (n == 0) return B200_OK;
    const T4 a = view(src0), b = view(src1), d = view(dst);
    const bool rows = src1->ne[0] == dst->ne[0] && dst->ne[0] % 4 == 0 && src0->nb[0] == 4 && src1->nb[0] == 4 && dst->nb[0] == 4 && aligned16(src0) &&
                      aligned16(src1) && aligned16(dst);
    const int grid = grid_for(rows ? n / 4 : n, 256, ctx->sm_count);
#define B200_BIN(OP)                                                                          \
    if (rows) launch_k(ctx, binary_rows_kernel<OP>, grid, 256, a, b, d);                 \
    else launch_k(ctx, binary_generic_kernel<OP>, grid, 256, a, b, d)
    if (op == B200_OP_ADD) { B200_BIN(B200_OP_ADD); }
    else if (op == B200_OP_MUL) { B200_BIN(B200_OP_MUL); }
    else { B200_BIN(B200_OP_DIV); }
#undef B200_BIN
    return finish(ctx, "binary");
}

static int elementwise(b200_ctx *ctx, int mode, const b200_tensor *src0, const b200_tensor *dst, int op, float s) {
    B200_REQUIRE(ctx, src0 && dst, B200_ERR_INVALID);
    B200_REQUIRE(ctx, src0->type == B200_TYPE_F32 && dst->type == B200_TYPE_F32, B200_ERR_UNSUPPORTED);
    B200_REQUIRE(ctx, same_shape(src0, dst), B200_ERR_INVALID);
    B200_REQUIRE(ctx, contiguous(src0) && contiguous(dst), B200_ERR_UNSUPPORTED);
    const int64_t n = nelements(dst);
    if (n == 0) return B200_OK;
    const float *x = static_cast<const float *>(src0->data);
    float *y = static_cast<float *>(dst->data);
    const int grid = grid_for(n, 256, ctx->sm_count);
    if (mode == 0) launch_k(ctx, elementwise_kernel<0>, grid, 256, x, y, n, op, s, src0->ne[0], src0->ne[1]);
    else if (mode == 1) launch_k(ctx, elementwise_kernel<1>, grid, 256, x, y, n, op, s, src0->ne[0], src0->ne[1]);
    else launch_k(ctx, elementwise_kernel<2>, grid, 256, x, y, n, op, s, src0->ne[0], src0->ne[1]);
    return finish(ctx, "elementwise");
}

int b200_op_unary(b200_ctx *ctx, int op, const b200_tensor *src0, const b200_tensor *dst) {
    OPS_ENTER(ctx);
    B200_REQUIRE(ctx, op >= 0 && op < B200_UNARY_COUNT, B200_ERR_UNSUPPORTED);
    return elementwise(ctx, 0, src0, dst, op, 0.f);
}
int b200_op_scale(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *dst, float s) {
    OPS_ENTER(ctx);
    return elementwise(ctx, 1, src0, dst, 0, s);
}
int b200_op_diag_mask_inf(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *dst, int n_past) {
    OPS_ENTER(ctx);
    B200_REQUIRE(ctx, n_past >= 0, B200_ERR_INVALID);
    return elementwise(ctx, 2, src0, dst, n_past, 0.f);
}

int b200_op_norm(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *gain, const b200_tensor *bias, const b200_tensor *dst, float eps, int rms) {
    OPS_ENTER(ctx);
    B200_REQUIRE(ctx, src0 && dst, B200_ERR_INVALID);
    B200_REQUIRE(ctx, src0->type == B200_TYPE_F32 && dst->type == B200_TYPE_F32, B200_ERR_UNSUPPORTED);
    B200_REQUIRE(ctx, same_shape(src0, dst) && src0->nb[0] == 4 && dst->nb[0] == 4, B200_ERR_INVALID);
    B200_REQUIRE(ctx, rms ? eps >= 0.0f : eps > 0.0f, B200_ERR_INVALID);
    for (const b200_tensor *v : {gain, bias})
        if (v) B200_REQUIRE(ctx, v->type == B200_TYPE_F32 && v->ne[0] == src0->ne[0] && v->ne[1] * v->ne[2] * v->ne[3] == 1 && v->nb[0] == 4, B200_ERR_UNSUPPORTED);
    const int64_t nr = nrows(src0);
    if (nr == 0 || src0->ne[0] == 0) return B200_OK;
    B200_REQUIRE(ctx, nr < (1ll << 31), B200_ERR_UNSUPPORTED);
    const int threads = src0->ne[0] <= 64 ? 32 : (src0->ne[0] <= 256 ? 64 : (src0->ne[0] <= 1024 ? 128 : 256));
    const float *g = gain ? static_cast<const float *>(gain->data) : nullptr, *b = bias ? static_cast<const float *>(bias->data) : nullptr;
    const int64_t ne0 = src0->ne[0];
    if (ne0 % 4 == 0 && ne0 <= 1024 && aligned16(src0) && aligned16(dst) && ((uintptr_t)g & 15) == 0 && ((uintptr_t)b & 15) == 0) {
        const unsigned blocks = (unsigned)((nr + 3) / 4);
        const T4 a = view(src0), d = view(dst);
#define B200_NORM_WARP(R, V) launch_k(ctx, norm_warp_kernel<R, V>, blocks, 128, a, d, g, b, eps, nr)
        if (ne0 <= 256) { if (rms) B200_NORM_WARP(true, 2); else B200_NORM_WARP(false, 2); }
        else { if (rms) B200_NORM_WARP(true, 8); else B200_NORM_WARP(false, 8); }
#undef B200_NORM_WARP
        return finish(ctx, "norm");
    }
    if (ne0 % 4 == 0 && ne0 <= 8192 && aligned16(src0) && aligned16(dst) && ((uintptr_t)g & 15) == 0 && ((uintptr_t)b & 15) == 0) {
        const T4 a = view(src0), d = view(dst);
#define B200_NORM_BLOCK(R, V) launch_k(ctx, norm_block_kernel<R, V>, (unsigned)nr, 256, a, d, g, b, eps)
        if (ne0 <= 2048) { if (rms) B200_NORM_BLOCK(true, 2); else B200_NORM_BLOCK(false, 2); }
        else if (ne0 <= 4096) { if (rms) B200_NORM_BLOCK(true, 4); else B200_NORM_BLOCK(false, 4); }
        else { if (rms) B200_NORM_BLOCK(true, 8); else B200_NORM_BLOCK(false, 8); }
#undef B200_NORM_BLOCK
        return finish(ctx, "norm");
    }
    if (rms) launch_k(ctx, norm_kernel<true>, (unsigned)nr, threads, view(src0), view(dst), g, b, eps);
    else launch_k(ctx, norm_kernel<false>, (unsigned)nr, threads, view(src0), view(dst), g, b, eps);
    return finish(ctx, "norm");
}

int b200_op_soft_max(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *mask, const b200_tensor *dst, float scale, float max_bias, int n_past) {
    OPS_ENTER(ctx);
    B200_REQUIRE(ctx, src0 && dst, B200_ERR_INVALID);
    B200_REQUIRE(ctx, src0->type == B200_TYPE_F32 && dst->type == B200_TYPE_F32, B200_ERR_UNSUPPORTED);
    B200_REQUIRE(ctx, same_shape(src0, dst), B200_ERR_INVALID);
    B200_REQUIRE(ctx, contiguous(src0) && contiguous(dst), B200_ERR_UNSUPPORTED);
    if (mask) {
        B200_REQUIRE(ctx, (mask->type == B200_TYPE_F32 || mask->type == B200_TYPE_F16) && contiguous(mask), B200_ERR_UNSUPPORTED);
        B200_REQUIRE(ctx, mask->ne[0] == src0->ne[0] && mask->ne[1] >= src0->ne[1], B200_ERR_INVALID);
    }
    const int64_t nr = nrows(src0), nc = src0->ne[0];
    if (nr == 0 || nc == 0) return B200_OK;
    B200_REQUIRE(ctx, nr < (1ll << 31), B200_ERR_UNSUPPORTED);
    const uint32_t n_head = (uint32_t)src0->ne[2];
    const uint32_t n_head_log2 = 1u << (uint32_t)floor(log2((double)n_head));
    const float m0 = powf(2.0f, -(max_bias) / n_head_log2), m1 = powf(2.0f, -(max_bias / 2.0f) / n_head_log2);      // src/ggml.c:13426-13430
    const int threads = nc <= 32 ? 32 : (nc <= 128 ? 64 : (nc <= 1024 ? 128 : 256));
    const float *x = static_cast<const float *>(src0->data);
    float *y = static_cast<float *>(dst->data);
    if (mask && mask->type == B200_TYPE_F16)
        launch_k(ctx, soft_max_kernel<__half>, (unsigned)nr, threads, x, y, static_cast<const __half *>(mask->data), nc, src0->ne[1], src0->ne[2], scale, max_bias,
                                                                          m0, m1, n_head_log2, n_past);
    else
        launch_k(ctx, soft_max_kernel<float>, (unsigned)nr, threads, x, y, mask ? static_cast<const float *>(mask->data) : nullptr, nc, src0->ne[1], src0->ne[2],
                                                                         scale, max_bias, m0, m1, n_head_log2, n_past);
    return finish(ctx, "soft_max");
}

int b200_op_copy(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *dst) {
    OPS_ENTER(ctx);
    B200_REQUIRE(ctx, src0 && dst, B200_ERR_INVALID);
    const int64_t n = nelements(src0);
    B200_REQUIRE(ctx, n == nelements(dst), B200_ERR_INVALID);
    const int es = elt_size(src0->type), ed = elt_size(dst->type);
    B200_REQUIRE(ctx, es && ed, B200_ERR_UNSUPPORTED);
    const bool fp_s = src0->type == B200_TYPE_F32 || src0->type == B200_TYPE_F16, fp_d = dst->type == B200_TYPE_F32 || dst->type == B200_TYPE_F16;
    B200_REQUIRE(ctx, src0->type == dst->type || (fp_s && fp_d), B200_ERR_UNSUPPORTED);
    if (n == 0) return B200_OK;
    if (src0->data == dst->data && src0->type == dst->type && contiguous(src0) && contiguous(dst)) return B200_OK;
    const T4 a = view(src0), d = view(dst);
    if (src0->type == dst->type) {
        if (contiguous(src0) && contiguous(dst)) {
            B200_CUDA_TRY(ctx, cudaMemcpyAsync(dst->data, src0->data, (size_t)n * es, cudaMemcpyDeviceToDevice, ctx->stream));
            ctx->launches++;
            return B200_OK;
        }
        const int64_t row_bytes = src0->ne[0] * es;
        if (same_shape(src0, dst) && src0->nb[0] == es && dst->nb[0] == es && row_bytes % 16 == 0 && aligned16(src0) && aligned16(dst)) {
            launch_k(ctx, copy_rows16_kernel, grid_for(n * es / 16, 256, ctx->sm_count), 256, a, d, row_bytes);
            return finish(ctx, "copy");
        }
    }
    const int grid = grid_for(n, 256, ctx->sm_count);
    if (src0->type == dst->type && es == 4) launch_k(ctx, copy_kernel<uint32_t, uint32_t>, grid, 256, a, d, n);
    else if (src0->type == dst->type) launch_k(ctx, copy_kernel<uint16_t, uint16_t>, grid, 256, a, d, n);
    else if (src0->type == B200_TYPE_F32) launch_k(ctx, copy_kernel<float, __half>, grid, 256, a, d, n);
    else launch_k(ctx, copy_kernel<__half, float>, grid, 256, a, d, n);
    return finish(ctx, "copy");
}

int b200_op_rope(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *pos, const b200_tensor *dst, const b200_rope_params *rp) {
    OPS_ENTER(ctx);
    B200_REQUIRE(ctx, src0 && pos && dst && rp, B200_ERR_INVALID);
    B200_REQUIRE(ctx, (src0->type == B200_TYPE_F32 || src0->type == B200_TYPE_F16) && pos->type == B200_TYPE_I32 &&
                          (dst->type == src0->type || (src0->type == B200_TYPE_F32 && dst->type == B200_TYPE_F16)), B200_ERR_UNSUPPORTED);
    B200_REQUIRE(ctx, (rp->mode & 4) == 0, B200_ERR_UNSUPPORTED);                                  // GLM layout: not built
    B200_REQUIRE(ctx, same_shape(src0, dst) && src0->nb[0] == elt_size(src0->type) && dst->nb[0] == elt_size(dst->type), B200_ERR_INVALID);
    B200_REQUIRE(ctx, rp->n_dims > 0 && rp->n_dims % 2 == 0 && rp->n_dims <= src0->ne[0] && src0->ne[0] % 2 == 0, B200_ERR_INVALID);
    B200_REQUIRE(ctx, pos->ne[0] >= src0->ne[2] && pos->nb[0] == 4, B200_ERR_INVALID);
    const int64_t n = nelements(src0);
    if (n == 0) return B200_OK;
    RopeK r;
    r.n_dims = rp->n_dims;
    r.neox = (rp->mode & 2) != 0;
    r.theta_scale = powf(rp->freq_base, -2.0f / rp->n_dims);
    r.freq_scale = rp->freq_scale; r.ext_factor = rp->ext_factor; r.attn_factor = rp->attn_factor;
    // ggml_rope_yarn_corr_dims (src/ggml.c:13765)
    const float two_pi = 2.0f * 3.14159265358979323846f;
    const float lo = floorf(rp->n_dims * logf(rp->n_orig_ctx / (rp->beta_fast * two_pi)) / (2 * logf(rp->freq_base)));
    const float hi = ceilf(rp->n_dims * logf(rp->n_orig_ctx / (rp->beta_slow * two_pi)) / (2 * logf(rp->freq_base)));
    r.corr0 = lo > 0 ? lo : 0;
    r.corr1 = hi < rp->n_dims - 1 ? hi : (float)(rp->n_dims - 1);
    r.xpos_base = rp->xpos_base; r.xpos_down = rp->xpos_down;
    const T4 a = view(src0), d = view(dst);
    const int grid = grid_for(n / 2, 256, ctx->sm_count);
    if (src0->type == B200_TYPE_F32 && dst->type == B200_TYPE_F16) launch_k(ctx, rope_kernel<float, __half>, grid, 256, a, d, static_cast<const int32_t *>(pos->data), r);
    else if (src0->type == B200_TYPE_F32) launch_k(ctx, rope_kernel<float, float>, grid, 256, a, d, static_cast<const int32_t *>(pos->data), r);
    else launch_k(ctx, rope_kernel<__half, __half>, grid, 256, a, d, static_cast<const int32_t *>(pos->data), r);
    return finish(ctx, "rope");
}

int b200_op_repeat(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *dst) {
    OPS_ENTER(ctx);
    B200_REQUIRE(ctx, src0 && dst, B200_ERR_INVALID);
    const int es = elt_size(src0->type);
    B200_REQUIRE(ctx, es && src0->type == dst->type, B200_ERR_UNSUPPORTED);
    for (int i = 0; i < 4; ++i) B200_REQUIRE(ctx, src0->ne[i] > 0 && dst->ne[i] % src0->ne[i] == 0, B200_ERR_INVALID);
    const int64_t n = nelements(dst);
    if (n == 0) return B200_OK;
    const T4 a = view(src0), d = view(dst);
    const int grid = grid_for(n, 256, ctx->sm_count);
    if (es == 4) launch_k(ctx, repeat_kernel<uint32_t>, grid, 256, a, d, n);
    else launch_k(ctx, repeat_kernel<uint16_t>, grid, 256, a, d, n);
    return finish(ctx, "repeat");
}

int b200_op_attention_decode(b200_ctx *ctx, const b200_tensor *q, const b200_tensor *k, const b200_tensor *v, const b200_tensor *dst, float scale, int n_past) {
    OPS_ENTER(ctx);
    B200_REQUIRE(ctx, q && k && v && dst, B200_ERR_INVALID);
    B200_REQUIRE(ctx, q->type == B200_TYPE_F32 && dst->type == B200_TYPE_F32 && (k->type == B200_TYPE_F32 || k->type == B200_TYPE_F16) &&
                          (v->type == B200_TYPE_F32 || v->type == B200_TYPE_F16), B200_ERR_UNSUPPORTED);
    const int64_t hd = q->ne[0], N = q->ne[1], H = q->ne[2], T = k->ne[1];
    B200_REQUIRE(ctx, q->ne[3] == 1 && k->ne[3] == 1 && v->ne[3] == 1 && dst->ne[3] == 1, B200_ERR_UNSUPPORTED);
    B200_REQUIRE(ctx, k->ne[0] == hd && k->ne[2] == H && v->ne[0] == T && v->ne[1] == hd && v->ne[2] == H, B200_ERR_INVALID);
    B200_REQUIRE(ctx, dst->ne[0] == hd && dst->ne[1] == H && dst->ne[2] == N, B200_ERR_INVALID);
    B200_REQUIRE(ctx, q->nb[0] == 4 && dst->nb[0] == 4 && k->nb[0] == elt_size(k->type) && v->nb[0] == elt_size(v->type), B200_ERR_INVALID);
    if (N > 8 || T > 1024 || H > 65535 || hd > 1024) {
        b200_set_error(ctx, "b200_op_attention_decode: a decode-sized problem only (N <= 8, T <= 1024)");
        return B200_ERR_UNSUPPORTED;
    }
    if (hd == 0 || N == 0 || H == 0 || T == 0) return B200_OK;
    const T4 a = view(q), b = view(k), c = view(v), d = view(dst);
    const dim3 grid((unsigned)H, (unsigned)N);
    const size_t smem = (size_t)T * sizeof(float);
    const bool kh = k->type == B200_TYPE_F16, vh = v->type == B200_TYPE_F16;
    const int threads = hd >= 128 ? 1024 : (hd >= 64 ? 512 : 256);       // about one warp per 4 output rows
    if (kh && vh) launch_ks(ctx, attention_decode_kernel<__half, __half>, grid, threads, DynSmem{smem}, a, b, c, d, scale, n_past);
    else if (kh) launch_ks(ctx, attention_decode_kernel<__half, float>, grid, threads, DynSmem{smem}, a, b, c, d, scale, n_past);
    else if (vh) launch_ks(ctx, attention_decode_kernel<float, __half>, grid, threads, DynSmem{smem}, a, b, c, d, scale, n_past);
    else launch_ks(ctx, attention_decode_kernel<float, float>, grid, threads, DynSmem{smem}, a, b, c, d, scale, n_past);
    return finish(ctx, "attention_decode");
}

int b200_op_mul_mat_dense(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *src1, const b200_tensor *dst) {
    OPS_ENTER(ctx);
    B200_REQUIRE(ctx, src0 && src1 && dst, B200_ERR_INVALID);
    B200_REQUIRE(ctx, (src0->type == B200_TYPE_F32 || src0->type == B200_TYPE_F16) && src1->type == B200_TYPE_F32 && dst->type == B200_TYPE_F32, B200_ERR_UNSUPPORTED);
    B200_REQUIRE(ctx, src0->ne[0] == src1->ne[0] && dst->ne[0] == src0->ne[1] && dst->ne[1] == src1->ne[1] && dst->ne[2] == src1->ne[2] && dst->ne[3] == src1->ne[3],
                 B200_ERR_INVALID);                                                                                   // ggml_can_mul_mat + the dst shape
    B200_REQUIRE(ctx, src0->ne[2] > 0 && src0->ne[3] > 0 && src1->ne[2] % src0->ne[2] == 0 && src1->ne[3] % src0->ne[3] == 0, B200_ERR_INVALID);
    B200_REQUIRE(ctx, src0->nb[0] == elt_size(src0->type) && src1->nb[0] == 4 && dst->nb[0] == 4, B200_ERR_UNSUPPORTED);
    if (nelements(dst) == 0) return B200_OK;
    const int64_t M = src0->ne[1], N = src1->ne[1], batch = dst->ne[2] * dst->ne[3];
    B200_REQUIRE(ctx, batch <= 65535, B200_ERR_UNSUPPORTED);
    const T4 a = view(src0), b = view(src1), d = view(dst);
    const int64_t r2 = src1->ne[2] / src0->ne[2], r3 = src1->ne[3] / src0->ne[3];
    if (src0->ne[0] == 0) {                      // empty contraction: zeros
        for (int64_t i3 = 0; i3 < dst->ne[3]; i3++)
            for (int64_t i2 = 0; i2 < dst->ne[2]; i2++)
                for (int64_t i1 = 0; i1 < dst->ne[1]; i1++)
                    B200_CUDA_TRY(ctx, cudaMemsetAsync(d.p + i3 * d.nb3 + i2 * d.nb2 + i1 * d.nb1, 0, (size_t)M * 4, ctx->stream));
        return B200_OK;
    }
    if (N <= 8) {
        const dim3 grid((unsigned)((M + 7) / 8), (unsigned)batch, 1);
#define B200_ROWS(NC)                                                                                                           \
    if (src0->type == B200_TYPE_F32) launch_k(ctx, mul_mat_dense_rows_kernel<float, NC>, grid, 256, a, b, d, r2, r3);     \
    else launch_k(ctx, mul_mat_dense_rows_kernel<__half, NC>, grid, 256, a, b, d, r2, r3)
        if (N == 1) { B200_ROWS(1); } else if (N <= 4) { B200_ROWS(4); } else { B200_ROWS(8); }
#undef B200_ROWS
    } else {
        const dim3 grid((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64), (unsigned)batch);
        if (src0->type == B200_TYPE_F32) launch_k(ctx, mul_mat_dense_kernel<float, 64, 64, 4, 4>, grid, 256, a, b, d, r2, r3);
        else launch_k(ctx, mul_mat_dense_kernel<__half, 64, 64, 4, 4>, grid, 256, a, b, d, r2, r3);
    }
    return finish(ctx, "mul_mat_dense");
}

}  // extern "C"
