// b200_wire_formats.cu -- the sibling 32-element block formats that share the Q8_0 activation path: Q5_0 and IQ4_NL.
//
// ggml_compute_forward_mul_mat picks, per src0 type, a vec_dot against Q8_0-quantized activations (type_traits: src/ggml.c:673-684 for Q5_0,
// :867-878 for IQ4_NL): ggml_vec_dot_q5_0_q8_0 and ggml_vec_dot_iq4_nl_q8_0 (src/ggml-quants.c:11733).  Both are "decode the block's 32
// integers, exact int32 dot with the activation block, times d_w * d_x, accumulate in fp32" -- the arithmetic of the Q4_0 / Q8_0 path with
// another decoder:
//   Q5_0   {fp16 d; u32 qh; u8 qs[16]} (22 B): element j < 16 = (low nibble of qs[j] | bit j of qh << 4) - 16, element j + 16 = (high nibble |
//          bit j + 16 << 4) - 16                                                     (dequantize_row_q5_0, src/ggml-quants.c:1021-1045)
//   IQ4_NL {fp16 d; u8 qs[16]} (18 B): element j / j + 16 = kvalues_iq4nl[low / high nibble of qs[j]]   (src/ggml-quants.c:3321-3339)
// These tensors stay in WIRE format on the device (no repack: set_tensor / get_tensor are plain copies); this is the correctness path of the
// widening step, not a bandwidth-tuned one -- one CTA quantizes an activation column into shared memory with the bit-exact
// quantize_row_q8_0 arithmetic of b200_quantize.cu, then its warps take rows, lanes take blocks.
#include "b200_internal.cuh"

namespace {

__constant__ int8_t kIq4nl[16] = {-127, -104, -83, -65, -49, -35, -22, -10, 1, 13, 25, 38, 53, 69, 89, 113};

template <int TYPE> struct Wire;
template <> struct Wire<B200_TYPE_Q5_0> { static constexpr int kBytes = 22; };
template <> struct Wire<B200_TYPE_IQ4_NL> { static constexpr int kBytes = 18; };

__device__ __forceinline__ float half_bits_to_float(const uint8_t *p) {
    const unsigned short h = (unsigned short)p[0] | ((unsigned short)p[1] << 8);
    return __half2float(__ushort_as_half(h));
}

// the 32 integers of one block -> w[32] (element order of the format)
template <int TYPE> __device__ __forceinline__ void decode_block(const uint8_t *blk, int (&w)[32]) {
    if (TYPE == B200_TYPE_Q5_0) {
        const uint32_t qh = (uint32_t)blk[2] | ((uint32_t)blk[3] << 8) | ((uint32_t)blk[4] << 16) | ((uint32_t)blk[5] << 24);
        const uint8_t *qs = blk + 6;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            w[j] = (int)((qs[j] & 0x0F) | (((qh >> j) & 1u) << 4)) - 16;
            w[j + 16] = (int)((qs[j] >> 4) | (((qh >> (j + 16)) & 1u) << 4)) - 16;
        }
    } else {
        const uint8_t *qs = blk + 2;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            w[j] = kIq4nl[qs[j] & 0x0F];
            w[j + 16] = kIq4nl[qs[j] >> 4];
        }
    }
}

// CTA = one activation column x a run of rows (of one (i2, i3) slice).  smem: int8 q[k], float d[k / 32]
template <int TYPE>
__global__ void __launch_bounds__(256) gemv_wire_kernel(const uint8_t *__restrict__ w, int64_t k, int64_t m, int64_t ne02, int64_t ne03, const char *__restrict__ x,
                                                        int64_t n, int64_t ne12, size_t nb11, size_t nb12, size_t nb13, float *__restrict__ dst, int rows_per_cta) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int64_t nb = k >> 5;
    int8_t *q = reinterpret_cast<int8_t *>(smem);
    float *dx = reinterpret_cast<float *>(smem + ((k + 15) & ~(int64_t)15));
    const int64_t col = blockIdx.y, batch = blockIdx.z, i12 = batch % ne12, i13 = batch / ne12;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // quantize_row_q8_0 of this column (src/ggml-quants.c:535-618: amax, d = amax / 127 -> fp16, id = 127 / amax, round half to even)
    const float *xc = reinterpret_cast<const float *>(x + col * nb11 + i12 * nb12 + i13 * nb13);
    for (int64_t b = warp; b < nb; b += 8) {
        const float v = xc[b * 32 + lane];
        float amax = fabsf(v);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        const float id = amax != 0.0f ? __fdiv_rn(127.f, amax) : 0.0f;
        q[b * 32 + lane] = (int8_t)__float2int_rn(__fmul_rn(v, id));
        if (lane == 0) dx[b] = __half2float(__float2half_rn(__fdiv_rn(amax, 127.f)));
    }
    __syncthreads();
    const int64_t r2 = ne12 / ne02, r3 = (gridDim.z / ne12) / ne03;
    const uint8_t *wb = w + (((i13 / r3) * ne02 + (i12 / r2)) * m) * nb * Wire<TYPE>::kBytes;
    const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
    for (int64_t row = row0 + warp; row < min(row0 + (int64_t)rows_per_cta, m); row += 8) {
        float acc = 0.0f;
        for (int64_t b = lane; b < nb; b += 32) {
            const uint8_t *blk = wb + (row * nb + b) * Wire<TYPE>::kBytes;
            int wv[32];
            decode_block<TYPE>(blk, wv);
            int sumi = 0;
#pragma unroll
            for (int j = 0; j < 32; j++) sumi += wv[j] * (int)q[b * 32 + j];
            acc += (half_bits_to_float(blk) * dx[b]) * (float)sumi;          // src/ggml-quants.c: sumf += (d_x * d_w) * sumi
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) dst[(batch * n + col) * m + row] = acc;
    }
}

// GET_ROWS: one thread per output element
template <int TYPE>
__global__ void __launch_bounds__(256) get_rows_wire_kernel(const uint8_t *__restrict__ src0, int64_t nb01, int64_t nb02, int64_t nb03, int64_t ne01,
                                                            const char *__restrict__ rows, int64_t nb10, int64_t nb11, int64_t nb12, int64_t ne10, int64_t ne11,
                                                            char *__restrict__ dst, int64_t nb1, int64_t nb2, int64_t nb3, int64_t nc, int64_t nr) {
    const int64_t total = nr * nc;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / nc, c = t - r * nc;
        const int64_t i12 = r / (ne11 * ne10), i11 = (r - i12 * ne11 * ne10) / ne10, i10 = r - i12 * ne11 * ne10 - i11 * ne10;
        const int64_t row = *reinterpret_cast<const int32_t *>(rows + i10 * nb10 + i11 * nb11 + i12 * nb12);
        float v = 0.0f;
        if (row >= 0 && row < ne01) {
            const uint8_t *blk = src0 + row * nb01 + i11 * nb02 + i12 * nb03 + (c >> 5) * Wire<TYPE>::kBytes;
            const int j = (int)(c & 31);
            int wq;
            if (TYPE == B200_TYPE_Q5_0) {
                const uint32_t qh = (uint32_t)blk[2] | ((uint32_t)blk[3] << 8) | ((uint32_t)blk[4] << 16) | ((uint32_t)blk[5] << 24);
                const uint8_t qb = blk[6 + (j & 15)];
                wq = (int)((j < 16 ? (qb & 0x0F) : (qb >> 4)) | (((qh >> j) & 1u) << 4)) - 16;
            } else {
                const uint8_t qb = blk[2 + (j & 15)];
                wq = kIq4nl[j < 16 ? (qb & 0x0F) : (qb >> 4)];
            }
            v = (float)wq * half_bits_to_float(blk);
        }
        *reinterpret_cast<float *>(dst + i10 * nb1 + i11 * nb2 + i12 * nb3 + c * 4) = v;
    }
}

}  // namespace

// dst [ne13][ne12][n][m] = src0 (wire-format blocks, [ne03][ne02][m][k / 32]) x src1; same shape rules as b200_mul_mat
int b200_launch_gemv_wire(b200_ctx *ctx, const b200_mul_mat_args *a) {
    const int64_t k = a->ne00, m = a->ne01, n = a->ne11, batch = a->ne12 * a->ne13;
    B200_REQUIRE(ctx, a->type == B200_TYPE_Q5_0 || a->type == B200_TYPE_IQ4_NL, B200_ERR_UNSUPPORTED);
    B200_REQUIRE(ctx, k > 0 && k % B200_QK == 0 && m > 0 && n > 0 && batch > 0, B200_ERR_INVALID);
    B200_REQUIRE(ctx, a->ne12 % a->ne02 == 0 && a->ne13 % a->ne03 == 0, B200_ERR_INVALID);
    B200_REQUIRE(ctx, n <= 65535 && batch <= 65535 && k <= 131072, B200_ERR_UNSUPPORTED);
    const size_t smem = (size_t)((k + 15) & ~(int64_t)15) + (size_t)(k / 32) * 4;
    const int wire = a->type == B200_TYPE_Q5_0 ? 22 : 18;
    const uint8_t *w = (const uint8_t *)a->src0_dev + a->src0_block_off * wire;
    // enough CTAs per column to fill the device, at least 8 rows (one per warp) each
    int64_t ctas = (int64_t)ctx->sm_count * 4 / (n * batch > 0 ? n * batch : 1);
    if (ctas < 1) ctas = 1;
    int64_t rows_per_cta = (m + ctas - 1) / ctas;
    if (rows_per_cta < 8) rows_per_cta = 8;
    const dim3 grid((unsigned)((m + rows_per_cta - 1) / rows_per_cta), (unsigned)n, (unsigned)batch);
    if (a->type == B200_TYPE_Q5_0) {
        B200_SMEM_LIMIT_ONCE(ctx, gemv_wire_kernel<B200_TYPE_Q5_0>, 200 * 1024);
        gemv_wire_kernel<B200_TYPE_Q5_0><<<grid, 256, smem, ctx->stream>>>(w, k, m, a->ne02, a->ne03, (const char *)a->src1_dev, n, a->ne12, a->nb11, a->nb12, a->nb13,
                                                                          a->dst_dev, (int)rows_per_cta);
    } else {
        B200_SMEM_LIMIT_ONCE(ctx, gemv_wire_kernel<B200_TYPE_IQ4_NL>, 200 * 1024);
        gemv_wire_kernel<B200_TYPE_IQ4_NL><<<grid, 256, smem, ctx->stream>>>(w, k, m, a->ne02, a->ne03, (const char *)a->src1_dev, n, a->ne12, a->nb11, a->nb12, a->nb13,
                                                                            a->dst_dev, (int)rows_per_cta);
    }
    B200_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return B200_OK;
}

int b200_launch_get_rows_wire(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *rows, const b200_tensor *dst) {
    const int64_t nc = src0->ne[0], nr = rows->ne[0] * rows->ne[1] * rows->ne[2];
    const int wire = src0->type == B200_TYPE_Q5_0 ? 22 : 18;
    B200_REQUIRE(ctx, nc % B200_QK == 0 && src0->nb[0] == wire, B200_ERR_UNSUPPORTED);
    int64_t g = (nr * nc + 255) / 256;
    if (g > (int64_t)ctx->sm_count * 16) g = (int64_t)ctx->sm_count * 16;
    if (g < 1) g = 1;
    if (src0->type == B200_TYPE_Q5_0)
        get_rows_wire_kernel<B200_TYPE_Q5_0><<<(unsigned)g, 256, 0, ctx->stream>>>((const uint8_t *)src0->data, src0->nb[1], src0->nb[2], src0->nb[3], src0->ne[1],
                                                                                  (const char *)rows->data, rows->nb[0], rows->nb[1], rows->nb[2], rows->ne[0], rows->ne[1],
                                                                                  (char *)dst->data, dst->nb[1], dst->nb[2], dst->nb[3], nc, nr);
    else
        get_rows_wire_kernel<B200_TYPE_IQ4_NL><<<(unsigned)g, 256, 0, ctx->stream>>>((const uint8_t *)src0->data, src0->nb[1], src0->nb[2], src0->nb[3], src0->ne[1],
                                                                                    (const char *)rows->data, rows->nb[0], rows->nb[1], rows->nb[2], rows->ne[0], rows->ne[1],
                                                                                    (char *)dst->data, dst->nb[1], dst->nb[2], dst->nb[3], nc, nr);
    B200_CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return B200_OK;
}
