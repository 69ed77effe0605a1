// b200_quantize.cu -- F32 activations -> Q8_0, bit-exact with the reference's runtime from_float.
//
// Reference: quantize_row_q8_0, src/ggml-quants.c:465 with the AVX2 body :535-618 (the branch the x86
// reference build runs): per 32 floats  amax = max|x|;  d = amax / 127 -> fp16 (RNE);
// id = amax != 0 ? 127 / amax : 0;  q = int8(round_half_even(x * id)).
// (The scalar quantize_row_q8_0_reference :440-463 uses 1/d and roundf and differs in a few elements per
//  million; it is NOT what mul_mat runs, src/ggml.c:697-712.)
//
// Mapping: 8 lanes per block, one 128-bit load (4 floats) per lane, amax by 3 xor-shuffles, one 32-bit
// store of 4 packed int8 per lane -> every warp reads 512 contiguous bytes and writes 128.
// Arithmetic uses the explicitly rounded intrinsics so -use_fast_math or contraction can never change it.
#include "b200_internal.cuh"

namespace {

__device__ __forceinline__ float4 ld_f4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

// quantize 4 consecutive floats of a block given the block amax; returns packed int8x4
__device__ __forceinline__ uint32_t q8_pack4(float4 v, float id) {
    const int q0 = __float2int_rn(__fmul_rn(v.x, id));
    const int q1 = __float2int_rn(__fmul_rn(v.y, id));
    const int q2 = __float2int_rn(__fmul_rn(v.z, id));
    const int q3 = __float2int_rn(__fmul_rn(v.w, id));
    return (uint32_t)(q0 & 0xff) | ((uint32_t)(q1 & 0xff) << 8) | ((uint32_t)(q2 & 0xff) << 16) | ((uint32_t)(q3 & 0xff) << 24);
}

// WIRE == false: planar output qs[row][k] + d[row][k/32];  WIRE == true: block_q8_0 records (34 B)
template <bool WIRE>
__global__ void __launch_bounds__(256) quantize_q8_0_kernel(const float *__restrict__ x, int64_t k, int64_t nrows, size_t row_stride,
                                                            int8_t *__restrict__ qs, uint16_t *__restrict__ dpl) {
    const int64_t nb = k / 32;
    const int64_t total = nrows * nb * 8;  // lane-tasks (8 per block)
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < b200_align_up(total, 32);
         t += (int64_t)gridDim.x * blockDim.x) {
        const bool live = t < total;
        const int64_t blk = (live ? t : total - 1) >> 3;
        const int sub = (int)(t & 7);
        const int64_t row = blk / nb, b = blk - row * nb;
        const float *src = reinterpret_cast<const float *>(reinterpret_cast<const char *>(x) + row * row_stride) + b * 32 + sub * 4;
        const float4 v = ld_f4(src);
        float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
        const float id = amax != 0.0f ? __fdiv_rn(127.f, amax) : 0.0f;
        const uint32_t packed = q8_pack4(v, id);
        if (!live) continue;
        if (WIRE) {
            // 34-byte records: 2-byte aligned only -> 16-bit stores
            uint16_t *rec = reinterpret_cast<uint16_t *>(reinterpret_cast<uint8_t *>(qs) + blk * B200_Q8_0_BYTES);
            rec[1 + sub * 2] = (uint16_t)(packed & 0xffffu);
            rec[2 + sub * 2] = (uint16_t)(packed >> 16);
            if (sub == 0) rec[0] = __half_as_ushort(__float2half_rn(__fdiv_rn(amax, 127.f)));
        } else {
            reinterpret_cast<uint32_t *>(qs)[blk * 8 + sub] = packed;
            if (sub == 0) dpl[blk] = __half_as_ushort(__float2half_rn(__fdiv_rn(amax, 127.f)));
        }
    }
}

}  // namespace

static int launch_quantize(b200_ctx *ctx, bool wire, const float *x_dev, int64_t k, int64_t nrows, size_t row_stride_bytes,
                           int8_t *qs_dev, uint16_t *d_dev) {
    B200_REQUIRE(ctx, ctx != NULL, B200_ERR_INVALID);
    B200_REQUIRE(ctx, k > 0 && k % 32 == 0 && nrows >= 0, B200_ERR_INVALID);
    B200_REQUIRE(ctx, ((uintptr_t)x_dev & 15) == 0 && (row_stride_bytes & 15) == 0, B200_ERR_UNSUPPORTED);
    if (nrows == 0) return B200_OK;
    B200_CUDA_TRY(ctx, b200_use_device(ctx->device));
    const int64_t total = nrows * (k / 32) * 8;
    int64_t grid = (total + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (grid > cap) grid = cap;
    if (wire)
        quantize_q8_0_kernel<true><<<(unsigned)grid, 256, 0, ctx->stream>>>(x_dev, k, nrows, row_stride_bytes, qs_dev, d_dev);
    else
        quantize_q8_0_kernel<false><<<(unsigned)grid, 256, 0, ctx->stream>>>(x_dev, k, nrows, row_stride_bytes, qs_dev, d_dev);
    ctx->launches++;
    B200_CUDA_TRY(ctx, cudaGetLastError());
    return B200_OK;
}

extern "C" {

int b200_quantize_q8_0(b200_ctx *ctx, const float *x_dev, int64_t k, int64_t nrows, size_t row_stride_bytes, int8_t *qs_dev,
                       uint16_t *d_dev) {
    return launch_quantize(ctx, false, x_dev, k, nrows, row_stride_bytes, qs_dev, d_dev);
}

int b200_quantize_q8_0_blocks(b200_ctx *ctx, const float *x_dev, int64_t k, int64_t nrows, size_t row_stride_bytes,
                              void *blocks_dev) {
    return launch_quantize(ctx, true, x_dev, k, nrows, row_stride_bytes, (int8_t *)blocks_dev, NULL);
}

}  // extern "C"
