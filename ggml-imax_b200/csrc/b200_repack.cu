// b200_repack.cu -- wire format <-> plane layout of Q4_0 / Q8_0 tensors.
//
// ggml stores a quantized row as an array of blocks {fp16 d; qs[]} (src/ggml-common.h:144-149,
// :186-191): 18-byte (Q4_0) / 34-byte (Q8_0) records, so no block payload is 16-byte aligned.
// On the device the same ggml_nbytes() hold two planes instead (see include/ggml_b200.h):
//     qs plane  : block b -> 16 B (Q4_0) / 32 B (Q8_0) at b * qs_bytes      (128-bit loadable, TMA-able)
//     d  plane  : block b -> fp16 at nblocks_total * qs_bytes + 2 * b
// The nibble order inside a Q4_0 block is kept (byte j = element j | element j+16 << 4): masking with
// 0x0F0F0F0F / shifting by 4 yields element order 0..15 / 16..31, which is what dp4a against the Q8_0
// activation words and the int8 expansion for the tensor-core path both want.
//
// These kernels run once per set_tensor / get_tensor (ggml_backend_buffer_i.set_tensor/get_tensor,
// src/ggml-backend-impl.h:43-44); they are HBM-trivial (one read + one write of the tensor).
#include "b200_internal.cuh"

namespace {

constexpr int kBlocksPerCta = 256;

template <int WIRE, int QS>
__global__ void __launch_bounds__(kBlocksPerCta) repack_kernel(const uint16_t *__restrict__ wire, uint8_t *__restrict__ qs_plane,
                                                                uint16_t *__restrict__ d_plane, int64_t nblocks) {
    constexpr int H = WIRE / 2;  // u16 per wire block
    __shared__ uint16_t sm[kBlocksPerCta * H];
    const int64_t b0 = (int64_t)blockIdx.x * kBlocksPerCta;
    const int64_t nb = min((int64_t)kBlocksPerCta, nblocks - b0);
    const uint16_t *src = wire + b0 * H;
    for (int i = threadIdx.x; i < nb * H; i += kBlocksPerCta) sm[i] = src[i];
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= nb) return;
    const uint16_t *blk = sm + t * H;
    d_plane[b0 + t] = blk[0];
    uint32_t w[QS / 4];
#pragma unroll
    for (int i = 0; i < QS / 4; i++) w[i] = (uint32_t)blk[1 + 2 * i] | ((uint32_t)blk[2 + 2 * i] << 16);
    uint4 *dst = reinterpret_cast<uint4 *>(qs_plane + (b0 + t) * QS);
#pragma unroll
    for (int i = 0; i < QS / 16; i++) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

template <int WIRE, int QS>
__global__ void __launch_bounds__(kBlocksPerCta) unrepack_kernel(uint16_t *__restrict__ wire, const uint8_t *__restrict__ qs_plane,
                                                                  const uint16_t *__restrict__ d_plane, int64_t nblocks) {
    constexpr int H = WIRE / 2;
    __shared__ uint16_t sm[kBlocksPerCta * H];
    const int64_t b0 = (int64_t)blockIdx.x * kBlocksPerCta;
    const int64_t nb = min((int64_t)kBlocksPerCta, nblocks - b0);
    const int t = threadIdx.x;
    if (t < nb) {
        uint16_t *blk = sm + t * H;
        blk[0] = d_plane[b0 + t];
        const uint4 *src = reinterpret_cast<const uint4 *>(qs_plane + (b0 + t) * QS);
#pragma unroll
        for (int i = 0; i < QS / 16; i++) {
            const uint4 v = src[i];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                blk[1 + 2 * (4 * i + j)] = (uint16_t)(w[j] & 0xffffu);
                blk[2 + 2 * (4 * i + j)] = (uint16_t)(w[j] >> 16);
            }
        }
    }
    __syncthreads();
    uint16_t *dst = wire + b0 * H;
    for (int i = threadIdx.x; i < nb * H; i += kBlocksPerCta) dst[i] = sm[i];
}

}  // namespace

int b200_launch_repack(b200_ctx *ctx, int type, void *tensor_dev, int64_t nblocks_total, const void *wire_dev,
                       int64_t block_off, int64_t nblocks) {
    if (nblocks <= 0) return B200_OK;
    const int qsb = b200_qs_bytes(type);
    uint8_t *qs = (uint8_t *)tensor_dev + block_off * qsb;
    uint16_t *d = (uint16_t *)((uint8_t *)tensor_dev + nblocks_total * qsb) + block_off;
    const unsigned grid = (unsigned)((nblocks + kBlocksPerCta - 1) / kBlocksPerCta);
    if (type == B200_TYPE_Q4_0)
        repack_kernel<B200_Q4_0_BYTES, 16><<<grid, kBlocksPerCta, 0, ctx->stream>>>((const uint16_t *)wire_dev, qs, d, nblocks);
    else
        repack_kernel<B200_Q8_0_BYTES, 32><<<grid, kBlocksPerCta, 0, ctx->stream>>>((const uint16_t *)wire_dev, qs, d, nblocks);
    ctx->launches++;
    B200_CUDA_TRY(ctx, cudaGetLastError());
    return B200_OK;
}

int b200_launch_unrepack(b200_ctx *ctx, int type, const void *tensor_dev, int64_t nblocks_total, void *wire_dev,
                         int64_t block_off, int64_t nblocks) {
    if (nblocks <= 0) return B200_OK;
    const int qsb = b200_qs_bytes(type);
    const uint8_t *qs = (const uint8_t *)tensor_dev + block_off * qsb;
    const uint16_t *d = (const uint16_t *)((const uint8_t *)tensor_dev + nblocks_total * qsb) + block_off;
    const unsigned grid = (unsigned)((nblocks + kBlocksPerCta - 1) / kBlocksPerCta);
    if (type == B200_TYPE_Q4_0)
        unrepack_kernel<B200_Q4_0_BYTES, 16><<<grid, kBlocksPerCta, 0, ctx->stream>>>((uint16_t *)wire_dev, qs, d, nblocks);
    else
        unrepack_kernel<B200_Q8_0_BYTES, 32><<<grid, kBlocksPerCta, 0, ctx->stream>>>((uint16_t *)wire_dev, qs, d, nblocks);
    ctx->launches++;
    B200_CUDA_TRY(ctx, cudaGetLastError());
    return B200_OK;
}

static bool quant_type_ok(int type) { return type == B200_TYPE_Q4_0 || type == B200_TYPE_Q8_0; }

// chunk of wire blocks staged on the device per copy (18 MiB / 34 MiB)
static const int64_t kStageBlocks = 1 << 20;

extern "C" {

int b200_repack_from_device(b200_ctx *ctx, int type, void *tensor_dev, int64_t nblocks_total, const void *src_wire_dev,
                            int64_t block_off, int64_t nblocks) {
    B200_REQUIRE(ctx, ctx && quant_type_ok(type), B200_ERR_INVALID);
    B200_REQUIRE(ctx, block_off >= 0 && nblocks >= 0 && block_off + nblocks <= nblocks_total, B200_ERR_INVALID);
    B200_REQUIRE(ctx, ((uintptr_t)tensor_dev & 15) == 0 && ((uintptr_t)src_wire_dev & 1) == 0, B200_ERR_INVALID);
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return b200_launch_repack(ctx, type, tensor_dev, nblocks_total, src_wire_dev, block_off, nblocks);
}

int b200_unrepack_to_device(b200_ctx *ctx, int type, const void *tensor_dev, int64_t nblocks_total, void *dst_wire_dev,
                            int64_t block_off, int64_t nblocks) {
    B200_REQUIRE(ctx, ctx && quant_type_ok(type), B200_ERR_INVALID);
    B200_REQUIRE(ctx, block_off >= 0 && nblocks >= 0 && block_off + nblocks <= nblocks_total, B200_ERR_INVALID);
    B200_REQUIRE(ctx, ((uintptr_t)tensor_dev & 15) == 0 && ((uintptr_t)dst_wire_dev & 1) == 0, B200_ERR_INVALID);
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return b200_launch_unrepack(ctx, type, tensor_dev, nblocks_total, dst_wire_dev, block_off, nblocks);
}

int b200_set_quantized(b200_ctx *ctx, int type, void *tensor_dev, int64_t nblocks_total, const void *src_host,
                       int64_t block_off, int64_t nblocks) {
    B200_REQUIRE(ctx, ctx && quant_type_ok(type), B200_ERR_INVALID);
    B200_REQUIRE(ctx, block_off >= 0 && nblocks >= 0 && block_off + nblocks <= nblocks_total, B200_ERR_INVALID);
    B200_REQUIRE(ctx, ((uintptr_t)tensor_dev & 15) == 0, B200_ERR_INVALID);
    if (nblocks == 0) return B200_OK;
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int wire = b200_wire_bytes(type);
    const int64_t chunk = nblocks < kStageBlocks ? nblocks : kStageBlocks;
    int rc = b200_stage_reserve(ctx, (size_t)chunk * wire);
    if (rc != B200_OK) return rc;
    for (int64_t done = 0; done < nblocks; done += chunk) {
        const int64_t nb = (nblocks - done) < chunk ? (nblocks - done) : chunk;
        B200_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage, (const uint8_t *)src_host + done * wire, (size_t)nb * wire,
                                           cudaMemcpyHostToDevice, ctx->stream));
        rc = b200_launch_repack(ctx, type, tensor_dev, nblocks_total, ctx->stage, block_off + done, nb);
        if (rc != B200_OK) return rc;
    }
    B200_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

int b200_get_quantized(b200_ctx *ctx, int type, const void *tensor_dev, int64_t nblocks_total, void *dst_host,
                       int64_t block_off, int64_t nblocks) {
    B200_REQUIRE(ctx, ctx && quant_type_ok(type), B200_ERR_INVALID);
    B200_REQUIRE(ctx, block_off >= 0 && nblocks >= 0 && block_off + nblocks <= nblocks_total, B200_ERR_INVALID);
    B200_REQUIRE(ctx, ((uintptr_t)tensor_dev & 15) == 0, B200_ERR_INVALID);
    if (nblocks == 0) return B200_OK;
    B200_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int wire = b200_wire_bytes(type);
    const int64_t chunk = nblocks < kStageBlocks ? nblocks : kStageBlocks;
    int rc = b200_stage_reserve(ctx, (size_t)chunk * wire);
    if (rc != B200_OK) return rc;
    for (int64_t done = 0; done < nblocks; done += chunk) {
        const int64_t nb = (nblocks - done) < chunk ? (nblocks - done) : chunk;
        rc = b200_launch_unrepack(ctx, type, tensor_dev, nblocks_total, ctx->stage, block_off + done, nb);
        if (rc != B200_OK) return rc;
        B200_CUDA_TRY(ctx, cudaMemcpyAsync((uint8_t *)dst_host + done * wire, ctx->stage, (size_t)nb * wire,
                                           cudaMemcpyDeviceToHost, ctx->stream));
    }
    B200_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return B200_OK;
}

}  // extern "C"
