// b200_gemv_stream.cu -- the decode GEMV as a streaming kernel: dst[m, n<=8] = W[m,k] x X[k,n], 2-D case.
//
// Same arithmetic as b200_gemv.cu (fused quantize_row_q8_0 of the activations, src/ggml-quants.c:535-618;
// per-block exact int32 dp4a dots scaled by d_w*d_x and accumulated in fp32, src/ggml-quants.c:3858-3869 /
// :5010-5015) but organised around what bounds a chain of microsecond-sized mul_mats on B200: keeping HBM
// busy ACROSS kernel boundaries.
//
//  * one persistent CTA per SM; CTA c owns a contiguous run of weight rows, i.e. one contiguous byte range of
//    the qs plane and one of the fp16 scale plane (the repacked layout makes both 16-byte aligned);
//  * a producer thread streams that range into a shared-memory ring with 1-D bulk async copies
//    (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) -- no registers are tied up by loads in flight,
//    ~100 KB per SM can be in flight;
//  * programmatic dependent launch: the kernel is co-resident with its predecessor (2 x ~100 KB of shared memory
//    per SM), fills its ring BEFORE griddepcontrol.wait because weights never depend on the previous mul_mat,
//    and only then reads the activations.  HBM therefore keeps streaming the next matrix while the current one
//    is being quantized/reduced/stored;
//  * n == 1: every lane keeps the int8 activations of "its" blocks (lane + 32 i) in registers for the whole
//    kernel, so a weight block costs 16 B + 2 B of shared-memory reads; long rows (k > 4096) are split across
//    2/4/8 warps by k-segment and combined through shared memory in a fixed order;
//  * n in 2..8: activations are read from shared memory per block.
// Shapes outside (k % 256 == 0, k <= 32768, no batch dims) use the generic kernel in b200_gemv.cu.
#include "b200_internal.cuh"

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kThreads = kConsumerThreads + 32;  // + producer warp
constexpr int kMaxStages = 8;
constexpr int kSegBlocks = 128;                  // blocks of k handled by one warp (4 per lane)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void consumer_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct StreamGeom {
    int rs;           // rows per stage
    int pr;           // rows whose k-split partials are parked before one combine
    int stages;       // ring depth
    int g;            // warps per row (k-segments)
    int stage_qs;     // bytes of qs per full stage
    int stage_sc;     // bytes of scales per full stage
    int stage_bytes;  // aligned total
    int act_col;      // bytes of activation scratch per column
    int ring_off, act_off, part_off, bar_off, total;
};

__host__ __device__ inline size_t stream_act_col_bytes(int type, int k) {
    const size_t nb = (size_t)(k >> 5);
    const size_t raw = (size_t)k + nb * 4 + (type == B200_TYPE_Q4_0 ? nb * 4 : 0);
    return (raw + 15) & ~(size_t)15;
}

// block dot of one 16/32-byte weight block against the activation block held as two uint4 (elements 0..15, 16..31)
template <int TYPE>
__device__ __forceinline__ int block_dot(const uint4 &w0, const uint4 &w1, const uint4 &alo, const uint4 &ahi, int s8) {
    int sumi;
    if (TYPE == B200_TYPE_Q4_0) {
        sumi = -s8;  // (nib - 8) . q == nib . q - 8 * sum(q)
        sumi = __dp4a((int)(w0.x & 0x0F0F0F0Fu), (int)alo.x, sumi);
        sumi = __dp4a((int)(w0.y & 0x0F0F0F0Fu), (int)alo.y, sumi);
        sumi = __dp4a((int)(w0.z & 0x0F0F0F0Fu), (int)alo.z, sumi);
        sumi = __dp4a((int)(w0.w & 0x0F0F0F0Fu), (int)alo.w, sumi);
        sumi = __dp4a((int)((w0.x >> 4) & 0x0F0F0F0Fu), (int)ahi.x, sumi);
        sumi = __dp4a((int)((w0.y >> 4) & 0x0F0F0F0Fu), (int)ahi.y, sumi);
        sumi = __dp4a((int)((w0.z >> 4) & 0x0F0F0F0Fu), (int)ahi.z, sumi);
        sumi = __dp4a((int)((w0.w >> 4) & 0x0F0F0F0Fu), (int)ahi.w, sumi);
    } else {
        sumi = __dp4a((int)w0.x, (int)alo.x, 0);
        sumi = __dp4a((int)w0.y, (int)alo.y, sumi);
        sumi = __dp4a((int)w0.z, (int)alo.z, sumi);
        sumi = __dp4a((int)w0.w, (int)alo.w, sumi);
        sumi = __dp4a((int)w1.x, (int)ahi.x, sumi);
        sumi = __dp4a((int)w1.y, (int)ahi.y, sumi);
        sumi = __dp4a((int)w1.z, (int)ahi.z, sumi);
        sumi = __dp4a((int)w1.w, (int)ahi.w, sumi);
    }
    return sumi;
}

template <int TYPE, int NCOLS, bool DOTS>
__global__ void __launch_bounds__(kThreads, 2) gemv_stream_kernel(const b200_gemv_params p, const StreamGeom g) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int QSB = TYPE == B200_TYPE_Q4_0 ? 16 : 32;
    const int k = (int)p.k, nb = k >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row_qs = nb * QSB, row_sc = nb * 2;

    unsigned char *ring = smem + g.ring_off;
    unsigned char *act = smem + g.act_off;
    float *part = reinterpret_cast<float *>(smem + g.part_off);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + g.bar_off);
    uint64_t *empty_bar = full_bar + kMaxStages;

    // this CTA's contiguous run of rows
    const int64_t r_begin = (int64_t)blockIdx.x * p.m / gridDim.x;
    const int64_t r_end = (int64_t)(blockIdx.x + 1) * p.m / gridDim.x;
    const int nrows = (int)(r_end - r_begin);
    const int nstage_iters = (nrows + g.rs - 1) / g.rs;

    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_launch_dependents();

    if (warp == kConsumerWarps) {
        // ===== producer: stream this CTA's byte ranges into the ring (does not wait for the previous grid) =====
        if (lane == 0) {
            const uint8_t *gq = p.qs + r_begin * row_qs;
            const uint8_t *gs = reinterpret_cast<const uint8_t *>(p.d) + r_begin * row_sc;
            for (int it = 0; it < nstage_iters; it++) {
                const int s = it % g.stages;
                const uint32_t ph = (uint32_t)(it / g.stages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                const int rows = min(g.rs, nrows - it * g.rs);
                unsigned char *dst = ring + (size_t)s * g.stage_bytes;
                mbar_expect_tx(&full_bar[s], (uint32_t)(rows * (row_qs + row_sc)));
                bulk_g2s(dst, gq + (size_t)it * g.rs * row_qs, (uint32_t)(rows * row_qs), &full_bar[s]);
                bulk_g2s(dst + g.stage_qs, gs + (size_t)it * g.rs * row_sc, (uint32_t)(rows * row_sc), &full_bar[s]);
            }
        }
        return;
    }

    // ===== consumers =====
    pdl_wait();  // activations (and dst, for write-after-read) belong to the previous grid until here

    // ---- quantize the activation columns into shared memory: quantize_row_q8_0, bit-exact ----
    // 8 lanes per block, one float4 per lane.  Loads are issued in batches of kQB per thread BEFORE any use, so a
    // column costs ceil(k / (4*256*kQB)) L2 round trips instead of one per 1024 elements.  tasks-per-column is a
    // multiple of 64 (k % 256 == 0), so every warp is uniformly live or dead and the shuffles see all 32 lanes.
    {
        constexpr int kQB = 8;
        const int tpc = nb * 8;
#pragma unroll 1
        for (int c = 0; c < NCOLS; c++) {
            const float *xcol = reinterpret_cast<const float *>(reinterpret_cast<const char *>(p.x) + (size_t)c * p.nb11);
            unsigned char *col = act + (size_t)c * g.act_col;
#pragma unroll 1
            for (int base = 0; base < tpc; base += kConsumerThreads * kQB) {
                float4 v[kQB];
#pragma unroll
                for (int u = 0; u < kQB; u++) {
                    const int t = base + u * kConsumerThreads + (int)threadIdx.x;
                    if (t < tpc) v[u] = *reinterpret_cast<const float4 *>(xcol + (size_t)t * 4);
                }
#pragma unroll
                for (int u = 0; u < kQB; u++) {
                    const int t = base + u * kConsumerThreads + (int)threadIdx.x;
                    if (t < tpc) {
                        const int b = t >> 3, sub = t & 7;
                        float amax = fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w)));
                        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
                        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
                        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
                        const float id = amax != 0.0f ? __fdiv_rn(127.f, amax) : 0.0f;
                        const int q0 = __float2int_rn(__fmul_rn(v[u].x, id)), q1 = __float2int_rn(__fmul_rn(v[u].y, id));
                        const int q2 = __float2int_rn(__fmul_rn(v[u].z, id)), q3 = __float2int_rn(__fmul_rn(v[u].w, id));
                        int sq = q0 + q1 + q2 + q3;
                        sq += __shfl_xor_sync(0xffffffffu, sq, 1);
                        sq += __shfl_xor_sync(0xffffffffu, sq, 2);
                        sq += __shfl_xor_sync(0xffffffffu, sq, 4);
                        const uint32_t packed = (uint32_t)(q0 & 0xff) | ((uint32_t)(q1 & 0xff) << 8) | ((uint32_t)(q2 & 0xff) << 16) | ((uint32_t)(q3 & 0xff) << 24);
                        // two planes (elements 0..15 / 16..31 of every block) -> 16-byte reads at stride 16 per lane
                        *reinterpret_cast<uint32_t *>(col + (size_t)(sub >> 2) * (k >> 1) + b * 16 + (sub & 3) * 4) = packed;
                        if (sub == 0) {
                            reinterpret_cast<float *>(col + k)[b] = __half2float(__float2half_rn(__fdiv_rn(amax, 127.f)));
                            if (TYPE == B200_TYPE_Q4_0) reinterpret_cast<int *>(col + k + (size_t)nb * 4)[b] = 8 * sq;
                        }
                    }
                }
            }
        }
    }
    consumer_bar_sync();

    // ---- which (row-in-pass, k-segment) this warp serves ----
    const int G = g.g;
    const int seg = warp % G;
    const int row_in_pass = warp / G;
    const int rows_per_pass = kConsumerWarps / G;
    const int b0 = seg * kSegBlocks;  // first block of the segment

    // n == 1: the lane's activation blocks live in registers for the whole kernel
    uint4 alo[4], ahi[4];
    float da[4];
    int s8[4];
    if (NCOLS == 1) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int b = min(b0 + lane + 32 * i, nb - 1);
            alo[i] = *reinterpret_cast<const uint4 *>(act + (size_t)b * 16);
            ahi[i] = *reinterpret_cast<const uint4 *>(act + (size_t)(k >> 1) + (size_t)b * 16);
            da[i] = reinterpret_cast<const float *>(act + k)[b];
            s8[i] = TYPE == B200_TYPE_Q4_0 ? reinterpret_cast<const int *>(act + k + (size_t)nb * 4)[b] : 0;
        }
    }

    int rows_in_chunk = 0, chunk_row0 = 0, cpar = 0;
    for (int it = 0; it < nstage_iters; it++) {
        const int s = it % g.stages;
        const uint32_t ph = (uint32_t)(it / g.stages) & 1u;
        const int rows = min(g.rs, nrows - it * g.rs);
        const unsigned char *sq = ring + (size_t)s * g.stage_bytes;
        const unsigned char *ssc = sq + g.stage_qs;
        mbar_wait(&full_bar[s], ph);

        for (int r0 = 0; r0 < rows; r0 += rows_per_pass) {
            const int r = r0 + row_in_pass;
            const bool row_live = r < rows;
            const int rr = row_live ? r : rows - 1;
            const unsigned char *wrow = sq + (size_t)rr * row_qs;
            const __half *srow = reinterpret_cast<const __half *>(ssc + (size_t)rr * row_sc);
            const int64_t grow = r_begin + (int64_t)it * g.rs + rr;
            float acc[NCOLS];
#pragma unroll
            for (int c = 0; c < NCOLS; c++) acc[c] = 0.0f;

#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int bb = b0 + lane + 32 * i;
                const bool live = bb < nb && bb < b0 + kSegBlocks;
                const int b = live ? bb : nb - 1;
                const uint4 w0 = *reinterpret_cast<const uint4 *>(wrow + (size_t)b * QSB);
                uint4 w1 = make_uint4(0, 0, 0, 0);
                if (TYPE == B200_TYPE_Q8_0) w1 = *reinterpret_cast<const uint4 *>(wrow + (size_t)b * QSB + 16);
                const float dw = __half2float(srow[b]);
                if (NCOLS == 1) {
                    const int sumi = block_dot<TYPE>(w0, w1, alo[i], ahi[i], s8[i]);
                    if (DOTS) {
                        if (live && row_live) p.dots[grow * nb + b] = sumi;
                    } else if (live) {
                        acc[0] = fmaf((float)sumi, dw * da[i], acc[0]);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < NCOLS; c++) {
                        const unsigned char *col = act + (size_t)c * g.act_col;
                        const uint4 xlo = *reinterpret_cast<const uint4 *>(col + (size_t)b * 16);
                        const uint4 xhi = *reinterpret_cast<const uint4 *>(col + (size_t)(k >> 1) + (size_t)b * 16);
                        const float dx = reinterpret_cast<const float *>(col + k)[b];
                        const int sx = TYPE == B200_TYPE_Q4_0 ? reinterpret_cast<const int *>(col + k + (size_t)nb * 4)[b] : 0;
                        const int sumi = block_dot<TYPE>(w0, w1, xlo, xhi, sx);
                        if (DOTS) {
                            if (live && row_live) p.dots[((int64_t)c * p.m + grow) * nb + b] = sumi;
                        } else if (live) {
                            acc[c] = fmaf((float)sumi, dw * dx, acc[c]);
                        }
                    }
                }
            }

            if (!DOTS) {
#pragma unroll
                for (int c = 0; c < NCOLS; c++) {
                    float v = acc[c];
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    acc[c] = v;
                }
                if (G == 1) {
                    if (row_live && lane < NCOLS) {
                        float v = 0.0f;
#pragma unroll
                        for (int c = 0; c < NCOLS; c++)
                            if (c == lane) v = acc[c];
                        p.dst[(int64_t)lane * p.m + grow] = v;
                    }
                } else if (lane < NCOLS) {
                    // k-split: park this segment's partial; combined below in segment order (deterministic)
                    float v = 0.0f;
#pragma unroll
                    for (int c = 0; c < NCOLS; c++)
                        if (c == lane) v = acc[c];
                    if (row_live) part[((cpar * g.pr + rows_in_chunk + rr) * kConsumerWarps + seg) * NCOLS + lane] = v;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);  // this warp is done reading stage s

        if (!DOTS && G > 1) {
            // k-split: partials of up to g.pr rows are parked in shared memory (double-buffered by chunk parity) and
            // combined in segment order after ONE barrier per chunk -- for decode shapes that is once per kernel
            rows_in_chunk += rows;
            if (it == nstage_iters - 1 || rows_in_chunk + g.rs > g.pr) {
                consumer_bar_sync();
                for (int t = threadIdx.x; t < rows_in_chunk * NCOLS; t += kConsumerThreads) {
                    const int r = t / NCOLS, c = t - r * NCOLS;
                    float v = 0.0f;
                    for (int sg = 0; sg < G; sg++) v += part[((cpar * g.pr + r) * kConsumerWarps + sg) * NCOLS + c];
                    p.dst[(int64_t)c * p.m + r_begin + chunk_row0 + r] = v;
                }
                chunk_row0 += rows_in_chunk;
                rows_in_chunk = 0;
                cpar ^= 1;
            }
        }
    }
}

bool stream_geometry(const b200_gemv_params &p, StreamGeom *g) {
    const int k = (int)p.k, nb = k >> 5;
    const int qsb = p.type == B200_TYPE_Q4_0 ? 16 : 32;
    const int row_qs = nb * qsb, row_sc = nb * 2, row_bytes = row_qs + row_sc;
    int G = (nb + kSegBlocks - 1) / kSegBlocks;
    if (G > 8) return false;
    if (G == 3) G = 4;
    if (G > 4 && G < 8) G = 8;
    const int rows_per_pass = kConsumerWarps / G;
    int rs = (16 * 1024) / row_bytes;
    if (rs < rows_per_pass) rs = rows_per_pass;
    if (rs > 64) rs = 64;
    rs = rs / rows_per_pass * rows_per_pass;
    g->rs = rs;
    g->g = G;
    g->stage_qs = rs * row_qs;
    g->stage_sc = rs * row_sc;
    g->stage_bytes = (int)b200_align_up((size_t)g->stage_qs + g->stage_sc, 128);
    g->act_col = (int)stream_act_col_bytes(p.type, k);
    const int act_bytes = (int)b200_align_up((size_t)g->act_col * p.n, 128);
    int pr = 64 / (int)p.n;
    if (pr < rs) pr = rs;
    g->pr = pr;
    const int part_bytes = G > 1 ? (int)b200_align_up((size_t)2 * pr * kConsumerWarps * p.n * 4, 128) : 128;
    const int bar_bytes = 2 * kMaxStages * 8;
    // two of these kernels must be co-resident per SM (current + programmatic dependent): <= ~110 KB each
    const int budget = 110 * 1024 - act_bytes - part_bytes - bar_bytes - 256;
    int stages = budget / g->stage_bytes;
    if (stages < 2) return false;
    if (stages > kMaxStages) stages = kMaxStages;
    g->stages = stages;
    g->ring_off = 0;
    g->act_off = stages * g->stage_bytes;
    g->part_off = g->act_off + act_bytes;
    g->bar_off = g->part_off + part_bytes;
    g->total = g->bar_off + bar_bytes;
    return true;
}

template <int TYPE, int NCOLS>
int launch_stream_typed(b200_ctx *ctx, const b200_gemv_params &p, const StreamGeom &g, bool dots) {
    auto kern = dots ? gemv_stream_kernel<TYPE, NCOLS, true> : gemv_stream_kernel<TYPE, NCOLS, false>;
    B200_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    int64_t ctas = ctx->sm_count;
    if (ctas > p.m) ctas = p.m;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)ctas, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = (size_t)g.total;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ctx->opt_pdl ? 1 : 0;
    B200_CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, kern, p, g));
    ctx->launches++;
    return B200_OK;
}

template <int TYPE>
int launch_stream_cols(b200_ctx *ctx, const b200_gemv_params &p, const StreamGeom &g, bool dots) {
    switch (p.n) {
        case 1: return launch_stream_typed<TYPE, 1>(ctx, p, g, dots);
        case 2: return launch_stream_typed<TYPE, 2>(ctx, p, g, dots);
        case 3: return launch_stream_typed<TYPE, 3>(ctx, p, g, dots);
        case 4: return launch_stream_typed<TYPE, 4>(ctx, p, g, dots);
        case 5: return launch_stream_typed<TYPE, 5>(ctx, p, g, dots);
        case 6: return launch_stream_typed<TYPE, 6>(ctx, p, g, dots);
        case 7: return launch_stream_typed<TYPE, 7>(ctx, p, g, dots);
        case 8: return launch_stream_typed<TYPE, 8>(ctx, p, g, dots);
    }
    return B200_ERR_INVALID;
}

}  // namespace

// returns true when the streaming kernel takes this shape; *rc then holds the launch status
bool b200_try_launch_gemv_stream(b200_ctx *ctx, const b200_gemv_params &p, int *rc) {
    if (p.ne12 != 1 || p.ne13 != 1 || p.ne02 != 1 || p.ne03 != 1) return false;
    if (p.k % 256 != 0 || p.k > 32768 || p.n < 1 || p.n > 8) return false;
    if (((uintptr_t)p.qs & 15) != 0 || ((uintptr_t)p.d & 15) != 0) return false;
    if (p.dots == NULL && p.dst_n != p.n) return false;  // column-chunked dst keeps the generic addressing
    StreamGeom g;
    if (!stream_geometry(p, &g)) return false;
    const bool dots = p.dots != NULL;
    if (p.type == B200_TYPE_Q4_0) *rc = launch_stream_cols<B200_TYPE_Q4_0>(ctx, p, g, dots);
    else if (p.type == B200_TYPE_Q8_0) *rc = launch_stream_cols<B200_TYPE_Q8_0>(ctx, p, g, dots);
    else return false;
    return true;
}
