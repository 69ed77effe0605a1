// b200_gemv.cu -- decode path: dst[m, n<=8] = W[m,k] (Q4_0/Q8_0 planes) x X[k, n] (F32), HBM-bound.
//
// Stands in for the COMPUTE phase of ggml_compute_forward_mul_mat (src/ggml.c:12056-12096) calling
// ggml_vec_dot_q4_0_q8_0 / ggml_vec_dot_q8_0_q8_0 (src/ggml-quants.c:3469, :4819), with the INIT-phase
// quantize_row_q8_0 of src1 (src/ggml.c:11952-11974) fused into the prologue: every CTA quantizes the
// (tiny, L2-resident) activation columns into shared memory itself, so a decode mul_mat is ONE launch.
//
// No tensor cores (north_star): weights stream once from HBM with 128-bit loads, one 16-byte chunk per
// lane (512 contiguous bytes per warp request); integer dots are dp4a against the int8 activations held
// in shared memory; each block's exact int32 partial is converted and multiplied by d_w * d_x (the
// product of two fp16 values is exact in fp32) and accumulated in fp32; lanes are combined with xor-shuffles.
//   Q4_0: chunk == block.  (nib - 8) . q  ==  nib . q  -  8 * sum(q);  8*sum(q) is precomputed per block.
//   Q8_0: chunk == half a block; the two halves are added (shuffle) BEFORE scaling so the int32 partial
//         is per 32-wide block as in the reference.
// Grid: persistent, a multiple of the SM count; warps stride over row pairs.
// Programmatic dependent launch: the kernel signals launch_dependents at once and only waits for its
// predecessor (griddepcontrol.wait) after it has prefetched its first weight rows into L2 -- weights never
// depend on the previous mul_mat, only the activations do.  All global writes come after the wait.
#include "b200_internal.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRows = 2;  // rows per warp iteration

__device__ __forceinline__ uint4 ldg_stream(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// shared memory per activation column (stride col_bytes, 16-byte aligned):
//   [q bytes: k][d_x: nb floats][8*sum(q): nb ints (Q4_0 only)]
__host__ __device__ __forceinline__ size_t gemv_col_bytes(int type, int64_t k) {
    const size_t nb = (size_t)(k >> 5);
    const size_t raw = (size_t)k + nb * 4 + (type == B200_TYPE_Q4_0 ? nb * 4 : 0);
    return (raw + 15) & ~(size_t)15;
}

__device__ __forceinline__ int dp4a_ss(int a, int b, int c) { return __dp4a(a, b, c); }
// nibbles 0..15 are non-negative as int8 too, so the signed form serves both operand kinds
__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) { return __dp4a((int)a, b, c); }

template <int TYPE, int NCOLS, bool DOTS>
__global__ void __launch_bounds__(kThreads) gemv_kernel(const b200_gemv_params p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int k = (int)p.k, nb = k >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t col_bytes = gemv_col_bytes(TYPE, k);

    // batch / broadcast indices (src/ggml.c:11848-11849, :12063-12065)
    const int64_t i12 = blockIdx.y % p.ne12, i13 = blockIdx.y / p.ne12;
    const int64_t i02 = i12 / (p.ne12 / p.ne02), i03 = i13 / (p.ne13 / p.ne03);
    const int64_t wrow0 = (i03 * p.ne02 + i02) * p.m;  // first weight row of this 2-D slice
    constexpr int QSB = TYPE == B200_TYPE_Q4_0 ? 16 : 32;
    const uint8_t *wq = p.qs + wrow0 * (int64_t)nb * QSB;
    const __half *wd = p.d + wrow0 * (int64_t)nb;
    const int64_t row_bytes = (int64_t)nb * QSB;

    const int64_t gw = (int64_t)blockIdx.x * kWarps + warp;
    const int64_t nw = (int64_t)gridDim.x * kWarps;

    pdl_launch_dependents();
    // weights do not depend on the previous kernel: pull this warp's first rows towards L2 now
    {
        const int64_t row = gw * kRows;
        if (row < p.m) {
            const int64_t bytes = min((int64_t)kRows, p.m - row) * row_bytes;
            const uint8_t *base = wq + row * row_bytes;
            for (int64_t off = (int64_t)lane * 128; off < bytes; off += 32 * 128) prefetch_l2(base + off);
        }
    }
    pdl_wait();

    // ---- prologue: quantize the NCOLS activation columns into shared memory (quantize_row_q8_0) ----
    {
        const char *xbase = reinterpret_cast<const char *>(p.x) + i13 * p.nb13 + i12 * p.nb12;
        const int tasks = NCOLS * nb * 8;
        for (int t = threadIdx.x; t < ((tasks + 31) & ~31); t += kThreads) {
            const bool live = t < tasks;
            const int tt = live ? t : tasks - 1;
            const int c = tt / (nb * 8);
            const int r = tt - c * (nb * 8);
            const int b = r >> 3, sub = r & 7;
            const float4 v = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(xbase + (size_t)c * p.nb11) + b * 32 + sub * 4);
            float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
            const float id = amax != 0.0f ? __fdiv_rn(127.f, amax) : 0.0f;
            const int q0 = __float2int_rn(__fmul_rn(v.x, id)), q1 = __float2int_rn(__fmul_rn(v.y, id));
            const int q2 = __float2int_rn(__fmul_rn(v.z, id)), q3 = __float2int_rn(__fmul_rn(v.w, id));
            int s = q0 + q1 + q2 + q3;
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (live) {
                const uint32_t packed = (uint32_t)(q0 & 0xff) | ((uint32_t)(q1 & 0xff) << 8) | ((uint32_t)(q2 & 0xff) << 16) | ((uint32_t)(q3 & 0xff) << 24);
                unsigned char *col = smem + (size_t)c * col_bytes;
                if (TYPE == B200_TYPE_Q4_0) {
                    // two planes so a lane's two 16-byte reads are bank-conflict free:
                    //   plane 0 = elements 0..15 of every block, plane 1 = elements 16..31
                    const int plane = sub >> 2;
                    *reinterpret_cast<uint32_t *>(col + (size_t)plane * (k >> 1) + b * 16 + (sub & 3) * 4) = packed;
                } else {
                    *reinterpret_cast<uint32_t *>(col + b * 32 + sub * 4) = packed;
                }
                if (sub == 0) {
                    const __half dh = __float2half_rn(__fdiv_rn(amax, 127.f));  // the fp16 the reference stores in y[i].d
                    reinterpret_cast<float *>(col + k)[b] = __half2float(dh);
                    if (TYPE == B200_TYPE_Q4_0) reinterpret_cast<int *>(col + k + (size_t)nb * 4)[b] = 8 * s;
                }
            }
        }
    }
    __syncthreads();

    // ---- main loop ----
    for (int64_t row = gw * kRows; row < p.m; row += nw * kRows) {
        float acc[kRows][NCOLS];
#pragma unroll
        for (int r = 0; r < kRows; r++)
#pragma unroll
            for (int c = 0; c < NCOLS; c++) acc[r][c] = 0.0f;
        const uint8_t *rq[kRows];
        const __half *rd[kRows];
#pragma unroll
        for (int r = 0; r < kRows; r++) {
            const int64_t rr = min(row + r, p.m - 1);
            rq[r] = wq + rr * row_bytes;
            rd[r] = wd + rr * nb;
        }
        // pull the rows this warp will need next towards L2 while it works on the current ones
        {
            const int64_t nrow = row + nw * kRows;
            if (nrow < p.m) {
                const int64_t bytes = min((int64_t)kRows, p.m - nrow) * row_bytes;
                const uint8_t *base = wq + nrow * row_bytes;
                for (int64_t off = (int64_t)lane * 128; off < bytes; off += 32 * 128) prefetch_l2(base + off);
            }
        }

        if (TYPE == B200_TYPE_Q4_0) {
#pragma unroll 4
            for (int b0 = 0; b0 < nb; b0 += 32) {
                const bool live = b0 + lane < nb;
                const int b = live ? b0 + lane : nb - 1;
                uint4 w[kRows];
                float dw[kRows];
#pragma unroll
                for (int r = 0; r < kRows; r++) {
                    w[r] = ldg_stream(rq[r] + (size_t)b * 16);
                    dw[r] = __half2float(rd[r][b]);
                }
#pragma unroll
                for (int c = 0; c < NCOLS; c++) {
                    const unsigned char *col = smem + (size_t)c * col_bytes;
                    const uint4 qlo = *reinterpret_cast<const uint4 *>(col + (size_t)b * 16);
                    const uint4 qhi = *reinterpret_cast<const uint4 *>(col + (size_t)(k >> 1) + (size_t)b * 16);
                    const float da = reinterpret_cast<const float *>(col + k)[b];
                    const int s8 = reinterpret_cast<const int *>(col + k + (size_t)nb * 4)[b];
#pragma unroll
                    for (int r = 0; r < kRows; r++) {
                        int sumi = -s8;
                        sumi = dp4a_us(w[r].x & 0x0F0F0F0Fu, (int)qlo.x, sumi);
                        sumi = dp4a_us(w[r].y & 0x0F0F0F0Fu, (int)qlo.y, sumi);
                        sumi = dp4a_us(w[r].z & 0x0F0F0F0Fu, (int)qlo.z, sumi);
                        sumi = dp4a_us(w[r].w & 0x0F0F0F0Fu, (int)qlo.w, sumi);
                        sumi = dp4a_us((w[r].x >> 4) & 0x0F0F0F0Fu, (int)qhi.x, sumi);
                        sumi = dp4a_us((w[r].y >> 4) & 0x0F0F0F0Fu, (int)qhi.y, sumi);
                        sumi = dp4a_us((w[r].z >> 4) & 0x0F0F0F0Fu, (int)qhi.z, sumi);
                        sumi = dp4a_us((w[r].w >> 4) & 0x0F0F0F0Fu, (int)qhi.w, sumi);
                        if (DOTS) {
                            if (live && row + r < p.m) p.dots[((int64_t)c * p.m + row + r) * nb + b] = sumi;
                        } else if (live) {
                            acc[r][c] = fmaf((float)sumi, dw[r] * da, acc[r][c]);
                        }
                    }
                }
            }
        } else {
            const int nch = nb * 2;
#pragma unroll 4
            for (int c0 = 0; c0 < nch; c0 += 32) {
                // uniform trip count: the xor-shuffle below needs all 32 lanes; nch is even, so a lane and
                // its partner lane^1 are live or dead together
                const bool live = c0 + lane < nch;
                const int ch = live ? c0 + lane : nch - 1;
                uint4 w[kRows];
                float dw[kRows];
#pragma unroll
                for (int r = 0; r < kRows; r++) {
                    w[r] = ldg_stream(rq[r] + (size_t)ch * 16);
                    dw[r] = __half2float(rd[r][ch >> 1]);
                }
#pragma unroll
                for (int c = 0; c < NCOLS; c++) {
                    const unsigned char *col = smem + (size_t)c * col_bytes;
                    const uint4 q = *reinterpret_cast<const uint4 *>(col + (size_t)ch * 16);
                    const float da = reinterpret_cast<const float *>(col + k)[ch >> 1];
#pragma unroll
                    for (int r = 0; r < kRows; r++) {
                        int sumi = dp4a_ss((int)w[r].x, (int)q.x, 0);
                        sumi = dp4a_ss((int)w[r].y, (int)q.y, sumi);
                        sumi = dp4a_ss((int)w[r].z, (int)q.z, sumi);
                        sumi = dp4a_ss((int)w[r].w, (int)q.w, sumi);
                        // nch is even and lanes own consecutive chunks: lane^1 holds the other half-block
                        sumi += __shfl_xor_sync(0xffffffffu, sumi, 1);
                        if (DOTS) {
                            if (live && !(lane & 1) && row + r < p.m) p.dots[((int64_t)c * p.m + row + r) * nb + (ch >> 1)] = sumi;
                        } else if (live && !(lane & 1)) {
                            acc[r][c] = fmaf((float)sumi, dw[r] * da, acc[r][c]);
                        }
                    }
                }
            }
        }

        if (!DOTS) {
#pragma unroll
            for (int r = 0; r < kRows; r++)
#pragma unroll
                for (int c = 0; c < NCOLS; c++) {
                    float v = acc[r][c];
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    acc[r][c] = v;
                }
            // dst[i13][i12][c][row]  (dst->ne0 == m contiguous, src/ggml.c:4834-4838)
            float *dst = p.dst + ((i13 * p.ne12 + i12) * p.dst_n) * p.m;
            if (lane < kRows * NCOLS) {
                const int r = lane / NCOLS, c = lane % NCOLS;
                float v = 0.0f;
#pragma unroll
                for (int rr = 0; rr < kRows; rr++)
#pragma unroll
                    for (int cc = 0; cc < NCOLS; cc++)
                        if (rr == r && cc == c) v = acc[rr][cc];
                if (row + r < p.m) dst[(int64_t)c * p.m + row + r] = b200_gemv_epilogue(p, v, row + r, (int64_t)c * p.m + row + r);
            }
        }
    }
}

template <int TYPE, int NCOLS>
int launch_typed(b200_ctx *ctx, const b200_gemv_params &p, bool dots) {
    const size_t smem = (size_t)NCOLS * gemv_col_bytes(TYPE, p.k);
    B200_REQUIRE(ctx, smem <= 200 * 1024, B200_ERR_UNSUPPORTED);
    auto kern = dots ? gemv_kernel<TYPE, NCOLS, true> : gemv_kernel<TYPE, NCOLS, false>;
    if (smem > 48 * 1024) {
        if (dots) B200_SMEM_LIMIT_ONCE(ctx, (gemv_kernel<TYPE, NCOLS, true>), 200 * 1024);
        else B200_SMEM_LIMIT_ONCE(ctx, (gemv_kernel<TYPE, NCOLS, false>), 200 * 1024);
    }
    // persistent grid: a multiple of the SM count, no more CTAs than there are row pairs
    const int64_t row_groups = (p.m + kRows - 1) / kRows;
    int64_t ctas = (row_groups + kWarps - 1) / kWarps;
    int per_sm = 4;
    if (smem > 48 * 1024) per_sm = 2;
    if (smem > 100 * 1024) per_sm = 1;
    const int64_t cap = (int64_t)ctx->sm_count * per_sm;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)ctas, (unsigned)(p.ne12 * p.ne13), 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ctx->opt_pdl ? 1 : 0;
    B200_CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, kern, p));
    ctx->launches++;
    return B200_OK;
}

template <int TYPE>
int launch_cols(b200_ctx *ctx, const b200_gemv_params &p, bool dots) {
    switch (p.n) {
        case 1: return launch_typed<TYPE, 1>(ctx, p, dots);
        case 2: return launch_typed<TYPE, 2>(ctx, p, dots);
        case 3: return launch_typed<TYPE, 3>(ctx, p, dots);
        case 4: return launch_typed<TYPE, 4>(ctx, p, dots);
        case 5: return launch_typed<TYPE, 5>(ctx, p, dots);
        case 6: return launch_typed<TYPE, 6>(ctx, p, dots);
        case 7: return launch_typed<TYPE, 7>(ctx, p, dots);
        case 8: return launch_typed<TYPE, 8>(ctx, p, dots);
    }
    return B200_ERR_INVALID;
}

}  // namespace

// n <= 8 columns per launch; callers chunk larger n.
int b200_launch_gemv(b200_ctx *ctx, const b200_gemv_params &p) {
    B200_REQUIRE(ctx, p.n >= 1 && p.n <= 8, B200_ERR_INVALID);
    B200_REQUIRE(ctx, p.k > 0 && p.k % 32 == 0 && p.m > 0, B200_ERR_INVALID);
    B200_REQUIRE(ctx, p.ne12 * p.ne13 <= 65535, B200_ERR_UNSUPPORTED);
    B200_REQUIRE(ctx, ((uintptr_t)p.x & 15) == 0 && (p.nb11 & 15) == 0 && (p.nb12 & 15) == 0 && (p.nb13 & 15) == 0, B200_ERR_UNSUPPORTED);
    const bool dots = p.dots != NULL;
    if (ctx->opt_gemv_stream) {  // the streaming kernel (b200_gemv_stream.cu) takes the 2-D, k % 256 == 0 shapes
        int rc = B200_OK;
        if (b200_try_launch_gemv_stream(ctx, p, &rc)) return rc;
    }
    if (p.type == B200_TYPE_Q4_0) return launch_cols<B200_TYPE_Q4_0>(ctx, p, dots);
    if (p.type == B200_TYPE_Q8_0) return launch_cols<B200_TYPE_Q8_0>(ctx, p, dots);
    return B200_ERR_UNSUPPORTED;
}
