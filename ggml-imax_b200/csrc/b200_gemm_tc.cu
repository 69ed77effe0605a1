// b200_gemm_tc.cu -- the EXACT prefill kernel: dst[n][m] = W[m,k] (int8, per-32 fp16 scales) x Xq[n,k] (Q8_0 activations),
// tcgen05.mma kind::i8 with int32 accumulators in TMEM, one MMA per 32-wide quant block.  It is what b200_block_dots(path 1)
// reads the bit-exact per-block partials from, what serves 9 <= n < 32, and what the "gemm_exact" option selects for any n;
// the default for n >= 32 is the fp16 contraction in b200_gemm_f16.cu (the per-block fp32 scaling on CUDA cores bounds this
// kernel at ~9 % of the int8 tensor peak, profiles/r01_gemm_experiments.md).
//
// Stands in for the COMPUTE phase of ggml_compute_forward_mul_mat (src/ggml.c:12056-12096) at n >= 9:
//   dst[n][m] = sum_kb  d_w[m][kb] * d_x[n][kb] * ( sum_{j<32} w[m][32 kb + j] * q[n][32 kb + j] )
// The inner integer sum is exactly the per-block partial of ggml_vec_dot_q4_0_q8_0 / _q8_0_q8_0
// (src/ggml-quants.c:3858-3869, :5010-5015).  One tcgen05.mma has K = 32 for 8-bit operands == one quant block,
// so every MMA produces the exact int32 block partials of a 128 x 128 tile; they are read back from TMEM,
// converted, multiplied by the product of the two fp16 scales and accumulated in fp32 registers.
//
// CTA = one 128(m) x 128(n) output tile, 20 warps:
//   warp 0      TMA producer: A tile (128 rows x 128 B) and B tile (128 x 128 B) per stage, SWIZZLE_128B, 4 k-blocks
//   warp 1      MMA issuer: 4 x tcgen05.mma (K=32 each) per stage, each into its own 128-column TMEM buffer
//               (4 buffers = all 512 columns), tcgen05.commit -> mbarriers (TMEM full, smem stage empty)
//   warp 2      TMEM allocator / deallocator
//   warps 4..   epilogue (B200_GEMM_EPI_WARPS = 8 or 16): warp w reads TMEM lanes 32*(w%4).. (its hardware quadrant), columns 32*((w-4)/4)..;
//               thread = one weight row: d_w is a per-thread scalar, d_x a warp-uniform (broadcast) load;
//               32 fp32 accumulators per thread; final store is coalesced along m.
// Q4_0 weights are expanded to int8 (nib - 8) by expand_q4_0_kernel into scratch first (v1; fusing the expansion
// behind the TMA load is the next step).
#include "b200_tc_common.cuh"

using namespace b200tc;

namespace {

constexpr int BM = 128, BN = 128, BK = 128;  // BK bytes of int8 = one 128-byte swizzle atom = 4 quant blocks
constexpr int kStages = 4;
constexpr int kTmemBufs = 4;
#ifndef B200_GEMM_EPI_WARPS
#define B200_GEMM_EPI_WARPS 8
#endif
constexpr int kEpiWarps = B200_GEMM_EPI_WARPS;   // a multiple of 4 (one per TMEM lane quadrant); each takes BN / (kEpiWarps/4) columns
constexpr int kGemmThreads = 128 + kEpiWarps * 32;
constexpr int kScaleBytes = (BK / 32) * BN * 4;            // activation scales of the stage's 4 k-blocks: [4][128] fp32
constexpr int kStageBytes = BM * BK + BN * BK + kScaleBytes;  // 34 KB (a multiple of 1024: operand tiles stay 1024-aligned)
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;

// instruction descriptor: dense, no saturate, D = S32, A = B = signed 8-bit, both K-major, N = 128, M = 128
constexpr uint32_t kIdescI8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct GemmArgs {
    const __half *dw;   // weight scales [m][nb]
    const float *dxT;   // activation scales, transposed fp32 [nb][ldn]
    int64_t ldn;
    float *dst;         // [n][m]
    int32_t *dots;      // parity dump [n][m][nb] or null
    int m, n, k;
};

template <bool DOTS>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_i8_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_s, const GemmArgs g) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kStages * kStageBytes);
    uint64_t *full_bar = bars;                       // [kStages]   TMA -> MMA
    uint64_t *empty_bar = bars + kStages;            // [kStages]   MMA -> TMA
    uint64_t *tfull_bar = bars + 2 * kStages;        // [kTmemBufs] MMA -> epilogue
    uint64_t *tempty_bar = tfull_bar + kTmemBufs;    // [kTmemBufs] epilogue -> MMA
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty_bar + kTmemBufs);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int nb = g.k >> 5;
    const int kiters = (g.k + BK - 1) / BK;          // TMA zero-fills a ragged last stage

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_s) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1 + kEpiWarps); }
        for (int b = 0; b < kTmemBufs; b++) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int it = 0; it < kiters; it++) {
                const int s = it % kStages;
                const uint32_t ph = (uint32_t)(it / kStages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                unsigned char *sa = smem + s * kStageBytes;
                unsigned char *sb = sa + BM * BK;
                mbar_expect_tx(&full_bar[s], kStageBytes);
                tma_load_2d(sa, &map_a, it * BK, m0, &full_bar[s]);
                tma_load_2d(sb, &map_b, it * BK, n0, &full_bar[s]);
                tma_load_2d(sb + BN * BK, &map_s, n0, it * (BK / 32), &full_bar[s]);   // d_x[4 k-blocks][128 columns]
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            for (int it = 0; it < kiters; it++) {
                const int s = it % kStages;
                const uint32_t ph = (uint32_t)(it / kStages) & 1u;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * kStageBytes);
                const uint32_t sb = sa + BM * BK;
                const uint64_t da = make_desc_sw128(sa), db = make_desc_sw128(sb);
                // Two k-blocks per hand-shake: TMEM buffers {0,1} and {2,3} form two groups that ping-pong between this thread
                // and the epilogue.  One wait, two MMAs (each a fresh accumulator: one quant block), one commit per group --
                // the per-k-block barrier traffic of both sides is what bounded the kernel (profiles/r01_gemm_experiments.md).
                const uint32_t tph = (uint32_t)it & 1u;                  // every group is used once per stage
#pragma unroll
                for (int grp = 0; grp < 2; grp++) {
                    const int kb0 = it * (BK / 32) + 2 * grp;
                    if (kb0 < nb) {
                        mbar_wait(&tempty_bar[grp], tph ^ 1u);           // epilogue has drained both buffers of the group
                        tc_fence_after();
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int j = 2 * grp + h;
                            // K advance inside the 128-byte swizzle atom: +32 bytes = +2 in the (>>4) start-address field
                            if (kb0 + h < nb) tc_mma_i8(tmem_base + (uint32_t)(j * BN), da + (uint64_t)(j * 2), db + (uint64_t)(j * 2), kIdescI8, 0u);
                        }
                        tc_commit(&tfull_bar[grp]);                      // arrives when both MMAs have written TMEM
                    }
                }
                tc_commit(&empty_bar[s]);                        // smem stage reusable once its MMAs have read it
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue =====
        const int ew = warp - 4;
        const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
        const int cgrp = ew >> 2;                  // which kCols-wide column group of the tile
        const int row = m0 + quad * 32 + lane;     // weight row owned by this thread
        const int row_c = row < g.m ? row : g.m - 1;
        const __half *dwp = g.dw + (int64_t)row_c * nb;
        constexpr int kCols = BN / (kEpiWarps / 4);
        static_assert(kCols % 32 == 0, "whole tcgen05.ld.32x32b.x32 loads per k-block per warp");
        float acc[kCols];
        unsigned long long acc2[kCols / 2];   // packed pairs of fp32 accumulators
#pragma unroll
        for (int j = 0; j < kCols / 2; j++) acc2[j] = 0ull;

        // One iteration = one smem stage = 4 k-blocks = 4 TMEM buffers (kTmemBufs == BK/32), so buffer indices are
        // compile-time, the mbarrier parity is one bit per iteration, and the 4 weight scales of the stage are one
        // 8-byte load issued a full iteration ahead.  The scale math is packed f32x2 (FMUL2 / FFMA2 on sm_100).
        static_assert(kTmemBufs == BK / 32, "one TMEM buffer per k-block of a stage");
        const uint32_t smem_a = smem_u32(smem);
        const bool dw_vec = (nb & 3) == 0;                 // rows of d_w are 8-byte aligned
        uint2 dwq = make_uint2(0, 0);
        auto load_dw = [&](int it) -> uint2 {
            const int kb0 = it * 4;
            if (dw_vec) return *reinterpret_cast<const uint2 *>(dwp + kb0);
            uint16_t h[4];
#pragma unroll
            for (int j = 0; j < 4; j++) h[j] = kb0 + j < nb ? __half_as_ushort(dwp[kb0 + j]) : (uint16_t)0;
            return make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
        };
        dwq = load_dw(0);
        static_assert(kCols == 64, "software pipeline below is written for two 32-column halves per k-block");
        const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cgrp * kCols);
        // scale math for 32 columns of one k-block: acc += float(p) * (d_w * d_x), packed f32x2
        auto scale_acc = [&](const uint32_t (&pv)[32], uint32_t sdx_kb, unsigned long long dw2, int half) {
#pragma unroll
            for (int c4 = 0; c4 < 8; c4++) {
                unsigned long long d01, d23;   // d_x of 4 columns: warp-uniform shared-memory broadcast
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(d01), "=l"(d23) : "r"(sdx_kb + (uint32_t)(half * 128 + c4 * 16)));
                const unsigned long long s01 = mul2(dw2, d01), s23 = mul2(dw2, d23);
                const unsigned long long v01 = pack2((float)(int32_t)pv[c4 * 4 + 0], (float)(int32_t)pv[c4 * 4 + 1]);
                const unsigned long long v23 = pack2((float)(int32_t)pv[c4 * 4 + 2], (float)(int32_t)pv[c4 * 4 + 3]);
                acc2[half * 16 + c4 * 2 + 0] = fma2(v01, s01, acc2[half * 16 + c4 * 2 + 0]);
                acc2[half * 16 + c4 * 2 + 1] = fma2(v23, s23, acc2[half * 16 + c4 * 2 + 1]);
            }
        };
        auto dump_dots = [&](const uint32_t (&pv)[32], int kb, int half) {
            if (row < g.m) {
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    const int col = n0 + cgrp * kCols + half * 32 + c;
                    if (col < g.n) g.dots[((int64_t)col * g.m + row) * nb + kb] = (int32_t)pv[c];
                }
            }
        };
        // Software pipeline over (k-block, half): the TMEM load of the NEXT half is in flight while the scale math of the
        // current one runs (two register sets pa / pb); the TMEM buffer is handed back to the MMA warp as soon as its second
        // half has landed in registers.
        uint32_t pa[32], pb[32];
        int st = 0;
        uint32_t tph = 0;
        mbar_wait(&tfull_bar[0], 0);
        tc_fence_after();
        tc_ld32(tcol, pa);                                   // (kb 0, half 0) in flight
        for (int it = 0; it < kiters; it++) {
            const uint2 dwcur = dwq;
            if (it + 1 < kiters) dwq = load_dw(it + 1);
            const uint32_t sdx = smem_a + (uint32_t)(st * kStageBytes + BM * BK + BN * BK + cgrp * kCols * 4);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int kb = it * 4 + j;
                if (kb < nb) {
                    const uint32_t hbits = (j < 2 ? dwcur.x : dwcur.y) >> ((j & 1) * 16);
                    const float dw = __half2float(__ushort_as_half((unsigned short)(hbits & 0xffffu)));
                    const unsigned long long dw2 = pack2(dw, dw);
                    const uint32_t tbuf = tcol + (uint32_t)(j * BN);
                    tc_wait_ld();                            // pa = (kb, half 0)
                    tc_ld32(tbuf + 32, pb);                  // (kb, half 1) in flight
                    if (DOTS) dump_dots(pa, kb, 0); else scale_acc(pa, sdx + (uint32_t)(j * BN * 4), dw2, 0);
                    tc_wait_ld();                            // pb = (kb, half 1): this warp is done with TMEM buffer j
                    if ((j & 1) == 1 || kb + 1 >= nb) {      // ... and, after the group's second k-block, with the group
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[j >> 1]);
                    }
                    if (kb + 1 < nb) {                       // next k-block's first half: buffer (j + 1) % 4
                        const int jn = (j + 1) & 3;
                        if ((jn & 1) == 0) {                 // first k-block of the next group: one wait per two k-blocks
                            mbar_wait(&tfull_bar[jn >> 1], jn == 0 ? (tph ^ 1u) : tph);
                            tc_fence_after();
                        }
                        tc_ld32(tcol + (uint32_t)(jn * BN), pa);
                    }
                    if (DOTS) dump_dots(pb, kb, 1); else scale_acc(pb, sdx + (uint32_t)(j * BN * 4), dw2, 1);
                }
            }
            // done with this stage's scales: let the producer refill it
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[st]);
            tph ^= 1u;
            if (++st == kStages) st = 0;
        }
#pragma unroll
        for (int c2 = 0; c2 < kCols / 2; c2++) unpack2(acc2[c2], acc[c2 * 2], acc[c2 * 2 + 1]);
        if (!DOTS && row < g.m) {
#pragma unroll
            for (int j = 0; j < kCols; j++) {
                const int c = n0 + cgrp * kCols + j;
                if (c < g.n) g.dst[(int64_t)c * g.m + row] = acc[j];   // 32 lanes -> 128 contiguous bytes
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// Q4_0 qs plane (packed nibbles) -> int8 (nib - 8) rows [m][k]; one thread per block
__global__ void __launch_bounds__(256) expand_q4_0_kernel(const uint4 *__restrict__ qs, uint4 *__restrict__ out, int64_t nblocks) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const uint4 v = qs[b];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        // per byte: (nib | 0x80) - 8, then ^ 0x80  ==  nib - 8 in two's complement, no borrow across bytes
        lo[i] = (((w[i] & 0x0F0F0F0Fu) | 0x80808080u) - 0x08080808u) ^ 0x80808080u;
        hi[i] = ((((w[i] >> 4) & 0x0F0F0F0Fu) | 0x80808080u) - 0x08080808u) ^ 0x80808080u;
    }
    out[2 * b] = make_uint4(lo[0], lo[1], lo[2], lo[3]);       // elements 0..15
    out[2 * b + 1] = make_uint4(hi[0], hi[1], hi[2], hi[3]);   // elements 16..31
}

// fp16 activation scales [n][nb] -> fp32 transposed [nb][ldn] (zero padded) for warp-uniform loads in the epilogue
__global__ void __launch_bounds__(256) transpose_scales_kernel(const __half *__restrict__ d, float *__restrict__ dT, int n, int nb, int64_t ldn) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nb * ldn) return;
    const int kb = (int)(t / ldn), c = (int)(t - (int64_t)kb * ldn);
    dT[t] = c < n ? __half2float(d[(int64_t)c * nb + kb]) : 0.0f;
}


}  // namespace

PFN_cuTensorMapEncodeTiled_v12000 b200tc::get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
        else
            (void)cudaGetLastError();
    }
    return fn;
}

namespace {

// rows x k bytes of int8, row pitch k; box = 128 rows x 128 bytes, 128-byte swizzle, out-of-bounds -> zeros
bool make_map(CUtensorMap *map, const void *base, int64_t rows, int64_t k) {
    auto fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)k};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// transposed activation scales [nb][ldn] fp32; box = 4 k-blocks x 128 columns, rows past nb -> zeros
bool make_scale_map(CUtensorMap *map, const void *base, int64_t nb, int64_t ldn) {
    auto fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)ldn, (cuuint64_t)nb};
    cuuint64_t strides[1] = {(cuuint64_t)ldn * 4};
    cuuint32_t box[2] = {(cuuint32_t)BN, (cuuint32_t)(BK / 32)};
    cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool b200_gemm_available(void) { return get_encode_fn() != nullptr; }

// extra scratch the GEMM needs beyond the quantized activations: transposed scales (+ expanded Q4_0 weights)
size_t b200_gemm_scratch_bytes(int type, int64_t k, int64_t m, int64_t n) {
    const int64_t nb = k / 32, ldn = (n + BN - 1) / BN * BN;
    size_t b = b200_align_up((size_t)nb * ldn * 4, 256);
    if (type == B200_TYPE_Q4_0) b += b200_align_up((size_t)m * k, 256);
    return b;
}

int b200_launch_gemm(b200_ctx *ctx, const b200_gemm_params &p) {
    B200_REQUIRE(ctx, p.k % 32 == 0 && p.k >= 32 && p.m >= 1 && p.n >= 1, B200_ERR_INVALID);
    B200_REQUIRE(ctx, p.scratch != NULL, B200_ERR_INVALID);
    B200_REQUIRE(ctx, p.m < (1 << 30) && p.n < (1 << 30) && p.k < (1 << 30), B200_ERR_UNSUPPORTED);
    const int64_t nb = p.k / 32, ldn = (p.n + BN - 1) / BN * BN;
    float *dxT = (float *)p.scratch;
    const int8_t *a8 = (const int8_t *)p.qs;
    {
        const int64_t total = nb * ldn;
        transpose_scales_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(p.ad, dxT, (int)p.n, (int)nb, ldn);
        ctx->launches++;
    }
    if (p.type == B200_TYPE_Q4_0) {
        int8_t *w8 = (int8_t *)((uint8_t *)p.scratch + b200_align_up((size_t)nb * ldn * 4, 256));
        const int64_t nblocks = p.m * nb;
        expand_q4_0_kernel<<<(unsigned)((nblocks + 255) / 256), 256, 0, ctx->stream>>>((const uint4 *)p.qs, (uint4 *)w8, nblocks);
        ctx->launches++;
        a8 = w8;
    }
    B200_CUDA_TRY(ctx, cudaGetLastError());
    CUtensorMap map_a, map_b, map_s;
    if (!make_map(&map_a, a8, p.m, p.k) || !make_map(&map_b, p.aq, p.n, p.k) || !make_scale_map(&map_s, dxT, nb, ldn)) {
        b200_set_error(ctx, "cuTensorMapEncodeTiled failed (m=%lld n=%lld k=%lld)", (long long)p.m, (long long)p.n, (long long)p.k);
        return B200_ERR_CUDA;
    }
    GemmArgs g;
    g.dw = p.d;
    g.dxT = dxT;
    g.ldn = ldn;
    g.dst = p.dst;
    g.dots = p.dots;
    g.m = (int)p.m;
    g.n = (int)p.n;
    g.k = (int)p.k;
    dim3 grid((unsigned)((p.m + BM - 1) / BM), (unsigned)((p.n + BN - 1) / BN), 1);
    if (p.dots) {
        B200_CUDA_TRY(ctx, cudaFuncSetAttribute(gemm_i8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        gemm_i8_kernel<true><<<grid, kGemmThreads, kSmemBytes, ctx->stream>>>(map_a, map_b, map_s, g);
    } else {
        B200_CUDA_TRY(ctx, cudaFuncSetAttribute(gemm_i8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        gemm_i8_kernel<false><<<grid, kGemmThreads, kSmemBytes, ctx->stream>>>(map_a, map_b, map_s, g);
    }
    ctx->launches++;
    B200_CUDA_TRY(ctx, cudaGetLastError());
    return B200_OK;
}
