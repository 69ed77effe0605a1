// b200_mul_mat.cu -- the mul_mat entry points of include/ggml_b200.h: argument checks that mirror the
// reference's asserts (src/ggml.c:11832-11849, src/ggml-quants.c:3473) and dispatch between the decode
// GEMV (b200_gemv.cu) and the prefill tensor-core GEMM (b200_gemm_tc.cu).  No CPU path exists here.
#include "b200_internal.cuh"

static bool quant_type_ok(int type) { return type == B200_TYPE_Q4_0 || type == B200_TYPE_Q8_0; }

// n > gemv_max_n goes to the tcgen05 GEMM when it can take the shape, else column chunks through the GEMV.
static bool use_gemm(const b200_ctx *ctx, const b200_mul_mat_args *a) {
    if (a->flags & B200_MM_FORCE_GEMV) return false;
    if (!b200_gemm_available()) return false;
    if (a->flags & B200_MM_FORCE_GEMM) return true;
    if (!ctx->opt_gemm) return false;
    if (a->ne11 > ctx->opt_gemv_max_n) return true;
    // The streaming GEMV keeps its quantized activation columns in shared memory next to the weight ring; past n * k of 64 KB (Q4_0) /
    // 32 KB (Q8_0: twice the ring bytes per row) it falls back to the generic kernel, and the tensor-core path -- its columns padded to
    // one tile -- is as fast or faster from there on (profiles/r02_sweep_n.log: 4096 x 16384 Q4_0, n = 8: 96 us against 48 us).
    return a->ne11 >= 2 && a->ne11 * a->ne00 >= (a->type == B200_TYPE_Q8_0 ? 32768 : 65536);
}

static int run_gemv_chunks(b200_ctx *ctx, const b200_mul_mat_args *a, const uint8_t *qs, const __half *d, const b200_epilogue *epi = nullptr) {
    // column chunks of <= 8; each chunk re-streams the weights (only used when the GEMM cannot serve the shape).  The generic
    // GEMV keeps the chunk's quantized columns in shared memory (k + k/32 * 8 bytes each, 200 KB in all): long rows take
    // narrower chunks, so that every shape supports_op accepts (k <= 131072) really runs.
    const size_t col_bytes = (size_t)a->ne00 + (size_t)(a->ne00 / B200_QK) * 8 + 16;
    int64_t chunk = 8;
    while (chunk > 1 && (size_t)chunk * col_bytes > 200 * 1024) chunk--;
    for (int64_t c0 = 0; c0 < a->ne11; c0 += chunk) {
        b200_gemv_params p;
        memset(&p, 0, sizeof(p));
        p.type = a->type;
        p.qs = qs;
        p.d = d;
        p.k = a->ne00;
        p.m = a->ne01;
        p.ne02 = a->ne02;
        p.ne03 = a->ne03;
        p.x = reinterpret_cast<const float *>(reinterpret_cast<const char *>(a->src1_dev) + c0 * a->nb11);
        p.n = (a->ne11 - c0) < chunk ? (a->ne11 - c0) : chunk;
        p.ne12 = a->ne12;
        p.ne13 = a->ne13;
        p.nb11 = a->nb11;
        p.nb12 = a->nb12;
        p.nb13 = a->nb13;
        p.dst = a->dst_dev + c0 * a->ne01;
        // dst batch stride must stay ne11*m: tell the kernel the full n for addressing via a second field
        // (the kernel addresses dst as ((i13*ne12+i12)*n_total + c)*m + row; n_total is carried in dst_n)
        p.dst_n = a->ne11;
        if (epi) {
            p.bias = epi->bias_dev;
            p.residual = epi->residual_dev;
            p.residual2 = epi->residual2_dev;
            p.act = epi->act;
        }
        int rc = b200_launch_gemv(ctx, p);
        if (rc != B200_OK) return rc;
    }
    return B200_OK;
}

size_t b200_prefill_ws_bytes(int type, int64_t k, int64_t m, int64_t n) {
    const size_t exact = b200_align_up((size_t)n * k, 256) + b200_align_up((size_t)n * (k / 32) * 2, 256) + b200_gemm_scratch_bytes(type, k, m, n);
    const size_t f16 = b200_gemm_f16_scratch_bytes(k, m, n, 160);      // (an upper bound on the pairs of any sm_100 part)
    return exact > f16 ? exact : f16;
}

// n at which the fp16 tensor-core contraction takes over from the exact per-block kernel.  It was 32 ("the 256-column tiles are mostly
// padding below that") until the sweep over n showed the exact kernel taking 67 us on 4096 x 4096 at n = 31 where the fp16 one takes
// 29 us at n = 32 (profiles/r02_sweep_n.log): padding costs nothing next to streaming and dequantizing the weights.
static const int64_t kGemmF16MinN = 1;

static int run_gemm_f16(b200_ctx *ctx, const b200_mul_mat_args *a, const uint8_t *qs, const __half *d) {
    const int64_t k = a->ne00, m = a->ne01, n = a->ne11, nb = k / 32;
    int rc = b200_ws_reserve(ctx, b200_gemm_f16_scratch_bytes(k, m, n, ctx->sm_count));
    if (rc != B200_OK) return rc;
    const int64_t r2 = a->ne12 / a->ne02, r3 = a->ne13 / a->ne03;
    const int qsb = b200_qs_bytes(a->type);
    for (int64_t i13 = 0; i13 < a->ne13; i13++)
        for (int64_t i12 = 0; i12 < a->ne12; i12++) {
            const float *x = reinterpret_cast<const float *>(reinterpret_cast<const char *>(a->src1_dev) + i13 * a->nb13 + i12 * a->nb12);
            const int64_t wrow0 = ((i13 / r3) * a->ne02 + (i12 / r2)) * m;
            rc = b200_launch_gemm_f16(ctx, a->type, qs + wrow0 * nb * qsb, d + wrow0 * nb, k, m, x, n, a->nb11,
                                      a->dst_dev + (i13 * a->ne12 + i12) * n * m, ctx->ws);
            if (rc != B200_OK) return rc;
        }
    return B200_OK;
}

static int run_gemm(b200_ctx *ctx, const b200_mul_mat_args *a, const uint8_t *qs, const __half *d) {
    const int64_t k = a->ne00, m = a->ne01, n = a->ne11, nb = k / 32;
    if (!ctx->opt_gemm_exact && n >= kGemmF16MinN) return run_gemm_f16(ctx, a, qs, d);
    // scratch: int8 plane [n][k] + fp16 scales [n][nb] + GEMM scratch, reused per (i12,i13) slice
    const size_t q_bytes = b200_align_up((size_t)n * k, 256);
    const size_t d_bytes = b200_align_up((size_t)n * nb * 2, 256);
    int rc = b200_ws_reserve(ctx, b200_prefill_ws_bytes(a->type, k, m, n));
    if (rc != B200_OK) return rc;
    int8_t *aq = (int8_t *)ctx->ws;
    uint16_t *ad = (uint16_t *)((uint8_t *)ctx->ws + q_bytes);
    void *scratch = (uint8_t *)ctx->ws + q_bytes + d_bytes;
    const int64_t r2 = a->ne12 / a->ne02, r3 = a->ne13 / a->ne03;
    const int qsb = b200_qs_bytes(a->type);
    for (int64_t i13 = 0; i13 < a->ne13; i13++)
        for (int64_t i12 = 0; i12 < a->ne12; i12++) {
            const float *x = reinterpret_cast<const float *>(reinterpret_cast<const char *>(a->src1_dev) + i13 * a->nb13 + i12 * a->nb12);
            rc = b200_quantize_q8_0(ctx, x, k, n, a->nb11, aq, ad);
            if (rc != B200_OK) return rc;
            const int64_t wrow0 = ((i13 / r3) * a->ne02 + (i12 / r2)) * m;
            b200_gemm_params g;
            memset(&g, 0, sizeof(g));
            g.type = a->type;
            g.qs = qs + wrow0 * nb * qsb;
            g.d = d + wrow0 * nb;
            g.k = k;
            g.m = m;
            g.aq = aq;
            g.ad = reinterpret_cast<const __half *>(ad);
            g.n = n;
            g.dst = a->dst_dev + (i13 * a->ne12 + i12) * n * m;
            g.scratch = scratch;
            rc = b200_launch_gemm(ctx, g);
            if (rc != B200_OK) return rc;
        }
    return B200_OK;
}

extern "C" {

int b200_mul_mat(b200_ctx *ctx, const b200_mul_mat_args *a) {
    B200_REQUIRE(ctx, ctx && a, B200_ERR_INVALID);
    if (a->type == B200_TYPE_Q5_0 || a->type == B200_TYPE_IQ4_NL) {
        // the sibling formats: wire-format blocks at src0_dev (+ src0_block_off blocks), plain correctness path
        B200_REQUIRE(ctx, a->src0_dev && a->src1_dev && a->dst_dev && a->src0_block_off >= 0, B200_ERR_INVALID);
        B200_REQUIRE(ctx, ((uintptr_t)a->src1_dev & 3) == 0 && a->nb11 % 4 == 0 && a->nb12 % 4 == 0 && a->nb13 % 4 == 0, B200_ERR_UNSUPPORTED);
        B200_CUDA_TRY(ctx, b200_use_device(ctx->device));
        return b200_launch_gemv_wire(ctx, a);
    }
    B200_REQUIRE(ctx, quant_type_ok(a->type), B200_ERR_UNSUPPORTED);
    B200_REQUIRE(ctx, a->ne00 > 0 && a->ne00 % B200_QK == 0, B200_ERR_INVALID);   // assert(n % qk == 0)
    B200_REQUIRE(ctx, a->ne01 > 0 && a->ne02 > 0 && a->ne03 > 0, B200_ERR_INVALID);
    B200_REQUIRE(ctx, a->ne11 > 0 && a->ne12 > 0 && a->ne13 > 0, B200_ERR_INVALID);
    B200_REQUIRE(ctx, a->ne12 % a->ne02 == 0 && a->ne13 % a->ne03 == 0, B200_ERR_INVALID);  // ggml_can_mul_mat
    B200_REQUIRE(ctx, a->src0_dev && a->src1_dev && a->dst_dev, B200_ERR_INVALID);
    const int64_t nb = a->ne00 / B200_QK;
    const int64_t nblk = nb * a->ne01 * a->ne02 * a->ne03;
    B200_REQUIRE(ctx, a->src0_block_off >= 0 && a->src0_block_off + nblk <= a->src0_nblocks_total, B200_ERR_INVALID);
    B200_CUDA_TRY(ctx, b200_use_device(ctx->device));
    const int qsb = b200_qs_bytes(a->type);
    const uint8_t *qs = (const uint8_t *)a->src0_dev + a->src0_block_off * qsb;
    const __half *d = (const __half *)((const uint8_t *)a->src0_dev + a->src0_nblocks_total * qsb) + a->src0_block_off;
    if (use_gemm(ctx, a)) return run_gemm(ctx, a, qs, d);
    return run_gemv_chunks(ctx, a, qs, d);
}

int b200_mul_mat_fused(b200_ctx *ctx, const b200_mul_mat_args *a, const b200_epilogue *epi) {
    B200_REQUIRE(ctx, ctx && a, B200_ERR_INVALID);
    if (!epi || (!epi->bias_dev && !epi->residual_dev && !epi->residual2_dev && epi->act == B200_EPI_NONE)) return b200_mul_mat(ctx, a);
    B200_REQUIRE(ctx, epi->act == B200_EPI_NONE || epi->act == B200_EPI_GELU, B200_ERR_INVALID);
    B200_REQUIRE(ctx, quant_type_ok(a->type), B200_ERR_UNSUPPORTED);
    B200_REQUIRE(ctx, a->ne00 > 0 && a->ne00 % B200_QK == 0 && a->ne01 > 0 && a->ne11 > 0, B200_ERR_INVALID);
    B200_REQUIRE(ctx, a->src0_dev && a->src1_dev && a->dst_dev, B200_ERR_INVALID);
    // one GEMV launch only: 2-D, at most 8 columns (a chunked or tensor-core mul_mat keeps its operators separate)
    if (a->ne02 != 1 || a->ne03 != 1 || a->ne12 != 1 || a->ne13 != 1 || a->ne11 > 8 || (a->flags & B200_MM_FORCE_GEMM) ||
        (a->ne11 > ctx->opt_gemv_max_n && !(a->flags & B200_MM_FORCE_GEMV))) {
        b200_set_error(ctx, "b200_mul_mat_fused: not a single-launch decode shape");
        return B200_ERR_UNSUPPORTED;
    }
    const size_t col_bytes = (size_t)a->ne00 + (size_t)(a->ne00 / B200_QK) * 8 + 16;
    if ((size_t)a->ne11 * col_bytes > 200 * 1024) {
        b200_set_error(ctx, "b200_mul_mat_fused: the columns do not fit one launch");
        return B200_ERR_UNSUPPORTED;
    }
    const int64_t nblk = (a->ne00 / B200_QK) * a->ne01;
    B200_REQUIRE(ctx, a->src0_block_off >= 0 && a->src0_block_off + nblk <= a->src0_nblocks_total, B200_ERR_INVALID);
    B200_CUDA_TRY(ctx, b200_use_device(ctx->device));
    const int qsb = b200_qs_bytes(a->type);
    const uint8_t *qs = (const uint8_t *)a->src0_dev + a->src0_block_off * qsb;
    const __half *d = (const __half *)((const uint8_t *)a->src0_dev + a->src0_nblocks_total * qsb) + a->src0_block_off;
    return run_gemv_chunks(ctx, a, qs, d, epi);
}

static bool decode_params(const b200_ctx *ctx, const b200_mul_mat_args *a, b200_gemv_params *p) {
    if (!quant_type_ok(a->type) || (a->flags & B200_MM_FORCE_GEMM)) return false;
    if (a->ne11 != 1 || a->ne12 != 1 || a->ne13 != 1 || a->ne02 != 1 || a->ne03 != 1) return false;
    if (a->ne00 <= 0 || a->ne00 % B200_QK != 0 || a->ne01 <= 0 || !a->src0_dev || !a->src1_dev || !a->dst_dev) return false;
    const int64_t nb = a->ne00 / B200_QK;
    if (a->src0_block_off < 0 || a->src0_block_off + nb * a->ne01 > a->src0_nblocks_total) return false;
    if (((uintptr_t)a->src1_dev & 15) != 0) return false;
    (void)ctx;
    const int qsb = b200_qs_bytes(a->type);
    memset(p, 0, sizeof(*p));
    p->type = a->type;
    p->qs = (const uint8_t *)a->src0_dev + a->src0_block_off * qsb;
    p->d = (const __half *)((const uint8_t *)a->src0_dev + a->src0_nblocks_total * qsb) + a->src0_block_off;
    p->k = a->ne00; p->m = a->ne01; p->ne02 = 1; p->ne03 = 1;
    p->x = a->src1_dev;
    p->n = 1; p->ne12 = 1; p->ne13 = 1;
    p->nb11 = a->nb11; p->nb12 = a->nb12; p->nb13 = a->nb13;
    p->dst = a->dst_dev;
    p->dst_n = 1;
    return true;
}

int b200_mul_mat_batch(b200_ctx *ctx, const b200_mul_mat_args *args, int count) {
    B200_REQUIRE(ctx, ctx && args && count >= 0, B200_ERR_INVALID);
    int i = 0;
    while (i < count) {
        // longest run (<= 4) starting at i that one streaming launch can take
        b200_gemv_params ps[4];
        int run = 0;
        while (run < 4 && i + run < count && ctx->opt_gemv_stream && decode_params(ctx, &args[i + run], &ps[run]) &&
               ps[run].type == ps[0].type && ps[run].k == ps[0].k && ps[run].x == ps[0].x)
            run++;
        int rc = B200_OK;
        if (run >= 2) {
            B200_CUDA_TRY(ctx, b200_use_device(ctx->device));
            if (b200_try_launch_gemv_stream_batch(ctx, ps, run, &rc)) {
                if (rc != B200_OK) return rc;
                i += run;
                continue;
            }
        }
        rc = b200_mul_mat(ctx, &args[i]);
        if (rc != B200_OK) return rc;
        i++;
    }
    return B200_OK;
}

static bool gather_ok(const b200_gather *g) {
    if (!g || g->world < 1 || g->world > B200_MAX_RANKS || g->rank < 0 || g->rank >= g->world || g->slot < 0 || g->slot >= 1024 ||
        g->wait_slot >= 1024 || !g->state)
        return false;
    for (int r = 0; r < g->world; r++)
        if (!g->peer_dst[r]) return false;
    return true;
}

int b200_mul_mat_gather(b200_ctx *ctx, const b200_mul_mat_args *a, const b200_gather *g) {
    B200_REQUIRE(ctx, ctx && a && gather_ok(g), B200_ERR_INVALID);
    B200_REQUIRE(ctx, quant_type_ok(a->type), B200_ERR_UNSUPPORTED);
    // decode only: one column, 2-D weights; shapes the streaming kernel takes (k % 256 == 0, k <= 32768)
    B200_REQUIRE(ctx, a->ne11 == 1 && a->ne12 == 1 && a->ne13 == 1 && a->ne02 == 1 && a->ne03 == 1, B200_ERR_UNSUPPORTED);
    B200_REQUIRE(ctx, a->ne00 > 0 && a->ne00 % 256 == 0 && a->ne00 <= 32768 && a->ne01 > 0, B200_ERR_UNSUPPORTED);
    const int64_t nb = a->ne00 / B200_QK;
    B200_REQUIRE(ctx, a->src0_block_off >= 0 && a->src0_block_off + nb * a->ne01 <= a->src0_nblocks_total, B200_ERR_INVALID);
    B200_REQUIRE(ctx, ((uintptr_t)a->src1_dev & 15) == 0, B200_ERR_UNSUPPORTED);
    B200_CUDA_TRY(ctx, b200_use_device(ctx->device));
    const int qsb = b200_qs_bytes(a->type);
    b200_gemv_params p;
    memset(&p, 0, sizeof(p));
    p.type = a->type;
    p.qs = (const uint8_t *)a->src0_dev + a->src0_block_off * qsb;
    p.d = (const __half *)((const uint8_t *)a->src0_dev + a->src0_nblocks_total * qsb) + a->src0_block_off;
    p.k = a->ne00; p.m = a->ne01; p.ne02 = 1; p.ne03 = 1;
    p.x = a->src1_dev;
    p.n = 1; p.ne12 = 1; p.ne13 = 1;
    p.nb11 = (size_t)a->ne00 * 4; p.nb12 = p.nb11; p.nb13 = p.nb11;
    p.dst = NULL;   // results go out as LL elements through gather->peer_dst
    p.dst_n = 1;
    p.gather = g;
    int rc = B200_OK;
    if (!b200_try_launch_gemv_stream(ctx, p, &rc)) {
        b200_set_error(ctx, "b200_mul_mat_gather: shape not served by the streaming GEMV");
        return B200_ERR_UNSUPPORTED;
    }
    return rc;
}

int b200_mul_mat_gather_batch(b200_ctx *ctx, const b200_mul_mat_args *args, const b200_gather *gathers, int count) {
    B200_REQUIRE(ctx, ctx && args && gathers && count >= 1, B200_ERR_INVALID);
    int i = 0;
    while (i < count) {
        b200_gemv_params ps[4];
        int run = 0;
        while (run < 4 && i + run < count && gather_ok(&gathers[i + run]) && decode_params(ctx, &args[i + run], &ps[run]) &&
               ps[run].k % 256 == 0 && ps[run].type == ps[0].type && ps[run].k == ps[0].k && ps[run].x == ps[0].x) {
            ps[run].gather = &gathers[i + run];
            ps[run].dst = NULL;
            run++;
        }
        int rc = B200_OK;
        if (run >= 2) {
            B200_CUDA_TRY(ctx, b200_use_device(ctx->device));
            if (b200_try_launch_gemv_stream_batch(ctx, ps, run, &rc)) {
                if (rc != B200_OK) return rc;
                i += run;
                continue;
            }
        }
        rc = b200_mul_mat_gather(ctx, &args[i], &gathers[i]);
        if (rc != B200_OK) return rc;
        i++;
    }
    return B200_OK;
}

int b200_gather_finish(b200_ctx *ctx, const b200_gather *g, const void *ll_src_dev, float *dense_out_dev, int64_t count) {
    B200_REQUIRE(ctx, ctx && gather_ok(g) && g->wait_slot >= 0 && ll_src_dev && dense_out_dev && count > 0, B200_ERR_INVALID);
    B200_CUDA_TRY(ctx, b200_use_device(ctx->device));
    return b200_launch_gather_finish(ctx, *g, ll_src_dev, dense_out_dev, count);
}

int b200_block_dots(b200_ctx *ctx, int type, const void *src0_dev, int64_t k, int64_t m, const float *src1_dev, int64_t n,
                    int32_t *out_dev, int path) {
    B200_REQUIRE(ctx, ctx && quant_type_ok(type), B200_ERR_INVALID);
    B200_REQUIRE(ctx, k > 0 && k % B200_QK == 0 && m > 0 && n > 0, B200_ERR_INVALID);
    B200_CUDA_TRY(ctx, b200_use_device(ctx->device));
    const int64_t nb = k / 32;
    const int qsb = b200_qs_bytes(type);
    const uint8_t *qs = (const uint8_t *)src0_dev;
    const __half *d = (const __half *)((const uint8_t *)src0_dev + m * nb * qsb);
    if (path == 0) {
        for (int64_t c0 = 0; c0 < n; c0 += 8) {
            b200_gemv_params p;
            memset(&p, 0, sizeof(p));
            p.type = type; p.qs = qs; p.d = d; p.k = k; p.m = m; p.ne02 = 1; p.ne03 = 1;
            p.x = src1_dev + c0 * k;
            p.n = (n - c0) < 8 ? (n - c0) : 8;
            p.ne12 = 1; p.ne13 = 1;
            p.nb11 = (size_t)k * 4; p.nb12 = p.nb11 * n; p.nb13 = p.nb12;
            p.dst = NULL;
            p.dst_n = n;
            p.dots = out_dev + c0 * m * nb;
            int rc = b200_launch_gemv(ctx, p);
            if (rc != B200_OK) return rc;
        }
        return B200_OK;
    }
    B200_REQUIRE(ctx, b200_gemm_available(), B200_ERR_UNSUPPORTED);
    const size_t q_bytes = b200_align_up((size_t)n * k, 256);
    const size_t d_bytes = b200_align_up((size_t)n * nb * 2, 256);
    int rc = b200_ws_reserve(ctx, b200_prefill_ws_bytes(type, k, m, n));
    if (rc != B200_OK) return rc;
    int8_t *aq = (int8_t *)ctx->ws;
    uint16_t *ad = (uint16_t *)((uint8_t *)ctx->ws + q_bytes);
    rc = b200_quantize_q8_0(ctx, src1_dev, k, n, (size_t)k * 4, aq, ad);
    if (rc != B200_OK) return rc;
    b200_gemm_params g;
    memset(&g, 0, sizeof(g));
    g.type = type; g.qs = qs; g.d = d; g.k = k; g.m = m; g.aq = aq; g.ad = reinterpret_cast<const __half *>(ad); g.n = n;
    g.dst = NULL;
    g.dots = out_dev;
    g.scratch = (uint8_t *)ctx->ws + q_bytes + d_bytes;
    return b200_launch_gemm(ctx, g);
}

int b200_mul_mat_host(b200_ctx *ctx, int type, const void *src0_dev, int64_t k, int64_t m, const float *src1_host, int64_t n,
                      float *dst_host) {
    B200_REQUIRE(ctx, ctx && quant_type_ok(type), B200_ERR_INVALID);
    B200_REQUIRE(ctx, k > 0 && k % B200_QK == 0 && m > 0 && n > 0, B200_ERR_INVALID);
    B200_CUDA_TRY(ctx, b200_use_device(ctx->device));
    // host-visible io scratch lives behind the activation scratch (the GEMM path grows ws itself, so use stage)
    const size_t x_bytes = b200_align_up((size_t)n * k * 4, 256), y_bytes = b200_align_up((size_t)n * m * 4, 256);
    int rc = b200_stage_reserve(ctx, x_bytes + y_bytes);
    if (rc != B200_OK) return rc;
    float *x_dev = (float *)ctx->stage;
    float *y_dev = (float *)((uint8_t *)ctx->stage + x_bytes);
    rc = b200_upload_async(ctx, x_dev, src1_host, (size_t)n * k * 4);
    if (rc != B200_OK) return rc;
    b200_mul_mat_args a;
    memset(&a, 0, sizeof(a));
    a.type = type;
    a.src0_dev = src0_dev;
    a.src0_nblocks_total = m * (k / 32);
    a.ne00 = k; a.ne01 = m; a.ne02 = 1; a.ne03 = 1;
    a.src1_dev = x_dev;
    a.ne11 = n; a.ne12 = 1; a.ne13 = 1;
    a.nb11 = (size_t)k * 4; a.nb12 = a.nb11 * n; a.nb13 = a.nb12;
    a.dst_dev = y_dev;
    rc = b200_mul_mat(ctx, &a);
    if (rc != B200_OK) return rc;
    return b200_download(ctx, dst_host, y_dev, (size_t)n * m * 4);
}

}  // extern "C"
