/*
 * qmm_oracle.c -- plain-C CPU restatement of ggml's Q4_0/Q8_0 mul_mat path.
 * TEST INFRASTRUCTURE ONLY (see qmm_oracle.h).  Written from the algorithm, not copied:
 * scalar C, no intrinsics; compile with -O2 -ffp-contract=off (IEEE, no fused contraction
 * except where fmaf() is written explicitly).
 */
#include "qmm_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- fp16 ------------------------------------------------------------------------- */

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float    u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* IEEE binary32 -> binary16, round to nearest even; what _cvtss_sh(x, 0) does
 * (src/ggml-impl.h:453). */
uint16_t oracle_fp32_to_fp16(float f) {
    const uint32_t u = f2u(f);
    const uint32_t sign = (u >> 16) & 0x8000u;
    const uint32_t absu = u & 0x7fffffffu;
    if (absu >= 0x7f800000u) {                       /* inf / nan */
        return (uint16_t)(sign | 0x7c00u | (absu > 0x7f800000u ? (0x0200u | ((absu >> 13) & 0x3ffu)) : 0));
    }
    if (absu >= 0x477ff000u) {                       /* >= 65520 rounds to inf */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (absu < 0x38800000u) {                        /* result is subnormal (or zero): |f| < 2^-14 */
        if (absu < 0x33000000u) return (uint16_t)sign;            /* < 2^-25 -> 0 (2^-25 itself ties to even 0) */
        const int e = (int)(absu >> 23);                           /* biased exponent, 102..112 */
        const uint32_t m = (absu & 0x7fffffu) | 0x800000u;         /* 24-bit significand */
        const int shift = 126 - e;                                 /* 14..24: value = m * 2^(e-150); half subnormal unit 2^-24 */
        const uint32_t q = m >> shift;
        const uint32_t rem = m & ((1u << shift) - 1u);
        const uint32_t half = 1u << (shift - 1);
        uint32_t r = q;
        if (rem > half || (rem == half && (q & 1u))) r++;
        return (uint16_t)(sign | r);
    }
    {
        const uint32_t e = (absu >> 23) - 112u;                    /* half biased exponent 1..30 */
        const uint32_t m = absu & 0x7fffffu;
        uint32_t r = (e << 10) | (m >> 13);
        const uint32_t rem = m & 0x1fffu;
        if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) r++;   /* carry may bump the exponent: correct */
        return (uint16_t)(sign | r);
    }
}

float oracle_fp16_to_fp32(uint16_t h) {
    const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    const uint32_t e = (h >> 10) & 0x1fu;
    const uint32_t m = h & 0x3ffu;
    if (e == 0) {
        if (m == 0) return u2f(sign);
        /* subnormal: m * 2^-24 */
        const float v = (float)m * 5.9604644775390625e-8f;
        return sign ? -v : v;
    }
    if (e == 31) return u2f(sign | 0x7f800000u | (m << 13));
    return u2f(sign | ((e + 112u) << 23) | (m << 13));
}

/* ---- block access (wire format, src/ggml-common.h:144-149, :186-191) ---------------- */

static inline uint16_t q4_d(const uint8_t *b)  { uint16_t d; memcpy(&d, b, 2); return d; }
static inline const uint8_t *q4_qs(const uint8_t *b) { return b + 2; }
static inline uint16_t q8_d(const uint8_t *b)  { uint16_t d; memcpy(&d, b, 2); return d; }
static inline const int8_t *q8_qs(const uint8_t *b) { return (const int8_t *)(b + 2); }

/* ---- quantize ---------------------------------------------------------------------- */

/* AVX2 branch of quantize_row_q8_0 (src/ggml-quants.c:535-618): d = amax/127 -> fp16;
 * id = amax ? 127/amax : 0; q = cvtps_epi32(round_nearest(x*id)) i.e. ties-to-even. */
void oracle_quantize_row_q8_0(const float *x, void *vy, int64_t k) {
    uint8_t *y = (uint8_t *)vy;
    const int64_t nb = k / ORACLE_QK;
    for (int64_t i = 0; i < nb; i++) {
        const float *xb = x + i * ORACLE_QK;
        float amax = 0.0f;
        for (int j = 0; j < ORACLE_QK; j++) {
            const float a = fabsf(xb[j]);
            if (a > amax) amax = a;                 /* _mm256_max_ps on |x| */
        }
        const float d = amax / 127.f;
        const uint16_t dh = oracle_fp32_to_fp16(d);
        const float id = (amax != 0.0f) ? 127.f / amax : 0.0f;
        uint8_t *blk = y + i * ORACLE_Q8_0_BYTES;
        memcpy(blk, &dh, 2);
        for (int j = 0; j < ORACLE_QK; j++) {
            const float v = xb[j] * id;
            /* nearbyintf under the default rounding mode == round-half-to-even */
            ((int8_t *)(blk + 2))[j] = (int8_t)(int)nearbyintf(v);
        }
    }
}

/* src/ggml-quants.c:440-463 */
void oracle_quantize_row_q8_0_reference(const float *x, void *vy, int64_t k) {
    uint8_t *y = (uint8_t *)vy;
    const int64_t nb = k / ORACLE_QK;
    for (int64_t i = 0; i < nb; i++) {
        const float *xb = x + i * ORACLE_QK;
        float amax = 0.0f;
        for (int j = 0; j < ORACLE_QK; j++) {
            const float a = fabsf(xb[j]);
            amax = amax > a ? amax : a;
        }
        const float d = amax / 127;
        const float id = d ? 1.0f / d : 0.0f;
        const uint16_t dh = oracle_fp32_to_fp16(d);
        uint8_t *blk = y + i * ORACLE_Q8_0_BYTES;
        memcpy(blk, &dh, 2);
        for (int j = 0; j < ORACLE_QK; j++) {
            ((int8_t *)(blk + 2))[j] = (int8_t)roundf(xb[j] * id);
        }
    }
}

/* src/ggml-quants.c:260-295 */
void oracle_quantize_row_q4_0_reference(const float *x, void *vy, int64_t k) {
    uint8_t *y = (uint8_t *)vy;
    const int64_t nb = k / ORACLE_QK;
    for (int64_t i = 0; i < nb; i++) {
        const float *xb = x + i * ORACLE_QK;
        float amax = 0.0f, max = 0.0f;
        for (int j = 0; j < ORACLE_QK; j++) {
            const float v = xb[j];
            if (amax < fabsf(v)) { amax = fabsf(v); max = v; }
        }
        const float d = max / -8;
        const float id = d ? 1.0f / d : 0.0f;
        const uint16_t dh = oracle_fp32_to_fp16(d);
        uint8_t *blk = y + i * ORACLE_Q4_0_BYTES;
        memcpy(blk, &dh, 2);
        for (int j = 0; j < ORACLE_QK / 2; j++) {
            const float x0 = xb[j] * id;
            const float x1 = xb[ORACLE_QK / 2 + j] * id;
            int a0 = (int8_t)(x0 + 8.5f); if (a0 > 15) a0 = 15;
            int a1 = (int8_t)(x1 + 8.5f); if (a1 > 15) a1 = 15;
            blk[2 + j] = (uint8_t)((uint8_t)a0 | ((uint8_t)a1 << 4));
        }
    }
}

/* src/ggml-quants.c:980-998 */
void oracle_dequantize_row_q4_0(const void *vx, float *y, int64_t k) {
    const uint8_t *x = (const uint8_t *)vx;
    const int64_t nb = k / ORACLE_QK;
    for (int64_t i = 0; i < nb; i++) {
        const uint8_t *blk = x + i * ORACLE_Q4_0_BYTES;
        const float d = oracle_fp16_to_fp32(q4_d(blk));
        for (int j = 0; j < ORACLE_QK / 2; j++) {
            const int x0 = (q4_qs(blk)[j] & 0x0F) - 8;
            const int x1 = (q4_qs(blk)[j] >> 4) - 8;
            y[i * ORACLE_QK + j] = x0 * d;
            y[i * ORACLE_QK + j + ORACLE_QK / 2] = x1 * d;
        }
    }
}

/* src/ggml-quants.c:1074-1088 */
void oracle_dequantize_row_q8_0(const void *vx, float *y, int64_t k) {
    const uint8_t *x = (const uint8_t *)vx;
    const int64_t nb = k / ORACLE_QK;
    for (int64_t i = 0; i < nb; i++) {
        const uint8_t *blk = x + i * ORACLE_Q8_0_BYTES;
        const float d = oracle_fp16_to_fp32(q8_d(blk));
        for (int j = 0; j < ORACLE_QK; j++) y[i * ORACLE_QK + j] = q8_qs(blk)[j] * d;
    }
}

/* ---- block dots -------------------------------------------------------------------- */

/* integer part of src/ggml-quants.c:3858-3869 */
static inline int32_t blockdot_q4(const uint8_t *xb, const uint8_t *yb) {
    int32_t sumi = 0;
    const uint8_t *qx = q4_qs(xb);
    const int8_t *qy = q8_qs(yb);
    for (int j = 0; j < ORACLE_QK / 2; j++) {
        const int v0 = (qx[j] & 0x0F) - 8;
        const int v1 = (qx[j] >> 4) - 8;
        sumi += v0 * qy[j] + v1 * qy[j + ORACLE_QK / 2];
    }
    return sumi;
}

/* integer part of src/ggml-quants.c:5010-5015 */
static inline int32_t blockdot_q8(const uint8_t *xb, const uint8_t *yb) {
    int32_t sumi = 0;
    const int8_t *qx = q8_qs(xb);
    const int8_t *qy = q8_qs(yb);
    for (int j = 0; j < ORACLE_QK; j++) sumi += qx[j] * qy[j];
    return sumi;
}

void oracle_block_dots_q4_0_q8_0(int64_t k, const void *vx, const void *vy, int32_t *out) {
    const uint8_t *x = (const uint8_t *)vx, *y = (const uint8_t *)vy;
    for (int64_t i = 0; i < k / ORACLE_QK; i++)
        out[i] = blockdot_q4(x + i * ORACLE_Q4_0_BYTES, y + i * ORACLE_Q8_0_BYTES);
}

void oracle_block_dots_q8_0_q8_0(int64_t k, const void *vx, const void *vy, int32_t *out) {
    const uint8_t *x = (const uint8_t *)vx, *y = (const uint8_t *)vy;
    for (int64_t i = 0; i < k / ORACLE_QK; i++)
        out[i] = blockdot_q8(x + i * ORACLE_Q8_0_BYTES, y + i * ORACLE_Q8_0_BYTES);
}

/* ---- vec_dot ----------------------------------------------------------------------- */

/* src/ggml-quants.c:3855-3872: sumf += sumi * d_x * d_y, left-associated, float. */
float oracle_vec_dot_q4_0_q8_0_scalar(int64_t k, const void *vx, const void *vy) {
    const uint8_t *x = (const uint8_t *)vx, *y = (const uint8_t *)vy;
    float sumf = 0.0f;
    for (int64_t i = 0; i < k / ORACLE_QK; i++) {
        const uint8_t *xb = x + i * ORACLE_Q4_0_BYTES, *yb = y + i * ORACLE_Q8_0_BYTES;
        const int sumi = blockdot_q4(xb, yb);
        sumf += sumi * oracle_fp16_to_fp32(q4_d(xb)) * oracle_fp16_to_fp32(q8_d(yb));
    }
    return sumf;
}

/* src/ggml-quants.c:5006-5021: sumf += sumi * (d_x * d_y) */
float oracle_vec_dot_q8_0_q8_0_scalar(int64_t k, const void *vx, const void *vy) {
    const uint8_t *x = (const uint8_t *)vx, *y = (const uint8_t *)vy;
    float sumf = 0.0f;
    for (int64_t i = 0; i < k / ORACLE_QK; i++) {
        const uint8_t *xb = x + i * ORACLE_Q8_0_BYTES, *yb = y + i * ORACLE_Q8_0_BYTES;
        const int sumi = blockdot_q8(xb, yb);
        sumf += sumi * (oracle_fp16_to_fp32(q8_d(xb)) * oracle_fp16_to_fp32(q8_d(yb)));
    }
    return sumf;
}

/* hsum_float_8, src/ggml-quants.c:43-49: ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7)) */
static inline float hsum8(const float a[8]) {
    const float r0 = a[4] + a[0], r1 = a[5] + a[1], r2 = a[6] + a[2], r3 = a[7] + a[3];
    const float s0 = r0 + r2, s1 = r1 + r3;
    return s0 + s1;
}

/* AVX2 body src/ggml-quants.c:3600-3623.  The 32 expanded bytes sit as [low nibbles 0..15 | high
 * nibbles 0..15] (bytes_from_nibbles_32 :99-105) which is element order 0..31; maddubs+madd leave
 * eight int32 lanes, lane l = sum of products of elements 4l..4l+3 (:107-124); each lane is
 * converted to float and fused-multiply-added with d = d_x*d_y. */
float oracle_vec_dot_q4_0_q8_0_avx2order(int64_t k, const void *vx, const void *vy) {
    const uint8_t *x = (const uint8_t *)vx, *y = (const uint8_t *)vy;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = 0; i < k / ORACLE_QK; i++) {
        const uint8_t *xb = x + i * ORACLE_Q4_0_BYTES, *yb = y + i * ORACLE_Q8_0_BYTES;
        const float d = oracle_fp16_to_fp32(q4_d(xb)) * oracle_fp16_to_fp32(q8_d(yb));
        const uint8_t *qx = q4_qs(xb);
        const int8_t *qy = q8_qs(yb);
        for (int l = 0; l < 8; l++) {
            int32_t s = 0;
            for (int e = 4 * l; e < 4 * l + 4; e++) {
                const int w = (e < 16) ? ((qx[e] & 0x0F) - 8) : ((qx[e - 16] >> 4) - 8);
                s += w * qy[e];
            }
            acc[l] = fmaf(d, (float)s, acc[l]);
        }
    }
    return hsum8(acc);
}

/* AVX2 body src/ggml-quants.c:4925-4946 */
float oracle_vec_dot_q8_0_q8_0_avx2order(int64_t k, const void *vx, const void *vy) {
    const uint8_t *x = (const uint8_t *)vx, *y = (const uint8_t *)vy;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = 0; i < k / ORACLE_QK; i++) {
        const uint8_t *xb = x + i * ORACLE_Q8_0_BYTES, *yb = y + i * ORACLE_Q8_0_BYTES;
        const float d = oracle_fp16_to_fp32(q8_d(xb)) * oracle_fp16_to_fp32(q8_d(yb));
        const int8_t *qx = q8_qs(xb), *qy = q8_qs(yb);
        for (int l = 0; l < 8; l++) {
            int32_t s = 0;
            for (int e = 4 * l; e < 4 * l + 4; e++) s += qx[e] * qy[e];
            acc[l] = fmaf(d, (float)s, acc[l]);
        }
    }
    return hsum8(acc);
}

/* ---- mul_mat driver ---------------------------------------------------------------- */

static size_t row_bytes(int type, int64_t k) {
    return (size_t)(k / ORACLE_QK) * (type == ORACLE_TYPE_Q4_0 ? ORACLE_Q4_0_BYTES : ORACLE_Q8_0_BYTES);
}

typedef float (*vecdot_fn)(int64_t, const void *, const void *);

static vecdot_fn pick_vecdot(int type, int avx2_order) {
    if (type == ORACLE_TYPE_Q4_0) return avx2_order ? oracle_vec_dot_q4_0_q8_0_avx2order : oracle_vec_dot_q4_0_q8_0_scalar;
    return avx2_order ? oracle_vec_dot_q8_0_q8_0_avx2order : oracle_vec_dot_q8_0_q8_0_scalar;
}

int oracle_mul_mat(int type, const void *src0, int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
                   const float *src1, int64_t ne11, int64_t ne12, int64_t ne13,
                   size_t nb11, size_t nb12, size_t nb13,
                   float *dst, int avx2_order) {
    if (type != ORACLE_TYPE_Q4_0 && type != ORACLE_TYPE_Q8_0) return -1;
    if (ne00 % ORACLE_QK != 0) return -1;                       /* assert(n % qk == 0), ggml-quants.c:3473 */
    if (ne02 <= 0 || ne03 <= 0 || ne12 % ne02 != 0 || ne13 % ne03 != 0) return -1; /* ggml.c:2714-2720 */
    const vecdot_fn vecdot = pick_vecdot(type, avx2_order);
    const size_t rs0 = row_bytes(type, ne00);                   /* nb01 of a contiguous src0 */
    const size_t rs1 = row_bytes(ORACLE_TYPE_Q8_0, ne00);       /* ggml_row_size(vec_dot_type, ne10) */
    const int64_t nrows1 = ne11 * ne12 * ne13;
    uint8_t *wdata = (uint8_t *)malloc(rs1 * (size_t)(nrows1 > 0 ? nrows1 : 1));
    if (!wdata) return -1;
    /* INIT: quantize every src1 row, contiguous [ne13][ne12][ne11] (ggml.c:11956-11971) */
    {
        uint8_t *w = wdata;
        for (int64_t i13 = 0; i13 < ne13; i13++)
            for (int64_t i12 = 0; i12 < ne12; i12++)
                for (int64_t i11 = 0; i11 < ne11; i11++) {
                    const float *row = (const float *)((const char *)src1 + i13 * nb13 + i12 * nb12 + i11 * nb11);
                    oracle_quantize_row_q8_0(row, w, ne00);
                    w += rs1;
                }
    }
    /* COMPUTE (ggml.c:12056-12096); tiling order does not change any result, so plain loops */
    const int64_t r2 = ne12 / ne02, r3 = ne13 / ne03;
    for (int64_t i13 = 0; i13 < ne13; i13++)
        for (int64_t i12 = 0; i12 < ne12; i12++)
            for (int64_t i11 = 0; i11 < ne11; i11++) {
                const int64_t i03 = i13 / r3, i02 = i12 / r2;
                const uint8_t *w0 = (const uint8_t *)src0 + (size_t)(i02 + i03 * ne02) * (size_t)ne01 * rs0;
                const uint8_t *col = wdata + (size_t)(i11 + i12 * ne11 + i13 * ne12 * ne11) * rs1;
                float *dcol = dst + (size_t)(i11 + i12 * ne11 + i13 * ne12 * ne11) * (size_t)ne01;
                for (int64_t ir0 = 0; ir0 < ne01; ir0++) dcol[ir0] = vecdot(ne00, w0 + (size_t)ir0 * rs0, col);
            }
    free(wdata);
    return 0;
}

/* ---- multi-threaded timing variant (CPU baseline "port") ---------------------------- */

struct mt_job {
    int type; const uint8_t *src0; int64_t k, m, n; const uint8_t *wdata; float *dst;
    int64_t r0, r1;
};

static void *mt_worker(void *arg) {
    struct mt_job *j = (struct mt_job *)arg;
    const vecdot_fn vecdot = pick_vecdot(j->type, 1);
    const size_t rs0 = row_bytes(j->type, j->k), rs1 = row_bytes(ORACLE_TYPE_Q8_0, j->k);
    for (int64_t c = 0; c < j->n; c++)
        for (int64_t r = j->r0; r < j->r1; r++)
            j->dst[c * j->m + r] = vecdot(j->k, j->src0 + (size_t)r * rs0, j->wdata + (size_t)c * rs1);
    return NULL;
}

int oracle_mul_mat_mt(int type, const void *src0, int64_t ne00, int64_t ne01,
                      const float *src1, int64_t ne11, float *dst, void *wdata, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    const size_t rs1 = row_bytes(ORACLE_TYPE_Q8_0, ne00);
    for (int64_t c = 0; c < ne11; c++)                           /* serial INIT like the reference (ggml.c:11953) */
        oracle_quantize_row_q8_0(src1 + c * ne00, (uint8_t *)wdata + (size_t)c * rs1, ne00);
    pthread_t th[256];
    struct mt_job jobs[256];
    const int64_t dr = (ne01 + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        int64_t r0 = dr * t, r1 = r0 + dr; if (r1 > ne01) r1 = ne01; if (r0 > ne01) r0 = ne01;
        jobs[t] = (struct mt_job){type, (const uint8_t *)src0, ne00, ne01, ne11, (const uint8_t *)wdata, dst, r0, r1};
        if (t > 0) pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
    }
    mt_worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}
