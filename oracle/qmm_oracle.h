/*
 * qmm_oracle.h -- CPU restatement of ggml's quantized mul_mat hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (include/, ggml-imax_b200/)
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the checker.
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 * Parity status: PINNED -- tests/test_oracle_pin.py checks every function here against
 * (a) the reference itself compiled into oracle/_ref/libggml_cpu.so when that file is present
 * and (b) golden vectors under tests/golden/ that were generated from that same reference
 * build by oracle/make_golden.py (the reference ships no golden vectors of its own,
 * SURVEY.md section 8c).
 */
#ifndef QMM_ORACLE_H
#define QMM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_QK 32            /* QK4_0 == QK8_0 == 32, src/ggml-common.h:143,185 */
#define ORACLE_Q4_0_BYTES 18    /* sizeof(block_q4_0), src/ggml-common.h:144-149 */
#define ORACLE_Q8_0_BYTES 34    /* sizeof(block_q8_0), src/ggml-common.h:186-191 */

enum oracle_type { ORACLE_TYPE_Q4_0 = 2, ORACLE_TYPE_Q8_0 = 8 }; /* values of enum ggml_type, include/ggml/ggml.h:347-355 */

/* fp16 <-> fp32, IEEE round-to-nearest-even == F16C _cvtss_sh(x,0)/_cvtsh_ss, src/ggml-impl.h:452-453 */
uint16_t oracle_fp32_to_fp16(float f);
float    oracle_fp16_to_fp32(uint16_t h);

/* src/ggml-quants.c:465 + AVX2 body :535-618 (the runtime from_float of Q8_0 on x86) */
void oracle_quantize_row_q8_0(const float *x, void *y, int64_t k);
/* src/ggml-quants.c:440-463 (scalar "_reference": id = 1/d, roundf) -- used by ggml_quantize_chunk for weights */
void oracle_quantize_row_q8_0_reference(const float *x, void *y, int64_t k);
/* src/ggml-quants.c:260-295 */
void oracle_quantize_row_q4_0_reference(const float *x, void *y, int64_t k);
/* src/ggml-quants.c:980-998, :1074-1088 */
void oracle_dequantize_row_q4_0(const void *x, float *y, int64_t k);
void oracle_dequantize_row_q8_0(const void *x, float *y, int64_t k);

/* per-block int32 partial sums: the integer inner loop of src/ggml-quants.c:3858-3869 and :5010-5015.
 * out[i] = sum_j (nib_j - 8) * q8_j   resp.  sum_j x_j * y_j   for block i < k/32. */
void oracle_block_dots_q4_0_q8_0(int64_t k, const void *x_q4_0, const void *y_q8_0, int32_t *out);
void oracle_block_dots_q8_0_q8_0(int64_t k, const void *x_q8_0, const void *y_q8_0, int32_t *out);

/* vec_dot, scalar summation order: src/ggml-quants.c:3855-3872 / :5006-5021 */
float oracle_vec_dot_q4_0_q8_0_scalar(int64_t k, const void *x, const void *y);
float oracle_vec_dot_q8_0_q8_0_scalar(int64_t k, const void *x, const void *y);
/* vec_dot, AVX2 summation order (8 float lanes + fmadd + hsum_float_8):
 * src/ggml-quants.c:3600-3623 / :4925-4946 with helpers :43-49, :99-124 */
float oracle_vec_dot_q4_0_q8_0_avx2order(int64_t k, const void *x, const void *y);
float oracle_vec_dot_q8_0_q8_0_avx2order(int64_t k, const void *x, const void *y);

/*
 * ggml_compute_forward_mul_mat, src/ggml.c:11808-12097 (INIT :11952-11974 + COMPUTE :12056-12096).
 *   src0: quantized weights [ne00=k, ne01=m, ne02, ne03] contiguous rows of blocks
 *   src1: F32 [k, n=ne11, ne12, ne13], byte strides nb11/nb12/nb13 (nb10 == 4)
 *   dst : F32 [m, n, ne12, ne13] dense
 * Broadcast i02 = i12/(ne12/ne02), i03 = i13/(ne13/ne03) (:11848-11849, :12063-12065).
 * avx2_order selects the fp32 summation order of vec_dot (1 == what the x86 reference build runs).
 * Returns 0, or -1 on a shape the reference would assert on.
 */
int oracle_mul_mat(int type, const void *src0, int64_t ne00, int64_t ne01, int64_t ne02, int64_t ne03,
                   const float *src1, int64_t ne11, int64_t ne12, int64_t ne13,
                   size_t nb11, size_t nb12, size_t nb13,
                   float *dst, int avx2_order);

/* multi-threaded variant used only as the CPU baseline timer (rows of src0 split over nthreads). */
int oracle_mul_mat_mt(int type, const void *src0, int64_t ne00, int64_t ne01,
                      const float *src1, int64_t ne11, float *dst, void *wdata, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
