/*
 * ggml_b200.h -- thin C ABI of the B200 (sm_100a) quantized mul_mat path.
 *
 * This is the drop-in boundary between host code written in C (the ggml backend in
 * ggml-imax_b200/host/ggml-b200.c, or any FFI: ctypes, cgo, JNI ...) and the CUDA kernels in
 * ggml-imax_b200/csrc/.  Plain pointers and sizes only; no C++/torch types.  Every entry point
 * names the reference interface it stands in for (paths relative to the reference checkout).
 *
 * There is NO CPU fallback behind any of these calls: without a usable sm_100 device they return
 * B200_ERR_CUDA / B200_ERR_UNSUPPORTED and leave the reason in b200_last_error().
 *
 * All *_dev pointers are device pointers obtained from b200_malloc() (or any CUDA allocation on
 * the context's device).  Unless stated otherwise calls are asynchronous on the context's stream;
 * b200_upload/b200_download/b200_set_quantized/b200_get_quantized are synchronous like the
 * reference's buffer set_tensor/get_tensor (src/ggml-backend.c:221-247).
 */
#ifndef GGML_B200_H
#define GGML_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#  define B200_API __attribute__((visibility("default")))
#else
#  define B200_API
#endif

/* status codes */
#define B200_OK               0
#define B200_ERR_CUDA        (-1)   /* a CUDA runtime/driver call failed (see b200_last_error) */
#define B200_ERR_INVALID     (-2)   /* bad argument / shape the reference would GGML_ASSERT on  */
#define B200_ERR_UNSUPPORTED (-3)   /* valid ggml, but outside this path (supports_op == false)   */
#define B200_ERR_ALLOC       (-4)   /* device allocation failed (GGML_STATUS_ALLOC_FAILED)        */

/* tensor types: numeric values of enum ggml_type (include/ggml/ggml.h:347-355) */
#define B200_TYPE_F32   0
#define B200_TYPE_F16   1
#define B200_TYPE_Q4_0  2
#define B200_TYPE_Q5_0  6    /* sibling 32-element formats that share the Q8_0 activation path: kept in WIRE format on the device, */
#define B200_TYPE_Q8_0  8
#define B200_TYPE_IQ4_NL 20  /* served by a plain correctness path (b200_wire_formats.cu), not by the streaming / tensor-core kernels */
#define B200_TYPE_I16   25
#define B200_TYPE_I32   26

/* wire formats (src/ggml-common.h:144-149, :186-191) */
#define B200_QK            32
#define B200_Q4_0_BYTES    18
#define B200_Q8_0_BYTES    34

typedef struct b200_ctx b200_ctx;   /* one per (device, stream); not thread-safe, like a ggml_backend */

/* ---- devices / contexts ------------------------------------------------------------------
 * replaces ggml_backend_cuda_get_device_count / _get_device_description / _get_device_memory
 * (src/ggml-cuda.h:33-36) and the per-backend context of ggml_backend_cuda_init (src/ggml-cuda.h:19) */
B200_API int  b200_device_count(void);
B200_API int  b200_device_info(int device, char *name, size_t name_len, size_t *free_bytes, size_t *total_bytes,
                               int *sm_count, int *cc_major, int *cc_minor);
B200_API int  b200_ctx_create(int device, b200_ctx **out);
/* same, but launches on a stream owned by the caller (e.g. torch's current stream, for plumbing) */
B200_API int  b200_ctx_create_on_stream(int device, void *cuda_stream, b200_ctx **out);
B200_API void b200_ctx_destroy(b200_ctx *ctx);
B200_API const char *b200_last_error(const b200_ctx *ctx);   /* ctx may be NULL: last global error */
B200_API void *b200_ctx_stream(const b200_ctx *ctx);         /* the cudaStream_t launches go to     */
B200_API int  b200_ctx_device(const b200_ctx *ctx);
/* knobs (the library reads NO environment variables; every knob is a per-context option with a measured default):
 *   "pdl" (0/1 programmatic dependent launch on the decode GEMV), "gemm" (0 = never use the tcgen05 GEMM, 1 = auto),
 *   "gemv_max_n" (largest n served by the GEMV), "gemv_stream" (0 = generic GEMV only),
 *   "gemm_exact" (1 = prefill GEMM keeps exact int32 block dots + fp32 scaling; 0 = fp16 tensor-core path, default),
 *   decode plans created afterwards: "plan_pub_min_k" / "plan_pub_dist" (an in-plan src1 of at least that many elements whose
 *   producer lies at least that many ops back is quantized once per GPU; 0 = never), "plan_l2_window" (ring slots the L2
 *   prefetcher runs ahead of the weight stream; 0 = off), "plan_slots" (ring slots; 0 = as many as fit), "plan_trace"
 *   (record a device-side timeline, see b200_plan_trace), "plan_timeout_ms" (bound of every in-kernel wait; 0 = default:
 *   10 s on one GPU, 120 s for row-split plans).
 * returns B200_ERR_INVALID if unknown */
B200_API int  b200_ctx_set_option(b200_ctx *ctx, const char *key, int64_t value);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
B200_API int64_t b200_ctx_launch_count(const b200_ctx *ctx);

/* device-side timeline for profiling the decode kernels (the reference has only host-side GGML_PERF counters,
 * src/ggml.c:19198-19205): trace_dev = room for max_launches * 160 * 8 uint64 (%globaltimer ns stamps per CTA:
 * 0 entry, 1 ring primed, 2 predecessor done (griddepcontrol.wait), 3 activations quantized, 4 first weights
 * landed, 5 last row done).  NULL switches it off.  Launch i of the context after this call uses slot i. */
B200_API int  b200_ctx_set_trace(b200_ctx *ctx, void *trace_dev, int64_t max_launches);

/* ---- device buffers -----------------------------------------------------------------------
 * replaces ggml_backend_buffer_type_i.alloc_buffer and ggml_backend_buffer_i.{free_buffer,clear,
 * set_tensor,get_tensor,cpy_tensor} for non-quantized data (src/ggml-backend-impl.h:18-48) */
B200_API int b200_malloc(b200_ctx *ctx, void **dptr, size_t size);
B200_API int b200_free(b200_ctx *ctx, void *dptr);
B200_API int b200_memset(b200_ctx *ctx, void *dptr, int value, size_t size);            /* sync */
B200_API int b200_upload(b200_ctx *ctx, void *dst_dev, const void *src_host, size_t size);   /* sync */
B200_API int b200_download(b200_ctx *ctx, void *dst_host, const void *src_dev, size_t size); /* sync */
B200_API int b200_upload_async(b200_ctx *ctx, void *dst_dev, const void *src_host, size_t size);
B200_API int b200_download_async(b200_ctx *ctx, void *dst_host, const void *src_dev, size_t size);
B200_API int b200_copy_d2d(b200_ctx *ctx, void *dst_dev, const void *src_dev, size_t size);  /* async; the two may live on different devices */
/* `height` rows of `width` bytes, rows dpitch / spitch bytes apart (a dst slice of a row-split mul_mat with n > 1); async */
B200_API int b200_copy_2d(b200_ctx *ctx, void *dst_dev, size_t dpitch, const void *src_dev, size_t spitch, size_t width, size_t height);
/* let kernels and copies of this context's device address memory of `peer_device` directly (NVLink / NVSwitch peer access): what
 * a single-process row split needs where the one-process-per-GPU path uses b200_ipc_* (the reference enables it the same way,
 * src/ggml-cuda.cu:1313-1351).  B200_OK if already enabled. */
B200_API int b200_enable_peer_access(b200_ctx *ctx, int peer_device);
/* ggml_backend_i.synchronize.  Also where a persistent kernel that gave up waiting (a dead peer rank, mismatched launch
 * sequences) surfaces: B200_ERR_CUDA with the reason in b200_last_error; the results of that launch are invalid. */
B200_API int b200_synchronize(b200_ctx *ctx);
/* pinned host staging (ggml_backend_cuda_host_buffer_type, src/ggml-cuda.h:31) */
/* events: ggml_backend_i.event_new / _free / _record / _wait / _synchronize (src/ggml-backend-impl.h:112-116; used by ggml_backend_sched
 * to overlap the copies of one split with the compute of another).  Record on one context's stream, wait from another context's stream
 * (same or another device) or from the host. */
typedef struct b200_event b200_event;
B200_API int  b200_event_create(b200_ctx *ctx, b200_event **out);
B200_API void b200_event_destroy(b200_event *ev);
B200_API int  b200_event_record(b200_ctx *ctx, b200_event *ev);        /* after everything enqueued on ctx's stream so far */
B200_API int  b200_event_wait(b200_ctx *ctx, b200_event *ev);          /* work enqueued on ctx's stream from now on waits for it */
B200_API int  b200_event_synchronize(b200_event *ev);                  /* the host waits */
B200_API int b200_host_malloc(void **hptr, size_t size);
B200_API int b200_host_free(void *hptr);

/* ---- launch graphs --------------------------------------------------------------------------
 * A launch-bound sequence (the whole-model decode step of a small model: GPT-2 117M issues ~170 kernels of 2-5 us) pays the host's
 * per-launch cost on every kernel.  b200_graph_begin/end bracket any sequence of the asynchronous calls of this header on the
 * context's stream and turn it into one replayable CUDA graph; b200_graph_launch replays it.  Between begin and end every call
 * is recorded instead of executed; a call that has to wait for the device (b200_synchronize, b200_download, a scratch area that
 * must grow: call b200_reserve_workspace first) makes b200_graph_end return B200_ERR_UNSUPPORTED, and nothing is kept.  The
 * recorded addresses and scalar arguments are frozen: record again when they change.  The backend's ggml_backend_graph_plan_*
 * uses this (cf. the reference's opt-in GGML_CUDA_USE_GRAPHS, src/ggml-cuda.cu:2461-2709). */
typedef struct b200_graph b200_graph;
B200_API int     b200_graph_begin(b200_ctx *ctx);
B200_API int     b200_graph_end(b200_ctx *ctx, b200_graph **out);       /* out == NULL: discard what was recorded */
B200_API int     b200_graph_launch(b200_ctx *ctx, b200_graph *graph);
B200_API int64_t b200_graph_node_count(const b200_graph *graph);         /* nodes (kernels, copies) one launch replays */
B200_API void    b200_graph_destroy(b200_graph *graph);
/* make sure the scratch of a prefill mul_mat (type, k, m, n) exists (no-op if already large enough) */
B200_API int  b200_reserve_workspace(b200_ctx *ctx, int type, int64_t k, int64_t m, int64_t n);

/* ---- quantized tensor storage ("repack once at set_tensor") --------------------------------
 * A Q4_0/Q8_0 tensor of nblocks_total blocks occupies exactly its ggml_nbytes() on the device but
 * laid out as planes:   [ qs plane: nblocks_total * 16 B (Q4_0, packed nibbles, original nibble order)
 *                                 or nblocks_total * 32 B (Q8_0, int8) ]
 *                       [ d plane : nblocks_total * 2 B  (fp16 block scales) ]
 * set = ggml_backend_buffer_i.set_tensor for a quantized tensor: src_host holds `nblocks` blocks in
 * wire format destined for blocks [block_off, block_off+nblocks).  get is the exact inverse, so
 * get(set(x)) == x byte for byte (needed by ggml_backend_graph_copy, src/ggml-backend.c:1974-2060). */
B200_API int b200_set_quantized(b200_ctx *ctx, int type, void *tensor_dev, int64_t nblocks_total,
                                const void *src_host, int64_t block_off, int64_t nblocks);
B200_API int b200_get_quantized(b200_ctx *ctx, int type, const void *tensor_dev, int64_t nblocks_total,
                                void *dst_host, int64_t block_off, int64_t nblocks);
/* same repack but from wire-format blocks already on the device (tensor copy between layouts) */
B200_API int b200_repack_from_device(b200_ctx *ctx, int type, void *tensor_dev, int64_t nblocks_total,
                                     const void *src_wire_dev, int64_t block_off, int64_t nblocks);
B200_API int b200_unrepack_to_device(b200_ctx *ctx, int type, const void *tensor_dev, int64_t nblocks_total,
                                     void *dst_wire_dev, int64_t block_off, int64_t nblocks);

/* ---- activation quantization ---------------------------------------------------------------
 * replaces quantize_row_q8_0 (src/ggml-quants.c:465, AVX2 body :535-618), bit-exact.
 * x_dev: nrows rows of k floats, rows row_stride_bytes apart.
 * planar form (what the kernels consume): qs_dev [nrows][k] int8, d_dev [nrows][k/32] fp16.
 * blocks form: nrows * (k/32) block_q8_0 in wire format (34 B), what the reference writes to wdata
 * (src/ggml.c:11956-11971). */
B200_API int b200_quantize_q8_0(b200_ctx *ctx, const float *x_dev, int64_t k, int64_t nrows, size_t row_stride_bytes,
                                int8_t *qs_dev, uint16_t *d_dev);
B200_API int b200_quantize_q8_0_blocks(b200_ctx *ctx, const float *x_dev, int64_t k, int64_t nrows,
                                       size_t row_stride_bytes, void *blocks_dev);

/* ---- mul_mat -------------------------------------------------------------------------------
 * replaces ggml_compute_forward_mul_mat (src/ggml.c:11808-12097) for src0 in {Q4_0,Q8_0} (repacked
 * storage above), src1 F32, dst F32.  Same argument meaning as the ggml tensors:
 *   src0 [ne00=k, ne01=m, ne02, ne03] contiguous;   src1 [k, ne11=n, ne12, ne13] with byte strides
 *   nb11/nb12/nb13 (nb10 == 4);   dst [m, n, ne12, ne13] dense;   ne12 % ne02 == 0, ne13 % ne03 == 0.
 * src0_nblocks_total / src0_block_off locate src0 inside its repacked tensor (0 / total for a
 * whole tensor; a row-range view uses the parent's total and its own first block). */
typedef struct b200_mul_mat_args {
    int32_t      type;                 /* B200_TYPE_Q4_0 or B200_TYPE_Q8_0 */
    int32_t      flags;                /* B200_MM_* */
    const void  *src0_dev;             /* base of the repacked tensor that holds src0 */
    int64_t      src0_nblocks_total;
    int64_t      src0_block_off;
    int64_t      ne00, ne01, ne02, ne03;
    const float *src1_dev;
    int64_t      ne11, ne12, ne13;
    size_t       nb11, nb12, nb13;
    float       *dst_dev;
} b200_mul_mat_args;

#define B200_MM_FORCE_GEMV  1   /* use the dp4a GEMV for any n (column chunks of <= 8) */
#define B200_MM_FORCE_GEMM  2   /* use the tcgen05 GEMM for any n */
#define B200_MM_EXPORT      4   /* b200_plan_create on a row-split plan: leave the COMPLETE dst (all ranks' slices) in dst_dev */

B200_API int b200_mul_mat(b200_ctx *ctx, const b200_mul_mat_args *args);
/* the same with the operators that follow a decode mul_mat in a transformer block folded into its epilogue:
 *     dst = (act(src0 x src1 + bias) + residual) + residual2
 * -- the ADD of a bias row, the GELU and the ADD of the residual stream (examples/gpt-2/main-backend.cpp:614-625, :659-672, :683-699); GPT-J sums its
 * MLP branch and the residual stream onto the attention projection (examples/gpt-j/main.cpp:556-559), hence the second one --
 * for 2-D decode shapes that take the GEMV (ne11 <= 8, no batch dims).  bias_dev [ne01] or NULL, residual_dev dense like dst or NULL (may be
 * dst_dev itself), act = B200_EPI_NONE / B200_EPI_GELU.  Same bits as b200_mul_mat followed by the separate operators.
 * B200_ERR_UNSUPPORTED for any other shape: the caller runs the operators one by one. */
#define B200_EPI_NONE 0
#define B200_EPI_GELU 1
typedef struct b200_epilogue {
    const float *bias_dev;
    const float *residual_dev;
    int32_t      act;
    int32_t      reserved;
    const float *residual2_dev;     /* added after residual_dev, or NULL (may be dst_dev itself, like residual_dev) */
} b200_epilogue;
B200_API int b200_mul_mat_fused(b200_ctx *ctx, const b200_mul_mat_args *args, const b200_epilogue *epilogue);
/* `count` mul_mats with NO data dependencies among them (e.g. the q/k/v/fc_in projections of a GPT-J block, which all
 * read the same normalised activations: examples/gpt-j/main.cpp:462-467, :535-538).  Decode-shaped entries that share
 * type, k and src1 are streamed by ONE launch (one activation quantization, the grid divided among the matrices);
 * everything else falls back to b200_mul_mat one by one.  Same result as `count` separate calls. */
B200_API int b200_mul_mat_batch(b200_ctx *ctx, const b200_mul_mat_args *args, int count);

/* ---- row-split decode across GPUs: GEMV fused with its all-gather -----------------------------------
 * Replaces the reference's multi-device mul_mat (split buffer + peer memcpys to a main GPU,
 * src/ggml-cuda.cu:578-975, :1360-1647) for the decode case (n == 1), one process per GPU.
 * Each rank holds a row slice of src0.  Instead of "kernel, then collective", the GEMV epilogue stores every
 * finished dst element straight into the activation vector of EVERY rank over NVLink (peer-mapped pointers from
 * b200_ipc_export / b200_ipc_import on b200_malloc'ed buffers).  No flags, no fences: an element travels as one
 * 8-byte store {fp32 value, u32 tag} ("LL" layout, the tag is the execution count of the launch), which the fabric
 * delivers atomically; the consuming kernel's activation loads simply re-read until the tag is the one it expects.
 * Tags are monotonic and every launch slot keeps its execution count on the device, so a captured CUDA graph replays
 * without resetting anything.
 *   peer_dst[r] : rank r's LL activation vector for this launch's dst: uint64[>= row0 + ne01], zero-initialised,
 *                 same layout on every rank; [rank] is the local one
 *   state       : local uint32 [n_slots][2], zero-initialised ({arrived CTAs, executions})
 *   slot        : index of this launch in the repeated sequence (unique per launch, including b200_gather_finish)
 *   wait_slot   : >= 0: src1 of this launch is the LL vector produced by that slot (args->src1_dev points to it);
 *                 -1: src1 is a plain local fp32 vector */
#define B200_MAX_RANKS 8
typedef struct b200_gather {
    int32_t   world, rank;
    int32_t   slot, wait_slot;
    int64_t   row0;
    void     *peer_dst[B200_MAX_RANKS];
    uint32_t *state;
} b200_gather;
B200_API int b200_mul_mat_gather(b200_ctx *ctx, const b200_mul_mat_args *args, const b200_gather *gather);
/* b200_mul_mat_batch for the fused path: independent same-input slices share one launch when their gather descriptions
 * agree (same group, state, wait_slot; LL vectors at the same relative offset on every rank); else one by one */
B200_API int b200_mul_mat_gather_batch(b200_ctx *ctx, const b200_mul_mat_args *args, const b200_gather *gathers, int count);
/* end of a sequence: wait for the LL vector ll_src_dev (count elements, produced by gather->wait_slot) to be complete
 * and write it out as plain fp32 (dense_out_dev), e.g. the logits before they are read back */
B200_API int b200_gather_finish(b200_ctx *ctx, const b200_gather *gather, const void *ll_src_dev, float *dense_out_dev, int64_t count);
/* CUDA IPC plumbing for the peer pointers (handle = 64 bytes, exchange it with any host-side transport) */
B200_API int b200_ipc_export(b200_ctx *ctx, void *dptr, void *handle64_out);
B200_API int b200_ipc_import(b200_ctx *ctx, const void *handle64, void **peer_ptr_out);
B200_API int b200_ipc_close(b200_ctx *ctx, void *peer_ptr);

/* ---- decode plans: a dependent sequence of decode mul_mats as ONE persistent launch -----------------------
 * Replaces ggml_backend_graph_plan_create / _compute / _free (src/ggml-backend-impl.h:94-99; the reference's CUDA backend
 * replays a captured CUDA graph of the cgraph instead, src/ggml-cuda.cu:2461-2709) for graphs whose compute nodes are
 * MUL_MAT{Q4_0|Q8_0} x F32 with ONE activation column (ne11 == 1, 2-D weights, ne00 % 256 == 0, ne00 <= 32768).
 * args[0..count) are executed with ggml's dataflow semantics: op j reads the result of the latest earlier op i whose
 * dst_dev == args[j].src1_dev (and ne01 of i == ne00 of j), otherwise src1_dev is a vector produced outside the plan.
 * One persistent CTA per SM walks the list; a producer thread per CTA streams the weights of op 0, 1, 2, ... back to
 * back through a shared-memory ring (HBM never idles at an op boundary), and results travel between ops as tagged
 * 8-byte elements, so there is no grid-wide barrier anywhere.  Results are bit-identical to `count` b200_mul_mat calls.
 * Every dst_dev is still written as plain fp32 (may be NULL for intermediates nobody outside reads), EXCEPT a dst whose memory
 * a later op's dst reuses (what ggml_gallocr does with dead intermediates, src/ggml-alloc.c): that store is dropped, see
 * b200_plan_plain_stores.  The launch is cooperative (all CTAs resident or none), every in-kernel wait is bounded.
 * B200_ERR_UNSUPPORTED: a shape outside the above, a src1 that is a partial view of another op's dst, or a dst that overwrites
 * an outside input some CTA may still have to read -- the caller then runs the nodes one by one through b200_mul_mat /
 * b200_mul_mat_batch.
 *
 * Row-split plans (split != NULL, one process per GPU): args[i] describes THIS rank's row slice of op i (ne01 = local
 * rows, src0 = local slice), split->row0[i] / m_total[i] place it in the whole matrix.  The tagged stores go to every
 * rank's arena over NVLink (peer_arena[r] from b200_ipc_export/import of a zero-initialised b200_malloc'ed buffer of
 * b200_plan_arena_bytes() on each rank), i.e. the all-gather is part of the GEMV epilogue.  dst_dev receives only the
 * local rows unless the op carries B200_MM_EXPORT.  The LAST op of a row-split plan must carry B200_MM_EXPORT (that is
 * also what keeps ranks within one token of each other), and EVERY rank must pass the SAME export set: an exported op is one
 * whose rows every rank sends to every rank (a rank that exports an op its peers do not would wait for rows that never come --
 * until the plan's timeout). */
typedef struct b200_plan b200_plan;
typedef struct b200_plan_split {
    int32_t        world, rank;
    void          *peer_arena[B200_MAX_RANKS];
    const int64_t *row0;      /* [count] */
    const int64_t *m_total;   /* [count] */
} b200_plan_split;
B200_API size_t b200_plan_arena_bytes(const b200_mul_mat_args *args, int count, const b200_plan_split *split);
B200_API int  b200_plan_create(b200_ctx *ctx, const b200_mul_mat_args *args, int count, const b200_plan_split *split, b200_plan **out);
/* the host-only analysis b200_plan_create starts with (no device needed): same status codes; src_op_out[i] (may be NULL)
 * receives the index of the op whose dst is op i's src1, or -1 for a vector from outside the plan */
B200_API int  b200_plan_analyze(const b200_mul_mat_args *args, int count, const b200_plan_split *split, int32_t *src_op_out);
/* host-only as well: published_out[i] = 1 when op i takes its src1 from the once-per-GPU quantization of an in-plan vector
 * (the publisher / fetcher warps of the plan kernel) on a device with sm_count SMs, else 0.  min_k / dist = the context options
 * "plan_pub_min_k" / "plan_pub_dist"; <= 0 selects their defaults (4096, 2) */
B200_API int  b200_plan_published(const b200_mul_mat_args *args, int count, const b200_plan_split *split, int sm_count,
                                  int min_k, int dist, int32_t *published_out);
/* host-only: plain_out[i] = 1 when the plan stores op i's result in args[i].dst_dev as plain fp32, 0 when it does not: dst_dev
 * is NULL, or a LATER op's dst occupies the same memory (buffer reuse by a graph allocator such as ggml_gallocr,
 * src/ggml-alloc.c: the earlier tensor is dead by the end of the graph; inside the plan its value travels as a tagged vector) */
B200_API int  b200_plan_plain_stores(const b200_mul_mat_args *args, int count, const b200_plan_split *split, int32_t *plain_out);
/* a plan must not be launched concurrently with itself (its hand-off arena and launch counter are per plan) */
B200_API int  b200_plan_launch(b200_ctx *ctx, b200_plan *plan);      /* asynchronous on the context's stream */
B200_API void b200_plan_destroy(b200_plan *plan);
/* device-side timeline of the last launch (plan created with option "plan_trace" = 1): [nops + 1][grid][4] ns stamps
 * (src1 complete, activations quantized, first weights landed, last row done; the last row holds per-CTA totals: ns the
 * producer spent blocked on a full ring, ns one consumer warp spent blocked on an empty ring, ns spent in quantization
 * phases); B200_ERR_UNSUPPORTED when not tracing */
B200_API int  b200_plan_trace(b200_ctx *ctx, b200_plan *plan, unsigned long long *out_host, size_t capacity_u64, int *nops, int *grid);

/* ---- the operators either side of the path ---------------------------------------------------------------
 * What a GPT-2 / GPT-J graph computes between its quantized mul_mats, so that the reference's gpt-2-backend runs its WHOLE graph on
 * this backend (examples/gpt-2/main-backend.cpp:442-717, computed by one ggml_backend_graph_compute at :768).  Each call stands in
 * for one ggml_compute_forward_* of src/ggml.c (cited per entry) and is asynchronous on the context's stream.  A b200_tensor is
 * the part of struct ggml_tensor (include/ggml/ggml.h:565-604) a kernel needs: device address, type, ne[], nb[] (bytes).  For a
 * repacked Q4_0 / Q8_0 tensor `data` is the start of the ROOT tensor's qs plane, q_total_blocks the blocks of the root and
 * q_block_off the first block of this tensor (as in b200_mul_mat_args); nb[] stay the wire strides of ggml.
 * Every operator may run in place (dst->data == src0->data), as ggml_gallocr arranges it.
 * B200_ERR_UNSUPPORTED = a type / layout the backend's supports_op must have declined. */
typedef struct b200_tensor {
    void   *data;
    int32_t type;             /* B200_TYPE_* */
    int32_t reserved;
    int64_t ne[4];
    int64_t nb[4];
    int64_t q_total_blocks;   /* repacked quantized tensors only */
    int64_t q_block_off;
} b200_tensor;

/* numeric values of enum ggml_unary_op (include/ggml/ggml.h:514-530) */
enum { B200_UNARY_ABS = 0, B200_UNARY_SGN, B200_UNARY_NEG, B200_UNARY_STEP, B200_UNARY_TANH, B200_UNARY_ELU, B200_UNARY_RELU, B200_UNARY_SIGMOID,
       B200_UNARY_GELU, B200_UNARY_GELU_QUICK, B200_UNARY_SILU, B200_UNARY_HARDSWISH, B200_UNARY_HARDSIGMOID, B200_UNARY_COUNT };
enum { B200_OP_ADD = 0, B200_OP_MUL = 1, B200_OP_DIV = 2 };

/* GGML_OP_GET_ROWS (ggml_compute_forward_get_rows, src/ggml.c:13049): dst[:, i10, i11, i12] = src0[:, rows[i10, i11, i12], i11, i12] as
 * F32; src0 F32, F16 or repacked Q4_0 / Q8_0 (dequantize_row_q4_0 / _q8_0, src/ggml-quants.c:980-998, :1074-1088), rows I32 */
B200_API int b200_op_get_rows(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *rows, const b200_tensor *dst);
/* GGML_OP_ADD / MUL / DIV on F32 (ggml_compute_forward_add_f32, src/ggml.c:8568; mul_f32 :9687): src1 is repeated over src0 in every
 * dimension (ggml_can_repeat) */
B200_API int b200_op_binary(b200_ctx *ctx, int op, const b200_tensor *src0, const b200_tensor *src1, const b200_tensor *dst);
/* GGML_OP_UNARY on contiguous F32 (ggml_compute_forward_unary; gelu = the tanh form of src/ggml.c:1966 in fp32 -- the CPU rounds x and
 * the result through its fp16 table, :1978-1991) */
B200_API int b200_op_unary(b200_ctx *ctx, int op, const b200_tensor *src0, const b200_tensor *dst);
/* GGML_OP_NORM / GGML_OP_RMS_NORM over dim 0 (ggml_compute_forward_norm_f32, src/ggml.c:11353), optionally fused with the MUL by a
 * gain row and the ADD of a bias row that follow it in a transformer block (examples/gpt-2/main-backend.cpp:490-498): dst = norm(src0)
 * [* gain] [+ bias]; gain / bias: F32 vectors of ne[0] elements or NULL */
B200_API int b200_op_norm(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *gain, const b200_tensor *bias, const b200_tensor *dst,
                          float eps, int rms);
/* GGML_OP_SCALE on contiguous F32 (ggml_compute_forward_scale_f32, src/ggml.c:12637) */
B200_API int b200_op_scale(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *dst, float s);
/* GGML_OP_DIAG_MASK_INF on contiguous F32 (ggml_compute_forward_diag_mask_f32, src/ggml.c:13301): dst[i, j, ...] = i > n_past + j ? -inf : src0 */
B200_API int b200_op_diag_mask_inf(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *dst, int n_past);
/* GGML_OP_SOFT_MAX over dim 0 of contiguous F32 (ggml_compute_forward_soft_max_f32, src/ggml.c:13393): softmax(src0 * scale + slope(head) *
 * mask[:, i1]); mask F32 / F16 [ne0, >= ne1] or NULL; max_bias > 0 = ALiBi slopes per head (dim 2).  n_past >= 0 additionally applies the
 * causal mask of a GGML_OP_DIAG_MASK_INF(n_past) in front (fused SCALE + DIAG_MASK_INF + SOFT_MAX, main-backend.cpp:567-583); -1 = none */
B200_API int b200_op_soft_max(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *mask, const b200_tensor *dst, float scale, float max_bias,
                              int n_past);
/* GGML_OP_CPY / DUP / CONT (ggml_compute_forward_dup, src/ggml.c:8535): element i of the flattened src0 to element i of the flattened dst,
 * both arbitrarily strided; F32 <-> F16 conversions, or any same-type pair of 2- / 4-byte elements (F32, F16, I32, I16) */
B200_API int b200_op_copy(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *dst);
/* GGML_OP_ROPE, forward (src/ggml.c:13775, :13953; op_params of ggml_rope_custom, src/ggml.c:5866-5889): src0 F32 or F16 [ne0][heads][tokens][b],
   pos I32 [tokens].  mode bit 1 = NeoX pairing; bit 2 (GLM) is B200_ERR_UNSUPPORTED.  dst has src0's type, or F16 for an F32 src0 (the CPY that
   stores the rotated k into an F16 KV cache, examples/gpt-j/main.cpp:473-484, folded into the store: same bits as ROPE then CPY). */
typedef struct b200_rope_params {
    int32_t n_dims, mode, n_ctx, n_orig_ctx;
    float   freq_base, freq_scale, ext_factor, attn_factor, beta_fast, beta_slow, xpos_base;
    int32_t xpos_down;
} b200_rope_params;
B200_API int b200_op_rope(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *pos, const b200_tensor *dst, const b200_rope_params *params);
/* GGML_OP_REPEAT (src/ggml.c:10323): dst tiles src0 along every dimension (dst->ne[i] a multiple of src0->ne[i]); F32, F16, I16, I32 */
B200_API int b200_op_repeat(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *dst);
/* The attention of a decode step in one launch -- what the six graph nodes MUL_MAT(K, Q) -> SCALE -> DIAG_MASK_INF -> SOFT_MAX -> MUL_MAT(V, .) ->
 * CPY / CONT of the PERMUTEd result compute (examples/gpt-j/main.cpp:490-530, examples/gpt-2/main-backend.cpp:567-610):
 *   dst[:, h, i] = sum_t V[t, :, h] * softmax_t(scale * K[:, t, h] . Q[:, i, h]  masked for t > n_past + i)
 * q F32 [hd][N][H], k F16/F32 [hd][T][H], v F16/F32 [T][hd][H] (any row / head strides: views of the KV cache), dst F32 [hd][H][N];
 * n_past < 0 = no causal mask.  Decode-sized problems only (N <= 8, T <= 1024): B200_ERR_UNSUPPORTED otherwise, the caller runs the nodes. */
B200_API int b200_op_attention_decode(b200_ctx *ctx, const b200_tensor *q, const b200_tensor *k, const b200_tensor *v, const b200_tensor *dst, float scale,
                                      int n_past);
/* GGML_OP_MUL_MAT with an F32 / F16 src0 (ggml_compute_forward_mul_mat, src/ggml.c:11808): dst[n][m] = sum_k src0[m][k] * src1[n][k] per
 * (i2, i3) with the broadcast of src0 over src1's batch dims; both operands k-contiguous, any row / batch strides (K*Q and V*softmax(KQ) on
 * permuted views of the KV cache, main-backend.cpp:567, :597); fp32 accumulation */
B200_API int b200_op_mul_mat_dense(b200_ctx *ctx, const b200_tensor *src0, const b200_tensor *src1, const b200_tensor *dst);

/* parity instrumentation: per-block int32 partial sums exactly as the kernels form them.
 * out_dev [n][m][k/32] int32.  path 0 = GEMV inner loop (dp4a), path 1 = GEMM (tcgen05 accumulators).
 * Compared bit-for-bit with the integer loop of src/ggml-quants.c:3858-3869 / :5010-5015. */
B200_API int b200_block_dots(b200_ctx *ctx, int type, const void *src0_dev, int64_t k, int64_t m,
                             const float *src1_dev, int64_t n, int32_t *out_dev, int path);

/* host-buffer convenience used by end-to-end timing and smoke: upload src1 (n*k floats, dense),
 * mul_mat against a resident repacked src0 [k, m], download dst (n*m floats).  Synchronous. */
B200_API int b200_mul_mat_host(b200_ctx *ctx, int type, const void *src0_dev, int64_t k, int64_t m,
                               const float *src1_host, int64_t n, float *dst_host);

#ifdef __cplusplus
}
#endif
#endif /* GGML_B200_H */
