/*
 * ref_shim.c -- flat C entry points onto the UNMODIFIED reference (libggml_cpu.so built from
 * /root/reference by oracle/Makefile).  Our code, compiled against the reference's public headers;
 * output goes to oracle/_ref/ only.  TEST INFRASTRUCTURE: used to pin oracle/qmm_oracle.c, to
 * generate tests/golden/, and as bench.py's "reference" CPU baseline.
 */
#include "ggml.h"
#include "ggml-alloc.h"
#include "ggml-backend.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* GGML_FP16_TO_FP32 is a table lookup on x86 (src/ggml-impl.h:561-587); the table is filled by the
 * first ggml_init() (src/ggml.c:2897-2920).  Make sure that has happened before any vec_dot. */
static void ref_init(void) {
    static int done = 0;
    if (done) return;
    struct ggml_init_params ip = { 1024, NULL, true };
    struct ggml_context *c = ggml_init(ip);
    ggml_free(c);
    done = 1;
}

/* Which backend the graph helpers below run on: NULL = the reference CPU backend; otherwise a name from the reference's
 * backend registry (src/ggml-backend.c:397-470).  The drop-in build of this file (_ref/libdropin_shim.so, linked against
 * the reference core + our backend) uses "B2000" to run the very same ggml graph on the B200 backend. */
static char g_backend_name[64] = "";
void ref_select_backend(const char *name) { snprintf(g_backend_name, sizeof(g_backend_name), "%s", name ? name : ""); }
static ggml_backend_t ref_backend_new(int n_threads) {
    if (g_backend_name[0] == 0) {
        ggml_backend_t b = ggml_backend_cpu_init();
        ggml_backend_cpu_set_n_threads(b, n_threads);
        return b;
    }
    const size_t n = ggml_backend_reg_get_count();
    for (size_t i = 0; i < n; i++)
        if (strcmp(ggml_backend_reg_get_name(i), g_backend_name) == 0) return ggml_backend_reg_init_backend(i, NULL);
    fprintf(stderr, "ref_shim: backend %s is not in the registry\n", g_backend_name);
    return NULL;
}
#ifdef REF_SHIM_DROPIN
#include "ggml-b200.h"
#endif

/* runtime from_float of a type: type_traits[type].from_float (src/ggml.c:617-632, :697-712) */
void ref_from_float(int type, const float *x, void *y, int64_t k) {
    ggml_type_traits_t tt = ggml_internal_get_type_traits((enum ggml_type)type);
    tt.from_float(x, y, k);
}

void ref_from_float_reference(int type, const float *x, void *y, int64_t k) {
    ggml_type_traits_t tt = ggml_internal_get_type_traits((enum ggml_type)type);
    tt.from_float_reference(x, y, k);
}

void ref_to_float(int type, const void *x, float *y, int64_t k) {
    ref_init();
    ggml_type_traits_t tt = ggml_internal_get_type_traits((enum ggml_type)type);
    tt.to_float(x, y, k);
}

/* type_traits[type].vec_dot(n, &s, 0, x, 0, y, 0, 1) (src/ggml.c:12084) */
float ref_vec_dot(int type, int64_t k, const void *x, const void *y) {
    ref_init();
    ggml_type_traits_t tt = ggml_internal_get_type_traits((enum ggml_type)type);
    float s = 0.0f;
    tt.vec_dot((int)k, &s, 0, x, 0, y, 0, 1);
    return s;
}

/* ggml_quantize_chunk(type, src, dst, 0, nrows, n_per_row, NULL) (src/ggml.c:21594-21623) */
size_t ref_quantize_chunk(int type, const float *src, void *dst, int64_t nrows, int64_t n_per_row) {
    ggml_quantize_init((enum ggml_type)type);
    return ggml_quantize_chunk((enum ggml_type)type, src, dst, 0, nrows, n_per_row, NULL);
}

size_t ref_row_size(int type, int64_t k) { return ggml_row_size((enum ggml_type)type, k); }

/*
 * One MUL_MAT node through the reference CPU backend, exactly as test-backend-ops does it
 * (tests/test-backend-ops.cpp:940-946 + src/ggml-backend.c:800-820).
 * a: type [k, m, ne02, ne03]; b: F32 [k, n, ne12, ne13] dense; dst F32 [m, n, ne12, ne13].
 * reps > 1 repeats graph_compute (for timing); returns best wall time in microseconds.
 */
typedef struct ref_mm {
    struct ggml_context *ctx;
    ggml_backend_t backend;
    ggml_backend_buffer_t buf;
    struct ggml_tensor *a, *b, *out;
    struct ggml_cgraph *gf;
} ref_mm;

ref_mm *ref_mm_create(int type, int64_t k, int64_t m, int64_t ne02, int64_t ne03,
                      int64_t n, int64_t ne12, int64_t ne13, int n_threads) {
    ref_mm *h = (ref_mm *)calloc(1, sizeof(ref_mm));
    struct ggml_init_params ip = { ggml_tensor_overhead() * 16 + ggml_graph_overhead(), NULL, true };
    h->ctx = ggml_init(ip);
    h->a = ggml_new_tensor_4d(h->ctx, (enum ggml_type)type, k, m, ne02, ne03);
    h->b = ggml_new_tensor_4d(h->ctx, GGML_TYPE_F32, k, n, ne12, ne13);
    h->out = ggml_mul_mat(h->ctx, h->a, h->b);
    h->gf = ggml_new_graph(h->ctx);
    ggml_build_forward_expand(h->gf, h->out);
    h->backend = ggml_backend_cpu_init();
    ggml_backend_cpu_set_n_threads(h->backend, n_threads);
    h->buf = ggml_backend_alloc_ctx_tensors(h->ctx, h->backend);
    return h;
}

void ref_mm_set_a(ref_mm *h, const void *data) { ggml_backend_tensor_set(h->a, data, 0, ggml_nbytes(h->a)); }
void ref_mm_set_b(ref_mm *h, const void *data) { ggml_backend_tensor_set(h->b, data, 0, ggml_nbytes(h->b)); }
void ref_mm_get_out(ref_mm *h, void *data)     { ggml_backend_tensor_get(h->out, data, 0, ggml_nbytes(h->out)); }

double ref_mm_compute(ref_mm *h, int reps) {
    double best = 1e30;
    for (int r = 0; r < (reps < 1 ? 1 : reps); r++) {
        const int64_t t0 = ggml_time_us();
        ggml_backend_graph_compute(h->backend, h->gf);
        const int64_t t1 = ggml_time_us();
        if ((double)(t1 - t0) < best) best = (double)(t1 - t0);
    }
    return best;
}

void ref_mm_free(ref_mm *h) {
    if (!h) return;
    ggml_backend_buffer_free(h->buf);
    ggml_backend_free(h->backend);
    ggml_free(h->ctx);
    free(h);
}

void ref_time_init(void) { ref_init(); ggml_time_init(); }

/*
 * A chain of MUL_MAT nodes in ONE ggml graph on the reference CPU backend: cur = W[i] x cur, the way a model
 * graph holds them (one ggml_backend_graph_compute per token, examples/gpt-2/main-backend.cpp:768).
 * mats i = 0..n_mats-1 use weight tensor wid[i] (so a few distinct weight sets can be cycled); weight j has
 * shape [wk[j], wm[j]].  x: F32 [wk[wid[0]], n].
 */
typedef struct ref_chain {
    struct ggml_context *ctx;
    ggml_backend_t backend;
    ggml_backend_buffer_t buf;
    struct ggml_tensor **w;
    struct ggml_tensor *x, *out;
    struct ggml_cgraph *gf;
    ggml_backend_graph_plan_t plan;
    int n_weights;
    struct ggml_context *ctx2;     /* ref_dag_create_galloc: the compute nodes live here, allocated by ggml_gallocr */
    ggml_gallocr_t galloc;
    ggml_backend_buffer_t buf2;    /* ref_dag_create_split: the activations' buffer */
} ref_chain;

ref_chain *ref_chain_create(int type, int n_mats, const int *wid, int n_weights, const int64_t *wk, const int64_t *wm,
                            int64_t n, int n_threads) {
    ref_init();
    ref_chain *h = (ref_chain *)calloc(1, sizeof(ref_chain));
    struct ggml_init_params ip = { ggml_tensor_overhead() * (size_t)(n_mats + n_weights + 8) + ggml_graph_overhead_custom(n_mats + n_weights + 64, false), NULL, true };
    h->ctx = ggml_init(ip);
    h->n_weights = n_weights;
    h->w = (struct ggml_tensor **)calloc((size_t)n_weights, sizeof(*h->w));
    for (int j = 0; j < n_weights; j++) h->w[j] = ggml_new_tensor_2d(h->ctx, (enum ggml_type)type, wk[j], wm[j]);
    h->x = ggml_new_tensor_2d(h->ctx, GGML_TYPE_F32, wk[wid[0]], n);
    struct ggml_tensor *cur = h->x;
    for (int i = 0; i < n_mats; i++) cur = ggml_mul_mat(h->ctx, h->w[wid[i]], cur);
    h->out = cur;
    h->gf = ggml_new_graph_custom(h->ctx, (size_t)(n_mats + n_weights + 64), false);
    ggml_build_forward_expand(h->gf, h->out);
    h->backend = ref_backend_new(n_threads);
    if (!h->backend) return NULL;
    h->buf = ggml_backend_alloc_ctx_tensors(h->ctx, h->backend);
    return h;
}

void ref_chain_set_weight(ref_chain *h, int j, const void *data) { ggml_backend_tensor_set(h->w[j], data, 0, ggml_nbytes(h->w[j])); }
void ref_chain_set_x(ref_chain *h, const float *x) { ggml_backend_tensor_set(h->x, x, 0, ggml_nbytes(h->x)); }
void ref_chain_get_out(ref_chain *h, float *out) { ggml_backend_tensor_get(h->out, out, 0, ggml_nbytes(h->out)); }
int64_t ref_chain_out_elements(ref_chain *h) { return ggml_nelements(h->out); }

double ref_chain_compute(ref_chain *h) {
    const int64_t t0 = ggml_time_us();
    ggml_backend_graph_compute(h->backend, h->gf);
    return (double)(ggml_time_us() - t0);
}

/* the same graph through the explicit plan API: ggml_backend_graph_plan_create once, _compute per call (src/ggml-backend.c:257-273) */
double ref_chain_compute_planned(ref_chain *h) {
    if (!h->plan) h->plan = ggml_backend_graph_plan_create(h->backend, h->gf);
    const int64_t t0 = ggml_time_us();
    ggml_backend_graph_plan_compute(h->backend, h->plan);
    ggml_backend_synchronize(h->backend);
    return (double)(ggml_time_us() - t0);
}

void ref_chain_free(ref_chain *h) {
    if (!h) return;
    if (h->plan) ggml_backend_graph_plan_free(h->backend, h->plan);
    if (h->galloc) ggml_gallocr_free(h->galloc);
    if (h->buf2) ggml_backend_buffer_free(h->buf2);
    ggml_backend_buffer_free(h->buf);
    ggml_backend_free(h->backend);
    if (h->ctx2) ggml_free(h->ctx2);
    ggml_free(h->ctx);
    free(h->w);
    free(h);
}

/*
 * A DAG of MUL_MAT nodes in ONE ggml graph on the reference CPU backend (one ggml_backend_graph_compute per token, the
 * way examples/gpt-j/main.cpp:593 evaluates its graph): node i = W[node_w[i]] x (node_src[i] < 0 ? x : node[node_src[i]]).
 * Every node is expanded into the graph (q and k of a block have no consumer among the mul_mats but are computed, as in the
 * model); the result read back is the last node.  x: F32 [wk[node_w[first node reading x]], n].
 */
ref_chain *ref_dag_create(int type, int n_nodes, const int *node_w, const int *node_src, int n_weights, const int64_t *wk,
                          const int64_t *wm, int64_t n, int n_threads) {
    ref_init();
    ref_chain *h = (ref_chain *)calloc(1, sizeof(ref_chain));
    struct ggml_init_params ip = { ggml_tensor_overhead() * (size_t)(n_nodes + n_weights + 8) + ggml_graph_overhead_custom(n_nodes + n_weights + 64, false), NULL, true };
    h->ctx = ggml_init(ip);
    h->n_weights = n_weights;
    h->w = (struct ggml_tensor **)calloc((size_t)n_weights, sizeof(*h->w));
    for (int j = 0; j < n_weights; j++) h->w[j] = ggml_new_tensor_2d(h->ctx, (enum ggml_type)type, wk[j], wm[j]);
    int64_t kx = 0;
    for (int i = 0; i < n_nodes; i++) if (node_src[i] < 0) { kx = wk[node_w[i]]; break; }
    h->x = ggml_new_tensor_2d(h->ctx, GGML_TYPE_F32, kx, n);
    struct ggml_tensor **nodes = (struct ggml_tensor **)calloc((size_t)n_nodes, sizeof(*nodes));
    h->gf = ggml_new_graph_custom(h->ctx, (size_t)(n_nodes + n_weights + 64), false);
    for (int i = 0; i < n_nodes; i++) {
        nodes[i] = ggml_mul_mat(h->ctx, h->w[node_w[i]], node_src[i] < 0 ? h->x : nodes[node_src[i]]);
        ggml_build_forward_expand(h->gf, nodes[i]);
    }
    h->out = nodes[n_nodes - 1];
    free(nodes);
    h->backend = ref_backend_new(n_threads);
    if (!h->backend) return NULL;
    h->buf = ggml_backend_alloc_ctx_tensors(h->ctx, h->backend);
    return h;
}

/*
 * The same DAG the way the reference's examples hold a model graph (examples/gpt-2/main-backend.cpp:442-770): weights and the
 * input in a buffer of their own (ggml_backend_alloc_ctx_tensors), the compute nodes allocated by ggml_gallocr_alloc_graph, which
 * REUSES the memory of dead intermediates for later nodes.  Only the last node (ggml_set_output) is meaningful afterwards.
 */
ref_chain *ref_dag_create_galloc(int type, int n_nodes, const int *node_w, const int *node_src, int n_weights, const int64_t *wk,
                                 const int64_t *wm, int64_t n, int n_threads) {
    ref_init();
    ref_chain *h = (ref_chain *)calloc(1, sizeof(ref_chain));
    struct ggml_init_params ip = { ggml_tensor_overhead() * (size_t)(n_weights + 8), NULL, true };
    h->ctx = ggml_init(ip);
    h->n_weights = n_weights;
    h->w = (struct ggml_tensor **)calloc((size_t)n_weights, sizeof(*h->w));
    for (int j = 0; j < n_weights; j++) h->w[j] = ggml_new_tensor_2d(h->ctx, (enum ggml_type)type, wk[j], wm[j]);
    int64_t kx = 0;
    for (int i = 0; i < n_nodes; i++) if (node_src[i] < 0) { kx = wk[node_w[i]]; break; }
    h->x = ggml_new_tensor_2d(h->ctx, GGML_TYPE_F32, kx, n);
    ggml_set_input(h->x);
    h->backend = ref_backend_new(n_threads);
    if (!h->backend) return NULL;
    h->buf = ggml_backend_alloc_ctx_tensors(h->ctx, h->backend);
    struct ggml_init_params ip2 = { ggml_tensor_overhead() * (size_t)(n_nodes + 8) + ggml_graph_overhead_custom(n_nodes + n_weights + 64, false), NULL, true };
    h->ctx2 = ggml_init(ip2);
    struct ggml_tensor **nodes = (struct ggml_tensor **)calloc((size_t)n_nodes, sizeof(*nodes));
    h->gf = ggml_new_graph_custom(h->ctx2, (size_t)(n_nodes + n_weights + 64), false);
    for (int i = 0; i < n_nodes; i++) {
        nodes[i] = ggml_mul_mat(h->ctx2, h->w[node_w[i]], node_src[i] < 0 ? h->x : nodes[node_src[i]]);
        if (i == n_nodes - 1) ggml_set_output(nodes[i]);
        ggml_build_forward_expand(h->gf, nodes[i]);
    }
    h->out = nodes[n_nodes - 1];
    free(nodes);
    h->galloc = ggml_gallocr_new(ggml_backend_get_default_buffer_type(h->backend));
    if (!ggml_gallocr_alloc_graph(h->galloc, h->gf)) return NULL;
    return h;
}
/* how many of the graph's nodes share their data pointer with an earlier node (what the allocator reused) */
int ref_chain_aliased_nodes(ref_chain *h) {
    int n = 0;
    for (int i = 0; i < h->gf->n_nodes; i++)
        for (int j = 0; j < i; j++)
            if (h->gf->nodes[i]->data == h->gf->nodes[j]->data) { n++; break; }
    return n;
}

/* any node of the last graph, by position in the cgraph (to compare intermediates between backends) */
int ref_chain_n_nodes(ref_chain *h) { return h->gf->n_nodes; }
int64_t ref_chain_node_elements(ref_chain *h, int i) { return ggml_nelements(h->gf->nodes[i]); }
void ref_chain_get_node(ref_chain *h, int i, float *out) { ggml_backend_tensor_get(h->gf->nodes[i], out, 0, ggml_nbytes(h->gf->nodes[i])); }
#ifdef REF_SHIM_DROPIN
/*
 * The same DAG with the WEIGHTS in the split buffer type of the dropped-in backend (ggml_backend_cuda_split_buffer_type,
 * src/ggml-cuda.h:28-29: rows of every matrix divided across the devices of this process, equal shares) and the activations in
 * the backend's default buffer -- how a host written against the reference uses several GPUs (examples/gpt-2/main-sched.cpp has
 * the per-layer variant; llama.cpp's --split-mode row is the consumer of this interface).
 */
#include "ggml-cuda.h"
ref_chain *ref_dag_create_split(int type, int n_nodes, const int *node_w, const int *node_src, int n_weights, const int64_t *wk,
                                const int64_t *wm, int64_t n, int n_threads) {
    ref_init();
    ggml_backend_buffer_type_t split = ggml_backend_cuda_split_buffer_type(NULL);
    if (!split) return NULL;
    ref_chain *h = (ref_chain *)calloc(1, sizeof(ref_chain));
    struct ggml_init_params ip = { ggml_tensor_overhead() * (size_t)(n_weights + 8), NULL, true };
    h->ctx = ggml_init(ip);
    h->n_weights = n_weights;
    h->w = (struct ggml_tensor **)calloc((size_t)n_weights, sizeof(*h->w));
    for (int j = 0; j < n_weights; j++) h->w[j] = ggml_new_tensor_2d(h->ctx, (enum ggml_type)type, wk[j], wm[j]);
    h->backend = ref_backend_new(n_threads);
    if (!h->backend) return NULL;
    h->buf = ggml_backend_alloc_ctx_tensors_from_buft(h->ctx, split);
    if (!h->buf) return NULL;
    struct ggml_init_params ip2 = { ggml_tensor_overhead() * (size_t)(n_nodes + 8) + ggml_graph_overhead_custom(n_nodes + n_weights + 64, false), NULL, true };
    h->ctx2 = ggml_init(ip2);
    int64_t kx = 0;
    for (int i = 0; i < n_nodes; i++) if (node_src[i] < 0) { kx = wk[node_w[i]]; break; }
    h->x = ggml_new_tensor_2d(h->ctx2, GGML_TYPE_F32, kx, n);
    struct ggml_tensor **nodes = (struct ggml_tensor **)calloc((size_t)n_nodes, sizeof(*nodes));
    h->gf = ggml_new_graph_custom(h->ctx2, (size_t)(n_nodes + n_weights + 64), false);
    for (int i = 0; i < n_nodes; i++) {
        nodes[i] = ggml_mul_mat(h->ctx2, h->w[node_w[i]], node_src[i] < 0 ? h->x : nodes[node_src[i]]);
        ggml_build_forward_expand(h->gf, nodes[i]);
    }
    h->out = nodes[n_nodes - 1];
    free(nodes);
    h->buf2 = ggml_backend_alloc_ctx_tensors(h->ctx2, h->backend);
    return h->buf2 ? h : NULL;
}
/* how many of the graph_compute calls on this handle's backend went down as one persistent launch (decode plan) */
int64_t ref_chain_plan_launches(ref_chain *h) { return ggml_backend_is_b200(h->backend) ? ggml_backend_b200_plan_launch_count(h->backend) : -1; }
int64_t ref_chain_kernel_launches(ref_chain *h) { return ggml_backend_is_b200(h->backend) ? ggml_backend_b200_launch_count(h->backend) : -1; }
#endif
