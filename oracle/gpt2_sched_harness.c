/*
 * gpt2_sched_harness.c -- BASELINE.json configs[2]: GPT-2 117M, Q4_0 (or Q8_0) matrices, random-init weights, a 128-token prompt
 * followed by decode steps, on the reference's own graph machinery.  TEST INFRASTRUCTURE (our code against the reference's public
 * API; built by oracle/Makefile into oracle/_ref/, linked with the reference core that has the B200 backend dropped in).
 *
 * Three arms compute the same model on the same tokens:
 *   cpu    every tensor and every op on ggml_backend_cpu (the reference path);
 *   b200   every tensor and every op on the B200 backend, the way examples/gpt-2/main-backend.cpp:744-768 runs a model: compute
 *          tensors from ggml_gallocr on the backend's default buffer type, ONE ggml_backend_graph_compute per step;
 *   sched  ggml_backend_sched over {B200, CPU} (src/ggml-backend.c:1683-1830, the way examples/gpt-2/main-sched.cpp:869-936 drives
 *          it): every GGML_OP_MUL_MAT with a quantized weight is pinned to the B200 backend (ggml_backend_sched_set_tensor_backend,
 *          src/ggml-backend.c:1874), the glue ops (GET_ROWS, ADD, MUL, NORM, SCALE, DIAG_MASK_INF, SOFT_MAX, GELU, CPY / CONT, the
 *          F32 attention mul_mats) stay on the CPU backend; the quantized matrices -- including wte, which is BOTH the token
 *          embedding read by GET_ROWS on the CPU and the tied lm_head multiplied on the GPU -- live in a B200 buffer in the
 *          repacked layout, so the scheduler's copy of wte to the CPU goes through the exact un-repack of get_tensor.
 * The model follows GPT-2's architecture as the reference's example builds it (examples/gpt-2/main-backend.cpp:442-640: learned
 * positions, pre-LN blocks, fused qkv projection, F32 KV cache, GELU MLP, final LN, tied head); written from the architecture,
 * not from that file.  Prints one JSON line: NMSE of the logits between the arms for the prompt and for each decode step.
 */
#include "ggml.h"
#include "ggml-alloc.h"
#include "ggml-backend.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define N_VOCAB 50257
#define N_CTX   1024
#define N_EMBD  768
#define N_HEAD  12
#define N_LAYER 12
#define MAX_NODES 4096

/* instrumentation of the B200 backend (ggml-imax_b200/host/ggml-b200.h) */
extern int64_t ggml_backend_b200_launch_count(ggml_backend_t backend);
extern int64_t ggml_backend_b200_fused_node_count(ggml_backend_t backend);
extern int64_t ggml_backend_b200_graph_plan_kernels(ggml_backend_graph_plan_t plan);
extern int ggml_backend_b200_set_option(ggml_backend_t backend, const char *key, int64_t value);

struct layer {
    struct ggml_tensor *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    struct ggml_tensor *qkv_w, *qkv_b, *proj_w, *proj_b, *fc_w, *fc_b, *out_w, *out_b;
};
struct model {
    struct ggml_context *ctx_q, *ctx_f;      /* quantized matrices / everything else */
    ggml_backend_buffer_t buf_q, buf_f;
    struct ggml_tensor *wte, *wpe, *lnf_g, *lnf_b, *mem_k, *mem_v;
    struct layer L[N_LAYER];
};

static uint64_t rng_state = 88172645463325252ull;
static float frand(void) {     /* xorshift64*: uniform in [-1, 1) */
    rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
    return (float)((double)((rng_state * 2685821657736338717ull) >> 11) / 9007199254740992.0 * 2.0 - 1.0);
}

/* fills a tensor with scale * U(-1,1) (+ offset), quantizing with ggml_quantize_chunk when its type asks for it */
static void fill(struct ggml_tensor *t, float scale, float offset) {
    const int64_t n = ggml_nelements(t);
    float *f = (float *)malloc((size_t)n * sizeof(float));
    for (int64_t i = 0; i < n; i++) f[i] = offset + scale * frand();
    if (t->type == GGML_TYPE_F32) {
        ggml_backend_tensor_set(t, f, 0, ggml_nbytes(t));
    } else {
        void *q = malloc(ggml_nbytes(t));
        ggml_quantize_chunk(t->type, f, q, 0, t->ne[1], t->ne[0], NULL);
        ggml_backend_tensor_set(t, q, 0, ggml_nbytes(t));
        free(q);
    }
    free(f);
}

static void model_build(struct model *m, enum ggml_type qtype, ggml_backend_buffer_type_t buft_q, ggml_backend_buffer_type_t buft_f) {
    struct ggml_init_params ip = { ggml_tensor_overhead() * 256, NULL, true };
    m->ctx_q = ggml_init(ip);
    m->ctx_f = ggml_init(ip);
    m->wte = ggml_new_tensor_2d(m->ctx_q, qtype, N_EMBD, N_VOCAB);
    m->wpe = ggml_new_tensor_2d(m->ctx_f, GGML_TYPE_F32, N_EMBD, N_CTX);
    m->lnf_g = ggml_new_tensor_1d(m->ctx_f, GGML_TYPE_F32, N_EMBD);
    m->lnf_b = ggml_new_tensor_1d(m->ctx_f, GGML_TYPE_F32, N_EMBD);
    m->mem_k = ggml_new_tensor_1d(m->ctx_f, GGML_TYPE_F32, (int64_t)N_LAYER * N_CTX * N_EMBD);
    m->mem_v = ggml_new_tensor_1d(m->ctx_f, GGML_TYPE_F32, (int64_t)N_LAYER * N_CTX * N_EMBD);
    for (int l = 0; l < N_LAYER; l++) {
        struct layer *L = &m->L[l];
        L->ln1_g = ggml_new_tensor_1d(m->ctx_f, GGML_TYPE_F32, N_EMBD);
        L->ln1_b = ggml_new_tensor_1d(m->ctx_f, GGML_TYPE_F32, N_EMBD);
        L->ln2_g = ggml_new_tensor_1d(m->ctx_f, GGML_TYPE_F32, N_EMBD);
        L->ln2_b = ggml_new_tensor_1d(m->ctx_f, GGML_TYPE_F32, N_EMBD);
        L->qkv_w = ggml_new_tensor_2d(m->ctx_q, qtype, N_EMBD, 3 * N_EMBD);
        L->qkv_b = ggml_new_tensor_1d(m->ctx_f, GGML_TYPE_F32, 3 * N_EMBD);
        L->proj_w = ggml_new_tensor_2d(m->ctx_q, qtype, N_EMBD, N_EMBD);
        L->proj_b = ggml_new_tensor_1d(m->ctx_f, GGML_TYPE_F32, N_EMBD);
        L->fc_w = ggml_new_tensor_2d(m->ctx_q, qtype, N_EMBD, 4 * N_EMBD);
        L->fc_b = ggml_new_tensor_1d(m->ctx_f, GGML_TYPE_F32, 4 * N_EMBD);
        L->out_w = ggml_new_tensor_2d(m->ctx_q, qtype, 4 * N_EMBD, N_EMBD);
        L->out_b = ggml_new_tensor_1d(m->ctx_f, GGML_TYPE_F32, N_EMBD);
    }
    m->buf_q = ggml_backend_alloc_ctx_tensors_from_buft(m->ctx_q, buft_q);
    m->buf_f = ggml_backend_alloc_ctx_tensors_from_buft(m->ctx_f, buft_f);
    if (!m->buf_q || !m->buf_f) { fprintf(stderr, "gpt2 harness: buffer allocation failed\n"); exit(2); }
    ggml_backend_buffer_set_usage(m->buf_q, GGML_BACKEND_BUFFER_USAGE_WEIGHTS);
    ggml_backend_buffer_clear(m->buf_f, 0);        /* the KV cache starts out zeroed */
    /* the same pseudo-random stream in both arms: identical weights */
    rng_state = 88172645463325252ull;
    fill(m->wte, 0.04f, 0.0f);
    fill(m->wpe, 0.02f, 0.0f);
    fill(m->lnf_g, 0.1f, 1.0f);
    fill(m->lnf_b, 0.02f, 0.0f);
    for (int l = 0; l < N_LAYER; l++) {
        struct layer *L = &m->L[l];
        fill(L->ln1_g, 0.1f, 1.0f); fill(L->ln1_b, 0.02f, 0.0f);
        fill(L->ln2_g, 0.1f, 1.0f); fill(L->ln2_b, 0.02f, 0.0f);
        fill(L->qkv_w, 0.06f, 0.0f); fill(L->qkv_b, 0.02f, 0.0f);
        fill(L->proj_w, 0.04f, 0.0f); fill(L->proj_b, 0.02f, 0.0f);
        fill(L->fc_w, 0.06f, 0.0f); fill(L->fc_b, 0.02f, 0.0f);
        fill(L->out_w, 0.03f, 0.0f); fill(L->out_b, 0.02f, 0.0f);
    }
}

/* LayerNorm with gain and bias */
static struct ggml_tensor *layer_norm(struct ggml_context *c, struct ggml_tensor *x, struct ggml_tensor *g, struct ggml_tensor *b) {
    return ggml_add(c, ggml_mul(c, ggml_norm(c, x, 1e-5f), g), b);
}

struct graph {
    struct ggml_context *ctx;
    struct ggml_cgraph *gf;
    struct ggml_tensor *tokens, *positions, *logits;
};

static struct graph build_graph(struct model *m, int n_past, int N) {
    struct graph G;
    struct ggml_init_params ip = { ggml_tensor_overhead() * MAX_NODES + ggml_graph_overhead_custom(MAX_NODES, false), NULL, true };
    struct ggml_context *c = ggml_init(ip);
    G.ctx = c;
    G.gf = ggml_new_graph_custom(c, MAX_NODES, false);
    G.tokens = ggml_new_tensor_1d(c, GGML_TYPE_I32, N);
    G.positions = ggml_new_tensor_1d(c, GGML_TYPE_I32, N);
    ggml_set_input(G.tokens);
    ggml_set_input(G.positions);
    const int hd = N_EMBD / N_HEAD;
    const size_t esz = sizeof(float);
    struct ggml_tensor *x = ggml_add(c, ggml_get_rows(c, m->wte, G.tokens), ggml_get_rows(c, m->wpe, G.positions));
    for (int l = 0; l < N_LAYER; l++) {
        struct layer *L = &m->L[l];
        /* attention */
        struct ggml_tensor *h = layer_norm(c, x, L->ln1_g, L->ln1_b);
        struct ggml_tensor *qkv = ggml_add(c, ggml_mul_mat(c, L->qkv_w, h), L->qkv_b);                 /* [3E, N] */
        struct ggml_tensor *q = ggml_view_2d(c, qkv, N_EMBD, N, qkv->nb[1], 0);
        struct ggml_tensor *k = ggml_view_2d(c, qkv, N_EMBD, N, qkv->nb[1], (size_t)N_EMBD * esz);
        struct ggml_tensor *v = ggml_view_2d(c, qkv, N_EMBD, N, qkv->nb[1], (size_t)2 * N_EMBD * esz);
        const size_t layer_off = (size_t)l * N_CTX * N_EMBD * esz;
        ggml_build_forward_expand(G.gf, ggml_cpy(c, k, ggml_view_1d(c, m->mem_k, (int64_t)N * N_EMBD, layer_off + (size_t)n_past * N_EMBD * esz)));
        ggml_build_forward_expand(G.gf, ggml_cpy(c, v, ggml_view_1d(c, m->mem_v, (int64_t)N * N_EMBD, layer_off + (size_t)n_past * N_EMBD * esz)));
        struct ggml_tensor *Q = ggml_permute(c, ggml_cont_3d(c, q, hd, N_HEAD, N), 0, 2, 1, 3);          /* [hd, N, heads] */
        struct ggml_tensor *K = ggml_permute(c, ggml_reshape_3d(c, ggml_view_1d(c, m->mem_k, (int64_t)(n_past + N) * N_EMBD, layer_off), hd, N_HEAD, n_past + N), 0, 2, 1, 3);
        struct ggml_tensor *att = ggml_soft_max(c, ggml_diag_mask_inf(c, ggml_scale(c, ggml_mul_mat(c, K, Q), 1.0f / sqrtf((float)hd)), n_past));
        struct ggml_tensor *Vt = ggml_cont_3d(c, ggml_permute(c, ggml_reshape_3d(c, ggml_view_1d(c, m->mem_v, (int64_t)(n_past + N) * N_EMBD, layer_off), hd, N_HEAD, n_past + N), 1, 2, 0, 3),
                                              n_past + N, hd, N_HEAD);
        struct ggml_tensor *ctxv = ggml_cont_2d(c, ggml_permute(c, ggml_mul_mat(c, Vt, att), 0, 2, 1, 3), N_EMBD, N);
        x = ggml_add(c, ggml_add(c, ggml_mul_mat(c, L->proj_w, ctxv), L->proj_b), x);
        /* MLP */
        h = layer_norm(c, x, L->ln2_g, L->ln2_b);
        h = ggml_gelu(c, ggml_add(c, ggml_mul_mat(c, L->fc_w, h), L->fc_b));
        x = ggml_add(c, ggml_add(c, ggml_mul_mat(c, L->out_w, h), L->out_b), x);
    }
    x = layer_norm(c, x, m->lnf_g, m->lnf_b);
    G.logits = ggml_mul_mat(c, m->wte, x);          /* the tied head */
    ggml_set_output(G.logits);
    ggml_build_forward_expand(G.gf, G.logits);
    return G;
}

static double nmse(const float *a, const float *b, int64_t n) {
    double num = 0, den = 0;
    for (int64_t i = 0; i < n; i++) { const double d = (double)a[i] - (double)b[i]; num += d * d; den += (double)b[i] * (double)b[i]; }
    return den > 0 ? num / den : num;
}

static bool is_quantized_mul_mat(const struct ggml_tensor *t) {
    return t->op == GGML_OP_MUL_MAT && t->src[0] && ggml_is_quantized(t->src[0]->type);
}

int main(int argc, char **argv) {
    const enum ggml_type qtype = (argc > 1 && strcmp(argv[1], "q8_0") == 0) ? GGML_TYPE_Q8_0 : GGML_TYPE_Q4_0;
    const int n_prompt = argc > 2 ? atoi(argv[2]) : 128, n_decode = argc > 3 ? atoi(argv[3]) : 3, n_threads = argc > 4 ? atoi(argv[4]) : 8;
    ggml_time_init();
    ggml_backend_t cpu_a = ggml_backend_cpu_init(), cpu_b = ggml_backend_cpu_init();
    ggml_backend_cpu_set_n_threads(cpu_a, n_threads);
    ggml_backend_cpu_set_n_threads(cpu_b, n_threads);
    ggml_backend_t gpu = NULL;
    for (size_t i = 0; i < ggml_backend_reg_get_count(); i++)
        if (strncmp(ggml_backend_reg_get_name(i), "B200", 4) == 0) { gpu = ggml_backend_reg_init_backend(i, NULL); break; }
    if (!gpu) { printf("{\"error\": \"no B200 backend in the registry\"}\n"); return 2; }

    static struct model ma, mb, mc;
    model_build(&ma, qtype, ggml_backend_cpu_buffer_type(), ggml_backend_cpu_buffer_type());                          /* arm cpu   */
    model_build(&mb, qtype, ggml_backend_get_default_buffer_type(gpu), ggml_backend_cpu_buffer_type());              /* arm sched */

    model_build(&mc, qtype, ggml_backend_get_default_buffer_type(gpu), ggml_backend_get_default_buffer_type(gpu));   /* arm b200  */
    if (argc > 5) ggml_backend_b200_set_option(gpu, "fuse", atoi(argv[5]));
    ggml_gallocr_t galloc_c = ggml_gallocr_new(ggml_backend_get_default_buffer_type(gpu));
    float *lc = (float *)malloc(sizeof(float) * N_VOCAB);
    int64_t launches_c = 0, nodes_c = 0;

    ggml_backend_t backends[2] = { gpu, cpu_b };
    /* argv[6] = 1: the scheduler's pipelined mode (several copies of the split inputs, ggml_backend_event_* and the asynchronous tensor accessors) */
    const bool parallel = argc > 6 && atoi(argv[6]) != 0;
    ggml_backend_sched_t sched = ggml_backend_sched_new(backends, NULL, 2, MAX_NODES, parallel);
    ggml_gallocr_t galloc = ggml_gallocr_new(ggml_backend_cpu_buffer_type());

    int32_t *tokens = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_prompt + n_decode));
    uint64_t ts = 1234;
    for (int i = 0; i < n_prompt + n_decode; i++) { ts = ts * 6364136223846793005ull + 1442695040888963407ull; tokens[i] = (int32_t)((ts >> 33) % N_VOCAB); }
    float *la = (float *)malloc(sizeof(float) * N_VOCAB), *lb = (float *)malloc(sizeof(float) * N_VOCAB);

    printf("{\"model\": \"gpt-2 117M, %s matrices, random-init\", \"prompt_tokens\": %d, \"threads\": %d, \"steps\": [", ggml_type_name(qtype), n_prompt, n_threads);
    int n_past = 0, ok = 1, qmm_nodes = 0, gpu_splits_last = 0;
    for (int step = 0; step <= n_decode; step++) {
        const int N = step == 0 ? n_prompt : 1;
        int32_t pos[1024];
        for (int i = 0; i < N; i++) pos[i] = n_past + i;
        /* arm cpu */
        struct graph Ga = build_graph(&ma, n_past, N);
        ggml_gallocr_alloc_graph(galloc, Ga.gf);
        ggml_backend_tensor_set(Ga.tokens, tokens + n_past, 0, sizeof(int32_t) * (size_t)N);
        ggml_backend_tensor_set(Ga.positions, pos, 0, sizeof(int32_t) * (size_t)N);
        int64_t t0 = ggml_time_us();
        ggml_backend_graph_compute(cpu_a, Ga.gf);
        const double ms_cpu = (double)(ggml_time_us() - t0) / 1e3;
        ggml_backend_tensor_get(Ga.logits, la, (size_t)(N - 1) * N_VOCAB * sizeof(float), sizeof(float) * N_VOCAB);
        /* arm sched: quantized mul_mats pinned to the B200 backend, everything else to the CPU backend */
        struct graph Gb = build_graph(&mb, n_past, N);
        qmm_nodes = 0;
        ggml_backend_sched_reset(sched);                  /* (clears the assignments of the previous graph) */
        for (int i = 0; i < Gb.gf->n_nodes; i++) {
            struct ggml_tensor *node = Gb.gf->nodes[i];
            const bool on_gpu = is_quantized_mul_mat(node);
            qmm_nodes += on_gpu;
            ggml_backend_sched_set_tensor_backend(sched, node, on_gpu ? gpu : cpu_b);
        }
        if (!ggml_backend_sched_alloc_graph(sched, Gb.gf)) { printf("], \"error\": \"sched alloc failed\"}\n"); return 3; }
        ggml_backend_tensor_set(Gb.tokens, tokens + n_past, 0, sizeof(int32_t) * (size_t)N);
        ggml_backend_tensor_set(Gb.positions, pos, 0, sizeof(int32_t) * (size_t)N);
        t0 = ggml_time_us();
        const enum ggml_status st = ggml_backend_sched_graph_compute(sched, Gb.gf);
        const double ms_sched = (double)(ggml_time_us() - t0) / 1e3;
        if (st != GGML_STATUS_SUCCESS) { printf("], \"error\": \"sched compute failed (%d)\"}\n", (int)st); return 4; }
        gpu_splits_last = ggml_backend_sched_get_n_splits(sched);
        ggml_backend_tensor_get(Gb.logits, lb, (size_t)(N - 1) * N_VOCAB * sizeof(float), sizeof(float) * N_VOCAB);
        /* arm b200: the whole graph on the B200 backend */
        struct graph Gc = build_graph(&mc, n_past, N);
        if (!ggml_gallocr_alloc_graph(galloc_c, Gc.gf)) { printf("], \"error\": \"gallocr on the B200 buffer type failed\"}\n"); return 5; }
        ggml_backend_tensor_set(Gc.tokens, tokens + n_past, 0, sizeof(int32_t) * (size_t)N);
        ggml_backend_tensor_set(Gc.positions, pos, 0, sizeof(int32_t) * (size_t)N);
        const int64_t l0 = ggml_backend_b200_launch_count(gpu);
        t0 = ggml_time_us();
        const enum ggml_status stc = ggml_backend_graph_compute_async(gpu, Gc.gf);
        const double ms_b200_enqueue = (double)(ggml_time_us() - t0) / 1e3;      /* host time to issue the step's launches */
        ggml_backend_synchronize(gpu);
        const double ms_b200 = (double)(ggml_time_us() - t0) / 1e3;
        if (stc != GGML_STATUS_SUCCESS) { printf("], \"error\": \"B200 graph_compute failed (%d)\"}\n", (int)stc); return 6; }
        launches_c = ggml_backend_b200_launch_count(gpu) - l0;
        nodes_c = Gc.gf->n_nodes;
        ggml_backend_tensor_get(Gc.logits, lc, (size_t)(N - 1) * N_VOCAB * sizeof(float), sizeof(float) * N_VOCAB);
        /* the same step as a ggml_backend_graph_plan: computed node by node once, recorded once, then replayed (one CUDA graph launch) */
        double ms_plan = 0.0;
        long long plan_kernels = 0;
        int plan_equal = 1;
        if (N == 1) {
            ggml_backend_graph_plan_t gplan = ggml_backend_graph_plan_create(gpu, Gc.gf);
            const int reps = 20;
            for (int r = 0; r < 3 + reps; r++) {
                if (r == 3) { ggml_backend_synchronize(gpu); t0 = ggml_time_us(); }
                ggml_backend_tensor_set_async(gpu, Gc.tokens, tokens + n_past, 0, sizeof(int32_t));       /* on the backend's stream, in front of the replay */
                ggml_backend_tensor_set_async(gpu, Gc.positions, pos, 0, sizeof(int32_t));
                if (ggml_backend_graph_plan_compute(gpu, gplan) != GGML_STATUS_SUCCESS) { printf("], \"error\": \"graph_plan_compute failed\"}\n"); return 7; }
                ggml_backend_synchronize(gpu);
            }
            ms_plan = (double)(ggml_time_us() - t0) / 1e3 / reps;
            plan_kernels = (long long)ggml_backend_b200_graph_plan_kernels(gplan);
            float *lp = (float *)malloc(sizeof(float) * N_VOCAB);
            ggml_backend_tensor_get(Gc.logits, lp, 0, sizeof(float) * N_VOCAB);
            plan_equal = memcmp(lp, lc, sizeof(float) * N_VOCAB) == 0;
            free(lp);
            ggml_backend_graph_plan_free(gpu, gplan);
            if (!plan_equal) ok = 0;
        }
        const double e = nmse(lb, la, N_VOCAB), ec = nmse(lc, la, N_VOCAB);
        int fin = 1;
        for (int i = 0; i < N_VOCAB; i++) if (!isfinite(lb[i]) || !isfinite(lc[i])) fin = 0;
        if (!(e <= 5e-4) || !(ec <= 5e-4) || !fin) ok = 0;
        printf("%s{\"n_past\": %d, \"n\": %d, \"logits_nmse_vs_cpu\": %.3e, \"b200_whole_graph_logits_nmse_vs_cpu\": %.3e, \"finite\": %s, \"ms_cpu\": %.2f, "
               "\"ms_sched_b200\": %.2f, \"ms_b200_whole_graph\": %.3f, \"ms_b200_enqueue\": %.3f, \"b200_launches\": %lld, \"graph_nodes\": %lld, "
               "\"ms_b200_graph_plan\": %.3f, \"graph_plan_kernels\": %lld, \"graph_plan_equals_node_by_node\": %s}", step ? ", " : "", n_past, N, e, ec,
               fin ? "true" : "false", ms_cpu, ms_sched, ms_b200, ms_b200_enqueue, (long long)launches_c, (long long)nodes_c, ms_plan, plan_kernels,
               plan_equal ? "true" : "false");
        n_past += N;
        ggml_free(Ga.ctx);
        ggml_free(Gb.ctx);
        ggml_free(Gc.ctx);
    }
    printf("], \"quantized_mul_mat_nodes_on_b200\": %d, \"graph_splits\": %d, \"b200_fused_nodes_total\": %lld, \"ok\": %s}\n", qmm_nodes, gpu_splits_last,
           (long long)ggml_backend_b200_fused_node_count(gpu), ok ? "true" : "false");
    ggml_backend_sched_free(sched);
    return ok ? 0 : 1;
}
