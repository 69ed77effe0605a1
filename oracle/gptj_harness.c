/*
 * gptj_harness.c -- the north-star model as a WHOLE ggml graph: GPT-J (BASELINE.json's 6B decode case is its mul_mats) with Q4_0 or
 * Q8_0 matrices, random-init weights, a prompt followed by decode steps, every tensor and every op on one backend.  TEST
 * INFRASTRUCTURE (our code against the reference's public API; built by oracle/Makefile into oracle/_ref/, linked with the reference
 * core that has the B200 backend dropped in).
 *
 * Two arms compute the same model on the same tokens:
 *   cpu    ggml_backend_cpu (the reference path, n_threads threads);
 *   b200   the B200 backend: compute tensors from ggml_gallocr on its buffer type, one ggml_backend_graph_compute per step, and for the
 *          decode steps also a ggml_backend_graph_plan (recorded once, replayed as one CUDA graph launch).
 * The graph is GPT-J's as the reference's example computes it (examples/gpt-j/main.cpp:440-586): LayerNorm with the gain and bias
 * broadcast through GGML_OP_REPEAT, separate q / k / v projections without bias, GGML_OP_ROPE (n_rot dimensions, mode 0) on q and k,
 * an F16 KV cache written through strided CPY (v transposed), K*Q and V*softmax on permuted F16 views of the cache, the attention
 * and the MLP both reading the same normalized input and summed into the residual, final LayerNorm, lm_head with bias.  Written from
 * the architecture, not from that file.
 *
 * Shapes are arguments so the same program is a small parity case in the GPU tests and the full 6B model in a profile run:
 *   gptj-harness <q4_0|q8_0> <n_layer> <n_embd> <n_head> <n_rot> <n_vocab> <n_ctx> <n_prompt> <n_decode> <threads> [fuse] [graphs] [taps]
 * taps = 1 adds two extra graph outputs that READ intermediates a fusion would like to skip (the first block's LayerNorm product before
 * its bias, and its MLP pre-activation): the backend has to notice the second reader, compute those nodes, and the taps have to match the CPU's.
 * With n_layer > 2 every block reuses the first block's random matrices (one generation + quantization pass instead of 28; each
 * block still owns its own tensors and its own device memory, so the traffic is the real model's).
 * Prints one JSON line: NMSE of the last token's logits against the CPU arm per step, and the times.
 */
#include "ggml.h"
#include "ggml-alloc.h"
#include "ggml-backend.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_LAYER 64
#define MAX_NODES 8192

extern int64_t ggml_backend_b200_launch_count(ggml_backend_t backend);
extern int64_t ggml_backend_b200_fused_node_count(ggml_backend_t backend);
extern int64_t ggml_backend_b200_graph_plan_kernels(ggml_backend_graph_plan_t plan);
extern int ggml_backend_b200_set_option(ggml_backend_t backend, const char *key, int64_t value);

struct hparams { int n_layer, n_embd, n_head, n_rot, n_vocab, n_ctx; };
struct block {
    struct ggml_tensor *ln_g, *ln_b, *wq, *wk, *wv, *wo, *fc_w, *fc_b, *proj_w, *proj_b;
};
struct model {
    struct hparams hp;
    struct ggml_context *ctx;
    ggml_backend_buffer_t buf;
    struct ggml_tensor *wte, *lnf_g, *lnf_b, *lmh_w, *lmh_b, *mem_k, *mem_v;
    struct block B[MAX_LAYER];
};

static uint64_t rng_state;
static float frand(void) {     /* xorshift64*: uniform in [-1, 1) */
    rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
    return (float)((double)((rng_state * 2685821657736338717ull) >> 11) / 9007199254740992.0 * 2.0 - 1.0);
}

/* host images of the tensors, generated once (slot = which tensor of which distinct block) and uploaded into every arm */
#define MAX_SLOTS 64
static void *slot_data[MAX_SLOTS];
static void upload(struct ggml_tensor *t, int slot, float scale, float offset) {
    if (!slot_data[slot]) {
        const int64_t n = ggml_nelements(t);
        float *f = (float *)malloc((size_t)n * sizeof(float));
        rng_state = 88172645463325252ull + 7919ull * (uint64_t)slot;
        for (int64_t i = 0; i < n; i++) f[i] = offset + scale * frand();
        if (t->type == GGML_TYPE_F32) {
            slot_data[slot] = f;
        } else {
            slot_data[slot] = malloc(ggml_nbytes(t));
            ggml_quantize_chunk(t->type, f, slot_data[slot], 0, t->ne[1], t->ne[0], NULL);
            free(f);
        }
    }
    ggml_backend_tensor_set(t, slot_data[slot], 0, ggml_nbytes(t));
}

static void model_build(struct model *m, const struct hparams *hp, enum ggml_type qtype, ggml_backend_buffer_type_t buft) {
    m->hp = *hp;
    const int E = hp->n_embd;
    struct ggml_init_params ip = { ggml_tensor_overhead() * (size_t)(16 + 12 * hp->n_layer), NULL, true };
    m->ctx = ggml_init(ip);
    struct ggml_context *c = m->ctx;
    m->wte = ggml_new_tensor_2d(c, qtype, E, hp->n_vocab);
    m->lnf_g = ggml_new_tensor_1d(c, GGML_TYPE_F32, E);
    m->lnf_b = ggml_new_tensor_1d(c, GGML_TYPE_F32, E);
    m->lmh_w = ggml_new_tensor_2d(c, qtype, E, hp->n_vocab);
    m->lmh_b = ggml_new_tensor_1d(c, GGML_TYPE_F32, hp->n_vocab);
    m->mem_k = ggml_new_tensor_1d(c, GGML_TYPE_F16, (int64_t)hp->n_layer * hp->n_ctx * E);
    m->mem_v = ggml_new_tensor_1d(c, GGML_TYPE_F16, (int64_t)hp->n_layer * hp->n_ctx * E);
    for (int l = 0; l < hp->n_layer; l++) {
        struct block *b = &m->B[l];
        b->ln_g = ggml_new_tensor_1d(c, GGML_TYPE_F32, E);
        b->ln_b = ggml_new_tensor_1d(c, GGML_TYPE_F32, E);
        b->wq = ggml_new_tensor_2d(c, qtype, E, E);
        b->wk = ggml_new_tensor_2d(c, qtype, E, E);
        b->wv = ggml_new_tensor_2d(c, qtype, E, E);
        b->wo = ggml_new_tensor_2d(c, qtype, E, E);
        b->fc_w = ggml_new_tensor_2d(c, qtype, E, 4 * E);
        b->fc_b = ggml_new_tensor_1d(c, GGML_TYPE_F32, 4 * E);
        b->proj_w = ggml_new_tensor_2d(c, qtype, 4 * E, E);
        b->proj_b = ggml_new_tensor_1d(c, GGML_TYPE_F32, E);
    }
    m->buf = ggml_backend_alloc_ctx_tensors_from_buft(c, buft);
    if (!m->buf) { fprintf(stderr, "gptj harness: buffer allocation failed\n"); exit(2); }
    ggml_backend_buffer_clear(m->buf, 0);          /* the KV cache starts out zeroed */
    const float s1 = 1.7320508f / sqrtf((float)E), s4 = 1.7320508f / sqrtf((float)(4 * E));      /* unit gain per projection */
    upload(m->wte, 0, 1.0f, 0.0f);
    upload(m->lnf_g, 1, 0.1f, 1.0f);
    upload(m->lnf_b, 2, 0.02f, 0.0f);
    upload(m->lmh_w, 3, s1, 0.0f);
    upload(m->lmh_b, 4, 0.02f, 0.0f);
    for (int l = 0; l < hp->n_layer; l++) {
        struct block *b = &m->B[l];
        const int base = 8 + 10 * (l < 2 ? l : l % 2);     /* beyond two blocks: reuse their host images */
        upload(b->ln_g, base + 0, 0.1f, 1.0f);
        upload(b->ln_b, base + 1, 0.02f, 0.0f);
        upload(b->wq, base + 2, s1, 0.0f);
        upload(b->wk, base + 3, s1, 0.0f);
        upload(b->wv, base + 4, s1, 0.0f);
        upload(b->wo, base + 5, 0.5f * s1, 0.0f);
        upload(b->fc_w, base + 6, s1, 0.0f);
        upload(b->fc_b, base + 7, 0.02f, 0.0f);
        upload(b->proj_w, base + 8, 0.5f * s4, 0.0f);
        upload(b->proj_b, base + 9, 0.02f, 0.0f);
    }
}

/* LayerNorm the GPT-J way: gain and bias broadcast explicitly */
static struct ggml_tensor *last_ln_product;       /* the MUL inside the most recent layer_norm (for the taps) */
static struct ggml_tensor *layer_norm(struct ggml_context *c, struct ggml_tensor *x, struct ggml_tensor *g, struct ggml_tensor *b) {
    struct ggml_tensor *n = ggml_norm(c, x, 1e-5f);
    last_ln_product = ggml_mul(c, ggml_repeat(c, g, n), n);
    return ggml_add(c, last_ln_product, ggml_repeat(c, b, n));
}

struct graph {
    struct ggml_context *ctx;
    struct ggml_cgraph *gf;
    struct ggml_tensor *tokens, *positions, *logits, *tap[2];
};

static struct graph build_graph(struct model *m, int n_past, int N, int taps) {
    const struct hparams *hp = &m->hp;
    const int E = hp->n_embd, H = hp->n_head, hd = E / H, T = n_past + N;
    struct graph G;
    struct ggml_init_params ip = { ggml_tensor_overhead() * MAX_NODES + ggml_graph_overhead_custom(MAX_NODES, false), NULL, true };
    struct ggml_context *c = ggml_init(ip);
    G.ctx = c;
    G.gf = ggml_new_graph_custom(c, MAX_NODES, false);
    G.tap[0] = G.tap[1] = NULL;
    G.tokens = ggml_new_tensor_1d(c, GGML_TYPE_I32, N);
    G.positions = ggml_new_tensor_1d(c, GGML_TYPE_I32, N);
    ggml_set_input(G.tokens);
    ggml_set_input(G.positions);
    const size_t e16 = ggml_element_size(m->mem_k);
    struct ggml_tensor *x = ggml_get_rows(c, m->wte, G.tokens);
    for (int l = 0; l < hp->n_layer; l++) {
        struct block *b = &m->B[l];
        struct ggml_tensor *h = layer_norm(c, x, b->ln_g, b->ln_b);
        if (taps && l == 0) { G.tap[0] = ggml_scale(c, last_ln_product, 2.0f); ggml_set_output(G.tap[0]); }
        /* attention: q and k rotated in place, k and v into the cache (v transposed: one cache row per embedding dimension) */
        struct ggml_tensor *q = ggml_rope_inplace(c, ggml_reshape_3d(c, ggml_mul_mat(c, b->wq, h), hd, H, N), G.positions, hp->n_rot, 0, 0);
        struct ggml_tensor *k = ggml_rope_inplace(c, ggml_reshape_3d(c, ggml_mul_mat(c, b->wk, h), hd, H, N), G.positions, hp->n_rot, 0, 0);
        struct ggml_tensor *vt = ggml_transpose(c, ggml_mul_mat(c, b->wv, h));                                          /* [N, E] */
        const size_t layer_off = (size_t)l * hp->n_ctx * E * e16;
        ggml_build_forward_expand(G.gf, ggml_cpy(c, k, ggml_view_1d(c, m->mem_k, (int64_t)N * E, layer_off + (size_t)n_past * E * e16)));
        ggml_build_forward_expand(G.gf, ggml_cpy(c, vt, ggml_view_2d(c, m->mem_v, N, E, (size_t)hp->n_ctx * e16, layer_off + (size_t)n_past * e16)));
        struct ggml_tensor *Q = ggml_permute(c, q, 0, 2, 1, 3);                                                          /* [hd, N, H] */
        struct ggml_tensor *K = ggml_permute(c, ggml_reshape_3d(c, ggml_view_1d(c, m->mem_k, (int64_t)T * E, layer_off), hd, H, T), 0, 2, 1, 3);
        struct ggml_tensor *att = ggml_soft_max_inplace(c, ggml_diag_mask_inf_inplace(c, ggml_scale_inplace(c, ggml_mul_mat(c, K, Q), 1.0f / sqrtf((float)hd)), n_past));
        struct ggml_tensor *V = ggml_view_3d(c, m->mem_v, T, hd, H, (size_t)hp->n_ctx * e16, (size_t)hp->n_ctx * e16 * hd, layer_off);
        struct ggml_tensor *merged = ggml_permute(c, ggml_mul_mat(c, V, att), 0, 2, 1, 3);
        struct ggml_tensor *a = ggml_mul_mat(c, b->wo, ggml_cpy(c, merged, ggml_new_tensor_2d(c, GGML_TYPE_F32, E, N)));
        /* MLP on the same normalized input */
        struct ggml_tensor *f = ggml_mul_mat(c, b->fc_w, h);
        f = ggml_add(c, ggml_repeat(c, b->fc_b, f), f);
        if (taps && l == 0) { G.tap[1] = ggml_scale(c, f, 3.0f); ggml_set_output(G.tap[1]); }
        f = ggml_gelu(c, f);
        f = ggml_mul_mat(c, b->proj_w, f);
        f = ggml_add(c, ggml_repeat(c, b->proj_b, f), f);
        x = ggml_add(c, ggml_add(c, f, a), x);
    }
    x = layer_norm(c, x, m->lnf_g, m->lnf_b);
    x = ggml_mul_mat(c, m->lmh_w, x);
    G.logits = ggml_add(c, ggml_repeat(c, m->lmh_b, x), x);
    ggml_set_output(G.logits);
    ggml_build_forward_expand(G.gf, G.logits);
    for (int t = 0; t < 2; t++)
        if (G.tap[t]) ggml_build_forward_expand(G.gf, G.tap[t]);
    return G;
}

static double nmse(const float *a, const float *b, int64_t n) {
    double num = 0, den = 0;
    for (int64_t i = 0; i < n; i++) { const double d = (double)a[i] - (double)b[i]; num += d * d; den += (double)b[i] * (double)b[i]; }
    return den > 0 ? num / den : num;
}

int main(int argc, char **argv) {
    if (argc < 11) {
        fprintf(stderr, "usage: %s <q4_0|q8_0> n_layer n_embd n_head n_rot n_vocab n_ctx n_prompt n_decode threads [fuse] [graphs]\n", argv[0]);
        return 64;
    }
    const enum ggml_type qtype = strcmp(argv[1], "q8_0") == 0 ? GGML_TYPE_Q8_0 : GGML_TYPE_Q4_0;
    struct hparams hp = { atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]) };
    const int n_prompt = atoi(argv[8]), n_decode = atoi(argv[9]), n_threads = atoi(argv[10]);
    if (hp.n_layer < 1 || hp.n_layer > MAX_LAYER || hp.n_embd % 32 || hp.n_embd % hp.n_head || hp.n_rot > hp.n_embd / hp.n_head || n_prompt + n_decode > hp.n_ctx ||
        n_prompt < 1) {
        fprintf(stderr, "gptj harness: bad shape\n");
        return 64;
    }
    ggml_time_init();
    ggml_backend_t cpu = ggml_backend_cpu_init();
    ggml_backend_cpu_set_n_threads(cpu, n_threads);
    ggml_backend_t gpu = NULL;
    for (size_t i = 0; i < ggml_backend_reg_get_count(); i++)
        if (strncmp(ggml_backend_reg_get_name(i), "B200", 4) == 0) { gpu = ggml_backend_reg_init_backend(i, NULL); break; }
    if (!gpu) { printf("{\"error\": \"no B200 backend in the registry\"}\n"); return 2; }
    if (argc > 11) ggml_backend_b200_set_option(gpu, "fuse", atoi(argv[11]));
    if (argc > 12) ggml_backend_b200_set_option(gpu, "graphs", atoi(argv[12]));
    const int taps = argc > 13 ? atoi(argv[13]) : 0;

    static struct model ma, mc;
    int64_t t0 = ggml_time_us();
    model_build(&ma, &hp, qtype, ggml_backend_cpu_buffer_type());
    model_build(&mc, &hp, qtype, ggml_backend_get_default_buffer_type(gpu));
    const double s_build = (double)(ggml_time_us() - t0) / 1e6;
    ggml_gallocr_t galloc_a = ggml_gallocr_new(ggml_backend_cpu_buffer_type());
    ggml_gallocr_t galloc_c = ggml_gallocr_new(ggml_backend_get_default_buffer_type(gpu));

    int32_t *tokens = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_prompt + n_decode));
    uint64_t ts = 1234;
    for (int i = 0; i < n_prompt + n_decode; i++) { ts = ts * 6364136223846793005ull + 1442695040888963407ull; tokens[i] = (int32_t)((ts >> 33) % (uint64_t)hp.n_vocab); }
    int32_t *pos = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_prompt + 1));
    float *la = (float *)malloc(sizeof(float) * (size_t)hp.n_vocab), *lc = (float *)malloc(sizeof(float) * (size_t)hp.n_vocab),
          *lp = (float *)malloc(sizeof(float) * (size_t)hp.n_vocab);
    double weight_bytes = 0;
    for (struct ggml_tensor *t = ggml_get_first_tensor(mc.ctx); t; t = ggml_get_next_tensor(mc.ctx, t))
        if (ggml_is_quantized(t->type) && t != mc.wte) weight_bytes += (double)ggml_nbytes(t);

    printf("{\"model\": \"gpt-j, %s matrices, random-init\", \"n_layer\": %d, \"n_embd\": %d, \"n_head\": %d, \"n_rot\": %d, \"n_vocab\": %d, \"n_ctx\": %d, "
           "\"prompt_tokens\": %d, \"threads\": %d, \"mul_mat_weight_bytes_per_token\": %.0f, \"model_build_s\": %.1f, \"steps\": [",
           ggml_type_name(qtype), hp.n_layer, hp.n_embd, hp.n_head, hp.n_rot, hp.n_vocab, hp.n_ctx, n_prompt, n_threads, weight_bytes, s_build);
    int n_past = 0, ok = 1;
    for (int step = 0; step <= n_decode; step++) {
        const int N = step == 0 ? n_prompt : 1;
        for (int i = 0; i < N; i++) pos[i] = n_past + i;
        /* arm cpu */
        struct graph Ga = build_graph(&ma, n_past, N, taps);
        if (!ggml_gallocr_alloc_graph(galloc_a, Ga.gf)) { printf("], \"error\": \"gallocr on the CPU buffer type failed\"}\n"); return 3; }
        ggml_backend_tensor_set(Ga.tokens, tokens + n_past, 0, sizeof(int32_t) * (size_t)N);
        ggml_backend_tensor_set(Ga.positions, pos, 0, sizeof(int32_t) * (size_t)N);
        t0 = ggml_time_us();
        ggml_backend_graph_compute(cpu, Ga.gf);
        const double ms_cpu = (double)(ggml_time_us() - t0) / 1e3;
        ggml_backend_tensor_get(Ga.logits, la, (size_t)(N - 1) * hp.n_vocab * sizeof(float), sizeof(float) * (size_t)hp.n_vocab);
        /* arm b200 */
        struct graph Gc = build_graph(&mc, n_past, N, taps);
        if (!ggml_gallocr_alloc_graph(galloc_c, Gc.gf)) { printf("], \"error\": \"gallocr on the B200 buffer type failed\"}\n"); return 5; }
        ggml_backend_tensor_set(Gc.tokens, tokens + n_past, 0, sizeof(int32_t) * (size_t)N);
        ggml_backend_tensor_set(Gc.positions, pos, 0, sizeof(int32_t) * (size_t)N);
        const int64_t l0 = ggml_backend_b200_launch_count(gpu);
        t0 = ggml_time_us();
        const enum ggml_status stc = ggml_backend_graph_compute_async(gpu, Gc.gf);
        const double ms_enqueue = (double)(ggml_time_us() - t0) / 1e3;
        ggml_backend_synchronize(gpu);
        double ms_b200 = (double)(ggml_time_us() - t0) / 1e3;
        if (stc != GGML_STATUS_SUCCESS) { printf("], \"error\": \"B200 graph_compute failed (%d)\"}\n", (int)stc); return 6; }
        long long launches = (long long)(ggml_backend_b200_launch_count(gpu) - l0);
        double ms_b200_first = ms_b200;
        if (N > 1) {       /* the prompt once more: the first call of a new shape also pays for module loading, scratch growth and kernel attributes */
            ggml_backend_tensor_set(Gc.tokens, tokens + n_past, 0, sizeof(int32_t) * (size_t)N);
            ggml_backend_tensor_set(Gc.positions, pos, 0, sizeof(int32_t) * (size_t)N);
            const int64_t l1 = ggml_backend_b200_launch_count(gpu);
            t0 = ggml_time_us();
            if (ggml_backend_graph_compute(gpu, Gc.gf) != GGML_STATUS_SUCCESS) { printf("], \"error\": \"B200 graph_compute (second prompt run) failed\"}\n"); return 6; }
            ms_b200 = (double)(ggml_time_us() - t0) / 1e3;
            launches = (long long)(ggml_backend_b200_launch_count(gpu) - l1);
        }
        ggml_backend_tensor_get(Gc.logits, lc, (size_t)(N - 1) * hp.n_vocab * sizeof(float), sizeof(float) * (size_t)hp.n_vocab);
        /* the same decode step as a graph plan: node by node once, recorded once, then replayed */
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
            ggml_backend_tensor_get(Gc.logits, lp, 0, sizeof(float) * (size_t)hp.n_vocab);
            plan_equal = memcmp(lp, lc, sizeof(float) * (size_t)hp.n_vocab) == 0;
            ggml_backend_graph_plan_free(gpu, gplan);
            if (!plan_equal) ok = 0;
        }
        double e_tap = 0.0;
        for (int t = 0; t < 2 && taps; t++) {
            const size_t nb = ggml_nbytes(Ga.tap[t]);
            float *ta = (float *)malloc(nb), *tc = (float *)malloc(nb);
            ggml_backend_tensor_get(Ga.tap[t], ta, 0, nb);
            ggml_backend_tensor_get(Gc.tap[t], tc, 0, nb);
            const double et = nmse(tc, ta, (int64_t)(nb / sizeof(float)));
            if (et > e_tap) e_tap = et;
            free(ta);
            free(tc);
        }
        if (!(e_tap <= 1e-6)) ok = 0;
        const double e = nmse(lc, la, hp.n_vocab);
        int fin = 1;
        for (int i = 0; i < hp.n_vocab; i++) if (!isfinite(lc[i]) || !isfinite(la[i])) fin = 0;
        if (!(e <= 5e-4) || !fin) ok = 0;
        printf("%s{\"n_past\": %d, \"n\": %d, \"taps_nmse_vs_cpu\": %.3e, \"logits_nmse_vs_cpu\": %.3e, \"finite\": %s, \"ms_cpu\": %.2f, \"ms_b200\": %.3f, \"ms_b200_first_call\": %.3f, \"ms_b200_enqueue\": %.3f, "
               "\"b200_launches\": %lld, \"graph_nodes\": %d, \"ms_b200_graph_plan\": %.3f, \"graph_plan_kernels\": %lld, \"graph_plan_equals_node_by_node\": %s}",
               step ? ", " : "", n_past, N, e_tap, e, fin ? "true" : "false", ms_cpu, ms_b200, ms_b200_first, ms_enqueue, launches, Gc.gf->n_nodes, ms_plan, plan_kernels,
               plan_equal ? "true" : "false");
        fflush(stdout);
        n_past += N;
        ggml_free(Ga.ctx);
        ggml_free(Gc.ctx);
    }
    printf("], \"b200_fused_nodes_total\": %lld, \"ok\": %s}\n", (long long)ggml_backend_b200_fused_node_count(gpu), ok ? "true" : "false");
    return ok ? 0 : 1;
}
