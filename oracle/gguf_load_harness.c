/*
 * gguf_load_harness.c -- SURVEY.md 8(f)-4: a GGUF checkpoint (docs/gguf.md; written and parsed by the reference's own gguf_* API,
 * src/ggml.c) loaded straight into the B200 backend's repacked device layout through a pinned host staging buffer
 * (ggml_backend_cuda_host_buffer_type, src/ggml-cuda.cu:977-1056 -- here the alias of ggml_backend_b200_host_buffer_type).  TEST
 * INFRASTRUCTURE: our code against the reference's public API, built by oracle/Makefile into oracle/_ref/.
 *
 *   1. writes a GGUF file with the matrices of `blocks` GPT-J-shaped transformer blocks (q, k, v, o: E x E; fc: 4E x E; proj: E x 4E)
 *      in Q4_0, Q8_0 and -- one matrix each -- Q5_0 and IQ4_NL (kept in wire format by the backend), plus F32 bias rows and metadata;
 *   2. LOADER: gguf_init_from_file(no_alloc) -> tensor metadata in a ggml context -> ggml_backend_alloc_ctx_tensors on the B200 backend ->
 *      for every tensor: read() its bytes from the file into the pinned staging buffer, ggml_backend_tensor_set (Q4_0 / Q8_0 are
 *      repacked into the qs / d planes on the device as they arrive);
 *   3. checks: ggml_backend_tensor_get of every tensor returns the file's bytes exactly; every matrix multiplied by a random
 *      activation on the B200 backend matches the reference CPU backend computing on the file's bytes (NMSE <= 5e-4);
 *   4. reports the load rate (file already in the page cache: read + host-to-device copy + repack).
 *
 *   gguf-load-harness <n_embd> <blocks> <path>
 */
#include "ggml.h"
#include "ggml-alloc.h"
#include "ggml-backend.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

extern ggml_backend_buffer_type_t ggml_backend_cuda_host_buffer_type(void);

static uint64_t rng_state = 88172645463325252ull;
static float frand(void) {
    rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
    return (float)((double)((rng_state * 2685821657736338717ull) >> 11) / 9007199254740992.0 * 2.0 - 1.0);
}
static double nmse(const float *a, const float *b, int64_t n) {
    double num = 0, den = 0;
    for (int64_t i = 0; i < n; i++) { const double d = (double)a[i] - (double)b[i]; num += d * d; den += (double)b[i] * (double)b[i]; }
    return den > 0 ? num / den : num;
}

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s n_embd blocks path\n", argv[0]); return 64; }
    const int E = atoi(argv[1]), blocks = atoi(argv[2]);
    const char *path = argv[3];
    if (E % 32 || E < 32 || blocks < 1 || blocks > 64) { fprintf(stderr, "bad shape\n"); return 64; }
    ggml_time_init();

    /* ---- 1. write the file --------------------------------------------------------------------------------------------- */
    size_t n_tensors = 0, file_payload = 0;
    {
        const size_t per_block = (size_t)12 * E * E;       /* elements */
        struct ggml_init_params ip = { ggml_tensor_overhead() * (size_t)(8 * blocks + 8) + per_block * blocks * 2 + ((size_t)16 << 20), NULL, false };
        struct ggml_context *wc = ggml_init(ip);
        if (!wc) { printf("{\"error\": \"host context for the writer\"}\n"); return 2; }
        struct gguf_context *gw = gguf_init_empty();
        gguf_set_val_str(gw, "general.architecture", "gptj");
        gguf_set_val_u32(gw, "gptj.embedding_length", (uint32_t)E);
        gguf_set_val_u32(gw, "gptj.block_count", (uint32_t)blocks);
        float *tmp = (float *)malloc(sizeof(float) * (size_t)4 * E * E);
        for (int l = 0; l < blocks; l++) {
            const char *names[6] = { "attn_q", "attn_k", "attn_v", "attn_output", "ffn_up", "ffn_down" };
            for (int w = 0; w < 6; w++) {
                const int64_t k = w == 5 ? 4 * E : E, m = w == 4 ? 4 * E : E;
                enum ggml_type type = (l + w) % 2 ? GGML_TYPE_Q8_0 : GGML_TYPE_Q4_0;
                if (l == 0 && w == 1) type = GGML_TYPE_Q5_0;
                if (l == 0 && w == 2) type = GGML_TYPE_IQ4_NL;
                struct ggml_tensor *t = ggml_new_tensor_2d(wc, type, k, m);
                char name[64];
                snprintf(name, sizeof(name), "blk.%d.%s.weight", l, names[w]);
                ggml_set_name(t, name);
                const float s = 1.7320508f / sqrtf((float)k);
                for (int64_t i = 0; i < k * m; i++) tmp[i] = s * frand();
                ggml_quantize_chunk(type, tmp, t->data, 0, m, k, NULL);
                gguf_add_tensor(gw, t);
                n_tensors++;
                file_payload += ggml_nbytes(t);
            }
            struct ggml_tensor *b = ggml_new_tensor_1d(wc, GGML_TYPE_F32, 4 * E);
            char name[64];
            snprintf(name, sizeof(name), "blk.%d.ffn_up.bias", l);
            ggml_set_name(b, name);
            for (int i = 0; i < 4 * E; i++) ((float *)b->data)[i] = 0.02f * frand();
            gguf_add_tensor(gw, b);
            n_tensors++;
            file_payload += ggml_nbytes(b);
        }
        free(tmp);
        gguf_write_to_file(gw, path, false);
        gguf_free(gw);
        ggml_free(wc);
    }

    /* ---- backends ------------------------------------------------------------------------------------------------------- */
    ggml_backend_t cpu = ggml_backend_cpu_init();
    ggml_backend_cpu_set_n_threads(cpu, 8);
    ggml_backend_t gpu = NULL;
    for (size_t i = 0; i < ggml_backend_reg_get_count(); i++)
        if (strncmp(ggml_backend_reg_get_name(i), "B200", 4) == 0) { gpu = ggml_backend_reg_init_backend(i, NULL); break; }
    if (!gpu) { printf("{\"error\": \"no B200 backend in the registry\"}\n"); return 2; }

    /* ---- 2. the loader ---------------------------------------------------------------------------------------------------- */
    struct ggml_context *meta = NULL;
    struct gguf_init_params gp = { /*.no_alloc =*/ true, /*.ctx =*/ &meta };
    struct gguf_context *gr = gguf_init_from_file(path, gp);
    if (!gr || !meta) { printf("{\"error\": \"gguf_init_from_file failed\"}\n"); return 3; }
    if ((size_t)gguf_get_n_tensors(gr) != n_tensors || strcmp(gguf_get_val_str(gr, gguf_find_key(gr, "general.architecture")), "gptj") != 0 ||
        gguf_get_val_u32(gr, gguf_find_key(gr, "gptj.embedding_length")) != (uint32_t)E) {
        printf("{\"error\": \"metadata of the file does not read back\"}\n");
        return 3;
    }
    ggml_backend_buffer_t wbuf = ggml_backend_alloc_ctx_tensors(meta, gpu);
    if (!wbuf) { printf("{\"error\": \"device buffer for the weights\"}\n"); return 4; }
    ggml_backend_buffer_set_usage(wbuf, GGML_BACKEND_BUFFER_USAGE_WEIGHTS);
    size_t biggest = 0;
    for (struct ggml_tensor *t = ggml_get_first_tensor(meta); t; t = ggml_get_next_tensor(meta, t))
        if (ggml_nbytes(t) > biggest) biggest = ggml_nbytes(t);
    ggml_backend_buffer_t stage = ggml_backend_buft_alloc_buffer(ggml_backend_cuda_host_buffer_type(), biggest);
    if (!stage) { printf("{\"error\": \"pinned staging buffer\"}\n"); return 4; }
    void *stage_p = ggml_backend_buffer_get_base(stage);
    FILE *f = fopen(path, "rb");
    if (!f) { printf("{\"error\": \"cannot reopen the file\"}\n"); return 3; }
    const size_t data_off = gguf_get_data_offset(gr);
    int64_t t0 = ggml_time_us();
    size_t loaded = 0;
    for (int i = 0; i < gguf_get_n_tensors(gr); i++) {
        struct ggml_tensor *t = ggml_get_tensor(meta, gguf_get_tensor_name(gr, i));
        const size_t nb = ggml_nbytes(t);
        if (fseek(f, (long)(data_off + gguf_get_tensor_offset(gr, i)), SEEK_SET) != 0 || fread(stage_p, 1, nb, f) != nb) {
            printf("{\"error\": \"short read of %s\"}\n", t->name);
            return 3;
        }
        ggml_backend_tensor_set(t, stage_p, 0, nb);
        loaded += nb;
    }
    ggml_backend_synchronize(gpu);
    const double s_load = (double)(ggml_time_us() - t0) / 1e6;

    /* ---- 3. checks ------------------------------------------------------------------------------------------------------ */
    /* the reference's own view of the file: data in host memory */
    struct ggml_context *hostc = NULL;
    struct gguf_init_params gp2 = { /*.no_alloc =*/ false, /*.ctx =*/ &hostc };
    struct gguf_context *gh = gguf_init_from_file(path, gp2);
    if (!gh || !hostc) { printf("{\"error\": \"gguf_init_from_file (host copy) failed\"}\n"); return 3; }
    int exact = 1, mm_ok = 1, n_mm = 0;
    double worst = 0.0;
    void *back = malloc(biggest);
    struct ggml_init_params ip = { ggml_tensor_overhead() * 16 + ggml_graph_overhead(), NULL, true };
    for (struct ggml_tensor *t = ggml_get_first_tensor(meta); t; t = ggml_get_next_tensor(meta, t)) {
        struct ggml_tensor *h = ggml_get_tensor(hostc, t->name);
        if (!h || h->type != t->type || ggml_nbytes(h) != ggml_nbytes(t)) { exact = 0; continue; }
        ggml_backend_tensor_get(t, back, 0, ggml_nbytes(t));
        if (memcmp(back, h->data, ggml_nbytes(t)) != 0) exact = 0;
        if (!ggml_is_quantized(t->type)) continue;
        /* y = W x on both backends, x the same random column */
        const int64_t k = t->ne[0], m = t->ne[1];
        float *x = (float *)malloc(sizeof(float) * (size_t)k), *ya = (float *)malloc(sizeof(float) * (size_t)m), *yc = (float *)malloc(sizeof(float) * (size_t)m);
        for (int64_t i = 0; i < k; i++) x[i] = frand();
        for (int arm = 0; arm < 2; arm++) {
            struct ggml_context *c = ggml_init(ip);
            struct ggml_tensor *xt = ggml_new_tensor_1d(c, GGML_TYPE_F32, k);
            struct ggml_tensor *y = ggml_mul_mat(c, arm == 0 ? h : t, xt);
            struct ggml_cgraph *gf = ggml_new_graph(c);
            ggml_build_forward_expand(gf, y);
            ggml_backend_buffer_t cb = ggml_backend_alloc_ctx_tensors(c, arm == 0 ? cpu : gpu);
            ggml_backend_tensor_set(xt, x, 0, sizeof(float) * (size_t)k);
            if (ggml_backend_graph_compute(arm == 0 ? cpu : gpu, gf) != GGML_STATUS_SUCCESS) mm_ok = 0;
            ggml_backend_tensor_get(y, arm == 0 ? ya : yc, 0, sizeof(float) * (size_t)m);
            ggml_backend_buffer_free(cb);
            ggml_free(c);
        }
        const double e = nmse(yc, ya, m);
        if (e > worst) worst = e;
        if (!(e <= 5e-4)) mm_ok = 0;
        n_mm++;
        free(x); free(ya); free(yc);
    }
    free(back);
    const int ok = exact && mm_ok;
    printf("{\"file\": \"gguf v%d\", \"n_embd\": %d, \"blocks\": %d, \"tensors\": %zu, \"payload_bytes\": %zu, \"loaded_bytes\": %zu, \"load_s\": %.4f, "
           "\"load_GBps\": %.2f, \"tensor_get_equals_file\": %s, \"mul_mats_checked\": %d, \"worst_nmse_vs_cpu\": %.3e, \"ok\": %s}\n",
           gguf_get_version(gr), E, blocks, n_tensors, file_payload, loaded, s_load, (double)loaded / s_load / 1e9, exact ? "true" : "false", n_mm, worst,
           ok ? "true" : "false");
    fclose(f);
    remove(path);
    return ok ? 0 : 1;
}
