/* Host-side cost of one glue-operator launch through the C ABI: N x b200_op_scale on a 4 KB tensor, enqueue time and total time.
 * build: gcc -O2 -Iinclude tools/launch_overhead.c -Lggml-imax_b200/lib -lggml_b200 -Wl,-rpath,$PWD/ggml-imax_b200/lib -o tools/_build/launch_overhead */
#define _GNU_SOURCE
#include <stdio.h>
#include <string.h>
#include <time.h>
#include "ggml_b200.h"

static double now_us(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec / 1e3;
}

int main(void) {
    b200_ctx *ctx;
    if (b200_ctx_create(0, &ctx) != B200_OK) { fprintf(stderr, "no device: %s\n", b200_last_error(NULL)); return 1; }
    void *buf;
    b200_malloc(ctx, &buf, 1 << 20);
    b200_memset(ctx, buf, 0, 1 << 20);
    b200_tensor t;
    memset(&t, 0, sizeof(t));
    t.data = buf; t.type = B200_TYPE_F32;
    t.ne[0] = 1024; t.ne[1] = t.ne[2] = t.ne[3] = 1;
    t.nb[0] = 4; t.nb[1] = t.nb[2] = t.nb[3] = 4096;
    for (int rep = 0; rep < 3; rep++) {
        const int N = 20000;
        b200_synchronize(ctx);
        const double t0 = now_us();
        for (int i = 0; i < N; i++) b200_op_scale(ctx, &t, &t, 1.0f);
        const double t1 = now_us();
        b200_synchronize(ctx);
        const double t2 = now_us();
        printf("b200_op_scale x %d: enqueue %.2f us per call, enqueue + drain %.2f us per call\n", N, (t1 - t0) / N, (t2 - t0) / N);
    }
    b200_ctx_destroy(ctx);
    return 0;
}
