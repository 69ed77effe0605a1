/*
 * ggml-b200.c -- the B200 ggml backend: host side, plain C, on top of the thin C ABI in
 * include/ggml_b200.h.  Implements the reference's three plugin vtables
 * (src/ggml-backend-impl.h:18-28 buffer type, :38-48 buffer, :78-117 backend) and registers through
 * ggml_backend_register (:135-137).  Written against the SPI, not derived from src/ggml-cuda.cu.
 *
 * Scope: GGML_OP_MUL_MAT with src0 in {Q4_0, Q8_0}, src1 F32, dst F32.  Everything
 * else is reported through supports_op == false; graph_compute fails loudly on an unsupported node.
 * There is NO CPU fallback in here.
 *
 * Quantized tensors are stored repacked (qs plane + fp16 scale plane inside the same ggml_nbytes, see
 * include/ggml_b200.h); the repack happens in set_tensor and is undone in get_tensor, so callers such as
 * ggml_backend_graph_copy (src/ggml-backend.c:1974-2060) read back the exact wire bytes.
 */
#include "ggml-b200.h"

#include "ggml-backend-impl.h"
#include "ggml_b200.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define B200_BUFFER_ALIGNMENT 256
#define B200_MAX_RUN 8

/* ---- per-device shared state ----------------------------------------------------------------- */

struct b200_device_state {
    b200_ctx *io;                                /* stream + staging used by buffer set/get/clear */
    struct ggml_backend_buffer_type buft;
    bool buft_init;
    char buft_name[32];
};

static struct b200_device_state g_dev[GGML_B200_MAX_DEVICES];

static b200_ctx *b200_io_ctx(int device) {
    if (device < 0 || device >= GGML_B200_MAX_DEVICES) return NULL;
    if (g_dev[device].io == NULL) {
        if (b200_ctx_create(device, &g_dev[device].io) != B200_OK) {
            fprintf(stderr, "ggml-b200: cannot create context on device %d: %s\n", device, b200_last_error(NULL));
            return NULL;
        }
    }
    return g_dev[device].io;
}

#define B200_CHECK(ctx, call)                                                                       \
    do {                                                                                            \
        int rc__ = (call);                                                                          \
        if (rc__ != B200_OK) {                                                                      \
            fprintf(stderr, "ggml-b200: %s failed (%d): %s\n", #call, rc__, b200_last_error(ctx));  \
            GGML_ASSERT(!"ggml-b200 error");                                                        \
        }                                                                                           \
    } while (0)

static bool b200_type_is_repacked(enum ggml_type t) { return t == GGML_TYPE_Q4_0 || t == GGML_TYPE_Q8_0; }
static int64_t b200_wire_bytes(enum ggml_type t) { return t == GGML_TYPE_Q4_0 ? B200_Q4_0_BYTES : B200_Q8_0_BYTES; }

/*
 * Where a (possibly view) quantized tensor lives inside its repacked root tensor.
 * A view must be a contiguous run of whole blocks of a contiguous root.
 */
struct b200_qloc {
    void   *base;          /* device address of the root tensor (start of its qs plane) */
    int64_t total_blocks;  /* blocks in the root tensor */
    int64_t block_off;     /* first block of this tensor */
    int64_t nblocks;       /* blocks in this tensor */
};

static bool b200_locate_quantized(const struct ggml_tensor *t, struct b200_qloc *loc) {
    const struct ggml_tensor *root = t->view_src ? t->view_src : t;
    if (!b200_type_is_repacked(t->type) || root->type != t->type) return false;
    if (!ggml_is_contiguous(t) || !ggml_is_contiguous(root)) return false;
    const int64_t wire = b200_wire_bytes(t->type);
    const size_t offs = t->view_src ? t->view_offs : 0;
    if (offs % wire != 0) return false;
    loc->base = root->data;
    loc->total_blocks = ggml_nelements(root) / B200_QK;
    loc->block_off = (int64_t)(offs / wire);
    loc->nblocks = ggml_nelements(t) / B200_QK;
    return loc->block_off + loc->nblocks <= loc->total_blocks;
}

/* ---- buffer ---------------------------------------------------------------------------------- */

struct b200_buffer_context {
    int   device;
    void *base;
    char  name[32];
};

GGML_CALL static const char *b200_buffer_get_name(ggml_backend_buffer_t buffer) {
    return ((struct b200_buffer_context *)buffer->context)->name;
}

static bool b200_buffer_is_ours(ggml_backend_buffer_t buffer) {
    return buffer && buffer->iface.get_name == b200_buffer_get_name;
}

GGML_CALL static void b200_buffer_free(ggml_backend_buffer_t buffer) {
    struct b200_buffer_context *bc = (struct b200_buffer_context *)buffer->context;
    b200_ctx *io = b200_io_ctx(bc->device);
    if (io) B200_CHECK(io, b200_free(io, bc->base));
    free(bc);
}

GGML_CALL static void *b200_buffer_get_base(ggml_backend_buffer_t buffer) {
    return ((struct b200_buffer_context *)buffer->context)->base;
}

GGML_CALL static void b200_buffer_init_tensor(ggml_backend_buffer_t buffer, struct ggml_tensor *tensor) {
    /* nothing to attach: the repacked layout is a pure function of (type, nelements) */
    GGML_UNUSED(buffer);
    GGML_UNUSED(tensor);
}

GGML_CALL static void b200_buffer_set_tensor(ggml_backend_buffer_t buffer, struct ggml_tensor *tensor, const void *data,
                                             size_t offset, size_t size) {
    struct b200_buffer_context *bc = (struct b200_buffer_context *)buffer->context;
    b200_ctx *io = b200_io_ctx(bc->device);
    GGML_ASSERT(io != NULL);
    if (b200_type_is_repacked(tensor->type)) {
        struct b200_qloc loc;
        const int64_t wire = b200_wire_bytes(tensor->type);
        GGML_ASSERT(b200_locate_quantized(tensor, &loc) && "ggml-b200: quantized tensor must be contiguous");
        GGML_ASSERT(offset % wire == 0 && size % wire == 0 && "ggml-b200: quantized set_tensor must be block aligned");
        B200_CHECK(io, b200_set_quantized(io, (int)tensor->type, loc.base, loc.total_blocks, data,
                                          loc.block_off + (int64_t)(offset / wire), (int64_t)(size / wire)));
    } else {
        B200_CHECK(io, b200_upload(io, (char *)tensor->data + offset, data, size));
    }
}

GGML_CALL static void b200_buffer_get_tensor(ggml_backend_buffer_t buffer, const struct ggml_tensor *tensor, void *data,
                                             size_t offset, size_t size) {
    struct b200_buffer_context *bc = (struct b200_buffer_context *)buffer->context;
    b200_ctx *io = b200_io_ctx(bc->device);
    GGML_ASSERT(io != NULL);
    if (b200_type_is_repacked(tensor->type)) {
        struct b200_qloc loc;
        const int64_t wire = b200_wire_bytes(tensor->type);
        GGML_ASSERT(b200_locate_quantized(tensor, &loc) && "ggml-b200: quantized tensor must be contiguous");
        GGML_ASSERT(offset % wire == 0 && size % wire == 0 && "ggml-b200: quantized get_tensor must be block aligned");
        B200_CHECK(io, b200_get_quantized(io, (int)tensor->type, loc.base, loc.total_blocks, data,
                                          loc.block_off + (int64_t)(offset / wire), (int64_t)(size / wire)));
    } else {
        B200_CHECK(io, b200_download(io, data, (const char *)tensor->data + offset, size));
    }
}

GGML_CALL static bool b200_buffer_cpy_tensor(ggml_backend_buffer_t buffer, const struct ggml_tensor *src, struct ggml_tensor *dst) {
    /* same-device copies only; anything else goes through the core's get+set fallback
     * (src/ggml-backend.c:313-334), which also re-does the repack correctly. */
    struct b200_buffer_context *bc = (struct b200_buffer_context *)buffer->context;
    ggml_backend_buffer_t sbuf = src->view_src ? src->view_src->buffer : src->buffer;
    if (!b200_buffer_is_ours(sbuf)) return false;
    if (((struct b200_buffer_context *)sbuf->context)->device != bc->device) return false;
    if (b200_type_is_repacked(src->type)) {
        /* plane layout is relative to the whole tensor: raw byte copies are only valid tensor-to-tensor */
        if (src->view_src || dst->view_src) return false;
        if (!ggml_is_contiguous(src) || !ggml_is_contiguous(dst)) return false;
    }
    b200_ctx *io = b200_io_ctx(bc->device);
    if (!io) return false;
    B200_CHECK(io, b200_copy_d2d(io, dst->data, src->data, ggml_nbytes(src)));
    B200_CHECK(io, b200_synchronize(io));
    return true;
}

GGML_CALL static void b200_buffer_clear(ggml_backend_buffer_t buffer, uint8_t value) {
    struct b200_buffer_context *bc = (struct b200_buffer_context *)buffer->context;
    b200_ctx *io = b200_io_ctx(bc->device);
    GGML_ASSERT(io != NULL);
    B200_CHECK(io, b200_memset(io, bc->base, value, buffer->size));
}

static struct ggml_backend_buffer_i b200_buffer_interface = {
    /* .get_name    = */ b200_buffer_get_name,
    /* .free_buffer = */ b200_buffer_free,
    /* .get_base    = */ b200_buffer_get_base,
    /* .init_tensor = */ b200_buffer_init_tensor,
    /* .set_tensor  = */ b200_buffer_set_tensor,
    /* .get_tensor  = */ b200_buffer_get_tensor,
    /* .cpy_tensor  = */ b200_buffer_cpy_tensor,
    /* .clear       = */ b200_buffer_clear,
    /* .reset       = */ NULL,
};

/* ---- buffer type ----------------------------------------------------------------------------- */

GGML_CALL static const char *b200_buft_get_name(ggml_backend_buffer_type_t buft) {
    return g_dev[(int)(intptr_t)buft->context].buft_name;
}

GGML_CALL static ggml_backend_buffer_t b200_buft_alloc_buffer(ggml_backend_buffer_type_t buft, size_t size) {
    const int device = (int)(intptr_t)buft->context;
    b200_ctx *io = b200_io_ctx(device);
    if (!io) return NULL;
    struct b200_buffer_context *bc = (struct b200_buffer_context *)calloc(1, sizeof(*bc));
    if (!bc) return NULL;
    bc->device = device;
    snprintf(bc->name, sizeof(bc->name), "%s%d", GGML_B200_NAME, device);
    if (size == 0) size = 1;
    if (b200_malloc(io, &bc->base, size) != B200_OK) {
        fprintf(stderr, "ggml-b200: allocating %.2f MiB on device %d failed: %s\n", size / 1024.0 / 1024.0, device,
                b200_last_error(io));
        free(bc);
        return NULL;
    }
    return ggml_backend_buffer_init(buft, b200_buffer_interface, bc, size);
}

GGML_CALL static size_t b200_buft_get_alignment(ggml_backend_buffer_type_t buft) {
    GGML_UNUSED(buft);
    return B200_BUFFER_ALIGNMENT;
}

GGML_CALL static size_t b200_buft_get_alloc_size(ggml_backend_buffer_type_t buft, const struct ggml_tensor *tensor) {
    /* the repacked planes fit the wire size exactly (16+2 == 18, 32+2 == 34) */
    GGML_UNUSED(buft);
    return ggml_nbytes(tensor);
}

GGML_CALL static bool b200_buft_supports_backend(ggml_backend_buffer_type_t buft, ggml_backend_t backend);

GGML_CALL ggml_backend_buffer_type_t ggml_backend_b200_buffer_type(int device) {
    if (device < 0 || device >= GGML_B200_MAX_DEVICES || device >= b200_device_count()) return NULL;
    struct b200_device_state *ds = &g_dev[device];
    if (!ds->buft_init) {
        snprintf(ds->buft_name, sizeof(ds->buft_name), "%s%d", GGML_B200_NAME, device);
        ds->buft.iface.get_name = b200_buft_get_name;
        ds->buft.iface.alloc_buffer = b200_buft_alloc_buffer;
        ds->buft.iface.get_alignment = b200_buft_get_alignment;
        ds->buft.iface.get_max_size = NULL;
        ds->buft.iface.get_alloc_size = b200_buft_get_alloc_size;
        ds->buft.iface.supports_backend = b200_buft_supports_backend;
        ds->buft.iface.is_host = NULL;
        ds->buft.context = (void *)(intptr_t)device;
        ds->buft_init = true;
    }
    return &ds->buft;
}

/* ---- pinned host buffer type (ggml_backend_cuda_host_buffer_type, src/ggml-cuda.h:31) ---------- */

GGML_CALL static const char *b200_host_buft_name(ggml_backend_buffer_type_t buft) {
    GGML_UNUSED(buft);
    return GGML_B200_NAME "_Host";
}

GGML_CALL static const char *b200_host_buffer_name(ggml_backend_buffer_t buffer) {
    GGML_UNUSED(buffer);
    return GGML_B200_NAME "_Host";
}

GGML_CALL static void b200_host_buffer_free(ggml_backend_buffer_t buffer) { b200_host_free(buffer->context); }

GGML_CALL static ggml_backend_buffer_t b200_host_buft_alloc(ggml_backend_buffer_type_t buft, size_t size) {
    void *p = NULL;
    if (b200_host_malloc(&p, size) != B200_OK) {
        /* like the reference: fall back to pageable host memory for the HOST buffer (src/ggml-cuda.cu:1021-1026) */
        return ggml_backend_buft_alloc_buffer(ggml_backend_cpu_buffer_type(), size);
    }
    ggml_backend_buffer_t buffer = ggml_backend_cpu_buffer_from_ptr(p, size);
    buffer->buft = buft;
    buffer->iface.get_name = b200_host_buffer_name;
    buffer->iface.free_buffer = b200_host_buffer_free;
    return buffer;
}

GGML_CALL ggml_backend_buffer_type_t ggml_backend_b200_host_buffer_type(void) {
    static struct ggml_backend_buffer_type t;
    static bool init = false;
    if (!init) {
        t.iface = ggml_backend_cpu_buffer_type()->iface;
        t.iface.get_name = b200_host_buft_name;
        t.iface.alloc_buffer = b200_host_buft_alloc;
        t.context = NULL;
        init = true;
    }
    return &t;
}

/* ---- backend --------------------------------------------------------------------------------- */

/* a decode plan cached for a cgraph: valid while the graph's MUL_MAT nodes keep their addresses and shapes */
#define B200_PLAN_CACHE 4
struct b200_cached_plan {
    uint64_t   key;        /* hash over (src0, src1, dst addresses, shapes) of the graph's MUL_MAT nodes; 0 = empty slot */
    int        n_nodes;
    b200_plan *plan;       /* NULL: this graph cannot run as a plan (remembered, so it is not analysed again) */
};

struct b200_backend_context {
    int       device;
    b200_ctx *ctx;
    char      name[32];
    struct b200_cached_plan plans[B200_PLAN_CACHE];
    int       plan_next;   /* round-robin eviction */
    int       opt_plans;   /* 0: never build decode plans (ggml_backend_b200_set_option "plans") */
    int64_t   plan_launches;
    int       failed;      /* an asynchronous error surfaced in synchronize (which cannot return one): the next graph_compute reports it */
};

static ggml_guid_t b200_backend_guid(void) {
    static ggml_guid guid = {0xb2, 0x00, 0x51, 0x0a, 0x71, 0x4d, 0x4a, 0x11, 0x9c, 0x3e, 0x67, 0x67, 0x6d, 0x6c, 0x71, 0x6d};
    return &guid;
}

GGML_CALL bool ggml_backend_is_b200(ggml_backend_t backend) {
    return backend != NULL && ggml_guid_matches(backend->guid, b200_backend_guid());
}

GGML_CALL static bool b200_buft_supports_backend(ggml_backend_buffer_type_t buft, ggml_backend_t backend) {
    if (!ggml_backend_is_b200(backend)) return false;
    return ((struct b200_backend_context *)backend->context)->device == (int)(intptr_t)buft->context;
}

GGML_CALL static const char *b200_backend_name(ggml_backend_t backend) {
    return ((struct b200_backend_context *)backend->context)->name;
}

GGML_CALL static void b200_backend_free(ggml_backend_t backend) {
    struct b200_backend_context *bc = (struct b200_backend_context *)backend->context;
    b200_synchronize(bc->ctx);
    for (int i = 0; i < B200_PLAN_CACHE; i++)
        if (bc->plans[i].plan) b200_plan_destroy(bc->plans[i].plan);
    b200_ctx_destroy(bc->ctx);
    free(bc);
    free(backend);
}

GGML_CALL static ggml_backend_buffer_type_t b200_backend_default_buft(ggml_backend_t backend) {
    return ggml_backend_b200_buffer_type(((struct b200_backend_context *)backend->context)->device);
}

GGML_CALL static void b200_backend_synchronize(ggml_backend_t backend) {
    /* ggml_backend_i.synchronize returns nothing (src/ggml-backend-impl.h:90); a failure here (e.g. a decode plan that gave up
     * waiting for a peer) must not abort the host: it is logged, remembered, and returned by the next graph_compute */
    struct b200_backend_context *bc = (struct b200_backend_context *)backend->context;
    const int rc = b200_synchronize(bc->ctx);
    if (rc != B200_OK) {
        fprintf(stderr, "ggml-b200: synchronize failed (%d): %s\n", rc, b200_last_error(bc->ctx));
        bc->failed = 1;
    }
}

static bool b200_op_is_noop(enum ggml_op op) {
    return op == GGML_OP_NONE || op == GGML_OP_RESHAPE || op == GGML_OP_VIEW || op == GGML_OP_PERMUTE || op == GGML_OP_TRANSPOSE;
}

/* the shapes/layouts the kernels take; mirrors the asserts of ggml_compute_forward_mul_mat
 * (src/ggml.c:11832-11845) plus 16-byte alignment of activation rows for 128-bit loads */
static bool b200_mul_mat_supported(const struct ggml_tensor *dst) {
    const struct ggml_tensor *a = dst->src[0], *b = dst->src[1];
    if (!a || !b) return false;
    if (!b200_type_is_repacked(a->type) || b->type != GGML_TYPE_F32 || dst->type != GGML_TYPE_F32) return false;
    struct b200_qloc loc;
    if (!b200_locate_quantized(a, &loc)) return false;
    if (b->nb[0] != sizeof(float)) return false;
    if (b->nb[1] % 16 != 0 || b->nb[2] % 16 != 0 || b->nb[3] % 16 != 0) return false;
    /* activation rows are read with 128-bit loads: a view must start on a 16-byte boundary of its (256-byte aligned) parent;
     * once allocated, the address itself is checked */
    if (b->view_src != NULL && b->view_offs % 16 != 0) return false;
    if (b->data != NULL && ((uintptr_t)b->data & 15) != 0) return false;
    if (!ggml_is_contiguous(dst)) return false;
    if (b->ne[2] % a->ne[2] != 0 || b->ne[3] % a->ne[3] != 0) return false;
    if (b->ne[2] * b->ne[3] > 65535) return false;
    if (a->ne[0] > 131072) return false; /* activation column must fit the GEMV's shared memory */
    return true;
}

GGML_CALL static bool b200_backend_supports_op(ggml_backend_t backend, const struct ggml_tensor *op) {
    GGML_UNUSED(backend);
    if (b200_op_is_noop(op->op)) return true;
    if (op->op == GGML_OP_MUL_MAT) return b200_mul_mat_supported(op);
    return false;
}

static bool b200_fill_mul_mat_args(struct ggml_tensor *dst, b200_mul_mat_args *args) {
    const struct ggml_tensor *a = dst->src[0], *b = dst->src[1];
    if (!b200_mul_mat_supported(dst)) {
        fprintf(stderr, "ggml-b200: MUL_MAT %s x %s not supported by this backend (no CPU fallback)\n",
                a ? ggml_type_name(a->type) : "?", b ? ggml_type_name(b->type) : "?");
        return false;
    }
    struct b200_qloc loc;
    b200_locate_quantized(a, &loc);
    memset(args, 0, sizeof(*args));
    args->type = (int32_t)a->type;
    args->src0_dev = loc.base;
    args->src0_nblocks_total = loc.total_blocks;
    args->src0_block_off = loc.block_off;
    args->ne00 = a->ne[0]; args->ne01 = a->ne[1]; args->ne02 = a->ne[2]; args->ne03 = a->ne[3];
    args->src1_dev = (const float *)b->data;
    args->ne11 = b->ne[1]; args->ne12 = b->ne[2]; args->ne13 = b->ne[3];
    args->nb11 = b->nb[1]; args->nb12 = b->nb[2]; args->nb13 = b->nb[3];
    args->dst_dev = (float *)dst->data;
    return true;
}

/* nodes[0..n) are MUL_MAT nodes that read the SAME src1 tensor and none of which feeds another (their src0 are
 * weights): e.g. the q/k/v projections of a transformer block.  They go down as one batch so that decode-shaped ones
 * share a launch (b200_mul_mat_batch). */
static enum ggml_status b200_compute_mul_mat_run(struct b200_backend_context *bc, struct ggml_tensor **nodes, int n) {
    b200_mul_mat_args args[B200_MAX_RUN];
    for (int i = 0; i < n; i++)
        if (!b200_fill_mul_mat_args(nodes[i], &args[i])) return GGML_STATUS_FAILED;
    const int rc = n == 1 ? b200_mul_mat(bc->ctx, &args[0]) : b200_mul_mat_batch(bc->ctx, args, n);
    if (rc != B200_OK) {
        fprintf(stderr, "ggml-b200: b200_mul_mat failed (%d): %s\n", rc, b200_last_error(bc->ctx));
        return rc == B200_ERR_ALLOC ? GGML_STATUS_ALLOC_FAILED : GGML_STATUS_FAILED;
    }
    return GGML_STATUS_SUCCESS;
}

static bool b200_node_is_decode_mul_mat(const struct ggml_tensor *node) {
    if (node->op != GGML_OP_MUL_MAT) return false;
    const struct ggml_tensor *a = node->src[0], *b = node->src[1];
    return a && b && b->ne[1] == 1 && b->ne[2] == 1 && b->ne[3] == 1 && a->ne[2] == 1 && a->ne[3] == 1 && a->ne[0] % 256 == 0 &&
           a->ne[0] <= 32768 && b200_mul_mat_supported(node);
}

/* [first, last) = a maximal run of compute nodes that are all decode-shaped MUL_MATs (no-op nodes in between are skipped);
 * returns how many such nodes it holds and a hash over their addresses and shapes */
static int b200_decode_run(const struct ggml_cgraph *cgraph, int first, int *last_out, uint64_t *key_out) {
    int n = 0, i = first;
    uint64_t key = 1469598103934665603ull;     /* FNV-1a */
    for (; i < cgraph->n_nodes; i++) {
        struct ggml_tensor *node = cgraph->nodes[i];
        if (ggml_is_empty(node) || b200_op_is_noop(node->op)) continue;
        if (!b200_node_is_decode_mul_mat(node)) break;
        const struct ggml_tensor *a = node->src[0], *b = node->src[1];
        const uint64_t words[6] = {(uint64_t)(uintptr_t)a->data, (uint64_t)(uintptr_t)b->data, (uint64_t)(uintptr_t)node->data,
                                   (uint64_t)a->ne[0], (uint64_t)a->ne[1], (uint64_t)a->type};
        for (int w = 0; w < 6; w++)
            for (int sh = 0; sh < 64; sh += 8) key = (key ^ ((words[w] >> sh) & 0xff)) * 1099511628211ull;
        n++;
    }
    if (key == 0) key = 1;
    *last_out = i;
    if (key_out) *key_out = key;
    return n;
}

/* args of the n decode MUL_MAT nodes in [first, last) -> b200_plan_create.  *out stays NULL when they cannot run as a plan
 * (B200_ERR_UNSUPPORTED); returns -1 on a hard error, 0 otherwise. */
static int b200_build_plan(struct b200_backend_context *bc, const struct ggml_cgraph *cgraph, int first, int last, int n, b200_plan **out) {
    *out = NULL;
    b200_mul_mat_args *args = (b200_mul_mat_args *)malloc(sizeof(b200_mul_mat_args) * (size_t)n);
    if (!args) return -1;
    int k = 0;
    bool ok = true;
    for (int i = first; i < last && ok; i++) {
        struct ggml_tensor *node = cgraph->nodes[i];
        if (ggml_is_empty(node) || b200_op_is_noop(node->op)) continue;
        ok = k < n && b200_fill_mul_mat_args(node, &args[k++]);
    }
    int ret = 0;
    if (ok && k == n) {
        const int rc = b200_plan_create(bc->ctx, args, n, NULL, out);
        if (rc != B200_OK) {
            *out = NULL;
            if (rc != B200_ERR_UNSUPPORTED) {
                fprintf(stderr, "ggml-b200: b200_plan_create failed (%d): %s\n", rc, b200_last_error(bc->ctx));
                ret = -1;
            }
        }
    }
    free(args);
    return ret;
}

/*
 * Decode runs: a maximal run of two or more decode-shaped MUL_MAT nodes (one activation column, 2-D weights) -- a whole
 * mul_mat-only graph, or the part of a mixed graph between two other ops -- goes down as ONE persistent launch (b200_plan_*,
 * include/ggml_b200.h): what ggml_backend_graph_plan_create / _compute would be for this backend, done transparently and cached
 * per run like the reference's CUDA-graph replay (src/ggml-cuda.cu:2461-2709).  Tensors that share memory the way ggml_gallocr
 * arranges them are fine (the plan keeps dead intermediates out of plain memory).  Returns 1 when the run was computed that
 * way, 0 when it has to go node by node, -1 on a hard error.
 */
static int b200_try_run_as_plan(struct b200_backend_context *bc, struct ggml_cgraph *cgraph, int first, int last, int n, uint64_t key) {
    if (!bc->opt_plans || n < 2) return 0;
    struct b200_cached_plan *slot = NULL;
    for (int i = 0; i < B200_PLAN_CACHE; i++)
        if (bc->plans[i].key == key && bc->plans[i].n_nodes == n) slot = &bc->plans[i];
    if (!slot) {
        slot = &bc->plans[bc->plan_next];
        bc->plan_next = (bc->plan_next + 1) % B200_PLAN_CACHE;
        if (slot->plan) {
            b200_synchronize(bc->ctx);
            b200_plan_destroy(slot->plan);
        }
        slot->key = key;
        slot->n_nodes = n;
        slot->plan = NULL;
        if (b200_build_plan(bc, cgraph, first, last, n, &slot->plan) < 0) return -1;
    }
    if (!slot->plan) return 0;
    const int rc = b200_plan_launch(bc->ctx, slot->plan);
    if (rc == B200_ERR_UNSUPPORTED) return 0;          /* the grid cannot be co-resident right now: node by node */
    if (rc != B200_OK) {
        fprintf(stderr, "ggml-b200: b200_plan_launch failed (%d): %s\n", rc, b200_last_error(bc->ctx));
        return -1;
    }
    bc->plan_launches++;
    return 1;
}

static enum ggml_status b200_graph_compute_nodes(struct b200_backend_context *bc, struct ggml_cgraph *cgraph, int first, int last);

GGML_CALL static enum ggml_status b200_backend_graph_compute(ggml_backend_t backend, struct ggml_cgraph *cgraph) {
    struct b200_backend_context *bc = (struct b200_backend_context *)backend->context;
    if (bc->failed) {
        bc->failed = 0;
        return GGML_STATUS_FAILED;
    }
    int i = 0;
    while (i < cgraph->n_nodes) {
        struct ggml_tensor *node = cgraph->nodes[i];
        if (ggml_is_empty(node) || b200_op_is_noop(node->op)) { i++; continue; }
        int last = i;
        uint64_t key = 0;
        const int n = b200_decode_run(cgraph, i, &last, &key);
        if (n >= 2) {
            const int as_plan = b200_try_run_as_plan(bc, cgraph, i, last, n, key);
            if (as_plan < 0) return GGML_STATUS_FAILED;
            if (as_plan == 0) {
                const enum ggml_status st = b200_graph_compute_nodes(bc, cgraph, i, last);
                if (st != GGML_STATUS_SUCCESS) return st;
            }
            i = last;
            continue;
        }
        /* anything else, up to the next decode run: node by node */
        int stop = n == 1 ? last : i + 1;
        const enum ggml_status st = b200_graph_compute_nodes(bc, cgraph, i, stop);
        if (st != GGML_STATUS_SUCCESS) return st;
        i = stop;
    }
    return GGML_STATUS_SUCCESS;
}

/* ggml_backend_graph_plan_create / _free / _compute (src/ggml-backend-impl.h:94-99, src/ggml-backend.c:257-273): the explicit
 * form of the above.  Like the reference CPU backend's plan (src/ggml-backend.c:761-790) it keeps a shallow copy of the cgraph,
 * so the graph has to outlive the plan; the persistent-launch plans of its decode runs are built at the first compute and
 * cached in the backend like any other. */
struct b200_graph_plan {
    struct ggml_cgraph cgraph;
};

GGML_CALL static ggml_backend_graph_plan_t b200_backend_graph_plan_create(ggml_backend_t backend, const struct ggml_cgraph *cgraph) {
    GGML_UNUSED(backend);
    struct b200_graph_plan *gp = (struct b200_graph_plan *)calloc(1, sizeof(*gp));
    if (!gp) return NULL;
    gp->cgraph = *cgraph;
    return (ggml_backend_graph_plan_t)gp;
}

GGML_CALL static void b200_backend_graph_plan_free(ggml_backend_t backend, ggml_backend_graph_plan_t plan) {
    GGML_UNUSED(backend);
    free(plan);
}

GGML_CALL static enum ggml_status b200_backend_graph_plan_compute(ggml_backend_t backend, ggml_backend_graph_plan_t plan) {
    struct b200_graph_plan *gp = (struct b200_graph_plan *)plan;
    return b200_backend_graph_compute(backend, &gp->cgraph);
}

static enum ggml_status b200_graph_compute_nodes(struct b200_backend_context *bc, struct ggml_cgraph *cgraph, int first, int last) {
    for (int i = first; i < last; i++) {
        struct ggml_tensor *node = cgraph->nodes[i];
        if (ggml_is_empty(node) || b200_op_is_noop(node->op)) continue;
        if (node->op == GGML_OP_MUL_MAT) {
            /* gather the run of consecutive MUL_MAT nodes that share this node's src1 and do not depend on one another */
            struct ggml_tensor *run[B200_MAX_RUN];
            int n = 0;
            run[n++] = node;
            while (n < B200_MAX_RUN && i + 1 < last) {
                struct ggml_tensor *next = cgraph->nodes[i + 1];
                if (next->op != GGML_OP_MUL_MAT || next->src[1] != node->src[1] || ggml_is_empty(next)) break;
                bool dep = false;
                for (int j = 0; j < n; j++)
                    if (next->src[0] == run[j] || next->src[0]->view_src == run[j]) dep = true;
                /* (a graph allocator may have given `next` the memory of an earlier result of this run: keep them apart) */
                for (int j = 0; j < n; j++)
                    if (next->data == run[j]->data) dep = true;
                if (dep) break;
                run[n++] = next;
                i++;
            }
            const enum ggml_status st = b200_compute_mul_mat_run(bc, run, n);
            if (st != GGML_STATUS_SUCCESS) return st;
            continue;
        }
        fprintf(stderr, "ggml-b200: op %s is outside this backend's path (supports_op is false for it); no CPU fallback\n",
                ggml_op_name(node->op));
        return GGML_STATUS_FAILED;
    }
    return GGML_STATUS_SUCCESS;
}

static struct ggml_backend_i b200_backend_interface = {
    /* .get_name                = */ b200_backend_name,
    /* .free                    = */ b200_backend_free,
    /* .get_default_buffer_type = */ b200_backend_default_buft,
    /* .set_tensor_async        = */ NULL,
    /* .get_tensor_async        = */ NULL,
    /* .cpy_tensor_async        = */ NULL,
    /* .synchronize             = */ b200_backend_synchronize,
    /* .graph_plan_create       = */ b200_backend_graph_plan_create,
    /* .graph_plan_free         = */ b200_backend_graph_plan_free,
    /* .graph_plan_compute      = */ b200_backend_graph_plan_compute,
    /* .graph_compute           = */ b200_backend_graph_compute,
    /* .supports_op             = */ b200_backend_supports_op,
    /* .offload_op              = */ NULL,
    /* .event_new               = */ NULL,
    /* .event_free              = */ NULL,
    /* .event_record            = */ NULL,
    /* .event_wait              = */ NULL,
    /* .event_synchronize       = */ NULL,
};

GGML_CALL ggml_backend_t ggml_backend_b200_init(int device) {
    if (device < 0 || device >= b200_device_count() || device >= GGML_B200_MAX_DEVICES) {
        fprintf(stderr, "ggml-b200: invalid device %d (%d visible)\n", device, b200_device_count());
        return NULL;
    }
    struct b200_backend_context *bc = (struct b200_backend_context *)calloc(1, sizeof(*bc));
    if (!bc) return NULL;
    bc->device = device;
    bc->opt_plans = 1;
    snprintf(bc->name, sizeof(bc->name), "%s%d", GGML_B200_NAME, device);
    if (b200_ctx_create(device, &bc->ctx) != B200_OK) {
        fprintf(stderr, "ggml-b200: %s\n", b200_last_error(NULL));
        free(bc);
        return NULL;
    }
    ggml_backend_t backend = (ggml_backend_t)malloc(sizeof(struct ggml_backend));
    if (!backend) {
        b200_ctx_destroy(bc->ctx);
        free(bc);
        return NULL;
    }
    backend->guid = b200_backend_guid();
    backend->iface = b200_backend_interface;
    backend->context = bc;
    return backend;
}

GGML_CALL int ggml_backend_b200_get_device_count(void) {
    const int n = b200_device_count();
    return n > GGML_B200_MAX_DEVICES ? GGML_B200_MAX_DEVICES : n;
}

GGML_CALL void ggml_backend_b200_get_device_description(int device, char *description, size_t description_size) {
    if (b200_device_info(device, description, description_size, NULL, NULL, NULL, NULL, NULL) != B200_OK && description_size)
        snprintf(description, description_size, "unknown");
}

GGML_CALL void ggml_backend_b200_get_device_memory(int device, size_t *free_b, size_t *total_b) {
    if (free_b) *free_b = 0;
    if (total_b) *total_b = 0;
    b200_device_info(device, NULL, 0, free_b, total_b, NULL, NULL, NULL);
}

GGML_CALL int64_t ggml_backend_b200_launch_count(ggml_backend_t backend) {
    GGML_ASSERT(ggml_backend_is_b200(backend));
    return b200_ctx_launch_count(((struct b200_backend_context *)backend->context)->ctx);
}

GGML_CALL int64_t ggml_backend_b200_plan_launch_count(ggml_backend_t backend) {
    GGML_ASSERT(ggml_backend_is_b200(backend));
    return ((struct b200_backend_context *)backend->context)->plan_launches;
}

GGML_CALL int ggml_backend_b200_set_option(ggml_backend_t backend, const char *key, int64_t value) {
    GGML_ASSERT(ggml_backend_is_b200(backend));
    if (strcmp(key, "plans") == 0) {
        ((struct b200_backend_context *)backend->context)->opt_plans = value != 0;
        return B200_OK;
    }
    return b200_ctx_set_option(((struct b200_backend_context *)backend->context)->ctx, key, value);
}

GGML_CALL static ggml_backend_t b200_reg_init(const char *params, void *user_data) {
    GGML_UNUSED(params);
    return ggml_backend_b200_init((int)(intptr_t)user_data);
}

GGML_CALL int ggml_backend_b200_reg_devices(void) {
    const int n = ggml_backend_b200_get_device_count();
    for (int i = 0; i < n; i++) {
        char name[32];
        snprintf(name, sizeof(name), "%s%d", GGML_B200_NAME, i);
        ggml_backend_register(name, b200_reg_init, ggml_backend_b200_buffer_type(i), (void *)(intptr_t)i);
    }
    return n;
}

/* ---- drop-in aliases for the src/ggml-cuda.h symbol set ----------------------------------------
 * With the reference core compiled -DGGML_USE_CUDA and src/ggml-cuda.cu left out of the link, these
 * make the registry (src/ggml-backend.c:423-426) and the examples' ggml_backend_cuda_init(0)
 * (examples/gpt-2/main-backend.cpp:200-208) land on this backend unchanged. */
#ifndef GGML_B200_NO_CUDA_ALIASES
GGML_API GGML_CALL ggml_backend_t ggml_backend_cuda_init(int device) { return ggml_backend_b200_init(device); }
GGML_API GGML_CALL bool ggml_backend_is_cuda(ggml_backend_t backend) { return ggml_backend_is_b200(backend); }
GGML_API GGML_CALL ggml_backend_buffer_type_t ggml_backend_cuda_buffer_type(int device) { return ggml_backend_b200_buffer_type(device); }
GGML_API GGML_CALL ggml_backend_buffer_type_t ggml_backend_cuda_host_buffer_type(void) { return ggml_backend_b200_host_buffer_type(); }
/* src/ggml-cuda.h:28-29.  The reference splits a matrix by rows across the devices of ONE process (src/ggml-cuda.cu:578-975).
 * Here rows are split across processes, one per GPU, and the slices are exchanged by the decode plan itself
 * (b200_plan_create with a b200_plan_split, include/ggml_b200.h): there is no single-process split buffer to hand out.
 * The symbol exists so that hosts written against ggml-cuda.h link; NULL = "not available", which callers must handle
 * like any failed buffer-type lookup. */
GGML_API GGML_CALL ggml_backend_buffer_type_t ggml_backend_cuda_split_buffer_type(const float *tensor_split) {
    GGML_UNUSED(tensor_split);
    fprintf(stderr, "ggml-b200: ggml_backend_cuda_split_buffer_type: row split is one process per GPU in this backend "
                    "(b200_plan_create + b200_plan_split); no single-process split buffer type\n");
    return NULL;
}
GGML_API GGML_CALL int ggml_backend_cuda_get_device_count(void) { return ggml_backend_b200_get_device_count(); }
GGML_API GGML_CALL void ggml_backend_cuda_get_device_description(int device, char *description, size_t description_size) {
    ggml_backend_b200_get_device_description(device, description, description_size);
}
GGML_API GGML_CALL void ggml_backend_cuda_get_device_memory(int device, size_t *free_b, size_t *total_b) {
    ggml_backend_b200_get_device_memory(device, free_b, total_b);
}
GGML_API GGML_CALL bool ggml_backend_cuda_register_host_buffer(void *buffer, size_t size) {
    GGML_UNUSED(buffer);
    GGML_UNUSED(size);
    return false; /* optional optimisation in the reference (GGML_CUDA_REGISTER_HOST); not needed on this path */
}
GGML_API GGML_CALL void ggml_backend_cuda_unregister_host_buffer(void *buffer) { GGML_UNUSED(buffer); }
GGML_API GGML_CALL int ggml_backend_cuda_reg_devices(void) { return ggml_backend_b200_reg_devices(); }
#endif
