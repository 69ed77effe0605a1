/*
 * ggml-b200.c -- the B200 ggml backend: host side, plain C, on top of the thin C ABI in
 * include/ggml_b200.h.  Implements the reference's three plugin vtables
 * (src/ggml-backend-impl.h:18-28 buffer type, :38-48 buffer, :78-117 backend) and registers through
 * ggml_backend_register (:135-137).  Written against the SPI, not derived from src/ggml-cuda.cu.
 *
 * Scope: GGML_OP_MUL_MAT with src0 in {Q4_0, Q8_0}, src1 F32, dst F32, plus the operators either side
 * of it in a GPT-2 / GPT-J graph (GET_ROWS, ADD/MUL/DIV, NORM/RMS_NORM, SCALE, DIAG_MASK_INF, SOFT_MAX, UNARY,
 * CPY/DUP/CONT, MUL_MAT with an F32/F16 src0) so that examples/gpt-2/main-backend.cpp computes its whole graph here.  Everything
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
    /* when this device serves row slices of tensors in a split buffer (ggml_backend_b200_split_buffer_type) */
    b200_ctx *split_ctx;                         /* its compute context */
    void *xstage, *ystage;                       /* src1 copy / dense dst slice [n][rows] */
    size_t xstage_size, ystage_size;
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

GGML_CALL bool ggml_backend_is_b200(ggml_backend_t backend);
GGML_CALL int ggml_backend_b200_get_device_count(void);

static bool b200_type_is_repacked(enum ggml_type t) { return t == GGML_TYPE_Q4_0 || t == GGML_TYPE_Q8_0; }
/* the sibling 32-element formats: kept in wire format, served by the plain path of b200_wire_formats.cu */
static bool b200_type_is_wire(enum ggml_type t) { return t == GGML_TYPE_Q5_0 || t == GGML_TYPE_IQ4_NL; }
_Static_assert((int)GGML_TYPE_Q5_0 == B200_TYPE_Q5_0 && (int)GGML_TYPE_IQ4_NL == B200_TYPE_IQ4_NL, "type numbering");
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
    /* device-to-device copies between our buffers, on one device or across two (unified addressing; the copy goes over NVLink when
     * peer access is enabled, else through the host); anything else goes through the core's get+set fallback
     * (src/ggml-backend.c:313-334), which also re-does the repack correctly. */
    struct b200_buffer_context *bc = (struct b200_buffer_context *)buffer->context;
    ggml_backend_buffer_t sbuf = src->view_src ? src->view_src->buffer : src->buffer;
    if (!b200_buffer_is_ours(sbuf)) return false;
    if (!ggml_is_contiguous(src) || !ggml_is_contiguous(dst) || ggml_nbytes(src) != ggml_nbytes(dst)) return false;
    if (b200_type_is_repacked(src->type)) {
        /* plane layout is relative to the whole tensor: raw byte copies are only valid tensor-to-tensor */
        if (src->view_src || dst->view_src || dst->type != src->type) return false;
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


/* ---- split buffer type: rows of a matrix divided across the devices of ONE process ----------------------
 * The reference's interface for multi-GPU weights (ggml_backend_cuda_split_buffer_type, src/ggml-cuda.h:28-29; implementation
 * src/ggml-cuda.cu:578-975): a buffer type whose tensors are not in one device's memory -- every device holds a contiguous range
 * of rows, per-device pointers live in tensor->extra, get_base is a dummy address, views are not supported and set/get_tensor take
 * whole tensors only (:727-790).  Same contract here; each slice is stored repacked on its device.  MUL_MAT with such a src0 runs
 * on every device that owns rows: a decode run as one persistent row-split plan per device (tagged peer stores over NVLink, the
 * all-gather fused into the GEMV epilogue), anything else slice by slice with the dst rows copied into place. */
struct b200_split_buft_context {
    float cum[GGML_B200_MAX_DEVICES + 1];        /* cumulative row fractions: device d owns rows [cum[d], cum[d+1]) * nrows */
    int   ndev;
    char  name[48];
    struct ggml_backend_buffer_type buft;
    bool  used;
};
struct b200_split_extra {
    int      ndev;
    void    *ptr[GGML_B200_MAX_DEVICES];          /* repacked slice on device d (NULL: no rows) */
    int64_t  row0[GGML_B200_MAX_DEVICES + 1];
};
struct b200_split_buffer_context {
    struct b200_split_buft_context *bt;
    struct b200_split_extra **extras;
    int n, cap;
};
#define B200_MAX_SPLIT_TYPES 8
static struct b200_split_buft_context g_split[B200_MAX_SPLIT_TYPES];

GGML_CALL static const char *b200_split_buffer_get_name(ggml_backend_buffer_t buffer) {
    GGML_UNUSED(buffer);
    return GGML_B200_NAME "_Split";
}
static bool b200_buffer_is_split(ggml_backend_buffer_t buffer) { return buffer && buffer->iface.get_name == b200_split_buffer_get_name; }
static bool b200_tensor_is_split(const struct ggml_tensor *t) { return t && b200_buffer_is_split(t->buffer) && t->extra != NULL; }

GGML_CALL static void b200_split_buffer_free(ggml_backend_buffer_t buffer) {
    struct b200_split_buffer_context *sc = (struct b200_split_buffer_context *)buffer->context;
    for (int i = 0; i < sc->n; i++) {
        struct b200_split_extra *e = sc->extras[i];
        for (int d = 0; d < e->ndev; d++)
            if (e->ptr[d]) {
                b200_ctx *io = b200_io_ctx(d);
                if (io) b200_free(io, e->ptr[d]);
            }
        free(e);
    }
    free(sc->extras);
    free(sc);
}
GGML_CALL static void *b200_split_buffer_get_base(ggml_backend_buffer_t buffer) {
    GGML_UNUSED(buffer);
    return (void *)0x1000;          /* tensors are addressed through tensor->extra (src/ggml-cuda.cu:720-725 does the same) */
}
GGML_CALL static void b200_split_buffer_init_tensor(ggml_backend_buffer_t buffer, struct ggml_tensor *tensor) {
    struct b200_split_buffer_context *sc = (struct b200_split_buffer_context *)buffer->context;
    GGML_ASSERT(tensor->view_src == NULL && "ggml-b200: views of split tensors are not supported");
    GGML_ASSERT(b200_type_is_repacked(tensor->type) && tensor->ne[2] == 1 && tensor->ne[3] == 1 &&
                "ggml-b200: the split buffer type holds 2-D Q4_0 / Q8_0 matrices");
    struct b200_split_extra *e = (struct b200_split_extra *)calloc(1, sizeof(*e));
    GGML_ASSERT(e != NULL);
    const int64_t nrows = tensor->ne[1], nb = tensor->ne[0] / B200_QK, wire = b200_wire_bytes(tensor->type);
    e->ndev = sc->bt->ndev;
    for (int d = 0; d <= e->ndev; d++) e->row0[d] = d == e->ndev ? nrows : (int64_t)((double)sc->bt->cum[d] * (double)nrows);
    for (int d = 0; d < e->ndev; d++) {
        const int64_t rows = e->row0[d + 1] - e->row0[d];
        if (rows <= 0) continue;
        b200_ctx *io = b200_io_ctx(d);
        GGML_ASSERT(io != NULL);
        B200_CHECK(io, b200_malloc(io, &e->ptr[d], (size_t)(rows * nb * wire)));
    }
    if (sc->n == sc->cap) {
        sc->cap = sc->cap ? sc->cap * 2 : 64;
        sc->extras = (struct b200_split_extra **)realloc(sc->extras, sizeof(*sc->extras) * (size_t)sc->cap);
        GGML_ASSERT(sc->extras != NULL);
    }
    sc->extras[sc->n++] = e;
    tensor->extra = e;
}
GGML_CALL static void b200_split_buffer_set_tensor(ggml_backend_buffer_t buffer, struct ggml_tensor *tensor, const void *data, size_t offset, size_t size) {
    GGML_UNUSED(buffer);
    GGML_ASSERT(offset == 0 && size == ggml_nbytes(tensor) && "ggml-b200: split tensors are set whole (src/ggml-cuda.cu:776-778)");
    const struct b200_split_extra *e = (const struct b200_split_extra *)tensor->extra;
    const int64_t nb = tensor->ne[0] / B200_QK, wire = b200_wire_bytes(tensor->type);
    for (int d = 0; d < e->ndev; d++) {
        const int64_t rows = e->row0[d + 1] - e->row0[d];
        if (rows <= 0) continue;
        b200_ctx *io = b200_io_ctx(d);
        B200_CHECK(io, b200_set_quantized(io, (int)tensor->type, e->ptr[d], rows * nb, (const char *)data + e->row0[d] * nb * wire, 0, rows * nb));
    }
}
GGML_CALL static void b200_split_buffer_get_tensor(ggml_backend_buffer_t buffer, const struct ggml_tensor *tensor, void *data, size_t offset, size_t size) {
    GGML_UNUSED(buffer);
    GGML_ASSERT(offset == 0 && size == ggml_nbytes(tensor) && "ggml-b200: split tensors are read whole");
    const struct b200_split_extra *e = (const struct b200_split_extra *)tensor->extra;
    const int64_t nb = tensor->ne[0] / B200_QK, wire = b200_wire_bytes(tensor->type);
    for (int d = 0; d < e->ndev; d++) {
        const int64_t rows = e->row0[d + 1] - e->row0[d];
        if (rows <= 0) continue;
        b200_ctx *io = b200_io_ctx(d);
        B200_CHECK(io, b200_get_quantized(io, (int)tensor->type, e->ptr[d], rows * nb, (char *)data + e->row0[d] * nb * wire, 0, rows * nb));
    }
}
GGML_CALL static void b200_split_buffer_clear(ggml_backend_buffer_t buffer, uint8_t value) {
    GGML_UNUSED(buffer);
    GGML_UNUSED(value);
}
static struct ggml_backend_buffer_i b200_split_buffer_interface = {
    /* .get_name    = */ b200_split_buffer_get_name,
    /* .free_buffer = */ b200_split_buffer_free,
    /* .get_base    = */ b200_split_buffer_get_base,
    /* .init_tensor = */ b200_split_buffer_init_tensor,
    /* .set_tensor  = */ b200_split_buffer_set_tensor,
    /* .get_tensor  = */ b200_split_buffer_get_tensor,
    /* .cpy_tensor  = */ NULL,
    /* .clear       = */ b200_split_buffer_clear,
    /* .reset       = */ NULL,
};
GGML_CALL static const char *b200_split_buft_get_name(ggml_backend_buffer_type_t buft) {
    return ((struct b200_split_buft_context *)buft->context)->name;
}
GGML_CALL static ggml_backend_buffer_t b200_split_buft_alloc_buffer(ggml_backend_buffer_type_t buft, size_t size) {
    struct b200_split_buffer_context *sc = (struct b200_split_buffer_context *)calloc(1, sizeof(*sc));
    if (!sc) return NULL;
    sc->bt = (struct b200_split_buft_context *)buft->context;
    return ggml_backend_buffer_init(buft, b200_split_buffer_interface, sc, size);     /* memory is allocated per tensor, in init_tensor */
}
GGML_CALL static size_t b200_split_buft_get_alignment(ggml_backend_buffer_type_t buft) {
    GGML_UNUSED(buft);
    return 128;
}
GGML_CALL static size_t b200_split_buft_get_alloc_size(ggml_backend_buffer_type_t buft, const struct ggml_tensor *tensor) {
    GGML_UNUSED(buft);
    return ggml_nbytes(tensor);
}
GGML_CALL static bool b200_split_buft_supports_backend(ggml_backend_buffer_type_t buft, ggml_backend_t backend) {
    GGML_UNUSED(buft);
    return ggml_backend_is_b200(backend);
}
GGML_CALL static bool b200_split_buft_is_host(ggml_backend_buffer_type_t buft) {
    GGML_UNUSED(buft);
    return false;
}

/* tensor_split: GGML_B200_MAX_DEVICES proportions (0 = device unused), or NULL for equal shares of every visible device.  One
 * buffer type object per distinct split (at most B200_MAX_SPLIT_TYPES); NULL when no device is visible. */
GGML_CALL ggml_backend_buffer_type_t ggml_backend_b200_split_buffer_type(const float *tensor_split) {
    const int ndev = ggml_backend_b200_get_device_count();
    if (ndev <= 0) return NULL;
    float cum[GGML_B200_MAX_DEVICES + 1];
    double total = 0;
    for (int d = 0; d < ndev; d++) total += tensor_split ? (tensor_split[d] > 0 ? tensor_split[d] : 0) : 1.0;
    if (total <= 0) return NULL;
    double run = 0;
    for (int d = 0; d < ndev; d++) {
        cum[d] = (float)(run / total);
        run += tensor_split ? (tensor_split[d] > 0 ? tensor_split[d] : 0) : 1.0;
    }
    cum[ndev] = 1.0f;
    for (int i = 0; i < B200_MAX_SPLIT_TYPES; i++) {
        struct b200_split_buft_context *bt = &g_split[i];
        if (bt->used && bt->ndev == ndev && memcmp(bt->cum, cum, sizeof(float) * (size_t)(ndev + 1)) == 0) return &bt->buft;
        if (!bt->used) {
            bt->used = true;
            bt->ndev = ndev;
            memcpy(bt->cum, cum, sizeof(float) * (size_t)(ndev + 1));
            snprintf(bt->name, sizeof(bt->name), "%s_Split%d", GGML_B200_NAME, i);
            bt->buft.iface.get_name = b200_split_buft_get_name;
            bt->buft.iface.alloc_buffer = b200_split_buft_alloc_buffer;
            bt->buft.iface.get_alignment = b200_split_buft_get_alignment;
            bt->buft.iface.get_max_size = NULL;
            bt->buft.iface.get_alloc_size = b200_split_buft_get_alloc_size;
            bt->buft.iface.supports_backend = b200_split_buft_supports_backend;
            bt->buft.iface.is_host = b200_split_buft_is_host;
            bt->buft.context = bt;
            return &bt->buft;
        }
    }
    fprintf(stderr, "ggml-b200: more than %d distinct tensor splits\n", B200_MAX_SPLIT_TYPES);
    return NULL;
}

/* ---- backend --------------------------------------------------------------------------------- */

/* a decode plan cached for a cgraph: valid while the graph's MUL_MAT nodes keep their addresses and shapes */
#define B200_PLAN_CACHE 4
#define B200_MAX_XIN 8
struct b200_cached_plan {
    uint64_t   key;        /* hash over (src0, src1, dst addresses, shapes) of the run's MUL_MAT nodes; 0 = empty slot */
    int        n_nodes;
    b200_plan *plan;       /* NULL: this run cannot go down as a plan (remembered, so it is not analysed again) */
    /* a run whose weights live in a split buffer: one row-split plan per device, launched together */
    int        ndev;                                  /* 0: single-device plan */
    b200_plan *dplan[GGML_B200_MAX_DEVICES];
    void      *arena[GGML_B200_MAX_DEVICES];          /* exchange arenas (tagged vectors), peer-accessible */
    void      *scratch[GGML_B200_MAX_DEVICES];        /* devices other than the backend's: their own (partial) dst vectors */
    void      *xin[GGML_B200_MAX_DEVICES];            /* ... and copies of the vectors that come from outside the run */
    int        n_xin;
    const void *xin_src[B200_MAX_XIN];
    size_t     xin_bytes[B200_MAX_XIN], xin_off[B200_MAX_XIN];
};

#define B200_MAX_DEFERRED 64
struct b200_backend_context {
    int       device;
    b200_ctx *ctx;
    char      name[32];
    struct b200_cached_plan plans[B200_PLAN_CACHE];
    int       plan_next;   /* round-robin eviction */
    int       opt_plans;   /* 0: never build decode plans (ggml_backend_b200_set_option "plans") */
    int       opt_fuse;    /* 0: one kernel per glue node (ggml_backend_b200_set_option "fuse") */
    int       opt_graphs;  /* 0: graph plans never record a CUDA graph (ggml_backend_b200_set_option "graphs") */
    int64_t   plan_launches;
    int64_t   fused_nodes; /* graph nodes that did not need a launch of their own */
    /* consumers of every tensor of the graph being computed (b200_uses_*), built on demand for the fusions that skip whole nodes */
    const struct ggml_tensor **use_keys;
    int32_t  *use_cnt;
    size_t    use_size;
    int       use_valid;
    int       graph_managed;   /* this graph_compute has seen the allocator place a node in a dead intermediate's memory: unflagged intermediates are not the caller's to read */
    /* REPEATs of a row whose only reader comes later in the graph: computed right before that reader unless it folds them in */
    struct ggml_tensor *deferred[B200_MAX_DEFERRED];
    int       n_deferred;
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

/* compute context of device d for this backend: its own on the backend's device, a per-device one elsewhere (split tensors) */
static b200_ctx *b200_device_compute_ctx(struct b200_backend_context *bc, int d) {
    if (d == bc->device) return bc->ctx;
    if (d < 0 || d >= GGML_B200_MAX_DEVICES) return NULL;
    if (g_dev[d].split_ctx == NULL && b200_ctx_create(d, &g_dev[d].split_ctx) != B200_OK) {
        fprintf(stderr, "ggml-b200: cannot create a context on device %d: %s\n", d, b200_last_error(NULL));
        return NULL;
    }
    return g_dev[d].split_ctx;
}

static void b200_cached_plan_release(struct b200_backend_context *bc, struct b200_cached_plan *slot) {
    if (slot->plan) {
        b200_synchronize(bc->ctx);
        b200_plan_destroy(slot->plan);
    }
    for (int d = 0; d < slot->ndev; d++) {
        b200_ctx *cd = b200_device_compute_ctx(bc, d);
        if (!cd) continue;
        b200_synchronize(cd);
        if (slot->dplan[d]) b200_plan_destroy(slot->dplan[d]);
        if (slot->arena[d]) b200_free(cd, slot->arena[d]);
        if (slot->scratch[d]) b200_free(cd, slot->scratch[d]);
        if (slot->xin[d]) b200_free(cd, slot->xin[d]);
    }
    memset(slot, 0, sizeof(*slot));
}

GGML_CALL static void b200_backend_free(ggml_backend_t backend) {
    struct b200_backend_context *bc = (struct b200_backend_context *)backend->context;
    b200_synchronize(bc->ctx);
    for (int i = 0; i < B200_PLAN_CACHE; i++) b200_cached_plan_release(bc, &bc->plans[i]);
    b200_ctx_destroy(bc->ctx);
    free(bc->use_keys);
    free(bc->use_cnt);
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
static bool b200_wire_mul_mat_supported(const struct ggml_tensor *dst) {
    const struct ggml_tensor *a = dst->src[0], *b = dst->src[1];
    if (!a || !b || !b200_type_is_wire(a->type) || b->type != GGML_TYPE_F32 || dst->type != GGML_TYPE_F32) return false;
    if (!ggml_is_contiguous(a) || !ggml_is_contiguous(dst) || b->nb[0] != sizeof(float)) return false;
    if (b200_buffer_is_split(a->buffer) || b->ne[2] % a->ne[2] != 0 || b->ne[3] % a->ne[3] != 0) return false;
    return b->ne[1] <= 65535 && b->ne[2] * b->ne[3] <= 65535 && a->ne[0] <= 131072;
}

static bool b200_mul_mat_supported(const struct ggml_tensor *dst) {
    const struct ggml_tensor *a = dst->src[0], *b = dst->src[1];
    if (!a || !b) return false;
    if (!b200_type_is_repacked(a->type) || b->type != GGML_TYPE_F32 || dst->type != GGML_TYPE_F32) return false;
    if (b200_buffer_is_split(a->buffer)) {
        /* rows of src0 live on several devices: 2-D weights, one dense 2-D activation matrix */
        return a->view_src == NULL && a->ne[2] == 1 && a->ne[3] == 1 && b->ne[2] == 1 && b->ne[3] == 1 && b->nb[0] == sizeof(float) &&
               b->nb[1] % 16 == 0 && ggml_is_contiguous(dst) && a->ne[0] <= 131072 && (b->view_src == NULL || b->view_offs % 16 == 0);
    }
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


/* ---- the operators either side of the path (b200_ops.cu) --------------------------------------------------------------------- */

_Static_assert((int)GGML_UNARY_OP_ABS == B200_UNARY_ABS && (int)GGML_UNARY_OP_GELU == B200_UNARY_GELU && (int)GGML_UNARY_OP_SILU == B200_UNARY_SILU &&
               (int)GGML_UNARY_OP_HARDSIGMOID == B200_UNARY_HARDSIGMOID && (int)GGML_UNARY_OP_COUNT == B200_UNARY_COUNT, "unary op numbering");
_Static_assert((int)GGML_TYPE_F16 == B200_TYPE_F16 && (int)GGML_TYPE_I16 == B200_TYPE_I16 && (int)GGML_TYPE_I32 == B200_TYPE_I32, "type numbering");

static bool b200_in_device_buffer(const struct ggml_tensor *t) {
    /* unallocated (supports_op before allocation) or in one of our plain device buffers */
    ggml_backend_buffer_t buf = t->view_src ? t->view_src->buffer : t->buffer;
    return buf == NULL || b200_buffer_is_ours(buf);
}
static bool b200_is_f32(const struct ggml_tensor *t) { return t && t->type == GGML_TYPE_F32; }
/* ggml_can_repeat (static in src/ggml.c): every extent of `big` is a multiple of the matching extent of `small` */
static bool b200_can_repeat(const struct ggml_tensor *small, const struct ggml_tensor *big) {
    for (int i = 0; i < 4; i++)
        if (small->ne[i] <= 0 || big->ne[i] % small->ne[i] != 0) return false;
    return true;
}
static bool b200_views_16(const struct ggml_tensor *t) { return t->view_src == NULL || t->view_offs % 16 == 0; }

/* what the kernels see of a tensor; quantized tensors are addressed through their repacked root */
static bool b200_fill_tensor(const struct ggml_tensor *t, b200_tensor *out) {
    memset(out, 0, sizeof(*out));
    out->type = (int32_t)t->type;
    for (int i = 0; i < 4; i++) {
        out->ne[i] = t->ne[i];
        out->nb[i] = (int64_t)t->nb[i];
    }
    if (b200_type_is_repacked(t->type)) {
        const struct ggml_tensor *root = t->view_src ? t->view_src : t;
        const int64_t wire = b200_wire_bytes(t->type);
        const size_t offs = t->view_src ? t->view_offs : 0;
        if (root->type != t->type || !ggml_is_contiguous(root) || offs % wire != 0) return false;
        out->data = root->data;
        out->q_total_blocks = ggml_nelements(root) / B200_QK;
        out->q_block_off = (int64_t)(offs / wire);
        return true;
    }
    out->data = t->data;
    return true;
}

static bool b200_copy_supported(const struct ggml_tensor *src, const struct ggml_tensor *dst) {
    if (!src || ggml_nelements(src) != ggml_nelements(dst)) return false;
    const bool fs = src->type == GGML_TYPE_F32 || src->type == GGML_TYPE_F16, fd = dst->type == GGML_TYPE_F32 || dst->type == GGML_TYPE_F16;
    if (fs && fd) return true;
    return src->type == dst->type && (src->type == GGML_TYPE_I32 || src->type == GGML_TYPE_I16);
}

static bool b200_dense_mul_mat_supported(const struct ggml_tensor *dst) {
    const struct ggml_tensor *a = dst->src[0], *b = dst->src[1];
    if (!a || !b || (a->type != GGML_TYPE_F32 && a->type != GGML_TYPE_F16) || b->type != GGML_TYPE_F32 || dst->type != GGML_TYPE_F32) return false;
    if (a->nb[0] != ggml_type_size(a->type) || b->nb[0] != sizeof(float) || dst->nb[0] != sizeof(float)) return false;
    if (b->ne[2] % a->ne[2] != 0 || b->ne[3] % a->ne[3] != 0) return false;
    return dst->ne[2] * dst->ne[3] <= 65535;
}

static bool b200_glue_supported_srcs(const struct ggml_tensor *op) {
    for (int i = 0; i < GGML_MAX_SRC; i++)
        if (op->src[i] && (!b200_in_device_buffer(op->src[i]) || b200_tensor_is_split(op->src[i]))) return false;
    return true;
}

static bool b200_glue_supported(const struct ggml_tensor *op) {
    const struct ggml_tensor *a = op->src[0], *b = op->src[1];
    if (!b200_glue_supported_srcs(op)) return false;
    switch (op->op) {
    case GGML_OP_GET_ROWS:
        if (!a || !b || b->type != GGML_TYPE_I32 || op->type != GGML_TYPE_F32 || op->nb[0] != sizeof(float)) return false;
        if (b200_type_is_repacked(a->type)) {
            const struct ggml_tensor *root = a->view_src ? a->view_src : a;
            const size_t wire = (size_t)b200_wire_bytes(a->type);
            return root->type == a->type && ggml_is_contiguous(root) && (a->view_src == NULL || a->view_offs % wire == 0) && a->nb[0] == wire &&
                   a->nb[1] % wire == 0 && a->nb[2] % wire == 0 && a->nb[3] % wire == 0 && b200_views_16(op) && op->nb[1] % 16 == 0 &&
                   op->nb[2] % 16 == 0 && op->nb[3] % 16 == 0;
        }
        if (b200_type_is_wire(a->type)) return a->nb[0] == ggml_type_size(a->type) && a->ne[0] % B200_QK == 0;
        return (a->type == GGML_TYPE_F32 || a->type == GGML_TYPE_F16) && a->nb[0] == ggml_type_size(a->type);
    case GGML_OP_ADD:
    case GGML_OP_MUL:
    case GGML_OP_DIV:
        return b200_is_f32(a) && b200_is_f32(b) && op->type == GGML_TYPE_F32 && ggml_are_same_shape(a, op) && b200_can_repeat(b, a);
    case GGML_OP_UNARY:
        return b200_is_f32(a) && op->type == GGML_TYPE_F32 && ggml_is_contiguous(a) && ggml_is_contiguous(op) && (int)ggml_get_unary_op(op) < B200_UNARY_COUNT;
    case GGML_OP_NORM:
    case GGML_OP_RMS_NORM:
        return b200_is_f32(a) && op->type == GGML_TYPE_F32 && a->nb[0] == sizeof(float) && op->nb[0] == sizeof(float) && ggml_nrows(a) < (1ll << 31);
    case GGML_OP_SCALE:
    case GGML_OP_DIAG_MASK_INF:
        return b200_is_f32(a) && op->type == GGML_TYPE_F32 && ggml_is_contiguous(a) && ggml_is_contiguous(op);
    case GGML_OP_SOFT_MAX:
        if (!b200_is_f32(a) || op->type != GGML_TYPE_F32 || !ggml_is_contiguous(a) || !ggml_is_contiguous(op) || ggml_nrows(a) >= (1ll << 31)) return false;
        if (op->src[2] != NULL) return false;
        return b == NULL || ((b->type == GGML_TYPE_F32 || b->type == GGML_TYPE_F16) && ggml_is_contiguous(b) && b->ne[0] == a->ne[0] && b->ne[1] >= a->ne[1]);
    case GGML_OP_CPY:
    case GGML_OP_DUP:
    case GGML_OP_CONT:
        return b200_copy_supported(a, op);
    case GGML_OP_ROPE:          /* forward, normal and NeoX pairing (GPT-J: examples/gpt-j/main.cpp:473-474); the GLM layout is declined */
        return a && b && (a->type == GGML_TYPE_F32 || a->type == GGML_TYPE_F16) && op->type == a->type && b->type == GGML_TYPE_I32 &&
               (((const int32_t *)op->op_params)[2] & 4) == 0 && a->nb[0] == ggml_type_size(a->type) && op->nb[0] == a->nb[0] && a->ne[0] % 2 == 0 &&
               ggml_is_contiguous(b) && b->ne[0] >= a->ne[2];
    case GGML_OP_REPEAT:
        return a && a->type == op->type && (a->type == GGML_TYPE_F32 || a->type == GGML_TYPE_F16 || a->type == GGML_TYPE_I32 || a->type == GGML_TYPE_I16) &&
               b200_can_repeat(a, op);
    default:
        return false;
    }
}

static enum ggml_status b200_glue_status(struct b200_backend_context *bc, const struct ggml_tensor *node, int rc);
static bool b200_fill_mul_mat_args(struct ggml_tensor *dst, b200_mul_mat_args *args);

/* one glue node (no fusion) */
static enum ggml_status b200_compute_glue(struct b200_backend_context *bc, struct ggml_tensor *node) {
    b200_tensor a, b, d;
    if (node->op == GGML_OP_MUL_MAT && b200_wire_mul_mat_supported(node)) {
        const struct ggml_tensor *wa = node->src[0], *wb = node->src[1];
        b200_mul_mat_args args;
        memset(&args, 0, sizeof(args));
        args.type = (int32_t)wa->type;
        args.src0_dev = wa->data;
        args.src0_nblocks_total = ggml_nelements(wa) / B200_QK;
        args.ne00 = wa->ne[0]; args.ne01 = wa->ne[1]; args.ne02 = wa->ne[2]; args.ne03 = wa->ne[3];
        args.src1_dev = (const float *)wb->data;
        args.ne11 = wb->ne[1]; args.ne12 = wb->ne[2]; args.ne13 = wb->ne[3];
        args.nb11 = wb->nb[1]; args.nb12 = wb->nb[2]; args.nb13 = wb->nb[3];
        args.dst_dev = (float *)node->data;
        return b200_glue_status(bc, node, b200_mul_mat(bc->ctx, &args));
    }
    if (!b200_glue_supported(node) && !(node->op == GGML_OP_MUL_MAT && b200_dense_mul_mat_supported(node))) {
        fprintf(stderr, "ggml-b200: op %s (%s) is outside this backend's path (supports_op is false for it); no CPU fallback\n", ggml_op_name(node->op),
                ggml_type_name(node->type));
        return GGML_STATUS_FAILED;
    }
    if (!b200_fill_tensor(node, &d) || (node->src[0] && !b200_fill_tensor(node->src[0], &a)) || (node->src[1] && !b200_fill_tensor(node->src[1], &b)))
        return GGML_STATUS_FAILED;
    int rc;
    switch (node->op) {
    case GGML_OP_GET_ROWS: rc = b200_op_get_rows(bc->ctx, &a, &b, &d); break;
    case GGML_OP_ADD: rc = b200_op_binary(bc->ctx, B200_OP_ADD, &a, &b, &d); break;
    case GGML_OP_MUL: rc = b200_op_binary(bc->ctx, B200_OP_MUL, &a, &b, &d); break;
    case GGML_OP_DIV: rc = b200_op_binary(bc->ctx, B200_OP_DIV, &a, &b, &d); break;
    case GGML_OP_UNARY: rc = b200_op_unary(bc->ctx, (int)ggml_get_unary_op(node), &a, &d); break;
    case GGML_OP_NORM:
    case GGML_OP_RMS_NORM: {
        float eps;
        memcpy(&eps, node->op_params, sizeof(eps));
        rc = b200_op_norm(bc->ctx, &a, NULL, NULL, &d, eps, node->op == GGML_OP_RMS_NORM);
        break;
    }
    case GGML_OP_SCALE: {
        float v;
        memcpy(&v, node->op_params, sizeof(v));
        rc = b200_op_scale(bc->ctx, &a, &d, v);
        break;
    }
    case GGML_OP_DIAG_MASK_INF: rc = b200_op_diag_mask_inf(bc->ctx, &a, &d, ((const int32_t *)node->op_params)[0]); break;
    case GGML_OP_SOFT_MAX: {
        float scale, max_bias;
        memcpy(&scale, (const float *)node->op_params + 0, sizeof(scale));
        memcpy(&max_bias, (const float *)node->op_params + 1, sizeof(max_bias));
        rc = b200_op_soft_max(bc->ctx, &a, node->src[1] ? &b : NULL, &d, scale, max_bias, -1);
        break;
    }
    case GGML_OP_CPY:
    case GGML_OP_DUP:
    case GGML_OP_CONT: rc = b200_op_copy(bc->ctx, &a, &d); break;
    case GGML_OP_ROPE: {        /* op_params as ggml_rope_impl lays them out (src/ggml.c:5866-5889) */
        const int32_t *op = (const int32_t *)node->op_params;
        b200_rope_params rp;
        memset(&rp, 0, sizeof(rp));
        rp.n_dims = op[1]; rp.mode = op[2]; rp.n_ctx = op[3]; rp.n_orig_ctx = op[4];
        memcpy(&rp.freq_base, op + 5, sizeof(float));
        memcpy(&rp.freq_scale, op + 6, sizeof(float));
        memcpy(&rp.ext_factor, op + 7, sizeof(float));
        memcpy(&rp.attn_factor, op + 8, sizeof(float));
        memcpy(&rp.beta_fast, op + 9, sizeof(float));
        memcpy(&rp.beta_slow, op + 10, sizeof(float));
        memcpy(&rp.xpos_base, op + 11, sizeof(float));
        { bool down; memcpy(&down, op + 12, sizeof(bool)); rp.xpos_down = down ? 1 : 0; }
        rc = b200_op_rope(bc->ctx, &a, &b, &d, &rp);
        break;
    }
    case GGML_OP_REPEAT: rc = b200_op_repeat(bc->ctx, &a, &d); break;
    case GGML_OP_MUL_MAT: rc = b200_op_mul_mat_dense(bc->ctx, &a, &b, &d); break;
    default: rc = B200_ERR_UNSUPPORTED; break;
    }
    return b200_glue_status(bc, node, rc);
}

/*
 * Fusions, only where the graph allocator has ALREADY proven that the intermediates are dead: ggml_gallocr
 * computes an op in place of its parent exactly when nothing else reads the parent (src/ggml-alloc.c: n_children == 1, n_views == 0),
 * so "the next node reads this one and writes the same memory" means the fused kernel loses no visible state.
 *   NORM -> MUL(gain row) -> ADD(bias row)          one kernel (examples/gpt-2/main-backend.cpp:490-498, :634-642, :686-694)
 *   SCALE -> DIAG_MASK_INF -> SOFT_MAX              one kernel (:567-583)
 * Returns the number of nodes consumed (0 = no fusion here).
 */
static bool b200_row_vector(const struct ggml_tensor *v, int64_t ne0) {
    return v && v->type == GGML_TYPE_F32 && v->ne[0] == ne0 && ggml_nrows(v) == 1 && v->nb[0] == sizeof(float) && b200_in_device_buffer(v);
}
static bool b200_inplace_child(const struct ggml_tensor *child, const struct ggml_tensor *parent) {
    if (child->src[0] != parent || child->data != parent->data || !ggml_are_same_shape(child, parent) || (parent->flags & GGML_TENSOR_FLAG_OUTPUT)) return false;
    for (int i = 0; i < GGML_MAX_DIMS; i++)
        if (child->nb[i] != parent->nb[i]) return false;
    if (child->view_src == NULL) return true;               /* the allocator's reuse */
    /* the explicit form (ggml_scale_inplace, ggml_diag_mask_inf_inplace, ggml_soft_max_inplace: examples/gpt-j/main.cpp:500-509): the child is
     * ggml_view_tensor(parent) and overwrites it, so whatever reads that memory later sees the child's values with or without the fusion */
    return child->view_src == (parent->view_src ? parent->view_src : parent) && child->view_offs == parent->view_offs;
}
/*
 * Fusions that SKIP nodes whose result nobody else reads (the old-style graphs of examples/gpt-j/main.cpp, where every gain / bias row is
 * first broadcast by GGML_OP_REPEAT and nothing is computed in place of its first operand):
 *   NORM -> REPEAT(gain) -> MUL -> REPEAT(bias) -> ADD                                  one kernel (main.cpp:446-456, :559-567)
 *   MUL_MAT(decode) -> REPEAT(bias) -> ADD [-> GELU] [-> ADD(residual)]                 the GEMV's epilogue (main.cpp:530-553)
 * The proof that the skipped results are dead is a count of their readers over the WHOLE graph, so these are only tried when
 *   - the cgraph is a complete one (ggml_backend_sched hands over ggml_graph_view()s, whose readers may sit in another split: never fused),
 *   - the group's last node was placed by the graph allocator in the memory of one of the group's own intermediates (a graph whose
 *     tensors all own their memory may have any of them read back afterwards: never fused),
 *   - no intermediate carries GGML_TENSOR_FLAG_OUTPUT and every reader of every intermediate (views included) is inside the group,
 *   - the destination does not overlap an input the kernel still reads (the allocator may have recycled the activation's memory).
 */
static size_t b200_use_slot(const struct b200_backend_context *bc, const struct ggml_tensor *t) {
    size_t h = ((uintptr_t)t >> 4) * 0x9E3779B97F4A7C15ull >> 20;
    h &= bc->use_size - 1;
    while (bc->use_keys[h] && bc->use_keys[h] != t) h = (h + 1) & (bc->use_size - 1);
    return h;
}
static void b200_use_add(struct b200_backend_context *bc, const struct ggml_tensor *t) {
    const size_t h = b200_use_slot(bc, t);
    bc->use_keys[h] = t;
    bc->use_cnt[h]++;
}
/* readers of one node: its distinct sources, and the tensor it is a view of */
static int b200_refs_from(const struct ggml_tensor *node, const struct ggml_tensor *t) {
    int n = 0;
    for (int k = 0; k < GGML_MAX_SRC; k++) {
        if (node->src[k] != t) continue;
        bool seen = false;
        for (int j = 0; j < k; j++) seen |= node->src[j] == t;
        n += !seen;
    }
    return n + (node->view_src == t);
}
static bool b200_uses_build(struct b200_backend_context *bc, const struct ggml_cgraph *g) {
    if (bc->use_valid) return true;
    if (g->visited_hash_table.size == 0 || g->visited_hash_table.keys == NULL) return false;      /* a ggml_graph_view: readers may be elsewhere */
    size_t refs = 0;                                      /* an upper bound of the distinct tensors that get an entry */
    for (int i = 0; i < g->n_nodes; i++) {
        for (int k = 0; k < GGML_MAX_SRC; k++) refs += g->nodes[i]->src[k] != NULL;
        refs += g->nodes[i]->view_src != NULL;
    }
    size_t want = 1024;
    while (want < refs * 2 + 16) want <<= 1;              /* open addressing: never more than half full */
    if (want > bc->use_size) {
        free(bc->use_keys);
        free(bc->use_cnt);
        bc->use_keys = (const struct ggml_tensor **)malloc(want * sizeof(*bc->use_keys));
        bc->use_cnt = (int32_t *)malloc(want * sizeof(*bc->use_cnt));
        bc->use_size = bc->use_keys && bc->use_cnt ? want : 0;
        if (!bc->use_size) return false;
    }
    memset(bc->use_keys, 0, bc->use_size * sizeof(*bc->use_keys));
    memset(bc->use_cnt, 0, bc->use_size * sizeof(*bc->use_cnt));
    for (int i = 0; i < g->n_nodes; i++) {
        const struct ggml_tensor *node = g->nodes[i];
        for (int k = 0; k < GGML_MAX_SRC; k++) {
            if (!node->src[k]) continue;
            bool seen = false;
            for (int j = 0; j < k; j++) seen |= node->src[j] == node->src[k];
            if (!seen) b200_use_add(bc, node->src[k]);
        }
        if (node->view_src) b200_use_add(bc, node->view_src);
    }
    bc->use_valid = 1;
    return true;
}
/* every reader of t is one of nodes[i .. i + n) */
static bool b200_read_only_by_group(struct b200_backend_context *bc, const struct ggml_cgraph *g, const struct ggml_tensor *t, int i, int n) {
    if (t->flags & GGML_TENSOR_FLAG_OUTPUT) return false;
    int inside = 0;
    for (int j = i; j < i + n; j++) inside += b200_refs_from(g->nodes[j], t);
    return bc->use_cnt[b200_use_slot(bc, t)] == inside;
}
static bool b200_ranges_overlap(const struct ggml_tensor *a, const struct ggml_tensor *b) {
    const char *a0 = (const char *)a->data, *b0 = (const char *)b->data;
    return a0 < b0 + ggml_nbytes(b) && b0 < a0 + ggml_nbytes(a);
}
/* REPEAT of a device-resident F32 row of ne0 elements to the shape of `like` */
static bool b200_repeat_of_row(const struct ggml_tensor *rep, const struct ggml_tensor *like) {
    return rep->op == GGML_OP_REPEAT && rep->type == GGML_TYPE_F32 && ggml_are_same_shape(rep, like) && b200_row_vector(rep->src[0], like->ne[0]) &&
           rep->src[0]->data != NULL;
}
static bool b200_binary_of(const struct ggml_tensor *node, enum ggml_op op, const struct ggml_tensor *x, const struct ggml_tensor *y) {
    return node->op == op && node->type == GGML_TYPE_F32 && ((node->src[0] == x && node->src[1] == y) || (node->src[0] == y && node->src[1] == x));
}

/* A REPEAT of a row does not reference the tensor whose shape it takes (ggml_repeat, src/ggml.c:4390), so ggml_build_forward_expand may
 * emit it long before its reader (GPT-J: the bias REPEATs of a block's MLP come out before the previous block's nodes).  Such a node is
 * DEFERRED when it is met: its memory is reserved from there to its reader, so writing it at any point in between is the same program.
 * The reader either folds the row into its own kernel (and the REPEAT is never computed) or has it computed right before it runs. */
static int b200_deferred_index(const struct b200_backend_context *bc, const struct ggml_tensor *t) {
    for (int k = 0; k < bc->n_deferred; k++)
        if (bc->deferred[k] == t) return k;
    return -1;
}
static void b200_deferred_drop(struct b200_backend_context *bc, const struct ggml_tensor *t) {
    const int k = b200_deferred_index(bc, t);
    if (k >= 0) bc->deferred[k] = bc->deferred[--bc->n_deferred];
}
static bool b200_try_defer_repeat(struct b200_backend_context *bc, const struct ggml_cgraph *cgraph, struct ggml_tensor *rep) {
    if (!bc->opt_fuse || rep->op != GGML_OP_REPEAT || rep->type != GGML_TYPE_F32 || bc->n_deferred >= B200_MAX_DEFERRED) return false;
    if (!b200_row_vector(rep->src[0], rep->ne[0]) || !ggml_is_contiguous(rep) || rep->view_src != NULL || (rep->flags & GGML_TENSOR_FLAG_OUTPUT)) return false;
    if (!b200_uses_build(bc, cgraph)) return false;
    if (bc->use_cnt[b200_use_slot(bc, rep)] != 1) return false;          /* exactly one reader, inside this graph */
    bc->deferred[bc->n_deferred++] = rep;
    bc->fused_nodes++;                                                     /* (taken back if it has to be computed after all) */
    return true;
}
static enum ggml_status b200_compute_glue(struct b200_backend_context *bc, struct ggml_tensor *node);
/* before `node` runs on its own: the deferred REPEATs it reads */
static enum ggml_status b200_materialize_deferred_for(struct b200_backend_context *bc, const struct ggml_tensor *node) {
    if (bc->n_deferred == 0) return GGML_STATUS_SUCCESS;
    for (int k = 0; k <= GGML_MAX_SRC; k++) {
        struct ggml_tensor *s = k < GGML_MAX_SRC ? node->src[k] : node->view_src;
        if (!s || b200_deferred_index(bc, s) < 0) continue;
        b200_deferred_drop(bc, s);
        bc->fused_nodes--;
        const enum ggml_status st = b200_compute_glue(bc, s);
        if (st != GGML_STATUS_SUCCESS) return st;
    }
    return GGML_STATUS_SUCCESS;
}
/* a REPEAT of a row to the shape of `like` that this group may fold in: deferred earlier, or one of nodes[i .. i + n) */
static bool b200_foldable_repeat(const struct b200_backend_context *bc, const struct ggml_cgraph *cgraph, const struct ggml_tensor *rep, const struct ggml_tensor *like,
                                 int i, int n) {
    if (!rep || !b200_repeat_of_row(rep, like)) return false;
    if (b200_deferred_index(bc, rep) >= 0) return true;
    for (int j = i; j < i + n; j++)
        if (cgraph->nodes[j] == rep) return true;
    return false;
}

/* NORM -> MUL(REPEAT(gain)) -> ADD(REPEAT(bias)); REPEATs of this group may sit between its nodes */
static int b200_try_fuse_norm_repeat(struct b200_backend_context *bc, struct ggml_cgraph *cgraph, int i, int last, enum ggml_status *st) {
    struct ggml_tensor *n0 = cgraph->nodes[i], *m2 = NULL, *a4 = NULL;
    if (!bc->opt_fuse || (n0->op != GGML_OP_NORM && n0->op != GGML_OP_RMS_NORM) || !b200_glue_supported(n0) || !ggml_is_contiguous(n0)) return 0;
    int j = i + 1;
    while (j < last && cgraph->nodes[j]->op == GGML_OP_REPEAT) j++;
    if (j >= last || cgraph->nodes[j]->op != GGML_OP_MUL) return 0;
    m2 = cgraph->nodes[j++];
    while (j < last && cgraph->nodes[j]->op == GGML_OP_REPEAT) j++;
    if (j >= last || cgraph->nodes[j]->op != GGML_OP_ADD) return 0;
    a4 = cgraph->nodes[j];
    const int n = j - i + 1;
    if (n > 5) return 0;
    struct ggml_tensor *r1 = m2->src[0] == n0 ? m2->src[1] : m2->src[0], *r3 = a4->src[0] == m2 ? a4->src[1] : a4->src[0];
    if (!b200_foldable_repeat(bc, cgraph, r1, n0, i, n) || !b200_foldable_repeat(bc, cgraph, r3, n0, i, n) || r1 == r3) return 0;
    if (!b200_binary_of(m2, GGML_OP_MUL, r1, n0) || !b200_binary_of(a4, GGML_OP_ADD, m2, r3)) return 0;
    for (int k = i + 1; k < i + n; k++)                          /* no foreign REPEAT inside the window */
        if (cgraph->nodes[k]->op == GGML_OP_REPEAT && cgraph->nodes[k] != r1 && cgraph->nodes[k] != r3) return 0;
    if (!ggml_is_contiguous(a4) || !ggml_are_same_shape(a4, n0) || a4->view_src != NULL) return 0;
    if (a4->data != n0->data && a4->data != r1->data && a4->data != m2->data && a4->data != r3->data) return 0;      /* not an allocator-managed graph */
    if (!b200_uses_build(bc, cgraph)) return 0;
    if (!b200_read_only_by_group(bc, cgraph, n0, i, n) || !b200_read_only_by_group(bc, cgraph, r1, i, n) || !b200_read_only_by_group(bc, cgraph, m2, i, n) ||
        !b200_read_only_by_group(bc, cgraph, r3, i, n))
        return 0;
    const struct ggml_tensor *x = n0->src[0];
    if (b200_ranges_overlap(a4, r1->src[0]) || b200_ranges_overlap(a4, r3->src[0])) return 0;
    if (b200_ranges_overlap(a4, x) && !(a4->data == x->data && ggml_is_contiguous(x))) return 0;       /* row by row in place is fine, a shifted overlap is not */
    b200_tensor a, g, b, d;
    float eps;
    memcpy(&eps, n0->op_params, sizeof(eps));
    if (!b200_fill_tensor(x, &a) || !b200_fill_tensor(r1->src[0], &g) || !b200_fill_tensor(r3->src[0], &b) || !b200_fill_tensor(a4, &d)) return 0;
    b200_deferred_drop(bc, r1);
    b200_deferred_drop(bc, r3);
    bc->graph_managed = 1;
    *st = b200_glue_status(bc, a4, b200_op_norm(bc->ctx, &a, &g, &b, &d, eps, n0->op == GGML_OP_RMS_NORM));
    return n;
}

static int b200_next_real(const struct ggml_cgraph *cgraph, int j, int last) {
    while (j < last && (ggml_is_empty(cgraph->nodes[j]) || b200_op_is_noop(cgraph->nodes[j]->op))) j++;
    return j;
}
/* ROPE -> CPY into a contiguous F16 / F32 destination (the rotated k of a token into the KV cache, examples/gpt-j/main.cpp:473-484): the rotation
 * writes the cache directly.  The F32 result is skipped, so nothing else may read it -- nor, when the ROPE is the in-place form, the memory it
 * shares with the projection's output: every tensor up the chain of views to the producing node has exactly the next one as its reader. */
static int b200_try_fuse_rope_cpy(struct b200_backend_context *bc, struct ggml_cgraph *cgraph, int i, int last, enum ggml_status *st) {
    struct ggml_tensor *rope = cgraph->nodes[i];
    if (!bc->opt_fuse || !bc->graph_managed || rope->op != GGML_OP_ROPE || rope->type != GGML_TYPE_F32 || !b200_glue_supported(rope)) return 0;
    const int j = b200_next_real(cgraph, i + 1, last);
    if (j >= last) return 0;
    struct ggml_tensor *cp = cgraph->nodes[j];
    if (cp->op != GGML_OP_CPY || cp->src[0] != rope || (cp->type != GGML_TYPE_F16 && cp->type != GGML_TYPE_F32) || !ggml_is_contiguous(cp) ||
        ggml_nelements(cp) != ggml_nelements(rope) || !ggml_is_contiguous(rope) || cp->data == NULL || !b200_in_device_buffer(cp))
        return 0;
    const int n = j - i + 1;
    if (!b200_uses_build(bc, cgraph) || !b200_read_only_by_group(bc, cgraph, rope, i, n)) return 0;
    if (rope->data == rope->src[0]->data) {                 /* in place: the un-rotated values stay behind in that memory */
        const struct ggml_tensor *chain[10], *t = rope->src[0];
        int n_chain = 0;
        chain[n_chain++] = rope;
        while (t != NULL) {                                 /* (a view's view_src is the root of the chain, not its immediate parent: count every member) */
            int refs = 0;
            for (int k = 0; k < n_chain; k++) refs += b200_refs_from(chain[k], t);
            if ((t->flags & GGML_TENSOR_FLAG_OUTPUT) || bc->use_cnt[b200_use_slot(bc, t)] != refs) return 0;
            if (!b200_op_is_noop(t->op) || t->op == GGML_OP_NONE) break;         /* the producing node (or a leaf) */
            if (n_chain == 10) return 0;
            chain[n_chain++] = t;
            t = t->src[0];
        }
    }
    if (b200_ranges_overlap(cp, rope->src[0]) || b200_ranges_overlap(cp, rope->src[1])) return 0;
    b200_tensor a, pos, d;
    if (!b200_fill_tensor(rope->src[0], &a) || !b200_fill_tensor(rope->src[1], &pos)) return 0;
    memset(&d, 0, sizeof(d));
    d.type = (int32_t)cp->type;
    d.data = cp->data;
    const int64_t es = (int64_t)ggml_type_size(cp->type);
    for (int k = 0; k < 4; k++) d.ne[k] = rope->ne[k];
    d.nb[0] = es; d.nb[1] = d.nb[0] * d.ne[0]; d.nb[2] = d.nb[1] * d.ne[1]; d.nb[3] = d.nb[2] * d.ne[2];
    const int32_t *op = (const int32_t *)rope->op_params;
    b200_rope_params rp;
    memset(&rp, 0, sizeof(rp));
    rp.n_dims = op[1]; rp.mode = op[2]; rp.n_ctx = op[3]; rp.n_orig_ctx = op[4];
    memcpy(&rp.freq_base, op + 5, sizeof(float));
    memcpy(&rp.freq_scale, op + 6, sizeof(float));
    memcpy(&rp.ext_factor, op + 7, sizeof(float));
    memcpy(&rp.attn_factor, op + 8, sizeof(float));
    memcpy(&rp.beta_fast, op + 9, sizeof(float));
    memcpy(&rp.beta_slow, op + 10, sizeof(float));
    memcpy(&rp.xpos_base, op + 11, sizeof(float));
    { bool down; memcpy(&down, op + 12, sizeof(bool)); rp.xpos_down = down ? 1 : 0; }
    const int rc = b200_op_rope(bc->ctx, &a, &pos, &d, &rp);
    if (rc == B200_ERR_UNSUPPORTED) return 0;
    *st = b200_glue_status(bc, cp, rc);
    return n;
}

/* The attention of a decode step: MUL_MAT(K, Q) -> SCALE -> DIAG_MASK_INF -> SOFT_MAX (each in place of the one before) -> MUL_MAT(V, .) ->
 * PERMUTE(0, 2, 1, 3) -> CPY / CONT into a contiguous [n_embd][N] (examples/gpt-j/main.cpp:490-530, examples/gpt-2/main-backend.cpp:567-610), views
 * in between, as ONE launch (b200_op_attention_decode) under the reader-count rule above. */
static int b200_try_fuse_attention(struct b200_backend_context *bc, struct ggml_cgraph *cgraph, int i, int last, enum ggml_status *st) {
    struct ggml_tensor *kq = cgraph->nodes[i];
    if (!bc->opt_fuse || kq->op != GGML_OP_MUL_MAT || kq->type != GGML_TYPE_F32) return 0;
    const struct ggml_tensor *K = kq->src[0], *Q = kq->src[1];
    if (!K || !Q || Q->type != GGML_TYPE_F32 || (K->type != GGML_TYPE_F32 && K->type != GGML_TYPE_F16)) return 0;
    const int64_t hd = Q->ne[0], N = Q->ne[1], H = Q->ne[2], T = K->ne[1];
    if (K->ne[0] != hd || K->ne[2] != H || K->ne[3] != 1 || Q->ne[3] != 1 || N > 8 || T > 1024 || hd > 1024) return 0;
    if (K->nb[0] != ggml_type_size(K->type) || Q->nb[0] != sizeof(float) || !b200_glue_supported_srcs(kq)) return 0;
    int j = b200_next_real(cgraph, i + 1, last);
    if (j >= last) return 0;
    struct ggml_tensor *sc = cgraph->nodes[j];
    if (sc->op != GGML_OP_SCALE || !b200_inplace_child(sc, kq)) return 0;
    j = b200_next_real(cgraph, j + 1, last);
    if (j >= last) return 0;
    struct ggml_tensor *mk = cgraph->nodes[j];
    if (mk->op != GGML_OP_DIAG_MASK_INF || !b200_inplace_child(mk, sc)) return 0;
    j = b200_next_real(cgraph, j + 1, last);
    if (j >= last) return 0;
    struct ggml_tensor *sm = cgraph->nodes[j];
    if (sm->op != GGML_OP_SOFT_MAX || !b200_inplace_child(sm, mk) || sm->src[1] != NULL || sm->src[2] != NULL) return 0;
    float s, scale, max_bias;
    memcpy(&s, sc->op_params, sizeof(s));
    memcpy(&scale, (const float *)sm->op_params + 0, sizeof(scale));
    memcpy(&max_bias, (const float *)sm->op_params + 1, sizeof(max_bias));
    if (max_bias != 0.0f) return 0;
    j = b200_next_real(cgraph, j + 1, last);
    if (j >= last) return 0;
    struct ggml_tensor *pv = cgraph->nodes[j];
    if (pv->op != GGML_OP_MUL_MAT || pv->src[1] != sm || pv->type != GGML_TYPE_F32 || !ggml_is_contiguous(pv)) return 0;
    const struct ggml_tensor *V = pv->src[0];
    if (!V || (V->type != GGML_TYPE_F32 && V->type != GGML_TYPE_F16) || V->ne[0] != T || V->ne[1] != hd || V->ne[2] != H || V->ne[3] != 1 ||
        V->nb[0] != ggml_type_size(V->type) || !b200_in_device_buffer(V) || b200_tensor_is_split(V) || b200_deferred_index(bc, V) >= 0)
        return 0;
    j = b200_next_real(cgraph, j + 1, last);
    if (j >= last) return 0;
    struct ggml_tensor *cp = cgraph->nodes[j];
    if ((cp->op != GGML_OP_CPY && cp->op != GGML_OP_CONT) || cp->type != GGML_TYPE_F32 || !ggml_is_contiguous(cp) || ggml_nelements(cp) != hd * H * N) return 0;
    const struct ggml_tensor *perm = cp->src[0];
    if (!perm || perm->op != GGML_OP_PERMUTE || perm->src[0] != pv || perm->ne[0] != hd || perm->ne[1] != H || perm->ne[2] != N || perm->ne[3] != 1 ||
        perm->nb[0] != sizeof(float) || perm->nb[1] != pv->nb[2] || perm->nb[2] != pv->nb[1] || perm->data != pv->data)
        return 0;
    const int n = j - i + 1;
    if (!b200_uses_build(bc, cgraph)) return 0;
    const struct ggml_tensor *inter[6] = { kq, sc, mk, sm, pv, perm };
    for (int k = 0; k < 6; k++)
        if (!b200_read_only_by_group(bc, cgraph, inter[k], i, n)) return 0;
    if (b200_ranges_overlap(cp, K) || b200_ranges_overlap(cp, V) || b200_ranges_overlap(cp, Q)) return 0;
    b200_tensor tq, tk, tv, td;
    if (!b200_fill_tensor(Q, &tq) || !b200_fill_tensor(K, &tk) || !b200_fill_tensor(V, &tv)) return 0;
    memset(&td, 0, sizeof(td));
    td.type = (int32_t)GGML_TYPE_F32;
    td.data = cp->data;
    td.ne[0] = hd; td.ne[1] = H; td.ne[2] = N; td.ne[3] = 1;
    td.nb[0] = (int64_t)sizeof(float); td.nb[1] = td.nb[0] * hd; td.nb[2] = td.nb[1] * H; td.nb[3] = td.nb[2] * N;
    const int rc = b200_op_attention_decode(bc->ctx, &tq, &tk, &tv, &td, s * scale, ((const int32_t *)mk->op_params)[0]);
    if (rc == B200_ERR_UNSUPPORTED) return 0;
    *st = b200_glue_status(bc, cp, rc);
    /* launches saved: of the six real nodes one remains */
    return n;
}

/* MUL_MAT(decode) [-> ADD(REPEAT(bias))] [-> GELU] [-> ADD(residual)] [-> ADD(second residual)]; the REPEAT deferred earlier or right after the
 * mul_mat.  GPT-J: fc -> bias -> GELU; proj -> bias; attention projection -> + MLP branch -> + residual stream (examples/gpt-j/main.cpp:520-559) */
static int b200_try_fuse_mul_mat_repeat(struct b200_backend_context *bc, struct ggml_cgraph *cgraph, int i, int last, enum ggml_status *st) {
    struct ggml_tensor *mm = cgraph->nodes[i];
    if (!bc->opt_fuse || i + 1 >= last || mm->src[1]->ne[1] > 8 || mm->ne[2] != 1 || mm->ne[3] != 1 || b200_tensor_is_split(mm->src[0])) return 0;
    b200_epilogue epi;
    memset(&epi, 0, sizeof(epi));
    const struct ggml_tensor *inter[6], *rep = NULL, *residual[2] = { NULL, NULL };
    int n_inter = 0, n = 1, n_res = 0;
    struct ggml_tensor *cur = mm;
    {   /* bias */
        int k = n;
        if (i + k < last && cgraph->nodes[i + k]->op == GGML_OP_REPEAT) k++;
        if (i + k < last) {
            struct ggml_tensor *add = cgraph->nodes[i + k];
            if (add->op == GGML_OP_ADD && ggml_is_contiguous(add)) {
                struct ggml_tensor *r = add->src[0] == mm ? add->src[1] : add->src[0];
                if (b200_foldable_repeat(bc, cgraph, r, mm, i, k + 1) && b200_binary_of(add, GGML_OP_ADD, r, mm) && (k == n || cgraph->nodes[i + n] == r)) {
                    rep = r;
                    epi.bias_dev = (const float *)r->src[0]->data;
                    inter[n_inter++] = cur;
                    inter[n_inter++] = r;
                    cur = add;
                    n = k + 1;
                }
            }
        }
    }
    if (rep && i + n < last) {
        struct ggml_tensor *nx = cgraph->nodes[i + n];
        if (nx->op == GGML_OP_UNARY && ggml_get_unary_op(nx) == GGML_UNARY_OP_GELU && nx->src[0] == cur && ggml_is_contiguous(nx) && nx->type == GGML_TYPE_F32) {
            epi.act = B200_EPI_GELU;
            inter[n_inter++] = cur;
            cur = nx;
            n++;
        }
    }
    while (n_res < 2 && i + n < last) {
        struct ggml_tensor *nx = cgraph->nodes[i + n];
        if (nx->op != GGML_OP_ADD || nx->type != GGML_TYPE_F32 || (nx->src[0] != cur && nx->src[1] != cur) || nx->src[0] == nx->src[1] || !ggml_is_contiguous(nx)) break;
        const struct ggml_tensor *r = nx->src[0] == cur ? nx->src[1] : nx->src[0];
        if (r->type != GGML_TYPE_F32 || !ggml_are_same_shape(r, mm) || !ggml_is_contiguous(r) || !b200_in_device_buffer(r) || r->data == NULL ||
            b200_deferred_index(bc, r) >= 0)
            break;
        residual[n_res++] = r;
        inter[n_inter++] = cur;
        cur = nx;
        n++;
    }
    if (n == 1) return 0;
    epi.residual_dev = residual[0] ? (const float *)residual[0]->data : NULL;
    epi.residual2_dev = residual[1] ? (const float *)residual[1]->data : NULL;
    if (cur->view_src != NULL || !ggml_are_same_shape(cur, mm)) return 0;
    bool managed = false;
    for (int k = 0; k < n_inter; k++) managed |= cur->data == inter[k]->data;
    if (!managed) return 0;                                    /* not an allocator-managed graph */
    if (!b200_uses_build(bc, cgraph)) return 0;
    for (int k = 0; k < n_inter; k++)
        if (!b200_read_only_by_group(bc, cgraph, inter[k], i, n)) return 0;
    if (b200_ranges_overlap(cur, mm->src[1]) || (rep && b200_ranges_overlap(cur, rep->src[0]))) return 0;
    for (int k = 0; k < n_res; k++)
        if (b200_ranges_overlap(cur, residual[k]) && cur->data != residual[k]->data) return 0;
    b200_mul_mat_args args;
    if (!b200_fill_mul_mat_args(mm, &args)) return 0;
    args.dst_dev = (float *)cur->data;
    const int rc = b200_mul_mat_fused(bc->ctx, &args, &epi);
    if (rc == B200_ERR_UNSUPPORTED) return 0;          /* not a single-launch shape: the operators run one by one */
    if (rep) b200_deferred_drop(bc, rep);
    bc->graph_managed = 1;
    *st = b200_glue_status(bc, mm, rc);
    return n;
}

/*   MUL_MAT(quantized, decode shape) -> ADD(bias row) [-> GELU] [-> ADD(residual)]     the GEMV's epilogue (main-backend.cpp:614-625, :659-699)
 * under the same rule: every follower computes in place of its predecessor.  Returns the nodes consumed (0 = no fusion). */
static int b200_try_fuse_mul_mat(struct b200_backend_context *bc, struct ggml_cgraph *cgraph, int i, int last, enum ggml_status *st) {
    struct ggml_tensor *mm = cgraph->nodes[i];
    if (!bc->opt_fuse || i + 1 >= last || mm->src[1]->ne[1] > 8 || mm->ne[2] != 1 || mm->ne[3] != 1 || b200_tensor_is_split(mm->src[0])) return 0;
    b200_epilogue epi;
    memset(&epi, 0, sizeof(epi));
    struct ggml_tensor *cur = mm;
    int n = 1;
    struct ggml_tensor *nx = cgraph->nodes[i + n];
    if (nx->op == GGML_OP_ADD && b200_inplace_child(nx, cur) && b200_row_vector(nx->src[1], mm->ne[0]) && nx->src[1]->data != NULL) {
        epi.bias_dev = (const float *)nx->src[1]->data;
        cur = nx;
        n++;
    }
    if (i + n < last) {
        nx = cgraph->nodes[i + n];
        if (nx->op == GGML_OP_UNARY && ggml_get_unary_op(nx) == GGML_UNARY_OP_GELU && b200_inplace_child(nx, cur)) {
            epi.act = B200_EPI_GELU;
            cur = nx;
            n++;
        }
    }
    if (i + n < last) {
        nx = cgraph->nodes[i + n];
        const struct ggml_tensor *r = nx->src[1];
        if (nx->op == GGML_OP_ADD && b200_inplace_child(nx, cur) && r && r->type == GGML_TYPE_F32 && ggml_are_same_shape(r, mm) && ggml_is_contiguous(r) &&
            b200_in_device_buffer(r) && r->data != NULL) {
            epi.residual_dev = (const float *)r->data;
            cur = nx;
            n++;
        }
    }
    if (n == 1) return 0;
    b200_mul_mat_args args;
    if (!b200_fill_mul_mat_args(mm, &args)) return 0;
    const int rc = b200_mul_mat_fused(bc->ctx, &args, &epi);
    if (rc == B200_ERR_UNSUPPORTED) return 0;          /* not a single-launch shape: the operators run one by one */
    *st = b200_glue_status(bc, mm, rc);
    return n;
}

static int b200_try_fuse(struct b200_backend_context *bc, struct ggml_cgraph *cgraph, int i, int last, enum ggml_status *st) {
    struct ggml_tensor *n0 = cgraph->nodes[i];
    if (!bc->opt_fuse || i + 2 >= last) return 0;
    struct ggml_tensor *n1 = cgraph->nodes[i + 1], *n2 = cgraph->nodes[i + 2];
    b200_tensor a, g, b, d;
    if ((n0->op == GGML_OP_NORM || n0->op == GGML_OP_RMS_NORM) && n1->op == GGML_OP_MUL && n2->op == GGML_OP_ADD && b200_glue_supported(n0) &&
        b200_inplace_child(n1, n0) && b200_inplace_child(n2, n1) && b200_row_vector(n1->src[1], n0->ne[0]) && b200_row_vector(n2->src[1], n0->ne[0]) &&
        n2->nb[0] == sizeof(float)) {
        float eps;
        memcpy(&eps, n0->op_params, sizeof(eps));
        if (!b200_fill_tensor(n0->src[0], &a) || !b200_fill_tensor(n1->src[1], &g) || !b200_fill_tensor(n2->src[1], &b) || !b200_fill_tensor(n2, &d)) return 0;
        *st = b200_glue_status(bc, n2, b200_op_norm(bc->ctx, &a, &g, &b, &d, eps, n0->op == GGML_OP_RMS_NORM));
        return 3;
    }
    if (n0->op == GGML_OP_SCALE && n1->op == GGML_OP_DIAG_MASK_INF && n2->op == GGML_OP_SOFT_MAX && b200_glue_supported(n0) && b200_glue_supported(n2) &&
        b200_inplace_child(n1, n0) && b200_inplace_child(n2, n1) && n2->src[1] == NULL) {
        float s, scale, max_bias;
        memcpy(&s, n0->op_params, sizeof(s));
        memcpy(&scale, (const float *)n2->op_params + 0, sizeof(scale));
        memcpy(&max_bias, (const float *)n2->op_params + 1, sizeof(max_bias));
        if (max_bias != 0.0f) return 0;
        if (!b200_fill_tensor(n0->src[0], &a) || !b200_fill_tensor(n2, &d)) return 0;
        *st = b200_glue_status(bc, n2, b200_op_soft_max(bc->ctx, &a, NULL, &d, s * scale, 0.0f, ((const int32_t *)n1->op_params)[0]));
        return 3;
    }
    return 0;
}

/* ---- GGML_OP_MUL_MAT_ID (ggml_compute_forward_mul_mat_id, src/ggml.c:12101): dst[:, s, t] = as[:, :, ids[s, t]] x b[:, s % ne11, t].
 * The expert matrices are ordinary repacked Q4_0 / Q8_0 slices of `as`, so every (slot, token) pair is one mul_mat of the path; the pairs
 * of one token go down as ONE b200_mul_mat_batch (a decode token whose experts share src1 -- b broadcast over the slots -- is one launch).
 * The ids are read back to the host first, like the reference's CUDA backend does (src/ggml-cuda.cu: ggml_cuda_mul_mat_id). */
#define B200_MAX_EXPERTS_USED 16

static bool b200_mul_mat_id_supported(const struct ggml_tensor *dst) {
    const struct ggml_tensor *as = dst->src[0], *b = dst->src[1], *ids = dst->src[2];
    if (!as || !b || !ids || !(b200_type_is_repacked(as->type) || b200_type_is_wire(as->type)) || b->type != GGML_TYPE_F32 || ids->type != GGML_TYPE_I32 ||
        dst->type != GGML_TYPE_F32)
        return false;
    if (b200_buffer_is_split(as->buffer) || !b200_glue_supported_srcs(dst)) return false;
    struct b200_qloc loc;
    if (b200_type_is_wire(as->type) ? !ggml_is_contiguous(as) : !b200_locate_quantized(as, &loc)) return false;
    if (as->ne[3] != 1 || b->ne[3] != 1 || ids->ne[2] != 1 || ids->ne[3] != 1 || ids->ne[0] > B200_MAX_EXPERTS_USED) return false;
    if (b->nb[0] != sizeof(float) || b->nb[1] % 16 != 0 || b->nb[2] % 16 != 0 || !b200_views_16(b) || ids->nb[0] != sizeof(int32_t)) return false;
    return ggml_is_contiguous(dst) && as->ne[0] <= 131072;
}

static enum ggml_status b200_compute_mul_mat_id(struct b200_backend_context *bc, struct ggml_tensor *dst) {
    const struct ggml_tensor *as = dst->src[0], *b = dst->src[1], *ids = dst->src[2];
    struct b200_qloc loc;
    memset(&loc, 0, sizeof(loc));
    if (as && b200_type_is_wire(as->type) && b200_mul_mat_id_supported(dst)) {
        loc.base = as->data;                  /* wire-format blocks: the expert's slice starts e * m * nb blocks in */
        loc.total_blocks = ggml_nelements(as) / B200_QK;
    } else if (!b200_mul_mat_id_supported(dst) || !b200_locate_quantized(as, &loc)) {
        fprintf(stderr, "ggml-b200: MUL_MAT_ID %s is outside this backend's path; no CPU fallback\n", as ? ggml_type_name(as->type) : "?");
        return GGML_STATUS_FAILED;
    }
    const int64_t n_used = ids->ne[0], n_tok = ids->ne[1], n_as = as->ne[2], m = as->ne[1], nb = as->ne[0] / B200_QK;
    if (n_used == 0 || n_tok == 0) return GGML_STATUS_SUCCESS;
    /* the ids (possibly a column slice of a wider matrix: row pitch nb[1]) -> host; they may have been produced on this stream */
    const size_t span = (size_t)(n_tok - 1) * ids->nb[1] + (size_t)n_used * sizeof(int32_t);
    char *hids = (char *)malloc(span);
    if (!hids) return GGML_STATUS_ALLOC_FAILED;
    int rc = b200_synchronize(bc->ctx);
    if (rc == B200_OK) rc = b200_download(bc->ctx, hids, ids->data, span);
    enum ggml_status st = GGML_STATUS_SUCCESS;
    for (int64_t t = 0; t < n_tok && rc == B200_OK; t++) {
        b200_mul_mat_args args[B200_MAX_EXPERTS_USED];
        for (int64_t s = 0; s < n_used; s++) {
            const int32_t e = *(const int32_t *)(hids + (size_t)t * ids->nb[1] + (size_t)s * sizeof(int32_t));
            if (e < 0 || e >= n_as) {
                fprintf(stderr, "ggml-b200: MUL_MAT_ID expert id %d out of range [0, %lld)\n", (int)e, (long long)n_as);
                free(hids);
                return GGML_STATUS_FAILED;
            }
            b200_mul_mat_args *a = &args[s];
            memset(a, 0, sizeof(*a));
            a->type = (int32_t)as->type;
            a->src0_dev = loc.base;
            a->src0_nblocks_total = loc.total_blocks;
            a->src0_block_off = loc.block_off + (int64_t)e * m * nb;
            a->ne00 = as->ne[0]; a->ne01 = m; a->ne02 = 1; a->ne03 = 1;
            a->src1_dev = (const float *)((const char *)b->data + (size_t)(s % b->ne[1]) * b->nb[1] + (size_t)t * b->nb[2]);
            a->ne11 = 1; a->ne12 = 1; a->ne13 = 1;
            a->nb11 = b->nb[1]; a->nb12 = b->nb[1]; a->nb13 = b->nb[1];
            a->dst_dev = (float *)((char *)dst->data + (size_t)s * dst->nb[1] + (size_t)t * dst->nb[2]);
        }
        rc = b200_mul_mat_batch(bc->ctx, args, (int)n_used);
    }
    free(hids);
    if (rc != B200_OK) {
        fprintf(stderr, "ggml-b200: MUL_MAT_ID failed (%d): %s\n", rc, b200_last_error(bc->ctx));
        st = rc == B200_ERR_ALLOC ? GGML_STATUS_ALLOC_FAILED : GGML_STATUS_FAILED;
    }
    return st;
}

GGML_CALL static bool b200_backend_supports_op(ggml_backend_t backend, const struct ggml_tensor *op) {
    GGML_UNUSED(backend);
    if (b200_op_is_noop(op->op)) return true;
    if (op->op == GGML_OP_MUL_MAT)
        return b200_mul_mat_supported(op) || ((b200_dense_mul_mat_supported(op) || b200_wire_mul_mat_supported(op)) && b200_glue_supported_srcs(op));
    if (op->op == GGML_OP_MUL_MAT_ID) return b200_mul_mat_id_supported(op);
    return b200_glue_supported(op);
}

static bool b200_fill_mul_mat_args(struct ggml_tensor *dst, b200_mul_mat_args *args) {
    const struct ggml_tensor *a = dst->src[0], *b = dst->src[1];
    if (!b200_mul_mat_supported(dst)) {
        fprintf(stderr, "ggml-b200: MUL_MAT %s x %s not supported by this backend (no CPU fallback)\n",
                a ? ggml_type_name(a->type) : "?", b ? ggml_type_name(b->type) : "?");
        return false;
    }
    if (b200_tensor_is_split(a)) return false;          /* (handled by b200_compute_split_mul_mat / the row-split plan) */
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


/* ---- MUL_MAT with src0 in a split buffer ------------------------------------------------------------------------ */

static bool b200_reserve_dev(b200_ctx *cd, void **p, size_t *cur, size_t bytes) {
    if (*cur >= bytes) return true;
    if (*p) b200_free(cd, *p);
    *p = NULL;
    *cur = 0;
    if (b200_malloc(cd, p, bytes + bytes / 4 + 4096) != B200_OK) return false;
    *cur = bytes + bytes / 4 + 4096;
    return true;
}

/* one node, any n: every device that owns rows multiplies its slice (its own stream, src1 copied over NVLink), the dst rows are
 * copied into place on the backend's device (strided for n > 1: dst is [n][m] with m contiguous, the layout issue the reference
 * notes at src/ggml-cuda.cu:1592-1608).  Synchronous: the reference's loop waits on per-device events the same way (:1618-1647). */
static enum ggml_status b200_compute_split_mul_mat(struct b200_backend_context *bc, struct ggml_tensor *dst) {
    const struct ggml_tensor *a = dst->src[0], *b = dst->src[1];
    if (!b200_mul_mat_supported(dst) || !b200_tensor_is_split(a)) {
        fprintf(stderr, "ggml-b200: MUL_MAT on a split tensor with an unsupported layout\n");
        return GGML_STATUS_FAILED;
    }
    const struct b200_split_extra *e = (const struct b200_split_extra *)a->extra;
    const int64_t k = a->ne[0], m = a->ne[1], n = b->ne[1], nb = k / B200_QK;
    if (b200_synchronize(bc->ctx) != B200_OK) return GGML_STATUS_FAILED;        /* src1 is complete on the backend's device */
    b200_ctx *used[GGML_B200_MAX_DEVICES];
    int n_used = 0;
    for (int d = 0; d < e->ndev; d++) {
        const int64_t rows = e->row0[d + 1] - e->row0[d];
        if (rows <= 0) continue;
        b200_ctx *cd = b200_device_compute_ctx(bc, d);
        if (!cd) return GGML_STATUS_FAILED;
        struct b200_device_state *ds = &g_dev[d];
        const bool local = d == bc->device;
        const float *x = (const float *)b->data;
        size_t nb11 = b->nb[1];
        if (!local) {
            if (!b200_reserve_dev(cd, &ds->xstage, &ds->xstage_size, (size_t)(n * k * 4))) return GGML_STATUS_ALLOC_FAILED;
            if (b200_copy_2d(cd, ds->xstage, (size_t)k * 4, b->data, b->nb[1], (size_t)k * 4, (size_t)n) != B200_OK) return GGML_STATUS_FAILED;
            x = (const float *)ds->xstage;
            nb11 = (size_t)k * 4;
        }
        float *y = (float *)dst->data + e->row0[d];
        const bool direct = local && n == 1;                                   /* a contiguous run of dst */
        if (!direct) {
            if (!b200_reserve_dev(cd, &ds->ystage, &ds->ystage_size, (size_t)(n * rows * 4))) return GGML_STATUS_ALLOC_FAILED;
            y = (float *)ds->ystage;
        }
        b200_mul_mat_args args;
        memset(&args, 0, sizeof(args));
        args.type = (int32_t)a->type;
        args.src0_dev = e->ptr[d];
        args.src0_nblocks_total = rows * nb;
        args.ne00 = k; args.ne01 = rows; args.ne02 = 1; args.ne03 = 1;
        args.src1_dev = x;
        args.ne11 = n; args.ne12 = 1; args.ne13 = 1;
        args.nb11 = nb11; args.nb12 = nb11 * (size_t)n; args.nb13 = args.nb12;
        args.dst_dev = y;
        int rc = b200_mul_mat(cd, &args);
        if (rc == B200_OK && !direct)
            rc = b200_copy_2d(cd, (float *)dst->data + e->row0[d], (size_t)m * 4, y, (size_t)rows * 4, (size_t)rows * 4, (size_t)n);
        if (rc != B200_OK) {
            fprintf(stderr, "ggml-b200: split mul_mat failed on device %d (%d): %s\n", d, rc, b200_last_error(cd));
            return rc == B200_ERR_ALLOC ? GGML_STATUS_ALLOC_FAILED : GGML_STATUS_FAILED;
        }
        used[n_used++] = cd;
    }
    for (int i = 0; i < n_used; i++)
        if (b200_synchronize(used[i]) != B200_OK) return GGML_STATUS_FAILED;
    return GGML_STATUS_SUCCESS;
}

/* A decode run whose weights all live in (the same kind of) split buffer: one row-split plan per device (b200_plan_create with a
 * b200_plan_split, peer arenas addressed directly: one process, peer access instead of CUDA IPC), launched together.  The
 * backend's device holds the real src / dst tensors; the other devices get copies of the outside inputs and scratch vectors for
 * their partial dsts.  *built = 0 when the run cannot go this way (then node by node).  Returns 0, or -1 on a hard error. */
static int b200_build_split_plan(struct b200_backend_context *bc, const struct ggml_cgraph *cgraph, int first, int last, int n,
                                 struct b200_cached_plan *slot, int *built) {
    *built = 0;
    struct ggml_tensor **nodes = (struct ggml_tensor **)malloc(sizeof(*nodes) * (size_t)n);
    b200_mul_mat_args *args = (b200_mul_mat_args *)malloc(sizeof(*args) * (size_t)n);
    int64_t *row0 = (int64_t *)malloc(sizeof(int64_t) * (size_t)n), *mtot = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    size_t *soff = (size_t *)malloc(sizeof(size_t) * (size_t)n);
    int32_t *plain = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int ret = 0, k = 0, ndev = 0;
    bool ok = nodes && args && row0 && mtot && soff && plain;
    for (int i = first; i < last && ok; i++) {
        struct ggml_tensor *node = cgraph->nodes[i];
        if (ggml_is_empty(node) || b200_op_is_noop(node->op)) continue;
        ok = k < n && b200_tensor_is_split(node->src[0]);
        if (ok) nodes[k++] = node;
    }
    ok = ok && k == n;
    if (ok) {
        ndev = ((const struct b200_split_extra *)nodes[0]->src[0]->extra)->ndev;
        ok = ndev >= 2 && ndev <= B200_MAX_RANKS && bc->device < ndev;
        for (int i = 0; i < n && ok; i++) ok = ((const struct b200_split_extra *)nodes[i]->src[0]->extra)->ndev == ndev;
    }
    /* scratch layout (other devices), outside inputs */
    size_t scratch_bytes = 0, xin_bytes = 0;
    slot->n_xin = 0;
    for (int i = 0; i < n && ok; i++) {
        soff[i] = scratch_bytes;
        scratch_bytes += (size_t)((nodes[i]->src[0]->ne[1] + 63) / 64 * 64) * 4;
        const struct ggml_tensor *b = nodes[i]->src[1];
        bool inside = false;
        for (int j = 0; j < i; j++) inside = inside || nodes[j] == b;
        if (!inside) {
            int f = -1;
            for (int x = 0; x < slot->n_xin; x++)
                if (slot->xin_src[x] == b->data) f = x;
            if (f < 0) {
                ok = slot->n_xin < B200_MAX_XIN;
                if (ok) {
                    slot->xin_src[slot->n_xin] = b->data;
                    slot->xin_bytes[slot->n_xin] = (size_t)b->ne[0] * 4;
                    slot->xin_off[slot->n_xin] = xin_bytes;
                    xin_bytes += (size_t)((b->ne[0] + 63) / 64 * 64) * 4;
                    slot->n_xin++;
                }
            }
        }
    }
    b200_ctx *cds[GGML_B200_MAX_DEVICES];
    for (int d = 0; d < ndev && ok; d++) {
        cds[d] = b200_device_compute_ctx(bc, d);
        ok = cds[d] != NULL;
    }
    for (int d = 0; d < ndev && ok; d++)
        for (int p2 = 0; p2 < ndev && ok; p2++)
            if (p2 != d) ok = b200_enable_peer_access(cds[d], p2) == B200_OK;
    if (ok) {
        slot->ndev = ndev;
        for (int r = 0; r < ndev && ok; r++) {
            const bool is_main = r == bc->device;
            if (!is_main) {
                ok = b200_malloc(cds[r], &slot->scratch[r], scratch_bytes + 256) == B200_OK && b200_malloc(cds[r], &slot->xin[r], xin_bytes + 256) == B200_OK;
                if (!ok) break;
            }
        }
    }
    /* per rank: its slices, the dataflow expressed through ITS addresses (the producer's dst address is the consumer's src1).
     * pass 0 = the backend's device only: which results the graph keeps in plain memory (b200_plan_plain_stores: not the dead
     * intermediates whose memory a later node reuses) and the size of the exchange arena; pass 1 = one plan per device.  Every
     * rank carries the SAME export set: an exported op is one whose rows every rank sends to every rank. */
    for (int pass = 0; pass < 2 && ok; pass++) {
        for (int r = 0; r < ndev && ok; r++) {
            const bool is_main = r == bc->device;
            if (pass == 0 && !is_main) continue;
            for (int i = 0; i < n; i++) {
                const struct ggml_tensor *a = nodes[i]->src[0], *b = nodes[i]->src[1];
                const struct b200_split_extra *e = (const struct b200_split_extra *)a->extra;
                const int64_t rows = e->row0[r + 1] - e->row0[r], nbk = a->ne[0] / B200_QK;
                b200_mul_mat_args *g = &args[i];
                memset(g, 0, sizeof(*g));
                g->type = (int32_t)a->type;
                g->src0_dev = rows > 0 ? e->ptr[r] : (const void *)nodes;       /* (no rows: never dereferenced) */
                g->src0_nblocks_total = rows * nbk;
                g->ne00 = a->ne[0]; g->ne01 = rows; g->ne02 = 1; g->ne03 = 1;
                g->ne11 = 1; g->ne12 = 1; g->ne13 = 1;
                g->nb11 = g->nb12 = g->nb13 = (size_t)a->ne[0] * 4;
                int prod = -1;
                for (int j = 0; j < i; j++)
                    if (nodes[j] == b) prod = j;
                if (is_main) {
                    g->src1_dev = (const float *)b->data;
                    g->dst_dev = (float *)nodes[i]->data;
                } else {
                    if (prod >= 0) g->src1_dev = (const float *)((char *)slot->scratch[r] + soff[prod]);
                    else
                        for (int x = 0; x < slot->n_xin; x++)
                            if (slot->xin_src[x] == b->data) g->src1_dev = (const float *)((char *)slot->xin[r] + slot->xin_off[x]);
                    g->dst_dev = (float *)((char *)slot->scratch[r] + soff[i]);
                }
                row0[i] = e->row0[r];
                mtot[i] = a->ne[1];
            }
            b200_plan_split sp;
            memset(&sp, 0, sizeof(sp));
            sp.world = ndev; sp.rank = r; sp.row0 = row0; sp.m_total = mtot;
            if (pass == 0) {
                if (b200_plan_plain_stores(args, n, &sp, plain) != B200_OK) { ok = false; break; }
                plain[n - 1] = 1;                                  /* (the last op is always exported: it keeps the ranks in step) */
                const size_t abytes = b200_plan_arena_bytes(args, n, &sp);
                for (int d = 0; d < ndev && ok; d++)
                    ok = b200_malloc(cds[d], &slot->arena[d], abytes) == B200_OK && b200_memset(cds[d], slot->arena[d], 0, abytes) == B200_OK;
                continue;
            }
            for (int d = 0; d < ndev; d++) sp.peer_arena[d] = slot->arena[d];
            for (int i = 0; i < n; i++)
                if (plain[i]) args[i].flags |= B200_MM_EXPORT;
            const int rc = b200_plan_create(cds[r], args, n, &sp, &slot->dplan[r]);
            if (rc != B200_OK) {
                if (rc != B200_ERR_UNSUPPORTED) {
                    fprintf(stderr, "ggml-b200: row-split b200_plan_create failed on device %d (%d): %s\n", r, rc, b200_last_error(cds[r]));
                    ret = -1;
                }
                ok = false;
            }
        }
    }
    if (ok) *built = 1;
    free(nodes); free(args); free(row0); free(mtot); free(soff); free(plain);
    return ret;
}

static int b200_launch_split_plan(struct b200_backend_context *bc, struct b200_cached_plan *slot) {
    if (b200_synchronize(bc->ctx) != B200_OK) return -1;                        /* the outside inputs are complete */
    const void *main_ptr = NULL;
    GGML_UNUSED(main_ptr);
    for (int d = 0; d < slot->ndev; d++) {
        if (d == bc->device) continue;
        b200_ctx *cd = b200_device_compute_ctx(bc, d);
        for (int x = 0; x < slot->n_xin; x++)
            if (b200_copy_d2d(cd, (char *)slot->xin[d] + slot->xin_off[x], slot->xin_src[x], slot->xin_bytes[x]) != B200_OK) return -1;
    }
    for (int d = 0; d < slot->ndev; d++) {
        b200_ctx *cd = b200_device_compute_ctx(bc, d);
        const int rc = b200_plan_launch(cd, slot->dplan[d]);
        if (rc != B200_OK) {
            fprintf(stderr, "ggml-b200: row-split plan launch failed on device %d (%d): %s\n", d, rc, b200_last_error(cd));
            return -1;
        }
    }
    int bad = 0;
    for (int d = 0; d < slot->ndev; d++)
        if (b200_synchronize(b200_device_compute_ctx(bc, d)) != B200_OK) bad = 1;
    return bad ? -1 : 0;
}

static bool b200_node_is_decode_mul_mat(const struct ggml_tensor *node) {
    if (node->op != GGML_OP_MUL_MAT) return false;
    const struct ggml_tensor *a = node->src[0], *b = node->src[1];
    return a && b && b->ne[1] == 1 && b->ne[2] == 1 && b->ne[3] == 1 && a->ne[2] == 1 && a->ne[3] == 1 && a->ne[0] % 256 == 0 &&
           a->ne[0] <= 32768 && b200_mul_mat_supported(node);
}

/* [first, last) = a maximal run of compute nodes that are all decode-shaped MUL_MATs (no-op nodes in between are skipped);
 * returns how many such nodes it holds and a hash over their addresses and shapes */
static int b200_decode_run(const struct ggml_cgraph *cgraph, int first, int *last_out, uint64_t *key_out, bool *split_out) {
    int n = 0, i = first;
    bool split = false;
    uint64_t key = 1469598103934665603ull;     /* FNV-1a */
    for (; i < cgraph->n_nodes; i++) {
        struct ggml_tensor *node = cgraph->nodes[i];
        if (ggml_is_empty(node) || b200_op_is_noop(node->op)) continue;
        if (!b200_node_is_decode_mul_mat(node)) break;
        const struct ggml_tensor *a = node->src[0], *b = node->src[1];
        if (n == 0) split = b200_tensor_is_split(a);
        else if (split != b200_tensor_is_split(a)) break;          /* weights in a split buffer and in one device's memory do not mix in a plan */
        const uint64_t words[6] = {(uint64_t)(uintptr_t)(split ? a->extra : a->data), (uint64_t)(uintptr_t)b->data, (uint64_t)(uintptr_t)node->data,
                                   (uint64_t)a->ne[0], (uint64_t)a->ne[1], (uint64_t)a->type};
        for (int w = 0; w < 6; w++)
            for (int sh = 0; sh < 64; sh += 8) key = (key ^ ((words[w] >> sh) & 0xff)) * 1099511628211ull;
        n++;
    }
    if (key == 0) key = 1;
    *last_out = i;
    if (key_out) *key_out = key;
    if (split_out) *split_out = split;
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
static int b200_try_run_as_plan(struct b200_backend_context *bc, struct ggml_cgraph *cgraph, int first, int last, int n, uint64_t key, bool split) {
    if (!bc->opt_plans || n < 2) return 0;
    struct b200_cached_plan *slot = NULL;
    for (int i = 0; i < B200_PLAN_CACHE; i++)
        if (bc->plans[i].key == key && bc->plans[i].n_nodes == n) slot = &bc->plans[i];
    if (!slot) {
        slot = &bc->plans[bc->plan_next];
        bc->plan_next = (bc->plan_next + 1) % B200_PLAN_CACHE;
        b200_cached_plan_release(bc, slot);
        slot->key = key;
        slot->n_nodes = n;
        if (split) {
            int built = 0;
            const int rc = b200_build_split_plan(bc, cgraph, first, last, n, slot, &built);
            if (!built) {
                b200_cached_plan_release(bc, slot);        /* frees what was allocated; remember the verdict */
                slot->key = key;
                slot->n_nodes = n;
            }
            if (rc < 0) return -1;
        } else if (b200_build_plan(bc, cgraph, first, last, n, &slot->plan) < 0) {
            return -1;
        }
    }
    if (slot->ndev > 0) {
        if (b200_launch_split_plan(bc, slot) < 0) return -1;
        bc->plan_launches++;
        return 1;
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

static enum ggml_status b200_graph_compute_nodes(struct b200_backend_context *bc, struct ggml_cgraph *cgraph, int first, int last, int *one_unit_next);

GGML_CALL static enum ggml_status b200_backend_graph_compute(ggml_backend_t backend, struct ggml_cgraph *cgraph) {
    struct b200_backend_context *bc = (struct b200_backend_context *)backend->context;
    if (bc->failed) {
        bc->failed = 0;
        return GGML_STATUS_FAILED;
    }
    bc->use_valid = 0;
    bc->graph_managed = 0;
    bc->n_deferred = 0;
    int i = 0;
    while (i < cgraph->n_nodes) {
        struct ggml_tensor *node = cgraph->nodes[i];
        if (ggml_is_empty(node) || b200_op_is_noop(node->op)) {
            if (b200_materialize_deferred_for(bc, node) != GGML_STATUS_SUCCESS) return GGML_STATUS_FAILED;
            i++;
            continue;
        }
        int last = i;
        uint64_t key = 0;
        bool split = false;
        const int n = b200_decode_run(cgraph, i, &last, &key, &split);
        if (n >= 2) {
            for (int j = i; j < last; j++)
                if (b200_materialize_deferred_for(bc, cgraph->nodes[j]) != GGML_STATUS_SUCCESS) return GGML_STATUS_FAILED;
            const int as_plan = b200_try_run_as_plan(bc, cgraph, i, last, n, key, split);
            if (as_plan < 0) return GGML_STATUS_FAILED;
            if (as_plan == 0) {
                const enum ggml_status st = b200_graph_compute_nodes(bc, cgraph, i, last, NULL);
                if (st != GGML_STATUS_SUCCESS) return st;
            }
            i = last;
            continue;
        }
        /* anything else: one unit at a time (a node, a same-input run of mul_mats, or a node with the in-place neighbours that fold
         * into its kernel), so that the next decode run is found where it starts */
        int next = i + 1;
        const enum ggml_status st = b200_graph_compute_nodes(bc, cgraph, i, cgraph->n_nodes, &next);
        if (st != GGML_STATUS_SUCCESS) return st;
        i = next;
    }
    while (bc->n_deferred > 0) {                    /* (every deferred node has a reader in this graph, so nothing is left; kept as a net) */
        struct ggml_tensor *rep = bc->deferred[--bc->n_deferred];
        bc->fused_nodes--;
        const enum ggml_status st = b200_compute_glue(bc, rep);
        if (st != GGML_STATUS_SUCCESS) return st;
    }
    return GGML_STATUS_SUCCESS;
}

/* ggml_backend_graph_plan_create / _free / _compute (src/ggml-backend-impl.h:94-99, src/ggml-backend.c:257-273): the explicit,
 * static form of the above.  Like the reference CPU backend's plan (src/ggml-backend.c:761-790) it keeps a shallow copy of the cgraph,
 * so the graph has to outlive the plan.  The first compute runs node by node (it sizes the scratch areas and builds the
 * persistent-launch plans of the decode runs); the second records the same sequence into a CUDA graph (b200_graph_begin / _end), and from
 * then on a compute is ONE graph launch -- the launch-bound whole-model decode step (GPT-2 117M: ~170 kernels of 2-5 us) stops
 * paying the host's per-launch cost.  The recording freezes addresses and scalar arguments, so every compute first fingerprints
 * what a kernel would read from the nodes (op, op_params, shapes, strides, addresses) and records again when that changed; a
 * graph that cannot be recorded (MUL_MAT_ID reads its routing back, split tensors drive several devices) stays node by node. */
struct b200_graph_plan {
    struct ggml_cgraph cgraph;
    b200_graph *graph;
    uint64_t fingerprint;
    int state;                     /* 0: never computed, 1: computed node by node once, 2: recorded, -1: cannot be recorded */
    int64_t plan_launches, fused_nodes;   /* what one replay adds to the backend's counters */
};

static uint64_t b200_fnv(uint64_t h, const void *p, size_t n) {
    const unsigned char *c = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) h = (h ^ c[i]) * 0x100000001b3ull;
    return h;
}
static uint64_t b200_graph_fingerprint(const struct ggml_cgraph *g, bool *recordable) {
    uint64_t h = 0xcbf29ce484222325ull;
    *recordable = true;
    h = b200_fnv(h, &g->n_nodes, sizeof(g->n_nodes));
    for (int i = 0; i < g->n_nodes; i++) {
        const struct ggml_tensor *t = g->nodes[i];
        if (t->op == GGML_OP_MUL_MAT_ID) *recordable = false;
        for (int k = -1; k < GGML_MAX_SRC; k++) {
            const struct ggml_tensor *x = k < 0 ? t : t->src[k];
            if (!x) continue;
            if (b200_tensor_is_split(x)) *recordable = false;
            h = b200_fnv(h, &x->data, sizeof(x->data));
            h = b200_fnv(h, x->ne, sizeof(x->ne));
            h = b200_fnv(h, x->nb, sizeof(x->nb));
            h = b200_fnv(h, &x->type, sizeof(x->type));
        }
        h = b200_fnv(h, &t->op, sizeof(t->op));
        h = b200_fnv(h, t->op_params, sizeof(t->op_params));
    }
    return h;
}

GGML_CALL static ggml_backend_graph_plan_t b200_backend_graph_plan_create(ggml_backend_t backend, const struct ggml_cgraph *cgraph) {
    GGML_UNUSED(backend);
    struct b200_graph_plan *gp = (struct b200_graph_plan *)calloc(1, sizeof(*gp));
    if (!gp) return NULL;
    gp->cgraph = *cgraph;
    return (ggml_backend_graph_plan_t)gp;
}

GGML_CALL static void b200_backend_graph_plan_free(ggml_backend_t backend, ggml_backend_graph_plan_t plan) {
    GGML_UNUSED(backend);
    struct b200_graph_plan *gp = (struct b200_graph_plan *)plan;
    if (gp) b200_graph_destroy(gp->graph);
    free(plan);
}

GGML_CALL static enum ggml_status b200_backend_graph_plan_compute(ggml_backend_t backend, ggml_backend_graph_plan_t plan) {
    struct b200_backend_context *bc = (struct b200_backend_context *)backend->context;
    struct b200_graph_plan *gp = (struct b200_graph_plan *)plan;
    if (!bc->opt_graphs || gp->state < 0) return b200_backend_graph_compute(backend, &gp->cgraph);
    bool recordable = true;
    const uint64_t fp = b200_graph_fingerprint(&gp->cgraph, &recordable);
    if (!recordable) {
        gp->state = -1;
        return b200_backend_graph_compute(backend, &gp->cgraph);
    }
    if (gp->state == 2 && fp != gp->fingerprint) {          /* the caller changed the graph under the plan: record again */
        b200_graph_destroy(gp->graph);
        gp->graph = NULL;
        gp->state = 1;
    }
    if (gp->state == 2) {
        if (bc->failed) {
            bc->failed = 0;
            return GGML_STATUS_FAILED;
        }
        if (b200_graph_launch(bc->ctx, gp->graph) == B200_OK) {
            bc->plan_launches += gp->plan_launches;
            bc->fused_nodes += gp->fused_nodes;
            return GGML_STATUS_SUCCESS;
        }
        fprintf(stderr, "ggml-b200: graph launch failed: %s\n", b200_last_error(bc->ctx));
        return GGML_STATUS_FAILED;
    }
    if (gp->state == 0) {                                    /* first compute: as graph_compute would (allocations, decode plans) */
        gp->state = 1;
        gp->fingerprint = fp;
        return b200_backend_graph_compute(backend, &gp->cgraph);
    }
    /* record, then launch what was recorded */
    gp->fingerprint = fp;
    if (b200_graph_begin(bc->ctx) != B200_OK) {
        gp->state = -1;
        return b200_backend_graph_compute(backend, &gp->cgraph);
    }
    const int64_t pl0 = bc->plan_launches, fn0 = bc->fused_nodes;
    const enum ggml_status st = b200_backend_graph_compute(backend, &gp->cgraph);
    gp->plan_launches = bc->plan_launches - pl0;
    gp->fused_nodes = bc->fused_nodes - fn0;
    bc->plan_launches = pl0;                                 /* recording launched nothing */
    bc->fused_nodes = fn0;
    b200_graph *g = NULL;
    const int rc = b200_graph_end(bc->ctx, st == GGML_STATUS_SUCCESS ? &g : NULL);
    if (st != GGML_STATUS_SUCCESS || rc != B200_OK || !g) {  /* nothing ran: compute it the plain way and stop trying */
        b200_graph_destroy(g);
        gp->state = -1;
        return b200_backend_graph_compute(backend, &gp->cgraph);
    }
    gp->graph = g;
    gp->state = 2;
    if (b200_graph_launch(bc->ctx, gp->graph) != B200_OK) return GGML_STATUS_FAILED;
    bc->plan_launches += gp->plan_launches;
    bc->fused_nodes += gp->fused_nodes;
    return GGML_STATUS_SUCCESS;
}

/* kernel nodes of the CUDA graph a plan replays (0: the plan computes node by node) */
GGML_CALL int64_t ggml_backend_b200_graph_plan_kernels(ggml_backend_graph_plan_t plan) {
    const struct b200_graph_plan *gp = (const struct b200_graph_plan *)plan;
    return gp && gp->state == 2 ? b200_graph_node_count(gp->graph) : 0;
}

/* nodes [first, last) one after the other; with one_unit_next: only the first unit of work found in that window (a node, a same-input
 * run of mul_mats, a fused group), reporting where the caller continues */
static enum ggml_status b200_graph_compute_nodes(struct b200_backend_context *bc, struct ggml_cgraph *cgraph, int first, int last, int *one_unit_next) {
    bool done_one = false;
    for (int i = first; i < last; i++) {
        struct ggml_tensor *node = cgraph->nodes[i];
        if (one_unit_next) {
            *one_unit_next = i;
            if (done_one) return GGML_STATUS_SUCCESS;
        }
        {
            const enum ggml_status st = b200_materialize_deferred_for(bc, node);
            if (st != GGML_STATUS_SUCCESS) return st;
        }
        if (ggml_is_empty(node) || b200_op_is_noop(node->op)) continue;
        done_one = true;
        if (node->op == GGML_OP_MUL_MAT && b200_tensor_is_split(node->src[0])) {
            const enum ggml_status st = b200_compute_split_mul_mat(bc, node);
            if (st != GGML_STATUS_SUCCESS) return st;
            continue;
        }
        if (node->op == GGML_OP_MUL_MAT && b200_type_is_repacked(node->src[0]->type)) {
            {
                enum ggml_status st = GGML_STATUS_SUCCESS;
                int fused = b200_try_fuse_mul_mat(bc, cgraph, i, last, &st);
                if (fused == 0) fused = b200_try_fuse_mul_mat_repeat(bc, cgraph, i, last, &st);
                if (fused > 0) {
                    if (st != GGML_STATUS_SUCCESS) return st;
                    bc->fused_nodes += fused - 1;
                    i += fused - 1;
                    continue;
                }
            }
            /* gather the run of consecutive MUL_MAT nodes that share this node's src1 and do not depend on one another */
            struct ggml_tensor *run[B200_MAX_RUN];
            int n = 0;
            run[n++] = node;
            while (n < B200_MAX_RUN && i + 1 < last) {
                struct ggml_tensor *next = cgraph->nodes[i + 1];
                if (next->op != GGML_OP_MUL_MAT || next->src[1] != node->src[1] || ggml_is_empty(next) || b200_tensor_is_split(next->src[0])) break;
                bool dep = false;
                for (int j = 0; j < n; j++)
                    if (next->src[0] == run[j] || next->src[0]->view_src == run[j]) dep = true;
                /* (a graph allocator may have given `next` the memory of an earlier result of this run: keep them apart) */
                for (int j = 0; j < n; j++)
                    if (next->data == run[j]->data) dep = true;
                if (dep) break;
                if (b200_materialize_deferred_for(bc, next) != GGML_STATUS_SUCCESS) return GGML_STATUS_FAILED;
                run[n++] = next;
                i++;
            }
            const enum ggml_status st = b200_compute_mul_mat_run(bc, run, n);
            if (st != GGML_STATUS_SUCCESS) return st;
            continue;
        }
        if (node->op == GGML_OP_MUL_MAT_ID) {
            const enum ggml_status st = b200_compute_mul_mat_id(bc, node);
            if (st != GGML_STATUS_SUCCESS) return st;
            continue;
        }
        {
            enum ggml_status st = GGML_STATUS_SUCCESS;
            if (node->op == GGML_OP_REPEAT && b200_try_defer_repeat(bc, cgraph, node)) continue;
            int fused = node->op == GGML_OP_MUL_MAT ? b200_try_fuse_attention(bc, cgraph, i, last, &st) : 0;
            if (fused == 0 && node->op == GGML_OP_ROPE) fused = b200_try_fuse_rope_cpy(bc, cgraph, i, last, &st);
            if (fused == 0) fused = b200_try_fuse(bc, cgraph, i, last, &st);
            if (fused == 0) fused = b200_try_fuse_norm_repeat(bc, cgraph, i, last, &st);
            if (fused > 0) {
                if (st != GGML_STATUS_SUCCESS) return st;
                bc->fused_nodes += fused - 1;
                i += fused - 1;
                continue;
            }
            st = b200_compute_glue(bc, node);
            if (st != GGML_STATUS_SUCCESS) return st;
        }
    }
    if (one_unit_next) *one_unit_next = last;
    return GGML_STATUS_SUCCESS;
}

static enum ggml_status b200_glue_status(struct b200_backend_context *bc, const struct ggml_tensor *node, int rc) {
    if (rc == B200_OK) return GGML_STATUS_SUCCESS;
    fprintf(stderr, "ggml-b200: %s failed (%d): %s\n", ggml_op_name(node->op), rc, b200_last_error(bc->ctx));
    return rc == B200_ERR_ALLOC ? GGML_STATUS_ALLOC_FAILED : GGML_STATUS_FAILED;
}

/* ---- optional SPI entries (src/ggml-backend-impl.h:86-90, :111-116): asynchronous tensor access on the backend's own stream, events ---- */

static bool b200_tensor_on_backend_device(struct b200_backend_context *bc, const struct ggml_tensor *t) {
    ggml_backend_buffer_t buf = t->view_src ? t->view_src->buffer : t->buffer;
    return b200_buffer_is_ours(buf) && ((struct b200_buffer_context *)buf->context)->device == bc->device;
}

GGML_CALL static void b200_backend_set_tensor_async(ggml_backend_t backend, struct ggml_tensor *tensor, const void *data, size_t offset, size_t size) {
    struct b200_backend_context *bc = (struct b200_backend_context *)backend->context;
    ggml_backend_buffer_t buf = tensor->view_src ? tensor->view_src->buffer : tensor->buffer;
    if (b200_tensor_on_backend_device(bc, tensor) && !b200_type_is_repacked(tensor->type)) {
        /* ordered with the graphs computed on this backend: the copy rides the compute stream */
        B200_CHECK(bc->ctx, b200_upload_async(bc->ctx, (char *)tensor->data + offset, data, size));
        return;
    }
    buf->iface.set_tensor(buf, tensor, data, offset, size);      /* repacked types (and foreign buffers) keep the synchronous path */
}

GGML_CALL static void b200_backend_get_tensor_async(ggml_backend_t backend, const struct ggml_tensor *tensor, void *data, size_t offset, size_t size) {
    struct b200_backend_context *bc = (struct b200_backend_context *)backend->context;
    ggml_backend_buffer_t buf = tensor->view_src ? tensor->view_src->buffer : tensor->buffer;
    if (b200_tensor_on_backend_device(bc, tensor) && !b200_type_is_repacked(tensor->type)) {
        B200_CHECK(bc->ctx, b200_download_async(bc->ctx, data, (const char *)tensor->data + offset, size));
        return;
    }
    buf->iface.get_tensor(buf, tensor, data, offset, size);
}

GGML_CALL static ggml_backend_event_t b200_backend_event_new(ggml_backend_t backend) {
    struct b200_backend_context *bc = (struct b200_backend_context *)backend->context;
    b200_event *ev = NULL;
    if (b200_event_create(bc->ctx, &ev) != B200_OK) return NULL;
    ggml_backend_event_t e = (ggml_backend_event_t)malloc(sizeof(struct ggml_backend_event));
    if (!e) {
        b200_event_destroy(ev);
        return NULL;
    }
    e->backend = backend;
    e->context = ev;
    return e;
}

GGML_CALL static void b200_backend_event_free(ggml_backend_event_t event) {
    b200_event_destroy((b200_event *)event->context);
    free(event);
}

GGML_CALL static void b200_backend_event_record(ggml_backend_event_t event) {
    struct b200_backend_context *bc = (struct b200_backend_context *)event->backend->context;
    B200_CHECK(bc->ctx, b200_event_record(bc->ctx, (b200_event *)event->context));
}

GGML_CALL static void b200_backend_event_wait(ggml_backend_t backend, ggml_backend_event_t event) {
    struct b200_backend_context *bc = (struct b200_backend_context *)backend->context;
    if (ggml_backend_is_b200(event->backend)) {
        B200_CHECK(bc->ctx, b200_event_wait(bc->ctx, (b200_event *)event->context));
    } else {
        ggml_backend_event_synchronize(event);      /* another backend's event: nothing to chain on the device, wait on the host */
    }
}

GGML_CALL static void b200_backend_event_synchronize(ggml_backend_event_t event) {
    struct b200_backend_context *bc = (struct b200_backend_context *)event->backend->context;
    B200_CHECK(bc->ctx, b200_event_synchronize((b200_event *)event->context));
}

static struct ggml_backend_i b200_backend_interface = {
    /* .get_name                = */ b200_backend_name,
    /* .free                    = */ b200_backend_free,
    /* .get_default_buffer_type = */ b200_backend_default_buft,
    /* .set_tensor_async        = */ b200_backend_set_tensor_async,
    /* .get_tensor_async        = */ b200_backend_get_tensor_async,
    /* .cpy_tensor_async        = */ NULL,        /* (the core then synchronizes both backends and copies through the buffer interface) */
    /* .synchronize             = */ b200_backend_synchronize,
    /* .graph_plan_create       = */ b200_backend_graph_plan_create,
    /* .graph_plan_free         = */ b200_backend_graph_plan_free,
    /* .graph_plan_compute      = */ b200_backend_graph_plan_compute,
    /* .graph_compute           = */ b200_backend_graph_compute,
    /* .supports_op             = */ b200_backend_supports_op,
    /* .offload_op              = */ NULL,
    /* .event_new               = */ b200_backend_event_new,
    /* .event_free              = */ b200_backend_event_free,
    /* .event_record            = */ b200_backend_event_record,
    /* .event_wait              = */ b200_backend_event_wait,
    /* .event_synchronize       = */ b200_backend_event_synchronize,
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
    bc->opt_fuse = 1;
    bc->opt_graphs = 1;
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

GGML_CALL int64_t ggml_backend_b200_fused_node_count(ggml_backend_t backend) {
    GGML_ASSERT(ggml_backend_is_b200(backend));
    return ((struct b200_backend_context *)backend->context)->fused_nodes;
}

GGML_CALL int ggml_backend_b200_set_option(ggml_backend_t backend, const char *key, int64_t value) {
    GGML_ASSERT(ggml_backend_is_b200(backend));
    if (strcmp(key, "plans") == 0) {
        ((struct b200_backend_context *)backend->context)->opt_plans = value != 0;
        return B200_OK;
    }
    if (strcmp(key, "graphs") == 0) {
        ((struct b200_backend_context *)backend->context)->opt_graphs = value != 0;
        return B200_OK;
    }
    if (strcmp(key, "fuse") == 0) {
        ((struct b200_backend_context *)backend->context)->opt_fuse = value != 0;
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
/* src/ggml-cuda.h:28-29: rows of a matrix split across the devices of one process */
GGML_API GGML_CALL ggml_backend_buffer_type_t ggml_backend_cuda_split_buffer_type(const float *tensor_split) {
    return ggml_backend_b200_split_buffer_type(tensor_split);
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
