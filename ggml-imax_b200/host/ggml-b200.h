/*
 * ggml-b200.h -- public entry points of the B200 ggml backend (host side, plain C).
 *
 * A sibling of the reference's src/ggml-cuda.h: same shape of API, one backend per device, behind the
 * unchanged ggml_backend_i / ggml_backend_buffer_i / ggml_backend_buffer_type_i plugin interfaces
 * (src/ggml-backend-impl.h:18-117).  It accelerates ONE path: GGML_OP_MUL_MAT with Q4_0/Q8_0 src0 and
 * F32 src1, plus the operators either side of it in a GPT-2 / GPT-J graph (GET_ROWS, ADD/MUL/DIV, NORM, SCALE,
 * DIAG_MASK_INF, SOFT_MAX, UNARY, CPY/DUP/CONT, F32/F16 MUL_MAT) so that gpt-2-backend runs on it unchanged;
 * supports_op is false for everything else; there is no CPU fallback inside the backend.
 *
 * The ggml_backend_cuda_* names of src/ggml-cuda.h:19-39 (+ ggml_backend_cuda_reg_devices,
 * src/ggml-cuda.cu:3031-3042) are exported as thin aliases so that a reference core compiled with
 * -DGGML_USE_CUDA -- registry (src/ggml-backend.c:423-426), examples/gpt-2/main-backend.cpp:200-208,
 * tests/test-backend-ops.cpp -- picks this backend up with zero source edits.
 */
#pragma once

#include "ggml.h"
#include "ggml-backend.h"

#ifdef __cplusplus
extern "C" {
#endif

#define GGML_B200_NAME "B200"
#define GGML_B200_MAX_DEVICES 16

GGML_API GGML_CALL ggml_backend_t             ggml_backend_b200_init(int device);              /* NULL on failure */
GGML_API GGML_CALL bool                       ggml_backend_is_b200(ggml_backend_t backend);
GGML_API GGML_CALL ggml_backend_buffer_type_t ggml_backend_b200_buffer_type(int device);      /* device buffers */
GGML_API GGML_CALL ggml_backend_buffer_type_t ggml_backend_b200_host_buffer_type(void);       /* pinned host   */
/* rows of 2-D Q4_0 / Q8_0 matrices split across the devices of this process (src/ggml-cuda.h:28-29): tensor_split = proportions
 * per device (GGML_B200_MAX_DEVICES floats) or NULL for equal shares; MUL_MAT with such a src0 runs on every device that owns rows */
GGML_API GGML_CALL ggml_backend_buffer_type_t ggml_backend_b200_split_buffer_type(const float *tensor_split);
GGML_API GGML_CALL int                        ggml_backend_b200_get_device_count(void);
GGML_API GGML_CALL void                       ggml_backend_b200_get_device_description(int device, char *description, size_t description_size);
GGML_API GGML_CALL void                       ggml_backend_b200_get_device_memory(int device, size_t *free, size_t *total);
GGML_API GGML_CALL int                        ggml_backend_b200_reg_devices(void);             /* registers "B200", "B2001", ... */
/* number of CUDA kernels the backend has launched (instrumentation for benches/tests) */
GGML_API GGML_CALL int64_t                    ggml_backend_b200_launch_count(ggml_backend_t backend);
/* number of cgraphs computed as ONE persistent launch (decode plan, see graph_compute in ggml-b200.c) */
GGML_API GGML_CALL int64_t                    ggml_backend_b200_plan_launch_count(ggml_backend_t backend);
/* graph nodes that were folded into the kernel of a neighbour (NORM+MUL+ADD, SCALE+DIAG_MASK_INF+SOFT_MAX, ...) */
GGML_API GGML_CALL int64_t                    ggml_backend_b200_fused_node_count(ggml_backend_t backend);
/* kernel nodes of the CUDA graph a ggml_backend_graph_plan_t of this backend replays; 0 while (or if) it computes node by node */
GGML_API GGML_CALL int64_t                    ggml_backend_b200_graph_plan_kernels(ggml_backend_graph_plan_t plan);
/* "plans" (0/1: compute runs of decode MUL_MATs as one persistent launch), "fuse" (0/1: fold in-place neighbours into one kernel);
 * everything else forwards to
 * b200_ctx_set_option (include/ggml_b200.h) */
GGML_API GGML_CALL int                        ggml_backend_b200_set_option(ggml_backend_t backend, const char *key, int64_t value);

#ifdef __cplusplus
}
#endif
