#!/usr/bin/env python3
"""TEST INFRASTRUCTURE.  Writes a random-init GPT-2 model in the reference's legacy ggml file format -- the layout
examples/gpt-2/main-backend.cpp:101-436 reads (magic, hparams, vocab, then per tensor: n_dims, name length, type, ne[], name, data) --
with the 2-D matrices quantized to Q4_0 / Q8_0 by the oracle's restatement of quantize_row_q4_0_reference / _q8_0_reference (pinned bit for
bit to the reference, tests/test_oracle_pin.py), the way examples/gpt-2/quantize.cpp leaves them.  BASELINE.json configs[2]: "GPT-2 117M
Q4_0 gpt-2-backend decode + 128-token prompt, random-init weights".

usage: make_gpt2_model.py OUT.bin [q4_0|q8_0] [n_layer] [seed]

The vocabulary is synthetic: ids 0..255 are the single bytes (so any prompt tokenizes, one token per character after the tokenizer's
own splitting), the rest are unique filler strings."""
import ctypes as C
import struct
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
N_VOCAB, N_CTX, N_EMBD, N_HEAD = 50257, 1024, 768, 12
GGML_FILE_MAGIC = 0x67676d6c
GGML_QNT_VERSION, GGML_QNT_VERSION_FACTOR = 2, 1000
FTYPE = {"q4_0": 2, "q8_0": 7}          # enum ggml_ftype (include/ggml/ggml.h:393-418)
TTYPE = {"f32": 0, "q4_0": 2, "q8_0": 8}
WIRE = {"q4_0": 18, "q8_0": 34}


def main():
    out = Path(sys.argv[1])
    q = sys.argv[2] if len(sys.argv) > 2 else "q4_0"
    n_layer = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    seed = int(sys.argv[4]) if len(sys.argv) > 4 else 1234
    lib = C.CDLL(str(ROOT / "_build" / "libqmm_oracle.so"))
    quant = lib.oracle_quantize_row_q4_0_reference if q == "q4_0" else lib.oracle_quantize_row_q8_0_reference
    rng = np.random.default_rng(seed)

    def tensor(f, name, arr, quantize):
        arr = np.ascontiguousarray(arr, np.float32)
        dims = arr.shape[::-1]                      # ggml order: ne[0] is the contiguous dimension
        nb = name.encode()
        f.write(struct.pack("<iii", len(dims), len(nb), TTYPE[q] if quantize else TTYPE["f32"]))
        for d in dims:
            f.write(struct.pack("<i", d))
        f.write(nb)
        if not quantize:
            f.write(arr.tobytes())
            return
        rows, k = arr.shape
        wire = np.zeros((rows, k // 32 * WIRE[q]), np.uint8)
        for r in range(rows):
            quant(arr[r].ctypes.data_as(C.c_void_p), wire[r].ctypes.data_as(C.c_void_p), C.c_int64(k))
        f.write(wire.tobytes())

    def u(shape, scale, offset=0.0):
        return (offset + scale * rng.uniform(-1, 1, shape)).astype(np.float32)

    with open(out, "wb") as f:
        f.write(struct.pack("<I", GGML_FILE_MAGIC))
        f.write(struct.pack("<iiiiii", N_VOCAB, N_CTX, N_EMBD, N_HEAD, n_layer, FTYPE[q] + GGML_QNT_VERSION * GGML_QNT_VERSION_FACTOR))
        f.write(struct.pack("<i", N_VOCAB))
        for i in range(N_VOCAB):
            w = bytes([i]) if i < 256 else b"<%d>" % i
            f.write(struct.pack("<I", len(w)) + w)
        tensor(f, "model/ln_f/g", u((N_EMBD,), 0.1, 1.0), False)
        tensor(f, "model/ln_f/b", u((N_EMBD,), 0.02), False)
        tensor(f, "model/wte", u((N_VOCAB, N_EMBD), 0.04), True)      # tied lm_head (main-backend.cpp:417-420)
        tensor(f, "model/wpe", u((N_CTX, N_EMBD), 0.02), False)
        for l in range(n_layer):
            p = f"model/h{l}/"
            tensor(f, p + "ln_1/g", u((N_EMBD,), 0.1, 1.0), False)
            tensor(f, p + "ln_1/b", u((N_EMBD,), 0.02), False)
            tensor(f, p + "ln_2/g", u((N_EMBD,), 0.1, 1.0), False)
            tensor(f, p + "ln_2/b", u((N_EMBD,), 0.02), False)
            tensor(f, p + "attn/c_attn/w", u((3 * N_EMBD, N_EMBD), 0.06), True)
            tensor(f, p + "attn/c_attn/b", u((3 * N_EMBD,), 0.02), False)
            tensor(f, p + "attn/c_proj/w", u((N_EMBD, N_EMBD), 0.04), True)
            tensor(f, p + "attn/c_proj/b", u((N_EMBD,), 0.02), False)
            tensor(f, p + "mlp/c_fc/w", u((4 * N_EMBD, N_EMBD), 0.06), True)
            tensor(f, p + "mlp/c_fc/b", u((4 * N_EMBD,), 0.02), False)
            tensor(f, p + "mlp/c_proj/w", u((N_EMBD, 4 * N_EMBD), 0.03), True)
            tensor(f, p + "mlp/c_proj/b", u((N_EMBD,), 0.02), False)
    print(f"{out}: GPT-2 {n_layer} layers, {q}, {out.stat().st_size / 1e6:.1f} MB")


if __name__ == "__main__":
    main()
