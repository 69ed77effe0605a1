"""The arithmetic behind the experimental fp16 prefill GEMM (b200_gemm_tc.cu), checked with numpy's IEEE
half type: v1 dequantizes a weight as fp16_rn(float(quant) * float(d)); v2 puts the integer into the mantissa of 1024.0
(bits 0x6400 | u), subtracts the bias in fp16 (exact) and multiplies by d in fp16 (one rounding).  Both must give the same
fp16 value for every nibble / int8 and every scale, or the two kernels would not be interchangeable."""
import numpy as np


def scales():
    rng = np.random.default_rng(0)
    return np.concatenate([rng.uniform(1e-4, 2, 20000), rng.uniform(-2, -1e-4, 2000), [6e-5, 1e-7, 0.0, 100.0]]).astype(np.float16)


def magic(u, bias, d):
    h = np.array([0x6400 | u], dtype=np.uint16).view(np.float16)[0]
    assert float(h) == 1024 + u                                   # mantissa LSB of 1024.0 is 1
    a = (np.float16(h) - np.float16(bias)).astype(np.float16)
    return a, (a * d).astype(np.float16)


def test_q4_0_nibbles():
    d = scales()
    for nib in range(16):
        a, v2 = magic(nib, 1032, d)
        assert float(a) == nib - 8
        v1 = (np.float32(nib - 8) * d.astype(np.float32)).astype(np.float16)
        assert np.array_equal(v1, v2)


def test_q8_0_bytes():
    d = scales()
    for q in range(-128, 128):
        a, v2 = magic((q & 0xFF) ^ 0x80, 1152, d)                 # int8 + 128 as an unsigned byte
        assert float(a) == q
        v1 = (np.float32(q) * d.astype(np.float32)).astype(np.float16)
        assert np.array_equal(v1, v2)
