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


def cvt_pair(bits, bias, d):
    """cvt_pair of gemm_f16_fused_kernel: two small unsigned integers at bits 0.. and 16.. -> (lo, hi) fp16"""
    hb = np.uint32(bits | 0x64006400)
    h = np.array([hb & 0xFFFF, hb >> 16], dtype=np.uint16).view(np.float16)
    return ((h - np.float16(bias)).astype(np.float16) * np.float16(d)).astype(np.float16)


def test_fused_kernel_bit_manipulation_reproduces_the_row():
    """The masks, shifts and chunk numbering of the v2 kernel's dequant warps, statement by statement, against the plain
    definition of the formats (src/ggml-common.h:144-149, :186-191): one weight row, one k-step = two blocks = 64 fp16 in
    eight 16-byte chunks (the swizzle of the chunk position is a separate, hardware-defined matter)."""
    rng = np.random.default_rng(1)
    d = np.array([0.0371, -0.52], dtype=np.float16)
    qs = rng.integers(0, 256, (2, 16), dtype=np.uint8)
    expect = np.zeros(64, np.float16)
    for blk in range(2):
        for j in range(16):
            expect[blk * 32 + j] = np.float16(np.float32((int(qs[blk, j]) & 0xF) - 8) * np.float32(d[blk]))
            expect[blk * 32 + 16 + j] = np.float16(np.float32((int(qs[blk, j]) >> 4) - 8) * np.float32(d[blk]))
    got = np.zeros(64, np.float16)
    for blk in range(2):
        w = qs[blk].view(np.uint32)
        for hi in range(2):
            for half8 in range(2):
                o = []
                for t in range(2):
                    x = int(w[half8 * 2 + t]) >> (hi * 4)
                    o.append(cvt_pair((x & 0x0000000F) | ((x & 0x00000F00) << 8), 1032, d[blk]))
                    o.append(cvt_pair(((x >> 16) & 0x0000000F) | ((x >> 8) & 0x000F0000), 1032, d[blk]))
                c = blk * 4 + hi * 2 + half8
                got[c * 8:(c + 1) * 8] = np.concatenate(o)
    assert np.array_equal(got, expect)
    q8 = rng.integers(-128, 128, (2, 32), dtype=np.int8)
    expect = np.concatenate([(q8[b].astype(np.float32) * np.float32(d[b])).astype(np.float16) for b in range(2)])
    got = np.zeros(64, np.float16)
    for blk in range(2):
        for h16 in range(2):
            w = q8[blk, h16 * 16:(h16 + 1) * 16].view(np.uint32) ^ np.uint32(0x80808080)
            for half8 in range(2):
                o = []
                for t in range(2):
                    x = int(w[half8 * 2 + t])
                    o.append(cvt_pair((x & 0x000000FF) | ((x & 0x0000FF00) << 8), 1152, d[blk]))
                    o.append(cvt_pair(((x >> 16) & 0x000000FF) | ((x >> 8) & 0x00FF0000), 1152, d[blk]))
                c = blk * 4 + h16 * 2 + half8
                got[c * 8:(c + 1) * 8] = np.concatenate(o)
    assert np.array_equal(got, expect)
