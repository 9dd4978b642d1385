"""CPU emulation of the index mapping of the EXPERIMENTAL `gemv_nf4_mma_kernel` (csrc/gemv.cu):
lane (g = lane/4, c = lane%4) owns columns 16c..16c+15 of every 64-column step for rows g and g+8
of two 16-row groups; bytes 2t, 2t+1 of its 8-byte chunk feed the A fragment of the t-th
mma.sync.m16n8k16 (PTX fragment layout: a0a1 = (row g, k 2c..2c+1), a2a3 = (row g+8, same k),
a4a5 / a6a7 = k + 8), the matching x pairs feed the B fragment, identical for all 8 columns; the
8 warps of a CTA interleave the steps.  The emulation must reproduce the plain dot product, i.e.
every (row, column) is used exactly once with its own x and its own block scale."""
import struct

import numpy as np


def test_fragment_mapping_reproduces_the_dot_product():
    rng = np.random.default_rng(0)
    m, k = 32, 256
    nib = rng.integers(0, 16, size=(m, k))
    code = np.linspace(-1, 1, 16)
    packed = ((nib[:, 0::2] << 4) | nib[:, 1::2]).astype(np.uint8)      # high nibble = even column
    absmax = rng.random((m, k // 64)) + 0.5
    x = rng.standard_normal(k)
    ref = np.einsum("rc,rc,c->r", code[nib], np.repeat(absmax, 64, axis=1), x)

    def lut(b):
        return code[b >> 4], code[b & 15]                                # (low half, high half)
    out = np.zeros(m)
    for warp in range(8):
        acc = np.zeros((32, 2, 2))
        for step in range(warp, k // 64, 8):
            col0 = step * 64
            for rg in range(2):
                D = np.zeros((16, 8))
                for t in range(4):
                    A = np.zeros((16, 16))
                    B = np.zeros((16, 8))
                    for lane in range(32):
                        g, c = lane >> 2, lane & 3
                        ch = [packed[r, col0 // 2 + c * 8: col0 // 2 + c * 8 + 8] for r in (rg * 16 + g, rg * 16 + g + 8)]
                        xs = x[col0 + c * 16: col0 + c * 16 + 16]
                        A[g, 2 * c], A[g, 2 * c + 1] = lut(ch[0][2 * t])
                        A[g + 8, 2 * c], A[g + 8, 2 * c + 1] = lut(ch[1][2 * t])
                        A[g, 2 * c + 8], A[g, 2 * c + 9] = lut(ch[0][2 * t + 1])
                        A[g + 8, 2 * c + 8], A[g + 8, 2 * c + 9] = lut(ch[1][2 * t + 1])
                        B[2 * c, g], B[2 * c + 1, g] = xs[4 * t], xs[4 * t + 1]
                        B[2 * c + 8, g], B[2 * c + 9, g] = xs[4 * t + 2], xs[4 * t + 3]
                    D += A @ B
                for lane in range(32):
                    g, c = lane >> 2, lane & 3
                    acc[lane, rg, 0] += absmax[rg * 16 + g, step] * D[g, 2 * c]
                    acc[lane, rg, 1] += absmax[rg * 16 + g + 8, step] * D[g + 8, 2 * c]
        for lane in range(0, 32, 4):                                      # c == 0 lanes publish
            for rg in range(2):
                for h in range(2):
                    out[rg * 16 + (lane >> 2) + 8 * h] += acc[lane, rg, h]
    assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max()


def test_table_offset_arithmetic():
    """offset of byte p of a 32-bit word in the per-lane table = byte * 128: one shift + one mask."""
    rng = np.random.default_rng(1)
    for _ in range(200):
        ch = bytes(int(v) for v in rng.integers(0, 256, size=8))
        w = struct.unpack("<II", ch)
        for t in range(4):
            u, sh0 = w[t >> 1], 16 * (t & 1) - 7
            o0 = ((u << 7) if sh0 < 0 else (u >> sh0)) & 0x7F80
            o1 = (u >> (sh0 + 8)) & 0x7F80
            assert o0 >> 7 == ch[2 * t] and o1 >> 7 == ch[2 * t + 1]
