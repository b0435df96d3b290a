"""Host-side restatement of two integer tricks of the resident kernel (divae_b200/csrc/umma_resident.cuh), checked on
random words: the bit-word -> bf16 expansion of the expander warps (a set unit becomes 2.0 = 0x4000, one shift + AND per
unit-byte and one PRMT per bf16 pair) and the exactness of the halved epilogue scale that goes with it."""
import numpy as np


def prmt(a, b, sel):
    """PTX prmt.b32 in generic mode: selector nibble i picks byte (nibble & 7) of {a, b}; bit 3 replicates its sign bit."""
    pool = [(a >> (8 * i)) & 0xFF for i in range(4)] + [(b >> (8 * i)) & 0xFF for i in range(4)]
    out = 0
    for i in range(4):
        nib = (sel >> (4 * i)) & 0xF
        byte = pool[nib & 7]
        if nib & 8:
            byte = 0xFF if byte & 0x80 else 0x00
        out |= byte << (8 * i)
    return out


def expand_word(word):
    """umma_resident.cuh, expander loop: 32 units of one word -> 16 registers, register c*4+pp = units 8c+2pp, 8c+2pp+1."""
    m = [((word << (6 - i)) & 0xFFFFFFFF) & 0x40404040 for i in range(7)] + [(word >> 1) & 0x40404040]
    o = [0] * 16
    for c in range(4):
        for pp in range(4):
            sel = (8 | c) | (c << 4) | ((12 | c) << 8) | ((4 | c) << 12)
            o[c * 4 + pp] = prmt(m[2 * pp], m[2 * pp + 1], sel)
    return o


def test_expansion_writes_bf16_two_for_every_set_unit():
    rng = np.random.default_rng(1)
    words = [0, 0xFFFFFFFF, 1, 0x80000000, 0x55555555, 0xAAAAAAAA] + [int(x) for x in rng.integers(0, 2 ** 32, 500, dtype=np.uint64)]
    for w in words:
        regs = expand_word(w)
        for col, r in enumerate(regs):
            lo, hi = r & 0xFFFF, r >> 16
            u0, u1 = 2 * col, 2 * col + 1                     # the MMA's K order: column col of the K block = units 2col, 2col+1
            assert lo == (0x4000 if (w >> u0) & 1 else 0), (hex(w), col)
            assert hi == (0x4000 if (w >> u1) & 1 else 0), (hex(w), col)


def test_bf16_0x4000_is_two_and_the_halved_scale_is_exact():
    # bf16 0x4000 = fp32 0x40000000 = 2.0
    assert np.array([0x40000000], dtype=np.uint32).view(np.float32)[0] == 2.0
    # arg = fma(acc, scale, nb) with acc doubled and scale halved is the same real product: halving an fp32 scale and
    # doubling an fp32 accumulator are exact (no subnormals here), so the fused result rounds identically
    rng = np.random.default_rng(2)
    acc = (rng.standard_normal(20000) * 30).astype(np.float32)
    nb = rng.standard_normal(20000).astype(np.float32)
    for beta in (1.0, 0.37, 1e-4):
        scale = np.float32(-1.4426950408889634) * np.float32(beta)
        half = np.float32(0.5) * scale
        assert np.float64(half) * 2 == np.float64(scale)
        ref = (acc.astype(np.float64) * np.float64(scale) + nb.astype(np.float64)).astype(np.float32)
        new = ((2 * acc).astype(np.float64) * np.float64(half) + nb.astype(np.float64)).astype(np.float32)
        assert np.array_equal(ref, new)
