"""Philox4x32-10 known-answer vectors (Random123 kat_vectors) and the draw contract."""
import numpy as np

from oracle import philox

KATS = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_random123_known_answers():
    for ctr, key, expect in KATS:
        out = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in out) == expect


def test_unit_mapping_is_a_bijection_per_32_units():
    blk, slot = philox.unit_block_slot(96)
    pairs = set(zip(blk.tolist(), slot.tolist()))
    assert len(pairs) == 96
    assert blk.max() == 11 and slot.max() == 7
    # the eight slots of one block are what one MMA thread owns: cols {2t, 2t+1} + 8j
    cols = np.arange(32)[blk[:32] == 1]
    assert cols.tolist() == [2, 3, 10, 11, 18, 19, 26, 27]


def test_uniforms_in_unit_interval_and_uniform():
    u = philox.draw_uniform(64, np.arange(512), half_step=3, tag=philox.TAG_GIBBS, seed=0x5EED)
    assert u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 4 * (1 / 12 ** 0.5) / (u.size ** 0.5)
    # different half-steps / chains / tags give different streams
    u2 = philox.draw_uniform(64, np.arange(512), half_step=4, tag=philox.TAG_GIBBS, seed=0x5EED)
    assert (u != u2).mean() > 0.99


def test_chain_offset_equals_slicing():
    a = philox.draw_u16(40, np.arange(100, 140), 7, philox.TAG_GIBBS, 123)
    b = philox.draw_u16(40, np.arange(0, 200), 7, philox.TAG_GIBBS, 123)[100:140]
    assert (a == b).all()
