"""Philox4x32-10 counter-based RNG and the uniform-draw contract (numpy).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates the published
algorithm of Salmon et al., "Parallel random numbers: as easy as 1, 2, 3"
(SC'11), Random123 v1.09 ``philox.h`` — a third-party algorithm that is not
vendored in the reference (the reference draws with ``torch.rand``,
models/samplers/pcd.py:35,49).  Known-answer vectors: Random123
``kat_vectors`` (checked in tests/test_philox.py).

Draw contract shared by every CUDA kernel in divae_b200/csrc (philox.cuh):

  key      = (seed & 0xffffffff, seed >> 32)
  counter  = (unit_block, chain_id, half_step, tag)
  unit -> (unit_block, slot):  with c the unit (column) index,
      unit_block = (c // 32) * 4 + (c % 8) // 2
      slot       = 2 * ((c % 32) // 8) + (c % 2)
  The 128 output bits are eight 16-bit fields: slot s lives in output word
  s // 2, low half if s is even, high half if odd.
  u16 = that field drawn with tag TAG_GIBBS (or TAG_INIT ...); e16 = the same
  field drawn with tag | TAG_EXTRA_BIT, consulted only when u16 alone cannot
  decide.  The uniform is u = (u16 * 65536 + e16) / 2**32 in [0, 1) and a unit
  fires iff p >= u (the reference's ``probs >= torch.rand``).
"""
import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

TAG_GIBBS = 0      # Bernoulli draws of Gibbs half-steps
TAG_INIT = 1       # chain (re-)initialisation, p = 1/2 (pcd.py:16-17)
TAG_AIS = 2        # AIS transition draws
TAG_AIS_INIT = 3   # AIS base-rate initial state
TAG_HITS = 4       # hit gate of the shower decoder (oracle/decoder_oracle.py)
TAG_EXTRA_BIT = 0x100  # OR-ed into the tag for the 16 refinement bits


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. All inputs broadcastable uint32-valued arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3)])
    c0 = c0 & MASK32; c1 = c1 & MASK32; c2 = c2 & MASK32; c3 = c3 & MASK32
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0 = p0 >> np.uint64(32); lo0 = p0 & MASK32
        hi1 = p1 >> np.uint64(32); lo1 = p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n1 = lo1
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def unit_block_slot(n_units):
    """unit index -> (unit_block, slot) per the contract above."""
    c = np.arange(n_units, dtype=np.int64)
    blk = (c // 32) * 4 + (c % 8) // 2
    slot = 2 * ((c % 32) // 8) + (c % 2)
    return blk, slot


def draw_u16(n_units, chain_ids, half_step, tag, seed):
    """16-bit fields for every (chain, unit): uint32 array [len(chain_ids), n_units]."""
    blk, slot = unit_block_slot(n_units)
    chain_ids = np.asarray(chain_ids, dtype=np.uint64)
    k0 = seed & 0xFFFFFFFF
    k1 = (seed >> 32) & 0xFFFFFFFF
    # evaluate one Philox call per distinct (chain, unit_block)
    nblk = int(blk.max()) + 1 if n_units > 0 else 0
    b = np.arange(nblk, dtype=np.uint64)
    w = philox4x32_10(b[None, :], chain_ids[:, None], np.uint64(half_step), np.uint64(tag), k0, k1)
    words = np.stack(w, axis=-1)  # [C, nblk, 4]
    word = words[:, blk, slot // 2]  # [C, n_units]
    shift = ((slot % 2) * 16).astype(np.uint32)
    return (word >> shift[None, :]) & np.uint32(0xFFFF)


def draw_uniform(n_units, chain_ids, half_step, tag, seed):
    """Full-resolution uniforms u = (u16*65536 + e16)/2**32 as float64 (exact)."""
    u16 = draw_u16(n_units, chain_ids, half_step, tag, seed).astype(np.float64)
    e16 = draw_u16(n_units, chain_ids, half_step, tag | TAG_EXTRA_BIT, seed).astype(np.float64)
    return (u16 * 65536.0 + e16) / 4294967296.0
