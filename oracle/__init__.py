"""CPU oracle for the RBM block-Gibbs hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and there only as the checker or as the
timed CPU baseline.  The product path (``divae_b200``) never imports it and
fails loudly if its CUDA library is missing.

Parity pin: the reference (ezeeEric/DiVAE) has no tests or golden vectors
for this path, so the oracle is pinned by (i) outputs of the
reference's own code executed in the build container with interposed
``torch.rand`` — committed as ``tests/golden/*.npz`` together with the
generating script ``oracle/gen_golden.py`` — (ii) exact enumeration of small
RBMs, and (iii) the published Random123 known-answer vectors for
Philox4x32-10.
"""
