"""Build-container only: re-run the real reference and compare with the committed fixtures.
Skipped wherever /root/reference is absent (e.g. the GPU box)."""
import numpy as np
import pytest

from oracle import _ref_shim
from helpers import unpack

pytestmark = pytest.mark.skipif(not _ref_shim.reference_available(), reason="reference tree not present")


def test_golden_pcd_regenerates_identically(golden, tmp_path, monkeypatch):
    from oracle import gen_golden
    _ref_shim.install()
    monkeypatch.setattr(gen_golden, "OUT", str(tmp_path))
    gen_golden.gen_pcd("pcd_64x64", 64, 64, 100, 10, 0.3, 0x5EED)
    new = dict(np.load(tmp_path / "pcd_64x64.npz"))
    old = golden("pcd_64x64")
    for key in ("W", "a", "c", "v", "h", "v0", "h1", "v1"):
        assert (new[key] == old[key]).all(), key


def test_torch_port_is_bit_identical_to_reference():
    import torch
    from oracle import gen_golden, ref_torch_port as port
    _ref_shim.install()
    from models.samplers.pcd import PCD
    W, a, c = gen_golden.make_params(32, 48, 0.3, 5)
    rbm = gen_golden.ref_rbm(W, a, c)
    torch.manual_seed(1)
    s = PCD(batch_size=17, RBM=rbm, n_gibbs_sampling_steps=4)
    with torch.no_grad():
        v, h = s.block_gibbs_sampling()
    torch.manual_seed(1)
    st = port.pcd_init_state(17, 32)
    with torch.no_grad():
        pv, ph, _ = port.block_gibbs_sampling(st, torch.from_numpy(W), torch.from_numpy(a), torch.from_numpy(c), 4)
    assert torch.equal(v, pv) and torch.equal(h, ph)
