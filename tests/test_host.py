"""Host-side mirror of the reference API: names, defaults, state_dict keys, config
instantiation, drop-in module paths, error behaviour, chain sharding and the two collectives
(gloo, world_size 2)."""
import importlib
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import divae_b200
from divae_b200 import dist as bdist


def test_rbm_matches_reference_construction():
    torch.manual_seed(32)
    r = divae_b200.RBM(n_visible=64, n_hidden=48)
    assert [n for n, _ in r.named_parameters()] == ["_weights", "_visible_bias", "_hidden_bias"]
    assert r.weights.shape == (64, 48) and r.weights.dtype == torch.float32 and r.weights.requires_grad
    assert torch.all(r.visible_bias == 0.5) and torch.all(r.hidden_bias == 0)
    assert abs(r.weights.std().item() - 0.01) < 2e-3
    assert r.get_logZ_value() == 33.65                      # models/rbm/rbm.py:54
    assert repr(r) == "RBM: n_vis=64, n_hid=48"
    assert r.get_weights() is r.weights
    # same RNG consumption as the reference: randn(nV, nH) is the first draw
    torch.manual_seed(32)
    assert torch.equal(r.weights.detach(), torch.randn(64, 48) * 0.01)


def test_sampler_state_dict_keys_match_reference():
    r = divae_b200.RBM(8, 8)
    s = divae_b200.PCD(batch_size=4, RBM=r, n_gibbs_sampling_steps=3)
    assert list(s.state_dict()) == ["_RBM._weights", "_RBM._visible_bias", "_RBM._hidden_bias"]
    assert list(s.named_buffers()) == []                     # _MCState is not a buffer (pcd.py:16)
    assert s.get_batch_size() == 4 and s.n_gibbs_sampling_steps == 3
    g = divae_b200.GibbsSampler(RBM=r, n_gibbs_sampling_steps=40)
    assert list(g.state_dict()) == ["_RBM._weights", "_RBM._visible_bias", "_RBM._hidden_bias"]
    assert divae_b200.BaseSampler(5).get_samples() is NotImplementedError  # baseSampler.py:25 returns the class
    with pytest.raises(NotImplementedError):
        divae_b200.BaseSampler(5).hidden_samples()


def _instantiate(cfg, **kwargs):
    cfg = dict(cfg)
    modname, clsname = cfg.pop("_target_").rsplit(".", 1)
    return getattr(importlib.import_module(modname), clsname)(**cfg, **kwargs)


def test_sampler_yaml_twins_instantiate():
    import yaml
    cfgdir = os.path.join(os.path.dirname(divae_b200.__file__), "configs", "sampler")
    r = divae_b200.RBM(16, 16)
    pcd = _instantiate(yaml.safe_load(open(os.path.join(cfgdir, "pcdSampler.yaml"))), RBM=r)
    assert isinstance(pcd, divae_b200.PCD) and pcd.get_batch_size() == 32 and pcd.n_gibbs_sampling_steps == 40
    gs = _instantiate(yaml.safe_load(open(os.path.join(cfgdir, "gibbsSampler.yaml"))), RBM=r)
    assert isinstance(gs, divae_b200.GibbsSampler) and gs.n_gibbs_sampling_steps == 40


def test_dropin_registers_reference_module_paths():
    from divae_b200 import dropin
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")}
    try:
        dropin.install()
        from models.rbm.rbm import RBM
        from models.samplers.pcd import PCD
        from models.samplers.gibbsSampler import GibbsSampler
        from models.samplers.baseSampler import BaseSampler
        assert RBM is divae_b200.RBM and PCD is divae_b200.PCD
        assert GibbsSampler is divae_b200.GibbsSampler and BaseSampler is divae_b200.BaseSampler
        # the reference's own construction calls (dvaepp.py:57, discreteVAE.py:122)
        prior = RBM(n_visible=32, n_hidden=32)
        PCD(batch_size=128, RBM=prior, n_gibbs_sampling_steps=40)
    finally:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_cpu_tensors_are_rejected_loudly():
    r = divae_b200.RBM(8, 8)
    s = divae_b200.PCD(batch_size=4, RBM=r, n_gibbs_sampling_steps=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.block_gibbs_sampling()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.hidden_samples(torch.zeros(4, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        r.energy([torch.zeros(4, 8), torch.zeros(4, 8)])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        divae_b200.GibbsSampler(RBM=r, n_gibbs_sampling_steps=2).gibbs_sampling(torch.zeros(4, 8), 1)


def test_seed_follows_torch_global_generator():
    r = divae_b200.RBM(8, 8)
    torch.manual_seed(5); a = divae_b200.PCD(4, r, n_gibbs_sampling_steps=1)._seed
    torch.manual_seed(5); b = divae_b200.PCD(4, r, n_gibbs_sampling_steps=1)._seed
    torch.manual_seed(6); c = divae_b200.PCD(4, r, n_gibbs_sampling_steps=1)._seed
    assert a == b and a != c
    assert divae_b200.PCD(4, r, n_gibbs_sampling_steps=1, seed=77)._seed == 77


def test_shard_chains_partitions_exactly():
    for total, world in ((65536, 8), (100, 8), (7, 2), (3, 4), (0, 2)):
        spans = [bdist.shard_chains(total, r, world) for r in range(world)]
        assert sum(n for _, n in spans) == total
        pos = 0
        for off, n in spans:
            assert off == pos
            pos += n
    with pytest.raises(ValueError):
        bdist.shard_chains(10, 2, 2)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        v = (torch.rand(40, 6) < 0.5).float()
        h = (torch.rand(40, 5) < 0.5).float()
        off, n = bdist.shard_chains(40, rank, world)
        vl, hl = v[off:off + n], h[off:off + n]
        # local statistics pre-scaled by 1/B_global, then ONE flat sum-allreduce
        S, sv, sh = bdist.allreduce_stats(vl.t() @ hl / 40, vl.sum(0) / 40, hl.sum(0) / 40)
        ok = torch.allclose(S, v.t() @ h / 40, atol=1e-6) and torch.allclose(sv, v.mean(0), atol=1e-6) \
            and torch.allclose(sh, h.mean(0), atol=1e-6)
        logw = torch.randn(40, dtype=torch.float64) * 5
        lse = bdist.logsumexp_allreduce(logw[off:off + n])
        ok = ok and abs(lse.item() - torch.logsumexp(logw, 0).item()) < 1e-9
        # negative_phase() under a process group: the ranks' batches (sizes may differ) are laid out one after
        # the other, so every rank draws its own chains and the statistics are scaled by the total count
        s = divae_b200.PCD(5 + 2 * rank, divae_b200.RBM(6, 5), n_gibbs_sampling_steps=1, seed=3)
        ok = ok and s._group_layout(None) == ((0, 12) if rank == 0 else (5, 12)) and not s._chain_offset_given
        ok = ok and divae_b200.PCD(5, divae_b200.RBM(6, 5), n_gibbs_sampling_steps=1, seed=3, chain_offset=9)._chain_offset_given
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_collectives_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_logsumexp_single_process():
    x = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)
    assert abs(bdist.logsumexp_allreduce(x).item() - torch.logsumexp(x, 0).item()) < 1e-12


def test_modules_survive_pickle_and_deepcopy():
    """torch.save(model) / copy.deepcopy(model) of reference-style models must keep working: the ctypes
    parameter cache is transient state."""
    import copy
    import io
    import pickle
    r = divae_b200.RBM(8, 6)
    s = divae_b200.PCD(batch_size=4, RBM=r, n_gibbs_sampling_steps=3, seed=5)
    g = divae_b200.GibbsSampler(RBM=r, n_gibbs_sampling_steps=2)
    r._pack._packed = object()           # pretend a call has populated the cache
    r2, s2, g2 = copy.deepcopy(r), copy.deepcopy(s), copy.deepcopy(g)
    assert torch.equal(r2.weights, r.weights) and s2._RBM is not s._RBM and s2._seed == 5
    buf = io.BytesIO()
    torch.save(s, buf)
    buf.seek(0)
    s3 = torch.load(buf, weights_only=False)
    assert s3.get_batch_size() == 4 and torch.equal(s3._RBM.weights, r.weights)
    assert pickle.loads(pickle.dumps(g)).n_gibbs_sampling_steps == 2
    s3.load_state_dict(s.state_dict())


def test_decoder_and_encoder_runners_validate_on_the_host():
    """DecoderRunner / EncoderRunner refuse what the B200 path does not implement, and CPU tensors (no CPU fallback)."""
    import torch
    from divae_b200.decoder import DecoderRunner, EncoderRunner

    class Dec(torch.nn.Module):
        def __init__(self, act):
            super().__init__()
            self._activation_fct = act
            self._layers = torch.nn.ModuleList([torch.nn.Linear(4, 6), torch.nn.Linear(6, 5)])

    with pytest.raises(RuntimeError, match="ReLU"):
        DecoderRunner(Dec(torch.nn.Tanh()))
    r = DecoderRunner(Dec(torch.nn.ReLU()))
    assert (r.in_features, r.out_features, r.two_heads) == (4, 5, False)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        r(torch.zeros(3, 4))
    with pytest.raises(RuntimeError, match="two-headed"):
        r.shower(torch.zeros(3, 4), seed=1)
    bad = Dec(torch.nn.ReLU()); bad._layers = torch.nn.ModuleList([torch.nn.Linear(4, 6, bias=False)])
    with pytest.raises(RuntimeError, match="nn.Linear layers with bias"):
        DecoderRunner(bad)

    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self._networks = torch.nn.ModuleList([torch.nn.Sequential(torch.nn.Linear(4, 6), torch.nn.ReLU(), torch.nn.Linear(6, 3), torch.nn.Identity())])

    with pytest.raises(RuntimeError, match="beta"):
        EncoderRunner(Enc())
    e = EncoderRunner(Enc(), beta=5.0)
    assert e.n_latent == 3 and e.beta == 5.0
    with pytest.raises(RuntimeError, match="CUDA-only"):
        e(torch.zeros(2, 4))
    tanh = Enc(); tanh._networks[0][1] = torch.nn.Tanh()
    with pytest.raises(RuntimeError, match="Linear/ReLU"):
        EncoderRunner(tanh, beta=5.0)
