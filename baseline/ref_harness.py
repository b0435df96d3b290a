"""Drive the UNMODIFIED reference (GumBolt on EngineDiVAEpp; GumBoltCaloV5 generation) from a staged copy.

TEST / BENCHMARK INFRASTRUCTURE — never imported by the product package.  Used by
tests/test_gpu_reference_engine.py, benchmarks/train_step_bench.py and bench.py's `train` leg.

Two arms are built from the same reference model / engine source files:
  "reference"  models.rbm.rbm.RBM + models.samplers.pcd.PCD are the reference's own classes (ATen op chain)
  "dropin"     divae_b200.dropin.install() has registered this repo's RBM / PCD under those module paths BEFORE the
               reference's model modules are imported, so `from models.samplers.pcd import PCD` (gumbolt.py:12,
               dvaepp.py:16) resolves to the B200 sampler; nothing else differs.
The engine is driven the way scripts/run.py:100-150 does it (model -> device, Adam over model.parameters(),
engine.fit(epoch)), minus hydra / wandb / plotting (stubs: oracle/_ref_shim.py and the `wandb` stub below).
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import _ref_shim  # noqa: E402

_REF_PKGS = ("models", "engine", "utils", "data")


def reference_available():
    return _ref_shim.reference_available()


def _wandb_stub():
    """engineDiVAEpp.py:76-96 builds wandb.Image lists and calls wandb.log; neither touches the arithmetic."""
    m = types.ModuleType("wandb")
    m.logged = []
    m.init = lambda *a, **k: None

    def log(d, *a, **k):
        m.logged.append({key: (float(v.detach()) if torch.is_tensor(v) and v.numel() == 1 else None) for key, v in d.items()})

    m.log = log
    m.Image = lambda img, caption=None: ("img", getattr(img, "shape", None), caption)
    m.run = types.SimpleNamespace(dir="/tmp")
    m.watch = lambda *a, **k: None
    return m


def _purge():
    for k in [k for k in sys.modules if k.split(".")[0] in _REF_PKGS]:
        del sys.modules[k]


def load_arm(arm):
    """Fresh import of the reference's model and engine modules for one arm; returns a namespace of classes."""
    assert arm in ("reference", "dropin")
    _ref_shim.install()
    sys.modules["wandb"] = _wandb_stub()
    _purge()
    if arm == "dropin":
        from divae_b200 import dropin
        dropin.install()
    from models.autoencoders.gumbolt import GumBolt
    from models.autoencoders.gumboltCaloV5 import GumBoltCaloV5
    from engine.engineDiVAEpp import EngineDiVAEpp
    import models.rbm.rbm as rbm_mod
    import models.samplers.pcd as pcd_mod
    ns = types.SimpleNamespace(arm=arm, GumBolt=GumBolt, GumBoltCaloV5=GumBoltCaloV5, EngineDiVAEpp=EngineDiVAEpp,
                               RBM=rbm_mod.RBM, PCD=pcd_mod.PCD, wandb=sys.modules["wandb"])
    _purge()   # the classes stay alive; the next arm re-imports from scratch
    return ns


def gumbolt_cfg(n_latent_nodes=64, n_epochs=1):
    """configs/model/gumbolt.yaml keys (n_latent_nodes overridable: 64 -> RBM 64x64 = BASELINE configs[1] size)."""
    ns = types.SimpleNamespace
    return ns(model=ns(n_latent_hierarchy_lvls=2, n_latent_nodes=n_latent_nodes, n_encoder_layer_nodes=200, n_encoder_layers=2,
                       decoder_hidden_nodes=[200, 200], encoder_hidden_nodes=[200, 200], beta_smoothing_fct=5,
                       output_smoothing_fct=8, activation_fct="relu", model_type="GumBolt"),
              engine=ns(weight_decay_factor=1e-4, n_gibbs_sampling_steps=40, n_epochs=n_epochs, learning_rate=1e-3),
              sampler=None)


def calo_cfg(n_latent_nodes=64):
    """configs/model/gumboltcaloV.yaml (the reference's default model: GumBoltCaloV5, 504 voxels)."""
    ns = types.SimpleNamespace
    return ns(model=ns(n_latent_hierarchy_lvls=2, n_latent_nodes=n_latent_nodes, n_encoder_layer_nodes=64, n_encoder_layers=2,
                       decoder_hidden_nodes=[200, 300, 400], encoder_hidden_nodes=[400, 300, 200], beta_smoothing_fct=7,
                       activation_fct="relu", model_type="GumBoltCaloV5"),
              engine=ns(weight_decay_factor=1e-4, n_gibbs_sampling_steps=40, n_epochs=1, learning_rate=1e-3),
              sampler=None)


def build_calo(arm_ns, device, n_latent_nodes=64, seed=0):
    """GumBoltCaloV5 at the CaloGAN 3-layer shape (288 + 144 + 72 = 504 voxels; BASELINE configs[2])."""
    torch.manual_seed(seed)
    cfg = calo_cfg(n_latent_nodes)
    model = arm_ns.GumBoltCaloV5(flat_input_size=[504], train_ds_mean=torch.full((504,), 0.3), activation_fct=torch.nn.ReLU(), cfg=cfg)
    model.create_networks()
    return model.to(device), cfg


class SyntheticData:
    """data_mgr stand-in (data/dataManager.py): binarised MNIST-shape batches [B, 1, 28, 28] with labels."""

    def __init__(self, n_batches, batch_size, seed=7, p=0.3):
        g = torch.Generator().manual_seed(seed)
        # a few fixed prototype "digits" plus pixel noise, so the autoencoder has something to learn
        protos = (torch.rand(10, 784, generator=g) < p).float()
        labels = torch.randint(0, 10, (n_batches * batch_size,), generator=g)
        flip = (torch.rand(n_batches * batch_size, 784, generator=g) < 0.05).float()
        x = (protos[labels] + flip).remainder(2.0).reshape(-1, 1, 28, 28)
        ds = torch.utils.data.TensorDataset(x, labels)
        self.train_loader = torch.utils.data.DataLoader(ds, batch_size=batch_size, shuffle=False)
        self.test_loader = self.train_loader
        self.mean = x.reshape(-1, 784).mean(0)


def build_gumbolt(arm_ns, device, n_latent_nodes=64, seed=0, sampler_kwargs=None):
    """GumBolt as models/modelCreator.py builds it: ctor, create_networks(), .to(device)."""
    torch.manual_seed(seed)
    cfg = gumbolt_cfg(n_latent_nodes)
    model = arm_ns.GumBolt(flat_input_size=[784], train_ds_mean=torch.full((784,), 0.3), activation_fct=torch.nn.ReLU(), cfg=cfg)
    model.create_networks()
    if sampler_kwargs:
        # BASELINE shapes are set by constructing the sampler directly (dvaepp.py:57 hard-codes 128 x 40)
        model.sampler = arm_ns.PCD(RBM=model.prior, **sampler_kwargs)
    return model.to(device), cfg


def build_engine(arm_ns, model, cfg, data, device, lr=1e-3):
    eng = arm_ns.EngineDiVAEpp(cfg=cfg)
    # engineDiVAEpp.py:20 skips EngineBase.__init__ through super(Engine, self): set what scripts/run.py sets
    eng._config = cfg
    eng.model = model
    eng.optimiser = torch.optim.Adam(model.parameters(), lr=lr)     # scripts/run.py:121
    eng.data_mgr = data
    eng.device = device
    return eng
