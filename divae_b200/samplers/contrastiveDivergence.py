"""CD-k for a vanilla RBM (BASELINE config 1: 784x128, CD-1, batch 128).

The reference's models/samplers/contrastiveDivergence.py does not parse; its
intended semantics survive in sandbox/rbm_standalone.py:18-121 (`Contrastive_Divergence`),
which this class mirrors: same constructor hyper-parameters, same attribute names
(`_weights`, `_visible_bias`, `_hidden_bias`, `weights_update`, ...), same update rule
(momentum, L2 weight cost, un-normalised learning rate) and the same returned loss
(sum of squared reconstruction errors).  All matmul/sigmoid/draw/statistics work runs in
librbm_b200.so (fp32 GENERAL kernels: the visible layer carries probabilities, not bits).

Draws: the sandbox draws ONE uniform vector of length n_hidden per draw and broadcasts it
over the batch (rbm_standalone.py:72,84 — SURVEY.md §9).  By default this class draws an
independent Philox uniform per unit and chain (statistically correct); pass
`uniforms=[k+1, n_hidden]` to `contrastive_divergence` to reproduce the sandbox bit for bit.
"""
import torch

from .. import _lib, _ops


class _Params:
    """Duck-typed RBM (weights / visible_bias / hidden_bias) for ParamPack."""

    def __init__(self, owner):
        self._o = owner

    @property
    def weights(self):
        return self._o._weights

    @property
    def visible_bias(self):
        return self._o._visible_bias

    @property
    def hidden_bias(self):
        return self._o._hidden_bias


class ContrastiveDivergence(object):
    def __init__(self, n_visible, n_hidden, learning_rate, momentum_coefficient, n_gibbs_sampling_steps,
                 weight_cost, device="cuda", seed=None, **kwargs):
        # rbm_standalone.py:18-44
        self.learning_rate = learning_rate
        self.momentum_coefficient = momentum_coefficient
        self.n_gibbs_sampling_steps = n_gibbs_sampling_steps
        self.weight_cost = weight_cost
        self.n_visible = n_visible
        self.n_hidden = n_hidden
        dev = torch.device(device)
        self.weights_update = torch.zeros(n_visible, n_hidden, device=dev)
        self.visible_bias_update = torch.zeros(n_visible, device=dev)
        self.hidden_bias_update = torch.zeros(n_hidden, device=dev)
        self._weights = (torch.randn(n_visible, n_hidden) * 0.01).to(dev)
        self._visible_bias = (torch.ones(n_visible) * 0.5).to(dev)
        self._hidden_bias = torch.zeros(n_hidden, device=dev)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        self._seed = int(seed)
        self._draws = 0
        self._pack = _ops.ParamPack()
        self._view = _Params(self)
        self._ws = None

    def _packed(self):
        # tensors are updated in place, so bump the cache key by hand
        self._pack._key = None
        self._pack._packed_key = None
        return self._pack.pack(self._view, need_bf16=False)

    def sample_from_hidden(self, probabilities_visible):
        """sigmoid(p_v W + c) (rbm_standalone.py:56-59)."""
        return _ops.half_step(self._packed(), 0, probabilities_visible, _lib.MODE_PROBS)

    def sample_from_visible(self, probabilities_hidden):
        """sigmoid(p_h W^T + a) (rbm_standalone.py:61-64)."""
        return _ops.half_step(self._packed(), 1, probabilities_hidden, _lib.MODE_PROBS)

    def _draw_hidden(self, visible, uniforms_row):
        """(sigmoid(v W + c) >= u).float(): one fused kernel, Philox or injected uniforms."""
        if uniforms_row is not None:
            return _ops.half_step(self._packed(), 0, visible, _lib.MODE_INJECTED, uniforms=uniforms_row.reshape(1, -1))
        t = self._draws
        self._draws += 1
        return _ops.half_step(self._packed(), 0, visible, _lib.MODE_SAMPLE, seed=self._seed, half_step_idx=t)

    def gibbs_sampling(self, output_hidden, uniforms=None):
        """k x { p_v = sigma(h W^T + a); p_h = sigma(p_v W + c); h = (p_h >= u) } (rbm_standalone.py:66-73)."""
        for step in range(self.n_gibbs_sampling_steps):
            probabilities_visible = self.sample_from_visible(output_hidden)
            probabilities_hidden = self.sample_from_hidden(probabilities_visible)
            output_hidden = self._draw_hidden(probabilities_visible, None if uniforms is None else uniforms[step])
        return probabilities_visible, probabilities_hidden

    def contrastive_divergence(self, input_data, uniforms=None):
        """One CD-k update (rbm_standalone.py:75-121) as ONE library call (rbm_b200_cd_step).
        Returns the summed squared reconstruction error (a 0-d tensor on the device)."""
        _lib.require_cuda(input_data, "input data")
        lib = _lib.load()
        v0 = input_data.float().contiguous()
        B = v0.shape[0]
        if v0.dim() != 2 or v0.shape[1] != self.n_visible:
            raise RuntimeError("size mismatch: data %s for %d visible units" % (tuple(v0.shape), self.n_visible))
        k = int(self.n_gibbs_sampling_steps)
        u = None
        if uniforms is not None:
            u = _ops._f32c(uniforms, "uniforms")
            if u.shape != (k + 1, self.n_hidden):
                raise RuntimeError("uniforms must be [k+1, n_hidden] = [%d, %d]" % (k + 1, self.n_hidden))
        nbytes = lib.rbm_b200_cd_workspace(self.n_visible, self.n_hidden, B)
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != v0.device:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=v0.device)
        loss = torch.empty((), dtype=torch.float32, device=v0.device)
        with torch.no_grad(), _ops._on_device(v0.device):
            _lib.check(lib.rbm_b200_cd_step(
                _ops._ptr(self._weights), _ops._ptr(self._visible_bias), _ops._ptr(self._hidden_bias),
                _ops._ptr(self.weights_update), _ops._ptr(self.visible_bias_update), _ops._ptr(self.hidden_bias_update),
                self.n_visible, self.n_hidden, _ops._ptr(v0), B, k, float(self.learning_rate),
                float(self.momentum_coefficient), float(self.weight_cost), _ops._ptr(u), self._seed, self._draws,
                _ops._ptr(loss), _ops._ptr(self._ws), nbytes, _ops._stream()), "cd_step")
        self._draws += k
        return loss
