"""Thin torch-side wrappers over the C ABI: raw pointers + current stream, autograd glue.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); every
forward computation on the path is a call into librbm_b200.so.
"""
import ctypes

import torch

from . import _lib


# raw handle of the current stream without building a torch.cuda.Stream object (5 us -> 0.3 us per call)
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _f32c(t, name):
    _lib.require_cuda(t, name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class ParamPack:
    """Device views of (W, a, c) as the C struct the library takes.

    Only the fp32 parameters are handed over: the bf16 kernels round W inside the library on every call
    (rbm_b200_params.W_bf16 = NULL), so in-place writes that do not bump `_version` (`p.data.copy_()`, legacy
    optimisers) can never leave a stale low-precision copy behind.  The packed struct itself only holds pointers
    and shapes; it is cached while the parameter tensors are the same objects at the same addresses."""

    def __init__(self):
        self._packed = None
        self._packed_key = None

    # the cache holds ctypes structs and device pointers: never pickled / deep-copied, rebuilt on first use
    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.__init__()

    def __deepcopy__(self, memo):
        return ParamPack()

    def pack(self, rbm, need_bf16=False):
        # nn.Module attribute access goes through __getattr__ (slow): read the parameter dict directly when there is one
        params = rbm.__dict__.get("_parameters") if hasattr(rbm, "__dict__") else None
        if params is not None and "_weights" in params:
            Wp, ap, cp = params["_weights"], params["_visible_bias"], params["_hidden_bias"]
        else:
            Wp, ap, cp = rbm.weights, rbm.visible_bias, rbm.hidden_bias
        pkey = (id(Wp), Wp.data_ptr(), id(ap), ap.data_ptr(), id(cp), cp.data_ptr(), Wp.dtype, Wp.is_contiguous())
        if pkey == self._packed_key:
            return self._packed
        W = _f32c(Wp.detach(), "RBM weights")
        a = _f32c(ap.detach(), "RBM visible bias")
        c = _f32c(cp.detach(), "RBM hidden bias")
        p = _lib.RbmParams(W.shape[0], W.shape[1], W.data_ptr(), 0, a.data_ptr(), c.data_ptr())
        self._packed = (p, (W, a, c))  # keep tensors alive for the duration of the call
        # non-fp32 / non-contiguous parameters are converted copies: never cache those
        direct = W.data_ptr() == Wp.data_ptr() and a.data_ptr() == ap.data_ptr() and c.data_ptr() == cp.data_ptr()
        self._packed_key = pkey if direct else None
        return self._packed


class _on_device:
    """torch.cuda.device(dev) only when dev is not already current (the context manager costs microseconds)."""

    def __init__(self, dev):
        self._ctx = None if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)

    def __enter__(self):
        if self._ctx is not None:
            self._ctx.__enter__()

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)


def half_step(rbm_pack, direction, x, mode, uniforms=None, seed=0, chain_offset=0, half_step_idx=0):
    """out = f(x W + c) (direction 0) or f(x W^T + a) (direction 1); see rbm_b200_half_step."""
    lib = _lib.load()
    p, keep = rbm_pack
    x = _f32c(x, "input state")
    squeeze = x.dim() == 1
    if squeeze:
        x = x.unsqueeze(0)
    n_in = p.n_visible if direction == 0 else p.n_hidden
    n_out = p.n_hidden if direction == 0 else p.n_visible
    if x.dim() != 2 or x.shape[1] != n_in:
        raise RuntimeError("size mismatch: got input %s, RBM side has %d units" % (tuple(x.shape), n_in))
    out = torch.empty((x.shape[0], n_out), dtype=torch.float32, device=x.device)
    u = None
    if mode == _lib.MODE_INJECTED:
        u = _f32c(uniforms, "uniforms").expand(x.shape[0], n_out).contiguous()
    with torch.cuda.device(x.device):
        _lib.check(lib.rbm_b200_half_step(ctypes.byref(p), direction, _ptr(x), _ptr(out), x.shape[0], mode,
                                          _ptr(u), seed, chain_offset, half_step_idx, _stream()), "half_step")
    return out.squeeze(0) if squeeze else out


_gemm_ws = {}   # (device index, stream) -> grow-only uint8 workspace shared by the GEMM wrappers below


def _workspace(dev, nbytes):
    key = (dev.index, _stream().value)
    ws = _gemm_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = _gemm_ws[key] = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=dev)
    return ws


def gemm(a, b, trans_b, bias=None, relu=False):
    """act(a @ b.T + bias) (trans_b=True, b is [N, K]) or act(a @ b + bias) (trans_b=False, b is [K, N]) for fp32
    row-major operands: ONE tcgen05 GEMM on bf16 high + low splits of both operands (rbm_b200_gemm; ~2^-16 relative)."""
    lib = _lib.load()
    a = _f32c(a, "gemm operand"); b = _f32c(b, "gemm operand")
    if a.dim() != 2 or b.dim() != 2:
        raise RuntimeError("gemm needs 2-D operands, got %s and %s" % (tuple(a.shape), tuple(b.shape)))
    M, K = a.shape
    N = b.shape[0] if trans_b else b.shape[1]
    if (b.shape[1] if trans_b else b.shape[0]) != K:
        raise RuntimeError("size mismatch: gemm got %s and %s (trans_b=%s)" % (tuple(a.shape), tuple(b.shape), trans_b))
    if bias is not None:
        bias = _f32c(bias, "gemm bias")
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    if M == 0:
        return out
    nbytes = lib.rbm_b200_gemm_workspace(M, N, K)
    ws = _workspace(a.device, nbytes)
    with _on_device(a.device):
        _lib.check(lib.rbm_b200_gemm(1 if trans_b else 0, M, N, K, _ptr(a), _ptr(b), _ptr(bias), 1 if relu else 0, _ptr(out),
                                     _ptr(ws), ws.numel(), _stream()), "gemm")
    return out


def gemm_tn(a, b):
    """a.T @ b for fp32 row-major a [R, M], b [R, N] (weight gradients: the reduction runs over the batch): three
    tcgen05 products of the bf16 high / low splits, split over the rows, partial tiles added in fp32 (rbm_b200_gemm_tn)."""
    lib = _lib.load()
    a = _f32c(a, "gemm operand"); b = _f32c(b, "gemm operand")
    if a.dim() != 2 or b.dim() != 2 or a.shape[0] != b.shape[0]:
        raise RuntimeError("size mismatch: gemm_tn got %s and %s" % (tuple(a.shape), tuple(b.shape)))
    R, M = a.shape
    N = b.shape[1]
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    nbytes = lib.rbm_b200_gemm_tn_workspace(R, M, N)
    ws = _workspace(a.device, nbytes)
    with _on_device(a.device):
        _lib.check(lib.rbm_b200_gemm_tn(_ptr(a), _ptr(b), R, M, N, _ptr(out), _ptr(ws), ws.numel(), _stream()), "gemm_tn")
    return out


class _Linear(torch.autograd.Function):
    """y = act(x W^T + b) (torch.nn.Linear + optional ReLU: the layers of hierarchicalEncoder.py:86-96 and
    basicCoders.py:28-84) with forward AND backward on the tensor cores:
    dx = g' W, dW = g'^T x, db = sum_rows g' with g' = g * [y > 0] behind a ReLU."""

    @staticmethod
    def forward(ctx, x, W, b, relu):
        y = gemm(x, W, True, bias=b, relu=relu)
        ctx.save_for_backward(x, W, y if relu else None)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, W, y = ctx.saved_tensors
        g = g.contiguous()
        if y is not None:
            g = g * (y > 0)
        dx = gemm(g, W, False) if ctx.needs_input_grad[0] else None
        dW = gemm_tn(g, x) if ctx.needs_input_grad[1] else None
        db = g.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx, dW, db, None


def linear(x, W, b=None, relu=False):
    return _Linear.apply(x, W, b, relu)


class _GumbelSmooth(torch.autograd.Function):
    """(clamp(l, -88, 88), GumbelMod(clamp(l), beta, is_training)) with Philox draws (hierarchicalEncoder.py:167-176,
    gumbelmod.py:14-24).  Backward: d l = (g_logits + g_samples * beta * z (1 - z)) * [-88 <= l <= 88] in training mode;
    the evaluation-mode gate has no gradient (heaviside)."""

    @staticmethod
    def forward(ctx, raw, beta, is_training, seed, chain_offset, step, unit_offset):
        lib = _lib.load()
        raw = _f32c(raw, "logits")
        logits = raw.clone()
        samples = torch.empty_like(logits)
        with _on_device(raw.device):
            _lib.check(lib.rbm_b200_gumbel_smooth(_ptr(logits), _ptr(samples), logits.shape[0], logits.shape[1], float(beta),
                                                  1 if is_training else 0, 1, int(seed), int(chain_offset), int(step),
                                                  int(unit_offset), _stream()), "gumbel_smooth")
        ctx.save_for_backward(raw, samples)
        ctx.beta, ctx.is_training = float(beta), bool(is_training)
        if not is_training:
            ctx.mark_non_differentiable(samples)
        return logits, samples

    @staticmethod
    def backward(ctx, g_logits, g_samples):
        raw, z = ctx.saved_tensors
        g = g_logits
        if ctx.is_training:
            g = g + g_samples * (ctx.beta * z * (1.0 - z))
        return g * ((raw >= -88.0) & (raw <= 88.0)), None, None, None, None, None, None


def gumbel_smooth(raw_logits, beta, is_training, seed=0, chain_offset=0, step=0, unit_offset=0):
    return _GumbelSmooth.apply(raw_logits, beta, is_training, seed, chain_offset, step, unit_offset)


class _MeanFieldHalfStep(torch.autograd.Function):
    """sigmoid(x W + c) / sigmoid(x W^T + a) with the forward in CUDA (gibbsSampler.py:20-28).

    Backward is the analytic gradient of the same expression (the reference gets it
    from autograd): g_pre = g * p (1 - p); dx = g_pre Wop^T; dW, dbias accordingly - the products through
    rbm_b200_gemm / rbm_b200_gemm_tn (tensor cores), not ATen matmuls.
    """

    @staticmethod
    def forward(ctx, x, W, bias, direction, pack):
        out = half_step(pack, direction, x, _lib.MODE_PROBS)
        ctx.save_for_backward(x, W, out)
        ctx.direction = direction
        return out

    @staticmethod
    def backward(ctx, g):
        x, W, p = ctx.saved_tensors
        gp = g * p * (1.0 - p)
        x2, gp2 = (x.unsqueeze(0), gp.unsqueeze(0)) if x.dim() == 1 else (x, gp)
        gp2 = gp2.contiguous()
        if ctx.direction == 0:          # p = sigmoid(x W + c): dx = g' W^T, dW = x^T g'   (tcgen05 GEMMs, split bf16)
            dx = gemm(gp2, W, True)
            dW = gemm_tn(x2, gp2)
        else:                           # p = sigmoid(x W^T + a): dx = g' W, dW = g'^T x
            dx = gemm(gp2, W, False)
            dW = gemm_tn(gp2, x2)
        dbias = gp2.sum(0)
        if x.dim() == 1:
            dx = dx.squeeze(0)
        return dx, dW, dbias, None, None


def meanfield_half_step(rbm, pack, direction, x):
    bias = rbm.hidden_bias if direction == 0 else rbm.visible_bias
    return _MeanFieldHalfStep.apply(x, rbm.weights, bias, direction, pack)


class _Energy(torch.autograd.Function):
    """E[b] = -v.a - h.c - sum_j (vW)_j h_j, forward in CUDA (models/rbm/rbm.py:56-72)."""

    @staticmethod
    def forward(ctx, v, h, W, a, c, pack):
        lib = _lib.load()
        p, keep = pack
        v = _f32c(v, "energy input"); h = _f32c(h, "energy input")
        if v.dim() != 2 or h.dim() != 2 or v.shape[1] != p.n_visible or h.shape[1] != p.n_hidden or v.shape[0] != h.shape[0]:
            # the reference fails here inside torch.matmul (rbm.py:68-70)
            raise RuntimeError("size mismatch: energy got v %s, h %s for an RBM of %d x %d units"
                               % (tuple(v.shape), tuple(h.shape), p.n_visible, p.n_hidden))
        B = v.shape[0]
        E = torch.empty(B, dtype=torch.float32, device=v.device)
        nbytes = lib.rbm_b200_energy_workspace(p.n_visible, p.n_hidden, B)
        ws = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=v.device)
        with torch.cuda.device(v.device):
            _lib.check(lib.rbm_b200_energy(ctypes.byref(p), _ptr(v), _ptr(h), _ptr(E), B, _ptr(ws), nbytes,
                                           _stream()), "energy")
        ctx.save_for_backward(v, h, W, a, c)
        return E

    @staticmethod
    def backward(ctx, g):
        v, h, W, a, c = ctx.saved_tensors
        # three tcgen05 GEMMs on split-bf16 operands (rbm_b200_gemm / rbm_b200_gemm_tn) instead of ATen matmuls
        gv = -(g[:, None] * gemm(h, W, True, bias=a)) if ctx.needs_input_grad[0] else None     # a + h W^T
        gh = -(g[:, None] * gemm(v, W, False, bias=c)) if ctx.needs_input_grad[1] else None    # c + v W
        gW = -gemm_tn(v * g[:, None], h) if ctx.needs_input_grad[2] else None
        ga = -(v * g[:, None]).sum(0)
        gc = -(h * g[:, None]).sum(0)
        return gv, gh, gW, ga, gc, None


def energy(rbm, pack, v, h):
    return _Energy.apply(v, h, rbm.weights, rbm.visible_bias, rbm.hidden_bias, pack)


def negative_phase_stats(v, h, scale=None):
    """(S_vh, S_v, S_h) = scale * (V^T H, sum V, sum H); scale defaults to 1/B (means)."""
    lib = _lib.load()
    v = _f32c(v, "visible states"); h = _f32c(h, "hidden states")
    if v.dim() != 2 or h.dim() != 2 or v.shape[0] != h.shape[0]:
        raise RuntimeError("size mismatch: statistics need v [B,nV] and h [B,nH]")
    B, nV = v.shape
    nH = h.shape[1]
    if scale is None:
        scale = 1.0 / max(B, 1)
    S = torch.empty((nV, nH), dtype=torch.float32, device=v.device)
    sv = torch.empty(nV, dtype=torch.float32, device=v.device)
    sh = torch.empty(nH, dtype=torch.float32, device=v.device)
    with torch.cuda.device(v.device):
        _lib.check(lib.rbm_b200_negative_phase(_ptr(v), _ptr(h), B, nV, nH, scale, _ptr(S), _ptr(sv), _ptr(sh),
                                               _stream()), "negative_phase")
    return S, sv, sh


class _EnergyExpFromStats(torch.autograd.Function):
    """mean_b E(v_b, h_b) = -(<W, S_vh> + a.S_v + c.S_h) with gradients -S_vh, -S_v, -S_h.

    Replaces GumBolt.energy_exp on detached PCD samples (gumbolt.py:102-134): same value,
    same autograd result, without the [B, nV, nH] broadcast.
    """

    @staticmethod
    def forward(ctx, W, a, c, S, sv, sh, value):
        ctx.save_for_backward(S, sv, sh)
        if value is not None:        # already computed by rbm_b200_block_gibbs_stats
            return value.clone()
        return -(torch.dot(W.reshape(-1), S.reshape(-1)) + torch.dot(a, sv) + torch.dot(c, sh))

    @staticmethod
    def backward(ctx, g):
        S, sv, sh = ctx.saved_tensors
        return -g * S, -g * sv, -g * sh, None, None, None, None


def energy_exp_from_stats(rbm, S, sv, sh, value=None):
    return _EnergyExpFromStats.apply(rbm.weights, rbm.visible_bias, rbm.hidden_bias, S, sv, sh, value)


def energy_exp(rbm, rbm_vis, rbm_hid):
    """mean_b(-v^T W h - a.v - c.h): drop-in for GumBolt.energy_exp / the DiVAEPP KL energy terms on
    POSTERIOR samples (gumbolt.py:97-99,109-134; dvaepp.py:188-200 — SURVEY.md §8f rank 1).

    Same value and the same gradients w.r.t. the (real-valued, smoothed) samples and (W, a, c) as the
    reference expression, without materialising W broadcast to [B, nV, nH]: forward is the fused energy
    kernel (rbm_b200_energy), backward three [B,n] x [n,n] products on the tensor cores (split-bf16 tcgen05 GEMMs)."""
    pack = rbm._pack.pack(rbm, need_bf16=False)
    return energy(rbm, pack, rbm_vis, rbm_hid).mean()
