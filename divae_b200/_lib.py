"""ctypes binding of librbm_b200.so (C ABI: include/rbm_b200.h).

There is no CPU fallback and no alternative backend: if the shared object is
missing and cannot be built, or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librbm_b200.so")

VARIANT_AUTO, VARIANT_GENERAL, VARIANT_SMEM, VARIANT_UMMA, VARIANT_UMMA_SPLIT = 0, 1, 2, 3, 4
VARIANTS = {"auto": VARIANT_AUTO, "general": VARIANT_GENERAL, "smem": VARIANT_SMEM, "umma": VARIANT_UMMA,
            "umma_split": VARIANT_UMMA_SPLIT}
MODE_SAMPLE, MODE_INJECTED, MODE_PROBS = 0, 1, 2
ERR_INVALID, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_NO_DEVICE = -1, -2, -3, -4


class RbmParams(ctypes.Structure):
    """rbm_b200_params (include/rbm_b200.h)."""
    _fields_ = [
        ("n_visible", ctypes.c_int32),
        ("n_hidden", ctypes.c_int32),
        ("W", ctypes.c_void_p),
        ("W_bf16", ctypes.c_void_p),
        ("vbias", ctypes.c_void_p),
        ("hbias", ctypes.c_void_p),
    ]


class Linear(ctypes.Structure):
    """rbm_b200_linear (include/rbm_b200.h): one nn.Linear, torch layout."""
    _fields_ = [
        ("in_features", ctypes.c_int32),
        ("out_features", ctypes.c_int32),
        ("weight", ctypes.c_void_p),
        ("bias", ctypes.c_void_p),
    ]


class Decoder(ctypes.Structure):
    """rbm_b200_decoder (include/rbm_b200.h): trunk + zero or two heads."""
    _fields_ = [
        ("n_trunk", ctypes.c_int32),
        ("n_head", ctypes.c_int32),
        ("trunk", ctypes.POINTER(Linear)),
        ("head_hits", ctypes.POINTER(Linear)),
        ("head_act", ctypes.POINTER(Linear)),
    ]


_P = ctypes.POINTER(RbmParams)
_D = ctypes.POINTER(Decoder)
_vp, _i32, _i64, _u64, _f32, _sz, _int = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64,
                                          ctypes.c_float, ctypes.c_size_t, ctypes.c_int)

# name -> (restype, argtypes); every prototype of include/rbm_b200.h
PROTOTYPES = {
    "rbm_b200_abi_version": (_int, []),
    "rbm_b200_last_error": (ctypes.c_char_p, []),
    "rbm_b200_device_info": (_int, [ctypes.POINTER(_i32)] * 3),
    "rbm_b200_half_step": (_int, [_P, _int, _vp, _vp, _i64, _int, _vp, _u64, _u64, _u64, _vp]),
    "rbm_b200_init_chains": (_int, [_vp, _i64, _i32, _u64, _u64, _u64, _vp]),
    "rbm_b200_block_gibbs_workspace": (_sz, [_i32, _i32, _i64, _int]),
    "rbm_b200_kernel_choice": (ctypes.c_char_p, [_i32, _i32, _i64, _int]),
    "rbm_b200_block_gibbs_stats_workspace": (_sz, [_i32, _i32, _i64, _int]),
    "rbm_b200_block_gibbs_launch_count": (_i64, [_i32, _i32, _i64, _i32, _int]),
    "rbm_b200_block_gibbs": (_int, [_P, _vp, _vp, _i64, _i32, _int, _u64, _u64, _u64, _int, _vp, _sz, _vp]),
    "rbm_b200_block_gibbs_stats": (_int, [_P, _vp, _vp, _i64, _i32, _int, _u64, _u64, _u64, _int, _vp, _sz, _f32, _vp, _vp,
                                           _vp, _vp, _vp]),
    "rbm_b200_block_gibbs_injected": (_int, [_P, _vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "rbm_b200_meanfield": (_int, [_P, _vp, _vp, _i64, _i32, _vp]),
    "rbm_b200_energy_workspace": (_sz, [_i32, _i32, _i64]),
    "rbm_b200_energy": (_int, [_P, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "rbm_b200_negative_phase": (_int, [_vp, _vp, _i64, _i32, _i32, _f32, _vp, _vp, _vp, _vp]),
    "rbm_b200_cd_workspace": (_sz, [_i32, _i32, _i64]),
    "rbm_b200_cd_step": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _i64, _i32, _f32, _f32, _f32, _vp, _u64, _u64,
                                 _vp, _vp, _sz, _vp]),
    "rbm_b200_ais_workspace": (_sz, [_i32, _i32, _i64]),
    "rbm_b200_ais_run": (_int, [_P, _i64, ctypes.POINTER(ctypes.c_double), _i32, _u64, _u64, _vp, _vp, _sz, _vp]),
    "rbm_b200_ais_run_ex": (_int, [_P, _i64, ctypes.POINTER(ctypes.c_double), _i32, _u64, _u64, _vp, _int, _vp, _sz, _vp]),
    "rbm_b200_decoder_workspace": (_sz, [_D, _i64]),
    "rbm_b200_decoder_forward": (_int, [_D, _vp, _i64, _vp, _vp, _vp, _u64, _u64, _u64, _int, _vp, _sz, _vp]),
    "rbm_b200_gemm_workspace": (_sz, [_i64, _i32, _i32]),
    "rbm_b200_gemm": (_int, [_int, _i64, _i32, _i32, _vp, _vp, _vp, _int, _vp, _vp, _sz, _vp]),
    "rbm_b200_gemm_tn_workspace": (_sz, [_i64, _i32, _i32]),
    "rbm_b200_gemm_tn": (_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _sz, _vp]),
    "rbm_b200_gumbel_smooth": (_int, [_vp, _vp, _i64, _i32, _f32, _int, _int, _u64, _u64, _u64, ctypes.c_uint32, _vp]),
}

_lib = None
_lock = threading.Lock()


def load(build_if_missing=True):
    """Load (building first if the .so is absent and nvcc is present). Raises on failure."""
    global _lib
    if _lib is not None:      # fast path, no lock
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        if not os.path.isfile(LIB_PATH):
            if not build_if_missing:
                raise RuntimeError("librbm_b200.so is missing (%s); run `python -m divae_b200.build`" % LIB_PATH)
            _build.build()
        elif _build.needs_build():
            # the sources or the header changed after the binary was built: same prototypes, different kernels
            if build_if_missing and _build.have_nvcc():
                _build.build()
            elif os.environ.get("RBM_B200_ALLOW_STALE") != "1":
                raise RuntimeError("librbm_b200.so is older than divae_b200/csrc or include/ (source digest mismatch); "
                                   "run `python -m divae_b200.build` (RBM_B200_ALLOW_STALE=1 loads it anyway)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI is incomplete: fail loudly
            fn.restype = res
            fn.argtypes = args
        if lib.rbm_b200_abi_version() != 1:
            raise RuntimeError("librbm_b200.so ABI version mismatch")
        _lib = lib
        return lib


def check(rc, what=""):
    if rc:
        msg = load().rbm_b200_last_error().decode("utf-8", "replace")
        raise RuntimeError("rbm_b200 %s failed (code %d): %s" % (what, rc, msg))


def require_cuda(t, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError(
            "divae_b200: %s lives on %s; this library is CUDA-only (sm_100a) and has no CPU fallback. "
            "Move the model to a CUDA device first." % (name, t.device))
