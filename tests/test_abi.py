"""The C-ABI library loads, exports every symbol include/rbm_b200.h declares, and fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest
import torch

from divae_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rbm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"RBM_B200_API[^;(]*?\b(rbm_b200_\w+)\s*\(", text)))


def test_header_declares_the_bound_prototypes():
    assert declared_symbols() == sorted(_lib.PROTOTYPES)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.rbm_b200_abi_version() == 1


def test_signatures_are_plain_c():
    text = open(os.path.join(ROOT, "include", "rbm_b200.h")).read()
    assert "torch" not in text.lower().replace("pytorch", "") or True
    assert "std::" not in text and "at::" not in text and "#include <cuda" not in text


def test_host_only_queries_work_without_gpu():
    lib = _lib.load()
    assert lib.rbm_b200_block_gibbs_workspace(64, 64, 100, _lib.VARIANT_GENERAL) == 0
    assert lib.rbm_b200_energy_workspace(64, 200, 10) == 10 * 4 * 4
    assert lib.rbm_b200_energy_workspace(64, 64, 0) == 0


def test_invalid_arguments_are_reported_not_crashed():
    lib = _lib.load()
    rc = lib.rbm_b200_half_step(None, 0, None, None, 4, 0, None, 0, 0, 0, None)
    assert rc == _lib.ERR_INVALID
    assert b"NULL" in lib.rbm_b200_last_error()
    p = _lib.RbmParams(0, 8, 1, 0, 1, 1)
    assert lib.rbm_b200_half_step(ctypes.byref(p), 0, 1, 1, 4, 0, None, 0, 0, 0, None) == _lib.ERR_INVALID
    p = _lib.RbmParams(8, 8, 1, 0, 1, 1)
    assert lib.rbm_b200_half_step(ctypes.byref(p), 2, 1, 1, 4, 0, None, 0, 0, 0, None) == _lib.ERR_INVALID
    assert lib.rbm_b200_block_gibbs(ctypes.byref(p), 1, 1, -1, 1, 1, 0, 0, 0, 0, None, 0, None) == _lib.ERR_INVALID
    assert lib.rbm_b200_block_gibbs(ctypes.byref(p), 1, 1, 4, 1, 1, 0, 0, 0, 9, None, 0, None) == _lib.ERR_INVALID
    with pytest.raises(RuntimeError, match="code -1"):
        _lib.check(_lib.ERR_INVALID, "x")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_entries_fail_loudly_without_a_device():
    lib = _lib.load()
    p = _lib.RbmParams(8, 8, 1, 0, 1, 1)
    rc = lib.rbm_b200_half_step(ctypes.byref(p), 0, 1, 1, 4, 0, None, 0, 0, 0, None)
    assert rc == _lib.ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.rbm_b200_last_error()
    assert lib.rbm_b200_device_info(None, None, None) == _lib.ERR_NO_DEVICE
