"""The boundary is usable from plain C: examples/c_abi_demo.c includes include/rbm_b200.h as C99, binds the shared
library with dlopen and samples through rbm_b200_block_gibbs without Python or torch in the process."""
import os
import shutil
import subprocess

import pytest

from divae_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build_demo(tmp_path):
    gcc = shutil.which("gcc")
    if gcc is None or not os.path.isfile(os.path.join(CUDA, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc or the CUDA runtime headers are not available")
    exe = str(tmp_path / "c_abi_demo")
    subprocess.check_call([gcc, "-std=c99", "-O2", "-Wall", "-Werror", os.path.join(ROOT, "examples", "c_abi_demo.c"),
                           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA, "include"),
                           "-L" + os.path.join(CUDA, "lib64"), "-lcudart", "-ldl", "-lm", "-o", exe])
    return exe


def test_header_compiles_as_c99_and_the_demo_links(tmp_path):
    assert os.path.isfile(build_demo(tmp_path))


@pytest.mark.gpu
def test_demo_samples_through_the_c_abi(tmp_path):
    _lib.load()          # makes sure the shared object exists
    exe = build_demo(tmp_path)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(CUDA, "lib64") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([exe, _lib.LIB_PATH], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "c_abi_demo ok" in out.stdout and out.stdout.count("short workspace -> rc -3") >= 1
