"""Build librbm_b200.so in-tree with nvcc for sm_100a (no torch headers, plain C ABI).

`python -m divae_b200.build` or `divae_b200.build.build()`.  The shared object is
git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librbm_b200.so")
STAMP = os.path.join(HERE, ".librbm_b200.stamp")
SOURCES = ["abi.cu", "gibbs_general.cu", "gibbs_smem.cu", "gibbs_umma.cu", "umma_decoder.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-cudart", "static",
] + os.environ.get("RBM_B200_NVCC_DEFS", "").split()


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    raise RuntimeError("nvcc not found")


def have_nvcc():
    try:
        nv = _nvcc()
    except RuntimeError:
        return False
    if nv == "nvcc":
        import shutil
        return shutil.which("nvcc") is not None
    return True


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode() + f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build():
    if not os.path.isfile(LIB) or not os.path.isfile(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _digest()


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into divae_b200/librbm_b200.so."""
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-Xptxas", "-v" if verbose else "-warn-spills", "-c",
               os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
    cmd = [_nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xlinker", "--exclude-libs,ALL", *objs, "-o", LIB]
    subprocess.check_call(cmd)
    with open(STAMP, "w") as f:
        f.write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
