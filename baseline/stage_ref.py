"""Stage the UNMODIFIED reference tree next to the repo so that it travels to the GPU box.

  python baseline/stage_ref.py            # /root/reference -> baseline/_ref/DiVAE (byte-for-byte copy)

TEST / BENCHMARK INFRASTRUCTURE.  `baseline/_ref/` is git-ignored (no reference source enters the history) but is
NOT gpurun-ignored, so the copy ships with the snapshot and the GPU box can run the reference's own classes:
the `--impl reference` arm and `cpu_baseline` of bench.py, the engine-level GPU tests (tests/test_gpu_reference_engine.py)
and benchmarks/train_step_bench.py import it through oracle/_ref_shim.py.  The reference is pure Python (no build
step); the `pip install --target baseline/_ref` recipe does not apply because the tree has no setup.py/pyproject.
A MANIFEST with the sha256 of every staged file is written so a stale or edited copy is detectable.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("DIVAE_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref", "DiVAE")
SKIP_DIRS = {".git", "__pycache__", "outputs", "wandb"}


def _files(root):
    for base, dirs, files in os.walk(root):
        dirs[:] = sorted(d for d in dirs if d not in SKIP_DIRS)
        for f in sorted(files):
            if f.endswith((".pyc", ".pyo")):
                continue
            yield os.path.relpath(os.path.join(base, f), root)


def manifest(root):
    out = []
    for rel in _files(root):
        if rel == "MANIFEST.sha256":
            continue
        with open(os.path.join(root, rel), "rb") as f:
            out.append("%s  %s" % (hashlib.sha256(f.read()).hexdigest(), rel))
    return "\n".join(out) + "\n"


def staged():
    return os.path.isfile(os.path.join(DST, "models", "rbm", "rbm.py"))


def stage(force=False):
    """Copy the reference tree; returns the staged root.  No-op when the source is absent and a copy exists."""
    if not os.path.isdir(SRC):
        if staged():
            return DST
        raise RuntimeError("reference tree not found at %s and nothing staged at %s" % (SRC, DST))
    want = manifest(SRC)
    mpath = os.path.join(DST, "MANIFEST.sha256")
    if not force and staged() and os.path.isfile(mpath):
        with open(mpath) as f:
            if f.read() == want:
                return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for rel in _files(SRC):
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
    with open(mpath, "w") as f:
        f.write(want)
    return DST


if __name__ == "__main__":
    root = stage(force="--force" in sys.argv)
    print(root, "(%d files)" % sum(1 for _ in _files(root)))
