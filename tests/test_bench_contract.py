"""bench.py contract pieces that run without a GPU: the reference arm's JSON line and helpers."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_required_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-chains", "64", "--cpu-steps", "1", "--n", "128"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "unit-updates/s" and d["value"] > 0
    # the staged / mounted reference tree when there is one, else the op-for-op port
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["scaling"] == "strong"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["vs_baseline"] is None and "workload" in d["config"]


def test_non_zero_ranks_of_the_reference_arm_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_graft_entry_build_produces_the_library():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    lib = ge.build()
    assert os.path.isfile(lib) and lib.endswith("librbm_b200.so")


def test_reference_arm_times_the_reference_classes_when_the_tree_is_present():
    sys.path.insert(0, ROOT)
    from oracle import _ref_shim
    if not _ref_shim.reference_available():
        import pytest
        pytest.skip("no reference tree")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-chains", "32", "--cpu-steps", "1", "--n", "64"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["kind"] == "reference" and "models/samplers/pcd.py" in d["config"]["sample"]


def test_ais_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "ais", "--steps", "1",
                          "--warmup", "0", "--cpu-chains", "16", "--cpu-steps", "8", "--n", "64"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and "configs[4]" in d["config"]["workload"]
