#!/usr/bin/env python
"""Benchmark of the RBM block-Gibbs hot path (BASELINE.json metric: unit-updates/s).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU torch sampler

Workload (config.workload): BASELINE.json configs[3] — "Large-prior shower generation:
RBM 2x1024, 65,536 chains x 1000 Gibbs steps, chains sharded over 1/2/4/8 B200", the
configuration the metric is quoted on at 1/2/4/8 GPUs.  One STEP = one
PCD.block_gibbs_sampling() call = 65,536 chains x 1000 Gibbs steps on a 1024x1024 prior
per GPU = 1.342e11 unit-updates (weak scaling: every rank runs 65,536 chains, chain ids
offset by rank so the draws are those of one 65,536*N-chain job).  Synthetic data:
reference-initialised parameters (models/rbm/rbm.py:28-34, torch.manual_seed(32)), chains
initialised Bernoulli(1/2) on the device.

Printed JSON (one line, rank 0):
  value      device-resident throughput: inputs (W, a, c) already in HBM, CUDA events around
             exactly K calls, max over ranks.
  e2e        the same metric through the public API with HOST buffers: every step copies the
             parameters host->device from pinned memory and the sampled (v, h) device->host.
  roofline   tensor-pipe roofline of the dominant kernel (umma_half_step_kernel):
             achieved = 2*B*nV*nH flops per launch / (timed call duration / 2k launches).
  cpu_baseline  the reference sampler's torch port (oracle/ref_torch_port.py) on the host cores,
             bounded sample of the same workload (rank 0, N=1 only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_VISIBLE = N_HIDDEN = 1024
CHAINS_PER_GPU = 65536
GIBBS_STEPS = 1000
WORKLOAD = "BASELINE configs[3]: RBM 1024x1024, 65536 chains/GPU x 1000 block-Gibbs steps per call"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(source="measured", hbm_gbs=p["hbm_gbs"], tflops_burst=p["bf16_tflops"],
                    tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]))
    return dict(source="fallback", hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=fd, stderr=subprocess.DEVNULL)
            os.close(fd)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        clocks, mx, reasons, power = [], [], set(), []
        try:
            with open(self.path) as f:
                for line in f:
                    parts = [x.strip() for x in line.split(",")]
                    if len(parts) < 9:
                        continue
                    try:
                        clocks.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
                    except ValueError:
                        continue
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                         parts[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if clocks:
            busy = sorted(clocks)[len(clocks) // 4:]  # drop idle samples at the edges
            out.update(sm_mhz=statistics.median(busy), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(clocks),
                       power_w=statistics.median(sorted(power)[len(power) // 4:]) if power else None)
        return out


def unit_updates(chains, k, nV, nH):
    return chains * k * (nV + nH)


# ------------------------------------------------------------------- CPU baseline / reference arm
def time_cpu_port(nV, nH, chains, k, repeats, threads):
    """Times oracle/ref_torch_port.block_gibbs_sampling (the reference's ATen op chain) on the host."""
    import torch
    from oracle import ref_torch_port as port
    torch.set_num_threads(threads)
    torch.manual_seed(32)
    W = torch.randn(nV, nH) * 0.01
    a = torch.ones(nV) * 0.5
    c = torch.zeros(nH)
    state = port.pcd_init_state(chains, nV)
    best = None
    with torch.no_grad():
        port.block_gibbs_sampling(state[:64], W, a, c, 1)  # warm-up
        for _ in range(repeats):
            t0 = time.perf_counter()
            v, h, state = port.block_gibbs_sampling(state, W, a, c, k)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return unit_updates(chains, k, nV, nH) / best, best


def cpu_baseline(args):
    threads = os.cpu_count() or 1
    chains, k = args.cpu_chains, args.cpu_steps
    rate, dt = time_cpu_port(args.n, args.n, chains, k, repeats=2, threads=threads)
    return {"value": rate, "unit": "unit-updates/s", "cores": threads, "kind": "port",
            "sample": "oracle/ref_torch_port (reference ATen op chain, pcd.py:51-69) on RBM %dx%d, %d chains x %d steps, "
                      "best of 2, %.2f s per call" % (args.n, args.n, chains, k, dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    chains, k = args.cpu_chains, args.cpu_steps
    import torch
    from oracle import ref_torch_port as port
    torch.set_num_threads(threads)
    torch.manual_seed(32)
    W = torch.randn(args.n, args.n) * 0.01
    a = torch.ones(args.n) * 0.5
    c = torch.zeros(args.n)
    state = port.pcd_init_state(chains, args.n)
    times = []
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            v, h, state = port.block_gibbs_sampling(state, W, a, c, k)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = unit_updates(chains, k, args.n, args.n) * len(times) / total
    sample = ("each step = reference sampler (torch port of models/samplers/pcd.py:51-69) on RBM %dx%d, %d chains x %d "
              "Gibbs steps, %d host threads" % (args.n, args.n, chains, k, threads))
    line = {"impl": "reference", "metric": "rbm_block_gibbs_unit_updates_per_s", "value": value, "unit": "unit-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "unit-updates/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "unit-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import divae_b200
    from divae_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load(build_if_missing=False)

    n, chains, k = args.n, args.chains, args.k
    torch.manual_seed(32)                                   # scripts/run.py:17
    rbm = divae_b200.RBM(n_visible=n, n_hidden=n).to(dev)   # reference init (rbm.py:28-34)
    sampler = divae_b200.PCD(batch_size=chains, RBM=rbm, n_gibbs_sampling_steps=k, seed=0x5EED,
                             variant=args.variant, chain_offset=rank * chains)
    units_per_step = unit_updates(chains, k, n, n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return x

    # ---------------- device-resident: `value`
    for _ in range(args.warmup):
        sampler.block_gibbs_sampling()
    barrier()
    sampler_clock = ClockSampler(local_rank)
    if rank == 0:
        sampler_clock.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        v, h = sampler.block_gibbs_sampling()
    e1.record()
    barrier()
    clocks = sampler_clock.stop() if rank == 0 else {}
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / args.steps
    value = world * units_per_step / (ms_per_step * 1e-3)
    launches_per_step = 2 * k + 6   # 2k half-step GEMMs + W pad + 2 bias + init + 2 output conversions

    # ---------------- end to end through the public API with host buffers: `e2e`
    W_host = rbm.weights.detach().cpu().pin_memory()
    a_host = rbm.visible_bias.detach().cpu().pin_memory()
    c_host = rbm.hidden_bias.detach().cpu().pin_memory()
    v_host = torch.empty((chains, n), dtype=torch.float32).pin_memory()
    h_host = torch.empty((chains, n), dtype=torch.float32).pin_memory()

    def e2e_step():
        with torch.no_grad():
            rbm._weights.copy_(W_host, non_blocking=True)
            rbm._visible_bias.copy_(a_host, non_blocking=True)
            rbm._hidden_bias.copy_(c_host, non_blocking=True)
        vv, hh = sampler.block_gibbs_sampling()
        v_host.copy_(vv, non_blocking=True)
        h_host.copy_(hh, non_blocking=True)

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_step()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / e2e_steps
    h2d = 4 * (n * n + 2 * n)
    d2h = 4 * 2 * chains * n
    assert set(torch.unique(v_host[:256]).tolist()) <= {0.0, 1.0}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    flops_per_launch = 2.0 * chains * n * n
    launch_ms = ms_per_step / (2 * k)
    achieved = flops_per_launch / (launch_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get("umma_half_step_dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": "rbm_block_gibbs_unit_updates_per_s", "value": value, "unit": "unit-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD if (n, chains, k) == (N_VISIBLE, CHAINS_PER_GPU, GIBBS_STEPS)
                   else "RBM %dx%d, %d chains/GPU x %d steps (non-default)" % (n, n, chains, k),
                   "n_visible": n, "n_hidden": n, "chains_per_gpu": chains, "gibbs_steps": k,
                   "variant": args.variant, "seed": "0x5EED",
                   "l2": "state tiles (2 x %d MiB bf16) + fp32 outputs exceed the 126 MB L2; no explicit flush"
                         % (chains * n * 2 >> 20)},
        "e2e": {"value": world * units_per_step / (e2e_ms * 1e-3), "unit": "unit-updates/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": {"sm_mhz": clocks.get("sm_mhz"), "sm_max_mhz": clocks.get("sm_max_mhz"),
                   "reasons": clocks.get("reasons", []), "samples": clocks.get("samples", 0),
                   "power_w": clocks.get("power_w")},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["tflops_sustained"], "traffic": traffic,
                     "kernel": "umma_half_step_kernel", "flops_per_launch": flops_per_launch,
                     "launch_ms": launch_ms, "peak_source": peaks["source"] + " bf16_tflops_sustained",
                     "frac_of_burst": achieved / peaks["tflops_burst"]},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_VISIBLE, help="units per RBM side")
    ap.add_argument("--chains", type=int, default=CHAINS_PER_GPU, help="chains per GPU")
    ap.add_argument("--k", type=int, default=GIBBS_STEPS, help="Gibbs steps per call")
    ap.add_argument("--variant", default="auto")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-chains", type=int, default=4096, help="bounded CPU sample: chains")
    ap.add_argument("--cpu-steps", type=int, default=10, help="bounded CPU sample: Gibbs steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
