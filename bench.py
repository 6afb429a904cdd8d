#!/usr/bin/env python
"""Benchmark of the hot path: batched expression-tree evaluation (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload) = BASELINE.json configs[1]: 1 000 random depth-8 trees,
5 features, Float32, 2^16 samples per GPU.  A step = one pass of the hot path over that
batch (all trees x all samples).  metric = node-ops/s = sum_trees count_nodes(tree) x
nsamples / second.  With N > 1 GPUs (launched by torchrun, one rank per GPU) every rank
evaluates its own 2^16-sample column block of X against the replicated population (the
sample axis shards with no data-path collective => "scaling": "weak"); the time is the max
over ranks and the value is the whole-job aggregate.

One JSON line on stdout (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel (eval_kernel<float>) vs the measured HBM copy bandwidth
  cpu_baseline  the CPU oracle (C port of the reference algorithm) on this box's cores
  e2e           the same metric through the host-buffer C-ABI entry point dex_eval_host
                (pinned host X in, host results out; copies inside the timed region)
`--impl reference` times the reference algorithm's CPU port (oracle/, OpenMP over trees —
Julia cannot run here) on the same config and prints the same line with impl=reference.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TREES, DEPTH, NFEATURES, NSAMPLES = 1000, 8, 5, 1 << 16
METRIC = "node-ops/sec (tree_nodes x samples / s), Float32"
UNIT = "node-ops/s"
BYTES_PER_UNIT = NFEATURES * 4 + 4        # X column read + result store = 24 B


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload():
    from dexb200 import treegen
    nodes, offsets = treegen.gen_population(N_TREES, DEPTH, 2, 4, NFEATURES, seed=0)
    return nodes, offsets


def make_X(rank):
    rng = np.random.default_rng([0, rank])
    return np.ascontiguousarray(rng.standard_normal((NSAMPLES, NFEATURES)).astype(np.float32))  # (N, F)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        # NVML / nvidia-smi enumerate physical GPUs: honour CUDA_VISIBLE_DEVICES when it is a list of indices
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        try:
            ids = [int(v) for v in vis.split(",") if v.strip() != ""]
            if ids and index < len(ids):
                index = ids[index]
        except ValueError:
            pass
        self.index = index
        self.rows = []
        self.stop = threading.Event()
        self.th = threading.Thread(target=self.run, daemon=True)

    def _nvml_open(self):
        """NVML is opened BEFORE the timed region starts (it takes tens of milliseconds), so that
        even a 10 ms region is sampled every 2 ms."""
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self._mx = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            self._R = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                       "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                       "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                       "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}
            self._nvml_sample()
            return True
        except Exception:
            self._nv = None
            return False

    def _nvml_sample(self):
        nv, h = self._nv, self._h
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        row = [str(self.index), str(sm), str(self._mx), "", ""]
        row += ["Active" if (mask & self._R[nm]) else "Not Active"
                for nm in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")]
        self.rows.append(row)

    def run(self):
        # NVML in-process (sub-millisecond per sample); nvidia-smi (the recipe's clocks line) as fallback
        if getattr(self, "_nv", None) is not None:
            while not self.stop.is_set():
                try:
                    self._nvml_sample()
                except Exception:
                    break
                self.stop.wait(0.002)
            return
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                   timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.splitlines()[0].split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self._nvml_open()
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    """Host threads this process may use (torchrun exports OMP_NUM_THREADS=1, so ask the OS)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline(nodes, offsets, opcodes, X_nf, budget_s=10.0):
    """The oracle timed on this box: all threads, the whole workload repeated for ~budget_s."""
    from oracle import oracle
    cores = host_cores()
    X = np.ascontiguousarray(X_nf.T)        # (F, N) view the wrapper expects
    out = np.empty((N_TREES, X.shape[1]), np.float32)
    oracle.eval_population(nodes, offsets, opcodes, X, nthreads=cores, out=out)      # warm-up
    reps, t_total = 0, 0.0
    while t_total < budget_s and reps < 200:
        t0 = time.perf_counter()
        oracle.eval_population(nodes, offsets, opcodes, X, nthreads=cores, out=out)
        t_total += time.perf_counter() - t0
        reps += 1
    nodeops = float(offsets[-1]) * X.shape[1] * reps
    t1 = time.perf_counter()
    oracle.eval_population(nodes, offsets, opcodes, X, nthreads=1, out=out)
    one = float(offsets[-1]) * X.shape[1] / (time.perf_counter() - t1)
    return {"value": nodeops / t_total, "unit": UNIT, "cores": cores, "kind": "port",
            "single_thread_value": one,
            "sample": f"the whole workload ({N_TREES} trees x {X.shape[1]} samples) x {reps} repetitions, "
                      f"{t_total:.1f} s; C port of src/Evaluate.jl (oracle/), OpenMP over trees on {cores} threads"}


def run_reference(args, rank, world):
    """`--impl reference`: the reference algorithm's CPU port on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import dexb200
    from dexb200 import treegen
    from oracle import oracle
    nodes, offsets = workload()
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    X = np.ascontiguousarray(make_X(0).T)
    cores = host_cores()
    # each step: the whole workload (it takes ~0.1 s on a multi-core host)
    ns = NSAMPLES
    Xs = np.ascontiguousarray(X[:, :ns])
    out = np.empty((N_TREES, ns), np.float32)
    for _ in range(args.warmup):
        oracle.eval_population(nodes, offsets, ops.opcodes, Xs, nthreads=cores, out=out)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.eval_population(nodes, offsets, ops.opcodes, Xs, nthreads=cores, out=out)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = float(offsets[-1]) * ns / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"the whole workload per step ({N_TREES} trees x {ns} samples); C port of the "
                                   f"reference algorithm (oracle/), OpenMP over trees; Julia is not installed"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def config_dict():
    return {"workload": "BASELINE.json configs[1]: 1k random depth-8 trees, 5 features, Float32, 2^16 samples "
                        "(per GPU), operator set (+,-,/,*),(cos,exp) of benchmark/benchmarks.jl:32-36",
            "n_trees": N_TREES, "depth": DEPTH, "nfeatures": NFEATURES, "nsamples_per_gpu": NSAMPLES,
            "tree_seed": 0, "parallelism": "sample-sharded, one rank per GPU, population replicated",
            "l2": "flushed between timed steps by writing a 512 MiB buffer; results (262 MB) exceed L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dexb200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import dexb200
    from dexb200 import device as D, treegen

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    dev = torch.device(f"cuda:{local_rank}")

    nodes, offsets = workload()
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    ctx = D.Context.get(local_rank)
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)
    total_nodes = int(pop.info["n_nodes"])
    node_ops_per_step = float(total_nodes) * NSAMPLES          # per GPU

    X_host = torch.from_numpy(make_X(rank)).pin_memory()       # (N, F) row-major == (F, N) column-major
    Xd = X_host.to(dev)
    Xview = Xd.T                                               # (F, N), strides (1, F): zero-copy
    out = torch.empty((N_TREES, NSAMPLES), dtype=torch.float32, device=dev)
    ok = torch.empty(N_TREES, dtype=torch.uint8, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------
    for _ in range(args.warmup):
        pop.eval(Xview, out=out, ok=ok)
    barrier()
    launches0 = ctx.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clk:
        barrier()
        t_wall0 = time.perf_counter()
        for a, b in evs:
            flush.fill_(1)                    # evict X / tapes from L2 (outside the event pair)
            a.record()
            pop.eval(Xview, out=out, ok=ok)
            b.record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count - launches0
    ms = [a.elapsed_time(b) for a, b in evs]
    ms_step = sum(ms) / len(ms)
    t = torch.tensor([ms_step], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step_max = float(t.item())
    value = node_ops_per_step * world / (ms_step_max * 1e-3)

    # ---- end to end through the host-buffer C-ABI entry point ---------------------------
    out_host = torch.empty((N_TREES, NSAMPLES), dtype=torch.float32).pin_memory()
    ok_host = torch.empty(N_TREES, dtype=torch.uint8).pin_memory()
    for _ in range(2):
        pop.eval_host(X_host, out_host, ok_host)
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pop.eval_host(X_host, out_host, ok_host)    # H2D X, kernels, D2H results + flags, sync
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = node_ops_per_step * world / float(te.item())

    # ---- the fused-loss entry point end to end (what optimiser-style callers consume): pinned X
    # and y in, P float64 losses + flags out — no (P x N) result matrix crosses PCIe
    y_host = torch.from_numpy(np.random.default_rng([1, rank]).standard_normal(NSAMPLES).astype(np.float32)).pin_memory()
    loss_host = torch.empty(N_TREES, dtype=torch.float64).pin_memory()

    def loss_step():
        xd = X_host.to(dev, non_blocking=True)
        yd = y_host.to(dev, non_blocking=True)
        loss, okl = pop.eval_loss(xd.T, yd)
        loss_host.copy_(loss, non_blocking=True)
        ok_host.copy_(okl, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(2):
        loss_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        loss_step()
    loss_s = (time.perf_counter() - t0) / e2e_steps
    tl = torch.tensor([loss_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tl, op=dist.ReduceOp.MAX)
    e2e_loss_value = node_ops_per_step * world / float(tl.item())

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = float(N_TREES) * NSAMPLES * BYTES_PER_UNIT      # per launch (per GPU)
        achieved = alg_bytes / (ms_step * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "eval_kernel_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        # instruction-issue roofline of the interpreter (SURVEY.md §8d caveat ii): warp instructions
        # of one launch (ncu, profiles/) at 148 SMs x 4 schedulers x the SM clock sampled above
        issue = None
        mpath = os.path.join(ROOT, "profiles", "r01_eval_kernel_ncu_metrics.json")
        clocks = clk.summary()
        if os.path.exists(mpath) and clocks.get("sm_mhz"):
            try:
                inst = float(json.load(open(mpath))["smsp__inst_executed.sum"]["value"])
                sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
                issue_peak = sm_count * 4 * clocks["sm_mhz"] * 1e6
                issue = {"bound": "issue", "warp_instructions_per_launch": inst,
                         "peak_warp_instructions_per_s": issue_peak, "ms_at_peak": inst / issue_peak * 1e3,
                         "frac": (inst / issue_peak * 1e3) / ms_step,
                         "source": "smsp__inst_executed.sum of profiles/r01_eval_kernel_ncu_metrics.json"}
            except Exception:
                issue = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step_max, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(),
            "population": {"total_nodes": total_nodes, "mean_nodes_per_tree": total_nodes / N_TREES,
                           "tape_instructions": int(pop.info["n_instructions"]),
                           "stack_rows": int(pop.info["max_stack"]),
                           "complete_fraction": float(ok.float().mean().item())},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "read_only_frac": (achieved * (NFEATURES * 4) / BYTES_PER_UNIT) / peak,
                         "kernel": "dex::eval_kernel<float,false,false>",
                         "note": "issue-bound interpreter: see DESIGN.md for the instruction roofline"},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(X_host.numel() * 4),
                    "d2h_bytes_per_step": int(out_host.numel() * 4 + ok_host.numel()),
                    "ms_per_step": float(te.item()) * 1e3, "entry": "dex_eval_host (pinned host buffers)"},
            "e2e_fused_loss": {"value": e2e_loss_value, "unit": UNIT,
                               "h2d_bytes_per_step": int(X_host.numel() * 4 + y_host.numel() * 4),
                               "d2h_bytes_per_step": int(loss_host.numel() * 8 + ok_host.numel()),
                               "ms_per_step": float(tl.item()) * 1e3,
                               "entry": "dex_eval_loss (per-tree MSE; the P x N results never leave the SM)"},
            "issue_roofline": issue,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "wall_s_timed_region": t_wall,
            "ms_per_step_min": min(ms), "ms_per_step_median": statistics.median(ms),
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(nodes, offsets, ops.opcodes, X_host.numpy())
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
