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
  roofline       dominant kernel (eval_kernel<float>) vs the measured HBM copy bandwidth
  cpu_baseline   the CPU oracle (C port of the reference algorithm) on this box's cores
  e2e            the same metric through the host-buffer C-ABI entry point dex_eval_host
                 (pinned host X in, host results out; copies inside the timed region)
  e2e_with_pack  as e2e, but every step also flattens + uploads the population from its wire
                 form (dex_population_pack): what a caller whose trees change every generation pays
  e2e_fused_loss pinned X and y in, P float64 losses out (dex_eval_loss)
  configs        the other BASELINE.json configurations on this GPU, device-resident: C3 (gradients),
                 C4 shard shape, C5 (parametric), C6 (north_star target), C6 fused loss, Float64
  c4             BASELINE.json configs[3] strong-scaled over the N ranks (10k depth-12 trees, 2^20/N
                 columns per rank): kernel time, fused evaluate + gather over NVLink peer memory,
                 NCCL all-gather, bit-equality of the gathered rows with an unsharded evaluation
`--impl reference` times the reference's own CPU path: the Julia package when a `julia` binary
with DynamicExpressions is on the box (benchmarks/reference_julia.jl), else the reference
algorithm's C port (oracle/, OpenMP over trees) on the same config; same line, impl=reference.
"""
import argparse
import json
import os
import shutil
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
    flags = oracle.use_native_build()          # -march=native build made on THIS box when gcc is here
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
            "single_thread_value": one, "compiler_flags": flags,
            "sample": f"the whole workload ({N_TREES} trees x {X.shape[1]} samples) x {reps} repetitions, "
                      f"{t_total:.1f} s; C port of src/Evaluate.jl (oracle/, {flags}), OpenMP over trees on {cores} "
                      f"threads; plain libm, not the SLEEF-style vector math of LoopVectorization"}


def julia_reference(args, nodes, offsets, ops):
    """The real reference when it can run: a `julia` binary with DynamicExpressions installed.
    benchmarks/reference_julia.jl rebuilds the trees from the wire dump and times
    [eval_tree_array(tree, X, operators; turbo, bumper) for tree in trees]
    (/root/reference/benchmark/benchmarks.jl:76-91).  Returns its JSON dict or None."""
    exe = shutil.which("julia")
    if not exe:
        return None
    import tempfile
    try:
        with tempfile.TemporaryDirectory() as d:
            np.save(os.path.join(d, "nodes.npy"), np.ascontiguousarray(nodes).view(np.uint8))
            np.save(os.path.join(d, "offsets.npy"), offsets)
            np.save(os.path.join(d, "X.npy"), np.ascontiguousarray(make_X(0).T))
            cmd = [exe, f"--threads={host_cores()}", os.path.join(ROOT, "benchmarks", "reference_julia.jl"), d,
                   str(args.steps), str(args.warmup)]
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
            if r.returncode != 0:
                return None
            return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        return None


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import dexb200
    from dexb200 import treegen
    from oracle import oracle
    nodes, offsets = workload()
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    cores = host_cores()
    jl = julia_reference(args, nodes, offsets, ops)
    if jl is not None:      # the real LoopVectorization / Bumper numbers
        value, dt = float(jl["node_ops_per_s"]), float(jl["ms_per_step"]) * 1e-3
        kind, sample = "reference", jl.get("sample", "DynamicExpressions.jl eval_tree_array over the population (Julia)")
        extra = {"julia": jl}
    else:
        flags = oracle.use_native_build()
        X = np.ascontiguousarray(make_X(0).T)
        ns = NSAMPLES       # each step: the whole workload (it takes ~0.1 s on a multi-core host)
        Xs = np.ascontiguousarray(X[:, :ns])
        out = np.empty((N_TREES, ns), np.float32)
        for _ in range(args.warmup):
            oracle.eval_population(nodes, offsets, ops.opcodes, Xs, nthreads=cores, out=out)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            oracle.eval_population(nodes, offsets, ops.opcodes, Xs, nthreads=cores, out=out)
        dt = (time.perf_counter() - t0) / max(args.steps, 1)
        value = float(offsets[-1]) * ns / dt
        kind = "port"
        sample = (f"the whole workload per step ({N_TREES} trees x {ns} samples); Julia reference not runnable "
                  f"(no `julia` on this box): stand-in = C port of the reference algorithm (oracle/, {flags}), "
                  f"OpenMP over trees")
        extra = {}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    line.update(extra)
    print(json.dumps(line), flush=True)


def config_dict():
    return {"workload": "BASELINE.json configs[1]: 1k random depth-8 trees, 5 features, Float32, 2^16 samples "
                        "(per GPU), operator set (+,-,/,*),(cos,exp) of benchmark/benchmarks.jl:32-36",
            "n_trees": N_TREES, "depth": DEPTH, "nfeatures": NFEATURES, "nsamples_per_gpu": NSAMPLES,
            "tree_seed": 0, "parallelism": "sample-sharded, one rank per GPU, population replicated",
            "l2": "flushed between timed steps by writing a 512 MiB buffer; results (262 MB) exceed L2"}


def bind_to_gpu_numa(index):
    """Pin this process to the CPUs next to its GPU before any pinned host buffer is allocated
    (first-touch puts the pages on that NUMA node): with 8 ranks copying 262 MB each to one
    host, remote-node buffers cap the aggregate D2H rate."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [int(v) for v in vis.split(",") if v.strip() != ""] if vis else []
        phys = ids[index] if index < len(ids) else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def device_time(f, reps, flush, sampler_index=None):
    """CUDA-event time of `f` per call (ms list), L2 flushed between calls; SM clocks sampled
    during the timed region when sampler_index is given."""
    import torch
    f()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    clk = ClockSampler(sampler_index) if sampler_index is not None else None
    if clk:
        clk.__enter__()
    for a, b in evs:
        flush.fill_(1)
        a.record()
        f()
        b.record()
    torch.cuda.synchronize()
    if clk:
        clk.__exit__()
    return [a.elapsed_time(b) for a, b in evs], (clk.summary() if clk else None)


def run_configs(local_rank, flush, peak, reps=3):
    """The other BASELINE.json configurations, device-resident on this GPU:
    ms (best / median), node-ops/s, fraction of the HBM roofline from the algorithmic bytes, clocks."""
    import torch
    import dexb200
    from dexb200 import device as D, treegen
    dev = torch.device(f"cuda:{local_rank}")
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    res = {}

    def emit(name, desc, ms, clocks, nodeops, alg_bytes, extra=None):
        best = min(ms)
        r = {"desc": desc, "ms": best, "ms_median": statistics.median(ms), "node_ops_per_s": nodeops / (best * 1e-3),
             "algorithmic_bytes": alg_bytes, "roofline_frac": alg_bytes / (peak * 1e9) / (best * 1e-3),
             "clocks": clocks}
        r.update(extra or {})
        res[name] = r

    def pop_of(P, depth, F, n_params=0, dtype=np.float32):
        nodes, offsets = treegen.gen_population(P, depth, 2, 4, F, seed=0, n_params=n_params, dtype=dtype)
        return D.Population(None, ops, dtype, wire=(nodes, offsets), ctx=D.Context.get(local_rank))

    P, F, N = N_TREES, NFEATURES, NSAMPLES
    g = torch.Generator(device=dev).manual_seed(1234)
    # ---- C3: eval_grad_tree_array d/dX on the C2 population ---------------------------------
    pop = pop_of(P, DEPTH, F)
    X = torch.randn((N, F), device=dev, generator=g)
    ms, clk = device_time(lambda: pop.eval_grad(X.T, D.GRAD_FEATURES), reps + 2, flush, local_rank)
    emit("C3", "configs[2]: eval_grad_tree_array d/dX (G=5) of the C2 population, Float32, 2^16 samples", ms, clk,
         pop.info["n_nodes"] * N, P * N * (F * 4 + (1 + F) * 4))
    o2 = torch.empty((P, N), device=dev)
    k2 = torch.empty(P, device=dev, dtype=torch.uint8)
    ms, clk = device_time(lambda: pop.eval(X.T, out=o2, ok=k2, early_exit=False), reps + 2, flush, local_rank)
    emit("C2-no-early-exit", "the C2 population with EvalContext(early_exit = false)", ms, clk,
         pop.info["n_nodes"] * N, P * N * (F * 4 + 4))
    del o2
    y = torch.randn(N, device=dev, generator=g)
    ms, clk = device_time(lambda: pop.eval_loss_grad(X.T, y, D.GRAD_CONSTANTS), reps + 2, flush, local_rank)
    emit("C2-loss-grad", "fused MSE + d/dconstants on the C2 population (what constant optimisation consumes)", ms, clk,
         pop.info["n_nodes"] * N, P * N * (F * 4 + 4))
    del pop
    # ---- Float64 (north_star: Float64 at 1e-6) ----------------------------------------------------
    pop64 = pop_of(P, DEPTH, F, dtype=np.float64)
    X64 = torch.randn((N, F), device=dev, dtype=torch.float64, generator=g)
    o64 = torch.empty((P, N), device=dev, dtype=torch.float64)
    k64 = torch.empty(P, device=dev, dtype=torch.uint8)
    ms, clk = device_time(lambda: pop64.eval(X64.T, out=o64, ok=k64), reps + 2, flush, local_rank)
    emit("C2-f64", "the C2 population in Float64", ms, clk, pop64.info["n_nodes"] * N, P * N * (F * 8 + 8))
    ms, clk = device_time(lambda: pop64.eval_grad(X64.T, D.GRAD_FEATURES), reps, flush, local_rank)
    emit("C3-f64", "d/dX of the C2 population in Float64", ms, clk, pop64.info["n_nodes"] * N, P * N * (F * 8 + (1 + F) * 8))
    del pop64, X64, o64
    # ---- C5: ParametricExpression ---------------------------------------------------------------
    npar, ncls, N5 = 3, 10, 1 << 18
    pop = pop_of(P, DEPTH, F, n_params=npar)
    X = torch.randn((N5, F), device=dev, generator=g)
    params = torch.randn((P, npar, ncls), device=dev, generator=g)
    cls = torch.randint(0, ncls, (N5,), device=dev, generator=g)
    out = torch.empty((P, N5), device=dev)
    ok = torch.empty(P, device=dev, dtype=torch.uint8)
    pcm = params.permute(0, 2, 1).contiguous()              # per-tree column-major (n_params x n_classes)
    cls32 = cls.to(torch.int32)
    ms, clk = device_time(lambda: pop.eval_parametric_prepared(X.T, pcm, cls32, npar, ncls, out=out, ok=ok),
                          reps + 2, flush, local_rank)
    emit("C5", "configs[4]: ParametricExpression, 1k trees, 3 parameters x 10 classes, Float32, 2^18 samples", ms, clk,
         pop.info["n_nodes"] * N5, P * N5 * (F * 4 + 4 + npar * 4 + 4))
    del pop, out
    # ---- C4 shard shape: 10k depth-12 trees, 10 features, 2^17 of the 2^20 samples -----------------
    P4, F4, N4 = 10000, 10, 1 << 17
    pop = pop_of(P4, 12, F4)
    X = torch.randn((N4, F4), device=dev, generator=g)
    big = torch.empty((10000, 1 << 20), device=dev)         # 41.9 GB: shared by C4 and C6
    out = big.view(-1)[: P4 * N4].view(P4, N4)
    ok = torch.empty(P4, device=dev, dtype=torch.uint8)
    ms, clk = device_time(lambda: pop.eval(X.T, out=out, ok=ok), reps, flush, local_rank)
    emit("C4-shard", "configs[3], one of 8 sample shards: 10k depth-12 trees, 10 features, Float32, 2^17 samples", ms, clk,
         pop.info["n_nodes"] * N4, P4 * N4 * (F4 * 4 + 4),
         {"mean_nodes_per_tree": pop.info["n_nodes"] / P4, "complete_fraction": float(ok.float().mean())})
    del pop
    # ---- C6: the north_star target sentence ---------------------------------------------------------
    P6, N6 = 10000, 1 << 20
    pop = pop_of(P6, DEPTH, F)
    X = torch.randn((N6, F), device=dev, generator=g)
    ms, clk = device_time(lambda: pop.eval(X.T, out=big, ok=ok), reps, flush, local_rank)
    emit("C6", "north_star target: 10k depth-8 trees, 5 features, Float32, 2^20 samples, 1 GPU (target >= 0.50)", ms, clk,
         pop.info["n_nodes"] * N6, P6 * N6 * (F * 4 + 4),
         {"read_only_roofline_frac": P6 * N6 * (F * 4) / (peak * 1e9) / (min(ms) * 1e-3)})
    y = torch.randn(N6, device=dev, generator=g)
    # what a caller of the materialising path still has to do to get the losses: one more pass
    # over the 42 GB of results (torch ops = the CALLER's side, timed only for this comparison)
    def reduce_pass():
        acc = torch.zeros(P6, device=dev, dtype=torch.float64)
        for lo in range(0, P6, 1000):
            blk = big[lo:lo + 1000]
            acc[lo:lo + 1000] = (blk - y).double().square_().mean(dim=1)
        return acc
    ms_red, _ = device_time(reduce_pass, 2, flush)
    res["C6"]["caller_side_loss_reduction_ms"] = min(ms_red)
    res["C6"]["materialise_then_reduce_ms"] = res["C6"]["ms"] + min(ms_red)
    ms, clk = device_time(lambda: pop.eval_loss(X.T, y), reps, flush, local_rank)
    emit("C6-loss", "fused MSE per tree on the C6 workload (no result matrix written); fraction vs the READ bytes only",
         ms, clk, pop.info["n_nodes"] * N6, P6 * N6 * (F * 4))
    del pop, big, out
    torch.cuda.empty_cache()
    return res


def run_c4(rank, local_rank, world, flush, peak, reps=2):
    """BASELINE.json configs[3] strong-scaled: 10k depth-12 trees, 10 features, 2^20 samples split
    into `world` contiguous column blocks.  Reports the kernel time (max over ranks), the fused
    evaluate + gather-to-root over NVLink peer memory, the NCCL all-gather of the row blocks, and
    whether the gathered rows equal an unsharded evaluation bit for bit (a sample of trees)."""
    import torch
    import torch.distributed as dist
    import dexb200
    from dexb200 import device as D, sharded, treegen
    dev = torch.device(f"cuda:{local_rank}")
    P, F, N = 10000, 10, 1 << 20
    ops = dexb200.OperatorEnum(treegen.OPSET_A)
    nodes, offsets = treegen.gen_population(P, 12, 2, 4, F, seed=0)
    ctx = D.Context.get(local_rank)
    pop = D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)
    Xfull = torch.randn((N, F), device=dev, generator=torch.Generator(device=dev).manual_seed(99))   # same on every rank
    s, e = sharded.column_block(N, rank, world)
    Xl = Xfull[s:e]
    nl = e - s
    ok = torch.empty(P, device=dev, dtype=torch.uint8)

    def maxed(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(f, n):
        best = 1e30
        for _ in range(n + 1):
            flush.fill_(1)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            f()
            b.record()
            torch.cuda.synchronize()
            best = min(best, maxed(a.elapsed_time(b)))
        return best

    out = torch.empty((P, nl), device=dev)
    t_kernel = timed(lambda: pop.eval(Xl.T, out=out, ok=ok), reps)
    nodeops = float(pop.info["n_nodes"]) * N
    alg = float(P) * N * (F * 4 + 4)
    res = {"workload": "BASELINE.json configs[3]: 10k depth-12 trees, 10 features, Float32, 2^20 samples, "
                       f"sample-sharded over {world} GPU(s) (strong scaling)",
           "n_gpus": world, "samples_per_gpu": nl, "kernel_ms": t_kernel, "node_ops_per_s": nodeops / (t_kernel * 1e-3),
           "roofline_frac_per_gpu": (alg / world) / (peak * 1e9) / (t_kernel * 1e-3),
           "result_bytes_per_gpu": P * nl * 4, "complete_fraction": float(ok.float().mean())}
    # rows of a tree sample from an unsharded evaluation on this rank (the N = 1 answer)
    sel = np.arange(0, P, 157)
    parts = [nodes[offsets[t]:offsets[t + 1]] for t in sel]
    soff = np.zeros(len(sel) + 1, dtype=np.int64)
    np.cumsum([len(q) for q in parts], out=soff[1:])
    spop = D.Population(None, ops, np.float32, wire=(np.concatenate(parts), soff), ctx=ctx)
    ref_rows, ref_ok = spop.eval(Xfull.T)
    keep = ref_ok.bool()     # rows of incomplete trees are unspecified (early exit), as in the reference
    res["rows_compared"] = int(keep.sum())
    same = lambda u, v: bool(torch.equal(u[keep], v[keep]))
    res["local_block_equals_unsharded_rows"] = same(out[torch.as_tensor(sel, device=dev)], ref_rows[:, s:e])
    if world > 1:
        del out
        fg = sharded.FusedGather(ctx, P, N, torch.float32, root=0)
        res["fused_peer_gather_ms"] = timed(lambda: fg.eval(pop, Xl.T), reps)
        full, okr = fg.eval(pop, Xl.T)
        if rank == 0:
            res["fused_gather_equals_unsharded_rows"] = same(full[torch.as_tensor(sel, device=dev)], ref_rows)
            extra = res["fused_peer_gather_ms"] - t_kernel
            res["gather_exposed_ms"] = extra          # what the gather adds on top of the kernel
            res["gather_bytes_into_root"] = P * (N - nl) * 4
            res["gather_effective_GBps"] = (P * (N - nl) * 4) / 1e9 / (res["fused_peer_gather_ms"] * 1e-3)
        del full
        fg.close()
        torch.cuda.empty_cache()
        # NCCL: all-gather of the (P x N/R) row blocks, layout (R, P, N/R)
        out = torch.empty((P, nl), device=dev)
        flat = torch.empty((world, P, nl), device=dev) if N % world == 0 else None
        if flat is not None:
            def nccl():
                pop.eval(Xl.T, out=out, ok=ok)
                dist.all_gather_into_tensor(flat.view(world * P, nl), out)
                sharded.allreduce_ok(ok)
            res["nccl_allgather_ms"] = timed(nccl, reps)
            r = int(np.random.default_rng(0).integers(world))
            rs, re_ = sharded.column_block(N, r, world)
            res["nccl_gather_equals_unsharded_rows"] = same(flat[r][torch.as_tensor(sel, device=dev)], ref_rows[:, rs:re_])
            del flat
        res["kernel_share_of_fused_gather"] = t_kernel / res["fused_peer_gather_ms"]
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dexb200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` and `c4` legs (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    numa_cpus = bind_to_gpu_numa(local_rank)
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

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
    ms_step_max = max_over_ranks(ms_step)
    value = node_ops_per_step * world / (ms_step_max * 1e-3)

    # ---- end to end through the host-buffer C-ABI entry point ---------------------------
    out_host = torch.empty((N_TREES, NSAMPLES), dtype=torch.float32).pin_memory()
    ok_host = torch.empty(N_TREES, dtype=torch.uint8).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    def host_timed(step):
        for _ in range(4):      # untimed: the first passes over freshly pinned host buffers run slower
            step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step()
        torch.cuda.synchronize()
        return max_over_ranks((time.perf_counter() - t0) / e2e_steps)

    e2e_s = host_timed(lambda: pop.eval_host(X_host, out_host, ok_host))    # H2D X, kernels, D2H results + flags, sync
    e2e_value = node_ops_per_step * world / e2e_s

    # ---- ... not transferring the rows of incomplete trees (DEX_EVAL_SKIP_INCOMPLETE: unspecified under
    # early exit — the reference returns at the first non-finite node — and D2H is what e2e is bound by)
    skip_s = host_timed(lambda: pop.eval_host(X_host, out_host, ok_host, skip_incomplete=True))
    skip_rows = int(ok_host.sum().item())

    # ---- ... including the flattening + upload of the population (callers whose trees change every
    # generation): wire arrays -> dex_population_pack -> dex_eval_host -> destroy, per step
    def pack_step():
        p2 = D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)
        p2.eval_host(X_host, out_host, ok_host)
        del p2

    pack_s = host_timed(pack_step)
    t0 = time.perf_counter()
    for _ in range(5):
        p2 = D.Population(None, ops, np.float32, wire=(nodes, offsets), ctx=ctx)
        del p2
    pack_only_ms = (time.perf_counter() - t0) / 5 * 1e3

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

    loss_s = host_timed(loss_step)
    e2e_loss_value = node_ops_per_step * world / loss_s

    peak, peak_src = peaks()
    complete_fraction = float(ok.float().mean().item())
    del out, out_host
    torch.cuda.empty_cache()
    configs = c4 = None
    if not args.no_configs:
        if rank == 0:
            configs = run_configs(local_rank, flush, peak)
        barrier()
        c4 = run_c4(rank, local_rank, world, flush, peak)

    if rank == 0:
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
        clocks = clk.summary()
        for mname in ("r02_eval_kernel_ncu_metrics.json", "r01_eval_kernel_ncu_metrics.json"):
            mpath = os.path.join(ROOT, "profiles", mname)
            if os.path.exists(mpath) and clocks.get("sm_mhz"):
                try:
                    inst = float(json.load(open(mpath))["smsp__inst_executed.sum"]["value"])
                    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
                    issue_peak = sm_count * 4 * clocks["sm_mhz"] * 1e6
                    issue = {"bound": "issue", "warp_instructions_per_launch": inst,
                             "peak_warp_instructions_per_s": issue_peak, "ms_at_peak": inst / issue_peak * 1e3,
                             "frac": (inst / issue_peak * 1e3) / ms_step,
                             "source": f"smsp__inst_executed.sum of profiles/{mname}"}
                    break
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
                           "complete_fraction": complete_fraction},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "read_only_frac": (achieved * (NFEATURES * 4) / BYTES_PER_UNIT) / peak,
                         "kernel": "dex::eval_kernel<float, 2, true, false, false, 256, true> (GX form: 64 registers, 4 CTAs per SM, feature rows beyond six read through L1)",
                         "note": "an ALGORITHMIC-bytes ratio: X is L2-resident, real DRAM traffic (`traffic`) is the "
                                 "result writes only; the kernel is issue-bound (issue_roofline)"},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(X_host.numel() * 4),
                    "d2h_bytes_per_step": int(N_TREES * NSAMPLES * 4 + ok_host.numel()),
                    "ms_per_step": e2e_s * 1e3, "entry": "dex_eval_host (pinned host buffers)"},
            "e2e_with_pack": {"value": node_ops_per_step * world / pack_s, "unit": UNIT, "ms_per_step": pack_s * 1e3,
                              "pack_and_upload_ms": pack_only_ms,
                              "h2d_bytes_per_step": int(X_host.numel() * 4 + pop.info["n_instructions"] * 32),
                              "d2h_bytes_per_step": int(N_TREES * NSAMPLES * 4 + ok_host.numel()),
                              "entry": "dex_population_pack (flatten + upload) + dex_eval_host + dex_population_destroy, every step"},
            "e2e_skip_incomplete": {"value": node_ops_per_step * world / skip_s, "unit": UNIT, "ms_per_step": skip_s * 1e3,
                                    "h2d_bytes_per_step": int(X_host.numel() * 4),
                                    "d2h_bytes_per_step": int(skip_rows * NSAMPLES * 4 + 2 * ok_host.numel()),
                                    "rows_transferred": skip_rows,
                                    "entry": "dex_eval_host with DEX_EVAL_SKIP_INCOMPLETE: rows of trees whose flag is 0 "
                                             "(unspecified under early exit, as in the reference) stay on the device"},
            "e2e_fused_loss": {"value": e2e_loss_value, "unit": UNIT,
                               "h2d_bytes_per_step": int(X_host.numel() * 4 + y_host.numel() * 4),
                               "d2h_bytes_per_step": int(loss_host.numel() * 8 + ok_host.numel()),
                               "ms_per_step": loss_s * 1e3,
                               "entry": "dex_eval_loss (per-tree MSE; the P x N results never leave the SM)"},
            "issue_roofline": issue,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "numa_bound_cpus": numa_cpus,
            "wall_s_timed_region": t_wall,
            "ms_per_step_min": min(ms), "ms_per_step_median": statistics.median(ms),
        }
        if configs is not None:
            line["configs"] = configs
        if c4 is not None:
            line["c4"] = c4
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(nodes, offsets, ops.opcodes, X_host.numpy())
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
