#!/usr/bin/env python
"""Benchmark of the MLMC estimation hot path (contract: see the task's bench.py section).

Workload (BASELINE.json configs[1]): 3-level SynthSimulation fine/coarse pairs, Legendre n_moments = 50,
1e7 samples per level, fused moments + mean / variance of the level differences, followed by the level-variance
regression and the n_samples allocation.  One "step" = one complete estimate over all levels.
Metric: level sample.moments / s = sum_l N_l * M * R / time.

  python bench.py [--gpus N] [--steps K] [--warmup W]           our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference [...]                        the reference's CPU algorithm (oracle port) on host cores

``value``  : inputs resident in HBM, all launches of one step captured in a CUDA graph, timed with CUDA events.
``e2e``    : the same estimate through the public API (``Estimate.estimate_diff_vars_regression`` + allocation)
             on a pinned-host ``Memory`` storage: H2D copies of every level and the D2H of the result are inside
             the timed region, every step.
Weak scaling under torchrun: every rank holds 1e7 samples per level (global N = world * 1e7 per level), level sums
are combined by one NCCL all-reduce per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_LEVELS = 3
N_PER_LEVEL = 10_000_000
N_MOMENTS = 50
STEP_RANGE = (0.5, 0.005)
TARGET_VAR = 1e-5
DOMAIN_Q = (1e-4, 1 - 1e-4)
METRIC = "level sample-moments/s (fused moments + mean/var of level differences)"
UNIT = "sample-moments/s"


def workload_name(n_per_level):
    return "cfg2: 3-level SynthSimulation pairs, Legendre R=50, %.0e samples/level, mean+var+regression+n_samples" \
        % n_per_level


def level_steps():
    return [STEP_RANGE[0] ** (1 - l / (N_LEVELS - 1)) * STEP_RANGE[1] ** (l / (N_LEVELS - 1)) for l in range(N_LEVELS)]


def n_ops_of(step):
    return (1 / step) ** 2 * np.log(max(1 / step, 2.0))


def domain():
    import scipy.stats
    return tuple(float(v) for v in scipy.stats.norm.ppf(DOMAIN_Q))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampling of SM clocks / throttle reasons DURING the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, sm_max, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                sm_max.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(sm_max) if sm_max else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(args):
    """Oracle (reference algorithm in NumPy) on one shard of every level -> per-level partial sums."""
    seed, n_rows, chunk_rows = args
    from oracle import mlmc_oracle as orc
    rng = np.random.default_rng(seed)
    steps = level_steps()
    basis = orc.Basis("legendre", N_MOMENTS, domain())
    levels = [orc.synth_level_rows(rng.normal(size=n_rows), steps[l], steps[l - 1] if l else None)
              for l in range(N_LEVELS)]
    t0 = time.perf_counter()
    est = orc.estimate_moments(levels, basis, chunk_rows=chunk_rows)
    reg = orc.regress_level_variances(est.l_vars, steps)
    orc.n_samples_for_target_variance(TARGET_VAR, reg, [n_ops_of(h) for h in steps], N_LEVELS)
    return time.perf_counter() - t0, int(est.n_samples.sum())


def cpu_reference_run(n_rows_per_level, n_procs, chunk_rows=65536, pool=None):
    """Time the oracle port on ``n_procs`` host processes, each with its own shard of ``n_rows_per_level`` rows per
    level (the reference itself is single-threaded NumPy; sharding over processes is the most generous way to let
    it use the host's cores).  Returns (sample-moments/s, slowest worker's seconds, wall seconds)."""
    import multiprocessing as mp
    jobs = [(1234 + 1000 * p, n_rows_per_level, chunk_rows) for p in range(n_procs)]
    t0 = time.perf_counter()
    if n_procs == 1:
        results = [_cpu_worker(jobs[0])]
    elif pool is not None:
        results = pool.map(_cpu_worker, jobs, chunksize=1)
    else:
        with mp.get_context("spawn").Pool(n_procs) as own:
            results = own.map(_cpu_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    inner = max(r[0] for r in results)
    units = N_LEVELS * n_rows_per_level * n_procs * N_MOMENTS
    return units / inner, inner, wall


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_procs = max(1, min(cores, 64))
    # per level per process: a bounded sample of the 1e7-row workload, ~2.5 s of NumPy per process and step, shrunk
    # for long runs so that the whole arm stays within a few minutes
    rows = 500_000 if args.steps <= 20 else max(50_000, 10_000_000 // args.steps)
    import multiprocessing as mp
    values, ms = [], []
    with mp.get_context("spawn").Pool(n_procs) as pool:
        for _ in range(max(args.warmup, 0)):
            cpu_reference_run(20_000, n_procs, pool=pool)
        for _ in range(args.steps):
            v, inner, _wall = cpu_reference_run(rows, n_procs, pool=pool)
            values.append(v)
            ms.append(inner * 1e3)
    value = float(np.median(values))
    sample = "%d processes x %d rows/level x %d levels of the cfg2 workload (oracle port, NumPy)" % (n_procs, rows, N_LEVELS)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.median(ms)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(N_PER_LEVEL), "sampled_rows_per_level": rows * n_procs},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def make_levels_on_device(torch, device, n_rows, seed_base):
    """SynthSimulation.sample_fn rows [n, 2, 1] per level, generated on the device (synthetic data)."""
    steps = level_steps()
    levels = []
    for l in range(N_LEVELS):
        gen = torch.Generator(device=device).manual_seed(seed_base + 1000 * l)
        x = torch.randn(n_rows, generator=gen, device=device, dtype=torch.float64)
        root = torch.sqrt(1e-4 + x.abs())
        fine = x + steps[l] * root
        if l > 0:
            levels.append(torch.stack([fine, x + steps[l - 1] * root], dim=1).unsqueeze(2).contiguous())
        else:                       # level 0 has no coarse simulation: rows [n, 1, 1] (the zero row is not stored)
            levels.append(fine.reshape(-1, 1, 1).contiguous())
        del x, root, fine
    return levels


def run_gpu_arm(args):
    import torch
    import torch.distributed as td
    from mlmc_b200 import _native as nat, dist as mdist
    from mlmc_b200.moments import Legendre
    from mlmc_b200.sample_storage import Memory
    from mlmc_b200.quantity.quantity import make_root_quantity
    from mlmc_b200.quantity.quantity_spec import QuantitySpec
    from mlmc_b200.estimator import Estimate, estimate_n_samples_for_target_variance

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    rank, world, local = mdist.init_from_env()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    nat.load()
    n_rows = args.samples_per_level
    steps = level_steps()
    n_ops = [n_ops_of(h) for h in steps]
    moments_fn = Legendre(N_MOMENTS, domain())
    basis = moments_fn.basis_struct()
    levels = make_levels_on_device(torch, device, n_rows, 1234 + rank)
    units_per_rank = N_LEVELS * n_rows * N_MOMENTS
    bytes_per_rank = sum(lv.numel() * 8 for lv in levels)          # 8 B per level-0 sample, 16 B per (fine, coarse) pair

    # ---------------- device-resident step: one CUDA graph = zero, 3 x (moments + reduce), finalize ----------
    acc = nat.LevelAccumulator(N_LEVELS, N_MOMENTS, device)
    views = []
    for l, rows in enumerate(levels):
        views.append(rows.permute(2, 0, 1))
    result = {}

    # the levels are independent until the finalize: each runs on its own stream (a fork / join inside the graph), so
    # the ramp and tail of one level's one-wave launch overlap the next level's kernel
    level_streams = [torch.cuda.Stream(device) for _ in range(N_LEVELS - 1)]
    concurrent_levels = os.environ.get("BENCH_CONCURRENT_LEVELS", "1") != "0"

    # N > 1: the sum of the level accumulators over the ranks rides in the finalize launch (NVLink peer memory,
    # mlmcb200_allreduce_finalize_levels); checked against the NCCL all-reduce once, NCCL stays the fallback
    use_peer = [False]

    def enqueue_step():
        acc.acc.zero_()
        main = torch.cuda.current_stream()
        if not concurrent_levels:
            for l in range(N_LEVELS):
                nat.moments_accumulate(basis, views[l], acc.level(l))
        else:
            fork = torch.cuda.Event()
            fork.record(main)
            joins = []
            for l in range(N_LEVELS):
                st = main if l == 0 else level_streams[l - 1]
                if st is not main:
                    st.wait_event(fork)
                with torch.cuda.stream(st):
                    nat.moments_accumulate(basis, views[l], acc.level(l))
                    if st is not main:
                        ev = torch.cuda.Event()
                        ev.record(st)
                        joins.append(ev)
            for ev in joins:
                main.wait_event(ev)
        if world > 1 and not use_peer[0]:
            td.all_reduce(acc.acc)
        result.update(acc.finalize(peer=mdist.peer_state(acc.acc.numel()) if use_peer[0] else None))

    if world > 1 and os.environ.get("BENCH_PEER_REDUCE", "1") != "0" and mdist.enable_peer_reduce():
        enqueue_step()
        torch.cuda.synchronize()
        want = result["packed"].clone()
        use_peer[0] = True
        enqueue_step()
        torch.cuda.synchronize()
        # (more than two ranks: NCCL adds in its own order, the peer kernel in rank order -- rounding-level differences
        # in the sums, amplified by the cancellation in the variances)
        good = torch.tensor([int(torch.allclose(result["packed"], want, rtol=1e-8, atol=1e-14, equal_nan=False)
                                 and not mdist.peer_error())], device=device)
        td.all_reduce(good, op=td.ReduceOp.MIN)
        use_peer[0] = bool(good.item())
        if not use_peer[0]:
            mdist.disable_peer_reduce()                  # the public API (e2e pass) stays on NCCL as well
    side = torch.cuda.Stream(device)
    with torch.cuda.stream(side):
        enqueue_step()                                   # warm up allocations outside the capture
        torch.cuda.synchronize()
        graph = None
        # N > 1: plain stream launches by default; BENCH_GRAPH_MULTI=1 captures the all-reduce with the kernels (NCCL
        # supports stream capture; measured 0.653 vs 0.659 ms per step at N = 2 -- not worth a capture failure mode)
        # (with the peer-memory reduce the step holds no NCCL call at all and is captured like the single-GPU step)
        if world == 1 or use_peer[0] or os.environ.get("BENCH_GRAPH_MULTI", "0") == "1":
            if world > 1:
                enqueue_step()                           # a second eager step: NCCL sets up its channels lazily
                torch.cuda.synchronize()
                td.barrier()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                enqueue_step()
    launches_before = nat.launch_count
    launches_per_step = 2 * N_LEVELS + 1                 # our kernels: (moments + reduce) per level, finalize

    def device_step():
        if graph is not None:
            graph.replay()
        else:
            enqueue_step()

    def host_tail():
        """level-variance regression + allocation on the [L, R] result (tiny host math, as in the reference)."""
        packed = result["packed"].cpu().numpy()
        l_vars = packed[N_LEVELS:2 * N_LEVELS]
        reg = Estimate(None, None)._all_moments_variance_regression(l_vars, np.array(steps))
        return estimate_n_samples_for_target_variance(TARGET_VAR, reg, n_ops, N_LEVELS)

    for _ in range(max(args.warmup, 3)):
        device_step()
    n_estimated = host_tail()
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    # per-launch time of the dominant kernel (level 1: fine + coarse) with CUDA events on the launching stream:
    # 10 back-to-back launches per measurement (each = moments kernel + its 5 us partial reduction), best of 3
    k_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    cur = torch.cuda.current_stream()
    kernel_ms = []
    for _ in range(3):
        k_ev[0].record(cur)
        for _rep in range(10):
            nat.moments_accumulate(basis, views[1], acc.level(1))
        k_ev[1].record(cur)
        torch.cuda.synchronize()
        kernel_ms.append(k_ev[0].elapsed_time(k_ev[1]) / 10)
    kernel_ms = float(np.min(kernel_ms))

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        td.all_reduce(ms_total, op=td.ReduceOp.MAX)
    ms_per_step = float(ms_total.item()) / args.steps
    value = units_per_rank * world / (ms_per_step * 1e-3)

    # ---------------- end to end through the public API, host buffers ----------------
    host_levels = [lv.cpu() for lv in levels]
    spec = [QuantitySpec(name="v", unit="", shape=(1, 1), times=[0.0], locations=["0"])]
    storage = Memory.from_arrays(host_levels, level_parameters=[[h] for h in steps], n_ops=n_ops, result_format=spec)
    storage.rows_are_local_shard = True      # weak scaling: every rank owns n_rows samples per level
    storage.resident_fraction = 0.0          # never keep a device copy: every step streams host -> HBM
    storage.device_chunk_bytes = 32 << 20    # 32 MB chunks, copy stream overlapped with the kernels
    del host_levels
    value_q = make_root_quantity(storage, spec)["v"][0.0]["0"][0, 0]
    estimator = Estimate(value_q, storage, moments_fn)
    d2h_bytes = (2 * N_LEVELS + 2) * N_MOMENTS * 8 + N_LEVELS * 16

    def e2e_step():
        storage.drop_device_copies()                                  # inputs start on the HOST every step
        variances, ops = estimator.estimate_diff_vars_regression(None)
        return estimate_n_samples_for_target_variance(TARGET_VAR, variances, ops, N_LEVELS)

    for _ in range(2):
        n_est_e2e = e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        n_est_e2e = e2e_step()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        td.all_reduce(e2e_s, op=td.ReduceOp.MAX)
    e2e_value = units_per_rank * world / (float(e2e_s.item()) / e2e_steps)
    clocks = sampler.stop() if rank == 0 else None

    # What bounds e2e: the pinned host -> device copy of the step's inputs, measured with ALL ranks copying at once
    # (the ranks of one box share PCIe switches and host memory).  Slowest rank's rate, like the e2e time.
    from mlmc_b200.sample_storage import stream_levels
    segments = [(l, storage._host_tensor(l)) for l in range(N_LEVELS)]
    h2d_ms = []
    for rep in range(3):
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        t0 = time.perf_counter()
        for _level, _rows in stream_levels(segments, device, storage.device_chunk_bytes):
            pass                                   # the same double-buffered chunk copies, no kernels
        torch.cuda.synchronize()
        h2d_ms.append((time.perf_counter() - t0) * 1e3)
    h2d_t = torch.tensor([min(h2d_ms[1:])], dtype=torch.float64, device=device)
    if world > 1:
        td.all_reduce(h2d_t, op=td.ReduceOp.MAX)
    h2d_gbs_slowest = bytes_per_rank / (float(h2d_t.item()) * 1e-3) / 1e9

    if rank != 0:
        if world > 1:
            td.barrier()
        return

    # ---------------- roofline of the dominant kernel + CPU baseline (rank 0, N = 1 only for the CPU) ----------
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    dfma_peak = nat.fp64_peak(0) / 1e12
    # algorithmic work of one launch over a level with a coarse part: 16 B per sample and 7 FP64 instructions per
    # sample-moment (2 x (DMUL + DFMA) recurrence, f - c, sum, FMA square), counted as 14 flop -- one FMA-equivalent
    # per instruction slot, the way the DFMA peak is counted
    alg_bytes = n_rows * 16.0
    alg_flop = n_rows * N_MOMENTS * 14.0
    achieved_tflops = alg_flop / (kernel_ms * 1e-3) / 1e12
    traffic = None
    try:                                   # dram__bytes_read + dram__bytes_write of this kernel, one ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "r1_roofline_traffic.json")) as f:
            cap = json.load(f)["moments_acc_kernel_coarse_R50"]
        if cap["samples"] == n_rows:
            traffic = cap["dram_bytes_read"] + cap["dram_bytes_write"]
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"kernel": "moments_acc_kernel<LEGENDRE, coarse> (level with fine+coarse, R=50)",
                "bound": "fp64", "achieved": achieved_tflops, "peak": dfma_peak, "unit": "TFLOP/s",
                "frac": achieved_tflops / dfma_peak, "peak_source": "DFMA micro-benchmark measured in this run",
                "traffic": traffic, "algorithmic_bytes": alg_bytes, "kernel_ms": kernel_ms,
                "hbm": {"achieved": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src},
                "note": "R=50 in fp64 is FP64-pipe bound (0.75*R flop/B >> ridge 5.6 flop/B); both terms reported"}
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        n_procs = max(1, min(os.cpu_count() or 1, 64))
        rows_cpu = 1_000_000                 # ~5 s of NumPy per process; ~10-20 s wall with all cores busy
        v, inner, _wall = cpu_reference_run(rows_cpu, n_procs)
        v1, inner1, _ = cpu_reference_run(rows_cpu, 1)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": n_procs, "kind": "port",
                        "sample": "%d processes x %d rows/level x %d levels (oracle port of the reference's NumPy "
                                  "algorithm, 65536-row chunks)" % (n_procs, rows_cpu, N_LEVELS),
                        "single_process_value": v1, "seconds": inner}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n_rows), "levels": N_LEVELS, "samples_per_level_per_gpu": n_rows,
                       "n_moments": N_MOMENTS, "l2_policy": "inputs (%.0f MB per step per GPU) exceed the 126 MB L2"
                       % (bytes_per_rank / 1e6), "parallelism": "sample-sharded x%d, one all-reduce of level sums" % world,
                       "reduce": ("none" if world == 1 else "NVLink peer memory, fused with the finalize launch"
                                  if use_peer[0] else "NCCL all-reduce"),
                       "launch": "CUDA graph" if graph is not None else "stream"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(bytes_per_rank * world),
                    "d2h_bytes_per_step": int(d2h_bytes * world), "steps": e2e_steps,
                    "api": "Estimate.estimate_diff_vars_regression + estimate_n_samples_for_target_variance on a "
                           "pinned-host Memory storage",
                    "ms_per_step": float(e2e_s.item()) / e2e_steps * 1e3,
                    "h2d_only_ms": float(h2d_t.item()),
                    "h2d_gbs_per_gpu_all_ranks_copying": h2d_gbs_slowest,
                    "bound": "PCIe / host memory: the bare pinned copy of the same bytes, all ranks at once"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "n_estimated": [int(v) for v in n_estimated], "n_estimated_e2e": [int(v) for v in n_est_e2e],
            "launch_count_native": nat.launch_count - launches_before}
    emit(line)
    if world > 1:
        td.barrier()


_JSON_OUT = None


def emit(line):
    """Print the ONE JSON line on the real stdout."""
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # Libraries (NCCL prints "NCCL version ..." when a communicator is created) may write to fd 1: keep a private
    # handle on the real stdout for the JSON line and point fd 1 at stderr for everything else.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples-per-level", type=int, default=N_PER_LEVEL)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
