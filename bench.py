#!/usr/bin/env python
"""Benchmark of the MLMC estimation hot path (contract: see the task's bench.py section).

Workload (BASELINE.json configs[1]): 3-level SynthSimulation fine/coarse pairs, Legendre n_moments = 50,
1e7 samples per level, fused moments + mean / variance of the level differences, followed by the level-variance
regression and the n_samples allocation.  One "step" = one complete estimate over all levels.
Metric: level sample.moments / s = sum_l N_l * M * R / time.

  python bench.py [--gpus N] [--steps K] [--warmup W]           our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference [...]                        the reference's CPU algorithm (oracle port) on host cores

``value``  : inputs resident in HBM, all launches of one step captured in a CUDA graph, timed with CUDA events; the
             D2H of the result and the host regression / allocation of every step are inside the timed region.
``e2e``    : the same estimate through the public API (``Estimate.estimate_diff_vars_regression`` + allocation)
             on a pinned-host ``Memory`` storage: H2D copies of every level and the D2H of the result are inside
             the timed region, every step.
``configs``: the other BASELINE.json configs in the same JSON line -- cfg3 (covariance, Legendre 100, 1e9 samples in
             total = STRONG scaling over the ranks, DMMA roofline), and at N = 1 cfg4 (max-ent fit ms), cfg1, cfg5.
Weak scaling of the headline under torchrun: every rank holds 1e7 samples per level (global N = world * 1e7 per level),
level sums are combined inside the finalize launch over NVLink peer memory (NCCL all-reduce as the fallback).
Before anything is timed the golden case of the unmodified reference (tests/golden/estimates.npz) is estimated, sharded
over the ranks, and compared with the stored reference outputs; a mismatch aborts the run.
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


def synth_pairs_on_device(torch, device, n_rows, h_fine, h_coarse, seed, piece=25_000_000, lognormal=False):
    """One level of (fine, coarse) rows [n, 2, 1] generated piecewise on the device (bounded temporaries)."""
    rows = torch.empty((n_rows, 2, 1), dtype=torch.float64, device=device)
    gen = torch.Generator(device=device).manual_seed(seed)
    for lo in range(0, n_rows, piece):
        hi = min(n_rows, lo + piece)
        x = torch.randn(hi - lo, generator=gen, device=device, dtype=torch.float64)
        if lognormal:
            x = torch.exp(x)
        root = torch.sqrt(1e-4 + x.abs())
        rows[lo:hi, 0, 0] = x + h_fine * root
        if h_coarse is None:
            rows[lo:hi, 1, 0] = 0.0
        else:
            rows[lo:hi, 1, 0] = x + h_coarse * root
        del x, root
    return rows


class Ctx:
    """What the legs share: torch / native modules, rank layout, helpers for device timing over all ranks."""

    def __init__(self, args):
        import torch
        import torch.distributed as td
        from mlmc_b200 import _native as nat, dist as mdist
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
        self.torch, self.td, self.nat, self.mdist, self.args = torch, td, nat, mdist, args
        self.rank, self.world, self.local = mdist.init_from_env()
        torch.cuda.set_device(self.local)
        self.device = torch.device("cuda", self.local)
        nat.load()
        self.peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                self.peaks = json.load(f)
        except OSError:
            pass

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.td.barrier()

    def max_over_ranks(self, value):
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.td.all_reduce(t, op=self.td.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, reps, warmup=1):
        """ms per call of ``fn`` (enqueues work on the current stream): CUDA events around ``reps`` calls, barrier +
        synchronize on both sides, max over ranks."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)) / reps

    def timed_wall(self, fn, reps, warmup=1):
        """The same for calls that end with their own device -> host copy (public API): wall clock, max over ranks."""
        res = None
        for _ in range(warmup):
            res = fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            res = fn()
        self.torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / reps
        return self.max_over_ranks(ms), res


def scalar_quantity(levels, steps, n_ops=None):
    from mlmc_b200.sample_storage import Memory
    from mlmc_b200.quantity.quantity import make_root_quantity
    from mlmc_b200.quantity.quantity_spec import QuantitySpec
    spec = [QuantitySpec(name="v", unit="", shape=(1, 1), times=[0.0], locations=["0"])]
    storage = Memory.from_arrays(levels, level_parameters=[[h] for h in steps], n_ops=n_ops, result_format=spec)
    return storage, make_root_quantity(storage, spec)["v"][0.0]["0"][0, 0]


def max_rel(got, want):
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    scale = float(np.max(np.abs(want))) + 1e-300
    return float(np.max(np.abs(got - want) / (np.abs(want) + 1e-12 * scale)))


def startup_parity_check(ctx):
    """Before anything is timed: the golden 3-level case of tests/golden/estimates.npz (inputs + outputs of the UNMODIFIED
    reference) estimated through the public API, rows sharded over the ranks when N > 1, against the stored reference
    results: level means / variances rel <= 1e-10 (scaled per level), covariance <= 1e-8, counts exact."""
    from mlmc_b200.moments import Legendre
    from mlmc_b200.quantity import quantity_estimate as qe
    g = np.load(os.path.join(ROOT, "tests", "golden", "estimates.npz"))
    levels = [g["A_rows%d" % l] for l in range(3)]
    _storage, value = scalar_quantity(levels, [float(h) for h in g["A_steps"]])
    domain = tuple(g["A_domain"])
    qm = qe.estimate_mean(qe.moments(value, Legendre(12, domain)))
    cm = qe.estimate_mean(qe.covariance(value, Legendre(8, domain)))
    lin = qe.estimate_mean(qe.covariance(value, Legendre(8, domain)), variance=False)

    def per_level(got, want):
        scale = np.max(np.abs(want), axis=1, keepdims=True)
        return float(np.max(np.abs(got - want) / (np.abs(want) + 1e-3 * scale)))
    out = {"l_means": per_level(qm.l_means, g["A_leg_l_means"]), "l_vars": per_level(qm.l_vars, g["A_leg_l_vars"]),
           "cov_mean": per_level(cm.mean, g["A_cov_mean"]), "cov_var": per_level(cm.var, g["A_cov_var"]),
           "cov_mean_linearised": per_level(lin.mean, g["A_cov_mean"]),
           "counts_exact": bool(np.array_equal(qm.n_samples, g["A_leg_n"]) and
                                np.array_equal(qm.n_rm_samples, g["A_leg_n_rm"]) and
                                np.array_equal(cm.n_samples, g["A_leg_n"]))}
    ok = (out["counts_exact"] and out["l_means"] <= 1e-10 and out["l_vars"] <= 1e-10 and out["cov_mean"] <= 1e-8 and
          out["cov_var"] <= 1e-8 and out["cov_mean_linearised"] <= 1e-8 and qm.mean[0] == 1.0 and qm.var[0] == 0.0)
    if not ok:
        raise SystemExit("bench.py: start-up parity check against the reference's golden outputs FAILED on rank %d: %s"
                         % (ctx.rank, out))
    out["against"] = "tests/golden/estimates.npz (outputs of the unmodified reference), rows sharded over %d rank(s)" \
        % ctx.world
    return out


def strict_peer_check(ctx, basis, views, acc):
    """The fused peer-memory reduce must reproduce, BIT FOR BIT, the rank-ordered fp64 sum of the per-rank accumulators
    (gathered with NCCL) and the finalize of that sum.  All ranks must agree, else the bench stays on NCCL."""
    torch, td, nat, mdist = ctx.torch, ctx.td, ctx.nat, ctx.mdist
    acc.acc.zero_()
    for l in range(N_LEVELS):
        nat.moments_accumulate(basis, views[l], acc.level(l))
    local = acc.acc.clone()
    gathered = [torch.empty_like(local) for _ in range(ctx.world)]
    td.all_gather(gathered, local)
    want = torch.zeros_like(local)
    for part in gathered:                                    # rank order, like the kernel
        want = want + part
    ref = nat.LevelAccumulator(N_LEVELS, N_MOMENTS, ctx.device)
    ref.acc.copy_(want)
    want_out = ref.finalize()["packed"].clone()
    got = acc.finalize(peer=mdist.peer_state(acc.acc.numel()))
    torch.cuda.synchronize()
    good = (torch.equal(acc.acc, want) and torch.equal(got["packed"], want_out) and float(got["status"].item()) == 0.0
            and not mdist.peer_error())
    flag = torch.tensor([int(good)], device=ctx.device)
    td.all_reduce(flag, op=td.ReduceOp.MIN)
    return bool(flag.item())


def host_topology(ctx):
    """NUMA placement of this rank and a host-memory copy bandwidth (all ranks copying at once): what the pinned
    host -> device stream of e2e competes for."""
    torch = ctx.torch
    info = {"numa_nodes_online": None, "gpu_numa_node": None, "cpus_allowed": len(os.sched_getaffinity(0))}
    try:
        with open("/sys/devices/system/node/online") as f:
            info["numa_nodes_online"] = f.read().strip()
        props = torch.cuda.get_device_properties(ctx.local)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            info["gpu_numa_node"] = int(f.read().strip())
    except (OSError, ValueError, AttributeError):
        pass
    src = torch.empty(32 << 20, dtype=torch.float64).pin_memory()                 # 256 MB
    dst = torch.empty_like(src)
    threads_before = torch.get_num_threads()
    torch.set_num_threads(max(1, info["cpus_allowed"] // ctx.world))              # torchrun sets OMP_NUM_THREADS=1
    src.fill_(1.0)
    dst.copy_(src)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        dst.copy_(src)
    gbs = 3 * 2 * src.numel() * 8 / (time.perf_counter() - t0) / 1e9            # read + write
    info["host_copy_gbs_per_rank_all_ranks_copying"] = -ctx.max_over_ranks(-gbs)   # the slowest rank's
    info["host_copy_threads"] = torch.get_num_threads()
    torch.set_num_threads(threads_before)
    return info


# ------------------------------------------------------------------------------------------------ cfg2 (headline)
def bench_cfg2(ctx, line_out):
    torch, td, nat, mdist, args = ctx.torch, ctx.td, ctx.nat, ctx.mdist, ctx.args
    from mlmc_b200.moments import Legendre
    from mlmc_b200.estimator import Estimate, estimate_n_samples_for_target_variance
    rank, world, local, device = ctx.rank, ctx.world, ctx.local, ctx.device
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
    views = [rows.permute(2, 0, 1) for rows in levels]
    result = {}
    # the levels are independent until the finalize: each runs on its own stream (a fork / join inside the graph), so
    # the ramp and tail of one level's one-wave launch overlap the next level's kernel
    level_streams = [torch.cuda.Stream(device) for _ in range(N_LEVELS - 1)]
    concurrent_levels = os.environ.get("BENCH_CONCURRENT_LEVELS", "1") != "0"
    # N > 1: the sum of the level accumulators over the ranks rides in the finalize launch (NVLink peer memory,
    # mlmcb200_allreduce_finalize_levels) once it has passed the strict check; NCCL stays the fallback
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

    peer_checked = None
    if world > 1 and os.environ.get("BENCH_PEER_REDUCE", "1") != "0" and mdist.enable_peer_reduce():
        peer_checked = strict_peer_check(ctx, basis, views, acc) and strict_peer_check(ctx, basis, views, acc)
        use_peer[0] = peer_checked
        if not use_peer[0]:
            mdist.disable_peer_reduce()                  # the public API (e2e pass) stays on NCCL as well
    side = torch.cuda.Stream(device)
    with torch.cuda.stream(side):
        enqueue_step()                                   # warm up allocations outside the capture
        torch.cuda.synchronize()
        graph = None
        # N > 1 on NCCL: plain stream launches by default (BENCH_GRAPH_MULTI=1 captures the all-reduce with the kernels);
        # with the peer-memory reduce the step holds no NCCL call at all and is captured like the single-GPU step
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

    # The tail of the workload -- device -> host copy of the [2L+2, R] result, level-variance regression and n_samples
    # allocation (tiny host math, as in the reference) -- is INSIDE the timed region: the copy of step k is enqueued
    # behind its kernels and its host math runs while the kernels of step k + 1 execute.
    packed_shape = (2 * N_LEVELS + 2, N_MOMENTS)
    tail_host = [torch.empty(packed_shape, dtype=torch.float64).pin_memory() for _ in range(2)]
    tail_done = [torch.cuda.Event() for _ in range(2)]
    regress = Estimate(None, None)._all_moments_variance_regression
    steps_arr = np.array(steps)

    def enqueue_tail(k):
        tail_host[k % 2].copy_(result["packed"], non_blocking=True)
        tail_done[k % 2].record()

    def host_tail(k):
        tail_done[k % 2].synchronize()
        l_vars = tail_host[k % 2].numpy()[N_LEVELS:2 * N_LEVELS]
        return estimate_n_samples_for_target_variance(TARGET_VAR, regress(l_vars, steps_arr), n_ops, N_LEVELS)

    def run_steps(n):
        n_est = None
        for k in range(n):
            device_step()
            enqueue_tail(k)
            if k > 0:
                n_est = host_tail(k - 1)
        return host_tail(n - 1) if n > 0 else n_est

    n_estimated = run_steps(max(args.warmup, 3))
    ctx.barrier()

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
    ctx.barrier()
    e0.record()
    n_estimated = run_steps(args.steps)
    e1.record()                                            # after the last step's host tail has finished
    ctx.barrier()
    ms_per_step = ctx.max_over_ranks(e0.elapsed_time(e1)) / args.steps
    value = units_per_rank * world / (ms_per_step * 1e-3)
    peer_errors = bool(use_peer[0] and mdist.peer_error())

    # ---------------- end to end through the public API, host buffers ----------------
    host_levels = [lv.cpu() for lv in levels]
    storage, value_q = scalar_quantity(host_levels, steps, n_ops)
    storage.rows_are_local_shard = True      # weak scaling: every rank owns n_rows samples per level
    storage.resident_fraction = 0.0          # never keep a device copy: every step streams host -> HBM
    storage.device_chunk_bytes = 32 << 20    # 32 MB chunks, copy stream overlapped with the kernels
    del host_levels
    estimator = Estimate(value_q, storage, moments_fn)
    d2h_bytes = (2 * N_LEVELS + 2) * N_MOMENTS * 8 + N_LEVELS * 16

    def e2e_step():
        storage.drop_device_copies()                                  # inputs start on the HOST every step
        variances, ops = estimator.estimate_diff_vars_regression(None)
        return estimate_n_samples_for_target_variance(TARGET_VAR, variances, ops, N_LEVELS)

    e2e_steps = max(1, min(args.steps, 20))       # 20 x ~8 ms: a single host hiccup no longer moves the mean by 10 %
    e2e_ms, n_est_e2e = ctx.timed_wall(e2e_step, e2e_steps, warmup=2)
    e2e_value = units_per_rank * world / (e2e_ms * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    # What bounds e2e: the pinned host -> device copy of the step's inputs, measured with ALL ranks copying at once
    # (the ranks of one box share PCIe switches and host memory).  Slowest rank's rate, like the e2e time.
    from mlmc_b200.sample_storage import stream_levels
    segments = [(l, storage._host_tensor(l)) for l in range(N_LEVELS)]
    h2d_ms = []
    for rep in range(3):
        ctx.barrier()
        t0 = time.perf_counter()
        for _level, _rows in stream_levels(segments, device, storage.device_chunk_bytes):
            pass                                   # the same double-buffered chunk copies, no kernels
        torch.cuda.synchronize()
        h2d_ms.append((time.perf_counter() - t0) * 1e3)
    h2d_t = ctx.max_over_ranks(min(h2d_ms[1:]))
    h2d_gbs_slowest = bytes_per_rank / (h2d_t * 1e-3) / 1e9
    topo = host_topology(ctx)
    gathered_nodes = [None] * world
    if world > 1:
        td.all_gather_object(gathered_nodes, topo["gpu_numa_node"])
    else:
        gathered_nodes = [topo["gpu_numa_node"]]
    topo["gpu_numa_node_per_rank"] = gathered_nodes

    # keep the resident levels for the data-driven density leg (cfg4)
    line_out["_levels"] = levels
    line_out["_steps"] = steps
    if rank != 0:
        return

    # ---------------- roofline of the dominant kernel + CPU baseline (rank 0, N = 1 only for the CPU) ----------
    peaks = ctx.peaks
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    dfma_peak = nat.fp64_peak(0) / 1e12
    # Algorithmic work of one launch over a level with a coarse part:
    # 16 B per sample and, per sample-moment, 12 flop by SURVEY's count (FMA = 2: recurrence 2 x 4, difference, sum,
    # FMA square).  The kernel issues 7 FP64 instructions per sample-moment (monic two-instruction recurrence), so the
    # issue-slot utilisation (7 slots x 2 / DFMA peak) is reported beside it.
    alg_bytes = n_rows * 16.0
    alg_flop = n_rows * N_MOMENTS * 12.0
    achieved_tflops = alg_flop / (kernel_ms * 1e-3) / 1e12
    traffic = None
    for name in ("r2_roofline_traffic.json", "r1_roofline_traffic.json"):
        try:                               # dram__bytes_read + dram__bytes_write of this kernel, one ncu --set full capture
            with open(os.path.join(ROOT, "profiles", name)) as f:
                cap = json.load(f)["moments_acc_kernel_coarse_R50"]
            if cap["samples"] == n_rows:
                traffic = cap["dram_bytes_read"] + cap["dram_bytes_write"]
                break
        except (OSError, KeyError, ValueError):
            pass
    roofline = {"kernel": "moments_acc_kernel<LEGENDRE, coarse> (level with fine+coarse, R=50)",
                "bound": "fp64", "achieved": achieved_tflops, "peak": dfma_peak, "unit": "TFLOP/s",
                "frac": achieved_tflops / dfma_peak, "flop_per_sample_moment": 12,
                "peak_source": "DFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "issue_slot_frac": n_rows * N_MOMENTS * 14.0 / (kernel_ms * 1e-3) / 1e12 / dfma_peak,
                "traffic": traffic, "algorithmic_bytes": alg_bytes, "kernel_ms": kernel_ms,
                "hbm": {"achieved": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src},
                "note": "R=50 in fp64 is FP64-pipe bound (0.75*R flop/B >> ridge 5.6 flop/B); both terms reported; "
                        "frac counts SURVEY 8d's 12 flop per sample-moment, issue_slot_frac the 7 issued FP64 "
                        "instructions x 2"}
    cpu_baseline = {"value": None, "unit": UNIT, "kind": "port",
                    "sample": "not timed at N > 1: see the reference arm (bench.py --impl reference) of this run"}
    if world == 1 and not args.no_cpu_baseline:
        n_procs = max(1, min(os.cpu_count() or 1, 64))
        rows_cpu = 1_000_000                 # ~5 s of NumPy per process; ~10-20 s wall with all cores busy
        v, inner, _wall = cpu_reference_run(rows_cpu, n_procs)
        v1, inner1, _ = cpu_reference_run(rows_cpu, 1)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": n_procs, "kind": "port",
                        "sample": "%d processes x %d rows/level x %d levels (oracle port of the reference's NumPy "
                                  "algorithm, 65536-row chunks)" % (n_procs, rows_cpu, N_LEVELS),
                        "single_process_value": v1, "seconds": inner}
    line_out.update({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n_rows), "levels": N_LEVELS, "samples_per_level_per_gpu": n_rows,
                   "n_moments": N_MOMENTS, "l2_policy": "inputs (%.0f MB per step per GPU) exceed the 126 MB L2"
                   % (bytes_per_rank / 1e6), "parallelism": "sample-sharded x%d, one all-reduce of level sums" % world,
                   "reduce": ("none" if world == 1 else "NVLink peer memory, fused with the finalize launch"
                              if use_peer[0] else "NCCL all-reduce"),
                   "peer_reduce_bitwise_check": peer_checked, "peer_reduce_errors": peer_errors,
                   "launch": "CUDA graph" if graph is not None else "stream",
                   "timed_region": "kernels + D2H of the result + level-variance regression + n_samples allocation "
                                   "(the host tail of step k overlaps the kernels of step k+1)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(bytes_per_rank * world),
                "d2h_bytes_per_step": int(d2h_bytes * world), "steps": e2e_steps,
                "api": "Estimate.estimate_diff_vars_regression + estimate_n_samples_for_target_variance on a "
                       "pinned-host Memory storage",
                "ms_per_step": e2e_ms, "h2d_only_ms": h2d_t,
                "h2d_gbs_per_gpu_all_ranks_copying": h2d_gbs_slowest, "host": topo,
                "bound": "PCIe / host memory: the bare pinned copy of the same bytes, all ranks at once"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "n_estimated": [int(v) for v in n_estimated], "n_estimated_e2e": [int(v) for v in n_est_e2e],
        "launch_count_native": nat.launch_count - launches_before})


# ------------------------------------------------------------------------------------------------ cfg3
def bench_cfg3(ctx):
    """BASELINE configs[2]: moment covariance, Legendre R = 100, 1e9 samples in total (STRONG scaling: 1e9 / N per rank),
    per-rank partial sums combined by one all-reduce inside the timed region.  Three device-resident variants:
    DMMA contraction (means), DMMA (means + entry variances), linearised means (199 moment sums, C applied once)."""
    torch, td, nat = ctx.torch, ctx.td, ctx.nat
    from mlmc_b200.moments import Legendre
    from mlmc_b200.estimator import Estimate
    R = 100
    n_total = ctx.args.cfg3_samples
    lo, hi = (ctx.rank * n_total) // ctx.world, ((ctx.rank + 1) * n_total) // ctx.world
    n_rank = hi - lo
    fn = Legendre(R, domain())
    basis = fn.basis_struct()
    rows = synth_pairs_on_device(torch, ctx.device, n_rank, 0.05, 0.5, 77 + ctx.rank)
    x = rows.permute(2, 0, 1)
    acc = nat.LevelAccumulator(1, R * R, ctx.device)
    ext_fn, c_t = fn.product_table()
    ext_basis = ext_fn.basis_struct()
    c_dev = torch.from_numpy(c_t).to(ctx.device)
    acc_m = nat.LevelAccumulator(1, ext_fn.size, ctx.device)
    out = {}

    def reduce(a):
        if ctx.world > 1:
            td.all_reduce(a.acc)

    def run_dmma(want_var):
        acc.acc.zero_()
        nat.gram_accumulate(basis, x, acc.level(0), want_var=want_var)
        reduce(acc)
        out["dmma"] = acc.finalize()

    def run_linear():
        acc_m.acc.zero_()
        nat.moments_accumulate(ext_basis, x, acc_m.level(0), sums_only=True)
        reduce(acc_m)
        out["lin"] = nat.level_sums_transform(acc_m, 1, c_dev).finalize()

    reps = max(1, min(ctx.args.steps, 3))
    ms_mean = ctx.timed(lambda: run_dmma(False), reps)
    dmma_mean = out["dmma"]["mean"].clone()
    ms_mean_var = ctx.timed(lambda: run_dmma(True), reps)
    ms_linear = ctx.timed(run_linear, max(reps, 3))
    # the two routes to the covariance means agree (the parity tests hold both to the oracle)
    scale = float(dmma_mean.abs().max())
    lin_vs_dmma = float((out["lin"]["mean"] - dmma_mean).abs().max()) / scale
    # kernel alone (no reduce / finalize), for the roofline
    ms_kernel = ctx.timed(lambda: nat.gram_accumulate(basis, x, acc.level(0), want_var=False), 1, warmup=0)

    # end to end through the public API from pinned host rows (a bounded share of the rank's rows when N = 1)
    n_e2e = min(n_rank, 125_000_000)
    # (level 0 of a storage has no coarse part: a token level 0 keeps the cfg3 rows a fine - coarse level)
    storage, value = scalar_quantity([rows[:1000].cpu(), rows[:n_e2e].cpu()], [0.5, 0.05])
    storage.rows_are_local_shard = True
    storage.resident_fraction = 0.0
    est = Estimate(value, storage, fn)

    def e2e(variance):
        storage.drop_device_copies()
        return est.estimate_covariance(variance=variance)
    e2e_lin_ms, _ = ctx.timed_wall(lambda: e2e(False), 2)
    e2e_var_ms, _ = ctx.timed_wall(lambda: e2e(True), 1)
    del storage, est, value
    if ctx.rank != 0:
        return None
    nb = (R + 7) // 8
    dmma_peak = nat.fp64_peak(1) / 1e12
    useful = 2.0 * R * R                                   # flop per sample: R(R+1)/2 products x 2 sides x FMA
    # DMMA.8x8x4 (512 flop) per 4 samples: two per off-diagonal block pair (S^T D + D^T S), one per diagonal block
    executed = (nb * (nb - 1) / 2 * 2 + nb) * 128.0
    ach = n_rank * useful / (ms_kernel * 1e-3) / 1e12
    units = float(n_total) * R
    return {
        "workload": "cfg3: moment covariance, Legendre R=100, %.0e samples in total, %d per rank (strong scaling), "
                    "all-reduce of the partial sums inside the timed region" % (n_total, n_rank),
        "scaling": "strong", "n_gpus": ctx.world, "samples_total": n_total, "samples_per_rank": n_rank, "reps": reps,
        "cov_mean_dmma_ms": ms_mean, "cov_mean_var_dmma_ms": ms_mean_var, "cov_mean_linearised_ms": ms_linear,
        "sample_moments_per_s": {"cov_mean_dmma": units / (ms_mean * 1e-3),
                                 "cov_mean_var_dmma": units / (ms_mean_var * 1e-3),
                                 "cov_mean_linearised": units / (ms_linear * 1e-3)},
        "linearised_vs_dmma_max_rel": lin_vs_dmma,
        "reduce": "none" if ctx.world == 1 else "NCCL all-reduce of [2 + 2 R^2] doubles (160 kB) / [2 + 2 * 199]",
        "roofline": {"kernel": "gram_kernel<coarse, sums> (DMMA m8n8k4), R=100", "bound": "fp64-tensor",
                     "achieved": ach, "peak": dmma_peak, "unit": "TFLOP/s", "frac": ach / dmma_peak,
                     "useful_flop_per_sample": useful, "executed_flop_per_sample": executed,
                     "executed_tflops": n_rank * executed / (ms_kernel * 1e-3) / 1e12,
                     "executed_frac": n_rank * executed / (ms_kernel * 1e-3) / 1e12 / dmma_peak,
                     "kernel_ms": ms_kernel, "algorithmic_bytes": n_rank * 16.0,
                     "hbm_gbs": n_rank * 16.0 / (ms_kernel * 1e-3) / 1e9,
                     "peak_source": "DMMA m8n8k4 micro-benchmark measured in this run"},
        "e2e": {"api": "Estimate.estimate_covariance on a pinned-host Memory storage (H2D of the rows + D2H of the "
                       "result inside the timed region)", "rows_per_rank": n_e2e,
                "h2d_bytes_per_step": (n_e2e * 16 + 8000) * ctx.world, "d2h_bytes_per_step": (4 * R * R * 2 + 64) * 8,
                "variance_false_ms": e2e_lin_ms, "variance_true_ms": e2e_var_ms,
                "sample_moments_per_s_variance_false": float(n_e2e) * ctx.world * R / (e2e_lin_ms * 1e-3),
                "sample_moments_per_s_variance_true": float(n_e2e) * ctx.world * R / (e2e_var_ms * 1e-3)}}


def cpu_cfg3(n=20_000):
    from oracle import mlmc_oracle as orc
    rng = np.random.default_rng(77)
    rows = orc.synth_level_rows(rng.normal(size=n), 0.05, 0.5)
    t0 = time.perf_counter()
    orc.estimate_covariance([np.zeros((1, 2, 1)), rows], orc.Basis("legendre", 100, domain()), chunk_rows=2048)
    dt = time.perf_counter() - t0
    return {"kind": "port", "cores": 1, "sample": "%d samples, Legendre 100 covariance (oracle port, 2048-row chunks)" % n,
            "samples_per_s": n / dt, "sample_moments_per_s": n * 100 / dt}


# ------------------------------------------------------------------------------------------------ cfg4
def bench_cfg4(ctx, levels, steps):
    """BASELINE configs[3]: max-ent fit, 50 moments, 4762 x 21 = 100 002 Gauss nodes (constructor -> normalised
    multipliers); plus the data-driven Estimate.construct_density on cfg2's resident samples."""
    import scipy.stats as stats
    from oracle import mlmc_oracle as orc
    from mlmc_b200.moments import Legendre
    from mlmc_b200.estimator import Estimate
    from mlmc_b200.tool.simple_distribution import (SimpleDistribution, construct_ortogonal_moments,
                                                    compute_semiexact_cov, compute_semiexact_moments)
    distr = stats.norm(loc=1, scale=2)
    dom = tuple(float(v) for v in distr.ppf([0.01, 0.99]))
    n_panels = 4762
    base = Legendre(50, dom, safe_eval=False)
    cov = compute_semiexact_cov(base, distr.pdf, n_panels=n_panels)
    orth, info = construct_ortogonal_moments(base, cov, tol=1e-4)
    mu = compute_semiexact_moments(orth, distr.pdf, n_panels=n_panels)
    data = np.stack([mu, np.ones_like(mu)], axis=1)

    def fit():
        sd = SimpleDistribution(orth, data, domain=dom, quad_panels=n_panels)
        return sd, sd.estimate_density_minimize(tol=1e-8, reg_param=0.0)
    times = []
    fit()
    for _ in range(7):
        ctx.torch.cuda.synchronize()
        t0 = time.perf_counter()
        sd, res = fit()
        ctx.torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
    before = ctx.nat.launch_count
    buf = sd._buffers
    ms_eval = ctx.timed(lambda: ctx.nat.maxent_fgh(buf.phi, buf.w, buf.lam, 7, buf.out, workspace=buf.workspace), 20)
    xs = np.linspace(dom[0], dom[1], 201)
    ob = orc.Basis("legendre", 50, dom, safe_eval=False, matrix=info[2])
    t0 = time.perf_counter()
    ofit = orc.maxent_fit(ob, data, dom, tol=1e-8, n_panels=n_panels)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    r = int(orth.size)
    q = n_panels * 21
    flop = 2.0 * q * r + 2.0 * q * r + 2.0 * q * r * r       # exponent, gradient, Hessian
    out = {"workload": "cfg4: max-ent fit, %d moments (orthogonalised Legendre 50), %d Gauss nodes, trust-ncg" % (r, q),
           "fit_ms": float(np.median(times)), "fit_ms_min": float(np.min(times)), "nit": int(res.nit),
           "success": bool(res.success), "device_evals": int(sd.n_device_evals), "fgh_eval_ms": ms_eval,
           "fgh_eval_tflops": flop / (ms_eval * 1e-3) / 1e12, "launches_per_eval": (ctx.nat.launch_count - before) // 21, "cuda_graph": buf.graph is not None,
           "cpu": {"kind": "port", "cores": "numpy/BLAS default threads", "fit_ms": cpu_ms,
                   "sample": "the same fit by the oracle (NumPy + scipy trust-ncg on the same fixed rule)"},
           "max_abs_multiplier_diff_vs_oracle": float(np.max(np.abs(sd.multipliers - ofit.multipliers))),
           "max_rel_pdf_vs_oracle": max_rel(sd.density(xs), orc.maxent_density(ob, ofit.multipliers, np.ones(r), xs)),
           "target_ms": 50.0}
    # data-driven variant on the cfg2 samples (3 x samples_per_level, resident in HBM)
    storage, value = scalar_quantity([lv.cpu() for lv in levels], steps)
    dom2 = tuple(float(v) for v in stats.norm.ppf([0.001, 0.999]))
    est = Estimate(value, storage, Legendre(25, dom2))
    est.construct_density(tol=1e-8, orth_moments_tol=1e-4)
    times = []
    for _ in range(5):
        ctx.torch.cuda.synchronize()
        t0 = time.perf_counter()
        dobj, _info, res_d, mom_d = est.construct_density(tol=1e-8, orth_moments_tol=1e-4)
        ctx.torch.cuda.synchronize()
        times.append((time.perf_counter() - t0) * 1e3)
    xs2 = np.linspace(dom2[0], dom2[1], 201)[10:-10]
    out["construct_density"] = {"samples": int(sum(len(lv) for lv in levels)), "base": "Legendre 25",
                                "ms": float(np.median(times)), "orthogonal_moments": int(mom_d.size),
                                "nit": int(res_d.nit), "fun_norm": float(res_d.fun_norm),
                                "max_abs_pdf_err_vs_normal_interior": float(np.max(np.abs(dobj.density(xs2)
                                                                                         - stats.norm.pdf(xs2)))),
                                "data_passes": 1}
    out["_storage_value"] = (storage, value)
    return out


# ------------------------------------------------------------------------------------------------ bootstrap
def bench_bootstrap(ctx, storage, value, n_rep=100):
    """Estimate.est_bootstrap (mlmc/estimator.py:171-218) on cfg2's resident levels, Legendre 50,
    100 replicates -- one pass per level (multiplicities x moment differences on DMMA tiles) against the per-replicate
    gather kernel; the reference re-runs the whole estimate per replicate."""
    from mlmc_b200.moments import Legendre
    from mlmc_b200.estimator import Estimate
    n = [int(v) for v in storage.get_n_collected()]
    est = Estimate(value, storage, Legendre(N_MOMENTS, domain()))
    out = {"workload": "est_bootstrap on cfg2's levels (%s rows), Legendre R=%d, %d replicates" % (n, N_MOMENTS, n_rep)}
    for method in ("weighted", "gather"):
        os.environ["MLMCB200_BOOTSTRAP"] = method
        try:
            before = ctx.nat.launch_count
            ms, _ = ctx.timed_wall(lambda: est.est_bootstrap(n_subsamples=n_rep, seed=1), 2, warmup=1)
            out[method + "_ms"] = ms
            out[method + "_launches"] = (ctx.nat.launch_count - before) // 3
        finally:
            os.environ.pop("MLMCB200_BOOTSTRAP", None)
    out["replicate_sample_moments_per_s"] = n_rep * float(sum(n)) * N_MOMENTS / (out["weighted_ms"] * 1e-3)
    # the two kernels of the one-pass path alone, on level 1 (fine + coarse): multiplicities, then the weighted sums
    # against the FP64 tensor roofline measured in this run
    torch, nat, dev = ctx.torch, ctx.nat, ctx.device
    rows = next(r for l, r in storage.device_chunks([1], dev, keep_resident=True))
    x = rows.permute(2, 0, 1)
    n1 = int(rows.shape[0])
    P = max(1, -(-n1 // 131072))
    edges = (np.arange(P + 1, dtype=np.int64) * n1) // P
    rng = np.random.default_rng(0)
    cum_h = np.zeros((n_rep, P + 1), dtype=np.int64)
    for b in range(n_rep):
        np.cumsum(rng.multinomial(n1, np.diff(edges) / n1), out=cum_h[b, 1:])
    cum = torch.from_numpy(cum_h).to(dev)
    basis = est._moments_fn.basis_struct()
    counts = nat.resample_counts(1, 1, n1, cum, n1, dev)
    acc = torch.zeros((n_rep, 2 + 2 * N_MOMENTS), dtype=torch.float64, device=dev)
    ms_counts = ctx.timed(lambda: nat.resample_counts(1, 1, n1, cum, n1, dev), 3, warmup=1)
    ms_w = ctx.timed(lambda: nat.moments_accumulate_weighted(basis, x, counts, acc), 3, warmup=1)
    peak = nat.fp64_peak(1)
    cols, reps = 8 * (-(-(2 + 2 * N_MOMENTS) // 8)), 8 * (-(-n_rep // 8))
    out["kernels_level1"] = {
        "resample_counts_ms": ms_counts, "weighted_sums_ms": ms_w,
        "roofline": {"bound": "fp64 tensor (DMMA m8n8k4)", "peak": peak / 1e12, "unit": "TFLOP/s",
                     "executed": 2.0 * n1 * cols * reps / (ms_w * 1e-3) / 1e12,
                     "useful": 2.0 * n1 * (2 * N_MOMENTS) * n_rep / (ms_w * 1e-3) / 1e12,
                     "frac": 2.0 * n1 * cols * reps / (ms_w * 1e-3) / peak,
                     "frac_useful": 2.0 * n1 * (2 * N_MOMENTS) * n_rep / (ms_w * 1e-3) / peak,
                     "algorithmic_bytes": n1 * (16 + n_rep),
                     "hbm_gbs": n1 * (16 + n_rep) / (ms_w * 1e-3) / 1e9,
                     "note": "useful = 2 flop x rows x 2R columns x replicates; executed adds the padding to 104 x 104"}}
    return out


# ------------------------------------------------------------------------------------------------ cfg1
def bench_cfg1(ctx):
    """BASELINE configs[0]: 1 level, 1e5 lognormal samples, Legendre 25 on the estimated log domain, estimate_moments."""
    from oracle import mlmc_oracle as orc
    from mlmc_b200.moments import Legendre
    from mlmc_b200.estimator import Estimate
    n = 100_000
    rows = synth_pairs_on_device(ctx.torch, ctx.device, n, 0.1, None, 1234, lognormal=True).cpu().numpy()
    storage, value = scalar_quantity([rows], [0.1])
    dom = Estimate.estimate_domain(value, storage, quantile=0.001)
    fn = Legendre(25, dom, log=True, safe_eval=True)
    est = Estimate(value, storage, fn)
    ms_res, (means, variances) = ctx.timed_wall(est.estimate_moments, 50, warmup=3)
    storage.resident_fraction = 0.0

    def staged():
        storage.drop_device_copies()
        return est.estimate_moments()
    ms_e2e, _ = ctx.timed_wall(staged, 50, warmup=3)
    t0 = time.perf_counter()
    o = orc.estimate_moments([rows], orc.Basis("legendre", 25, tuple(dom), log=True))
    cpu_ms = (time.perf_counter() - t0) * 1e3
    return {"workload": "cfg1: 1 level, 1e5 lognormal samples, Legendre R=25 (log domain), estimate_moments",
            "resident_ms": ms_res, "e2e_ms": ms_e2e, "h2d_bytes_per_step": n * 8,
            "sample_moments_per_s_resident": n * 25 / (ms_res * 1e-3), "sample_moments_per_s_e2e": n * 25 / (ms_e2e * 1e-3),
            "cpu": {"kind": "port", "cores": 1, "ms": cpu_ms, "sample_moments_per_s": n * 25 / (cpu_ms * 1e-3),
                    "sample": "the full config by the oracle port"},
            "max_rel_mean_vs_oracle": max_rel(means, o.mean), "max_rel_var_vs_oracle": max_rel(variances, o.var),
            "n_rm_samples": int(o.n_rm_samples[0])}


# ------------------------------------------------------------------------------------------------ cfg5
def bench_cfg5(ctx):
    """BASELINE configs[4]: vector quantity, 1e4 locations x 5 levels, Fourier R = 32, per-location mean / variance."""
    torch = ctx.torch
    from oracle import mlmc_oracle as orc
    from mlmc_b200.moments import Fourier
    from mlmc_b200.sample_storage import Memory
    from mlmc_b200.quantity.quantity import make_root_quantity
    from mlmc_b200.quantity.quantity_spec import QuantitySpec
    from mlmc_b200.quantity import quantity_estimate as qe
    M = 10_000
    n_levels = [4096, 2048, 1024, 512, 256]
    steps = orc.level_steps(5, (0.5, 0.005))
    offs = torch.arange(M, dtype=torch.float64, device=ctx.device) * 1e-4
    levels = []
    for l, n in enumerate(n_levels):
        base = synth_pairs_on_device(torch, ctx.device, n, steps[l], steps[l - 1] if l else None, 500 + l)   # [n, 2, 1]
        rows = base + offs[None, None, :]
        if l == 0:
            rows[:, 1, :] = 0
        levels.append(rows.cpu())
        del rows
    spec = [QuantitySpec(name="field", unit="", shape=(1, 1), times=[0.0], locations=[str(i) for i in range(M)])]
    storage = Memory.from_arrays(levels, level_parameters=[[h] for h in steps], result_format=spec)
    field = make_root_quantity(storage, spec)["field"][0.0]
    dom = (-4.2, 5.4)
    fn = Fourier(32, dom)
    ms_res, qm = ctx.timed_wall(lambda: qe.estimate_mean(qe.moments(field, fn)), 5, warmup=2)
    storage.resident_fraction = 0.0

    def staged():
        storage.drop_device_copies()
        return qe.estimate_mean(qe.moments(field, fn))
    # two warm-up calls: the 30 MB result goes to pooled pinned buffers and two of them alternate (the previous result
    # is still referenced while the next call runs); pinning the second one inside the timed calls cost 10-20 ms
    ms_e2e, _ = ctx.timed_wall(staged, 3, warmup=2)
    # the same estimate with the samples staged from an HDF5 file in the reference's layout (BASELINE configs[4]:
    # "per-location mean/var from HDF-staged samples"): Levels/<l>/collected_values chunks -> two pinned staging buffers
    # -> device -> kernels, nothing resident
    hdf = hdf_staged_cfg5(ctx, levels, steps, spec, fn, qm)
    # oracle on 50 locations spread over the field (sample mask of all locations applied first)
    ob = orc.Basis("fourier", 32, dom)
    pick = np.arange(0, M, 200)
    sl = []
    for l, lv in enumerate(levels):
        lv = lv.numpy()
        t = orc.to_ref_domain(ob, lv[:, :1, :] if l == 0 else lv)
        keep = ~np.isnan(t).any(axis=(1, 2))
        sl.append(np.ascontiguousarray(lv[keep][:, :, pick]))
    t0 = time.perf_counter()
    o = orc.estimate_moments(sl, ob, chunk_rows=512)
    cpu_s = time.perf_counter() - t0
    units = float(sum(n_levels)) * M * 32
    n_bytes = sum(n_levels[1:]) * M * 16 + n_levels[0] * M * 8
    got_means = qm.l_means.reshape(5, M, 32)[:, pick].reshape(5, -1)
    got_vars = qm.l_vars.reshape(5, M, 32)[:, pick].reshape(5, -1)
    return {"workload": "cfg5: field of 1e4 locations x 5 levels (4096..256 samples), Fourier R=32, per-location mean/var",
            "resident_ms": ms_res, "e2e_ms": ms_e2e, "h2d_bytes_per_step": n_bytes, "d2h_bytes_per_step": (2 * 5 + 2) * M * 32 * 8,
            "sample_moments_per_s_resident": units / (ms_res * 1e-3), "sample_moments_per_s_e2e": units / (ms_e2e * 1e-3),
            "hbm_gbs_resident": n_bytes / (ms_res * 1e-3) / 1e9, "hdf_staged": hdf,
            "cpu": {"kind": "port", "cores": 1, "sample_moments_per_s": sum(len(s) for s in sl) * len(pick) * 32 / cpu_s,
                    "sample": "50 of the 1e4 locations, all samples (oracle port, 512-row chunks)"},
            "n_samples": [int(v) for v in qm.n_samples], "n_rm_samples": [int(v) for v in qm.n_rm_samples],
            "max_rel_l_means_vs_oracle": max_rel(got_means, o.l_means), "max_rel_l_vars_vs_oracle": max_rel(got_vars, o.l_vars)}


def hdf_staged_cfg5(ctx, levels, steps, spec, fn, qm_resident):
    """cfg5 read from an HDF5 sample file (written here in the reference's structure by the fixture writer, read by
    ``SampleStorageHDF``: h5py when importable, else the built-in reader).  The file sits in the page cache: the number is
    the read side's own cost (HDF5 chunk lookup + copy into pinned staging) plus the PCIe copy, not the disk's."""
    import shutil
    import tempfile
    from mlmc_b200.sample_storage import SampleStorageHDF
    from mlmc_b200.tool.hdf5_min import write_mlmc_file
    from mlmc_b200.quantity.quantity import make_root_quantity
    from mlmc_b200.quantity import quantity_estimate as qe
    tmp = tempfile.mkdtemp(prefix="mlmcb200_bench_")
    try:
        t0 = time.perf_counter()
        path = write_mlmc_file(os.path.join(tmp, "cfg5.hdf5"), [lv.numpy() for lv in levels], [[h] for h in steps])
        write_s = time.perf_counter() - t0
        storage = SampleStorageHDF(path)
        storage.resident_fraction = 0.0
        field = make_root_quantity(storage, spec)["field"][0.0]
        times = []
        for rep in range(7):                                  # the first two calls map the file's pages: not timed
            ctx.torch.cuda.synchronize()
            t0 = time.perf_counter()
            qm = qe.estimate_mean(qe.moments(field, fn))
            ctx.torch.cuda.synchronize()
            if rep >= 2:
                times.append((time.perf_counter() - t0) * 1e3)
        ms = float(np.median(times))
        size = os.path.getsize(path)
        return {"ms": ms, "ms_min": float(np.min(times)), "ms_max": float(np.max(times)), "file_bytes": size,
                "file_gbs": size / (ms * 1e-3) / 1e9, "backend": storage.backend,
                "fixture_write_s": write_s, "page_cache": True,
                "max_rel_l_means_vs_resident": max_rel(qm.l_means, qm_resident.l_means),
                "max_rel_l_vars_vs_resident": max_rel(qm.l_vars, qm_resident.l_vars),
                "n_samples": [int(v) for v in qm.n_samples]}
    except Exception as exc:                                  # an optional leg must not cost the bench line
        return {"error": "%s: %s" % (type(exc).__name__, exc)}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_gpu_arm(args):
    ctx = Ctx(args)
    parity = startup_parity_check(ctx)
    line = {}
    bench_cfg2(ctx, line)
    levels, steps = line.pop("_levels"), line.pop("_steps")
    wanted = [c for c in args.configs.split(",") if c]
    configs = {}
    if "cfg3" in wanted:
        res = bench_cfg3(ctx)
        if ctx.rank == 0:
            configs["cfg3"] = res
            if ctx.world == 1 and not args.no_cpu_baseline:
                configs["cfg3"]["cpu"] = cpu_cfg3()
    ctx.barrier()
    # cfg1 / cfg4 / cfg5 are single-GPU configurations (max-ent: replicas only): timed at N = 1
    if ctx.world == 1:
        if "cfg4" in wanted:
            configs["cfg4"] = bench_cfg4(ctx, levels, steps)
            storage_value = configs["cfg4"].pop("_storage_value")
            if "bootstrap" in wanted:
                configs["bootstrap"] = bench_bootstrap(ctx, *storage_value)
            del storage_value
        if "cfg1" in wanted:
            configs["cfg1"] = bench_cfg1(ctx)
        if "cfg5" in wanted:
            del levels
            ctx.torch.cuda.empty_cache()
            configs["cfg5"] = bench_cfg5(ctx)
    elif ctx.rank == 0:
        configs["note"] = "cfg1 / cfg4 / cfg5 are single-GPU configurations: reported by the N = 1 run"
    if ctx.rank == 0:
        line["parity_check"] = parity
        line["configs"] = configs
        emit(line)
    ctx.barrier()


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
    ap.add_argument("--configs", default="cfg1,cfg3,cfg4,cfg5,bootstrap",
                    help="further BASELINE configs reported under `configs` in the same JSON line ('' = none)")
    ap.add_argument("--cfg3-samples", type=int, default=1_000_000_000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)
        try:                                        # leave the process group cleanly (NCCL warns otherwise)
            import torch.distributed as td
            if td.is_available() and td.is_initialized():
                td.barrier()
                td.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
