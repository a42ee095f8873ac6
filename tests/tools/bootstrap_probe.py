"""Fused bootstrap: all replicates of est_bootstrap on the cfg2 shape.

Times (CUDA events around the public call, device-resident storage):
  * Estimate.est_bootstrap(n_subsamples=B) on the fused path,
  * the kernel alone (row numbers already drawn),
  * B plain estimates of the full levels (the cost of re-running the hot path B times without re-sampling),
and one replicate of the CPU oracle on the same shape (gather + estimate), scaled to B.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mlmc_b200 import _native as nat
from mlmc_b200.estimator import Estimate
from mlmc_b200.moments import Legendre
from mlmc_b200.quantity.quantity import make_root_quantity
from mlmc_b200.quantity.quantity_spec import QuantitySpec
from mlmc_b200.sample_storage import Memory
from oracle import mlmc_oracle as orc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100
R = int(sys.argv[3]) if len(sys.argv) > 3 else 50
dev = torch.device("cuda:0")
steps = orc.level_steps(3, (0.5, 0.005))
levels = []
for l in range(3):
    rng = np.random.default_rng(1234 + 1000 * l)
    x = rng.normal(size=n)
    levels.append(orc.synth_level_rows(x, steps[l], steps[l - 1] if l else None))
spec = [QuantitySpec(name="v", unit="", shape=(1, 1), times=[0.0], locations=["0"])]
storage = Memory.from_arrays(levels, level_parameters=[[h] for h in steps], result_format=spec)
value = make_root_quantity(storage, spec)["v"][0.0]["0"][0, 0]
domain = (-3.719016485455709, 3.719016485455709)
fn = Legendre(R, domain)
est = Estimate(value, storage, fn)
est.estimate_moments()                                     # levels become resident


def timed(fn_, reps=3):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn_()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


out = {"n_per_level": n, "replicates": B, "moments": R}
est.est_bootstrap(n_subsamples=2, seed=1)                  # warm-up
out["est_bootstrap_ms"] = timed(lambda: est.est_bootstrap(n_subsamples=B, seed=1))
out["replicate_sample_moments_per_s"] = B * 3 * n * R / out["est_bootstrap_ms"] * 1e3

# kernel alone on the middle level
basis = fn.basis_struct()
rows = torch.from_numpy(levels[1]).to(dev)
x = rows.permute(2, 0, 1)
nb = min(B, 20)
idx = torch.randint(0, n, (nb, n), dtype=torch.int32, device=dev)
acc = torch.zeros((nb, 2 + 2 * R), dtype=torch.float64, device=dev)
nat.moments_accumulate_resampled(basis, x, idx, acc)
out["kernel_ms_per_replicate_level"] = timed(lambda: nat.moments_accumulate_resampled(basis, x, idx, acc)) / nb
plain = torch.zeros(2 + 2 * R, dtype=torch.float64, device=dev)
nat.moments_accumulate(basis, x, plain)
out["plain_kernel_ms_per_level"] = timed(lambda: nat.moments_accumulate(basis, x, plain))
out["randint_ms_per_replicate_level"] = timed(
    lambda: torch.randint(0, n, (nb, n), dtype=torch.int32, device=dev)) / nb
out["plain_estimates_ms"] = timed(lambda: [est.estimate_moments() for _ in range(min(B, 10))]) * B / min(B, 10)

# CPU oracle: one replicate on a bounded sample of the shape, scaled
m = min(n, 1_000_000)
rng = np.random.default_rng(0)
t0 = time.perf_counter()
picked = [lv[rng.integers(0, m, m)] for lv in (levels[0][:m], levels[1][:m], levels[2][:m])]
orc.estimate_moments(picked, orc.Basis("legendre", R, domain))
t_cpu = time.perf_counter() - t0
out["cpu_oracle_s_per_replicate_scaled"] = t_cpu * n / m
out["cpu_oracle_s_all_replicates_scaled"] = t_cpu * n / m * B
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/bootstrap_probe.json", "w") as f:
    json.dump(out, f, indent=1)
