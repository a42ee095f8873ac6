"""Bootstrap on the cfg2 shape (3 levels x n rows, Legendre R, B replicates): the one-pass weighted sums on DMMA tiles
(csrc/bootstrap.cu) against the per-replicate gather kernel; CUDA events, best of 3, levels resident.
usage: probe_bootstrap.py [n] [B] [R]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlmc_b200 import _native as nat  # noqa: E402
from mlmc_b200.estimator import Estimate  # noqa: E402
from mlmc_b200.moments import Legendre  # noqa: E402
from mlmc_b200.quantity.quantity import make_root_quantity  # noqa: E402
from mlmc_b200.quantity.quantity_spec import QuantitySpec  # noqa: E402
from mlmc_b200.sample_storage import Memory  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100
R = int(sys.argv[3]) if len(sys.argv) > 3 else 50
dev = torch.device("cuda:0")
levels = []
for l in range(3):
    rng = np.random.default_rng(1234 + 1000 * l)
    x = rng.normal(size=n)
    rows = np.stack([x, (x + (0.5 ** (l + 1)) * rng.normal(size=n)) if l else np.zeros(n)], axis=1)
    levels.append(np.ascontiguousarray(rows).reshape(n, 2, 1))
spec = [QuantitySpec(name="v", unit="", shape=(1, 1), times=[0.0], locations=["0"])]
storage = Memory.from_arrays(levels, level_parameters=[[0.5 ** l] for l in range(3)], result_format=spec)
value = make_root_quantity(storage, spec)["v"][0.0]["0"][0, 0]
fn = Legendre(R, (-3.719016485455709, 3.719016485455709))
est = Estimate(value, storage, fn)
est.estimate_moments()


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
res = {}
for method in ("weighted", "gather"):
    os.environ["MLMCB200_BOOTSTRAP"] = method
    est.est_bootstrap(n_subsamples=8, seed=1)
    out["est_bootstrap_ms_" + method] = timed(lambda: est.est_bootstrap(n_subsamples=B, seed=1))
    res[method] = (np.array(est.mean_bs_mean), np.array(est.mean_bs_var), np.array(est.mean_bs_l_vars))
os.environ.pop("MLMCB200_BOOTSTRAP")
out["max_rel_diff_between_methods"] = [float(np.max(np.abs(a - b) / (np.abs(b) + 1e-300))) for a, b in
                                       zip(res["weighted"], res["gather"])]
out["replicate_sample_moments_per_s_weighted"] = B * 3 * n * R / out["est_bootstrap_ms_weighted"] * 1e3

# the kernels alone on the middle level
basis = fn.basis_struct()
x = torch.from_numpy(levels[1]).to(dev).permute(2, 0, 1)
P = max(1, -(-n // 131072))
edges = (np.arange(P + 1, dtype=np.int64) * n) // P
rng = np.random.default_rng(0)
cum_h = np.zeros((B, P + 1), dtype=np.int64)
for b in range(B):
    np.cumsum(rng.multinomial(n, np.diff(edges) / n), out=cum_h[b, 1:])
cum = torch.from_numpy(cum_h).to(dev)
counts = nat.resample_counts(1, 1, n, cum, n, dev)
out["counts_kernel_ms"] = timed(lambda: nat.resample_counts(1, 1, n, cum, n, dev))
acc = torch.zeros((B, 2 + 2 * R), dtype=torch.float64, device=dev)
nat.moments_accumulate_weighted(basis, x, counts, acc)
t = timed(lambda: nat.moments_accumulate_weighted(basis, x, counts, acc))
out["weighted_kernel_ms"] = t
cols, reps = 8 * (-(-(2 + 2 * R) // 8)), 8 * (-(-B // 8))
peak = nat.fp64_peak(1)
out["dmma_peak_tflops"] = peak / 1e12
out["weighted_executed_tflops"] = 2.0 * n * cols * reps / t / 1e9
out["weighted_useful_tflops"] = 2.0 * n * (2 * R) * B / t / 1e9
out["weighted_executed_frac"] = out["weighted_executed_tflops"] * 1e12 / peak
print(json.dumps(out, indent=1))
