"""Where does the end-to-end time go?  Pinned H2D bandwidth + a timed breakdown of one public-API estimate."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mlmc_b200 import _native as nat  # noqa: E402
from mlmc_b200.moments import Legendre  # noqa: E402
from mlmc_b200.sample_storage import Memory, stream_rows  # noqa: E402
from mlmc_b200.quantity.quantity import make_root_quantity  # noqa: E402
from mlmc_b200.quantity.quantity_spec import QuantitySpec  # noqa: E402
from mlmc_b200.estimator import Estimate  # noqa: E402
from mlmc_b200.quantity import quantity_estimate as qe  # noqa: E402

dev = torch.device("cuda:0")
for mb in (32, 160, 480):
    h = torch.empty(mb << 17, dtype=torch.float64, pin_memory=True)
    d = torch.empty_like(h, device=dev)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("H2D pinned %4d MB: %.2f ms  %.1f GB/s" % (mb, dt * 1e3, (mb << 20) / dt / 1e9), flush=True)
    del h, d

n = 10_000_000
levels = [lv.cpu() for lv in bench.make_levels_on_device(torch, dev, n, 1234)]
steps = bench.level_steps()
spec = [QuantitySpec(name="v", unit="", shape=(1, 1), times=[0.0], locations=["0"])]
storage = Memory.from_arrays(levels, level_parameters=[[h] for h in steps], n_ops=[1, 2, 3], result_format=spec)
print("pinned:", storage._host_tensor(1).is_pinned())
value_q = make_root_quantity(storage, spec)["v"][0.0]["0"][0, 0]
fn = Legendre(50, bench.domain())
est = Estimate(value_q, storage, fn)
for chunk_mb in (8, 32, 160):
    storage.resident_fraction = 0.0
    storage.device_chunk_bytes = chunk_mb << 20
    for _ in range(2):
        est.estimate_diff_vars_regression(None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        est.estimate_diff_vars_regression(None)
    torch.cuda.synchronize()
    print("chunk %3d MB: e2e %.2f ms/step" % (chunk_mb, (time.perf_counter() - t0) / 5 * 1e3), flush=True)

# breakdown: stream only
for _ in range(2):
    for l in range(3):
        for rows in stream_rows(storage._host_tensor(l), dev, 32 << 20):
            pass
torch.cuda.synchronize()
t0 = time.perf_counter()
for l in range(3):
    for rows in stream_rows(storage._host_tensor(l), dev, 32 << 20):
        pass
torch.cuda.synchronize()
print("stream_rows only (3 levels): %.2f ms" % ((time.perf_counter() - t0) * 1e3))
t0 = time.perf_counter()
q = qe.moments(value_q, fn)
torch.cuda.synchronize()
print("build lazy quantity: %.3f ms" % ((time.perf_counter() - t0) * 1e3))
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
est.estimate_diff_vars_regression(None)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
