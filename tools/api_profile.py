"""Host-side profile of the public API on HBM-resident levels (cfg2 shape): where do the microseconds between the
kernels go?  Wall time per estimate_moments call vs the GPU time of its kernels, and a cProfile of 300 calls."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mlmc_b200.estimator import Estimate  # noqa: E402
from mlmc_b200.moments import Legendre  # noqa: E402
from mlmc_b200.quantity.quantity import make_root_quantity  # noqa: E402
from mlmc_b200.quantity.quantity_spec import QuantitySpec  # noqa: E402
from mlmc_b200.sample_storage import Memory  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda:0")
steps = bench.level_steps()
levels = [lv.cpu() for lv in bench.make_levels_on_device(torch, dev, n, 1234)]
spec = [QuantitySpec(name="v", unit="", shape=(1, 1), times=[0.0], locations=["0"])]
storage = Memory.from_arrays(levels, level_parameters=[[h] for h in steps], result_format=spec)
value = make_root_quantity(storage, spec)["v"][0.0]["0"][0, 0]
est = Estimate(value, storage, Legendre(bench.N_MOMENTS, bench.domain()))
for _ in range(5):
    est.estimate_moments()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    est.estimate_moments()
dt = (time.perf_counter() - t0) / 200
print("estimate_moments, resident levels: %.3f ms per call" % (dt * 1e3))
prof = cProfile.Profile()
prof.enable()
for _ in range(300):
    est.estimate_moments()
prof.disable()
pstats.Stats(prof).sort_stats("cumulative").print_stats(35)
