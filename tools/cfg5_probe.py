"""cfg5 (1e4 locations x 5 levels, Fourier 32) through the public API on resident levels: wall time per call, for an ncu
launch list (`ncu --metrics gpu__time_duration.sum --clock-control none python tools/cfg5_probe.py`)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle import mlmc_oracle as orc  # noqa: E402  (level steps only)
from mlmc_b200.moments import Fourier, Legendre  # noqa: E402
from mlmc_b200.sample_storage import Memory  # noqa: E402
from mlmc_b200.quantity.quantity import make_root_quantity  # noqa: E402
from mlmc_b200.quantity.quantity_spec import QuantitySpec  # noqa: E402
from mlmc_b200.quantity import quantity_estimate as qe  # noqa: E402

dev = torch.device("cuda:0")
M = int(os.environ.get("PROBE_M", "10000"))
n_levels = [4096, 2048, 1024, 512, 256]
steps = orc.level_steps(5, (0.5, 0.005))
offs = torch.arange(M, dtype=torch.float64, device=dev) * 1e-4
levels = []
for l, n in enumerate(n_levels):
    base = bench.synth_pairs_on_device(torch, dev, n, steps[l], steps[l - 1] if l else None, 500 + l)
    rows = base + offs[None, None, :]
    if l == 0:
        rows[:, 1, :] = 0
    levels.append(rows.cpu())
spec = [QuantitySpec(name="field", unit="", shape=(1, 1), times=[0.0], locations=[str(i) for i in range(M)])]
storage = Memory.from_arrays(levels, level_parameters=[[h] for h in steps], result_format=spec)
field = make_root_quantity(storage, spec)["field"][0.0]
for name, fn in (("fourier32", Fourier(32, (-4.2, 5.4))), ("legendre32", Legendre(32, (-4.2, 5.4)))):
    q = qe.moments(field, fn)
    for rep in range(int(os.environ.get("PROBE_REPS", "4"))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        qm = qe.estimate_mean(q)
        torch.cuda.synchronize()
        print("%s call %d: %.3f ms" % (name, rep, 1e3 * (time.perf_counter() - t0)), flush=True)
print(np.asarray(qm.n_samples))
