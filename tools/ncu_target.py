"""Small fixed workload for ncu: the cfg2 device step (3 levels, Legendre R=50, 1e7 samples/level) plus one
covariance launch (R=100, 1e6 samples) and one max-ent F/g/H evaluation (Q=100002, R=50), each run 3 times."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mlmc_b200 import _native as nat  # noqa: E402
from mlmc_b200.moments import Legendre  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0")
n = int(os.environ.get("NCU_ROWS", "10000000"))
levels = bench.make_levels_on_device(torch, dev, n, 1234)
fn = Legendre(bench.N_MOMENTS, bench.domain())
basis = fn.basis_struct()
acc = nat.LevelAccumulator(3, bench.N_MOMENTS, dev)
views = [lv.permute(2, 0, 1)[:, :, :1] if l == 0 else lv.permute(2, 0, 1) for l, lv in enumerate(levels)]
for rep in range(3):
    if which in ("all", "moments"):
        acc.acc.zero_()
        for l in range(3):
            nat.moments_accumulate(basis, views[l], acc.level(l))
        out = acc.finalize()
    if which in ("all", "gram"):
        b100 = Legendre(100, bench.domain()).basis_struct()
        acc_c = nat.LevelAccumulator(1, 100 * 100, dev)
        nat.gram_accumulate(b100, views[1][:, :1_000_000], acc_c.level(0), want_var=True)
        nat.gram_accumulate(b100, views[1][:, :1_000_000], acc_c.level(0), want_var=False)
    if which in ("all", "maxent"):
        Q, R = 100_002, 50
        phi = torch.randn(Q, R, device=dev, dtype=torch.float64) * 0.1
        w = torch.full((Q,), 1.0 / Q, device=dev, dtype=torch.float64)
        lam = torch.zeros(R, device=dev, dtype=torch.float64)
        nat.maxent_fgh(phi, w, lam, 7)
if which == "resampled":
    nb = 4
    for layout in ("random", "blocked"):
        if layout == "random":
            idx = torch.randint(0, n, (nb, n), dtype=torch.int32, device=dev)
        else:
            bs = n // 16
            seg = (torch.arange(n, device=dev) // bs).clamp(max=15).to(torch.int32)
            idx = (torch.randint(0, bs, (nb, n), dtype=torch.int32, device=dev) + seg[None, :] * bs).contiguous()
        acc_r = torch.zeros((nb, 2 + 2 * bench.N_MOMENTS), dtype=torch.float64, device=dev)
        nat.moments_accumulate_resampled(basis, views[1], idx, acc_r)
if which == "weighted":
    # all bootstrap replicates of the middle level in one pass: multiplicities (shared-memory histogram of the Philox
    # draws) + weighted sums on DMMA tiles (csrc/bootstrap.cu)
    B, P = 100, max(1, -(-n // 131072))
    edges = (np.arange(P + 1, dtype=np.int64) * n) // P
    rng = np.random.default_rng(0)
    cum_h = np.zeros((B, P + 1), dtype=np.int64)
    for b in range(B):
        np.cumsum(rng.multinomial(n, np.diff(edges) / n), out=cum_h[b, 1:])
    cum = torch.from_numpy(cum_h).to(dev)
    acc_w = torch.zeros((B, 2 + 2 * bench.N_MOMENTS), dtype=torch.float64, device=dev)
    for rep in range(2):
        counts = nat.resample_counts(1, 1, n, cum, n, dev)
        nat.moments_accumulate_weighted(basis, views[1], counts, acc_w)
torch.cuda.synchronize()
print("ok", float(out["mean"][1]) if which in ("all", "moments") else "")
