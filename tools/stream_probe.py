"""HBM-bound regime: few moments, many samples.  10 launches between CUDA events."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlmc_b200 import _native as nat

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(n, generator=g, device=dev, dtype=torch.float64)
rows = torch.stack([x + 0.05, x + 0.5], dim=1).unsqueeze(2).contiguous()
del x
xv = rows.permute(2, 0, 1)
x0 = rows[:, 0, :].reshape(n, 1, 1).contiguous().permute(2, 0, 1)
for name, R, kind in (("raw", 1, nat.RAW), ("leg", 2, nat.LEGENDRE), ("leg", 5, nat.LEGENDRE), ("leg", 8, nat.LEGENDRE),
                      ("leg", 12, nat.LEGENDRE), ("leg", 13, nat.LEGENDRE), ("leg", 25, nat.LEGENDRE)):
    basis = nat.RAW_BASIS if kind == nat.RAW else nat.make_basis(kind, R, (-3.72, 3.72), (-1.0, 1.0))
    acc = nat.LevelAccumulator(1, R, dev)
    for label, view, nbytes in (("pairs", xv, 16), ("level0", x0, 8)):
        for _ in range(3):
            nat.moments_accumulate(basis, view, acc.level(0))
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                nat.moments_accumulate(basis, view, acc.level(0))
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 10)
        print("%s R=%2d %-6s %.4f ms  %7.1f GB/s  %.3e sample-moments/s" % (name, R, label, best, n * nbytes / best / 1e6, n * R / best * 1e3), flush=True)
