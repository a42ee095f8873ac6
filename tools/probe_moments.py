"""Moments-kernel timing sweep (CUDA events, best of 5 regions of PROBE_B2B back-to-back calls): R in 8..199, fine+coarse / level 0, full and sums-only."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlmc_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
n = int(os.environ.get("PROBE_ROWS", "10000000"))
g = torch.Generator(device=dev).manual_seed(1)
x0 = torch.randn(n, generator=g, device=dev, dtype=torch.float64)
root = torch.sqrt(1e-4 + x0.abs())
rows = torch.stack([x0 + 0.05 * root, x0 + 0.5 * root], dim=1).unsqueeze(2).contiguous()
x = rows.permute(2, 0, 1)
peak = nat.fp64_peak(0) / 2          # FP64 lane-instructions / s


BACK_TO_BACK = int(os.environ.get("PROBE_B2B", "10"))      # launches per timed region (1: a lone call, launch latency exposed)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(BACK_TO_BACK):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / BACK_TO_BACK)
    return best


print("n = %d, DFMA peak %.2f TFLOP/s" % (n, peak * 2 / 1e12))
print("%5s %10s %10s %10s %10s | %s" % ("R", "pair ms", "lvl0 ms", "sums ms", "sums0 ms", "frac of pipe (7.25 / 4.25 / 6 / 3 instr)"))
for R in [int(v) for v in os.environ.get("PROBE_R", "8,12,16,25,32,50,64,100,150,199").split(",")]:
    basis = nat.make_basis(nat.LEGENDRE, R, (-3.72, 3.72), (-1.0, 1.0))
    acc = nat.LevelAccumulator(1, R, dev)
    t = [timed(lambda: nat.moments_accumulate(basis, x, acc.level(0))),
         timed(lambda: nat.moments_accumulate(basis, x[:, :, :1], acc.level(0))),
         timed(lambda: nat.moments_accumulate(basis, x, acc.level(0), sums_only=True)),
         timed(lambda: nat.moments_accumulate(basis, x[:, :, :1], acc.level(0), sums_only=True))]
    fr = [n * R * c / (ti * 1e-3) / peak for ti, c in zip(t, (7.25, 4.25, 6.0, 3.0))]
    print("%5d %10.4f %10.4f %10.4f %10.4f | %.2f %.2f %.2f %.2f   hbm %.0f GB/s" % (
        R, t[0], t[1], t[2], t[3], fr[0], fr[1], fr[2], fr[3], n * 16 / (t[0] * 1e-3) / 1e9), flush=True)
