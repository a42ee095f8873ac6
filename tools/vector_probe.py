"""Throughput of the moments path for vector quantities of M components (n x M = 1e7 values per side, Legendre R = 25):
the public call (mask inside the kernel for M <= 128, separate mask pass above), and the two passes timed one by one
(each of those includes ~0.05-0.1 ms of host launch latency).  CUDA events, best of 5."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlmc_b200 import _native as nat
dev = torch.device("cuda:0")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 25
basis = nat.make_basis(nat.LEGENDRE, R, (-4.5, 4.5), (-1.0, 1.0))
for M in (1, 2, 8, 24, 64, 127, 128, 256, 1000, 10000):
    n = 10_000_000 // M
    g = torch.Generator(device=dev).manual_seed(M)
    rows = torch.randn((n, 2, M), generator=g, device=dev, dtype=torch.float64)
    rows[:, 1, :] = rows[:, 0, :] + 0.01
    x = rows.permute(2, 0, 1)
    acc = torch.zeros(2 + 2 * M * R, dtype=torch.float64, device=dev)
    def run():                                     # what estimate_mean does: M <= 128 masks inside the kernel
        nat.moments_accumulate(basis, x, acc)
    for _ in range(2):
        run()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    parts = []
    valid = nat.sample_mask(basis, x) if M > 1 else None
    for fn in ((lambda: nat.sample_mask(basis, x)) if M > 1 else None, lambda: nat.moments_accumulate(basis, x, acc, valid=valid)):
        if fn is None:
            parts.append(0.0)
            continue
        b2 = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            b2 = min(b2, e0.elapsed_time(e1))
        parts.append(b2)
    print("M=%5d n=%8d  %.3f ms (mask %.3f, accumulate %.3f)  %.3e sample-moments/s"
          % (M, n, best, parts[0], parts[1], n * M * R / best * 1e3), flush=True)
