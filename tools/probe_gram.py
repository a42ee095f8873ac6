"""Covariance (DMMA) kernel timing: R in {25, 50, 100}, sums / sums + squares / difference Gram; executed and useful
TFLOP/s against the DMMA peak measured in the same run."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlmc_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
n = int(os.environ.get("PROBE_ROWS", "8000000"))
g = torch.Generator(device=dev).manual_seed(2)
x0 = torch.randn(n, generator=g, device=dev, dtype=torch.float64)
root = torch.sqrt(1e-4 + x0.abs())
rows = torch.stack([x0 + 0.05 * root, x0 + 0.5 * root], dim=1).unsqueeze(2).contiguous()
x = rows.permute(2, 0, 1)
peak = nat.fp64_peak(1) / 1e12


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


print("n = %d, DMMA peak %.2f TFLOP/s" % (n, peak))
for R in (25, 50, 100):
    basis = nat.make_basis(nat.LEGENDRE, R, (-3.72, 3.72), (-1.0, 1.0))
    acc = nat.LevelAccumulator(1, R * R, dev)
    nb = (R + 7) // 8
    off, dia = nb * (nb - 1) // 2, nb
    # DMMAs per 4 samples (fine + coarse tiles hold S, D: two per off-diagonal block and one per diagonal block for the
    # sums, three / two for the sums of squares)
    for name, kw, dmmas, useful in (("sums", dict(mode=0, want_var=False), 2 * off + dia, 2.0 * R * R),
                                    ("sums+squares", dict(mode=0, want_var=True), 5 * off + 3 * dia, 5.0 * R * R),
                                    ("diff gram", dict(mode=1, want_var=False), off + dia, 1.0 * R * R),
                                    ("level-0 sums", dict(mode=0, want_var=False), off + dia, 1.0 * R * R)):
        xx = x[:, :, :1] if name.startswith("level-0") else x
        ms = timed(lambda: nat.gram_accumulate(basis, xx, acc.level(0), **kw))
        ex = n * dmmas * 128.0 / (ms * 1e-3) / 1e12
        us = n * useful / (ms * 1e-3) / 1e12
        print("R=%3d %-13s %8.3f ms  executed %5.2f TFLOP/s (%.2f)  useful %5.2f TFLOP/s (%.2f)" % (
            R, name, ms, ex, ex / peak, us, us / peak), flush=True)

# vector quantity: all components in one launch (mlmcb200_gram_accumulate_comp) against one scalar launch per component
for M, R, nv in ((8, 25, 1_000_000), (64, 25, 100_000), (1000, 10, 4096)):
    basis = nat.make_basis(nat.LEGENDRE, R, (-3.72, 3.72), (-1.0, 1.0))
    rows_v = torch.randn(nv, 2, M, generator=g, device=dev, dtype=torch.float64)
    rows_v[:, 1] = rows_v[:, 0] + 0.1 * rows_v[:, 1]
    xv = rows_v.permute(2, 0, 1)
    acc_v = nat.LevelAccumulator(1, M * R * R, dev)
    acc_s = nat.LevelAccumulator(1, R * R, dev)
    valid = nat.sample_mask(basis, xv)
    ms_vec = timed(lambda: nat.gram_accumulate(basis, xv, acc_v.level(0), mode=0, want_var=True, valid=valid))

    def one_by_one():
        for m in range(M):
            nat.gram_accumulate(basis, xv[m:m + 1], acc_s.level(0), mode=0, want_var=True)
    ms_sep = timed(one_by_one)
    print("vector M=%4d R=%2d n=%7d sums+squares: one launch %8.3f ms, %d scalar launches %8.3f ms" % (
        M, R, nv, ms_vec, M, ms_sep), flush=True)
