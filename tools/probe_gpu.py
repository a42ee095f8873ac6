"""Quick on-GPU probe: FP64 pipe peaks and first kernel timings (CUDA events).  Writes gpurun_out/probe.json."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlmc_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
out = {"gpu": torch.cuda.get_device_name(0), "sms": nat.sm_count()}
out["dfma_tflops"] = nat.fp64_peak(0) / 1e12
out["dmma_tflops"] = nat.fp64_peak(1) / 1e12
print(out, flush=True)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def synth(n, h_f, h_c, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(n, generator=g, device=dev, dtype=torch.float64)
    root = torch.sqrt(1e-4 + x.abs())
    return torch.stack([x + h_f * root, x + h_c * root], dim=1).unsqueeze(2).contiguous()


for R in (5, 25, 50, 100):
    n = 10_000_000
    rows = synth(n, 0.05, 0.5, 1)
    basis = nat.make_basis(nat.LEGENDRE, R, (-3.72, 3.72), (-1.0, 1.0))
    acc = nat.LevelAccumulator(1, R, dev)
    x = rows.permute(2, 0, 1)
    med, best = timed(lambda: nat.moments_accumulate(basis, x, acc.level(0)))
    med0, best0 = timed(lambda: nat.moments_accumulate(basis, x[:, :, :1], acc.level(0)))
    out["moments_R%d" % R] = {"n": n, "ms_pair": med, "ms_pair_best": best, "ms_level0": med0,
                              "gsm_per_s": n * R / med / 1e6, "gbs": n * 16 / med / 1e6,
                              "tflops_7": n * R * 7 * 2 / med / 1e9 / 2}
    print(R, out["moments_R%d" % R], flush=True)

for kind, name in ((nat.MONOMIAL, "mono"), (nat.FOURIER, "four")):
    R = 32
    basis = nat.make_basis(kind, R, (-3.72, 3.72), (0.0, 1.0) if kind == nat.MONOMIAL else (0.0, 2 * np.pi))
    acc = nat.LevelAccumulator(1, R, dev)
    med, best = timed(lambda: nat.moments_accumulate(basis, x, acc.level(0)))
    out["moments_%s32" % name] = {"ms": med, "gsm_per_s": 10_000_000 * R / med / 1e6}
    print(name, out["moments_%s32" % name], flush=True)

# vector quantity cfg5-like: M = 1e4, n = 1024, Fourier 32
M, n = 10_000, 1024
g = torch.Generator(device=dev).manual_seed(3)
rows = torch.randn(n, 2, 1, generator=g, device=dev, dtype=torch.float64).expand(n, 2, M).contiguous()
rows += torch.arange(M, device=dev, dtype=torch.float64)[None, None, :] * 1e-4
basis = nat.make_basis(nat.FOURIER, 32, (-4.5, 5.5), (0.0, 2 * np.pi))
acc = nat.LevelAccumulator(1, 32 * M, dev)
xv = rows.permute(2, 0, 1)
med, best = timed(lambda: nat.moments_accumulate(basis, xv, acc.level(0)))
out["moments_vec_M1e4_four32"] = {"ms": med, "gsm_per_s": n * M * 32 / med / 1e6, "gbs": n * M * 16 / med / 1e6}
print("vec", out["moments_vec_M1e4_four32"], flush=True)
del rows, xv, acc

for R in (25, 50, 100):
    n = 2_000_000
    rows = synth(n, 0.05, 0.5, 2)
    basis = nat.make_basis(nat.LEGENDRE, R, (-3.72, 3.72), (-1.0, 1.0))
    acc = nat.LevelAccumulator(1, R * R, dev)
    x = rows.permute(2, 0, 1)
    nb = (R + 7) // 8
    blocks = nb * (nb + 1) // 2
    for want_var in (False, True):
        med, best = timed(lambda: nat.gram_accumulate(basis, x, acc.level(0), want_var=want_var), reps=3, warm=1)
        dmma_per_sample = blocks * (5 if want_var else 2) / 4
        out["gram_R%d_var%d" % (R, want_var)] = {
            "n": n, "ms": med, "msamples_per_s": n / med / 1e3,
            "dmma_tflops": n * dmma_per_sample * 512 / med / 1e9}
        print(R, want_var, out["gram_R%d_var%d" % (R, want_var)], flush=True)

# max-ent F/g/H at cfg4 size
Q, R = 100_002, 50
phi = torch.randn(Q, R, device=dev, dtype=torch.float64) * 0.1
w = torch.full((Q,), 1.0 / Q, device=dev, dtype=torch.float64)
lam = torch.randn(R, device=dev, dtype=torch.float64) * 0.1
res = torch.zeros(1 + R + R * R, device=dev, dtype=torch.float64)
med, best = timed(lambda: nat.maxent_fgh(phi, w, lam, 7, res), reps=10)
medg, _ = timed(lambda: nat.maxent_fgh(phi, w, lam, 3, res), reps=10)
out["maxent_fgh_Q1e5_R50"] = {"ms_fgh": med, "ms_fg": medg}
print(out["maxent_fgh_Q1e5_R50"], flush=True)
ref_h = (phi.T * (w * torch.exp(torch.clamp(-(phi @ lam), -200, 200)))) @ phi
print("maxent H max abs diff vs torch:", float((res[1 + R:].reshape(R, R) - ref_h).abs().max()))

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/probe.json", "w") as f:
    json.dump(out, f, indent=1)
