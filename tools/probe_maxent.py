"""Max-ent evaluation kernel: time per F+g+H launch (new kernel vs MLMCB200_MAXENT_V1=1) and a cProfile of one cfg4 fit."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import scipy.stats as stats
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlmc_b200 import _native as nat  # noqa: E402
from mlmc_b200.moments import Legendre  # noqa: E402
from mlmc_b200.tool.simple_distribution import (SimpleDistribution, construct_ortogonal_moments,  # noqa: E402
                                                compute_semiexact_cov, compute_semiexact_moments)

dev = torch.device("cuda:0")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for Q, R in ((100_002, 50), (100_002, 25), (5376, 20), (100_002, 100), (1_000_000, 50)):
    phi = torch.randn(Q, R, device=dev, dtype=torch.float64) * 0.1
    w = torch.full((Q,), 1.0 / Q, device=dev, dtype=torch.float64)
    lam = torch.randn(R, device=dev, dtype=torch.float64) * 0.1
    res = torch.zeros(1 + R + R * R, device=dev, dtype=torch.float64)
    ms = timed(lambda: nat.maxent_fgh(phi, w, lam, 7, res))
    ms_fg = timed(lambda: nat.maxent_fgh(phi, w, lam, 3, res))
    nat.maxent_fgh(phi, w, lam, 7, res)
    rho = w * torch.exp(torch.clamp(-(phi @ lam), -200, 200))
    h = (phi.T * rho) @ phi
    err_h = float((res[1 + R:].reshape(R, R) - h).abs().max() / h.abs().max())
    err_g = float((res[1:1 + R] - phi.T @ rho).abs().max())
    flop = 2.0 * Q * R * R + 4.0 * Q * R
    print("Q=%7d R=%3d: F+g+H %.4f ms (%.2f TFLOP/s useful)  F+g %.4f ms   rel err H %.1e, g %.1e, F %.1e" % (
        Q, R, ms, flop / ms / 1e9, ms_fg, err_h, err_g, abs(float(res[0] - rho.sum()))), flush=True)

distr = stats.norm(loc=1, scale=2)
dom = tuple(float(v) for v in distr.ppf([0.01, 0.99]))
n_panels = 4762
base = Legendre(50, dom, safe_eval=False)
cov = compute_semiexact_cov(base, distr.pdf, n_panels=n_panels)
orth, info = construct_ortogonal_moments(base, cov, tol=1e-4)
mu = compute_semiexact_moments(orth, distr.pdf, n_panels=n_panels)
data = np.stack([mu, np.ones_like(mu)], axis=1)


def fit():
    sd = SimpleDistribution(orth, data, domain=dom, quad_panels=n_panels)
    return sd, sd.estimate_density_minimize(tol=1e-8, reg_param=0.0)


for _ in range(3):
    fit()
ts = []
for _ in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sd, res_ = fit()
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print("cfg4 fit: median %.3f ms, min %.3f ms, nit %d, evals %d" % (np.median(ts), np.min(ts), res_.nit, sd.n_device_evals))
if "--profile" in sys.argv:
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        fit()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(40)
