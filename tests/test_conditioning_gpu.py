"""Conditioning of the level sums at large N (north_star: "Welford / Kahan-compensated accumulation"; SURVEY.md section 7
"parity vs a numerically naive reference").

The kernels accumulate PLAIN sums (short per-thread runs, then fixed-order pairwise merges: O(eps log N) error at no
extra FP64 instruction) and the level variance is the reference's own formula ``(sp - s^2 / n) / (n - 1)``
(mlmc/quantity/quantity_estimate.py:75), which loses ``mean^2 / var`` of the precision whatever the summation does.
Here both are measured against an extended-precision result:

* truth            : two passes -- the mean from ``np.longdouble`` accumulation of fp64 chunk sums, then the sums of the
                     CENTRED values ``d - mean`` and their squares (no cancellation left), again accumulated in
                     ``np.longdouble``, with the exact correction for the rounding of the centre;
* reference formula: fp64 sums over 65 536-row chunks added up chunk by chunk, then ``(sp - s^2 / n) / (n - 1)`` -- what
                     the reference's chunk loop does (quantity_estimate.py:43-77);
* GPU              : ``mlmcb200_moments_accumulate`` + ``mlmcb200_finalize_levels``.

Cases: (A) the ill-conditioned one, Monomial R = 8 of an offset variable (|mean| >> std), N = 1e8, one level;
(B) Legendre R = 100 of fine - coarse pairs at N = 1.25e8 (one GPU's share of cfg3).
The GPU must be at least as close to the truth as the reference formula is (up to a factor 2 for the luck of the
rounding), and within 1e-10 wherever the formula itself is.  The measured deviations go to gpurun_out/ (DESIGN.md 4.7).
"""
import json
import os

import numpy as np
import pytest
import torch

from test_kernels_gpu import dev, native

pytestmark = pytest.mark.gpu
LD = np.longdouble


def _differences(nat, basis, rows, R):
    """d [n', R] of the kept samples of a chunk of rows [n, S, 1] (bit-exact tables from mlmcb200_basis_eval)."""
    pf = nat.basis_eval(basis, rows[:, 0, 0].contiguous(), R)
    d = pf if rows.shape[1] == 1 else pf - nat.basis_eval(basis, rows[:, 1, 0].contiguous(), R)
    return d[~torch.isnan(d).any(dim=1)]


def _truth_and_formula(nat, basis, rows, R, chunk):
    n_rows = rows.shape[0]
    s = np.zeros(R, dtype=LD)
    n = 0
    s64 = torch.zeros(R, dtype=torch.float64, device=dev())          # the reference formula's running fp64 sums
    sp64 = torch.zeros(R, dtype=torch.float64, device=dev())
    for lo in range(0, n_rows, chunk):
        d = _differences(nat, basis, rows[lo:lo + chunk], R)
        n += d.shape[0]
        s += d.sum(dim=0).cpu().numpy().astype(LD)
        for lo2 in range(0, d.shape[0], 65536):
            part = d[lo2:lo2 + 65536]
            s64 += part.sum(dim=0)
            sp64 += (part * part).sum(dim=0)
    mean = s / LD(n)
    centre = mean.astype(np.float64)
    delta = mean - centre.astype(LD)                                # rounding of the centre, known exactly
    c_dev = torch.from_numpy(centre).to(dev())
    sr, srr = np.zeros(R, dtype=LD), np.zeros(R, dtype=LD)
    for lo in range(0, n_rows, chunk):
        r = _differences(nat, basis, rows[lo:lo + chunk], R) - c_dev
        sr += r.sum(dim=0).cpu().numpy().astype(LD)
        srr += (r * r).sum(dim=0).cpu().numpy().astype(LD)
    # sum (d - mean)^2 = sum (r - delta)^2
    m2 = srr - 2 * delta * sr + LD(n) * delta * delta
    truth_var = (m2 / LD(n - 1)).astype(np.float64)
    s_h, sp_h = s64.cpu().numpy(), sp64.cpu().numpy()
    formula_mean = s_h / n
    formula_var = (sp_h - s_h ** 2 / n) / (n - 1)
    return n, mean.astype(np.float64), truth_var, formula_mean, formula_var


def _rel(a, b, floor):
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def _run_case(name, basis, rows, R, chunk, report):
    nat = native()
    x = rows.permute(2, 0, 1)
    acc = nat.LevelAccumulator(1, R, dev())
    nat.moments_accumulate(basis, x, acc.level(0))
    fin = acc.finalize()
    gpu_mean, gpu_var = fin["l_means"][0].cpu().numpy(), fin["l_vars"][0].cpu().numpy()
    n, t_mean, t_var, f_mean, f_var = _truth_and_formula(nat, basis, rows, R, chunk)
    assert int(acc.acc[0, 0].item()) == n
    scale_m, scale_v = np.abs(t_mean).max(), np.abs(t_var).max()
    k = slice(1, None)                                              # moment 0 is exact (mean 1, variance 0)
    out = {"n": n, "R": R,
           "mean_rel_err_gpu": float(_rel(gpu_mean, t_mean, 1e-3 * scale_m)[k].max()),
           "mean_rel_err_reference_formula": float(_rel(f_mean, t_mean, 1e-3 * scale_m)[k].max()),
           "var_rel_err_gpu": float(_rel(gpu_var, t_var, 1e-6 * scale_v)[k].max()),
           "var_rel_err_reference_formula": float(_rel(f_var, t_var, 1e-6 * scale_v)[k].max()),
           "var_rel_diff_gpu_vs_reference_formula": float(_rel(gpu_var, f_var, 1e-6 * scale_v)[k].max()),
           "mean2_over_var_max": float((t_mean[k] ** 2 / np.maximum(t_var[k], 1e-300)).max())}
    report[name] = out
    assert gpu_mean[0] == 1.0 and gpu_var[0] == 0.0 if rows.shape[1] == 1 else gpu_var[0] == 0.0
    assert out["mean_rel_err_gpu"] <= max(2 * out["mean_rel_err_reference_formula"], 1e-13), out
    assert out["var_rel_err_gpu"] <= max(2 * out["var_rel_err_reference_formula"], 1e-12), out
    return out


def test_conditioning_against_extended_precision():
    from mlmc_b200.moments import Monomial, Legendre
    report = {}
    # (A) offset variable: t = x / 20 = 0.5 +- 5e-4, monomials t^k: mean^2 / var up to ~1e6
    n = 100_000_000
    g = torch.Generator(device=dev()).manual_seed(11)
    rows = (10.0 + 0.01 * torch.randn(n, generator=g, device=dev(), dtype=torch.float64)).reshape(n, 1, 1)
    a = _run_case("A_monomial8_offset_1e8", Monomial(8, (0.0, 20.0)).basis_struct(), rows, 8, 4_000_000, report)
    del rows
    # the formula is the limit, not the sums: both lose ~mean^2/var digits
    assert a["mean2_over_var_max"] > 1e5
    # (B) bounded basis, level differences: well conditioned, 1e-10 holds with a wide margin
    n = 125_000_000
    g = torch.Generator(device=dev()).manual_seed(77)
    x = torch.randn(n, generator=g, device=dev(), dtype=torch.float64)
    root = torch.sqrt(1e-4 + x.abs())
    rows = torch.stack([x + 0.05 * root, x + 0.5 * root], dim=1).unsqueeze(2).contiguous()
    del x, root
    b = _run_case("B_legendre100_pairs_1.25e8", Legendre(100, (-3.719016485455709, 3.719016485455709)).basis_struct(),
                  rows, 100, 1_000_000, report)
    assert b["var_rel_err_gpu"] < 1e-10 and b["mean_rel_err_gpu"] < 1e-10
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "r2_conditioning.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report, indent=1))
