"""Run the five BASELINE.json configs on one B200 through the public API, with a parity check against the oracle on
a bounded slice and the oracle's CPU time beside each.  Writes gpurun_out/configs.json (copy to profiles/).

  cfg1  1 level, 1e5 lognormal samples, Legendre R=25 (log domain), estimate_moments
  cfg2  3 levels x 1e7, Legendre R=50, mean + var + regression + n_samples          (bench.py headline)
  cfg3  covariance, Legendre R=100; 1.25e8 samples = the per-GPU share of 1e9 over 8 GPUs (means + entry variances)
  cfg4  max-ent fit, 50 moments, 4762 x 21 = 100 002 Gauss nodes
  cfg5  vector quantity: 1e4 locations x 5 levels, Fourier R=32, per-location mean / var
"""
import json
import os
import sys
import time

import numpy as np
import scipy.stats as stats
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import mlmc_oracle as orc  # noqa: E402
from mlmc_b200 import _native as nat  # noqa: E402
from mlmc_b200.moments import Legendre, Fourier  # noqa: E402
from mlmc_b200.sample_storage import Memory  # noqa: E402
from mlmc_b200.quantity.quantity import make_root_quantity  # noqa: E402
from mlmc_b200.quantity.quantity_spec import QuantitySpec  # noqa: E402
from mlmc_b200.quantity import quantity_estimate as qe  # noqa: E402
from mlmc_b200.estimator import Estimate, estimate_n_samples_for_target_variance  # noqa: E402
from mlmc_b200.tool.simple_distribution import (SimpleDistribution, construct_ortogonal_moments,  # noqa: E402
                                                compute_semiexact_cov, compute_semiexact_moments)

dev = torch.device("cuda:0")
only = set(sys.argv[1:])
out = {"gpu": torch.cuda.get_device_name(0), "host_cores": os.cpu_count()}


def want(name):
    return not only or name in only


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        res = fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), res


def max_rel(a, b, floor=1e-300):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = np.max(np.abs(b)) + floor
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-12 * scale)))


def scalar_quantity(levels, steps, n_ops=None):
    spec = [QuantitySpec(name="v", unit="", shape=(1, 1), times=[0.0], locations=["0"])]
    storage = Memory.from_arrays(levels, level_parameters=[[h] for h in steps], n_ops=n_ops, result_format=spec)
    return storage, make_root_quantity(storage, spec)["v"][0.0]["0"][0, 0]


def synth_device(n, h_f, h_c, seed, distr="norm"):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(n, generator=g, device=dev, dtype=torch.float64)
    if distr == "lognorm":
        x = torch.exp(x)
    root = torch.sqrt(1e-4 + x.abs())
    coarse = x + h_c * root if h_c else torch.zeros_like(x)
    return torch.stack([x + h_f * root, coarse], dim=1).unsqueeze(2).contiguous()


# ------------------------------------------------------------------------------------------------ cfg1
if want("cfg1"):
    rows = synth_device(100_000, 0.1, None, 1234, "lognorm").cpu().numpy()
    storage, value = scalar_quantity([rows], [0.1])
    dom = Estimate.estimate_domain(value, storage, quantile=0.001)
    fn = Legendre(25, dom, log=True, safe_eval=True)
    est = Estimate(value, storage, fn)
    t_gpu, (means, variances) = timed(lambda: est.estimate_moments())
    t0 = time.perf_counter()
    o = orc.estimate_moments([rows], orc.Basis("legendre", 25, tuple(dom), log=True))
    t_cpu = time.perf_counter() - t0
    out["cfg1"] = {"gpu_ms": t_gpu * 1e3, "cpu_ms": t_cpu * 1e3, "sample_moments_per_s": 1e5 * 25 / t_gpu,
                   "cpu_sample_moments_per_s": 1e5 * 25 / t_cpu, "max_rel_mean": max_rel(means, o.mean),
                   "max_rel_var": max_rel(variances, o.var), "n_rm": int(o.n_rm_samples[0]),
                   "domain_equal": bool(np.array_equal(np.array(dom), np.array(orc.estimate_domain([rows], 0.001))))}
    print("cfg1", out["cfg1"], flush=True)

# ------------------------------------------------------------------------------------------------ cfg2
if want("cfg2"):
    n = 10_000_000
    steps = orc.level_steps(3, (0.5, 0.005))
    n_ops = [orc.synth_n_ops(h) for h in steps]
    levels = [synth_device(n, steps[l], steps[l - 1] if l else None, 1234 + 1000 * l).cpu() for l in range(3)]
    storage, value = scalar_quantity(levels, steps, n_ops)
    domain = tuple(stats.norm.ppf([1e-4, 1 - 1e-4]))
    fn = Legendre(50, domain)
    est = Estimate(value, storage, fn)

    def run():
        variances, ops = est.estimate_diff_vars_regression(None)
        return variances, estimate_n_samples_for_target_variance(1e-5, variances, ops, 3)
    t_res, (reg, n_est) = timed(run)                       # levels resident in HBM after the first call
    storage.resident_fraction = 0.0
    storage.drop_device_copies()
    t_host, _ = timed(run)                                 # every call streams the 480 MB from pinned host memory
    sl = [lv[:200_000].numpy() for lv in levels]
    st2, v2 = scalar_quantity(sl, steps, n_ops)
    qm = qe.estimate_mean(qe.moments(v2, fn))
    t0 = time.perf_counter()
    o = orc.estimate_moments(sl, orc.Basis("legendre", 50, domain))
    t_cpu = time.perf_counter() - t0
    out["cfg2"] = {"resident_ms": t_res * 1e3, "host_staged_ms": t_host * 1e3,
                   "sample_moments_per_s_resident": 3 * n * 50 / t_res, "sample_moments_per_s_host": 3 * n * 50 / t_host,
                   "cpu_sample_moments_per_s_1proc": 3 * 200_000 * 50 / t_cpu, "n_estimated": [int(v) for v in n_est],
                   "slice_max_rel_l_means": max_rel(qm.l_means, o.l_means), "slice_max_rel_l_vars": max_rel(qm.l_vars, o.l_vars)}
    print("cfg2", out["cfg2"], flush=True)
    del levels, storage

# ------------------------------------------------------------------------------------------------ cfg3
if want("cfg3"):
    n = 125_000_000                                         # 1e9 / 8 GPUs
    domain = tuple(stats.norm.ppf([1e-4, 1 - 1e-4]))
    fn = Legendre(100, domain)
    basis = fn.basis_struct()
    rows = synth_device(n, 0.05, 0.5, 77)                   # 2 GB, generated on the device
    x = rows.permute(2, 0, 1)
    acc = nat.LevelAccumulator(1, 100 * 100, dev)

    def run(want_var):
        acc.acc.zero_()
        nat.gram_accumulate(basis, x, acc.level(0), want_var=want_var)
        return acc.finalize()
    t_mean, _ = timed(lambda: run(False), reps=2)
    t_var, res = timed(lambda: run(True), reps=2)
    # fused moments (mean/var of the 100 moments) on the same data
    acc_m = nat.LevelAccumulator(1, 100, dev)

    def run_m():
        acc_m.acc.zero_()
        nat.moments_accumulate(basis, x, acc_m.level(0))
        return acc_m.finalize()
    t_mom, _ = timed(run_m, reps=3)
    # parity on a 20k-sample slice against the oracle
    sl = rows[:20_000].cpu().numpy()
    acc_s = nat.LevelAccumulator(2, 100 * 100, dev)
    nat.gram_accumulate(basis, rows[:20_000].permute(2, 0, 1), acc_s.level(1), want_var=True)
    nat.gram_accumulate(basis, torch.zeros(1, 1, 1, dtype=torch.float64, device=dev), acc_s.level(0), want_var=True)
    fin = acc_s.finalize()
    t0 = time.perf_counter()
    o = orc.estimate_covariance([np.zeros((1, 2, 1)), sl], orc.Basis("legendre", 100, domain), chunk_rows=2048)
    t_cpu = time.perf_counter() - t0
    nb = 13
    out["cfg3"] = {"samples_per_gpu": n, "cov_mean_ms": t_mean * 1e3, "cov_mean_var_ms": t_var * 1e3,
                   "moments_meanvar_ms": t_mom * 1e3,
                   "samples_per_s_mean": n / t_mean, "samples_per_s_mean_var": n / t_var,
                   "sample_moments_per_s_cov_mean": n * 100 / t_mean, "sample_moments_per_s_moments": n * 100 / t_mom,
                   "dmma_tflops_mean": n * (nb * (nb + 1) // 2) * 2 / 4 * 512 / t_mean / 1e12,
                   "dmma_tflops_mean_var": n * (nb * (nb + 1) // 2) * 5 / 4 * 512 / t_var / 1e12,
                   "cpu_samples_per_s_1proc": 20_000 / t_cpu,
                   "slice_max_rel_cov_mean": max_rel(fin["l_means"][1].cpu().numpy(), o.l_means[1]),
                   "slice_max_rel_cov_var": max_rel(fin["l_vars"][1].cpu().numpy(), o.l_vars[1])}
    print("cfg3", out["cfg3"], flush=True)
    del rows, x

# ------------------------------------------------------------------------------------------------ cfg4
if want("cfg4"):
    distr = stats.norm(loc=1, scale=2)
    domain = tuple(distr.ppf([0.01, 0.99]))
    n_panels = 4762
    base = Legendre(50, domain, safe_eval=False)
    cov = compute_semiexact_cov(base, distr.pdf, n_panels=n_panels)
    orth, info = construct_ortogonal_moments(base, cov, tol=1e-4)
    mu = compute_semiexact_moments(orth, distr.pdf, n_panels=n_panels)
    data = np.stack([mu, np.ones_like(mu)], axis=1)

    def fit():
        sd = SimpleDistribution(orth, data, domain=domain, quad_panels=n_panels)
        res = sd.estimate_density_minimize(tol=1e-8, reg_param=0.0)
        return sd, res
    t_fit, (sd, res) = timed(fit, reps=5)
    xs = np.linspace(domain[0], domain[1], 201)
    pdf_err = float(np.max(np.abs(sd.density(xs) - distr.pdf(xs) / (0.98))))
    t0 = time.perf_counter()
    ofit = orc.maxent_fit(orc.Basis("legendre", 50, domain, safe_eval=False, matrix=info[2]), data, domain, tol=1e-8,
                          n_panels=n_panels)
    t_cpu = time.perf_counter() - t0
    out["cfg4"] = {"fit_ms": t_fit * 1e3, "cpu_fit_ms": t_cpu * 1e3, "nit": int(res.nit), "success": bool(res.success),
                   "device_evals": sd.n_device_evals, "n_moments": int(orth.size), "nodes": n_panels * 21,
                   "max_abs_multiplier_diff_vs_oracle": float(np.max(np.abs(sd.multipliers - ofit.multipliers))),
                   "max_rel_pdf_vs_oracle": max_rel(sd.density(xs), orc.maxent_density(
                       orc.Basis("legendre", 50, domain, safe_eval=False, matrix=info[2]), ofit.multipliers,
                       np.ones(len(mu)), xs)), "max_abs_pdf_err_vs_truncated_normal": pdf_err}
    print("cfg4", out["cfg4"], flush=True)

# ------------------------------------------------------------------------------------------------ cfg4 (data-driven)
if want("cfg4"):
    # Estimate.construct_density on cfg2-shaped samples: covariance pass (DMMA) -> orthogonalisation -> moments of
    # the orthogonal basis from the level sums -> max-ent fit; the whole call, levels resident in HBM
    n = 10_000_000
    steps = orc.level_steps(3, (0.5, 0.005))
    levels = [synth_device(n, steps[l], steps[l - 1] if l else None, 1234 + 1000 * l).cpu() for l in range(3)]
    storage, value = scalar_quantity(levels, steps, [orc.synth_n_ops(h) for h in steps])
    dom = tuple(stats.norm.ppf([0.001, 0.999]))
    xs = np.linspace(dom[0], dom[1], 201)
    inner = slice(10, -10)                                  # the fit has a boundary layer in the outer 5 % of the domain
    out["cfg4_data_driven"] = {"samples": 3 * n}
    for r_base in (12, 25):
        est = Estimate(value, storage, Legendre(r_base, dom))
        t_cd, (dobj, info_d, res_d, mom_d) = timed(lambda: est.construct_density(tol=1e-8, orth_moments_tol=1e-4),
                                                   reps=3)
        err = np.abs(dobj.density(xs) - stats.norm.pdf(xs))
        out["cfg4_data_driven"]["legendre_%d" % r_base] = {
            "construct_density_ms": t_cd * 1e3, "orthogonal_moments": int(mom_d.size), "success": bool(res_d.success),
            "fun_norm": float(res_d.fun_norm), "nit": int(res_d.nit),
            "max_abs_pdf_err_vs_normal_interior": float(np.max(err[inner])), "max_abs_pdf_err_vs_normal": float(np.max(err))}
    print("cfg4_data_driven", out["cfg4_data_driven"], flush=True)
    del levels, storage, value, est

# ------------------------------------------------------------------------------------------------ cfg5
if want("cfg5"):
    M = 10_000
    n_levels = [4096, 2048, 1024, 512, 256]
    steps = orc.level_steps(5, (0.5, 0.005))
    rng = np.random.default_rng(5)
    levels = []
    for l, n in enumerate(n_levels):
        base_rows = orc.synth_level_rows(rng.normal(size=n), steps[l], steps[l - 1] if l else None)     # [n, 2, 1]
        rows = np.repeat(base_rows, M, axis=2) + (np.arange(M) * 1e-4)[None, None, :]
        if l == 0:
            rows[:, 1, :] = 0
        levels.append(rows)
    spec = [QuantitySpec(name="field", unit="", shape=(1, 1), times=[0.0], locations=[str(i) for i in range(M)])]
    storage = Memory.from_arrays(levels, level_parameters=[[h] for h in steps], result_format=spec)
    field = make_root_quantity(storage, spec)["field"][0.0]
    fn = Fourier(32, (-4.2, 5.4))
    t_gpu, qm = timed(lambda: qe.estimate_mean(qe.moments(field, fn)), reps=3)
    storage.resident_fraction = 0.0
    storage.drop_device_copies()
    t_host, _ = timed(lambda: qe.estimate_mean(qe.moments(field, fn)), reps=5)
    n_loc = 50
    # oracle on the first 50 locations, with the sample mask of ALL locations (a sample is dropped if any of the
    # 1e4 locations leaves the domain): remove those samples from the slice first
    ob = orc.Basis("fourier", 32, (-4.2, 5.4))
    sl = []
    for l, lv in enumerate(levels):
        t = orc.to_ref_domain(ob, lv[:, :1, :] if l == 0 else lv)
        keep = ~np.isnan(t).any(axis=(1, 2))
        sl.append(lv[keep][:, :, :n_loc])
    t0 = time.perf_counter()
    o = orc.estimate_moments(sl, orc.Basis("fourier", 32, (-4.2, 5.4)), chunk_rows=512)
    t_cpu = time.perf_counter() - t0
    got_means = qm.l_means.reshape(5, M, 32)[:, :n_loc].reshape(5, -1)
    got_vars = qm.l_vars.reshape(5, M, 32)[:, :n_loc].reshape(5, -1)
    units = sum(n_levels) * M * 32
    # the oracle masks on the 50-location slice only; compare where the masks agree (no sample dropped here)
    out["cfg5"] = {"resident_ms": t_gpu * 1e3, "host_staged_ms": t_host * 1e3, "sample_moments_per_s_resident": units / t_gpu,
                   "sample_moments_per_s_host": units / t_host, "bytes": sum(n_levels) * M * 16,
                   "cpu_sample_moments_per_s_1proc": sum(n_levels) * n_loc * 32 / t_cpu,
                   "n_rm": [int(v) for v in qm.n_rm_samples], "n_samples": [int(v) for v in qm.n_samples],
                   "oracle_n_samples": [int(v) for v in o.n_samples],
                   "slice_max_rel_l_means": max_rel(got_means, o.l_means), "slice_max_rel_l_vars": max_rel(got_vars, o.l_vars)}
    print("cfg5", out["cfg5"], flush=True)

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/configs.json", "w") as f:
    json.dump(out, f, indent=1)
