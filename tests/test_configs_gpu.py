"""The five BASELINE.json configs through the public API on the CUDA path, each checked against the oracle (the
restated reference algorithm) on the same inputs: full size where the oracle finishes in seconds (cfg1, cfg4, cfg5's
kept locations), a slice plus size-independent properties where it would take minutes to hours (cfg3).

Tolerances (BASELINE.json north_star): moment means / variances rel <= 1e-10 (scaled per level), covariance <= 1e-8,
max-ent multipliers and PDF values <= 1e-6; sample counts exact.
"""
import numpy as np
import pytest
import scipy.stats as stats
import torch

from oracle import mlmc_oracle as orc
from test_api_gpu import scalar_setup
from test_kernels_gpu import dev, native, rel_close, to_struct

pytestmark = pytest.mark.gpu


def _synth(n, h_f, h_c, seed, lognormal=False):
    g = torch.Generator(device=dev()).manual_seed(seed)
    x = torch.randn(n, generator=g, device=dev(), dtype=torch.float64)
    if lognormal:
        x = torch.exp(x)
    root = torch.sqrt(1e-4 + x.abs())
    coarse = torch.zeros_like(x) if h_c is None else x + h_c * root
    return torch.stack([x + h_f * root, coarse], dim=1).unsqueeze(2).contiguous()          # [n, 2, 1]


# ------------------------------------------------------------------------------------------------ cfg1
def test_cfg1_lognormal_legendre25_log_domain():
    """1 level, 1e5 lognormal samples, Legendre(25, log=True) on the estimated 0.001 quantile domain: everything
    (domain, counts, means, variances) against the oracle on the full input."""
    from mlmc_b200.moments import Legendre
    from mlmc_b200.estimator import Estimate
    rows = _synth(100_000, 0.1, None, 1234, lognormal=True).cpu().numpy()
    storage, value = scalar_setup([rows], [[0.1]])
    dom = Estimate.estimate_domain(value, storage, quantile=0.001)
    assert np.array_equal(np.array(dom), np.array(orc.estimate_domain([rows], 0.001)))     # bit-identical percentiles
    fn = Legendre(25, dom, log=True, safe_eval=True)
    est = Estimate(value, storage, fn)
    means, variances = est.estimate_moments()
    l_vars, n_samples = est.estimate_diff_vars()
    o = orc.estimate_moments([rows], orc.Basis("legendre", 25, tuple(dom), log=True))
    assert means[0] == 1.0 and variances[0] == 0.0
    assert list(n_samples) == list(o.n_samples) and int(o.n_rm_samples[0]) > 0
    rel_close(means, o.mean, rtol=1e-10, atol_scale=1e-14)
    rel_close(variances, o.var, rtol=1e-10, atol_scale=1e-14)
    rel_close(l_vars, o.l_vars, rtol=1e-10, atol_scale=1e-14)


# ------------------------------------------------------------------------------------------------ cfg3
def test_cfg3_covariance_legendre100_slice_vs_oracle():
    """Legendre(100) covariance of a 2-level 20 000 + 20 000 sample slice of the cfg3 data: DMMA kernel (means +
    entry variances) and the linearised means, both against the oracle's per-sample outer products."""
    from mlmc_b200.moments import Legendre
    from mlmc_b200.estimator import Estimate
    domain = tuple(stats.norm.ppf([1e-4, 1 - 1e-4]))
    rows0 = _synth(20_000, 0.5, None, 76).cpu().numpy()
    rows1 = _synth(20_000, 0.05, 0.5, 77).cpu().numpy()
    storage, value = scalar_setup([rows0, rows1], [[0.5], [0.05]])
    est = Estimate(value, storage, Legendre(100, domain))
    cov_mean, cov_var = est.estimate_covariance()
    lin_mean, lin_var = est.estimate_covariance(variance=False)
    o = orc.estimate_covariance([rows0, rows1], orc.Basis("legendre", 100, domain), chunk_rows=2048)
    want = o.mean.reshape(100, 100)
    rel_close(cov_mean, want, rtol=1e-8, atol_scale=1e-12)
    rel_close(lin_mean, want, rtol=1e-8, atol_scale=1e-12)
    rel_close(cov_var, o.var.reshape(100, 100), rtol=1e-8, atol_scale=1e-12)
    assert np.isnan(lin_var).all()
    assert np.array_equal(cov_mean, cov_mean.T)
    assert np.abs(lin_mean - lin_mean.T).max() <= 1e-15 * np.abs(lin_mean).max()


def test_cfg3_full_size_linearised_equals_dmma():
    """1.25e8 samples (one GPU's share of 1e9), Legendre(100): the covariance level sums from the 199 moment sums
    (C . s) equal the DMMA contraction, same sample counts; first column = the moment sums themselves."""
    from mlmc_b200.moments import Legendre
    nat = native()
    n, R = 125_000_000, 100
    domain = tuple(stats.norm.ppf([1e-4, 1 - 1e-4]))
    rows = _synth(n, 0.05, 0.5, 77)
    x = rows.permute(2, 0, 1)
    fn = Legendre(R, domain)
    ext_fn, c_t = fn.product_table()
    assert ext_fn.size == 2 * R - 1
    mom = nat.LevelAccumulator(1, ext_fn.size, dev())
    nat.moments_accumulate(ext_fn.basis_struct(), x, mom.level(0), sums_only=True)
    lin = nat.level_sums_transform(mom, 1, torch.from_numpy(c_t).to(dev()))
    cov = nat.LevelAccumulator(1, R * R, dev())
    nat.gram_accumulate(fn.basis_struct(), x, cov.level(0), mode=0, want_var=False)
    a, b = lin.acc[0].cpu().numpy(), cov.acc[0].cpu().numpy()
    assert a[0] == b[0] and a[1] == b[1] and a[0] + a[1] == n
    ga, gb = a[2:2 + R * R].reshape(R, R), b[2:2 + R * R].reshape(R, R)
    scale = np.abs(gb).max()
    assert np.abs(ga - gb).max() < 1e-10 * scale, np.abs(ga - gb).max() / scale
    assert np.array_equal(ga[:, 0], mom.acc[0, 2:2 + R].cpu().numpy())
    assert np.isnan(a[2 + R * R:]).all()


# ------------------------------------------------------------------------------------------------ cfg4
def test_cfg4_maxent_50_moments_100002_nodes():
    """Max-ent fit for 50 moments of norm(1, 2) cut at q = 0.01 on the fixed 4762 x 21 = 100 002 node rule: multipliers,
    PDF and CDF against the oracle's NumPy fit on the same rule (<= 1e-6)."""
    from mlmc_b200.moments import Legendre
    from mlmc_b200.tool.simple_distribution import (SimpleDistribution, construct_ortogonal_moments,
                                                    compute_semiexact_cov, compute_semiexact_moments)
    distr = stats.norm(loc=1, scale=2)
    domain = tuple(distr.ppf([0.01, 0.99]))
    n_panels = 4762
    base = Legendre(50, domain, safe_eval=False)
    ob = orc.Basis("legendre", 50, domain, safe_eval=False)
    cov = compute_semiexact_cov(base, distr.pdf, n_panels=n_panels)
    nodes, w = orc.gauss_panels(domain, n_panels)
    phi = orc.basis_eval(ob, nodes)
    rel_close(cov, (phi.T * (distr.pdf(nodes) * w)) @ phi, rtol=1e-10, atol_scale=1e-13)
    orth, info = construct_ortogonal_moments(base, cov, tol=1e-4)
    mu = compute_semiexact_moments(orth, distr.pdf, n_panels=n_panels)
    data = np.stack([mu, np.ones_like(mu)], axis=1)
    sd = SimpleDistribution(orth, data, domain=domain, quad_panels=n_panels)
    res = sd.estimate_density_minimize(tol=1e-8, reg_param=0.0)
    assert res.success and res.fun_norm < 1e-7
    ob_t = orc.Basis("legendre", 50, domain, safe_eval=False, matrix=info[2])
    ofit = orc.maxent_fit(ob_t, data, domain, tol=1e-8, n_panels=n_panels)
    rel_close(sd.multipliers, ofit.multipliers, rtol=1e-6, atol_scale=1e-8)
    xs = np.linspace(domain[0], domain[1], 401)
    rel_close(sd.density(xs), orc.maxent_density(ob_t, ofit.multipliers, np.ones(len(mu)), xs), rtol=1e-6)
    # and the fit is the truncated normal renormalised on the domain, up to what the kept moments resolve
    assert np.max(np.abs(sd.density(xs) - distr.pdf(xs) / 0.98)) < 1e-2
    assert sd.n_device_evals <= 3 * (res.nit + 2)


# ------------------------------------------------------------------------------------------------ cfg5
def test_cfg5_field_1e4_locations_fourier32():
    """Vector quantity, 1e4 locations x 5 levels, Fourier(32): per-location mean / variance of 50 locations spread over
    the field against the oracle, with the sample mask of ALL locations (a sample is dropped when any location leaves
    the domain); counts exact."""
    from mlmc_b200.moments import Fourier
    from mlmc_b200.sample_storage import Memory
    from mlmc_b200.quantity.quantity import make_root_quantity
    from mlmc_b200.quantity.quantity_spec import QuantitySpec
    from mlmc_b200.quantity import quantity_estimate as qe
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
    dom = (-3.2, 4.4)
    qm = qe.estimate_mean(qe.moments(field, Fourier(32, dom)))
    ob = orc.Basis("fourier", 32, dom)
    pick = np.arange(0, M, 200)                                        # 50 locations across the field
    sl = []
    for l, lv in enumerate(levels):
        t = orc.to_ref_domain(ob, lv[:, :1, :] if l == 0 else lv)
        keep = ~np.isnan(t).any(axis=(1, 2))
        sl.append(lv[keep][:, :, pick])
        assert qm.n_samples[l] == int(keep.sum()) and qm.n_rm_samples[l] == int((~keep).sum())
    assert sum(qm.n_rm_samples) > 0
    o = orc.estimate_moments(sl, ob, chunk_rows=512)
    got_means = qm.l_means.reshape(5, M, 32)[:, pick].reshape(5, -1)
    got_vars = qm.l_vars.reshape(5, M, 32)[:, pick].reshape(5, -1)
    rel_close(got_means, o.l_means, rtol=1e-10, atol_scale=1e-13, per_level=True)
    rel_close(got_vars, o.l_vars, rtol=1e-10, atol_scale=1e-13, per_level=True)
    rel_close(qm.mean.reshape(M, 32)[pick].reshape(-1), o.mean, rtol=1e-10, atol_scale=1e-13)
    assert np.all(qm.mean.reshape(M, 32)[:, 0] == 1.0)


# ------------------------------------------------------------------------------------------------ linearised covariance
def test_linearised_covariance_matches_reference_golden(golden):
    """estimate_covariance(variance=False) on the golden 3-level case: Legendre / Monomial / Fourier / transformed and a
    vector quantity against the stored REFERENCE outputs (<= 1e-8)."""
    from mlmc_b200.moments import Legendre, Monomial, Fourier, TransformedMoments
    from mlmc_b200.quantity import quantity_estimate as qe
    g = golden("estimates")
    levels = [g["A_rows%d" % l] for l in range(3)]
    storage, value = scalar_setup(levels, [[h] for h in g["A_steps"]])
    domain = tuple(g["A_domain"])
    cov8 = qe.estimate_mean(qe.covariance(value, Legendre(8, domain)), variance=False)
    rel_close(cov8.mean, g["A_cov_mean"], rtol=1e-8, atol_scale=1e-13)
    assert np.isnan(cov8.var).all() and np.isnan(cov8.l_vars).all()
    full = qe.estimate_mean(qe.covariance(value, Legendre(8, domain)))
    assert np.array_equal(cov8.n_samples, full.n_samples) and np.array_equal(cov8.n_rm_samples, full.n_rm_samples)
    rel_close(cov8.l_means.reshape(3, -1), full.l_means.reshape(3, -1), rtol=1e-9, atol_scale=1e-13, per_level=True)
    rel_close(qe.estimate_mean(qe.covariance(value, Legendre(10, domain)), variance=False).mean, g["A_cov10_mean"],
              rtol=1e-8, atol_scale=1e-13)
    for fn, ob in ((Monomial(6, domain), orc.Basis("monomial", 6, domain)),
                   (Fourier(7, domain), orc.Basis("fourier", 7, domain)),
                   (Fourier(8, domain), orc.Basis("fourier", 8, domain))):
        got = qe.estimate_mean(qe.covariance(value, fn), variance=False)
        want = orc.estimate_covariance(levels, ob)
        rel_close(got.l_means.reshape(want.l_means.shape), want.l_means, rtol=1e-8, atol_scale=1e-13, per_level=True)
        assert np.array_equal(got.n_samples, want.n_samples)
    # transformed basis: covariance of L phi
    tm = TransformedMoments(Legendre(10, domain), g["A_orth_L"])
    got = qe.estimate_mean(qe.covariance(value, tm), variance=False)
    want = orc.estimate_covariance(levels, orc.Basis("legendre", 10, domain, matrix=g["A_orth_L"]))
    rel_close(got.l_means.reshape(want.l_means.shape), want.l_means, rtol=1e-8, atol_scale=1e-12, per_level=True)
    # vector quantity, both layouts
    levels_c = [g["C_rows%d" % l] for l in range(4)]
    _st, vec = scalar_setup(levels_c, n_comp=6)
    got = qe.estimate_mean(qe.covariance(vec, Legendre(3, tuple(g["C_domain"]))), variance=False)
    rel_close(np.ravel(got.mean), np.ravel(g["C_cov_mean"]), rtol=1e-8, atol_scale=1e-13)
    top = qe.estimate_mean(qe.covariance(vec, Legendre(3, tuple(g["C_domain"])), cov_at_bottom=False), variance=False)
    rel_close(np.ravel(top.mean).reshape(3, 3, -1), np.moveaxis(np.ravel(got.mean).reshape(-1, 3, 3), 0, -1), rtol=1e-12)


def test_construct_density_single_pass_equals_two_pass_chain():
    """construct_density now reads the samples once (covariance means by linearisation, moments of the orthogonal basis
    as L cov[:, 0]); the oracle runs the reference's two-pass chain on the same samples."""
    from mlmc_b200.moments import Legendre
    from mlmc_b200.estimator import Estimate
    from mlmc_b200 import _native
    rng = np.random.default_rng(7)
    steps = orc.level_steps(3, (0.5, 0.005))
    levels = [orc.synth_level_rows(rng.normal(size=n), steps[l], steps[l - 1] if l else None)
              for l, n in enumerate([100000, 10000, 1000])]
    storage, value = scalar_setup(levels, [[h] for h in steps])
    domain = tuple(stats.norm.ppf([0.001, 0.999]))
    est = Estimate(value, storage, Legendre(12, domain))
    before = _native.launch_count
    distr_obj, info, result, moments_obj = est.construct_density(tol=1e-8, orth_moments_tol=1e-4)
    ob = orc.Basis("legendre", 12, domain)
    oc = orc.estimate_covariance(levels, ob)
    l_mat, _evals, _thr = orc.orthogonalize_moments(oc.mean.reshape(12, 12), 1e-4)
    om = orc.estimate_moments(levels, orc.Basis("legendre", 12, domain, matrix=l_mat))
    rel_close(info[2], l_mat, rtol=1e-6, atol_scale=1e-9)
    rel_close(distr_obj.moment_means, om.mean, rtol=1e-7, atol_scale=1e-10)
    assert _native.launch_count - before < 200
