"""The reference-facing Python API (Moments / Quantity / Estimate / SimpleDistribution) on the CUDA path, checked
against the reference's golden outputs (tests/golden/*.npz) and its own tests' closed forms.

These read like the reference's tests on purpose: test/test_moments.py, test/test_quantity_concept.py::test_moments,
test/test_sampling_pools.py, test/test_distribution.py.
"""
import numpy as np
import pytest
import scipy.stats as stats

from oracle import mlmc_oracle as orc

pytestmark = pytest.mark.gpu


def rel_close(got, want, rtol, atol_scale=1e-15):
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    assert got.shape == want.shape, (got.shape, want.shape)
    scale = np.max(np.abs(want[np.isfinite(want)])) if np.isfinite(want).any() else 1.0
    ok = np.isclose(got, want, rtol=rtol, atol=atol_scale * max(scale, 1e-300), equal_nan=True)
    assert ok.all(), "max abs err %.3e (scale %.3e)" % (np.nanmax(np.abs(got - want)[~ok]), scale)


def scalar_setup(levels, level_parameters=None, n_ops=None, n_comp=1):
    from mlmc_b200.sample_storage import Memory
    from mlmc_b200.quantity.quantity import make_root_quantity
    from mlmc_b200.quantity.quantity_spec import QuantitySpec
    spec = [QuantitySpec(name="v", unit="", shape=(n_comp, 1), times=[0.0], locations=["0"])]
    storage = Memory.from_arrays(levels, level_parameters=level_parameters, n_ops=n_ops, result_format=spec)
    root = make_root_quantity(storage, spec)
    q = root["v"][0.0]["0"]
    return storage, (q[0, 0] if n_comp == 1 else q)


# ---------------------------------------------------------------- test/test_moments.py
def test_monomials():
    from mlmc_b200.moments import Monomial
    size = 5
    values = np.array([-2, -1, -0.5, 0, 0.5, 1, 2])
    ref = np.array([values ** r for r in range(size)]).T
    assert np.allclose(ref, Monomial(size, safe_eval=False)(values))
    a, b = (-1, 3)
    assert np.allclose(ref, Monomial(size, (a, b), safe_eval=False)((b - a) * values + a))
    sample = np.random.default_rng(0).normal(size=1000)
    assert np.abs(np.mean(Monomial(2, safe_eval=False)(sample)[:, 1])) < 0.1


def test_fourier():
    from mlmc_b200.moments import Fourier
    size = 6
    values = np.array([0.0, 0.25, 0.5, 0.75, 1.0])
    v = 2 * np.pi * values
    ref = np.array([np.ones_like(v), np.cos(v), np.sin(v), np.cos(2 * v), np.sin(2 * v), np.cos(3 * v)]).T
    assert np.allclose(ref, Fourier(size, (0, 1))(values))
    a, b = (-1, 3)
    assert np.allclose(ref, Fourier(size, (a, b))((b - a) * values + a))


def test_legendre():
    from mlmc_b200.moments import Legendre
    values = np.array([0.0, 0.25, 0.5, 0.75, 1.0])
    ref = np.array([np.ones_like(values), values, (3 * values ** 2 - 1.0) / 2.0, (5 * values ** 3 - 3 * values) / 2.0]).T
    assert np.allclose(ref, Legendre(4, (-1.0, 1.0))(values))


def test_transform_and_derivatives():
    from mlmc_b200.moments import Legendre, TransformedMoments
    size = 5
    values = np.array([0.0, 0.25, 0.5, 0.75, 1.0])
    fn = Legendre(size, [-1.0, 1.0], log=False, safe_eval=True)
    assert np.allclose(fn(values), TransformedMoments(fn, np.eye(size))(values))
    mat = np.ones((size, size))
    assert np.allclose(np.matmul(fn(values), mat.T), TransformedMoments(fn, mat)(values))
    # derivatives against numpy's Legendre calculus (moments.py:199-229)
    for deg, got in ((1, fn.eval_diff(values)), (2, fn.eval_diff2(values)), (1, fn.eval_all_der(values, degree=1))):
        want = np.array([np.polynomial.legendre.legval(values, np.polynomial.legendre.legder(np.eye(size)[s], deg))
                         for s in range(size)]).T
        assert np.allclose(got, want)
    assert fn == Legendre(size, [-1.0, 1.0]) and fn != Legendre(size + 1, [-1.0, 1.0])
    assert fn.change_size(3).size == 3
    assert np.allclose(fn.inv_transform(fn.transform(values)), values)
    assert np.isnan(fn.transform(np.array([2.0]))[0])
    # eval / eval_single_moment / shapes
    assert np.allclose(fn.eval(2, values), (3 * values ** 2 - 1) / 2)
    assert fn.eval_all(values.reshape(5, 1), 3).shape == (5, 1, 3)


# ---------------------------------------------------------------- test/test_sampling_pools.py
def test_sampling_pools_golden(golden):
    from mlmc_b200.sample_storage import Memory
    from mlmc_b200.quantity.quantity import make_root_quantity
    from mlmc_b200.quantity.quantity_spec import QuantitySpec
    from mlmc_b200.moments import Legendre
    from mlmc_b200.estimator import Estimate
    g = golden("sampling_pools")
    fmt = [QuantitySpec(name="length", unit="m", shape=(2, 1), times=[1, 2, 3], locations=["10", "20"]),
           QuantitySpec(name="width", unit="mm", shape=(2, 1), times=[1, 2, 3], locations=["30", "40"])]
    storage = Memory.from_arrays([g["rows%d" % l] for l in range(3)], level_parameters=[[0.01], [0.001], [0.0001]],
                                 result_format=fmt)
    quantity = make_root_quantity(storage=storage, q_specs=storage.load_result_format())
    value_quantity = quantity["length"][1]["10"][0]
    moments_fn = Legendre(5, stats.norm(loc=1, scale=2).ppf([0.0001, 0.9999]))
    means, variances = Estimate(quantity=value_quantity, sample_storage=storage, moments_fn=moments_fn).estimate_moments()
    assert means[0] == 1
    assert variances[0] == 0
    assert np.allclose(g["ref_means"], means, atol=1e-5)
    rel_close(means, g["means"], rtol=1e-10)
    rel_close(variances, g["vars"], rtol=1e-10)


# ---------------------------------------------------------------- Estimate on the multi-level golden case
def test_estimate_three_levels(golden):
    from mlmc_b200.moments import Legendre, Monomial, Fourier
    from mlmc_b200.estimator import Estimate, estimate_n_samples_for_target_variance
    g = golden("estimates")
    levels = [g["A_rows%d" % l] for l in range(3)]
    steps = g["A_steps"]
    storage, value = scalar_setup(levels, [[h] for h in steps], g["A_n_ops"])
    domain = tuple(g["A_domain"])
    for tag, fn in (("leg", Legendre(12, domain)), ("mono", Monomial(6, domain)), ("four", Fourier(7, domain)),
                    ("legraw", Legendre(8, (-20.0, 20.0), safe_eval=False))):
        est = Estimate(value, storage, fn)
        means, variances = est.estimate_moments()
        rel_close(means, g["A_%s_mean" % tag], rtol=1e-10)
        rel_close(variances, g["A_%s_var" % tag], rtol=1e-10)
        l_vars, n_samples = est.estimate_diff_vars()
        rel_close(l_vars, g["A_%s_l_vars" % tag], rtol=1e-10)
        assert np.array_equal(n_samples, g["A_%s_n" % tag])
    est = Estimate(value, storage, Legendre(12, domain))
    reg_vars, n_ops = est.estimate_diff_vars_regression(g["A_leg_n"])
    rel_close(reg_vars, g["A_leg_reg_vars"], rtol=1e-9)
    n_est = estimate_n_samples_for_target_variance(1e-5, reg_vars, n_ops, n_levels=3)
    assert np.array_equal(n_est, g["A_leg_n_estimated"])
    dom = Estimate.estimate_domain(value, storage, quantile=0.01)
    assert np.array_equal(np.array(dom), g["A_est_domain"])
    cov_mean, cov_var = Estimate(value, storage, Legendre(8, domain)).estimate_covariance()
    rel_close(cov_mean, g["A_cov_mean"], rtol=1e-8)
    rel_close(cov_var, g["A_cov_var"], rtol=1e-8)
    assert np.array_equal(cov_mean, cov_mean.T)


def test_quantity_mean_result_object(golden):
    from mlmc_b200.moments import Legendre
    from mlmc_b200.quantity import quantity_estimate as qe
    g = golden("estimates")
    levels = [g["A_rows%d" % l] for l in range(3)]
    storage, value = scalar_setup(levels, [[h] for h in g["A_steps"]])
    qm = qe.estimate_mean(qe.moments(value, Legendre(12, tuple(g["A_domain"]))))
    assert qm.l_means.shape == (3, 12) and qm.l_vars.shape == (3, 12)
    assert np.array_equal(qm.n_rm_samples, g["A_leg_n_rm"])
    rel_close(np.squeeze(qm[3].mean), g["A_leg_mean"][3], rtol=1e-10)
    # samples() keeps the reference contract (NumPy [M*R, n, 2])
    chunk = qe.moments(value, Legendre(4, tuple(g["A_domain"]))).samples(next(storage.chunks(level_id=1)))
    want = orc.moments_chunk(orc.Basis("legendre", 4, tuple(g["A_domain"])), levels[1].transpose(2, 0, 1))
    assert chunk.shape == want.shape and np.array_equal(chunk, want, equal_nan=True)
    masked, n_bad = qe.mask_nan_samples(chunk)
    assert n_bad == g["A_leg_n_rm"][1] and masked.shape[1] == chunk.shape[1] - n_bad


def test_transformed_moments_and_log_domain(golden):
    from mlmc_b200.moments import Legendre, TransformedMoments
    from mlmc_b200.quantity import quantity_estimate as qe
    from mlmc_b200.estimator import Estimate
    from mlmc_b200.tool.simple_distribution import construct_ortogonal_moments
    g = golden("estimates")
    levels = [g["A_rows%d" % l] for l in range(3)]
    storage, value = scalar_setup(levels, [[h] for h in g["A_steps"]])
    base = Legendre(10, tuple(g["A_domain"]))
    cov = qe.estimate_mean(qe.covariance(value, base)).mean
    rel_close(cov, g["A_cov10_mean"], rtol=1e-8)
    orth, info = construct_ortogonal_moments(base, g["A_cov10_mean"], tol=1e-4)
    rel_close(info[2], g["A_orth_L"], rtol=1e-9, atol_scale=1e-12)
    assert info[1] == int(g["A_orth_threshold"])
    tm = qe.estimate_mean(qe.moments(value, TransformedMoments(base, g["A_orth_L"])))
    rel_close(tm.mean, g["A_orth_mean"], rtol=1e-9, atol_scale=1e-13)
    rel_close(tm.l_means, g["A_orth_l_means"], rtol=1e-9, atol_scale=1e-13)
    rel_close(tm.l_vars, g["A_orth_l_vars"], rtol=1e-8, atol_scale=1e-12)
    # ||L cov L^T - I|| < 1e-10 in the centred sense asserted by test/test_distribution.py:180
    # log domain, one level (config 1 in miniature)
    storage_b, value_b = scalar_setup([g["B_rows0"]], [[0.1]])
    dom_b = Estimate.estimate_domain(value_b, storage_b, quantile=0.001)
    assert np.array_equal(np.array(dom_b), g["B_domain"])
    means, variances = Estimate(value_b, storage_b, Legendre(25, dom_b, log=True, safe_eval=True)).estimate_moments()
    rel_close(means, g["B_mean"], rtol=1e-10)
    rel_close(variances, g["B_var"], rtol=1e-10)


def test_vector_quantity_and_dag(golden):
    from mlmc_b200.moments import Legendre
    from mlmc_b200.quantity import quantity_estimate as qe
    g = golden("estimates")
    levels = [g["C_rows%d" % l] for l in range(4)]
    storage, vec = scalar_setup(levels, n_comp=6)
    fn = Legendre(5, tuple(g["C_domain"]))
    for tag, bottom in (("bottom", True), ("top", False)):
        qm = qe.estimate_mean(qe.moments(vec, fn, mom_at_bottom=bottom))
        assert qm.mean.shape == g["C_%s_mean" % tag].shape
        rel_close(qm.mean, g["C_%s_mean" % tag], rtol=1e-10)
        rel_close(qm.var, g["C_%s_var" % tag], rtol=1e-10)
        rel_close(qm.l_vars, g["C_%s_l_vars" % tag], rtol=1e-10)
        assert np.array_equal(qm.n_samples, g["C_%s_n" % tag]) and np.array_equal(qm.n_rm_samples, g["C_%s_n_rm" % tag])
    raw = qe.estimate_mean(vec)
    rel_close(raw.mean, g["C_raw_mean"], rtol=1e-10)
    rel_close(raw.var, g["C_raw_var"], rtol=1e-10)
    cov = qe.estimate_mean(qe.covariance(vec, Legendre(3, tuple(g["C_domain"]))))     # generic (vector) route
    rel_close(cov.mean, g["C_cov_mean"], rtol=1e-8)
    rel_close(cov.var, g["C_cov_var"], rtol=1e-8)
    # DAG identities of test/test_quantity_concept.py:574-582: mean(2 q) = 2 mean(q), mean(q + q) likewise
    m1 = qe.estimate_mean(vec).mean
    rel_close(qe.estimate_mean(2 * vec).mean, 2 * m1, rtol=1e-12)
    rel_close(qe.estimate_mean(vec + vec).mean, 2 * m1, rtol=1e-12)
    rel_close(qe.estimate_mean(vec / 2 - vec).mean, -0.5 * m1, rtol=1e-12)
    # indexing, ufuncs, selection
    comp = vec[2, 0]
    rel_close(np.squeeze(qe.estimate_mean(comp).mean), m1.reshape(-1)[2], rtol=1e-12)
    want = orc.estimate_mean(levels, lambda x: np.sin(x)).mean
    rel_close(qe.estimate_mean(np.sin(vec)).mean.reshape(-1), want, rtol=1e-11)
    sel = vec.select(vec < 1.0)
    want = orc.estimate_mean([lv[(lv[:, :1 if i == 0 else 2, :] < 1.0).all(axis=(1, 2))] for i, lv in enumerate(levels)]).mean
    rel_close(qe.estimate_mean(sel).mean.reshape(-1), want, rtol=1e-11)
    one = qe.moment(comp, fn, 2)
    rel_close(np.squeeze(qe.estimate_mean(one).mean), qe.estimate_mean(qe.moments(comp, fn)).mean[2], rtol=1e-11)


def test_all_samples_masked_raises():
    from mlmc_b200.moments import Legendre
    from mlmc_b200.quantity import quantity_estimate as qe
    storage, value = scalar_setup([np.full((20, 2, 1), 9.0)])
    with pytest.raises(Exception, match="All samples were masked"):
        qe.estimate_mean(qe.moments(value, Legendre(3, (0.0, 1.0))))


def test_npy_storage_streaming(tmp_path, golden):
    """The file-backed read side (HDF stand-in) streamed in small chunks gives the same estimate."""
    from mlmc_b200.sample_storage import NpyStorage
    from mlmc_b200.quantity.quantity import make_root_quantity
    from mlmc_b200.quantity.quantity_spec import QuantitySpec
    from mlmc_b200.moments import Legendre
    from mlmc_b200.estimator import Estimate
    g = golden("estimates")
    levels = [g["A_rows%d" % l] for l in range(3)]
    storage = NpyStorage.write(str(tmp_path / "st"), levels, [[h] for h in g["A_steps"]], g["A_n_ops"])
    storage.device_chunk_bytes = 16 * 300          # force ~14 chunks on level 0
    storage.resident_fraction = 0.0                # never keep resident: exercise the double-buffered stream
    spec = [QuantitySpec(name="v", unit="", shape=(1, 1), times=[0.0], locations=["0"])]
    value = make_root_quantity(storage, spec)["v"][0.0]["0"][0, 0]
    means, variances = Estimate(value, storage, Legendre(12, tuple(g["A_domain"]))).estimate_moments()
    rel_close(means, g["A_leg_mean"], rtol=1e-10)
    rel_close(variances, g["A_leg_var"], rtol=1e-10)


def test_hdf_storage_streaming(tmp_path, golden):
    """The reference's HDF5 layout (Levels/<l>/collected_values, chunked, (2, M) float64 elements) read by the built-in
    reader and streamed through the two pinned staging buffers in many small pieces; also the resident path and a
    vector quantity; same estimates as the golden reference outputs."""
    from mlmc_b200.sample_storage import SampleStorageHDF
    from mlmc_b200.tool.hdf5_min import write_mlmc_file
    from mlmc_b200.quantity.quantity import make_root_quantity
    from mlmc_b200.quantity.quantity_spec import QuantitySpec
    from mlmc_b200.quantity import quantity_estimate as qe
    from mlmc_b200.moments import Legendre
    from mlmc_b200.estimator import Estimate
    g = golden("estimates")
    levels = [g["A_rows%d" % l] for l in range(3)]
    path = write_mlmc_file(str(tmp_path / "mlmc.hdf5"), levels, [[h] for h in g["A_steps"]], g["A_n_ops"], chunk_rows=500)
    spec = [QuantitySpec(name="v", unit="", shape=(1, 1), times=[0.0], locations=["0"])]
    for resident in (0.0, 0.6):
        storage = SampleStorageHDF(path, backend="min")
        storage.device_chunk_bytes = 16 * 300      # ~14 pieces on level 0, each crossing HDF5 chunk boundaries
        storage.resident_fraction = resident
        value = make_root_quantity(storage, spec)["v"][0.0]["0"][0, 0]
        est = Estimate(value, storage, Legendre(12, tuple(g["A_domain"])))
        for _ in range(2):                         # second call: resident copy (or a fresh stream)
            means, variances = est.estimate_moments()
            rel_close(means, g["A_leg_mean"], rtol=1e-10)
            rel_close(variances, g["A_leg_var"], rtol=1e-10)
        reg_vars, n_ops = est.estimate_diff_vars_regression(g["A_leg_n"])
        rel_close(reg_vars, g["A_leg_reg_vars"], rtol=1e-9)
        assert np.allclose(n_ops, g["A_n_ops"])
        dom = Estimate.estimate_domain(value, storage, quantile=0.01)
        assert np.array_equal(np.array(dom), g["A_est_domain"])
    # vector quantity from a second file
    levels_c = [g["C_rows%d" % l] for l in range(4)]
    path_c = write_mlmc_file(str(tmp_path / "vec.hdf5"), levels_c, [[0.1]] * 4, chunk_rows=37)
    storage = SampleStorageHDF(path_c, backend="min")
    storage.resident_fraction = 0.0
    storage.device_chunk_bytes = 6 * 16 * 50
    spec_c = [QuantitySpec(name="v", unit="", shape=(6, 1), times=[0.0], locations=["0"])]
    vec = make_root_quantity(storage, spec_c)["v"][0.0]["0"]
    qm = qe.estimate_mean(qe.moments(vec, Legendre(5, tuple(g["C_domain"]))))
    rel_close(qm.mean, g["C_bottom_mean"], rtol=1e-10)
    rel_close(qm.var, g["C_bottom_var"], rtol=1e-10)
    assert np.array_equal(qm.n_samples, g["C_bottom_n"])


# ---------------------------------------------------------------- test/test_distribution.py
@pytest.mark.parametrize("size", [5, 15, 25])
def test_simple_distribution_matches_reference(golden, size):
    from mlmc_b200.moments import Legendre
    from mlmc_b200.tool.simple_distribution import (SimpleDistribution, construct_ortogonal_moments,
                                                    compute_semiexact_cov, compute_semiexact_moments)
    g = golden("maxent")
    t = "R%d_" % size
    domain = tuple(g["domain"])
    distr = stats.norm(loc=1, scale=2)
    base = Legendre(size, domain, safe_eval=False)
    cov = compute_semiexact_cov(base, distr.pdf, n_panels=int(g["n_panels"]))
    rel_close(cov, g[t + "cov"], rtol=1e-10, atol_scale=1e-14)
    orth, info = construct_ortogonal_moments(base, g[t + "cov"], tol=1e-4)
    rel_close(info[2], g[t + "L"], rtol=1e-8, atol_scale=1e-11)
    # || L cov L^T - I || < 1e-10 after centring is what test/test_distribution.py:180 asserts
    mu = compute_semiexact_moments(orth, distr.pdf, n_panels=int(g["n_panels"]))
    rel_close(mu, g[t + "mu"], rtol=1e-9, atol_scale=1e-12)
    data = np.stack([g[t + "mu"], np.ones(len(mu))], axis=1)
    sd = SimpleDistribution(orth, data, domain=domain, quad_panels=int(g["n_panels"]))
    res = sd.estimate_density_minimize(tol=1e-8, reg_param=0.0)
    assert res.success and res.nit == int(g[t + "nit"])
    # north_star tolerance: 1e-6 on multipliers and PDF values
    rel_close(sd.multipliers, g[t + "multipliers"], rtol=1e-6, atol_scale=1e-8)
    rel_close(sd.density(g[t + "density_x"]), g[t + "density"], rtol=1e-6)
    rel_close(sd.cdf(g[t + "density_x"]), g[t + "cdf"], rtol=1e-6, atol_scale=1e-9)
    rel_close(res.eigvals, g[t + "eigvals"], rtol=1e-6, atol_scale=1e-9)
    assert res.fun_norm < 1e-7 and hasattr(res, "solver_res")
    if (t + "adaptive_multipliers") in g.files:       # the unmodified adaptive reference, where it converges
        rel_close(sd.multipliers, g[t + "adaptive_multipliers"], rtol=1e-6, atol_scale=1e-8)


def test_construct_density_end_to_end():
    """Estimate.construct_density on synthetic normal samples: the fitted PDF integrates to 1, reproduces the
    estimated moments and is close to the true density."""
    from mlmc_b200.moments import Legendre
    from mlmc_b200.estimator import Estimate
    rng = np.random.default_rng(42)
    steps = orc.level_steps(3, (0.5, 0.005))
    levels = [orc.synth_level_rows(rng.normal(size=n), steps[l], steps[l - 1] if l else None)
              for l, n in enumerate([200000, 20000, 2000])]
    storage, value = scalar_setup(levels, [[h] for h in steps])
    domain = tuple(stats.norm.ppf([0.001, 0.999]))
    est = Estimate(value, storage, Legendre(12, domain))
    distr_obj, info, result, moments_obj = est.construct_density(tol=1e-8, orth_moments_tol=1e-4)
    # sampled moments: trust-ncg may stop at its 20-iteration cap a hair above tol (the reference does not assert
    # convergence either, test/test_distribution.py:215-228 is commented out); the residual must be tiny though
    assert result.success or result.fun_norm < 1e-6
    xs = np.linspace(domain[0], domain[1], 201)
    pdf = distr_obj.density(xs)
    assert abs(np.trapezoid(pdf, xs) - 1.0) < 5e-3
    assert np.max(np.abs(pdf - stats.norm.pdf(xs))) < 5e-2
    assert abs(distr_obj.cdf(np.array([domain[1] - 1e-9]))[0] - 1.0) < 5e-3
    # oracle on the same data
    oc = orc.estimate_covariance(levels, orc.Basis("legendre", 12, domain))
    l_mat, _, _ = orc.orthogonalize_moments(oc.mean.reshape(12, 12), 1e-4)
    rel_close(info[2], l_mat, rtol=1e-6, atol_scale=1e-9)


def test_bootstrap_and_subsample(golden):
    """est_bootstrap (estimator.py:171-205): statistics of sub-sampled estimates have the reference's shapes and the
    bootstrap mean of the moment means is close to the full-sample estimate (statistical check, like the reference's
    atol=1e-2 tests in test_quantity_concept.py)."""
    from mlmc_b200.moments import Legendre
    from mlmc_b200.estimator import Estimate
    from mlmc_b200.quantity import quantity_estimate as qe
    g = golden("estimates")
    levels = [g["A_rows%d" % l] for l in range(3)]
    storage, value = scalar_setup(levels, [[h] for h in g["A_steps"]], g["A_n_ops"])
    fn = Legendre(6, tuple(g["A_domain"]))
    est = Estimate(value, storage, fn)
    est.est_bootstrap(n_subsamples=12, sample_vector=[2000, 1000, 500])
    assert est.mean_bs_mean.shape == (6,) and est.mean_bs_l_vars.shape == (3, 6) and est.var_bs_l_means.shape == (3, 6)
    full, _ = est.estimate_moments()
    assert np.allclose(est.mean_bs_mean, full, atol=2e-2)
    assert est.mean_bs_mean[0] == 1.0
    sub = value.subsample(sample_vec=[100, 50, 25])
    qm = qe.estimate_mean(qe.moments(sub, fn))
    assert list(qm.n_samples + qm.n_rm_samples) == [100, 50, 25]


@pytest.mark.parametrize("method", ["gather", "weighted"])
def test_fused_bootstrap_matches_oracle_on_same_rows(golden, method):
    """qe.bootstrap_moments: every replicate equals the oracle's estimate on exactly the rows it drew -- by the per-replicate
    gather kernel and by the one-pass weighted sums on tensor tiles (multiplicities of the same draws)."""
    import torch
    from mlmc_b200.moments import Legendre
    from mlmc_b200.quantity import quantity_estimate as qe
    g = golden("estimates")
    levels = [g["A_rows%d" % l] for l in range(3)]
    storage, value = scalar_setup(levels, [[h] for h in g["A_steps"]], g["A_n_ops"])
    domain = tuple(g["A_domain"])
    fn = Legendre(8, domain)
    sample_vec = [1500, 700, 300]
    out = qe.bootstrap_moments(value, fn, sample_vec, 6, seed=11, return_indices=True, method=method)
    assert out["mean"].shape == (6, 8) and out["l_vars"].shape == (6, 3, 8)
    assert np.all(out["mean"][:, 0] == 1.0) and np.all(out["var"][:, 0] == 0.0)
    basis = orc.Basis("legendre", 8, domain)
    for b in range(6):
        picked = []
        for l in range(3):
            parts = [levels[l][off + idx[b - b0].cpu().numpy()] for off, b0, idx in out["indices"][l]
                     if b0 <= b < b0 + idx.shape[0]]
            picked.append(np.concatenate(parts))
        want = orc.estimate_moments(picked, basis)
        assert list(out["n_samples"][b]) == list(want.n_samples)
        rel_close(out["l_means"][b], want.l_means, rtol=1e-10, atol_scale=1e-14)
        rel_close(out["l_vars"][b], want.l_vars, rtol=1e-10, atol_scale=1e-13)
        rel_close(out["mean"][b], want.mean, rtol=1e-10, atol_scale=1e-14)
        rel_close(out["var"][b], want.var, rtol=1e-10, atol_scale=1e-13)
    # same seed -> same replicates
    again = qe.bootstrap_moments(value, fn, sample_vec, 6, seed=11, method=method)
    assert np.array_equal(again["mean"], out["mean"]) and np.array_equal(again["l_vars"], out["l_vars"])


def test_fused_bootstrap_statistics(golden):
    """est_bootstrap on the fused path: the bootstrap variance of the level means estimates l_var / n
    (estimator.py:207: ``var_bs_l_means * n``) -- statistical check with many replicates."""
    from mlmc_b200.moments import Legendre
    from mlmc_b200.estimator import Estimate
    g = golden("estimates")
    levels = [g["A_rows%d" % l] for l in range(3)]
    storage, value = scalar_setup(levels, [[h] for h in g["A_steps"]], g["A_n_ops"])
    fn = Legendre(5, tuple(g["A_domain"]))
    est = Estimate(value, storage, fn)
    est.est_bootstrap(n_subsamples=400, seed=3)
    full_means, full_vars = est.estimate_moments()
    l_vars, n = est.estimate_diff_vars()
    assert np.allclose(est.mean_bs_mean, full_means, atol=5 * np.sqrt(np.max(full_vars)) + 1e-12)
    ratio = est._bs_level_mean_variance[:, 1:] / l_vars[:, 1:]
    assert np.all((ratio > 0.7) & (ratio < 1.4)), ratio
    n_est = est.bs_target_var_n_estimated(1e-5)
    assert n_est.shape == (3,) and np.all(n_est >= 0)


def test_bases_beyond_the_fused_kernel_limits(golden):
    """More moments than the fused kernels hold (226 / 104): the generic device path (moments evaluated per row slice,
    RAW accumulate) still gives the reference's numbers."""
    from mlmc_b200.moments import Legendre
    from mlmc_b200.quantity import quantity_estimate as qe
    g = golden("estimates")
    levels = [g["A_rows%d" % l][:400] for l in range(3)]
    storage, value = scalar_setup(levels, [[h] for h in g["A_steps"]], g["A_n_ops"])
    domain = tuple(g["A_domain"])
    qm = qe.estimate_mean(qe.moments(value, Legendre(240, domain)))
    want = orc.estimate_moments(levels, orc.Basis("legendre", 240, domain))
    assert list(qm.n_samples) == list(want.n_samples) and list(qm.n_rm_samples) == list(want.n_rm_samples)
    rel_close(qm.l_means, want.l_means, rtol=1e-10, atol_scale=1e-14)
    rel_close(qm.l_vars, want.l_vars, rtol=1e-10, atol_scale=1e-13)
    old = qe._RAW_SLICE_BYTES
    qe._RAW_SLICE_BYTES = 1 << 22                        # several row slices per level
    try:
        cm = qe.estimate_mean(qe.covariance(value, Legendre(120, domain)))
    finally:
        qe._RAW_SLICE_BYTES = old
    wc = orc.estimate_covariance(levels, orc.Basis("legendre", 120, domain))
    rel_close(np.ravel(cm.mean), wc.mean, rtol=1e-8, atol_scale=1e-13)
    rel_close(np.ravel(cm.var), wc.var, rtol=1e-8, atol_scale=1e-13)


@pytest.mark.parametrize("method", ["gather", "weighted"])
def test_fused_bootstrap_streamed_chunks(golden, method):
    """A level that arrives in several device chunks: hypergeometric split of the draws over the chunks
    (quantity.py:317), ragged draws per replicate -- every replicate still equals the oracle on its own rows."""
    from mlmc_b200.moments import Legendre
    from mlmc_b200.quantity import quantity_estimate as qe
    g = golden("estimates")
    levels = [g["A_rows%d" % l] for l in range(3)]
    storage, value = scalar_setup(levels, [[h] for h in g["A_steps"]], g["A_n_ops"])
    storage.resident_fraction = 0.0            # stream from the host ...
    storage.device_chunk_bytes = 16 * 700      # ... 700 rows at a time
    domain = tuple(g["A_domain"])
    fn = Legendre(6, domain)
    sample_vec = [900, 400, 150]
    out = qe.bootstrap_moments(value, fn, sample_vec, 5, seed=2, return_indices=True, method=method)
    basis = orc.Basis("legendre", 6, domain)
    n_chunks = [len({off for off, _, _ in out["indices"][l]}) for l in range(3)]
    assert max(n_chunks) > 1
    for b in range(5):
        picked = []
        for l in range(3):
            parts = [levels[l][off + idx[b - b0].cpu().numpy()] for off, b0, idx in out["indices"][l]
                     if b0 <= b < b0 + idx.shape[0]]
            picked.append(np.concatenate(parts))
            assert len(picked[-1]) == sample_vec[l]
        want = orc.estimate_moments(picked, basis)
        assert list(out["n_samples"][b]) == list(want.n_samples)
        rel_close(out["l_means"][b], want.l_means, rtol=1e-10, atol_scale=1e-14)
        rel_close(out["l_vars"][b], want.l_vars, rtol=1e-10, atol_scale=1e-13)


@pytest.mark.parametrize("r_base", [12, 25])
def test_construct_density_equals_oracle_pipeline(r_base):
    """The whole data-driven chain (covariance pass -> orthogonalisation -> moments of the orthogonal basis from the
    level sums -> max-ent fit) against the same chain in the oracle, on the same samples: PDF values within 1e-6
    (north_star), in practice 1e-14."""
    from mlmc_b200.moments import Legendre
    from mlmc_b200.estimator import Estimate
    rng = np.random.default_rng(42)
    steps = orc.level_steps(3, (0.5, 0.005))
    levels = [orc.synth_level_rows(rng.normal(size=n), steps[l], steps[l - 1] if l else None)
              for l, n in enumerate([200000, 20000, 2000])]
    storage, value = scalar_setup(levels, [[h] for h in steps])
    domain = tuple(stats.norm.ppf([0.001, 0.999]))
    est = Estimate(value, storage, Legendre(r_base, domain))
    distr_obj, info, result, moments_obj = est.construct_density(tol=1e-8, orth_moments_tol=1e-4)
    b = orc.Basis("legendre", r_base, domain)
    oc = orc.estimate_covariance(levels, b)
    l_mat, _, _ = orc.orthogonalize_moments(oc.mean.reshape(r_base, r_base), 1e-4)
    assert l_mat.shape[0] == moments_obj.size
    ob = orc.Basis("legendre", r_base, domain, matrix=l_mat)
    om = orc.estimate_moments(levels, ob)
    data = np.stack([om.mean, np.ones_like(om.mean)], axis=1)
    ofit = orc.maxent_fit(ob, data, domain, tol=1e-8, n_panels=distr_obj._n_panels)
    xs = np.linspace(domain[0], domain[1], 201)
    want = orc.maxent_density(ob, ofit.multipliers, np.ones(len(om.mean)), xs)
    assert np.max(np.abs(distr_obj.density(xs) - want)) < 1e-6 * max(1.0, np.max(want))


def test_module_level_domain_and_helpers(golden):
    """estimator.estimate_domain (module level: every level's own fine samples), calc_level_params,
    _variance_of_variance."""
    from mlmc_b200 import estimator
    from mlmc_b200.moments import Legendre
    g = golden("estimates")
    levels = [g["A_rows%d" % l] for l in range(3)]
    storage, value = scalar_setup(levels, [[h] for h in g["A_steps"]], g["A_n_ops"])
    lo, hi = estimator.estimate_domain(value, storage, quantile=0.01)
    n0 = len(levels[0])
    with np.errstate(invalid="ignore"):
        want = np.array([np.percentile(lv[:n0, 0, 0][~np.isnan(lv[:n0, 0, 0])], [1, 99]) for lv in levels])
    assert lo == want[:, 0].min() and hi == want[:, 1].max()
    assert estimator.calc_level_params((0.5, 0.005), 3) == [[0.5], [0.5 ** 0.5 * 0.005 ** 0.5], [0.005]]
    est = estimator.Estimate(value, storage, Legendre(4, (lo, hi)))
    vv = est._variance_of_variance([10, 100, 1000])
    assert vv.shape == (3,) and np.all(np.diff(vv) < 0) and abs(vv[2] - 2 / 999) < 2e-4


def test_allocation_grows_as_target_variance_shrinks(golden):
    """Property of the reference's (skipped) test/test_estimate.py: the n-sample allocation grows monotonically while
    the target variance decreases; the covariance estimate is symmetric; the allocation follows the oracle's."""
    from mlmc_b200 import estimator
    from mlmc_b200.moments import Legendre
    g = golden("estimates")
    levels = [g["A_rows%d" % l] for l in range(3)]
    storage, value = scalar_setup(levels, [[h] for h in g["A_steps"]], g["A_n_ops"])
    domain = tuple(g["A_domain"])
    est = estimator.Estimate(value, storage, Legendre(15, domain))
    variances, n_ops = est.estimate_diff_vars_regression(None)
    prev = np.zeros(3)
    for target in (1e-3, 1e-5, 1e-7, 1e-9):
        n_est = estimator.estimate_n_samples_for_target_variance(target, variances, n_ops, n_levels=3)
        assert np.all(n_est >= prev) and np.any(n_est > prev)
        prev = n_est
    want_vars = orc.regress_level_variances(orc.estimate_moments(levels, orc.Basis("legendre", 15, domain)).l_vars,
                                            g["A_steps"])
    rel_close(variances, want_vars, rtol=1e-9)
    assert np.array_equal(n_est, orc.n_samples_for_target_variance(1e-9, want_vars, n_ops, 3))
    cov, _ = est.estimate_covariance()
    assert np.array_equal(cov, cov.T)


def _vector_levels(rng, n_comp, sizes, spread=1.0):
    """Correlated fine / coarse rows [N, 2, M] of a vector quantity, a few values far outside the unit domain."""
    levels = []
    for l, n in enumerate(sizes):
        base = rng.normal(size=(n, 1, n_comp)) * spread
        rows = np.concatenate([base + 0.3 * rng.normal(size=(n, 1, n_comp)) * 2.0 ** -l,
                               base if l > 0 else np.zeros_like(base)], axis=1)
        rows[rng.integers(0, n, size=max(1, n // 97)), rng.integers(0, 2 if l > 0 else 1), rng.integers(0, n_comp)] = 50.0
        levels.append(rows)
    return levels


@pytest.mark.parametrize("n_comp,size,kind", [(5, 20, "legendre"), (3, 41, "legendre"), (200, 9, "monomial"),
                                              (7, 13, "fourier")])
def test_vector_covariance_is_fused(n_comp, size, kind):
    """``covariance`` of a vector quantity (quantity_estimate.py:131-147 handles any shape): all components go through
    the DMMA kernel in one launch, with the mask of ``mask_nan_samples`` shared by the components; both layouts."""
    from mlmc_b200 import moments as mom
    from mlmc_b200.quantity import quantity_estimate as qe
    rng = np.random.default_rng(100 + n_comp)
    levels = _vector_levels(rng, n_comp, [700, 333, 130])
    storage, vec = scalar_setup(levels, n_comp=n_comp)
    domain = (-3.0, 3.0)
    fn = {"legendre": mom.Legendre, "monomial": mom.Monomial, "fourier": mom.Fourier}[kind](size, domain)
    want = orc.estimate_covariance(levels, orc.Basis(kind, size, domain), chunk_rows=4096)
    assert int(np.sum(want.n_rm_samples)) > 0
    for bottom in (True, False):
        quantity = qe.covariance(vec, fn, cov_at_bottom=bottom)
        assert qe._Plan(quantity).kind == "covariance"
        qm = qe.estimate_mean(quantity)
        assert list(qm.n_samples) == list(want.n_samples) and list(qm.n_rm_samples) == list(want.n_rm_samples)

        def layout(a):          # oracle: [.., M * R * R] in cov_at_bottom order
            a = np.asarray(a)
            if bottom:
                return a
            lead = a.shape[:-1]
            return np.moveaxis(a.reshape(lead + (n_comp, size * size)), -1, -2).reshape(lead + (-1,))
        for l in range(3):
            rel_close(np.reshape(qm.l_means, (3, -1))[l], layout(want.l_means)[l], rtol=1e-8, atol_scale=1e-13)
            rel_close(np.reshape(qm.l_vars, (3, -1))[l], layout(want.l_vars)[l], rtol=1e-8, atol_scale=1e-13)
        rel_close(np.ravel(qm.mean), layout(want.mean), rtol=1e-8, atol_scale=1e-13)
        rel_close(np.ravel(qm.var), layout(want.var), rtol=1e-8, atol_scale=1e-13)


def test_vector_covariance_component_batches():
    """More components than one launch takes (2048): the component batches land in the right slots."""
    from mlmc_b200.moments import Legendre
    from mlmc_b200.quantity import quantity_estimate as qe
    rng = np.random.default_rng(7)
    n_comp, size = 2100, 3
    levels = _vector_levels(rng, n_comp, [40, 24], spread=0.5)
    storage, vec = scalar_setup(levels, n_comp=n_comp)
    fn = Legendre(size, (-60.0, 60.0))
    qm = qe.estimate_mean(qe.covariance(vec, fn))
    want = orc.estimate_covariance(levels, orc.Basis("legendre", size, (-60.0, 60.0)))
    assert list(qm.n_samples) == [40, 24]
    rel_close(np.ravel(qm.mean), want.mean, rtol=1e-8, atol_scale=1e-13)
    rel_close(np.ravel(qm.var), want.var, rtol=1e-8, atol_scale=1e-13)


def test_vector_transformed_moments_are_fused():
    """``moments(vector quantity, TransformedMoments)``: sums and difference Gram of the base functions per component,
    transformed under the sums (moments.py:256-259)."""
    from mlmc_b200.moments import Legendre, TransformedMoments
    from mlmc_b200.quantity import quantity_estimate as qe
    rng = np.random.default_rng(11)
    n_comp, r0, r1 = 4, 12, 7
    levels = _vector_levels(rng, n_comp, [900, 400, 150])
    storage, vec = scalar_setup(levels, n_comp=n_comp)
    domain = (-3.0, 3.0)
    l_mat = rng.normal(size=(r1, r0)) / np.sqrt(r0)
    l_mat[0] = 0.0
    l_mat[0, 0] = 1.0
    fn = TransformedMoments(Legendre(r0, domain), l_mat)
    want = orc.estimate_moments(levels, orc.Basis("legendre", r0, domain, matrix=l_mat), chunk_rows=4096)
    assert int(np.sum(want.n_rm_samples)) > 0
    quantity = qe.moments(vec, fn)
    assert qe._Plan(quantity).kind == "transformed"
    qm = qe.estimate_mean(quantity)
    assert list(qm.n_samples) == list(want.n_samples) and list(qm.n_rm_samples) == list(want.n_rm_samples)
    for l in range(3):
        rel_close(np.reshape(qm.l_means, (3, -1))[l], want.l_means[l], rtol=1e-9, atol_scale=1e-13)
        rel_close(np.reshape(qm.l_vars, (3, -1))[l], want.l_vars[l], rtol=1e-8, atol_scale=1e-12)
