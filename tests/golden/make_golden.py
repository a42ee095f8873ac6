"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run here (build container, where ``/root/reference`` is mounted)::

    python tests/golden/make_golden.py

It imports the reference through ``ref_shim``, runs the hot path on small seeded inputs, checks that
``oracle/mlmc_oracle.py`` reproduces every output, and writes ``*.npz`` fixtures holding the inputs and the
REFERENCE outputs.  ``tests/test_oracle_golden.py`` (CPU) replays them against the oracle and the ``-m gpu``
parity tests replay them against the CUDA path; neither needs the reference at run time.
"""
import hashlib
import os
import sys
import warnings

import numpy as np
import scipy.stats as stats

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import ref_shim  # noqa: E402

ref_shim.install()

import mlmc.moments as rm  # noqa: E402
import mlmc.quantity.quantity as rq  # noqa: E402
import mlmc.quantity.quantity_estimate as rqe  # noqa: E402
import mlmc.estimator as rest  # noqa: E402
import mlmc.tool.simple_distribution as rsd  # noqa: E402
from mlmc.quantity.quantity_spec import QuantitySpec  # noqa: E402
from mlmc.sim.synth_simulation import SynthSimulationWorkspace  # noqa: E402
from mlmc.sampling_pool import SamplingPool  # noqa: E402

from oracle import mlmc_oracle as orc  # noqa: E402

warnings.simplefilter("ignore")


def close(a, b, rtol=1e-12, atol=1e-14, what=""):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    ok = np.isclose(a, b, rtol=rtol, atol=atol, equal_nan=True)
    assert ok.all(), (what, np.abs(a - b)[~ok].max())


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote %-28s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


def scalar_quantity(storage, n_comp=1):
    spec = [QuantitySpec(name="v", unit="", shape=(n_comp, 1), times=[0.0], locations=["0"])]
    storage.save_result_format(spec)
    root = rq.make_root_quantity(storage, spec)
    q = root["v"][0.0]["0"]
    return q[0, 0] if n_comp == 1 else q


# --------------------------------------------------------------------------------------------------
def basis_tables():
    """Reference ``eval_all`` tables, incl. the closed-form cases of test/test_moments.py:11-70."""
    rng = np.random.default_rng(2024)
    out = {}
    cases = []
    x_wide = np.concatenate([rng.normal(0.5, 2.0, 200), [-3.0, 5.0, np.nan, 0.0, 1.0, 2.0]])
    x_pos = np.concatenate([rng.lognormal(0.0, 1.0, 200), [0.05, 20.0, np.nan, 1e-3, 40.0]])
    for size in (1, 2, 5, 25, 100):
        cases.append(("legendre", size, (-3.0, 5.0), False, True, x_wide))
        cases.append(("monomial", size, (-3.0, 5.0), False, True, x_wide))
    cases.append(("legendre", 30, (0.05, 20.0), True, True, x_pos))
    cases.append(("legendre", 9, (-1.0, 1.0), False, False, np.array([0.0, 0.25, 0.5, 0.75, 1.0, -1.5, 3.0])))
    cases.append(("monomial", 5, (0.0, 1.0), False, False, np.array([-2, -1, -0.5, 0, 0.5, 1, 2.0])))
    cases.append(("monomial", 5, (-1.0, 3.0), False, False, 4 * np.array([-2, -1, -0.5, 0, 0.5, 1, 2.0]) - 1))
    cases.append(("monomial", 12, (0.05, 20.0), True, True, x_pos))
    for size in (1, 2, 6, 7, 32):
        cases.append(("fourier", size, (0.0, 1.0), False, True, np.array([0.0, 0.25, 0.5, 0.75, 1.0, 1.5, -0.1])))
        cases.append(("fourier", size, (-3.0, 5.0), False, True, x_wide))
    cases.append(("fourier", 9, (-1.0, 3.0), False, False, x_wide))
    ctor = {"legendre": rm.Legendre, "monomial": rm.Monomial, "fourier": rm.Fourier}
    for i, (kind, size, dom, log, safe, x) in enumerate(cases):
        ref = ctor[kind](size, dom, log=log, safe_eval=safe).eval_all(x)
        mine = orc.basis_eval(orc.Basis(kind, size, dom, log=log, safe_eval=safe), x)
        assert np.array_equal(ref, mine, equal_nan=True), ("basis table not bit-identical", kind, size)
        out["case%02d_meta" % i] = np.array([kind, str(size), repr(dom), str(int(log)), str(int(safe))])
        out["case%02d_x" % i] = x
        out["case%02d_ref" % i] = ref
    # closed forms asserted by the reference's own tests (test_moments.py:61-70 and :39-57)
    v = np.array([0.0, 0.25, 0.5, 0.75, 1.0])
    leg = rm.Legendre(4, (-1.0, 1.0))(v)
    close(leg, np.array([np.ones_like(v), v, (3 * v ** 2 - 1) / 2, (5 * v ** 3 - 3 * v) / 2]).T, what="legendre closed form")
    fo = rm.Fourier(6, (0, 1))(v)
    a = 2 * np.pi * v
    close(fo, np.array([np.ones_like(a), np.cos(a), np.sin(a), np.cos(2 * a), np.sin(2 * a), np.cos(3 * a)]).T,
          atol=1e-12, what="fourier closed form")
    # TransformedMoments (moments.py:232-259)
    mat = rng.normal(size=(6, 9))
    mat[0] = 0
    mat[0, 0] = 1
    base = rm.Legendre(9, (-3.0, 5.0))
    ref = rm.TransformedMoments(base, mat).eval_all(x_wide)
    mine = orc.basis_eval(orc.Basis("legendre", 9, (-3.0, 5.0), matrix=mat), x_wide)
    assert np.array_equal(ref, mine, equal_nan=True)
    out["transformed_matrix"] = mat
    out["transformed_x"] = x_wide
    out["transformed_ref"] = ref
    out["n_cases"] = np.array(len(cases))
    save("basis_tables", **out)


# --------------------------------------------------------------------------------------------------
def sampling_pools_vector():
    """The reference's only golden vector: ``ref_means`` of test/test_sampling_pools.py:18, 55-87.

    Rows are produced exactly as ``OneProcessPool`` would: sample ids ``L%02d_S%07d`` (sampler.py:114-120),
    seed = first uint32 of md5(id) (sampling_pool.py:75-84), ``SynthSimulationWorkspace.calculate`` with
    ``distr='norm'`` -> norm(1, 2) (synth_simulation.py:218-286)."""
    ref_means = np.array([1., -0.03814235, -0.42411443, 0.05103307, 0.2123083])
    step_range = [[0.01], [0.001], [0.0001]]
    sim = SynthSimulationWorkspace({"config_yaml": "unused"})
    fmt = sim.result_format()
    SynthSimulationWorkspace._read_config = staticmethod(lambda: {"distr": "norm", "nan_fraction": 0})
    rows = []
    for level_id in range(3):
        coarse_step = step_range[level_id - 1][0] if level_id > 0 else 0
        cfg = {"fine": {"step": step_range[level_id][0]}, "coarse": {"step": coarse_step}, "res_format": fmt}
        lvl = []
        for s in range(10):
            seed = SamplingPool.compute_seed("L{:02d}_S{:07d}".format(level_id, s))
            fine, coarse = SynthSimulationWorkspace.calculate(cfg, seed)
            lvl.append(np.stack([fine, coarse]))
        rows.append(np.array(lvl))                                    # [10, 2, 24]
    storage = ref_shim.array_storage(rows, level_parameters=step_range)
    storage.save_result_format(fmt)
    root = rq.make_root_quantity(storage, fmt)
    value = root["length"][1]["10"][0]
    domain = stats.norm(loc=1, scale=2).ppf([0.0001, 0.9999])
    mf = rm.Legendre(5, domain)
    means, variances = rest.Estimate(value, storage, mf).estimate_moments(mf)
    assert means[0] == 1 and variances[0] == 0
    assert np.allclose(ref_means, means, atol=1e-5), means
    # which storage column is length[1]['10'][0]? quantity-major, then time, location, shape
    column = 0
    o = orc.estimate_moments([r[:, :, column:column + 1] for r in rows], orc.Basis("legendre", 5, tuple(domain)))
    close(o.mean, means, what="golden means")
    close(o.var, variances, what="golden vars")
    save("sampling_pools", ref_means=ref_means, rows0=rows[0], rows1=rows[1], rows2=rows[2], domain=domain,
         column=np.array(column), means=means, vars=variances)


# --------------------------------------------------------------------------------------------------
def synth_levels(rng, n_per_level, step_range, distr="norm", n_comp=1):
    steps = orc.level_steps(len(n_per_level), step_range)
    levels = []
    for l, n in enumerate(n_per_level):
        x = rng.normal(size=n) if distr == "norm" else rng.lognormal(0.0, 1.0, size=n)
        rows = orc.synth_level_rows(x, steps[l], steps[l - 1] if l > 0 else None)
        if n_comp > 1:
            rows = rows + np.arange(n_comp)[None, None, :] * 1e-2 * (1 if l > 0 else 1)
            if l == 0:
                rows[:, 1, :] = 0
        levels.append(rows)
    return levels, steps


def estimates():
    """estimate_moments / covariance / regression / allocation on small seeded multi-level inputs."""
    rng = np.random.default_rng(1234)
    out = {}
    # --- case A: 3 levels, normal, Legendre + Monomial + (wrapped) Fourier, some samples out of domain
    levels, steps = synth_levels(rng, [4000, 2000, 1000], (0.5, 0.005))
    levels[1][7, 0, 0] = np.nan
    levels[2][11, 1, 0] = np.nan
    domain = tuple(stats.norm.ppf([0.005, 0.995]))
    n_ops = [orc.synth_n_ops(h) for h in steps]
    storage = ref_shim.array_storage(levels, level_parameters=[[h] for h in steps], n_ops=n_ops, chunk_rows=512)
    value = scalar_quantity(storage)
    for i, lv in enumerate(levels):
        out["A_rows%d" % i] = lv
    out["A_domain"] = np.array(domain)
    out["A_steps"] = np.array(steps)
    out["A_n_ops"] = np.array(n_ops)

    # the reference's Fourier only takes 1-D input (moments.py:153-161): wrap with ravel/reshape
    class FourierND(rm.Fourier):
        def _eval_all(self, value, size):
            value = np.asarray(value)
            return super()._eval_all(value.ravel(), size).reshape(value.shape + (size,))

    for tag, ref_fn, basis in (
            ("leg", rm.Legendre(12, domain), orc.Basis("legendre", 12, domain)),
            ("mono", rm.Monomial(6, domain), orc.Basis("monomial", 6, domain)),
            ("four", FourierND(7, domain), orc.Basis("fourier", 7, domain)),
            ("legraw", rm.Legendre(8, (-20.0, 20.0), safe_eval=False), orc.Basis("legendre", 8, (-20.0, 20.0), safe_eval=False))):
        est = rest.Estimate(value, storage, ref_fn)
        qm = rqe.estimate_mean(rqe.moments(value, ref_fn))
        o = orc.estimate_moments(levels, basis, chunk_rows=512)
        for k in ("l_means", "l_vars", "mean", "var"):
            close(getattr(o, k), getattr(qm, k), rtol=1e-11, atol=1e-15, what=tag + k)
        assert np.array_equal(o.n_samples, qm.n_samples) and np.array_equal(o.n_rm_samples, qm.n_rm_samples)
        out["A_%s_l_means" % tag] = qm.l_means
        out["A_%s_l_vars" % tag] = qm.l_vars
        out["A_%s_mean" % tag] = qm.mean
        out["A_%s_var" % tag] = qm.var
        out["A_%s_n" % tag] = qm.n_samples
        out["A_%s_n_rm" % tag] = qm.n_rm_samples
        if tag == "leg":
            reg_vars, ops = est.estimate_diff_vars_regression(qm.n_samples, ref_fn)
            n_est = rest.estimate_n_samples_for_target_variance(1e-5, reg_vars, ops, n_levels=3)
            close(orc.regress_level_variances(o.l_vars, steps), reg_vars, rtol=1e-10, what="regression")
            assert np.array_equal(orc.n_samples_for_target_variance(1e-5, reg_vars, ops, 3), n_est)
            out["A_leg_reg_vars"] = reg_vars
            out["A_leg_n_estimated"] = n_est
            dom = rest.Estimate.estimate_domain(value, storage, quantile=0.01)
            close(orc.estimate_domain(levels, 0.01), dom, what="domain")
            out["A_est_domain"] = np.array(dom)
    # covariance (R=8): mean + entry variance
    ref_fn = rm.Legendre(8, domain)
    cm = rqe.estimate_mean(rqe.covariance(value, ref_fn))
    oc = orc.estimate_covariance(levels, orc.Basis("legendre", 8, domain), chunk_rows=512)
    close(oc.mean.reshape(8, 8), cm.mean, rtol=1e-10, atol=1e-15, what="cov mean")
    close(oc.var.reshape(8, 8), cm.var, rtol=1e-10, atol=1e-18, what="cov var")
    out["A_cov_mean"] = cm.mean
    out["A_cov_var"] = cm.var
    out["A_cov_l_means"] = cm.l_means
    out["A_cov_l_vars"] = cm.l_vars
    # moments of an orthogonalised (TransformedMoments) basis + construct_ortogonal_moments
    ref_fn10 = rm.Legendre(10, domain)
    cov10 = rqe.estimate_mean(rqe.covariance(value, ref_fn10)).mean
    orth, info = rsd.construct_ortogonal_moments(ref_fn10, cov10, tol=1e-4)
    l_mine, evals, thr = orc.orthogonalize_moments(cov10, 1e-4)
    close(l_mine, info[2], rtol=1e-9, atol=1e-12, what="orth L")
    tm = rqe.estimate_mean(rqe.moments(value, orth))
    ot = orc.estimate_moments(levels, orc.Basis("legendre", 10, domain, matrix=info[2]), chunk_rows=512)
    close(ot.mean, tm.mean, rtol=1e-10, atol=1e-14, what="transformed mean")
    close(ot.l_vars, tm.l_vars, rtol=1e-9, atol=1e-16, what="transformed l_vars")
    out["A_cov10_mean"] = cov10
    out["A_orth_L"] = info[2]
    out["A_orth_evals"] = info[0]
    out["A_orth_threshold"] = np.array(info[1])
    out["A_orth_mean"] = tm.mean
    out["A_orth_var"] = tm.var
    out["A_orth_l_means"] = tm.l_means
    out["A_orth_l_vars"] = tm.l_vars

    # --- case B: 1 level lognormal, log domain, Legendre 25 (config 1 in miniature)
    levels_b, _ = synth_levels(rng, [5000], (0.1, 0.1), distr="lognorm")
    storage_b = ref_shim.array_storage(levels_b, level_parameters=[[0.1]], chunk_rows=4096)
    value_b = scalar_quantity(storage_b)
    dom_b = rest.Estimate.estimate_domain(value_b, storage_b, quantile=0.001)
    ref_fn = rm.Legendre(25, dom_b, log=True, safe_eval=True)
    qm = rqe.estimate_mean(rqe.moments(value_b, ref_fn))
    o = orc.estimate_moments(levels_b, orc.Basis("legendre", 25, tuple(dom_b), log=True), chunk_rows=4096)
    close(o.mean, qm.mean, rtol=1e-11, atol=1e-15, what="B mean")
    close(o.var, qm.var, rtol=1e-11, atol=1e-18, what="B var")
    out["B_rows0"] = levels_b[0]
    out["B_domain"] = np.array(dom_b)
    out["B_mean"] = qm.mean
    out["B_var"] = qm.var
    out["B_l_vars"] = qm.l_vars
    out["B_n"] = qm.n_samples
    out["B_n_rm"] = qm.n_rm_samples

    # --- case C: vector quantity, 6 components, 4 levels, per-component Legendre(5) moments, both layouts
    levels_c, steps_c = synth_levels(rng, [600, 300, 150, 80], (0.3, 0.003), n_comp=6)
    levels_c[2][5, 0, 3] = 50.0          # one component out of domain -> whole sample dropped
    storage_c = ref_shim.array_storage(levels_c, level_parameters=[[h] for h in steps_c], chunk_rows=256)
    value_c = scalar_quantity(storage_c, n_comp=6)
    dom_c = (-4.0, 4.1)
    ref_fn = rm.Legendre(5, dom_c)
    for bottom in (True, False):
        qm = rqe.estimate_mean(rqe.moments(value_c, ref_fn, mom_at_bottom=bottom))
        o = orc.estimate_moments(levels_c, orc.Basis("legendre", 5, dom_c), chunk_rows=256, mom_at_bottom=bottom)
        close(o.l_means, qm.l_means.reshape(4, -1), rtol=1e-11, atol=1e-15, what="C l_means")
        close(o.l_vars, qm.l_vars.reshape(4, -1), rtol=1e-11, atol=1e-18, what="C l_vars")
        tag = "bottom" if bottom else "top"
        out["C_%s_mean" % tag] = qm.mean
        out["C_%s_var" % tag] = qm.var
        out["C_%s_l_means" % tag] = qm.l_means
        out["C_%s_l_vars" % tag] = qm.l_vars
        out["C_%s_n" % tag] = qm.n_samples
        out["C_%s_n_rm" % tag] = qm.n_rm_samples
    # plain estimate_mean of the vector quantity itself (identity operation)
    qm = rqe.estimate_mean(value_c)
    o = orc.estimate_mean(levels_c, None, chunk_rows=256)
    close(o.l_means, qm.l_means.reshape(4, -1), rtol=1e-12, what="C raw means")
    out["C_raw_mean"] = qm.mean
    out["C_raw_var"] = qm.var
    cm = rqe.estimate_mean(rqe.covariance(value_c, rm.Legendre(3, dom_c)))
    oc = orc.estimate_covariance(levels_c, orc.Basis("legendre", 3, dom_c), chunk_rows=256)
    close(oc.mean, cm.mean.reshape(-1), rtol=1e-10, atol=1e-15, what="C cov")
    out["C_cov_mean"] = cm.mean
    out["C_cov_var"] = cm.var
    for i, lv in enumerate(levels_c):
        out["C_rows%d" % i] = lv
    out["C_domain"] = np.array(dom_c)
    save("estimates", **out)


# --------------------------------------------------------------------------------------------------
def maxent():
    """SimpleDistribution on a fixed composite Gauss rule (reference formulas, ``_update_quadrature`` replaced
    by a fixed node set) and, where it converges, the unmodified adaptive reference."""
    out = {}
    distr = stats.norm(loc=1, scale=2)
    domain = tuple(distr.ppf([0.01, 0.99]))
    n_panels = 60
    nodes, w = orc.gauss_panels(domain, n_panels)

    class FixedNodes(rsd.SimpleDistribution):
        def _update_quadrature(self, multipliers, force=False):
            if not force:
                return
            self._quad_points, self._quad_weights = nodes, w
            self._quad_moments = self.eval_moments(nodes)

        def _calculate_exact_moment(self, multipliers, m=0, full_output=0):
            rho = self._density_in_quads(multipliers)
            return np.dot(rho * self._quad_moments[:, m], w), None

    for size in (5, 15, 25):
        base = rm.Legendre(size, domain, safe_eval=False)
        phi = base.eval_all(nodes)
        pdf_w = distr.pdf(nodes) * w
        cov = (phi.T * pdf_w) @ phi
        orth, info = rsd.construct_ortogonal_moments(base, cov, tol=1e-4)
        l_mat = info[2]
        l_mine, _, thr = orc.orthogonalize_moments(cov, 1e-4)
        close(l_mine, l_mat, rtol=1e-8, atol=1e-11, what="maxent L")
        phi_o = orth.eval_all(nodes)
        mu = pdf_w @ phi_o
        data = np.stack([mu, np.ones_like(mu)], axis=1)
        ref = FixedNodes(orth, data, domain=domain)
        res = ref.estimate_density_minimize(tol=1e-8, reg_param=0.0)
        basis = orc.Basis("legendre", size, domain, safe_eval=False, matrix=l_mat)
        fit = orc.maxent_fit(basis, data, domain, tol=1e-8, n_panels=n_panels)
        close(fit.multipliers, ref.multipliers, rtol=1e-9, atol=1e-11, what="multipliers %d" % size)
        xs = np.linspace(domain[0], domain[1], 41)
        close(orc.maxent_density(basis, fit.multipliers, np.ones(len(mu)), xs), ref.density(xs), rtol=1e-8, what="pdf")
        # one F/g/H evaluation at a non-trivial lambda
        lam = 0.7 * ref.multipliers + 0.05
        f_ref = ref._calculate_functional(lam)
        g_ref = ref._calculate_gradient(lam)
        h_ref = ref._calculate_jacobian_matrix(lam)
        sig = np.ones(len(mu))
        close(orc.maxent_functional(phi_o, w, lam, mu, sig), f_ref, what="F")
        close(orc.maxent_gradient(phi_o, w, lam, mu, sig), g_ref, what="g")
        close(orc.maxent_hessian(phi_o, w, lam, mu, sig), h_ref, what="H")
        t = "R%d_" % size
        out[t + "cov"] = cov
        out[t + "L"] = l_mat
        out[t + "mu"] = mu
        out[t + "multipliers"] = ref.multipliers
        out[t + "nit"] = np.array(res.nit)
        out[t + "density_x"] = xs
        out[t + "density"] = ref.density(xs)
        out[t + "cdf"] = ref.cdf(xs)
        out[t + "lam"] = lam
        out[t + "F"] = np.array(f_ref)
        out[t + "g"] = g_ref
        out[t + "H"] = h_ref
        out[t + "eigvals"] = res.eigvals
        print("  maxent R=%d: nit=%d |g|=%.2e pdf err=%.2e" % (size, res.nit, res.fun_norm,
                                                              np.abs(ref.density(xs) - distr.pdf(xs)).max()))
        # unmodified adaptive reference (QUADPACK node set) where it converges: multipliers agree loosely
        if size <= 15:
            ada = rsd.SimpleDistribution(orth, data, domain=domain)
            ada.estimate_density_minimize(tol=1e-8, reg_param=0.0)
            out[t + "adaptive_multipliers"] = ada.multipliers
            print("  adaptive-vs-fixed multipliers max diff %.2e" % np.abs(ada.multipliers - ref.multipliers).max())
    out["domain"] = np.array(domain)
    out["n_panels"] = np.array(n_panels)
    save("maxent", **out)


if __name__ == "__main__":
    basis_tables()
    sampling_pools_vector()
    estimates()
    maxent()
