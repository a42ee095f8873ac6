"""``Estimate`` facade and sample-allocation helpers -- mirror of ``mlmc/estimator.py:11-452``.

The heavy lifting (moments, covariance, level variances) is ``quantity_estimate.estimate_mean`` on the fused CUDA
path; what remains here is the reference's small host-side ``[L, R]`` arithmetic: the log-variance regression
(``estimator.py:87-134``), the allocation formula (``:366-385``), level-parameter helpers and the orchestration
of ``construct_density`` (``:304-331``).
"""
import numpy as np
import torch

from .quantity import quantity_estimate as qe
from .quantity.quantity_types import ScalarType
from .tool import simple_distribution


class Estimate:
    """Wrapper methods for moments estimation, PDF approximation, ... (``estimator.py:11-341``)."""

    def __init__(self, quantity, sample_storage, moments_fn=None):
        self._quantity = quantity
        self._sample_storage = sample_storage
        self._moments_fn = moments_fn

    @property
    def quantity(self):
        return self._quantity

    @quantity.setter
    def quantity(self, quantity):
        self._quantity = quantity

    @property
    def n_moments(self):
        return self._moments_fn.size

    # ---- the hot path ----
    def estimate_moments(self, moments_fn=None):
        """-> (means of the moments, variances of these estimates), arrays of length n_moments (``:32-42``)."""
        moments_fn = self._moments_fn if moments_fn is None else moments_fn
        est = qe.estimate_mean(qe.moments(self._quantity, moments_fn))
        return est.mean, est.var

    def estimate_covariance(self, moments_fn=None, variance=True):
        """-> (covariance matrix of the moments, variance of its entries), both [R, R] (``:44-54``).

        ``variance=False`` (extension): the entry variances are not needed (they come back as NaN); the covariance
        means are then a linear map of 2R-1 moment means (one pass of the fused moments kernel, ~20x cheaper than the
        DMMA contraction at R = 100)."""
        moments_fn = self._moments_fn if moments_fn is None else moments_fn
        est = qe.estimate_mean(qe.covariance(self._quantity, moments_fn), variance=variance)
        return est.mean, est.var

    def estimate_diff_vars(self, moments_fn=None):
        """-> (variances of the level differences [L, R], n_samples [L]) (``:76-85``)."""
        moments_fn = self._moments_fn if moments_fn is None else moments_fn
        est = qe.estimate_mean(qe.moments(self._quantity, moments_fn))
        return est.l_vars, est.n_samples

    def estimate_diff_vars_regression(self, n_created_samples, moments_fn=None, raw_vars=None):
        """Level variances smoothed by the regression model -> (vars [L, R], n_ops [L]) (``:56-74``)."""
        self._n_created_samples = n_created_samples
        if raw_vars is None:
            raw_vars, _n = self.estimate_diff_vars(self._moments_fn if moments_fn is None else moments_fn)
        sim_steps = np.squeeze(self._sample_storage.get_level_parameters())
        return self._all_moments_variance_regression(raw_vars, sim_steps), self._sample_storage.get_n_ops()

    def _all_moments_variance_regression(self, raw_vars, sim_steps):
        """Regression of every moment column (``:87-93``).  All columns share the design matrix, so they are
        solved as ONE multi-right-hand-side least-squares problem instead of R - 1 separate ones."""
        raw_vars = np.asarray(raw_vars, dtype=float)
        reg_vars = raw_vars.copy()
        n_levels, n_moments = raw_vars.shape
        if n_levels >= 3 and n_moments > 1:
            cols = np.flatnonzero(~(np.abs(raw_vars) <= 1e-8).all(axis=0))      # np.isclose(raw_vars, 0)
            cols = cols[cols >= 1]
            if len(cols):
                log_h = np.log(np.asarray(sim_steps, dtype=float)[1:])
                design = np.column_stack([np.ones(n_levels - 1), log_h, log_h ** 2])
                coef = np.linalg.lstsq(design, np.log(raw_vars[1:][:, cols]), rcond=None)[0]
                reg_vars[1:, cols] = np.exp(design @ coef)
        assert np.all(np.abs(reg_vars[:, 0]) <= 1e-8)
        return reg_vars

    def _moment_variance_regression(self, raw_vars, sim_steps):
        """Least squares of ``log V_l = A + B log h_l + C log^2 h_l`` over levels 1..L-1, unit weights
        (``:95-134``; the chi^2 weights of ``:111-112`` are overwritten by ones at ``:113``)."""
        n_levels, = raw_vars.shape
        if n_levels < 3 or np.allclose(raw_vars, 0):
            return raw_vars
        log_h = np.log(np.asarray(sim_steps, dtype=float)[1:])
        design = np.column_stack([np.ones(n_levels - 1), log_h, log_h ** 2])
        coef = np.linalg.lstsq(design, np.log(raw_vars[1:]), rcond=None)[0]
        fitted = np.array(raw_vars, dtype=float, copy=True)
        fitted[1:] = np.exp(design @ coef)
        return fitted

    def _variance_of_variance(self, n_samples=None):
        """Variance of log(X), X ~ chi^2 with ``n - 1`` degrees of freedom scaled by ``1 / (n - 1)`` (``:136-169``;
        tiny host quadrature, the reference computes it for its regression weights and then overrides them)."""
        import scipy.integrate
        import scipy.stats
        if n_samples is None:
            n_samples = self._n_created_samples
        if hasattr(self, "_saved_var_var"):
            ns, var_var = self._saved_var_var
            if len(ns) == len(n_samples) and np.sum(np.abs(np.array(ns) - np.array(n_samples))) == 0:
                return var_var
        out = []
        for ns in n_samples:
            df = ns - 1
            std_est = np.sqrt(2 / df)

            def log_chi_pdf(x, df=df):
                return np.exp(x) * df * scipy.stats.chi2.pdf(np.exp(x) * df, df=df)

            def moment(m, std_est=std_est):
                return scipy.integrate.quad(lambda x: x ** m * log_chi_pdf(x), -100 * std_est, 100 * std_est)[0]
            mean = moment(1)
            out.append(moment(2) - mean ** 2)
        out = np.array(out)
        self._saved_var_var = (n_samples, out)
        return out

    # ---- bootstrap (``:171-218``).  The reference's version cannot run (its ``select(subsample(...))`` indexes
    # with a float array, quantity.py:168-169); here the sub-sampled quantity is used directly.  For plain bases all
    # replicates are fused on the device (``quantity_estimate.bootstrap_moments``): one launch per level. ----
    def est_bootstrap(self, n_subsamples=100, sample_vector=None, moments_fn=None, seed=None):
        if moments_fn is not None:
            self._moments_fn = moments_fn
        moments_fn = self._moments_fn
        sample_vector = determine_sample_vec(self._sample_storage.get_n_collected(),
                                             self._sample_storage.get_n_levels(), sample_vector)
        if qe.can_fuse_bootstrap(self.quantity, moments_fn):
            stats = qe.bootstrap_moments(self.quantity, moments_fn, sample_vector, n_subsamples, seed=seed,
                                         mom_at_bottom=False)
        else:
            stats = {"mean": [], "var": [], "l_means": [], "l_vars": []}
            for _ in range(n_subsamples):
                sub = self.quantity.subsample(sample_vec=sample_vector)
                q_mean = qe.estimate_mean(qe.moments(sub, moments_fn=moments_fn, mom_at_bottom=False))
                for key in stats:
                    stats[key].append(getattr(q_mean, key))
        self.mean_bs_mean, self.mean_bs_var = np.mean(stats["mean"], axis=0), np.mean(stats["var"], axis=0)
        self.mean_bs_l_means = np.mean(stats["l_means"], axis=0)
        self.mean_bs_l_vars = np.mean(stats["l_vars"], axis=0)
        self.var_bs_mean, self.var_bs_var = np.var(stats["mean"], axis=0, ddof=1), np.var(stats["var"], axis=0, ddof=1)
        self.var_bs_l_means = np.var(stats["l_means"], axis=0, ddof=1)
        self.var_bs_l_vars = np.var(stats["l_vars"], axis=0, ddof=1)
        self._bs_level_mean_variance = self.var_bs_l_means * \
            np.array(self._sample_storage.get_n_collected())[:, None]

    def bs_target_var_n_estimated(self, target_var, sample_vec=None):
        sample_vec = determine_sample_vec(self._sample_storage.get_n_collected(),
                                          self._sample_storage.get_n_levels(), sample_vec)
        self.est_bootstrap(n_subsamples=300, sample_vector=sample_vec)
        variances, n_ops = self.estimate_diff_vars_regression(sample_vec, raw_vars=self.mean_bs_l_vars)
        return estimate_n_samples_for_target_variance(target_var, variances, n_ops,
                                                      n_levels=self._sample_storage.get_n_levels())

    # ---- domain / density ----
    @staticmethod
    def estimate_domain(quantity, sample_storage, quantile=None):
        """Moments domain from sample quantiles (``:275-302``).

        Reference behaviour kept for parity of the resulting domain: ``next(storage.chunks(n_samples=N_l))``
        carries no level id, so every iteration reads the LEVEL-0 fine column truncated to ``N_l`` rows."""
        if quantile is None:
            quantile = 0.01
        ranges = []
        n_collected = sample_storage.get_n_collected()
        for level_id in range(sample_storage.get_n_levels()):
            chunk_spec = next(sample_storage.chunks(n_samples=int(n_collected[level_id])))
            storage_q = quantity.get_quantity_storage()
            fine = quantity.device_samples(storage_q.device_chunk(chunk_spec))[..., 0].reshape(-1)
            ranges.append(_percentiles(fine, [100 * quantile, 100 * (1 - quantile)]))
        ranges = np.array(ranges)
        return np.min(ranges[:, 0]), np.max(ranges[:, 1])

    def construct_density(self, tol=1e-8, reg_param=0.0, orth_moments_tol=1e-4, exact_pdf=None):
        """Max-entropy PDF from the samples (``:304-331``) -> (distribution, info, result, orthogonal moments).

        The reference makes two full data passes: the covariance (of which it reads the means only, ``:311-312``),
        then the moments of the orthogonalised basis ``L phi`` (whose variances it overwrites with ones, ``:323``).
        Both are linear in the level sums of the base functions: ``phi_0 = 1`` makes ``mean(phi_i) = cov[i, 0]`` and
        ``mean(L phi) = L cov[:, 0]``, so ONE pass (``estimate_mean(..., variance=False)``:
        moment sums of the 2R-1 function basis, no contraction) yields everything the fit consumes."""
        if not isinstance(self._quantity.qtype, ScalarType):
            raise NotImplementedError("Currently, we only support ScalarType quantities")
        cov_mat = qe.estimate_mean(qe.covariance(self._quantity, self._moments_fn), variance=False).mean
        moments_obj, info = simple_distribution.construct_ortogonal_moments(self._moments_fn, cov_mat,
                                                                           tol=orth_moments_tol)
        if self._moments_fn.transform_matrix() is None:
            est_moments = info[2] @ cov_mat[:, 0]
        else:                                                   # phi_0 of a transformed basis need not be the constant 1
            est_moments = qe.estimate_mean(qe.moments(self._quantity, moments_obj)).mean
        est_vars = np.ones(moments_obj.size)
        moments_data = np.stack((est_moments, est_vars), axis=1)
        distr_obj = simple_distribution.SimpleDistribution(moments_obj, moments_data, domain=moments_obj.domain)
        result = distr_obj.estimate_density_minimize(tol, reg_param)
        return distr_obj, info, result, moments_obj

    def get_level_samples(self, level_id, n_samples=None):
        """NumPy ``[M, N, 1]`` (level 0) / ``[M, N, 2]`` samples of a level (``:333-341``)."""
        chunk_spec = next(self._sample_storage.chunks(level_id=level_id, n_samples=n_samples))
        return self._quantity.samples(chunk_spec=chunk_spec)


def _percentiles(values, percents):
    """``np.percentile(values, percents)`` (default linear interpolation) of a CUDA vector, NaN entries ignored: the
    two neighbouring order statistics of every percentile come from the device (exact radix selection,
    ``mlmcb200_percentile_stats``), then numpy's own lerp (``_lerp`` in numpy/lib/_function_base_impl.py) on the
    host -- the result is bit-identical to ``np.percentile`` of the non-NaN values."""
    from . import _native
    fracs = np.true_divide(np.asarray(percents, dtype=float), 100)
    stats, n = _native.percentile_stats(values, fracs)
    if n == 0:
        raise Exception("All samples were masked")
    out = []
    with np.errstate(invalid="ignore"):                  # infinite order statistics give NaN, as in numpy
        for frac, (a, b) in zip(fracs, stats):
            pos = frac * (n - 1)
            t = pos - np.floor(pos)
            diff = b - a
            out.append(b - diff * (1 - t) if t >= 0.5 else a + diff * t)
    return np.array(out)


def estimate_domain(quantity, sample_storage, quantile=None):
    """Module-level variant (``:344-363``): percentiles of EVERY level's own fine samples (at most ``N_0`` rows of
    each).  The reference's version cannot run (it passes an ``n_samples`` keyword its ``ChunkSpec`` does not have);
    this is what it evidently means.  Order statistics from the device (``mlmcb200_percentile_stats``)."""
    if quantile is None:
        quantile = 0.01
    n_first = int(sample_storage.get_n_collected()[0])
    storage_q = quantity.get_quantity_storage()
    ranges = []
    for level_id in range(sample_storage.get_n_levels()):
        chunk_spec = next(sample_storage.chunks(level_id=level_id, n_samples=n_first))
        fine = quantity.device_samples(storage_q.device_chunk(chunk_spec))[..., 0].reshape(-1)
        ranges.append(_percentiles(fine, [100 * quantile, 100 * (1 - quantile)]))
    ranges = np.array(ranges)
    return np.min(ranges[:, 0]), np.max(ranges[:, 1])


def calc_level_params(step_range, n_levels):
    """``:388-398``: geometric sequence of level steps as ``[[h_0], [h_1], ...]``."""
    assert step_range[0] > step_range[1]
    level_parameters = []
    for i_level in range(n_levels):
        level_param = 1 if n_levels == 1 else i_level / (n_levels - 1)
        level_parameters.append([step_range[0] ** (1 - level_param) * step_range[1] ** level_param])
    return level_parameters


def estimate_n_samples_for_target_variance(target_variance, prescribe_vars, n_ops, n_levels):
    """Optimal samples per level for a target variance, maximum over moments (``estimator.py:366-385``):
    ``n_l = max_r clip(round(sqrt(V_lr / C_l) * sum_k sqrt(V_kr C_k) / eps), 2, V_lr L / eps)``."""
    variances = np.asarray(prescribe_vars, dtype=float)
    n_ops = np.asarray(n_ops, dtype=float)
    root_vc = np.sqrt(variances.T * n_ops)                       # [R, L]
    total = np.sum(root_vc, axis=1)                              # [R]
    estimate = np.round((root_vc / n_ops).T * total / target_variance).astype(int)       # [L, R]
    capped = np.maximum(np.minimum(estimate, variances * n_levels / target_variance), 2)
    return np.max(capped, axis=1).astype(int)


def determine_level_parameters(n_levels, step_range):
    """Geometric sequence of simulation steps between ``step_range[0] > step_range[1]`` (``:409-426``)."""
    assert step_range[0] > step_range[1]
    params = []
    for i_level in range(n_levels):
        frac = 1 if n_levels == 1 else i_level / (n_levels - 1)
        params.append([step_range[0] ** (1 - frac) * step_range[1] ** frac])
    return params


def determine_sample_vec(n_collected_samples, n_levels, sample_vector=None):
    if sample_vector is None:
        sample_vector = n_collected_samples
    if len(sample_vector) > n_levels:
        sample_vector = sample_vector[:n_levels]
    return np.array(sample_vector)


def determine_n_samples(n_levels, n_samples=None):
    """Target samples per level: geometric interpolation between the first and last level (``:429-450``)."""
    if n_samples is None:
        n_samples = [100, 3]
    n_samples = np.atleast_1d(n_samples)
    if len(n_samples) == 1:
        n_samples = np.array([n_samples[0], 3])
    if len(n_samples) == 2:
        n0, n_last = n_samples
        n_samples = np.round(np.exp2(np.linspace(np.log2(n0), np.log2(n_last), n_levels))).astype(int)
    return n_samples
