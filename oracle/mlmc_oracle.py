"""CPU oracle for the MLMC estimation hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-NumPy restatement of the reference algorithm (GeoMop/MLMC v1.0.2, ``/root/reference``) for the
path named by BASELINE.json ``north_star``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module, and only as the
checker or the timed CPU baseline.  The product (``mlmc_b200``) never imports it.

Pinning: this oracle is checked against the reference itself (imported through ``tests/golden/ref_shim.py``
in the build container) by ``tests/golden/make_golden.py`` and against the reference's own golden
vectors (``test/test_moments.py:39-70`` closed forms, ``test/test_sampling_pools.py:18`` ``ref_means``);
the resulting fixtures live in ``tests/golden/*.npz`` and ``tests/test_oracle_golden.py`` replays them
without the reference.  Parity is therefore PINNED for every function below.

Third-party arithmetic the reference leans on (not vendored under /root/reference, un-pinned in its
``requirements.txt:1-2``; versions in this image: numpy 2.3.5, scipy 1.18.1) is restated here from the
published formulas:
  * ``numpy.polynomial.legendre.legvander``:  v[0]=1, v[1]=x, v[i]=(v[i-1]*x*(2i-1) - v[i-2]*(i-1))/i
  * ``numpy.polynomial.polynomial.polyvander``: v[0]=1, v[i]=v[i-1]*x
  * ``numpy.ma.masked_outside`` + ``filled(nan)``: closed interval test, NaN passes through
``scipy.optimize.minimize(method='trust-ncg')`` and ``scipy.linalg.rq`` / ``numpy.linalg.eigh`` are
called as the reference calls them (host LAPACK/optimizer, not on the data-parallel path).

Layout conventions: storage rows per level ``float64[N, 2, M]`` (sample, fine/coarse,
component); a chunk is the transposed view ``[M, n, 2]`` (level 0: ``[M, n, 1]``).
"""
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple
import numpy as np

LEGENDRE, MONOMIAL, FOURIER = "legendre", "monomial", "fourier"
_DEFAULT_REF = {LEGENDRE: (-1.0, 1.0), MONOMIAL: (0.0, 1.0), FOURIER: (0.0, 2 * np.pi)}


# --------------------------------------------------------------------------------------------------
# a1-a5: moment bases
# --------------------------------------------------------------------------------------------------
@dataclass
class Basis:
    """Parameters of one ``mlmc.moments`` object (``moments.py:10-39``; subclasses ``:111-229``)."""
    kind: str
    size: int
    domain: Tuple[float, float]
    ref_domain: Optional[Tuple[float, float]] = None
    log: bool = False
    safe_eval: bool = True
    matrix: Optional[np.ndarray] = None      # TransformedMoments: new = matrix @ old (moments.py:232-259)
    scale: float = field(init=False)
    shift: float = field(init=False)

    def __post_init__(self):
        if self.ref_domain is None:
            self.ref_domain = _DEFAULT_REF[self.kind]
        lo, hi = (np.log(self.domain[0]), np.log(self.domain[1])) if self.log else self.domain
        width = hi - lo
        assert width > 0
        width = max(width, 1e-15)                                   # moments.py:24
        self.scale = (self.ref_domain[1] - self.ref_domain[0]) / width
        self.shift = lo
        if self.matrix is not None:
            self.matrix = np.asarray(self.matrix, dtype=np.float64)
            assert self.matrix.shape[1] == self.size

    @property
    def out_size(self):
        return self.size if self.matrix is None else self.matrix.shape[0]


def to_ref_domain(b: Basis, value):
    """``Moments.transform`` (moments.py:28-39, 58-73): optional log, affine map, optional clip->NaN."""
    v = np.asarray(value, dtype=np.float64)
    if b.log:
        with np.errstate(divide="ignore", invalid="ignore"):
            v = np.log(v)
    t = (v - b.shift) * b.scale + b.ref_domain[0]
    if b.safe_eval:
        outside = (t < b.ref_domain[0]) | (t > b.ref_domain[1])    # NaN compares False -> stays NaN
        t = np.where(outside, np.nan, t)
    return t


def legendre_table(t, size):
    """numpy ``legvander`` recurrence, moment index last (moments.py:195-197)."""
    t = np.asarray(t, dtype=np.float64)
    out = np.empty((size,) + t.shape)
    out[0] = t * 0 + 1
    if size > 1:
        out[1] = t
        for i in range(2, size):
            out[i] = (out[i - 1] * t * (2 * i - 1) - out[i - 2] * (i - 1)) / i
    return np.moveaxis(out, 0, -1)


def monomial_table(t, size):
    """numpy ``polyvander`` recurrence (moments.py:122-126)."""
    t = np.asarray(t, dtype=np.float64)
    out = np.empty((size,) + t.shape)
    out[0] = t * 0 + 1
    if size > 1:
        out[1] = t
        for i in range(2, size):
            out[i] = out[i - 1] * t
    return np.moveaxis(out, 0, -1)


def fourier_table(t, size):
    """``Fourier._eval_all`` (moments.py:145-162): col 0 = 1, odd cols cos(k t), even cols sin(k t).

    The reference only accepts 1-D input (``np.outer`` + ``len(t)``); the same formula is applied
    here over any shape (the cfg5 oracle ravels and reshapes)."""
    t = np.asarray(t, dtype=np.float64)
    half = int(size / 2)
    k = np.arange(1, half + 1)
    kx = t[..., None] * k
    out = np.empty(t.shape + (size,))
    out[..., 0] = 1
    out[..., 1::2] = np.cos(kx)
    out[..., 2::2] = np.sin(kx[..., : half - (1 - int(size % 2))])
    return out


def basis_eval(b: Basis, value, size=None):
    """``Moments.eval_all`` (moments.py:90-93): ``[...] -> [..., size]``."""
    n_base = b.size
    if b.matrix is None and size is not None:
        n_base = size
    t = to_ref_domain(b, np.atleast_1d(value))
    table = {LEGENDRE: legendre_table, MONOMIAL: monomial_table, FOURIER: fourier_table}[b.kind](t, n_base)
    if b.matrix is None:
        return table
    out = np.matmul(table, b.matrix.T)                              # moments.py:256-259
    return out[..., : (b.out_size if size is None else size)]


# --------------------------------------------------------------------------------------------------
# a7-a8: chunk operations of the lazy "moments" / "covariance" quantities
# --------------------------------------------------------------------------------------------------
def moments_chunk(b: Basis, x, mom_at_bottom=True):
    """``qe.moments.eval_moments`` (quantity_estimate.py:105-110): ``[M,n,S] -> [M*R, n, S]``."""
    phi = basis_eval(b, x)                                           # [M, n, S, R]
    phi = phi.transpose((0, 3, 1, 2)) if mom_at_bottom else phi.transpose((3, 0, 1, 2))
    return phi.reshape((-1,) + phi.shape[-2:])


def covariance_chunk(b: Basis, x, cov_at_bottom=True):
    """``qe.covariance.eval_cov`` (quantity_estimate.py:131-147): ``[M,n,S] -> [M*R*R, n, S]``."""
    phi = basis_eval(b, x)                                           # [M, n, S, R]
    sides = [np.einsum("...i,...j", phi[..., s, :], phi[..., s, :]) for s in range(phi.shape[-2])]
    cov = np.array(sides)                                            # [S, M, n, R, R]
    cov = cov.transpose((1, 3, 4, 2, 0)) if cov_at_bottom else cov.transpose((3, 4, 1, 2, 0))
    return cov.reshape((-1,) + cov.shape[-2:])


def drop_nan_samples(chunk):
    """``mask_nan_samples`` (quantity_estimate.py:6-14): drop a sample if ANY entry (fine or coarse) is NaN."""
    bad = np.isnan(chunk).any(axis=0).any(axis=1)
    return chunk[:, ~bad, :], int(np.count_nonzero(bad))


# --------------------------------------------------------------------------------------------------
# a9-a11: level sums -> means / variances
# --------------------------------------------------------------------------------------------------
@dataclass
class LevelEstimate:
    """What ``QuantityMean`` carries (quantity.py:568-630)."""
    l_means: np.ndarray      # [L, K]
    l_vars: np.ndarray       # [L, K]
    n_samples: np.ndarray    # [L]
    n_rm_samples: np.ndarray  # [L]

    @property
    def mean(self):
        return np.sum(self.l_means, axis=0)                          # quantity.py:592

    @property
    def var(self):
        return np.sum(self.l_vars / self.n_samples[:, None], axis=0)  # quantity.py:593


def level_chunks(rows, level_id, chunk_rows):
    """Yield chunk views ``[M, n, 2 or 1]`` of one level's storage rows ``[N, 2, M]``
    (``sample_storage_hdf.py:169-184``: level 0 drops the auxiliary zero coarse row)."""
    n = len(rows)
    for start in range(0, max(n, 1), chunk_rows):
        part = rows[start:start + chunk_rows]
        if level_id == 0:
            part = part[:, :1, :]
        yield part.transpose((2, 0, 1))


def estimate_mean(levels: Sequence[np.ndarray], chunk_op=None, chunk_rows=65536) -> LevelEstimate:
    """``estimate_mean`` (quantity_estimate.py:22-80) over array-backed levels.

    :param levels: per level ``float64[N_l, 2, M]`` storage rows
    :param chunk_op: ``[M,n,S] -> [K,n,S]`` (e.g. ``lambda x: moments_chunk(b, x)``); None = identity
    """
    n_levels = len(levels)
    sums = None
    sums_sq = None
    n_samples = [0] * n_levels
    n_removed = [0] * n_levels
    for level_id, rows in enumerate(levels):
        for x in level_chunks(rows, level_id, chunk_rows):
            y = x if chunk_op is None else chunk_op(x)
            y, n_bad = drop_nan_samples(y)
            n_samples[level_id] += y.shape[1]
            n_removed[level_id] += n_bad
            if y.shape[1] == 0:
                continue
            if sums is None:
                sums = [np.zeros(y.shape[0]) for _ in range(n_levels)]
                sums_sq = [np.zeros(y.shape[0]) for _ in range(n_levels)]
            d = y[:, :, 0] if level_id == 0 else y[:, :, 0] - y[:, :, 1]
            sums[level_id] += np.sum(d, axis=1)
            sums_sq[level_id] += np.sum(d ** 2, axis=1)
    if sums is None:
        raise Exception("All samples were masked")
    l_means, l_vars = [], []
    for s, sq, n in zip(sums, sums_sq, n_samples):
        with np.errstate(divide="ignore", invalid="ignore"):
            l_means.append(s / n)
            l_vars.append((sq - (s ** 2 / n)) / (n - 1) if n > 1 else np.full(len(s), np.inf))
    return LevelEstimate(np.array(l_means), np.array(l_vars), np.array(n_samples), np.array(n_removed))


def estimate_moments(levels, b: Basis, chunk_rows=65536, mom_at_bottom=True):
    """``Estimate.estimate_moments`` / ``estimate_diff_vars`` (estimator.py:32-42, 76-85)."""
    return estimate_mean(levels, lambda x: moments_chunk(b, x, mom_at_bottom), chunk_rows)


def estimate_covariance(levels, b: Basis, chunk_rows=65536):
    """``Estimate.estimate_covariance`` (estimator.py:44-54); entries flat m*R*R + i*R + j."""
    return estimate_mean(levels, lambda x: covariance_chunk(b, x), chunk_rows)


# --------------------------------------------------------------------------------------------------
# a13-a14: level-variance regression and sample allocation (tiny host math)
# --------------------------------------------------------------------------------------------------
def regress_level_variances(raw_vars, sim_steps):
    """``_all_moments_variance_regression`` + ``_moment_variance_regression`` (estimator.py:87-134).

    For every moment m >= 1 with L >= 3 and non-zero variances: least-squares fit of
    ``log V_l = A + B log h_l + C log^2 h_l`` on levels 1..L-1 with unit weights (the chi^2 weights
    are computed and then overwritten by ones at estimator.py:111-113)."""
    raw_vars = np.asarray(raw_vars, dtype=np.float64)
    sim_steps = np.asarray(sim_steps, dtype=np.float64)
    out = raw_vars.copy()
    n_levels = raw_vars.shape[0]
    for m in range(1, raw_vars.shape[1]):
        col = raw_vars[:, m]
        if n_levels < 3 or np.allclose(col, 0):
            continue
        log_h = np.log(sim_steps[1:])
        design = np.stack([np.ones(n_levels - 1), log_h, log_h ** 2], axis=1)
        coef = np.linalg.lstsq(design, np.log(col[1:]), rcond=None)[0]
        out[1:, m] = np.exp(design @ coef)
    assert np.allclose(out[:, 0], 0.0)
    return out


def n_samples_for_target_variance(target_variance, level_vars, n_ops, n_levels):
    """``estimate_n_samples_for_target_variance`` (estimator.py:366-385)."""
    level_vars = np.asarray(level_vars, dtype=np.float64)
    n_ops = np.asarray(n_ops, dtype=np.float64)
    root = np.sqrt(level_vars.T * n_ops)
    total = np.sum(root, axis=1)
    est = np.round((root / n_ops).T * total / target_variance).astype(int)
    safe = np.maximum(np.minimum(est, level_vars * n_levels / target_variance), 2)
    return np.max(safe, axis=1).astype(int)


def estimate_domain(levels, quantile=None, column=0):
    """``Estimate.estimate_domain`` (estimator.py:275-302) for one scalar column.

    Reference quirk kept on purpose: ``next(storage.chunks(n_samples=N_l))`` has no level id, so every
    iteration reads the LEVEL-0 fine column truncated to N_l rows."""
    if quantile is None:
        quantile = 0.01
    ranges = []
    for rows in levels:
        fine = levels[0][: len(rows), 0, column]
        fine = fine[~np.isnan(fine)]
        ranges.append(np.percentile(fine, [100 * quantile, 100 * (1 - quantile)]))
    ranges = np.array(ranges)
    return np.min(ranges[:, 0]), np.max(ranges[:, 1])


# --------------------------------------------------------------------------------------------------
# Synthetic inputs (SynthSimulation.sample_fn, mlmc/sim/synth_simulation.py:38-46, 76-131)
# --------------------------------------------------------------------------------------------------
def level_steps(n_levels, step_range):
    """``determine_level_parameters`` (estimator.py:409-426), flattened to a 1-D list."""
    out = []
    for i in range(n_levels):
        p = 1 if n_levels == 1 else i / (n_levels - 1)
        out.append(step_range[0] ** (1 - p) * step_range[1] ** p)
    return out


def synth_level_rows(x, h_fine, h_coarse):
    """fine = x + h sqrt(1e-4+|x|); coarse likewise with the coarser step, zeros on level 0."""
    x = np.asarray(x, dtype=np.float64)
    root = np.sqrt(1e-4 + np.abs(x))
    fine = x + h_fine * root
    coarse = np.zeros_like(x) if h_coarse is None or h_coarse == 0 else x + h_coarse * root
    return np.stack([fine, coarse], axis=1)[:, :, None]


def synth_n_ops(step, complexity=2):
    """``SynthSimulation.n_ops_estimate`` (synth_simulation.py:133-134)."""
    return (1 / step) ** complexity * np.log(max(1 / step, 2.0))


# --------------------------------------------------------------------------------------------------
# a15-a20: maximum-entropy PDF reconstruction on a FIXED quadrature rule
# --------------------------------------------------------------------------------------------------
def gauss_panels(domain, n_panels, degree=21):
    """Composite Gauss-Legendre rule: what ``_update_quadrature`` (simple_distribution.py:222-231) builds
    from QUADPACK's panels, here on ``n_panels`` equal panels (fixed-node restatement)."""
    pt, w = np.polynomial.legendre.leggauss(degree)
    edges = np.linspace(domain[0], domain[1], n_panels + 1)
    a = edges[:-1, None]
    b = edges[1:, None]
    nodes = (pt[None, :] + 1) / 2 * (b - a) + a
    weights = w[None, :] * (b - a) / 2
    return nodes.ravel(), weights.ravel()


def maxent_density_at_nodes(phi, lam, sigma):
    """``_density_in_quads`` (simple_distribution.py:254-257)."""
    power = -np.dot(phi, lam / sigma)
    return np.exp(np.minimum(np.maximum(power, -200), 200))


def maxent_functional(phi, w, lam, mu, sigma):
    """``_calculate_functional`` (simple_distribution.py:259-275), penalty coefficient 0 (``:48``)."""
    rho = maxent_density_at_nodes(phi, lam, sigma)
    return np.sum(mu * lam / sigma) + np.dot(rho, w)


def maxent_gradient(phi, w, lam, mu, sigma):
    """``_calculate_gradient`` (simple_distribution.py:277-291)."""
    rho = maxent_density_at_nodes(phi, lam, sigma)
    return mu / sigma - np.dot(phi.T * rho, w) / sigma


def maxent_hessian(phi, w, lam, mu, sigma):
    """``_calculate_jacobian_matrix`` (simple_distribution.py:293-327)."""
    rho_w = maxent_density_at_nodes(phi, lam, sigma) * w
    scaled = phi / sigma
    return (scaled.T * rho_w) @ scaled


@dataclass
class MaxEntFit:
    multipliers: np.ndarray
    nit: int
    success: bool
    fun_norm: float
    eigvals: np.ndarray
    n_evals: Tuple[int, int, int]


def maxent_fit(b: Basis, moment_data, domain=None, tol=1e-8, n_panels=100, degree=21, max_it=20):
    """``SimpleDistribution.estimate_density_minimize`` (simple_distribution.py:50-94) with the node set
    fixed to ``n_panels`` x ``degree`` Gauss points; the post-fit normalisation ``lambda_0 -= log int rho``
    (``:82-86``) uses the same rule."""
    import warnings
    import scipy.optimize
    moment_data = np.asarray(moment_data, dtype=np.float64)
    mu, sigma = moment_data[:, 0], np.sqrt(moment_data[:, 1])
    size = len(mu)
    if domain is None:
        domain = b.domain
    nodes, w = gauss_panels(domain, n_panels, degree)
    phi = basis_eval(b, nodes, size)
    lam0 = np.zeros(size)
    lam0[0] = -np.log(1 / (domain[1] - domain[0]))                   # simple_distribution.py:143-144
    counts = [0, 0, 0]

    def fun(lam):
        counts[0] += 1
        return maxent_functional(phi, w, lam, mu, sigma)

    def jac(lam):
        counts[1] += 1
        return maxent_gradient(phi, w, lam, mu, sigma)

    def hess(lam):
        counts[2] += 1
        return maxent_hessian(phi, w, lam, mu, sigma)

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = scipy.optimize.minimize(fun, lam0, method="trust-ncg", jac=jac, hess=hess,
                                      options={"tol": tol, "xtol": tol, "gtol": tol, "disp": False,
                                               "maxiter": max_it})
    lam = np.array(res.x)
    jac_norm = float(np.linalg.norm(res.jac))
    eigvals = np.linalg.eigvalsh(maxent_hessian(phi, w, lam, mu, sigma))
    moment_0 = np.dot(maxent_density_at_nodes(phi, lam, sigma) * phi[:, 0], w)
    lam[0] -= np.log(moment_0)
    return MaxEntFit(lam, max(int(res.nit), 1), bool(res.success or jac_norm < tol), jac_norm, eigvals,
                     tuple(counts))


def maxent_density(b: Basis, lam, sigma, value):
    """``SimpleDistribution.density`` (simple_distribution.py:96-105)."""
    phi = basis_eval(b, value, len(lam))
    power = -np.sum(phi * lam / sigma, axis=1)
    return np.exp(np.minimum(np.maximum(power, -200), 200))


def orthogonalize_moments(cov, tol):
    """``construct_ortogonal_moments`` (simple_distribution.py:756-841) with an explicit ``tol``:
    returns ``(L, eigenvalues, threshold)``; the new basis is ``TransformedMoments(base, L)``."""
    import scipy.linalg
    cov = np.asarray(cov, dtype=np.float64)
    size = cov.shape[0]
    center = np.eye(size)
    center[:, 0] = -cov[:, 0]
    cov_center = center @ cov @ center.T
    evals, evecs = np.linalg.eigh(cov_center)
    threshold = int(np.argmax(evals > tol))
    kept_vals = np.flip(evals[threshold:], axis=0)
    kept_vecs = np.flip(evecs[:, threshold:], axis=1)
    icov_sqrt_t = center.T @ kept_vecs * (1 / np.sqrt(kept_vals))[None, :]
    r_nm, _ = scipy.linalg.rq(icov_sqrt_t, mode="full")
    l_mn = r_nm.T
    if l_mn[0, 0] < 0:
        l_mn = -l_mn
    return l_mn, evals, threshold
