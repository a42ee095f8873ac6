"""Maximum-entropy PDF reconstruction -- mirror of ``mlmc/tool/simple_distribution.py:9-327, 349-464, 756-841``.

``SimpleDistribution`` keeps the reference's constructor, attributes and result object.  What changes is where
the quadrature lives: the reference re-runs QUADPACK on a Python integrand at (almost) every Newton callback to
choose Gauss panels (``:198-238``, 99 % of its fit time and the reason it collapses to one panel at R = 50,
SURVEY.md section 8c); here the composite 21-point Gauss-Legendre rule has a FIXED number of equal panels, the
moment table ``Phi[Q, R]`` of its nodes is evaluated once on the device, and every callback is one fused kernel
(``mlmcb200_maxent_fgh``: exponent, ``F``, ``g`` and the DMMA contraction ``H = Phi^T diag(w rho) Phi``) plus a
single small device->host copy.  The Newton / CG iteration itself is scipy's ``trust-ncg`` exactly as in the
reference (``:65-69``).  With enough panels the multipliers agree with the adaptive reference to ~1e-15 where the
latter converges (tests/golden/make_golden.py).
"""
import warnings

import numpy as np
import scipy as sc
import scipy.integrate
import scipy.linalg
import scipy.optimize
import torch

from .. import _native
from .. import moments as moments_mod

EXACT_QUAD_LIMIT = 1000


def _device():
    if not torch.cuda.is_available():
        raise _native.NativeError("mlmc_b200 needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


_rule_cache = {}        # (domain, n_panels, degree) -> (nodes, weights); + device -> CUDA copies.  A few entries.
_pinned_cache = {}      # size -> pinned staging vectors of the fit loop


def gauss_panels(domain, n_panels, degree=21):
    """Composite Gauss-Legendre nodes / weights on ``n_panels`` equal panels of ``domain`` (the construction of
    ``_update_quadrature``, ``:222-231``, on a fixed partition).  Read-only arrays, memoised."""
    key = (float(domain[0]), float(domain[1]), int(n_panels), int(degree))
    hit = _rule_cache.get(key)
    if hit is None:
        pt, w = np.polynomial.legendre.leggauss(degree)
        edges = np.linspace(domain[0], domain[1], n_panels + 1)
        a, b = edges[:-1, None], edges[1:, None]
        nodes = ((pt[None, :] + 1) / 2 * (b - a) + a).ravel()
        weights = (w[None, :] * (b - a) / 2).ravel()
        nodes.setflags(write=False)
        weights.setflags(write=False)
        if len(_rule_cache) >= 16:
            _rule_cache.clear()
        hit = _rule_cache[key] = (nodes, weights)
    return hit


def _gauss_panels_on(device, domain, n_panels, degree):
    """The same rule as CUDA tensors (memoised per device)."""
    key = (float(domain[0]), float(domain[1]), int(n_panels), int(degree), str(device))
    hit = _rule_cache.get(key)
    if hit is None:
        nodes, weights = gauss_panels(domain, n_panels, degree)
        hit = _rule_cache[key] = (torch.from_numpy(nodes.copy()).to(device), torch.from_numpy(weights.copy()).to(device))
    return hit


def _pinned_vectors(size):
    hit = _pinned_cache.get(size)
    if hit is None:
        if len(_pinned_cache) >= 16:
            _pinned_cache.clear()
        hit = _pinned_cache[size] = (torch.empty(1 + size + size * size, dtype=torch.float64).pin_memory(),
                                     torch.empty(size, dtype=torch.float64).pin_memory())
    return hit


class _FitBuffers:
    """Device / pinned buffers of one (device, n_nodes, n_moments) fit shape, kept across fits, and the CUDA graph of
    one functional evaluation on them: H2D of the scaled multipliers -> rho / H / sum kernels -> D2H of [F | g | H].
    A fit then costs one graph launch + one stream synchronisation per evaluation instead of ~10 host calls."""

    def __init__(self, device, n_nodes, size):
        self.phi = torch.empty((n_nodes, size), dtype=torch.float64, device=device)
        self.w = torch.empty(n_nodes, dtype=torch.float64, device=device)
        self.lam = torch.zeros(size, dtype=torch.float64, device=device)
        self.out = torch.zeros(1 + size + size * size, dtype=torch.float64, device=device)
        self.out_host, self.lam_host = _pinned_vectors(size)
        self.out_np, self.lam_np = self.out_host.numpy(), self.lam_host.numpy()
        self.workspace = _native.maxent_workspace(self.phi, size)
        self.owner = None
        self.graph = None
        self.n_evals = 0

    def _enqueue(self):
        self.lam.copy_(self.lam_host, non_blocking=True)
        _native.maxent_fgh(self.phi, self.w, self.lam, 7, self.out, workspace=self.workspace)
        self.out_host.copy_(self.out, non_blocking=True)

    def evaluate(self):
        """``lam_np`` holds the scaled multipliers -> ``out_np`` = [F | g | H] (valid until the next call)."""
        self.n_evals += 1
        if self.graph is None and self.n_evals > 2 and _USE_GRAPH:
            # third evaluation on these buffers: worth a capture (the first two ran eagerly and warmed the workspace)
            torch.cuda.synchronize()
            side = torch.cuda.Stream(self.phi.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(graph, stream=side):
                    self._enqueue()
            self.graph = graph
        if self.graph is not None:
            self.graph.replay()
        else:
            self._enqueue()
        torch.cuda.current_stream().synchronize()
        return self.out_np


_USE_GRAPH = True
_fit_buffers = {}


def _fit_buffers_for(device, n_nodes, size):
    key = (str(device), int(n_nodes), int(size))
    hit = _fit_buffers.get(key)
    if hit is None:
        if len(_fit_buffers) >= 8:
            _fit_buffers.clear()
        hit = _fit_buffers[key] = _FitBuffers(device, n_nodes, size)
    return hit


class SimpleDistribution:
    """Calculation of the distribution (``simple_distribution.py:9-327``)."""

    #: default number of equal Gauss panels (x 21 nodes); cfg4 of BASELINE.json uses 4762 (100 002 nodes)
    QUAD_PANELS = 256

    def __init__(self, moments_obj, moment_data, domain=None, force_decay=(True, True), verbose=False,
                 quad_panels=None):
        self.moments_fn = None
        if domain is None:
            domain = moments_obj.domain
        self.domain = domain
        self.decay_penalty = force_decay
        self._verbose = verbose
        if moment_data is not None:
            moment_data = np.asarray(moment_data, dtype=np.float64)
            self.moment_means = moment_data[:, 0]
            self.moment_errs = np.sqrt(moment_data[:, 1])
        self.multipliers = None
        self.approx_size = len(self.moment_means)
        assert moments_obj.size >= self.approx_size
        self.moments_fn = moments_obj
        self._gauss_degree = 21
        self._penalty_coef = 0
        self._n_panels = int(quad_panels) if quad_panels is not None else self.QUAD_PANELS
        self._cache_key = None
        self._cache = None
        #: number of fused device evaluations of the last fit
        self.n_device_evals = 0

    # ---------------------------------------------------------------- fit
    def estimate_density_minimize(self, tol=1e-5, reg_param=0.01):
        """Minimise the max-ent functional with trust-ncg (``:50-94``); returns scipy's result object extended by
        ``eigvals, solver_res, fun_norm`` like the reference."""
        self._initialize_params(self.approx_size, tol)
        max_it = 20
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")        # scipy >= 1.1x: "Unknown solver options: tol, xtol"
            result = sc.optimize.minimize(self._calculate_functional, self.multipliers, method="trust-ncg",
                                          jac=self._calculate_gradient, hess=self._calculate_jacobian_matrix,
                                          options={"tol": tol, "xtol": tol, "gtol": tol, "disp": False,
                                                   "maxiter": max_it})
        self.multipliers = np.array(result.x)
        jac_norm = np.linalg.norm(result.jac)
        if self._verbose:
            print("size: {} nits: {} tol: {:5.3g} res: {:5.3g} msg: {}".format(
                self.approx_size, result.nit, tol, jac_norm, result.message))
        jac = self._calculate_jacobian_matrix(self.multipliers)
        result.eigvals = np.linalg.eigvalsh(jac)
        result.solver_res = result.jac
        # fix normalisation: lambda_0 -= log int rho   (``:82-86``), on the same fixed rule
        moment_0, _ = self._calculate_exact_moment(self.multipliers, m=0)
        self.multipliers[0] -= np.log(moment_0)
        self._cache_key = None
        if result.success or jac_norm < tol:
            result.success = True
        result.nit = max(result.nit, 1)
        result.fun_norm = jac_norm
        return result

    def _initialize_params(self, size, tol=None):
        assert self.domain is not None
        assert tol is not None
        self._quad_tolerance = 1e-10
        self._moment_errs = self.moment_errs
        self.multipliers = np.zeros(size)
        self.multipliers[0] = -np.log(1 / (self.domain[1] - self.domain[0]))      # uniform start (``:143-144``)
        self._quad_log = []
        self.n_device_evals = 0
        self._end_point_diff = self.end_point_derivatives() if self._penalty_coef != 0 else np.zeros((2, size))
        self._update_quadrature(self.multipliers, force=True)

    def _update_quadrature(self, multipliers, force=False):
        """Build the node set and its moment table on the device.  The reference rebuilds it from QUADPACK's
        panels whenever the multipliers moved (``:198-238``); the fixed rule is built once per fit."""
        if not force:
            return
        dev = _device()
        nodes, weights = gauss_panels(self.domain, self._n_panels, self._gauss_degree)
        self._quad_points = nodes
        self._quad_weights = weights
        self._nodes_dev, self._weights_dev = _gauss_panels_on(dev, self.domain, self._n_panels, self._gauss_degree)
        # the moment table and the weights go into buffers that persist across fits of this shape (and with them the
        # CUDA graph of one evaluation); this fit owns them until another fit of the same shape starts
        buf = self._buffers = _fit_buffers_for(dev, len(nodes), self.approx_size)
        buf.owner = self
        self.moments_fn._eval_device(self._nodes_dev, self.approx_size, out=buf.phi)
        buf.w.copy_(self._weights_dev)
        self._quad_moments_dev = buf.phi
        self._cache_key = None

    @property
    def _quad_moments(self):
        """Moment table of the nodes as a NumPy array (reference attribute name)."""
        if self._buffers.owner is not self:
            self._update_quadrature(self.multipliers, force=True)
        return self._quad_moments_dev.cpu().numpy()

    def _integrals(self, multipliers):
        """(int rho, int rho phi_i, int rho phi_i phi_j) at ``multipliers``: one fused launch, memoised on the
        multiplier vector because scipy asks for F, g and H at the same point separately."""
        multipliers = np.asarray(multipliers, dtype=np.float64)
        key = multipliers.tobytes()
        if key != self._cache_key:
            r = self.approx_size
            buf = self._buffers
            if buf.owner is not self:                      # another fit of the same shape took the buffers: take them back
                self._update_quadrature(multipliers, force=True)
                buf = self._buffers
            np.divide(multipliers, self._moment_errs, out=buf.lam_np)
            out = buf.evaluate()
            self._cache = (float(out[0]), out[1:1 + r].copy(), out[1 + r:].reshape(r, r).copy())
            self._cache_key = key
            self.n_device_evals += 1
        return self._cache

    def _density_in_quads(self, multipliers):
        if self._buffers.owner is not self:
            self._update_quadrature(multipliers, force=True)
        power = -(self._quad_moments_dev @ torch.from_numpy(np.asarray(multipliers / self._moment_errs)).to(
            self._quad_moments_dev.device))
        return torch.exp(torch.clamp(power, -200, 200)).cpu().numpy()

    def _calculate_exact_moment(self, multipliers, m=0, full_output=0):
        """Moment ``m`` of the current density on the fixed rule (stands in for the adaptive ``:157-174``)."""
        _f, g_int, _h = self._integrals(multipliers)
        return g_int[m], None

    def _calculate_functional(self, multipliers):
        """``F = sum mu lambda / sigma + int rho`` (``:259-275``; end-point penalty coefficient is 0, ``:48``)."""
        f_int, _g, _h = self._integrals(multipliers)
        fun = np.sum(self.moment_means * multipliers / self._moment_errs) + f_int
        if self._penalty_coef != 0:
            end_diff = np.dot(self._end_point_diff, multipliers)
            fun = fun + np.abs(fun) * self._penalty_coef * np.sum(np.maximum(end_diff, 0) ** 2)
        return fun

    def _calculate_gradient(self, multipliers):
        """``g = mu / sigma - (Phi^T (w rho)) / sigma`` (``:277-291``)."""
        _f, g_int, _h = self._integrals(multipliers)
        integral = g_int / self._moment_errs
        gradient = self.moment_means / self._moment_errs - integral
        if self._penalty_coef != 0:
            end_diff = np.dot(self._end_point_diff, multipliers)
            penalty = 2 * np.dot(np.maximum(end_diff, 0), self._end_point_diff)
            fun = np.sum(self.moment_means * multipliers / self._moment_errs) + integral[0] * self._moment_errs[0]
            gradient = gradient + np.abs(fun) * self._penalty_coef * penalty
        return gradient

    def _calculate_jacobian_matrix(self, multipliers):
        """``H = (Phi / sigma)^T diag(w rho) (Phi / sigma)`` (``:293-327``), symmetric ``[R, R]``."""
        _f, _g, h_int = self._integrals(multipliers)
        jacobian_matrix = h_int / np.outer(self._moment_errs, self._moment_errs)
        if self._penalty_coef != 0:
            end_diff = np.dot(self._end_point_diff, multipliers)
            fun = np.sum(self.moment_means * multipliers / self._moment_errs) + \
                jacobian_matrix[0, 0] * self._moment_errs[0] ** 2
            for side in [0, 1]:
                if end_diff[side] > 0:
                    penalty = 2 * np.outer(self._end_point_diff[side], self._end_point_diff[side])
                    jacobian_matrix = jacobian_matrix + np.abs(fun) * self._penalty_coef * penalty
        return jacobian_matrix

    # ---------------------------------------------------------------- evaluation
    def eval_moments(self, x):
        return self.moments_fn.eval_all(x, self.approx_size)

    def density(self, value):
        """``exp(clip(-phi(x) . lambda / sigma, -200, 200))`` (``:96-105``); NumPy in -> NumPy out."""
        as_numpy = not isinstance(value, torch.Tensor)
        if as_numpy:
            value = torch.from_numpy(np.atleast_1d(np.asarray(value, dtype=np.float64))).to(_device())
        out = _native.density_eval(self.moments_fn.basis_struct(), value.reshape(-1).to(torch.float64),
                                   self._density_coefficients(value.device))
        return out.cpu().numpy() if as_numpy else out

    def _density_coefficients(self, device):
        """``lambda / sigma`` in terms of the BASE functions (a TransformedMoments basis ``L phi``: ``L[:n]^T lambda / sigma``),
        for the one-launch ``mlmcb200_density_eval`` (no moment table, no library matmul)."""
        lam = np.asarray(self.multipliers / self._moment_errs, dtype=np.float64)
        l_mat = self.moments_fn.transform_matrix()
        if l_mat is not None:
            lam = l_mat[:self.approx_size].T @ lam
        return torch.from_numpy(np.ascontiguousarray(lam)).to(device)

    def cdf(self, values):
        """Cumulative 10-point Gauss sums between consecutive values (``:108-125``), all intervals evaluated in
        one batched density call."""
        values = np.atleast_1d(values)
        pt, w = np.polynomial.legendre.leggauss(10)
        spans = []
        last_x = self.domain[0]
        for val in values:
            if self.domain[0] < val < self.domain[1]:
                spans.append((last_x, val))
                last_x = val
        pieces = np.zeros(0)
        if spans:
            a = np.array([s[0] for s in spans])[:, None]
            b = np.array([s[1] for s in spans])[:, None]
            x = (b - a) * (np.real(pt)[None, :] + 1) / 2.0 + a
            dens = self.density(x.ravel()).reshape(x.shape)
            pieces = (b[:, 0] - a[:, 0]) / 2.0 * np.sum(w[None, :] * dens, axis=-1)
        cdf_y = np.empty(len(values))
        last_y, k = 0, 0
        for i, val in enumerate(values):
            if val <= self.domain[0]:
                last_y = 0
            elif val >= self.domain[1]:
                last_y = 1
            else:
                last_y = last_y + pieces[k]
                k += 1
            cdf_y[i] = last_y
        return cdf_y

    def end_point_derivatives(self):
        """Finite-difference moment derivatives at the domain end points (``:240-252``), shape (2, n_moments)."""
        eps = 1e-10
        left_diff = right_diff = np.zeros((1, self.approx_size))
        if self.decay_penalty[0]:
            left_diff = self.eval_moments(self.domain[0] + eps) - self.eval_moments(self.domain[0])
        if self.decay_penalty[1]:
            right_diff = -self.eval_moments(self.domain[1]) + self.eval_moments(self.domain[1] - eps)
        return np.stack((left_diff[0, :], right_diff[0, :]), axis=0) / eps / self.moment_errs[None, :]


# --------------------------------------------------------------------------------------------------------------
def _weighted_contraction(moments_fn, density, n_panels):
    """(int p phi_i, int p phi_i phi_j) of a host density callable on the fixed composite Gauss rule, contracted
    by the same fused kernel (multipliers = 0 => rho = 1, weights = w * p)."""
    dev = _device()
    nodes, weights = gauss_panels(moments_fn.domain, n_panels)
    nodes_dev, _ = _gauss_panels_on(dev, moments_fn.domain, n_panels, 21)
    pdf_w = np.asarray(density(nodes), dtype=np.float64) * weights
    phi = moments_fn.eval_all(nodes_dev).contiguous()
    r = phi.shape[1]
    out = _native.maxent_fgh(phi, torch.from_numpy(pdf_w).to(dev), torch.zeros(r, dtype=torch.float64, device=dev))
    out = out.cpu().numpy()
    return out[1:1 + r], out[1 + r:].reshape(r, r)


def compute_semiexact_moments(moments_fn, density, tol=1e-10, n_panels=SimpleDistribution.QUAD_PANELS):
    """Moments of an exact density on the Gauss rule (``:349-378``)."""
    return _weighted_contraction(moments_fn, density, n_panels)[0]


def compute_semiexact_cov(moments_fn, density, tol=1e-10, n_panels=SimpleDistribution.QUAD_PANELS):
    """Covariance matrix of the moment functions under an exact density (``:404-438``)."""
    return _weighted_contraction(moments_fn, density, n_panels)[1]


compute_exact_moments = compute_semiexact_moments
compute_exact_cov = compute_semiexact_cov


def KL_divergence(prior_density, posterior_density, a, b):
    """``int P log(P / Q) - P + Q`` (``:443-459``); host diagnostic."""
    def integrand(x):
        p = prior_density(x)
        q = max(posterior_density(x), 1e-300)
        return p * np.log(p / q) - p + q
    value = sc.integrate.quad(integrand, a, b, epsabs=1e-10)
    return max(value[0], 1e-10)


def L2_distance(prior_density, posterior_density, a, b):
    """(``:462-464``); host diagnostic."""
    return np.sqrt(sc.integrate.quad(lambda x: (posterior_density(x) - prior_density(x)) ** 2, a, b))[0]


def construct_ortogonal_moments(moments, cov, tol=None):
    """Basis orthogonal with respect to the sample covariance (``:756-841``): centre the covariance, drop
    eigen-directions below ``tol``, whiten, RQ-factorise -> ``TransformedMoments(moments, L)``.
    Returns ``(orthogonal_moments, (eigenvalues, threshold, L))``.  R x R host LAPACK, as in the reference."""
    if tol is None:
        raise NotImplementedError("automatic threshold detection (detect_treshold_slope_change, "
                                  "simple_distribution.py:538-609) is a diagnostic outside the hot path; pass tol")
    cov = np.asarray(cov, dtype=np.float64)
    center = np.eye(moments.size)
    center[:, 0] = -cov[:, 0]
    cov_center = center @ cov @ center.T
    evals, evecs = np.linalg.eigh(cov_center)
    threshold = np.argmax(evals > tol)
    kept_vals = np.flip(evals[threshold:], axis=0)
    kept_vecs = np.flip(evecs[:, threshold:], axis=1)
    icov_sqrt_t = center.T @ kept_vecs * (1 / np.sqrt(kept_vals))[None, :]
    r_nm, _q_mm = sc.linalg.rq(icov_sqrt_t, mode="full")
    l_mn = r_nm.T
    if l_mn[0, 0] < 0:
        l_mn = -l_mn
    ortogonal_moments = moments_mod.TransformedMoments(moments, l_mn)
    return ortogonal_moments, (evals, threshold, l_mn)
