"""Generalized moment functions -- B200 mirror of ``mlmc/moments.py`` (reference v1.0.2).

Same class names, constructor signatures, attributes (``size, domain, ref_domain, _is_log, _is_clip,
_linear_scale, _linear_shift``) and methods (``__call__, eval_all, eval, eval_single_moment, eval_all_der,
eval_diff, eval_diff2, change_size, transform, inv_transform, __eq__``) as the reference
(``mlmc/moments.py:6-274``).  The tables are produced by the CUDA kernel ``mlmcb200_basis_eval``; the fused
estimators (``mlmc_b200.quantity.quantity_estimate``) never materialise them and only take the parameter block
``basis_struct()`` from these objects.

``eval_all`` accepts NumPy arrays / scalars (result: NumPy, like the reference) or CUDA tensors (result: CUDA
tensor, no host copy).  Unlike the reference's ``Fourier._eval_all`` (1-D only, ``moments.py:153-161``) every
basis accepts any input shape and appends the moment axis.
"""
import numpy as np
import torch

from . import _native

__all__ = ["Moments", "Monomial", "Fourier", "Legendre", "TransformedMoments"]


def _device():
    if not torch.cuda.is_available():
        raise _native.NativeError("mlmc_b200 needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class Moments:
    """Base class: affine (optionally log) map of ``domain`` to ``ref_domain`` with optional clipping to NaN
    (``mlmc/moments.py:10-39, 58-73``)."""

    _kind = None

    def __init__(self, size, domain, log=False, safe_eval=True):
        assert size > 0
        self.size = size
        self.domain = domain
        self._is_log = log
        self._is_clip = safe_eval
        lin_domain = (np.log(domain[0]), np.log(domain[1])) if log else domain
        width = lin_domain[1] - lin_domain[0]
        assert width > 0
        width = max(width, 1e-15)
        self._linear_scale = (self.ref_domain[1] - self.ref_domain[0]) / width
        self._linear_shift = lin_domain[0]

    # ---- host-side helpers kept for API compatibility (not on the data path) ----
    def transform(self, value):
        value = np.asarray(value, dtype=np.float64)
        if self._is_log:
            with np.errstate(divide="ignore", invalid="ignore"):
                value = np.log(value)
        ref = self.linear(value)
        return self.clip(ref) if self._is_clip else ref

    def inv_transform(self, ref):
        value = self.inv_linear(np.asarray(ref, dtype=np.float64))
        return np.exp(value) if self._is_log else value

    def clip(self, value):
        value = np.asarray(value, dtype=np.float64)
        outside = (value < self.ref_domain[0]) | (value > self.ref_domain[1])
        return np.where(outside, np.nan, value)

    def linear(self, value):
        return (value - self._linear_shift) * self._linear_scale + self.ref_domain[0]

    def inv_linear(self, value):
        return (value - self.ref_domain[0]) / self._linear_scale + self._linear_shift

    def __eq__(self, other):
        return type(self) is type(other) \
            and self.size == other.size \
            and np.all(np.asarray(self.domain) == np.asarray(other.domain)) \
            and self._is_log == other._is_log \
            and self._is_clip == other._is_clip

    __hash__ = object.__hash__

    def change_size(self, size):
        return self.__class__(size, self.domain, log=self._is_log, safe_eval=self._is_clip)

    # ---- device side ----
    def basis_struct(self, size=None):
        """Parameter block passed across the C ABI (``mlmcb200_basis_t``)."""
        return _native.BasisStruct(self._kind, int(self.size if size is None else size), int(bool(self._is_log)),
                                   int(bool(self._is_clip)), float(self._linear_shift), float(self._linear_scale),
                                   float(self.ref_domain[0]), float(self.ref_domain[1]))

    def transform_matrix(self):
        """Linear map applied on top of the base functions (None for plain bases)."""
        return None

    def base_moments(self):
        return self

    # ---- products of the basis functions as linear combinations of a longer basis of the same family ----
    def _product_coefficients(self):
        """-> (size of the extended basis E, C[R, R, E]) with ``phi_i phi_j = sum_k C[i, j, k] phi_k``; None where
        the family has no such closed form."""
        return None

    def product_table(self):
        """Linearisation of the outer product: ``(extended moments object, C_t)`` with
        ``C_t[k, i * R + j]`` float64 ``[E, R*R]`` such that ``phi_i(x) phi_j(x) = sum_k C_t[k, i R + j] phi^E_k(x)``
        for every x -- and, because both sides are NaN for exactly the same samples (the mask depends on the domain
        only), ``estimate_mean(covariance(q, fn)).mean = C . estimate_mean(moments(q, fn^E)).mean`` level by level.
        Memoised.  None if the family does not linearise."""
        hit = getattr(self, "_product_table", None)
        if hit is None:
            coef = self._product_coefficients()
            if coef is None:
                return None
            ext, c = coef
            r = self.size
            hit = (self.change_size(ext), np.ascontiguousarray(c.reshape(r * r, ext).T))
            self._product_table = hit
        return hit

    def _eval_device(self, value, size, out=None):
        """value: CUDA float64 tensor of any shape -> CUDA tensor ``value.shape + (size,)`` (``out``: optional
        contiguous ``[value.numel(), size]`` tensor to write into)."""
        flat = value.reshape(-1)
        out = _native.basis_eval(self.basis_struct(size), flat, size, out=out)
        return out.reshape(tuple(value.shape) + (size,))

    def _eval_all(self, value, size):
        if isinstance(value, torch.Tensor):
            if not value.is_cuda:
                value = value.to(_device())
            return self._eval_device(value.to(torch.float64), size)
        host = np.atleast_1d(np.asarray(value, dtype=np.float64))
        dev_val = torch.from_numpy(np.ascontiguousarray(host)).to(_device())
        return self._eval_device(dev_val, size).cpu().numpy()

    def __call__(self, value):
        return self._eval_all(value, self.size)

    def eval_all(self, value, size=None):
        return self._eval_all(value, self.size if size is None else size)

    def eval(self, i, value):
        return self._eval_all(value, i + 1)[..., -1]

    def eval_single_moment(self, i, value):
        return self._eval_all(value, i + 1)[..., i]

    def eval_all_der(self, value, size=None, degree=1):
        return self._eval_all_der(value, self.size if size is None else size, degree)

    def eval_diff(self, value, size=None):
        return self._eval_diff(value, self.size if size is None else size)

    def eval_diff2(self, value, size=None):
        return self._eval_diff2(value, self.size if size is None else size)


class Monomial(Moments):
    """``t^k`` on ref_domain (0, 1) (``mlmc/moments.py:111-130``)."""
    _kind = _native.MONOMIAL

    def __init__(self, size, domain=(0, 1), ref_domain=None, log=False, safe_eval=True):
        self.ref_domain = ref_domain if ref_domain is not None else (0, 1)
        super().__init__(size, domain, log=log, safe_eval=safe_eval)

    def eval(self, i, value):
        return self._eval_all(value, i + 1)[..., i]

    def _product_coefficients(self):
        """``t^i t^j = t^(i+j)``"""
        r = self.size
        ext = 2 * r - 1
        c = np.zeros((r, r, ext))
        i, j = np.meshgrid(np.arange(r), np.arange(r), indexing="ij")
        c[i, j, i + j] = 1.0
        return ext, c


class Fourier(Moments):
    """``1, cos t, sin t, cos 2t, ...`` on ref_domain (0, 2 pi) (``mlmc/moments.py:133-162``)."""
    _kind = _native.FOURIER

    def __init__(self, size, domain=(0, 2 * np.pi), ref_domain=None, log=False, safe_eval=True):
        self.ref_domain = ref_domain if ref_domain is not None else (0, 2 * np.pi)
        super().__init__(size, domain, log=log, safe_eval=safe_eval)

    def _product_coefficients(self):
        """Product-to-sum: column 2k-1 = cos(k t), column 2k = sin(k t), column 0 = 1.
        ``cos a cos b = (cos(a-b) + cos(a+b)) / 2``, ``sin a sin b = (cos(a-b) - cos(a+b)) / 2``,
        ``sin a cos b = (sin(a+b) + sin(a-b)) / 2``."""
        r = self.size
        k_max = r // 2                                   # highest harmonic present (cos only when r is even)
        # the extended basis must hold cos / sin of 2 k_max, except that sin(2 k_max) never appears with r even
        ext = 4 * k_max + 1 if r % 2 == 1 else 4 * k_max
        ext = max(ext, r)
        c = np.zeros((r, r, ext))

        def col(kind, k):                                # kind 0: cos, 1: sin
            if k == 0:
                return (0, 1.0) if kind == 0 else (None, 0.0)
            if k < 0:
                idx, sign = col(kind, -k)
                return idx, (sign if kind == 0 else -sign)
            return (2 * k - 1 if kind == 0 else 2 * k), 1.0

        def harmonic(i):
            return (0, 0) if i == 0 else ((i + 1) // 2, 0 if i % 2 == 1 else 1)

        for i in range(r):
            ki, ti = harmonic(i)
            for j in range(r):
                kj, tj = harmonic(j)
                if ti == 0 and tj == 0:
                    terms = [(0, ki - kj, 0.5), (0, ki + kj, 0.5)]
                elif ti == 1 and tj == 1:
                    terms = [(0, ki - kj, 0.5), (0, ki + kj, -0.5)]
                elif ti == 1 and tj == 0:
                    terms = [(1, ki + kj, 0.5), (1, ki - kj, 0.5)]
                else:
                    terms = [(1, ki + kj, 0.5), (1, kj - ki, 0.5)]
                for kind, k, w in terms:
                    idx, sign = col(kind, k)
                    if idx is not None:
                        c[i, j, idx] += w * sign
        return ext, c


class Legendre(Moments):
    """Legendre polynomials on ref_domain (-1, 1) (``mlmc/moments.py:174-229``)."""
    _kind = _native.LEGENDRE

    def __init__(self, size, domain, ref_domain=None, log=False, safe_eval=True):
        self.ref_domain = ref_domain if ref_domain is not None else (-1, 1)
        # derivative in the Legendre basis: P_n' = sum_{k < n, n-k odd} (2k+1) P_k   (moments.py:185-188)
        self.diff_mat = np.zeros((size, size))
        for k in range(size - 1):
            self.diff_mat[k, k + 1::2] = 2 * k + 1
        self.diff2_mat = self.diff_mat @ self.diff_mat
        super().__init__(size, domain, log, safe_eval)

    def _product_coefficients(self):
        """Adams' formula (1878): ``P_m P_n = sum_{r <= min(m, n)} A_r A_{m-r} A_{n-r} / A_{m+n-r}
        (2m + 2n - 4r + 1) / (2m + 2n - 2r + 1) P_{m+n-2r}``, ``A_r = (2r - 1)!! / r! = binom(2r, r) / 2^r``
        (the powers of two cancel).  All coefficients are positive and each (m, n) row sums to 1, so the combination
        is a convex one.  Evaluated in exact rational arithmetic, rounded once."""
        from fractions import Fraction
        from math import comb
        r_size = self.size
        ext = 2 * r_size - 1
        a = [comb(2 * k, k) for k in range(ext + 1)]
        c = np.zeros((r_size, r_size, ext))
        for m in range(r_size):
            for n in range(m, r_size):
                for r in range(m + 1):
                    val = Fraction(a[r] * a[m - r] * a[n - r] * (2 * m + 2 * n - 4 * r + 1),
                                   a[m + n - r] * (2 * m + 2 * n - 2 * r + 1))
                    c[m, n, m + n - 2 * r] = c[n, m, m + n - 2 * r] = float(val)
        return ext, c

    def _apply_matrix(self, value, size, mat):
        table = self._eval_all(value, size)
        if isinstance(table, torch.Tensor):
            return table @ torch.from_numpy(mat[:size, :size]).to(table.device)
        return table @ mat[:size, :size]

    def _eval_diff(self, value, size):
        return self._apply_matrix(value, size, self.diff_mat)

    def _eval_diff2(self, value, size):
        return self._apply_matrix(value, size, self.diff2_mat)

    def _eval_all_der(self, value, size, degree=1):
        mat = np.linalg.matrix_power(self.diff_mat[:size, :size], degree)
        return self._apply_matrix(value, size, mat)


class TransformedMoments(Moments):
    """``new_moments = matrix . old_moments`` (``mlmc/moments.py:232-274``); row 0 of ``matrix`` is expected to
    be (1, 0, ...) so that the first new moment is still the constant 1."""

    def __init__(self, other_moments, matrix):
        matrix = np.asarray(matrix, dtype=np.float64)
        n, m = matrix.shape
        assert m == other_moments.size
        self.size = n
        self.domain = other_moments.domain
        self._origin = other_moments
        self._transform = matrix
        self._matrix_dev = None

    @property
    def ref_domain(self):
        return self._origin.ref_domain

    @property
    def _is_log(self):
        return self._origin._is_log

    @property
    def _is_clip(self):
        return self._origin._is_clip

    def __eq__(self, other):
        return type(self) is type(other) \
            and self.size == other.size \
            and self._origin == other._origin \
            and np.all(self._transform == other._transform)

    __hash__ = object.__hash__

    def change_size(self, size):
        return TransformedMoments(self._origin, self._transform[:size])

    def transform(self, value):
        return self._origin.transform(value)

    def inv_transform(self, ref):
        return self._origin.inv_transform(ref)

    def base_moments(self):
        return self._origin.base_moments()

    def transform_matrix(self):
        inner = self._origin.transform_matrix()
        return self._transform if inner is None else self._transform @ inner

    def basis_struct(self, size=None):
        return self.base_moments().basis_struct()

    def product_table(self):
        """``(L phi)_i (L phi)_j = sum_ab L_ia L_jb phi_a phi_b``: the base family's table with ``L`` applied on both
        sides (same extended base moments object)."""
        hit = getattr(self, "_product_table", None)
        if hit is None:
            base = self.base_moments()
            base_table = base.product_table()
            if base_table is None:
                return None
            ext_fn, c_t = base_table
            r0, l_mat = base.size, self.transform_matrix()
            c = c_t.T.reshape(r0, r0, -1)
            c = np.einsum("ia,abk->ibk", l_mat, c)
            c = np.einsum("jb,ibk->ijk", l_mat, c)
            hit = (ext_fn, np.ascontiguousarray(c.reshape(self.size * self.size, -1).T))
            self._product_table = hit
        return hit

    def _matrix_on(self, device):
        if self._matrix_dev is None or self._matrix_dev.device != device:
            self._matrix_dev = torch.from_numpy(np.ascontiguousarray(self.transform_matrix())).to(device)
        return self._matrix_dev

    def _eval_device(self, value, size, out=None):
        flat = value.reshape(-1)
        out = _native.basis_eval(self.basis_struct(), flat, size, self._matrix_on(value.device), out=out)
        return out.reshape(tuple(value.shape) + (size,))

    def _derived(self, fn_name, value, size, **kw):
        base = getattr(self._origin, fn_name)(value, self._origin.size, **kw)
        mat = self._transform.T
        if isinstance(base, torch.Tensor):
            mat = torch.from_numpy(np.ascontiguousarray(mat)).to(base.device)
        return (base @ mat)[..., :size]

    def _eval_all_der(self, value, size, degree=1):
        return self._derived("eval_all_der", value, size, degree=degree)

    def _eval_diff(self, value, size):
        return self._derived("eval_diff", value, size)

    def _eval_diff2(self, value, size):
        return self._derived("eval_diff2", value, size)
