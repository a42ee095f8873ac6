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

    def _eval_device(self, value, size):
        """value: CUDA float64 tensor of any shape -> CUDA tensor ``value.shape + (size,)``."""
        flat = value.reshape(-1)
        out = _native.basis_eval(self.basis_struct(size), flat, size)
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


class Fourier(Moments):
    """``1, cos t, sin t, cos 2t, ...`` on ref_domain (0, 2 pi) (``mlmc/moments.py:133-162``)."""
    _kind = _native.FOURIER

    def __init__(self, size, domain=(0, 2 * np.pi), ref_domain=None, log=False, safe_eval=True):
        self.ref_domain = ref_domain if ref_domain is not None else (0, 2 * np.pi)
        super().__init__(size, domain, log=log, safe_eval=safe_eval)


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

    def _matrix_on(self, device):
        if self._matrix_dev is None or self._matrix_dev.device != device:
            self._matrix_dev = torch.from_numpy(np.ascontiguousarray(self.transform_matrix())).to(device)
        return self._matrix_dev

    def _eval_device(self, value, size):
        flat = value.reshape(-1)
        out = _native.basis_eval(self.basis_struct(), flat, size, self._matrix_on(value.device))
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
