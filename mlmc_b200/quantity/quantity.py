"""Lazy quantity DAG on device tensors -- mirror of ``mlmc/quantity/quantity.py:14-695``.

Same classes and methods as the reference (``Quantity``, ``QuantityConst``, ``QuantityMean``, ``QuantityStorage``,
``make_root_quantity``; indexing, arithmetic, comparisons -> masks, ``select``, ``subsample``, numpy-ufunc
overloading, ``QArray/QDict/QTimeSeries/QField``).  The difference is where a chunk lives: every operation maps
CUDA tensors ``[M, n, 2]`` to CUDA tensors (views for indexing, elementwise device ops for arithmetic), so a
derived quantity reaches the fused estimation kernels without a host round trip.  ``samples(chunk_spec)`` keeps
the reference contract and returns a NumPy array.

The chunk contract of the reference is unchanged (``quantity.py:686-692``): ``QuantityStorage`` serves the
transposed view ``[M, n, 2]`` of storage rows ``[n, 2, M]`` (level 0: ``[M, n, 1]``).
"""
import operator
from typing import List

import numpy as np
import scipy.stats
import torch

from ..sample_storage import SampleStorage
from .quantity_spec import QuantitySpec, ChunkSpec
from . import quantity_types as qt

RNG = np.random.default_rng()

_UFUNC_TO_TORCH = {
    "absolute": "abs", "fabs": "abs", "power": "pow", "arcsin": "asin", "arccos": "acos", "arctan": "atan",
    "arcsinh": "asinh", "arccosh": "acosh", "arctanh": "atanh", "true_divide": "true_divide",
    "divide": "true_divide", "multiply": "mul", "subtract": "sub", "negative": "neg", "mod": "remainder",
    "remainder": "remainder", "greater": "gt", "greater_equal": "ge", "less": "lt", "less_equal": "le",
    "equal": "eq", "not_equal": "ne", "arctan2": "atan2", "invert": "bitwise_not",
}


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("mlmc_b200 needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class DeviceChunk:
    """Rows ``[n, 2, M]`` of one level resident in HBM plus the identifiers the DAG needs."""
    __slots__ = ("level_id", "rows", "chunk_id")

    def __init__(self, level_id, rows, chunk_id=0):
        self.level_id = level_id
        self.rows = rows
        self.chunk_id = chunk_id


def make_root_quantity(storage: SampleStorage, q_specs: List[QuantitySpec]):
    """Root quantity over a storage; its type mirrors the result format (quantity.py:14-32):
    dict of quantities -> time series -> field of locations -> array of ``shape``."""
    entries = []
    for spec in q_specs:
        array_type = qt.ArrayType(spec.shape, qt.ScalarType(float))
        field_type = qt.FieldType([(loc, array_type) for loc in spec.locations])
        entries.append((spec.name, qt.TimeSeriesType(spec.times, field_type)))
    return QuantityStorage(storage, qt.DictType(entries))


class Quantity:
    def __init__(self, quantity_type, operation, input_quantities=[]):
        self.qtype = quantity_type
        self._operation = operation
        self._input_quantities = input_quantities
        self._storage = self.get_quantity_storage()
        self._selection_id = self.set_selection_id()
        self._check_selection_ids()

    # ---- graph bookkeeping (quantity.py:55-115) ----
    def get_quantity_storage(self):
        for in_quantity in self._input_quantities:
            storage = in_quantity.get_quantity_storage()
            if storage is not None:
                self._storage = storage
                return storage
        return None

    def set_selection_id(self):
        selection_id = None
        for in_quantity in self._input_quantities:
            if selection_id is None:
                selection_id = in_quantity.selection_id()
            elif in_quantity.selection_id() is not None and selection_id != in_quantity.selection_id():
                raise Exception("Different selection IDs among input quantities")
        return selection_id

    def _check_selection_ids(self):
        if self._storage is None:
            return
        for in_quantity in self._input_quantities:
            sel_id = in_quantity.selection_id()
            if sel_id is not None and sel_id != self.selection_id():
                raise AssertionError("Not all input quantities come from the same quantity storage")

    def selection_id(self):
        if self._selection_id is not None:
            return self._selection_id
        if self._storage is None:
            self._storage = self.get_quantity_storage()
        return id(self._storage)

    def size(self) -> int:
        return self.qtype.size()

    # ---- evaluation ----
    def device_samples(self, chunk: DeviceChunk):
        """CUDA tensor ``[M, n, 2]`` of this quantity on one device chunk."""
        inputs = [q.device_samples(chunk) for q in self._input_quantities]
        return self._operation(*inputs)

    def samples(self, chunk_spec):
        """Reference contract (quantity.py:126-135): NumPy ``[M, chunk size, 2]`` for a ChunkSpec."""
        storage_q = self.get_quantity_storage()
        if storage_q is None:
            return self.device_samples(DeviceChunk(chunk_spec.level_id, None)).cpu().numpy()
        chunk = storage_q.device_chunk(chunk_spec)
        return self.device_samples(chunk).cpu().numpy()

    # ---- selection / sub-sampling (quantity.py:149-364) ----
    def select(self, *args):
        masks = args[0]
        for quantity in args:
            if not isinstance(quantity.qtype.base_qtype(), qt.BoolType):
                raise Exception("Quantity: {} doesn't have BoolType, instead it has QType: {}"
                                .format(quantity, quantity.qtype.base_qtype()))
        for m in args[1:]:
            masks = np.logical_and(masks, m)

        def op(x, mask):
            return x[..., mask, :]
        q = Quantity(quantity_type=self.qtype, input_quantities=[self, masks], operation=op)
        q._selection_id = id(q)
        return q

    @staticmethod
    def pick_samples(chunk, subsample_params):
        """Hypergeometric chunk-wise sub-sampling (quantity.py:305-322; method S of Vitter's paper)."""
        size = int(scipy.stats.hypergeom(subsample_params.n, subsample_params.k, chunk.shape[1]).rvs(size=1)[0])
        idx = torch.from_numpy(RNG.integers(0, chunk.shape[1], size=size)).to(chunk.device)
        out = chunk.index_select(1, idx)
        subsample_params.k -= out.shape[1]
        subsample_params.n -= chunk.shape[1]
        return out

    def subsample(self, sample_vec):
        class SubsampleParams:
            def __init__(self, num_subsample, num_collected):
                self._orig_k = num_subsample
                self._orig_n = num_collected
                self.k = num_subsample
                self.n = num_collected

        params = {level: SubsampleParams(sample_vec[level], n)
                  for level, n in enumerate(self.get_quantity_storage().n_collected())}
        holder = QuantityConst(quantity_type=qt.ScalarType(), value=0.0)

        def per_level(_value, level_id):
            p = params[level_id]
            p.k, p.n = p._orig_k, p._orig_n
            return p
        holder._adjust_value = per_level
        return Quantity(quantity_type=self.qtype.replace_scalar(qt.BoolType()),
                        input_quantities=[self, holder], operation=Quantity.pick_samples)

    # ---- indexing ----
    def __getitem__(self, key):
        new_qtype, start = self.qtype.get_key(key)
        if not isinstance(self.qtype, qt.ArrayType):
            key = slice(start, start + new_qtype.size())

        def getitem_op(y):
            return self.qtype._make_getitem_op(y, key=key)
        return Quantity(quantity_type=new_qtype, input_quantities=[self], operation=getitem_op)

    def __getattr__(self, name):
        if name.startswith("__") or name in ("qtype", "_storage", "_input_quantities", "_operation",
                                             "_selection_id"):
            raise AttributeError(name)
        static_fun = getattr(self.qtype, name)      # only static QType functions are forwarded

        def apply_on_quantity(*attr, **d_attr):
            return static_fun(self, *attr, **d_attr)
        return apply_on_quantity

    # ---- arithmetic ----
    @staticmethod
    def create_quantity(quantities, operation):
        for quantity in quantities:
            if not isinstance(quantity, QuantityConst):
                return Quantity(quantity.qtype, operation=operation, input_quantities=quantities)
        return QuantityConst(quantities[0].qtype, value=operation(*[q._value for q in quantities]))

    _reduction_op = create_quantity

    add_op = staticmethod(operator.add)
    sub_op = staticmethod(operator.sub)
    mult_op = staticmethod(operator.mul)
    truediv_op = staticmethod(operator.truediv)
    mod_op = staticmethod(operator.mod)

    def __add__(self, other):
        return Quantity.create_quantity([self, Quantity.wrap(other)], Quantity.add_op)

    def __sub__(self, other):
        return Quantity.create_quantity([self, Quantity.wrap(other)], Quantity.sub_op)

    def __mul__(self, other):
        return Quantity.create_quantity([self, Quantity.wrap(other)], Quantity.mult_op)

    def __truediv__(self, other):
        return Quantity.create_quantity([self, Quantity.wrap(other)], Quantity.truediv_op)

    def __mod__(self, other):
        return Quantity.create_quantity([self, Quantity.wrap(other)], Quantity.mod_op)

    def __radd__(self, other):
        return Quantity.create_quantity([Quantity.wrap(other), self], Quantity.add_op)

    def __rsub__(self, other):
        return Quantity.create_quantity([Quantity.wrap(other), self], Quantity.sub_op)

    def __rmul__(self, other):
        return Quantity.create_quantity([Quantity.wrap(other), self], Quantity.mult_op)

    def __rtruediv__(self, other):
        return Quantity.create_quantity([Quantity.wrap(other), self], Quantity.truediv_op)

    def __rmod__(self, other):
        return Quantity.create_quantity([Quantity.wrap(other), self], Quantity.mod_op)

    # conveniences the reference leaves to ``np.negative(q)`` / ``np.abs(q)`` / ``np.power(q, p)``
    def __neg__(self):
        return Quantity._method(np.negative, "__call__", self)

    def __abs__(self):
        return Quantity._method(np.absolute, "__call__", self)

    def __pow__(self, other):
        return Quantity._method(np.power, "__call__", self, other)

    # ---- comparisons -> sample masks (quantity.py:245-303) ----
    @staticmethod
    def _process_mask(x, y, op):
        """A sample passes only if ALL its entries (components, fine and coarse) meet the condition."""
        mask = op(x, y)
        return mask.all(dim=0).all(dim=1) if mask.dim() == 3 else mask.reshape(-1, *mask.shape[-2:]).all(0).all(1)

    def _mask_quantity(self, other, op):
        new_qtype = self.qtype.replace_scalar(qt.BoolType())
        other = Quantity.wrap(other)
        if not isinstance(self.qtype.base_qtype(), qt.ScalarType) or \
                not isinstance(other.qtype.base_qtype(), qt.ScalarType):
            raise TypeError("Quantity has base qtype {}. Quantities with base qtype ScalarType are the only ones "
                            "that support comparison".format(self.qtype.base_qtype()))
        return Quantity(quantity_type=new_qtype, input_quantities=[self, other], operation=op)

    def __lt__(self, other):
        return self._mask_quantity(other, lambda x, y: Quantity._process_mask(x, y, operator.lt))

    def __le__(self, other):
        return self._mask_quantity(other, lambda x, y: Quantity._process_mask(x, y, operator.le))

    def __gt__(self, other):
        return self._mask_quantity(other, lambda x, y: Quantity._process_mask(x, y, operator.gt))

    def __ge__(self, other):
        return self._mask_quantity(other, lambda x, y: Quantity._process_mask(x, y, operator.ge))

    def __eq__(self, other):
        return self._mask_quantity(other, lambda x, y: Quantity._process_mask(x, y, operator.eq))

    def __ne__(self, other):
        return self._mask_quantity(other, lambda x, y: Quantity._process_mask(x, y, operator.ne))

    __hash__ = object.__hash__

    # ---- numpy ufuncs (quantity.py:173-175, 418-438): evaluated with the torch function of the same name ----
    def __array_ufunc__(self, ufunc, method, *args, **kwargs):
        return Quantity._method(ufunc, method, *args, **kwargs)

    @staticmethod
    def _method(ufunc, method, *args, **kwargs):
        if method == "reduce":
            return Quantity._reduce(ufunc, *args, **kwargs)
        if method != "__call__" or any(v is not None and v != (None,) for v in kwargs.values()):
            raise NotImplementedError("only plain ufunc calls and reductions over the components are supported")
        name = _UFUNC_TO_TORCH.get(ufunc.__name__, ufunc.__name__)
        torch_fn = getattr(torch, name, None)
        if torch_fn is None:
            raise NotImplementedError("numpy ufunc %s has no device implementation" % ufunc.__name__)

        def ufunc_call(*chunks):
            return torch_fn(*chunks)
        quantities = [Quantity.wrap(arg) for arg in args]
        result_qtype = Quantity._result_qtype(ufunc_call, quantities)
        return Quantity(quantity_type=result_qtype, input_quantities=list(quantities), operation=ufunc_call)

    @staticmethod
    def _reduce(ufunc, quantity, axis=0, keepdims=False, **kwargs):
        """``np.max / np.min / np.sum / np.prod / np.all / np.any`` of a quantity over its COMPONENT axis
        (``np.max(q, axis=0, keepdims=True)`` in test/test_quantity_concept.py:382-392) -> one-component quantity."""
        if any(v is not None and v != (None,) for v in kwargs.values()):
            raise NotImplementedError("reduction options %s are not supported on quantities" % sorted(kwargs))
        if axis not in (0, (0,)):
            raise NotImplementedError("quantities reduce over axis 0 (the components) only")
        name = {"maximum": "amax", "minimum": "amin", "add": "sum", "multiply": "prod", "logical_and": "all",
                "logical_or": "any"}.get(ufunc.__name__)
        if name is None:
            raise NotImplementedError("numpy reduction of %s has no device implementation" % ufunc.__name__)
        torch_fn = getattr(torch, name)

        def reduce_call(chunk):
            return torch_fn(chunk, dim=0, keepdim=True)          # a chunk always keeps its [M, n, S] shape
        quantity = Quantity.wrap(quantity)
        result_qtype = Quantity._result_qtype(reduce_call, [quantity])
        return Quantity(quantity_type=result_qtype, input_quantities=[quantity], operation=reduce_call)

    @staticmethod
    def wrap(value):
        if isinstance(value, Quantity):
            return value
        if isinstance(value, bool):
            return QuantityConst(quantity_type=qt.BoolType(), value=value)
        if isinstance(value, (int, float, np.integer, np.floating)):
            return QuantityConst(quantity_type=qt.ScalarType(), value=float(value))
        if isinstance(value, (list, np.ndarray)):
            value = np.array(value)
            return QuantityConst(quantity_type=qt.ArrayType(shape=value.shape, qtype=qt.ScalarType()), value=value)
        raise ValueError("Values {} are not flat, bool or array (list)".format(value))

    @staticmethod
    def _get_base_qtype(args_quantities):
        for quantity in args_quantities:
            if isinstance(quantity, Quantity) and type(quantity.qtype.base_qtype()) == qt.ScalarType:
                return qt.ScalarType()
        return qt.BoolType()

    @staticmethod
    def _result_qtype(method, quantities):
        """Result type from an evaluation on the first few samples (quantity.py:463-482)."""
        probe = None
        for q in quantities:
            storage_q = q.get_quantity_storage()
            if storage_q is not None:
                probe = storage_q.device_chunk(next(storage_q.chunks()), max_rows=8)
                break
        if probe is None:
            probe = DeviceChunk(None, None)
        try:
            result = method(*[q.device_samples(probe) for q in quantities])
        except RuntimeError as exc:              # torch's broadcasting error; numpy raises ValueError here
            raise ValueError(str(exc))
        return qt.ArrayType(shape=result.shape[0], qtype=Quantity._get_base_qtype(quantities))

    # ---- composition (quantity.py:396-523) ----
    @staticmethod
    def _concatenate(quantities, qtype, axis=0):
        def op_concatenate(*chunks):
            return torch.cat(tuple(chunks), dim=axis)
        return Quantity(qtype, input_quantities=[*quantities], operation=op_concatenate)

    @staticmethod
    def _check_same_qtype(quantities):
        qtype = quantities[0].qtype
        for quantity in quantities[1:]:
            if qtype != quantity.qtype:
                raise ValueError("Quantities don't have same QType")
        return qtype

    @staticmethod
    def QArray(quantities):
        arr = np.empty(np.shape(quantities), dtype=object)
        flat = []

        def collect(item):
            if isinstance(item, (list, tuple)):
                for sub in item:
                    collect(sub)
            else:
                flat.append(item)
        collect(quantities)
        qtype = Quantity._check_same_qtype(flat)
        return Quantity._concatenate(flat, qtype=qt.ArrayType(arr.shape, qtype))

    @staticmethod
    def QDict(key_quantity):
        dict_type = qt.DictType([(key, quantity.qtype) for key, quantity in key_quantity])
        return Quantity._concatenate([q for _, q in key_quantity], qtype=dict_type)

    @staticmethod
    def QTimeSeries(time_quantity):
        qtype = Quantity._check_same_qtype([q for _, q in time_quantity])
        times = [t for t, _ in time_quantity]
        return Quantity._concatenate([q for _, q in time_quantity], qtype=qt.TimeSeriesType(times=times, qtype=qtype))

    @staticmethod
    def QField(key_quantity):
        Quantity._check_same_qtype([q for _, q in key_quantity])
        field_type = qt.FieldType([(key, quantity.qtype) for key, quantity in key_quantity])
        return Quantity._concatenate([q for _, q in key_quantity], qtype=field_type)


class QuantityConst(Quantity):
    """Constant quantity, value broadcast as ``[M, 1, 1]`` (quantity.py:526-581)."""

    def __init__(self, quantity_type, value):
        self.qtype = quantity_type
        self._value = self._process_value(value)
        self._input_quantities = []
        self._selection_id = None
        self._storage = None
        self._operation = None

    @staticmethod
    def _process_value(value):
        if isinstance(value, torch.Tensor):
            return value
        if isinstance(value, (int, float, bool, np.integer, np.floating)):
            value = np.array([value])
        value = np.asarray(value)
        dtype = torch.bool if value.dtype == bool else torch.float64
        return torch.as_tensor(value.reshape(-1), dtype=dtype, device=_device())[:, None, None]

    def get_quantity_storage(self):
        return None

    def selection_id(self):
        return self._selection_id

    def _adjust_value(self, value, level_id=None):
        return value

    def device_samples(self, chunk):
        return self._adjust_value(self._value, chunk.level_id)

    def samples(self, chunk_spec):
        out = self._adjust_value(self._value, chunk_spec.level_id)
        return out.cpu().numpy() if isinstance(out, torch.Tensor) else out


class QuantityMean:
    """Result of ``estimate_mean`` (quantity.py:568-651): per-level means / variances and their MLMC totals."""

    def __init__(self, quantity_type, l_means, l_vars, n_samples, n_rm_samples, mean=None, var=None):
        self.qtype = quantity_type
        self._mean = mean           # optional: totals already formed on the device (same level order / arithmetic)
        self._var = var
        self._l_means = np.asarray(l_means)
        self._l_vars = np.asarray(l_vars)
        self._n_samples = np.array(n_samples)
        self._n_rm_samples = np.array(n_rm_samples)

    def _calculate_mean_var(self):
        self._mean = np.sum(self._l_means, axis=0)
        self._var = np.sum(self._l_vars / self._n_samples[:, None], axis=0)

    @property
    def mean(self):
        if self._mean is None or self._var is None:
            self._calculate_mean_var()
        return self._reshape(self._mean)

    @property
    def var(self):
        if self._mean is None or self._var is None:
            self._calculate_mean_var()
        return self._reshape(self._var)

    @property
    def l_means(self):
        return np.array([self._reshape(means) for means in self._l_means])

    @property
    def l_vars(self):
        return np.array([self._reshape(variances) for variances in self._l_vars])

    @property
    def n_samples(self):
        return self._n_samples

    @property
    def n_rm_samples(self):
        return self._n_rm_samples

    def _reshape(self, data):
        return self.qtype.reshape(data)

    def __getitem__(self, key):
        new_qtype, start = self.qtype.get_key(key)
        if not isinstance(self.qtype, qt.ArrayType):
            key = slice(start, start + new_qtype.size())
        l_means = self.l_means[:, key]
        l_vars = self.l_vars[:, key]
        return QuantityMean(quantity_type=new_qtype, l_means=l_means.reshape((l_means.shape[0], -1)),
                            l_vars=l_vars.reshape((l_vars.shape[0], -1)), n_samples=self._n_samples,
                            n_rm_samples=self._n_rm_samples)


class QuantityStorage(Quantity):
    """The only quantity that touches the sample storage (quantity.py:654-695)."""

    def __init__(self, storage, qtype):
        self._storage = storage
        self.qtype = qtype
        self._input_quantities = []
        self._operation = None
        self._selection_id = None

    def level_ids(self):
        return self._storage.get_level_ids()

    def selection_id(self):
        return id(self)

    def get_quantity_storage(self):
        return self

    def chunks(self, level_id=None):
        return self._storage.chunks(level_id)

    def n_collected(self):
        return self._storage.get_n_collected()

    def device_chunk(self, chunk_spec, max_rows=None):
        """Upload the rows named by a ChunkSpec (used by ``samples`` and by type probing)."""
        level_id = 0 if chunk_spec.level_id is None else int(chunk_spec.level_id)
        rows = self._storage.level_rows(level_id)
        if chunk_spec.chunk_slice is not None:
            rows = rows[chunk_spec.chunk_slice]
        if max_rows is not None:
            rows = rows[:max_rows]
        dev_rows = torch.from_numpy(np.ascontiguousarray(rows)).to(_device())
        return DeviceChunk(level_id, dev_rows, chunk_spec.chunk_id or 0)

    def device_samples(self, chunk):
        x = chunk.rows.permute(2, 0, 1)              # [M, n, 2] view in storage strides
        return x[:, :, :1] if chunk.level_id == 0 else x

    def samples(self, chunk_spec):
        return self._storage.sample_pairs_level(chunk_spec)
