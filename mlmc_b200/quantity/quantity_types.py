"""Quantity types -- mirror of ``mlmc/quantity/quantity_types.py:9-246``.

A QType describes how the flat component axis ``M`` of a chunk ``[M, n, 2]`` is structured
(dict of quantities -> time series -> field of locations -> array).  Chunks here are CUDA tensors; all type
operations are views / reshapes, no data moves to the host.
"""
import copy
from typing import List, Tuple

import numpy as np
import torch


def flatten_leading(chunk):
    """Always give a chunk the shape ``[M, n, S]`` (``QType.keep_dims``, quantity_types.py:35-49)."""
    if chunk.dim() == 2:
        return chunk.unsqueeze(0)
    if chunk.dim() > 2:
        return chunk.reshape((-1,) + tuple(chunk.shape[-2:]))
    raise ValueError("Chunk shape not supported")


class QType:
    def __init__(self, qtype):
        self._qtype = qtype

    def size(self) -> int:
        raise NotImplementedError

    def base_qtype(self):
        return self._qtype.base_qtype()

    def replace_scalar(self, substitute_qtype):
        """Copy of this type with the innermost ScalarType replaced (quantity_types.py:21-32).  The copy is shallow:
        only the chain of ``_qtype`` links is rebuilt, the (never mutated) shape / time / key lists are shared."""
        new_qtype = copy.copy(self)
        new_qtype._qtype = self._qtype.replace_scalar(substitute_qtype)
        return new_qtype

    keep_dims = staticmethod(flatten_leading)

    def _make_getitem_op(self, chunk, key):
        return flatten_leading(chunk[key])

    def reshape(self, data):
        return data


class ScalarType(QType):
    def __init__(self, qtype=float):
        self._qtype = qtype

    def base_qtype(self):
        if isinstance(self._qtype, BoolType):
            return self._qtype.base_qtype()
        return self

    def size(self) -> int:
        return self._qtype.size() if hasattr(self._qtype, "size") else 1

    def replace_scalar(self, substitute_qtype):
        return substitute_qtype


class BoolType(ScalarType):
    pass


class ArrayType(QType):
    def __init__(self, shape, qtype: QType):
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        self._shape = tuple(int(s) for s in shape)
        self._qtype = qtype

    def size(self) -> int:
        return int(np.prod(self._shape)) * self._qtype.size()

    def get_key(self, key):
        """Indexing by int / tuple / slices (quantity_types.py:106-125): one selected item is the inner type."""
        new_shape = np.empty(self._shape)[key].shape
        if len(new_shape) == 1 and new_shape[0] == 1:
            new_shape = ()
        q_type = ArrayType(new_shape, qtype=self._qtype) if len(new_shape) > 0 else self._qtype
        return q_type, 0

    def _make_getitem_op(self, chunk, key):
        chunk = chunk.reshape(self._shape + (-1,) + tuple(chunk.shape[-2:])) if self._qtype.size() > 1 \
            else chunk.reshape(self._shape + tuple(chunk.shape[-2:]))
        return flatten_leading(chunk[key])

    def reshape(self, data):
        if isinstance(self._qtype, ScalarType):
            return data.reshape(self._shape)
        return data.reshape(self._shape + (int(np.prod(data.shape)) // int(np.prod(self._shape)),))


class TimeSeriesType(QType):
    def __init__(self, times, qtype):
        if isinstance(times, np.ndarray):
            times = times.tolist()
        self._times = list(times)
        self._qtype = qtype

    def size(self) -> int:
        return len(self._times) * self._qtype.size()

    def get_key(self, key):
        position = self._times.index(key)
        return self._qtype, position * self._qtype.size()

    @staticmethod
    def time_interpolation(quantity, value):
        """Linear interpolation in time (quantity_types.py:161-174, scipy ``interp1d`` default kind)."""
        from . import quantity as q_mod
        times = np.asarray(quantity.qtype._times, dtype=float)
        inner = quantity.qtype._qtype
        hi = int(np.searchsorted(times, value, side="right"))
        hi = min(max(hi, 1), len(times) - 1)
        lo = hi - 1
        if not (times[0] <= value <= times[-1]):
            raise ValueError("A value in x_new is outside the interpolation range.")
        w = (value - times[lo]) / (times[hi] - times[lo])
        n_in = inner.size()

        def interp(y):
            a = y[lo * n_in:(lo + 1) * n_in]
            b = y[hi * n_in:(hi + 1) * n_in]
            return a + (b - a) * w
        return q_mod.Quantity(quantity_type=inner, input_quantities=[quantity], operation=interp)


class FieldType(QType):
    def __init__(self, args: List[Tuple[str, QType]]):
        self._dict = dict(args)
        self._qtype = args[0][1]
        assert all(q_type.size() == self._qtype.size() for _, q_type in args)

    def size(self) -> int:
        return len(self._dict) * self._qtype.size()

    def replace_scalar(self, substitute_qtype):
        """All locations share one inner type: replace it once and share the result (a deep copy of a field with
        1e4 locations would cost more than the estimate itself)."""
        inner = self._qtype.replace_scalar(substitute_qtype)
        new = FieldType.__new__(FieldType)
        new._dict = {key: inner for key in self._dict}
        new._qtype = inner
        return new

    def get_key(self, key):
        position = list(self._dict.keys()).index(key)
        return self._qtype, position * self._qtype.size()


class DictType(QType):
    def __init__(self, args: List[Tuple[str, QType]]):
        self._dict = dict(args)
        self._check_base_type()

    def _check_base_type(self):
        qtypes = list(self._dict.values())
        first = qtypes[0].base_qtype()
        for qtype in qtypes[1:]:
            if not isinstance(qtype.base_qtype(), type(first)):
                raise TypeError("qtype {} has base QType {}, expecting {}. All QTypes must have same base QType, "
                                "either ScalarType or BoolType".format(qtype, qtype.base_qtype(), first))

    def base_qtype(self):
        return next(iter(self._dict.values())).base_qtype()

    def size(self) -> int:
        return int(sum(q_type.size() for q_type in self._dict.values()))

    def get_qtypes(self):
        return self._dict.values()

    def replace_scalar(self, substitute_qtype):
        return DictType([(key, qtype.replace_scalar(substitute_qtype)) for key, qtype in self._dict.items()])

    def get_key(self, key):
        q_type = self._dict[key]
        start = 0
        for k, qt in self._dict.items():
            if k == key:
                break
            start += qt.size()
        return q_type, start
