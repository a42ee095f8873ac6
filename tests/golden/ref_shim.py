"""Import shim that lets the UNMODIFIED reference (``/root/reference``) run in this container.

Test infrastructure only.  It is used by ``tests/golden/make_golden.py`` to generate the golden
fixtures committed next to it; nothing in the product, the ``-m gpu`` tests, ``smoke()`` or
``bench.py`` imports it (``/root/reference`` does not exist on the GPU box).

What is missing here and how it is routed around:

* numpy >= 1.24 dropped ``np.float`` / ``np.VisibleDeprecationWarning`` that the reference still
  names (``mlmc/sample_storage.py:174``, ``mlmc/sample_storage_hdf.py:8``).
* ``h5py``, ``memoization``, ``ruamel.yaml``, ``matplotlib``, ``gstools`` are not installed; the
  reference imports them at module import time (``mlmc/tool/hdf5.py:2``, ``mlmc/quantity/quantity.py:4``,
  ``mlmc/tool/pbs_job.py:6``, ``mlmc/plot/plots.py:4-8``).  They are replaced by inert stubs; the only
  stub with behaviour is ``memoization.cached`` (a dict cache with ``cache_clear``).
* sample storage: an array-backed ``SampleStorage`` subclass serving ``float64[N, 2, M]`` rows per
  level with the ``sample_pairs_level`` contract of ``mlmc/sample_storage.py:261-285``.
"""
import sys
import types
import itertools
import numpy as np

REFERENCE_ROOT = "/root/reference"


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


class _Anything:
    """Object that swallows any attribute access / call (matplotlib stand-in)."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, item):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __getitem__(self, item):
        return _Anything()

    def __setitem__(self, key, value):
        pass

    def __iter__(self):
        return iter(())


def _cached(custom_key_maker=None, **_kw):
    def deco(fn):
        cache = {}

        def wrapper(*args, **kwargs):
            key = custom_key_maker(*args, **kwargs) if custom_key_maker else (args, tuple(kwargs.items()))
            if key not in cache:
                cache[key] = fn(*args, **kwargs)
            return cache[key]

        wrapper.cache_clear = cache.clear
        wrapper.__wrapped__ = fn
        return wrapper
    return deco


def install():
    """Install the stubs and put the reference on sys.path.  Idempotent."""
    if getattr(install, "_done", False):
        return
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "VisibleDeprecationWarning"):
        np.VisibleDeprecationWarning = np.exceptions.VisibleDeprecationWarning
    _stub("memoization", cached=_cached)
    _stub("h5py", File=_Anything, Group=_Anything, Dataset=_Anything)
    ruamel = _stub("ruamel")
    yaml = _stub("ruamel.yaml", YAML=_Anything, load=_Anything(), dump=_Anything())
    err = _stub("ruamel.yaml.error", ReusedAnchorWarning=Warning, UnsafeLoaderWarning=Warning)
    ruamel.yaml = yaml
    yaml.error = err
    mpl = _stub("matplotlib", rcParams={}, use=lambda *a, **k: None, cm=_Anything(), colors=_Anything())
    for sub in ("pyplot", "patches", "ticker", "cm", "colors", "lines", "gridspec"):
        m = _stub("matplotlib." + sub)
        m.__getattr__ = lambda name: _Anything()
        setattr(mpl, sub, m)
    for name in ("gstools", "seaborn", "statprof"):
        m = _stub(name)
        m.__getattr__ = lambda name: _Anything()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    install._done = True


def array_storage(level_rows, level_parameters=None, n_ops=None, chunk_rows=65536):
    """Array-backed reference ``SampleStorage``.

    :param level_rows: list over levels of float64[N_l, 2, M] (storage row order: sample, fine/coarse, component)
    """
    install()
    from mlmc.sample_storage import SampleStorage
    from mlmc.quantity.quantity_spec import ChunkSpec

    class ArrayStorage(SampleStorage):
        def __init__(self):
            self._rows = [np.ascontiguousarray(r, dtype=np.float64) for r in level_rows]
            self._level_parameters = level_parameters
            self._n_ops = n_ops
            self._format = []

        # -- read side (the feed of the hot path) --
        def _level_chunks(self, level_id, n_samples=None):
            n = len(self._rows[level_id][:n_samples])
            if n_samples is not None:       # both reference storages yield ONE chunk then (hdf5.py:359-360)
                yield ChunkSpec(chunk_id=0, chunk_slice=slice(0, n, 1), level_id=level_id)
                return
            for cid, start in enumerate(range(0, max(n, 1), chunk_rows)):
                yield ChunkSpec(chunk_id=cid, chunk_slice=slice(start, min(start + chunk_rows, n), 1),
                                level_id=level_id)

        def sample_pairs_level(self, chunk_spec):
            level_id = 0 if chunk_spec.level_id is None else int(chunk_spec.level_id)
            chunk = self._rows[level_id]
            if chunk_spec.chunk_slice is not None:
                chunk = chunk[chunk_spec.chunk_slice]
            if level_id == 0:
                chunk = chunk[:, :1, :]
            return chunk.transpose((2, 0, 1))

        def sample_pairs(self):
            return [self.sample_pairs_level(ChunkSpec(level_id=l)) for l in self.get_level_ids()]

        def get_level_ids(self):
            return list(range(len(self._rows)))

        def get_n_levels(self):
            return len(self._rows)

        def get_n_collected(self):
            return [len(r) for r in self._rows]

        def get_level_parameters(self):
            return self._level_parameters

        def get_n_ops(self):
            return self._n_ops

        def n_finished(self):
            return self.get_n_collected()

        def unfinished_ids(self):
            return []

        # -- write side: unused --
        def save_samples(self, successful_samples, failed_samples):
            raise NotImplementedError

        def save_result_format(self, res_spec):
            self._format = res_spec

        def load_result_format(self):
            return self._format

        def save_global_data(self, result_format, level_parameters=None):
            self._format = result_format
            self._level_parameters = level_parameters

        def save_scheduled_samples(self, level_id, samples):
            pass

        def load_scheduled_samples(self):
            return {}

        def save_n_ops(self, n_ops):
            pass

    return ArrayStorage()
