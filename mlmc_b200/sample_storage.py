"""Sample storages feeding the estimation path -- B200 mirror of ``mlmc/sample_storage.py`` and the read side of
``mlmc/sample_storage_hdf.py`` / ``mlmc/tool/hdf5.py``.

Contract kept from the reference:
  * per level, collected samples are rows ``float64[N, 2, M]`` (sample, fine/coarse, component); level 0 stores
    an auxiliary zero coarse row (``mlmc/tool/hdf5.py:311-320``, ``mlmc/sample_storage.py:163-190``);
  * ``sample_pairs_level(chunk_spec)`` returns the transposed view ``[M, n, 2]`` (level 0: ``[M, n, 1]``)
    (``mlmc/sample_storage.py:261-285``, ``mlmc/sample_storage_hdf.py:169-184``);
  * ``chunks(level_id, n_samples)`` yields ``ChunkSpec`` objects; with ``n_samples`` given a level is ONE chunk
    (``mlmc/tool/hdf5.py:353-363``).

B200 side: ``device_chunks(level_id, device)`` yields CUDA tensors ``[n, 2, M]`` of a level.  Rows are kept in
PINNED host memory when CUDA is available so that they go to HBM with plain ``cudaMemcpyAsync`` on a copy stream,
double-buffered against the compute stream; a level that was uploaded once stays resident (``keep_resident``)
as long as it fits the HBM budget.
"""
import itertools
import os
from abc import ABCMeta, abstractmethod

import numpy as np
import torch

from .quantity.quantity_spec import ChunkSpec


def _pinned_empty(shape):
    """float64 host tensor, pinned if a CUDA runtime is usable."""
    if torch.cuda.is_available():
        try:
            return torch.empty(shape, dtype=torch.float64, pin_memory=True)
        except RuntimeError:
            pass
    return torch.empty(shape, dtype=torch.float64)


class SampleStorage(metaclass=ABCMeta):
    """Abstract storage (``mlmc/sample_storage.py:9-131``)."""

    #: rows per streamed device chunk (x 16*M bytes); whole level if it is already resident
    device_chunk_bytes = 32 << 20
    #: fraction of free HBM a resident copy of all levels may take
    resident_fraction = 0.6
    #: multi-GPU: True if this process's storage already holds only ITS shard of every level (then every local row
    #: is reduced and only the level sums are combined); False = every rank sees all rows and takes a row range
    rows_are_local_shard = False

    # ------------------------------------------------------------------ write side (reference API)
    @abstractmethod
    def save_samples(self, successful_samples, failed_samples):
        """Store results of finished samples."""

    @abstractmethod
    def save_result_format(self, res_spec):
        """Store the list of QuantitySpec."""

    @abstractmethod
    def load_result_format(self):
        """Return the list of QuantitySpec."""

    @abstractmethod
    def save_global_data(self, result_format, level_parameters=None):
        """Store result format and level parameters."""

    @abstractmethod
    def save_scheduled_samples(self, level_id, samples):
        """Remember ids of scheduled samples."""

    @abstractmethod
    def load_scheduled_samples(self):
        """Dict[level_id, List[sample_id]]"""

    @abstractmethod
    def save_n_ops(self, n_ops):
        """n_ops: iterable of (level_id, (time, n_samples))"""

    # ------------------------------------------------------------------ read side
    @abstractmethod
    def level_rows(self, level_id):
        """Host rows ``float64[N, 2, M]`` of one level (NumPy array or memmap); level 0 may come as ``[N, 1, M]``
        (fine row only)."""

    @abstractmethod
    def get_level_ids(self):
        pass

    @abstractmethod
    def get_level_parameters(self):
        pass

    @abstractmethod
    def get_n_ops(self):
        pass

    def get_n_levels(self):
        return len(self.get_level_ids())

    def get_n_collected(self):
        return [len(self.level_rows(l)) for l in self.get_level_ids()]

    def n_finished(self):
        return np.array(self.get_n_collected(), dtype=float)

    def unfinished_ids(self):
        return []

    def chunks(self, level_id=None, n_samples=None):
        assert isinstance(n_samples, (type(None), int)), "n_samples param must be int"
        level_ids = self.get_level_ids() if level_id is None else [level_id]
        return itertools.chain(*[self._level_chunks(l, n_samples) for l in level_ids])

    def _level_chunks(self, level_id, n_samples=None):
        n = len(self.level_rows(level_id))
        if n_samples is not None:
            n = min(n, n_samples)
        yield ChunkSpec(chunk_id=0, chunk_slice=slice(0, n, 1), level_id=level_id)

    def sample_pairs_level(self, chunk_spec):
        level_id = 0 if chunk_spec.level_id is None else int(chunk_spec.level_id)
        rows = self.level_rows(level_id)
        if chunk_spec.chunk_slice is not None:
            rows = rows[chunk_spec.chunk_slice]
        rows = np.asarray(rows)
        if level_id == 0:
            rows = rows[:, :1, :]
        return rows.transpose((2, 0, 1))

    def sample_pairs(self):
        return [self.sample_pairs_level(ChunkSpec(level_id=l)) for l in self.get_level_ids()]

    # ------------------------------------------------------------------ device feed
    def _resident(self):
        if not hasattr(self, "_resident_rows"):
            self._resident_rows = {}
        return self._resident_rows

    def drop_device_copies(self):
        self._resident().clear()

    def _host_tensor(self, level_id):
        """Rows as a (preferably pinned) torch CPU tensor; subclasses with pinned storage override."""
        return torch.from_numpy(np.ascontiguousarray(self.level_rows(level_id)))

    def device_rows(self, level_id, device, keep_resident=True, free_bytes=None):
        """Whole level as one CUDA tensor ``[N, 2, M]`` (uploaded once, cached) or None if it should be streamed."""
        key = (level_id, str(device))
        cache = self._resident()
        rows = cache.get(key)
        n_now = len(self.level_rows(level_id))
        if rows is not None and rows.shape[0] == n_now:
            return rows
        if self.resident_fraction <= 0:
            return None
        host = self._host_tensor(level_id)
        n_bytes = host.numel() * 8
        if free_bytes is None:
            free_bytes, _total = torch.cuda.mem_get_info(device)
        if n_bytes > self.resident_fraction * free_bytes:
            return None
        rows = torch.empty(host.shape, dtype=torch.float64, device=device)
        rows.copy_(host, non_blocking=True)
        if keep_resident:
            cache[key] = rows
        return rows

    def device_chunks(self, level_ids, device, keep_resident=True, row_ranges=None):
        """Yield ``(level_id, rows)`` with ``rows`` a CUDA tensor ``[n, 2, M]``, covering the given levels (or
        ``row_ranges[level_id] = (start, stop)`` of them) in level order.

        Levels that are (or may become) resident in HBM are yielded whole.  Everything else is copied from
        pinned host memory in ``device_chunk_bytes`` pieces on a side stream into two alternating device buffers,
        ONE pipeline across all levels, so the copy engine never idles between levels and the consumer's kernels
        on the current stream overlap the next copy."""
        free_bytes = None     # queried only when a level has to be placed (cudaMemGetInfo costs ~0.6 ms)
        pending = []          # (level_id, host rows) still to be streamed, in order
        for level_id in level_ids:
            rng = None if row_ranges is None else row_ranges[level_id]
            rows = None
            if keep_resident:
                was_cached = (level_id, str(device)) in self._resident()
                if not was_cached and free_bytes is None and self.resident_fraction > 0:
                    free_bytes, _total = torch.cuda.mem_get_info(device)
                rows = self.device_rows(level_id, device, True, free_bytes)
                if rows is not None and not was_cached and free_bytes is not None:
                    free_bytes -= rows.numel() * 8
            if rows is not None:
                if pending:
                    yield from stream_levels(pending, device, self.device_chunk_bytes)
                    pending = []
                yield level_id, (rows if rng is None else rows[rng[0]:rng[1]])
            else:
                host = self._host_tensor(level_id)
                pending.append((level_id, host if rng is None else host[rng[0]:rng[1]]))
        if pending:
            yield from stream_levels(pending, device, self.device_chunk_bytes)


def stream_levels(segments, device, chunk_bytes=32 << 20):
    """Double-buffered H2D streaming of ``[(level_id, host rows [N, 2, M]), ...]`` (pinned => truly asynchronous).
    Yields ``(level_id, device rows)``; a yielded tensor stays valid until the second next ``next()``."""
    pieces = []
    max_elems = 0
    for level_id, host in segments:
        n = host.shape[0]
        if n == 0:
            continue
        row_elems = host[0].numel()
        chunk_rows = max(1, min(n, chunk_bytes // (8 * row_elems)))
        for lo in range(0, n, chunk_rows):
            hi = min(lo + chunk_rows, n)
            pieces.append((level_id, host, lo, hi))
            max_elems = max(max_elems, (hi - lo) * row_elems)
    if not pieces:
        return
    compute = torch.cuda.current_stream(device)
    copy_stream = _copy_stream(device)
    bufs = [torch.empty(max_elems, dtype=torch.float64, device=device) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    copy_stream.wait_stream(compute)          # the buffers may be recycled memory still in use on the compute stream

    def issue(k):
        b = k % 2
        level_id, host, lo, hi = pieces[k]
        view = bufs[b][: (hi - lo) * host[0].numel()].view((hi - lo,) + tuple(host.shape[1:]))
        with torch.cuda.stream(copy_stream):
            if k >= 2:
                copy_stream.wait_event(consumed[b])
            view.copy_(host[lo:hi], non_blocking=True)
            copied[b].record(copy_stream)
        return level_id, view

    nxt = issue(0)
    for k in range(len(pieces)):
        cur = nxt
        if k + 1 < len(pieces):
            nxt = issue(k + 1)
        b = k % 2
        compute.wait_event(copied[b])
        yield cur
        consumed[b].record(compute)
    for buf in bufs:                           # allocated on the compute stream, written on the copy stream
        buf.record_stream(copy_stream)


_copy_streams = {}


def _copy_stream(device):
    key = str(device)
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(device)
    return _copy_streams[key]


def stream_rows(host, device, chunk_bytes=32 << 20):
    """Single-level convenience wrapper of ``stream_levels`` yielding the device rows only."""
    for _level, rows in stream_levels([(0, host)], device, chunk_bytes):
        yield rows


class Memory(SampleStorage):
    """In-memory storage (``mlmc/sample_storage.py:134-338``), rows kept in pinned host memory."""

    def __init__(self):
        self._rows = {}            # level_id -> torch CPU tensor [cap, 2, M] (pinned), first _n[level] rows valid
        self._n = {}
        self._failed = {}
        self._successful_sample_ids = {}
        self._scheduled = {}
        self._result_specification = []
        self._n_ops = {}
        self._n_finished = {}
        self._level_parameters = []

    @classmethod
    def from_arrays(cls, level_rows, level_parameters=None, n_ops=None, result_format=None):
        """Build a storage directly from per-level arrays ``float64[N_l, 2, M]``."""
        st = cls()
        for level_id, rows in enumerate(level_rows):
            st._append_rows(level_id, rows)
            st._n_finished[level_id] = len(rows)
        st._level_parameters = level_parameters
        if n_ops is not None:
            st._n_ops = {l: float(v) for l, v in enumerate(n_ops)}
        if result_format is not None:
            st._result_specification = result_format
        return st

    def _append_rows(self, level_id, rows):
        if isinstance(rows, torch.Tensor):
            rows = rows.detach().cpu()
            new = rows.to(torch.float64)
        else:
            new = torch.from_numpy(np.ascontiguousarray(np.asarray(rows, dtype=np.float64)))
        if new.dim() == 2:
            new = new.unsqueeze(2)
        if int(level_id) == 0 and new.shape[1] == 2:
            # level 0 has no coarse simulation: the reference stores an auxiliary zero row and drops it on every read
            # (sample_storage.py:282-284); here it is dropped once, so it is neither kept nor copied to the GPU
            new = new[:, :1, :]
        n_old = self._n.get(level_id, 0)
        buf = self._rows.get(level_id)
        if buf is None or n_old + len(new) > buf.shape[0]:
            cap = max(n_old + len(new), 2 * n_old)
            grown = _pinned_empty((cap,) + tuple(new.shape[1:]))
            if buf is not None:
                grown[:n_old].copy_(buf[:n_old])
            self._rows[level_id] = buf = grown
        buf[n_old:n_old + len(new)].copy_(new)
        self._n[level_id] = n_old + len(new)
        self._resident().pop((level_id, None), None)

    # ---- write side
    def save_samples(self, successful_samples, failed_samples):
        for level_id, samples in successful_samples.items():
            if len(samples) == 0:
                continue
            ids = [sid for sid, _ in samples]
            rows = np.array([np.stack([np.ravel(res[0]), np.ravel(res[1])]) for _, res in samples], dtype=np.float64)
            self._successful_sample_ids.setdefault(level_id, []).extend(ids)
            self._n_finished[level_id] = self._n_finished.get(level_id, 0) + len(samples)
            self._append_rows(level_id, rows)
        for level_id, failed in failed_samples.items():
            self._failed.setdefault(level_id, []).extend(failed)
            self._n_finished[level_id] = self._n_finished.get(level_id, 0) + len(failed)

    def save_global_data(self, result_format, level_parameters=None):
        self.save_result_format(result_format)
        self._level_parameters = level_parameters

    def save_result_format(self, res_spec):
        self._result_specification = res_spec

    def load_result_format(self):
        return self._result_specification

    def save_scheduled_samples(self, level_id, samples):
        self._scheduled.setdefault(level_id, []).extend(samples)

    def load_scheduled_samples(self):
        return self._scheduled

    def save_n_ops(self, n_ops):
        for level, (time, n_samples) in n_ops:
            if n_samples != 0:
                self._n_ops[level] = self._n_ops.get(level, 0) + time / n_samples

    # ---- read side
    def level_rows(self, level_id):
        level_id = int(level_id)
        return self._rows[level_id][: self._n[level_id]].numpy()

    def _host_tensor(self, level_id):
        level_id = int(level_id)
        return self._rows[level_id][: self._n[level_id]]

    def n_finished(self):
        out = np.zeros(max(self._n_finished) + 1 if self._n_finished else 0)
        for level_id, n in self._n_finished.items():
            out[level_id] = n
        return out

    def get_n_ops(self):
        return [self._n_ops.get(l, 0) for l in range(max(self._n_ops) + 1)] if self._n_ops else []

    def get_level_ids(self):
        return sorted(self._rows.keys())

    def get_n_collected(self):
        return [self._n[l] for l in self.get_level_ids()]

    def get_level_parameters(self):
        return self._level_parameters


class NpyStorage(SampleStorage):
    """File-backed read side in the reference's on-disk row order: one ``level_<l>.npy`` of shape ``[N, 2, M]``
    per level plus ``meta.npz`` (level parameters, n_ops).  It stands in for ``SampleStorageHDF``'s
    ``Levels/<l>/collected_values`` datasets (h5py / libhdf5 are not part of this image); levels are
    memory-mapped and streamed to the GPU through pinned staging buffers."""

    def __init__(self, directory):
        self._dir = directory
        meta = np.load(os.path.join(directory, "meta.npz"), allow_pickle=False)
        self._level_parameters = meta["level_parameters"].tolist()
        self._n_ops = meta["n_ops"].tolist()
        self._maps = {}
        self._format = []

    @staticmethod
    def write(directory, level_rows, level_parameters, n_ops=None):
        os.makedirs(directory, exist_ok=True)
        for l, rows in enumerate(level_rows):
            np.save(os.path.join(directory, "level_%d.npy" % l), np.ascontiguousarray(rows, dtype=np.float64))
        np.savez(os.path.join(directory, "meta.npz"), level_parameters=np.asarray(level_parameters, dtype=float),
                 n_ops=np.asarray(n_ops if n_ops is not None else [0.0] * len(level_rows), dtype=float))
        return NpyStorage(directory)

    def level_rows(self, level_id):
        level_id = int(level_id)
        if level_id not in self._maps:
            self._maps[level_id] = np.load(os.path.join(self._dir, "level_%d.npy" % level_id), mmap_mode="r")
        return self._maps[level_id]

    def _host_tensor(self, level_id):
        rows = self.level_rows(level_id)
        staged = _pinned_empty(rows.shape)
        staged.numpy()[...] = rows          # disk -> pinned host buffer
        return staged

    def get_level_ids(self):
        return list(range(len(self._level_parameters)))

    def get_level_parameters(self):
        return self._level_parameters

    def get_n_ops(self):
        return self._n_ops

    def save_samples(self, successful_samples, failed_samples):
        raise NotImplementedError("NpyStorage is a read-side adapter")

    def save_result_format(self, res_spec):
        self._format = res_spec

    def load_result_format(self):
        return self._format

    def save_global_data(self, result_format, level_parameters=None):
        self._format = result_format

    def save_scheduled_samples(self, level_id, samples):
        raise NotImplementedError("NpyStorage is a read-side adapter")

    def load_scheduled_samples(self):
        return {}

    def save_n_ops(self, n_ops):
        raise NotImplementedError("NpyStorage is a read-side adapter")


class SampleStorageHDF(SampleStorage):
    """Read-side adapter for the reference's HDF5 files (``mlmc/sample_storage_hdf.py:169-184``,
    ``mlmc/tool/hdf5.py:365-376``): dataset ``Levels/<l>/collected_values`` of array dtype ``(2, M) f8``.
    Needs ``h5py``, which this image does not ship -- constructing it without h5py raises ImportError."""

    def __init__(self, file_path):
        try:
            import h5py  # noqa: F401
        except ImportError as exc:
            raise ImportError("SampleStorageHDF needs h5py (not installed in this image); "
                              "use NpyStorage or Memory for the same row layout") from exc
        self._path = file_path
        self._h5py = h5py
        with h5py.File(file_path, "r") as f:
            self._level_parameters = np.array(f.attrs["level_parameters"]).tolist()
            self._levels = sorted(int(k) for k in f["Levels"].keys())
        self._format = []

    def level_rows(self, level_id):
        with self._h5py.File(self._path, "r") as f:
            return np.asarray(f["Levels"][str(int(level_id))]["collected_values"][()], dtype=np.float64)

    def get_level_ids(self):
        return self._levels

    def get_level_parameters(self):
        return self._level_parameters

    def get_n_ops(self):
        out = []
        with self._h5py.File(self._path, "r") as f:
            for l in self._levels:
                est = f["Levels"][str(l)].attrs.get("n_ops_estimate", [0.0, 0.0])
                out.append(est[0] / est[1] if est[1] > 0 else 0)
        return out

    def save_samples(self, successful_samples, failed_samples):
        raise NotImplementedError("read-side adapter")

    def save_result_format(self, res_spec):
        self._format = res_spec

    def load_result_format(self):
        return self._format

    def save_global_data(self, result_format, level_parameters=None):
        self._format = result_format

    def save_scheduled_samples(self, level_id, samples):
        raise NotImplementedError("read-side adapter")

    def load_scheduled_samples(self):
        return {}

    def save_n_ops(self, n_ops):
        raise NotImplementedError("read-side adapter")
