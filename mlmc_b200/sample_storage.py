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
        return [self.level_n_rows(l) for l in self.get_level_ids()]

    def n_finished(self):
        return np.array(self.get_n_collected(), dtype=float)

    def unfinished_ids(self):
        return []

    def chunks(self, level_id=None, n_samples=None):
        assert isinstance(n_samples, (type(None), int)), "n_samples param must be int"
        level_ids = self.get_level_ids() if level_id is None else [level_id]
        return itertools.chain(*[self._level_chunks(l, n_samples) for l in level_ids])

    def _level_chunks(self, level_id, n_samples=None):
        n = self.level_n_rows(level_id)
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
        """Where the streaming feed reads a level from: a pinned torch CPU tensor (``Memory``) or a ``RowSource`` over
        file-backed rows (level 0 without its zero coarse row)."""
        return RowSource(self.level_rows(level_id), drop_coarse=int(level_id) == 0)

    def level_n_rows(self, level_id):
        """Number of collected rows of a level, WITHOUT reading them (array / memmap / HDF5 dataset shape)."""
        return int(self.level_rows(level_id).shape[0])

    def device_rows(self, level_id, device, keep_resident=True, free_bytes=None):
        """Whole level as one CUDA tensor ``[N, 2, M]`` (uploaded once, cached) or None if it should be streamed."""
        key = (level_id, str(device))
        cache = self._resident()
        rows = cache.get(key)
        n_now = self.level_n_rows(level_id)
        if rows is not None and rows.shape[0] == n_now:
            return rows
        if self.resident_fraction <= 0:
            return None
        host = self._host_tensor(level_id)
        n_bytes = 8 * int(np.prod(host.shape))
        if free_bytes is None:
            free_bytes, _total = torch.cuda.mem_get_info(device)
        if n_bytes > self.resident_fraction * free_bytes:
            return None
        rows = torch.empty(tuple(host.shape), dtype=torch.float64, device=device)
        if isinstance(host, torch.Tensor):
            rows.copy_(host, non_blocking=True)
        else:                                   # file-backed: through the staged pipeline, piece by piece
            at = 0
            for _level, piece in stream_levels([(level_id, host)], device, self.device_chunk_bytes):
                rows[at:at + piece.shape[0]].copy_(piece)
                at += piece.shape[0]
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
                if rng is not None:
                    host = host[rng[0]:rng[1]] if isinstance(host, torch.Tensor) else host.slice_rows(rng[0], rng[1])
                pending.append((level_id, host))
        if pending:
            yield from stream_levels(pending, device, self.device_chunk_bytes)


class RowSource:
    """Rows ``float64[N, S, M]`` of one level that live in a file (memory map, HDF5 dataset): ``read_into(lo, hi, out)``
    fills a C-contiguous NumPy buffer ``[hi - lo, S, M]`` -- a pinned staging buffer on the streaming path.  Level 0
    drops the auxiliary zero coarse row HERE (``S = 1``), so it is neither staged nor copied to the GPU."""

    def __init__(self, rows, drop_coarse=False, lo=0, hi=None):
        self._rows = rows                       # NumPy array / memmap [N, 2, M] or an object with read_rows(lo, hi, out)
        n, sides, m = rows.shape
        self._lo = lo
        self._hi = n if hi is None else hi
        self._drop = bool(drop_coarse) and sides == 2
        self.shape = (self._hi - self._lo, 1 if self._drop else sides, m)

    def __len__(self):
        return self.shape[0]

    def slice_rows(self, lo, hi):
        lo, hi = max(0, min(lo, len(self))), max(0, min(hi, len(self)))
        return RowSource(self._rows, self._drop, self._lo + lo, self._lo + max(lo, hi))

    def _read_part(self, lo, hi, out):
        if hasattr(self._rows, "read_rows"):
            if self._drop and getattr(self._rows, "keeps_items", False):
                self._rows.read_rows(lo, hi, out=out, keep_items=self.shape[2])    # fine row only, straight from the file
            elif self._drop:
                # the stored zero coarse row is skipped block by block (a temporary of a few MB, not of the whole piece)
                step = max(1, (4 << 20) // (16 * self.shape[2]))
                for a in range(lo, hi, step):
                    b = min(hi, a + step)
                    out[a - lo:b - lo] = self._rows.read_rows(a, b)[:, :1, :]
            else:
                self._rows.read_rows(lo, hi, out=out)
        else:
            rows = self._rows
            copy = _native_copy()
            n_bytes = (hi - lo) * self.shape[1] * self.shape[2] * 8
            if (copy is not None and n_bytes >= (2 << 20) and isinstance(rows, np.ndarray) and rows.flags.c_contiguous
                    and out.flags.c_contiguous and rows.dtype == np.float64 and out.dtype == np.float64):
                pitch = rows.shape[1] * rows.shape[2] * 8            # memory map of a .npy file: rows `pitch` bytes apart
                copy(out.ctypes.data, rows[lo:hi].ctypes.data, hi - lo, self.shape[1] * self.shape[2] * 8, pitch)
            else:
                out[...] = rows[lo:hi, :1, :] if self._drop else rows[lo:hi]

    def read_into(self, lo, hi, out):
        """Copies out of the page cache run at a few GB/s per thread -- a tenth of the PCIe rate behind them -- so a
        piece is cut into row ranges that a small thread pool copies concurrently (NumPy releases the GIL in the copy
        loops; the h5py backend serialises on its own lock and stays single-threaded)."""
        lo, hi = self._lo + lo, self._lo + hi
        n_bytes = (hi - lo) * self.shape[1] * self.shape[2] * 8
        threads = _read_threads()
        native = _native_copy() is not None and (isinstance(self._rows, np.ndarray) or
                                                 getattr(self._rows, "native_copy", False))
        if native or threads <= 1 or n_bytes < (8 << 20) or not getattr(self._rows, "parallel_reads", True):
            # (native: the copy itself is cut over host threads inside the library, without the interpreter lock)
            self._read_part(lo, hi, out)
            return
        parts = min(threads, max(1, n_bytes // (4 << 20)), hi - lo)
        edges = [lo + (hi - lo) * i // parts for i in range(parts + 1)]
        futures = [_read_pool().submit(self._read_part, edges[i], edges[i + 1], out[edges[i] - lo:edges[i + 1] - lo])
                   for i in range(parts) if edges[i + 1] > edges[i]]
        for f in futures:
            f.result()


_pool = {}


def _native_copy():
    """``f(dst, src, n_rows, keep_bytes, src_pitch)`` over ``mlmcb200_host_copy_rows`` with the configured number of
    threads, or None when the library cannot be loaded (the readers then copy with NumPy)."""
    if "copy" not in _pool:
        try:
            from . import _native
            _native.load()
            _pool["copy"] = lambda d, s, n, k, p: _native.host_copy_rows(d, s, n, k, p, _read_threads())
        except Exception:
            _pool["copy"] = None
    return _pool["copy"]


def _read_threads():
    """Threads of the file -> pinned staging copy (``MLMCB200_READ_THREADS``; default: up to 16, at most the cores --
    one thread copies ~11 GB/s out of the page cache, sixteen 40-58 GB/s: profiles/r2_feed_probe.txt)."""
    n = os.environ.get("MLMCB200_READ_THREADS")
    if n is not None:
        return max(1, int(n))
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    return max(1, min(16, cores))


def _read_pool():
    n = _read_threads()
    if _pool.get("n") != n:
        from concurrent.futures import ThreadPoolExecutor
        _pool["pool"], _pool["n"] = ThreadPoolExecutor(max_workers=n, thread_name_prefix="mlmcb200-read"), n
    return _pool["pool"]


_staging = {}


def _staging_buffers(n_bytes):
    """Two fixed pinned staging buffers (reused by every staged stream of this process)."""
    bufs = _staging.get("bufs")
    if bufs is None or bufs[0].numel() * 8 < n_bytes:
        n = max((int(n_bytes) + 7) // 8, 1 << 20)
        bufs = _staging["bufs"] = [_pinned_empty((n,)) for _ in range(2)]
    return bufs


def stream_levels(segments, device, chunk_bytes=32 << 20):
    """Double-buffered H2D streaming of ``[(level_id, host rows [N, S, M]), ...]``.
    Yields ``(level_id, device rows)``; a yielded tensor stays valid until the second next ``next()``.

    ``host rows`` is a torch CPU tensor (pinned => truly asynchronous copies straight from it: ``Memory``) or a
    ``RowSource`` (file-backed levels).  The latter go through a THREE-stage pipeline: file -> one of two fixed pinned
    staging buffers (CPU copy out of the page cache / HDF5 chunks) -> one of two device buffers (copy stream) -> the
    consumer's kernels (compute stream); while the kernels of piece k run, piece k + 1 is on the copy engine and the
    CPU reads piece k + 2.  No buffer of the size of a level is ever pinned."""
    pieces = []
    max_elems = 0
    staged = False
    for level_id, host in segments:
        n = host.shape[0]
        if n == 0:
            continue
        row_elems = int(host.shape[1]) * int(host.shape[2])
        chunk_rows = max(1, min(n, chunk_bytes // (8 * row_elems)))
        staged = staged or not isinstance(host, torch.Tensor)
        for lo in range(0, n, chunk_rows):
            hi = min(lo + chunk_rows, n)
            pieces.append((level_id, host, lo, hi))
            max_elems = max(max_elems, (hi - lo) * row_elems)
    if not pieces:
        return
    compute = torch.cuda.current_stream(device)
    copy_stream = _copy_stream(device)
    bufs = [torch.empty(max_elems, dtype=torch.float64, device=device) for _ in range(2)]
    for buf in bufs:
        # allocated on the compute stream, written on the copy stream: tell the caching allocator NOW, so that a consumer
        # that raises mid-iteration (the generator is then closed at the `yield`) cannot hand the memory out again while
        # a prefetch is still in flight
        buf.record_stream(copy_stream)
    stage = _staging_buffers(max_elems * 8) if staged else None
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    copy_stream.wait_stream(compute)          # the buffers may be recycled memory still in use on the compute stream

    def geometry(k):
        level_id, host, lo, hi = pieces[k]
        shape = (hi - lo,) + tuple(int(d) for d in host.shape[1:])
        return level_id, host, lo, hi, shape, shape[0] * shape[1] * shape[2]

    def read(k):
        """CPU stage (file-backed pieces; runs on the reader thread): fill staging buffer k % 2 with piece k."""
        b = k % 2
        level_id, host, lo, hi, shape, n_elems = geometry(k)
        if isinstance(host, torch.Tensor):
            return
        if k >= 2:
            copied[b].synchronize()           # the copy that last read this staging buffer has finished
        host.read_into(lo, hi, stage[b][:n_elems].view(shape).numpy())

    def enqueue(k):
        """Copy stage: piece k (pinned rows or its staging buffer) -> device buffer k % 2 on the copy stream."""
        b = k % 2
        level_id, host, lo, hi, shape, n_elems = geometry(k)
        view = bufs[b][:n_elems].view(shape)
        src = host[lo:hi] if isinstance(host, torch.Tensor) else stage[b][:n_elems].view(shape)
        with torch.cuda.stream(copy_stream):
            if k >= 2:
                copy_stream.wait_event(consumed[b])
            view.copy_(src, non_blocking=True)
            copied[b].record(copy_stream)
        return level_id, view

    # The CPU stage of piece k + 2 runs right before piece k is handed to the consumer (the copy engine then moves piece
    # k + 1 and the kernels of piece k - 1 may still run); MLMCB200_FEED_THREAD=1 moves it to a background thread instead.
    # Measured on cfg5 from an HDF5 file (profiles/r2_feed_probe.txt): the background thread gains nothing (39.0 against
    # 38.0 ms; in the bench it lost) -- it and the consumer hand the interpreter lock back and forth around every call.
    pending = {}

    background = staged and os.environ.get("MLMCB200_FEED_THREAD", "0") == "1"

    def submit(k):
        if staged and k < len(pieces):
            if background:
                pending[k] = _reader().submit(read, k)
            else:
                read(k)

    def collect(k):
        f = pending.pop(k, None)
        if f is not None:
            f.result()

    try:
        submit(0)
        collect(0)
        nxt = enqueue(0)
        submit(1)
        for k in range(len(pieces)):
            cur = nxt
            if k + 1 < len(pieces):
                collect(k + 1)
                nxt = enqueue(k + 1)
            submit(k + 2)                     # staging buffer k % 2 again: its last reader is the copy of piece k, enqueued
            b = k % 2
            compute.wait_event(copied[b])
            yield cur
            consumed[b].record(compute)
    finally:
        for f in list(pending.values()):      # nobody may still be writing the shared staging buffers ...
            try:
                f.result()
            except Exception:
                pass
        pending.clear()
        if staged:                            # ... or reading them
            copy_stream.synchronize()


def _reader():
    if "reader" not in _pool:
        from concurrent.futures import ThreadPoolExecutor
        _pool["reader"] = ThreadPoolExecutor(max_workers=1, thread_name_prefix="mlmcb200-feed")
    return _pool["reader"]


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
        cache = self._resident()
        for key in [k for k in cache if k[0] == level_id]:       # resident copies (any device) of this level are stale
            del cache[key]

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

    def level_n_rows(self, level_id):
        return self._n[int(level_id)]

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
    memory-mapped (row counts come from the ``.npy`` header) and streamed to the GPU through two fixed pinned staging
    buffers (``stream_levels``): file read, H2D copy and kernels overlap, no level-sized pinned allocation."""

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


class _H5pyRows:
    """``read_rows`` over an h5py dataset of one level (the file is opened per read, like the reference does)."""
    parallel_reads = False          # libhdf5 serialises on a global lock

    def __init__(self, h5py, path, name, shape):
        self._h5py, self._path, self._name, self.shape = h5py, path, name, shape

    def read_rows(self, lo=0, hi=None, out=None):
        hi = self.shape[0] if hi is None else hi
        with self._h5py.File(self._path, "r") as f:
            dset = f[self._name]
            if out is None:
                return np.asarray(dset[lo:hi], dtype=np.float64)
            dset.read_direct(out, np.s_[lo:hi])
            return out

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, key):
        if isinstance(key, slice):
            lo, hi, step = key.indices(self.shape[0])
            rows = self.read_rows(lo, hi)
            return rows if step == 1 else rows[::step]
        return self.read_rows()[key]

    def __array__(self, dtype=None, copy=None):
        return self.read_rows()


class _MinRows(_H5pyRows):
    """The same over ``mlmc_b200.tool.hdf5_min`` (memory-mapped file, no h5py)."""
    parallel_reads = True

    def __init__(self, dataset, shape):
        self._dset, self.shape = dataset, shape
        self.keeps_items = not dataset._filters       # level 0 without its zero coarse row, straight from the mapping
        from .tool import hdf5_min
        if hdf5_min.parallel_copy is None and _native_copy() is not None:
            hdf5_min.parallel_copy = _native_copy()
        self.native_copy = hdf5_min.parallel_copy is not None and not dataset._filters

    def read_rows(self, lo=0, hi=None, out=None, keep_items=None):
        return self._dset.read_rows(lo, hi, out=out, keep_items=keep_items)


class SampleStorageHDF(SampleStorage):
    """Read side of the reference's HDF5 sample files (``mlmc/sample_storage_hdf.py:144-184``,
    ``mlmc/tool/hdf5.py:311-320, 353-376``): dataset ``Levels/<l>/collected_values`` -- 1-D, chunked, element type
    ``(2, M)`` float64 -- root attribute ``level_parameters``, level attribute ``n_ops_estimate``.

    Opens with h5py when it is importable, otherwise with the built-in minimal reader
    (``mlmc_b200.tool.hdf5_min``; h5py / libhdf5 are not part of this image).  Nothing is read at construction but
    the metadata: row counts come from the dataset shapes (the reference's ``collected_n_items`` loads the whole
    dataset, ``hdf5.py:378-389``).  Rows reach the GPU through the staged pipeline of ``stream_levels`` (HDF5 chunks ->
    two pinned staging buffers -> device), level 0 without its stored zero coarse row."""

    def __init__(self, file_path, backend=None):
        self._path = file_path
        self._format = []
        if backend not in (None, "h5py", "min"):
            raise ValueError("backend must be None, 'h5py' or 'min'")
        h5py = None
        if backend != "min":
            try:
                import h5py
            except ImportError:
                if backend == "h5py":
                    raise
        self._rows = {}
        self._n_ops = {}
        if h5py is not None:
            self.backend = "h5py"
            with h5py.File(file_path, "r") as f:
                self._level_parameters = np.array(f.attrs["level_parameters"]).tolist()
                for key in f["Levels"].keys():
                    group = f["Levels"][key]
                    if "collected_values" not in group:
                        continue
                    dset = group["collected_values"]
                    shape = (int(dset.shape[0]),) + tuple(int(d) for d in dset.dtype.shape)
                    self._rows[int(key)] = _H5pyRows(h5py, file_path, "Levels/%s/collected_values" % key, shape)
                    self._n_ops[int(key)] = np.array(group.attrs.get("n_ops_estimate", [0.0, 0.0]), dtype=float)
        else:
            from .tool import hdf5_min
            self.backend = "min"
            self._file = hdf5_min.File(file_path)
            params = self._file.attrs.get("level_parameters")
            if params is None:
                raise Exception("'level_parameters' aren't store in HDF file, so unable to create level groups")
            self._level_parameters = np.asarray(params).tolist()
            levels = self._file["Levels"]
            for key in levels.keys():
                group = levels[key]
                if "collected_values" not in group:
                    continue
                dset = group["collected_values"]
                sub = dset.dtype.shape if dset.dtype.subdtype else ()
                if len(dset.shape) != 1 or len(sub) != 2 or dset.dtype.base != np.dtype("<f8"):
                    raise ValueError("Levels/%s/collected_values: expected a 1-D dataset of (2, M) float64 elements" % key)
                self._rows[int(key)] = _MinRows(dset, (int(dset.shape[0]),) + tuple(int(d) for d in sub))
                ops = group.attrs.get("n_ops_estimate")
                self._n_ops[int(key)] = np.asarray(ops if ops is not None else [0.0, 0.0], dtype=float).reshape(-1)
        self._levels = sorted(self._rows)

    def level_rows(self, level_id):
        """Lazy rows ``[N, 2, M]`` of a level: ``len()``, ``.shape``, slicing and ``np.asarray`` read from the file."""
        return self._rows[int(level_id)]

    def level_n_rows(self, level_id):
        return self._rows[int(level_id)].shape[0]

    def sample_pairs_level(self, chunk_spec):
        level_id = 0 if chunk_spec.level_id is None else int(chunk_spec.level_id)
        sl = chunk_spec.chunk_slice if chunk_spec.chunk_slice is not None else slice(None)
        rows = self._rows[level_id][sl]
        if level_id == 0:
            rows = rows[:, :1, :]
        return rows.transpose((2, 0, 1))

    def get_level_ids(self):
        return self._levels

    def get_level_parameters(self):
        return self._level_parameters

    def get_n_ops(self):
        """``n_ops_estimate`` = (accumulated time, number of samples) per level -> time per sample
        (``sample_storage_hdf.py:225-241``)."""
        out = []
        for l in self._levels:
            est = self._n_ops.get(l, np.zeros(2))
            out.append(float(est[0] / est[1]) if len(est) > 1 and est[1] > 0 else 0)
        return out

    def save_samples(self, successful_samples, failed_samples):
        raise NotImplementedError("read-side adapter (write with the reference, or hdf5_min.write_mlmc_file)")

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
