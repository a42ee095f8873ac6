"""ctypes binding of libmlmcb200.so (C ABI: include/mlmcb200.h).

PyTorch is used for device memory and streams only; every function here takes ``torch`` CUDA tensors (or raw
device pointers), passes plain pointers + sizes across the C ABI and launches on torch's current stream.
There is NO fallback: if the shared library is missing and cannot be built, importing the product path raises.
"""
import ctypes
import os
import threading

import torch

from . import build as _build

_c_i32, _c_i64, _c_dbl, _c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p

RAW, LEGENDRE, MONOMIAL, FOURIER = 0, 1, 2, 3
MAX_MOMENTS = 256
ABI_VERSION = 2


class BasisStruct(ctypes.Structure):
    """mlmcb200_basis_t"""
    _fields_ = [("kind", _c_i32), ("size", _c_i32), ("is_log", _c_i32), ("is_clip", _c_i32),
                ("shift", _c_dbl), ("scale", _c_dbl), ("ref_lo", _c_dbl), ("ref_hi", _c_dbl)]


class LevelStruct(ctypes.Structure):
    """mlmcb200_level_t"""
    _fields_ = [("pairs", _c_vp), ("n", _c_i64), ("stride_n", _c_i64), ("stride_side", _c_i64), ("stride_m", _c_i64),
                ("has_coarse", _c_i32), ("reserved", _c_i32), ("valid", _c_vp)]


RAW_BASIS = BasisStruct(RAW, 1, 0, 0, 0.0, 1.0, 0.0, 0.0)

_SIGNATURES = {
    "mlmcb200_abi_version": (ctypes.c_int, []),
    "mlmcb200_last_error": (ctypes.c_char_p, []),
    "mlmcb200_sm_count": (ctypes.c_int, []),
    "mlmcb200_basis_eval": (ctypes.c_int, [ctypes.POINTER(BasisStruct), _c_vp, _c_i64, _c_vp, _c_i32, _c_i32,
                                           _c_vp, _c_vp]),
    "mlmcb200_sample_mask": (ctypes.c_int, [ctypes.POINTER(BasisStruct), _c_vp, _c_i64, _c_i32, _c_i64, _c_i64,
                                            _c_i64, _c_i32, _c_vp, _c_vp]),
    "mlmcb200_moments_workspace_bytes": (_c_i64, [_c_i32, _c_i32]),
    "mlmcb200_moments_accumulate": (ctypes.c_int, [ctypes.POINTER(BasisStruct), _c_vp, _c_i64, _c_i32, _c_i64,
                                                   _c_i64, _c_i64, _c_i32, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp]),
    "mlmcb200_moments_accumulate_sums": (ctypes.c_int, [ctypes.POINTER(BasisStruct), _c_vp, _c_i64, _c_i32, _c_i64,
                                                        _c_i64, _c_i64, _c_i32, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp]),
    "mlmcb200_resample_indices": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_uint64, _c_i64, _c_i64, _c_i32, _c_i32,
                                                 _c_i32, _c_vp, _c_vp, _c_vp]),
    "mlmcb200_moments_resampled_workspace_bytes": (_c_i64, [_c_i32, _c_i32, _c_i32]),
    "mlmcb200_moments_accumulate_resampled": (ctypes.c_int, [ctypes.POINTER(BasisStruct), _c_vp, _c_i64, _c_i32,
                                                             _c_i64, _c_i64, _c_i64, _c_i32, _c_vp, _c_vp, _c_i64,
                                                             _c_i32, _c_vp, _c_i64, _c_vp, _c_i64, _c_vp]),
    "mlmcb200_resample_counts_block_rows": (_c_i32, []),
    "mlmcb200_resample_counts": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_uint64, _c_i64, _c_i64, _c_i32, _c_i32,
                                                _c_i32, _c_vp, _c_vp, _c_i64, _c_vp]),
    "mlmcb200_moments_weighted_max_size": (_c_i32, []),
    "mlmcb200_moments_weighted_workspace_bytes": (_c_i64, [_c_i64, _c_i32, _c_i32]),
    "mlmcb200_moments_accumulate_weighted": (ctypes.c_int, [ctypes.POINTER(BasisStruct), _c_vp, _c_i64, _c_i64, _c_i32,
                                                            _c_vp, _c_i64, _c_i32, _c_vp, _c_i64, _c_vp, _c_i64, _c_vp]),
    "mlmcb200_gram_workspace_bytes": (_c_i64, [_c_i32]),
    "mlmcb200_gram_accumulate": (ctypes.c_int, [ctypes.POINTER(BasisStruct), _c_vp, _c_i64, _c_i64, _c_i64,
                                                _c_i32, _c_i32, _c_i32, _c_vp, _c_vp, _c_i64, _c_vp]),
    "mlmcb200_gram_workspace_bytes_comp": (_c_i64, [_c_i32, _c_i32]),
    "mlmcb200_gram_accumulate_comp": (ctypes.c_int, [ctypes.POINTER(BasisStruct), _c_vp, _c_i64, _c_i32, _c_i64, _c_i64,
                                                     _c_i64, _c_i32, _c_vp, _c_i32, _c_i32, _c_vp, _c_vp, _c_i64,
                                                     _c_vp]),
    "mlmcb200_finalize_levels": (ctypes.c_int, [_c_vp, _c_i64, _c_i32, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp,
                                                _c_vp]),
    "mlmcb200_finalize_levels_batched": (ctypes.c_int, [_c_vp, _c_i64, _c_i32, _c_i64, _c_i32, _c_i64, _c_vp,
                                                        _c_vp]),
    "mlmcb200_estimate_moments_levels": (ctypes.c_int, [ctypes.POINTER(BasisStruct), ctypes.POINTER(LevelStruct), _c_i32,
                                                        _c_i32, _c_vp, _c_i64, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp]),
    "mlmcb200_level_sums_transform": (ctypes.c_int, [_c_vp, _c_i64, _c_i32, _c_i32, _c_i32, _c_vp, _c_i32, _c_vp,
                                                     _c_i64, _c_vp]),
    "mlmcb200_peer_buffer_bytes": (_c_i64, [_c_i32, _c_i64]),
    "mlmcb200_peer_alloc": (ctypes.c_int, [_c_i64, ctypes.POINTER(_c_vp), ctypes.POINTER(ctypes.c_ubyte)]),
    "mlmcb200_peer_open": (ctypes.c_int, [ctypes.POINTER(ctypes.c_ubyte), ctypes.POINTER(_c_vp)]),
    "mlmcb200_peer_close": (ctypes.c_int, [_c_vp]),
    "mlmcb200_peer_free": (ctypes.c_int, [_c_vp]),
    "mlmcb200_peer_error": (ctypes.c_int, [_c_vp, _c_i32, _c_i64, ctypes.POINTER(_c_i32)]),
    "mlmcb200_allreduce_finalize_levels": (ctypes.c_int, [_c_vp, _c_i64, _c_i32, _c_i64, _c_i32, _c_i32, _c_vp, _c_i64,
                                                          _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]),
    "mlmcb200_percentile_workspace_bytes": (_c_i64, [_c_i32]),
    "mlmcb200_percentile_stats": (ctypes.c_int, [_c_vp, _c_i64, _c_i64, ctypes.POINTER(_c_dbl), _c_i32, _c_vp, _c_vp,
                                                 _c_i64, _c_vp]),
    "mlmcb200_maxent_workspace_bytes": (_c_i64, [_c_i64, _c_i32]),
    "mlmcb200_maxent_fgh": (ctypes.c_int, [_c_vp, _c_i64, _c_vp, _c_vp, _c_i64, _c_i32, _c_i32, _c_vp, _c_vp,
                                           _c_i64, _c_vp]),
    "mlmcb200_density_eval": (ctypes.c_int, [ctypes.POINTER(BasisStruct), _c_vp, _c_i64, _c_vp, _c_i32, _c_vp, _c_vp]),
    "mlmcb200_fp64_peak": (ctypes.c_int, [_c_i32, ctypes.POINTER(_c_dbl), _c_vp]),
    "mlmcb200_host_copy_rows": (ctypes.c_int, [_c_vp, _c_vp, _c_i64, _c_i64, _c_i64, _c_i32]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_lock = threading.Lock()
#: number of kernel launches issued through this binding (bench.py reports it as ``gpu_launches``)
launch_count = 0


class NativeError(RuntimeError):
    pass


def library_path():
    return _build.LIB_PATH


def load():
    """Load (building first if the .so is absent).  Raises if neither works -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            path = library_path()
            if not os.path.exists(path):
                try:
                    _build.build()
                except Exception as exc:  # pragma: no cover - depends on toolchain
                    raise NativeError("libmlmcb200.so is missing and could not be built: %s" % exc)
            lib = ctypes.CDLL(path)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            if lib.mlmcb200_abi_version() != ABI_VERSION:
                raise NativeError("libmlmcb200.so ABI version mismatch")
            _lib = lib
    return _lib


def _check(rc, what):
    if rc != 0:
        raise NativeError("%s failed (%d): %s" % (what, rc, load().mlmcb200_last_error().decode()))


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _require_cuda(t, name, dtype=torch.float64):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype):
        raise NativeError("%s must be a CUDA tensor of dtype %s" % (name, dtype))


def sm_count():
    return load().mlmcb200_sm_count()


# ---------------------------------------------------------------------------------------------------------
def basis_eval(basis, x, n_out, matrix=None, out=None):
    """x: flat CUDA float64 tensor [n] -> Phi [n, n_out] (written into ``out`` if given: contiguous, that shape)."""
    global launch_count
    _require_cuda(x, "x")
    x = x.contiguous()
    if out is None:
        out = torch.empty((x.numel(), n_out), dtype=torch.float64, device=x.device)
    elif tuple(out.shape) != (x.numel(), n_out) or not out.is_contiguous() or out.dtype != torch.float64:
        raise NativeError("basis_eval: out must be a contiguous float64 tensor of shape (%d, %d)" % (x.numel(), n_out))
    n_rows = 0
    if matrix is not None:
        _require_cuda(matrix, "matrix")
        matrix = matrix.contiguous()
        n_rows = matrix.shape[0]
    with torch.cuda.device(x.device):
        _check(load().mlmcb200_basis_eval(ctypes.byref(basis), _ptr(x), x.numel(), _ptr(matrix), n_rows, n_out,
                                          _ptr(out), _stream()), "basis_eval")
    launch_count += 1
    return out


def _chunk_layout(x):
    """x: CUDA float64 tensor view [M, n, S] (S = 1 or 2) -> (M, n, has_coarse, stride_n, stride_side, stride_m)."""
    _require_cuda(x, "chunk")
    if x.dim() != 3 or x.shape[2] not in (1, 2):
        raise NativeError("chunk must have shape [M, n, 1|2], got %s" % (tuple(x.shape),))
    sm, sn, ss = x.stride()
    return x.shape[0], x.shape[1], int(x.shape[2] == 2), sn, ss, sm


class LevelAccumulator:
    """Device-side level sums for ``n_levels`` levels of ``K`` statistics each: [L, 2 + 2K] float64."""

    def __init__(self, n_levels, K, device, zero=True):
        self.n_levels, self.K = n_levels, K
        alloc = torch.zeros if zero else torch.empty
        self.acc = alloc((n_levels, 2 + 2 * K), dtype=torch.float64, device=device)

    def level(self, l):
        return self.acc[l]

    def finalize(self, peer=None):
        """-> dict of CUDA tensors l_means [L,K], l_vars [L,K], mean [K], var [K] (one launch).

        ``peer`` (``mlmc_b200.dist.peer_state()``): the launch first adds the accumulators of all ranks over NVLink
        peer memory (``mlmcb200_allreduce_finalize_levels``); ``self.acc`` then holds the global sums.  The result then
        carries ``status`` (CUDA tensor [1]): 1.0 if the launch gave up waiting for a peer -- ``self.acc`` still holds
        the LOCAL sums and the caller must reduce another way (``quantity_estimate.estimate_mean`` does)."""
        global launch_count
        dev = self.acc.device
        L, K = self.n_levels, self.K
        out = torch.empty((2 * L + 2, K), dtype=torch.float64, device=dev)
        l_means, l_vars, mean, var = out[:L], out[L:2 * L], out[2 * L], out[2 * L + 1]
        if peer is not None:
            status = torch.empty(1, dtype=torch.float64, device=dev)
            with torch.cuda.device(dev):
                _check(load().mlmcb200_allreduce_finalize_levels(
                    _ptr(self.acc), self.acc.stride(0), L, K, peer["rank"], peer["world"], _ptr(peer["ptrs"]),
                    peer["slot"], _ptr(l_means), _ptr(l_vars), _ptr(mean), _ptr(var), _ptr(status), _stream()),
                    "allreduce_finalize_levels")
            launch_count += 1
            return {"l_means": l_means, "l_vars": l_vars, "mean": mean, "var": var, "packed": out, "status": status}
        with torch.cuda.device(dev):
            _check(load().mlmcb200_finalize_levels(_ptr(self.acc), self.acc.stride(0), L, K, _ptr(l_means),
                                                   _ptr(l_vars), _ptr(mean), _ptr(var), _stream()),
                   "finalize_levels")
        launch_count += 1
        return {"l_means": l_means, "l_vars": l_vars, "mean": mean, "var": var, "packed": out}


_workspaces = {}


def _workspace(device, n_bytes):
    """Scratch buffer of the launching stream (partial sums): launches on one stream are ordered, launches on
    different streams (levels running concurrently) must not share it."""
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < n_bytes:
        if len(_workspaces) >= 32:                      # many short-lived streams: do not hoard their buffers
            _workspaces.clear()
        ws = torch.empty(max(int(n_bytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def sample_mask(basis, x):
    """valid[n] (uint8) for a vector chunk x [M, n, S]."""
    global launch_count
    M, n, has_coarse, sn, ss, sm = _chunk_layout(x)
    valid = torch.empty(n, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _check(load().mlmcb200_sample_mask(ctypes.byref(basis), _ptr(x), n, M, sn, ss, sm, has_coarse,
                                           _ptr(valid), _stream()), "sample_mask")
    launch_count += 1
    return valid


_ws_bytes_cache = {}


class _on_device:
    """``torch.cuda.device(dev)`` only when ``dev`` is not already current (the context manager costs ~20 us)."""

    def __init__(self, device):
        self._ctx = None
        if device.index is not None and device.index != torch.cuda.current_device():
            self._ctx = torch.cuda.device(device)

    def __enter__(self):
        if self._ctx is not None:
            self._ctx.__enter__()

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)


def moments_accumulate(basis, x, acc_row, valid=None, sums_only=False):
    """Add the chunk x [M, n, S] to ``acc_row`` ([2 + 2*M*R] float64, contiguous).  ``sums_only``: the caller does not
    read the sums of squares (they are then unspecified) -- ``mlmcb200_moments_accumulate_sums``."""
    global launch_count
    M, n, has_coarse, sn, ss, sm = _chunk_layout(x)
    _require_cuda(acc_row, "acc")
    K = M * basis.size
    if acc_row.numel() != 2 + 2 * K or not acc_row.is_contiguous():
        raise NativeError("accumulator must be contiguous with %d entries" % (2 + 2 * K))
    if n == 0:
        return
    lib = load()
    if M > 1 and valid is None and not (M <= 128 and basis.size <= 112):
        valid = sample_mask(basis, x)          # wide quantities: mask pass; narrow ones are masked inside the kernel
    with _on_device(x.device):
        key = (basis.size, M)
        ws_bytes = _ws_bytes_cache.get(key)
        if ws_bytes is None:
            ws_bytes = lib.mlmcb200_moments_workspace_bytes(basis.size, M)
            if ws_bytes < 0:
                raise NativeError("moments workspace: %s" % lib.mlmcb200_last_error().decode())
            _ws_bytes_cache[key] = ws_bytes
        ws = _workspace(x.device, ws_bytes)
        fn = lib.mlmcb200_moments_accumulate_sums if sums_only else lib.mlmcb200_moments_accumulate
        _check(fn(ctypes.byref(basis), _ptr(x), n, M, sn, ss, sm, has_coarse, _ptr(valid), _ptr(acc_row), _ptr(ws),
                  ws.numel(), _stream()), "moments_accumulate")
    launch_count += 2


def resample_indices(seed, stream_id, n_rows, n_draws, n_rep, device, block_cum=None, rep_offset=0):
    """int32 CUDA tensor [n_rep, n_draws] of row numbers drawn uniformly with replacement from ``range(n_rows)``;
    row b holds the draws of the GLOBAL replicate ``rep_offset + b`` (independent of the grouping into calls).
    ``block_cum`` ([n_rep, P + 1] int64 CUDA, cumulative multinomial block counts) lists each replicate's draws in
    row-block order (L2-friendly gather); reproducible for equal arguments."""
    global launch_count
    idx = torch.empty((n_rep, n_draws), dtype=torch.int32, device=device)
    n_blocks = 1
    if block_cum is not None:
        _require_cuda(block_cum, "block_cum", torch.int64)
        if block_cum.dim() != 2 or block_cum.shape[0] != n_rep or not block_cum.is_contiguous():
            raise NativeError("block_cum must be a contiguous [n_rep, n_blocks + 1] tensor")
        n_blocks = block_cum.shape[1] - 1
    with _on_device(idx.device):
        _check(load().mlmcb200_resample_indices(int(seed) & (2 ** 64 - 1), int(stream_id) & (2 ** 64 - 1), n_rows,
                                                n_draws, n_rep, int(rep_offset), n_blocks, _ptr(block_cum), _ptr(idx),
                                                _stream()),
               "resample_indices")
    launch_count += 1
    return idx


def moments_accumulate_resampled(basis, x, idx, acc_level, valid=None):
    """Bootstrap replicates of one level in one launch.

    x [M, n_rows, S]: the level's chunk; idx [B, n_draws] int32 (CUDA): rows drawn for each replicate;
    acc_level [B, 2 + 2*M*R]: view of the replicate accumulators of this level (row b may be strided by any
    replicate stride; each row contiguous).  Adds  sum over the drawn rows  exactly as ``moments_accumulate`` would
    for the gathered chunk ``x[:, idx[b], :]``."""
    global launch_count
    M, n_rows, has_coarse, sn, ss, sm = _chunk_layout(x)
    _require_cuda(acc_level, "acc")
    _require_cuda(idx, "idx", torch.int32)
    K = M * basis.size
    if idx.dim() != 2 or not idx.is_contiguous():
        raise NativeError("idx must be a contiguous [n_replicates, n_draws] tensor")
    B, n_draws = idx.shape
    if acc_level.dim() != 2 or acc_level.shape != (B, 2 + 2 * K) or acc_level.stride(1) != 1:
        raise NativeError("replicate accumulators must have shape [%d, %d] with contiguous rows" % (B, 2 + 2 * K))
    if n_draws == 0 or B == 0:
        return
    if n_rows == 0:
        raise NativeError("cannot draw from an empty chunk")
    lib = load()
    if M > 1 and valid is None and not (M <= 128 and basis.size <= 112):
        valid = sample_mask(basis, x)
    with _on_device(x.device):
        ws_bytes = lib.mlmcb200_moments_resampled_workspace_bytes(basis.size, M, B)
        if ws_bytes < 0:
            raise NativeError("moments workspace: %s" % lib.mlmcb200_last_error().decode())
        ws = _workspace(x.device, ws_bytes)
        _check(lib.mlmcb200_moments_accumulate_resampled(ctypes.byref(basis), _ptr(x), n_rows, M, sn, ss, sm,
                                                         has_coarse, _ptr(valid), _ptr(idx), n_draws, B,
                                                         _ptr(acc_level), acc_level.stride(0), _ptr(ws), ws.numel(),
                                                         _stream()),
               "moments_accumulate_resampled")
    launch_count += 2


def resample_counts(seed, stream_id, n_rows, block_cum, max_draws, device, rep_offset=0):
    """uint8 CUDA tensor [n_rep, stride] (stride = n_rows rounded up to 16): ``counts[b, i]`` = how many times replicate
    ``rep_offset + b`` drew row i -- the multiplicities of exactly the rows ``resample_indices`` lists for the same
    arguments.  ``block_cum`` ([n_rep, P + 1] int64 CUDA) is required (one block: ``[[0, k]]``)."""
    global launch_count
    _require_cuda(block_cum, "block_cum", torch.int64)
    if block_cum.dim() != 2 or not block_cum.is_contiguous() or block_cum.shape[1] < 2:
        raise NativeError("block_cum must be a contiguous [n_rep, n_blocks + 1] tensor")
    n_rep, n_blocks = block_cum.shape[0], block_cum.shape[1] - 1
    stride = -(-int(n_rows) // 16) * 16
    counts = torch.empty((n_rep, stride), dtype=torch.uint8, device=device)
    with _on_device(counts.device):
        _check(load().mlmcb200_resample_counts(int(seed) & (2 ** 64 - 1), int(stream_id) & (2 ** 64 - 1), int(n_rows),
                                               int(max_draws), n_rep, int(rep_offset), n_blocks, _ptr(block_cum),
                                               _ptr(counts), stride, _stream()), "resample_counts")
    launch_count += 1
    return counts


def counts_block_rows():
    return int(load().mlmcb200_resample_counts_block_rows())


def weighted_max_size():
    return int(load().mlmcb200_moments_weighted_max_size())


def moments_accumulate_weighted(basis, x, counts, acc_level):
    """All bootstrap replicates of one level in one pass (scalar quantity, any basis but RAW, size <=
    ``weighted_max_size()``).

    x [1, n_rows, S] in storage order; counts [B, stride] uint8 (``resample_counts``); acc_level [B, 2 + 2 R] (rows
    contiguous, any replicate stride).  Adds what ``moments_accumulate_resampled`` adds for the rows behind the counts."""
    global launch_count
    M, n_rows, has_coarse, sn, ss, sm = _chunk_layout(x)
    _require_cuda(acc_level, "acc")
    _require_cuda(counts, "counts", torch.uint8)
    if M != 1 or (has_coarse and (sn != 2 or ss != 1)):
        raise NativeError("moments_accumulate_weighted: scalar quantity in storage order expected")
    B = counts.shape[0]
    if counts.dim() != 2 or counts.stride(1) != 1 or counts.shape[1] < n_rows:
        raise NativeError("counts must be a [n_replicates, >= n_rows] tensor with contiguous rows")
    if acc_level.dim() != 2 or acc_level.shape != (B, 2 + 2 * basis.size) or acc_level.stride(1) != 1:
        raise NativeError("replicate accumulators must have shape [%d, %d] with contiguous rows" % (B, 2 + 2 * basis.size))
    if B == 0 or n_rows == 0:
        return
    lib = load()
    with _on_device(x.device):
        ws_bytes = lib.mlmcb200_moments_weighted_workspace_bytes(n_rows, B, basis.size)
        if ws_bytes < 0:
            raise NativeError("weighted moments workspace: %s" % lib.mlmcb200_last_error().decode())
        ws = _workspace(x.device, ws_bytes)
        _check(lib.mlmcb200_moments_accumulate_weighted(ctypes.byref(basis), _ptr(x), n_rows, sn, has_coarse,
                                                        _ptr(counts), counts.stride(0), B, _ptr(acc_level),
                                                        acc_level.stride(0), _ptr(ws), ws.numel(), _stream()),
               "moments_accumulate_weighted")
    launch_count += 2


def finalize_levels_batched(acc):
    """acc [B, L, 2 + 2K] (contiguous) -> CUDA tensor [B, 2*L*K + 2*K]: per replicate l_means | l_vars | mean | var."""
    global launch_count
    B, L, width = acc.shape
    K = (width - 2) // 2
    _require_cuda(acc, "acc")
    if not acc.is_contiguous():
        raise NativeError("replicate accumulators must be contiguous")
    out = torch.empty((B, 2 * L * K + 2 * K), dtype=torch.float64, device=acc.device)
    with _on_device(acc.device):
        _check(load().mlmcb200_finalize_levels_batched(_ptr(acc), acc.stride(1), L, K, B, acc.stride(0), _ptr(out),
                                                       _stream()), "finalize_levels_batched")
    launch_count += 1
    return out


class EstimatePlan:
    """Everything ``mlmcb200_estimate_moments_levels`` needs for one (quantity, basis) on resident levels, marshalled
    once: level descriptors, accumulator / result / pinned host buffers, workspace.  ``run()`` is ONE C call
    (memset + kernels + finalize + D2H + synchronise) and returns the packed host result as a NumPy view."""

    def __init__(self, basis, views, n_levels):
        """views: {level_id: CUDA tensor [M, n, S]} (levels without samples may be missing)."""
        first = next(iter(views.values()))
        self.device = first.device
        self.basis = basis
        self.M = int(first.shape[0])
        self.L = int(n_levels)
        self.K = self.M * basis.size
        if self.M > 128 or basis.size > 112 and self.M > 1:
            raise NativeError("EstimatePlan covers quantities the kernel masks itself (<= 128 components)")
        levels = (LevelStruct * self.L)()
        self.keys = []
        for l in range(self.L):
            x = views.get(l)
            if x is None:
                levels[l] = LevelStruct(None, 0, 0, 0, 0, 0, 0, None)
                self.keys.append(None)
                continue
            M, n, has_coarse, sn, ss, sm = _chunk_layout(x)
            if M != self.M:
                raise NativeError("EstimatePlan: levels differ in the number of components")
            levels[l] = LevelStruct(x.data_ptr(), n, sn, ss, sm, has_coarse, 0, None)
            self.keys.append((x.data_ptr(), tuple(x.shape), tuple(x.stride())))
        self.levels = levels
        self.views = views                       # keeps the rows alive
        self.acc = torch.zeros((self.L, 2 + 2 * self.K), dtype=torch.float64, device=self.device)
        n_out = 2 * self.L * self.K + 2 * self.K + 2 * self.L
        self.out = torch.empty(n_out, dtype=torch.float64, device=self.device)
        self.host = torch.empty(n_out, dtype=torch.float64).pin_memory()
        self.host_np = self.host.numpy()
        lib = load()
        ws_bytes = lib.mlmcb200_moments_workspace_bytes(basis.size, self.M)
        if ws_bytes < 0:
            raise NativeError("moments workspace: %s" % lib.mlmcb200_last_error().decode())
        self.ws = torch.empty(max(int(ws_bytes), 8), dtype=torch.uint8, device=self.device)
        self._args = (ctypes.byref(self.basis), self.levels, self.L, self.M, _ptr(self.acc), self.acc.stride(0),
                      _ptr(self.out), ctypes.c_void_p(self.host.data_ptr()), _ptr(self.ws), self.ws.numel())
        self._fn = lib.mlmcb200_estimate_moments_levels
        self.n_launches = 2 * sum(1 for k in self.keys if k is not None) + 1

    def matches(self, views):
        if len(views) != sum(1 for k in self.keys if k is not None):
            return False
        for l, key in enumerate(self.keys):
            x = views.get(l)
            if (x is None) != (key is None):
                return False
            if x is not None and key != (x.data_ptr(), tuple(x.shape), tuple(x.stride())):
                return False
        return True

    def run(self):
        global launch_count
        with _on_device(self.device):
            _check(self._fn(*self._args, _stream()), "estimate_moments_levels")
        launch_count += self.n_launches
        return self.host_np


def level_sums_transform(acc_in, n_comp, mat_t):
    """LevelAccumulator whose sums are ``mat_t^T`` applied to the sums of ``acc_in`` (per level and component);
    mat_t: CUDA float64 ``[K0, K1]``.  Counts are copied, sums of squares become NaN."""
    global launch_count
    _require_cuda(mat_t, "mat_t")
    K0, K1 = mat_t.shape
    if not mat_t.is_contiguous() or acc_in.K != n_comp * K0:
        raise NativeError("level_sums_transform: mat_t must be contiguous [K0, K1] with n_comp * K0 == %d" % acc_in.K)
    out = LevelAccumulator(acc_in.n_levels, n_comp * K1, acc_in.acc.device, zero=False)
    with _on_device(acc_in.acc.device):
        _check(load().mlmcb200_level_sums_transform(_ptr(acc_in.acc), acc_in.acc.stride(0), acc_in.n_levels, K0,
                                                    n_comp, _ptr(mat_t), K1, _ptr(out.acc), out.acc.stride(0),
                                                    _stream()), "level_sums_transform")
    launch_count += 1
    return out


def percentile_stats(values, fracs):
    """Neighbouring order statistics of a CUDA float64 vector (any stride) for ``np.percentile``-style fractions.

    Returns (stats [len(fracs), 2] NumPy, n_valid): ``stats[k] = sorted[floor(pos)], sorted[floor(pos) + 1]`` with
    ``pos = fracs[k] * (n_valid - 1)`` over the non-NaN entries -- exact values, found by radix selection."""
    global launch_count
    _require_cuda(values, "values")
    if values.dim() != 1 or values.numel() == 0:
        raise NativeError("percentile_stats needs a non-empty 1-D tensor")
    fracs = [float(f) for f in fracs]
    lib = load()
    arr = (_c_dbl * len(fracs))(*fracs)
    out = torch.empty(2 * len(fracs) + 1, dtype=torch.float64, device=values.device)
    with _on_device(values.device):
        ws_bytes = lib.mlmcb200_percentile_workspace_bytes(len(fracs))
        if ws_bytes < 0:
            raise NativeError("percentile_stats: at most 8 fractions per call")
        ws = _workspace(values.device, ws_bytes)
        _check(lib.mlmcb200_percentile_stats(_ptr(values), values.numel(), values.stride(0), arr, len(fracs),
                                             _ptr(out), _ptr(ws), ws.numel(), _stream()), "percentile_stats")
    launch_count += 12
    host = out.cpu().numpy()
    return host[:-1].reshape(len(fracs), 2), int(host[-1])


def gram_accumulate(basis, x, acc_row, mode=0, want_var=True, valid=None):
    """Add the chunk x [M, n, S] to the covariance accumulator ``acc_row`` ([2 + 2*M*R*R], flat index
    ``m*R*R + i*R + j``).  Vector quantities (M > 1) share one sample mask (``sample_mask``, computed here unless
    passed): a sample is dropped when any component leaves the domain."""
    global launch_count
    M, n, has_coarse, sn, ss, sm = _chunk_layout(x)
    _require_cuda(acc_row, "acc")
    R = basis.size
    if acc_row.numel() != 2 + 2 * M * R * R or not acc_row.is_contiguous():
        raise NativeError("accumulator must be contiguous with %d entries" % (2 + 2 * M * R * R))
    if n == 0:
        return
    lib = load()
    if M > 1 and valid is None:
        valid = sample_mask(basis, x)
    with _on_device(x.device):
        ws_bytes = lib.mlmcb200_gram_workspace_bytes_comp(R, M)
        if ws_bytes < 0:
            raise NativeError("gram workspace: %s" % lib.mlmcb200_last_error().decode())
        ws = _workspace(x.device, ws_bytes)
        _check(lib.mlmcb200_gram_accumulate_comp(ctypes.byref(basis), _ptr(x), n, M, sn, ss, sm, has_coarse,
                                                 _ptr(valid if M > 1 else None), mode, int(bool(want_var)),
                                                 _ptr(acc_row), _ptr(ws), ws.numel(), _stream()),
               "gram_accumulate")
    launch_count += 2


def maxent_workspace(phi, n_moments):
    """A private scratch tensor for ``maxent_fgh`` on ``phi`` (callers that capture the evaluation in a CUDA graph)."""
    n_bytes = load().mlmcb200_maxent_workspace_bytes(phi.shape[0], n_moments)
    if n_bytes < 0:
        raise NativeError("maxent workspace: %s" % load().mlmcb200_last_error().decode())
    return torch.empty(max(int(n_bytes), 8), dtype=torch.uint8, device=phi.device)


def maxent_fgh(phi, w, lam_scaled, what=7, out=None, workspace=None):
    """phi [Q, ld] (R = len(lam_scaled) <= ld), w [Q] -> out [1 + R + R*R] (F, g, H integral terms)."""
    global launch_count
    _require_cuda(phi, "phi")
    _require_cuda(w, "w")
    _require_cuda(lam_scaled, "lam_scaled")
    Q, ld = phi.shape
    R = lam_scaled.numel()
    if phi.stride(1) != 1 or phi.stride(0) != ld or R > ld:
        raise NativeError("phi must be row-major [Q, ld] with R <= ld")
    if out is None:
        out = torch.zeros(1 + R + R * R, dtype=torch.float64, device=phi.device)
    lib = load()
    with _on_device(phi.device):
        if workspace is not None:
            ws = workspace
        else:
            ws_bytes = lib.mlmcb200_maxent_workspace_bytes(Q, R)
            ws = _workspace(phi.device, ws_bytes)
        _check(lib.mlmcb200_maxent_fgh(_ptr(phi), ld, _ptr(w), _ptr(lam_scaled), Q, R, what, _ptr(out), _ptr(ws),
                                       ws.numel(), _stream()), "maxent_fgh")
    launch_count += 3
    return out


def density_eval(basis, x, coef):
    """exp(clip(-phi(x) . coef, -200, 200)) for a flat CUDA float64 tensor x; coef: CUDA float64 [n_coef <= basis.size]."""
    global launch_count
    _require_cuda(x, "x")
    _require_cuda(coef, "coef")
    x = x.contiguous()
    coef = coef.contiguous()
    out = torch.empty(x.numel(), dtype=torch.float64, device=x.device)
    with _on_device(x.device):
        _check(load().mlmcb200_density_eval(ctypes.byref(basis), _ptr(x), x.numel(), _ptr(coef), coef.numel(),
                                            _ptr(out), _stream()), "density_eval")
    launch_count += 1
    return out


def host_copy_rows(dst_address, src_address, n_rows, keep_bytes, src_pitch, n_threads):
    """Host copy of the staged feed: ``n_rows`` rows of ``keep_bytes`` (``src_pitch`` apart in the source) packed into
    ``dst`` by ``n_threads`` native threads (the call releases the interpreter lock).  Addresses are plain integers."""
    _check(load().mlmcb200_host_copy_rows(ctypes.c_void_p(int(dst_address)), ctypes.c_void_p(int(src_address)),
                                          int(n_rows), int(keep_bytes), int(src_pitch), int(n_threads)),
           "host_copy_rows")


def fp64_peak(kind):
    """Measured FP64 pipe throughput in FLOP/s: kind 0 = DFMA, 1 = DMMA m8n8k4."""
    val = _c_dbl(0.0)
    _check(load().mlmcb200_fp64_peak(kind, ctypes.byref(val), _stream()), "fp64_peak")
    return val.value


def make_basis(kind, size, domain, ref_domain, log=False, safe_eval=True):
    """mlmcb200_basis_t from ``Moments.__init__`` parameters (mlmc/moments.py:10-26)."""
    import math
    lo, hi = (math.log(domain[0]), math.log(domain[1])) if log else (float(domain[0]), float(domain[1]))
    width = hi - lo
    assert width > 0
    width = max(width, 1e-15)
    scale = (ref_domain[1] - ref_domain[0]) / width
    return BasisStruct(kind, int(size), int(bool(log)), int(bool(safe_eval)), lo, scale,
                       float(ref_domain[0]), float(ref_domain[1]))
