"""Estimation operators -- mirror of ``mlmc/quantity/quantity_estimate.py:6-156`` on the fused CUDA path.

``moments`` / ``covariance`` / ``moment`` build lazy quantities exactly like the reference.  ``estimate_mean``
recognises them and, instead of materialising the moment (or outer-product) chunks, streams the INPUT quantity's
device chunks ``[M, n, 2]`` through the fused kernels:

=====================================  ======================================================================
quantity                               kernels
=====================================  ======================================================================
``moments(q, basis)``                  ``mlmcb200_moments_accumulate`` (+ ``sample_mask`` when M > 128)
``moments(q, TransformedMoments)``     moments sums of the base functions + DMMA Gram of the differences,
                                       then mean' = L s, sum d'^2 = diag(L G L^T)   (per component of q)
``covariance(q, basis)``               ``mlmcb200_gram_accumulate_comp`` (DMMA), sums and sums of squares; vector
                                       quantities: all components in one launch, one shared sample mask
``covariance(q, basis)``, means only   ``estimate_mean(..., variance=False)``: products of basis functions are linear
                                       combinations of a longer basis of the same family, so the level sums are
                                       ``C . (moment sums of 2R-1 functions)``: ``mlmcb200_moments_accumulate`` +
                                       ``mlmcb200_level_sums_transform`` (any M, also TransformedMoments)
anything else                          generic: chunk evaluated by device ops, reduced by the RAW kernel
=====================================  ======================================================================

Level sums are reduced across ranks (``mlmc_b200.dist``) when sample sharding is active, then finalised on the
device (``mlmcb200_finalize_levels``) and copied to the host once.
"""
import os

import numpy as np
import torch

from .. import _native
from .. import dist as _dist
from ..moments import TransformedMoments
from . import quantity as q_mod
from . import quantity_types as qt


def mask_nan_samples(chunk):
    """Drop samples with a NaN in the fine or the coarse part (quantity_estimate.py:6-14).
    Accepts a NumPy array or a CUDA tensor ``[M, n, S]``; returns (compacted chunk, number dropped)."""
    if isinstance(chunk, torch.Tensor):
        bad = torch.isnan(chunk).any(dim=0).any(dim=1)
        return chunk[:, ~bad, :], int(bad.sum().item())
    bad = np.any(np.isnan(chunk), axis=0).any(axis=1)
    return chunk[..., ~bad, :], np.count_nonzero(bad)


def cache_clear():
    """The reference clears its per-chunk memoisation here (quantity_estimate.py:17-19); this implementation
    keeps no sample caches besides the storage's resident device copy, which stays valid."""
    return None


# ----------------------------------------------------------------------------------------------------------
class _MomentsQuantity(q_mod.Quantity):
    """``moments(quantity, moments_fn)``: lazy quantity that remembers what it is, so that ``estimate_mean`` can fuse."""

    def __init__(self, quantity_type, operation, input_quantity, moments_fn, at_bottom, kind):
        super().__init__(quantity_type, operation, [input_quantity])
        self._moments_fn = moments_fn
        self._at_bottom = at_bottom
        self._fused_kind = kind          # "moments" | "covariance"


def moment(quantity, moments_fn, i=0):
    """Quantity of a single moment function (quantity_estimate.py:83-93)."""
    def eval_moment(x):
        return moments_fn.eval_single_moment(i, value=x)
    return q_mod.Quantity(quantity_type=quantity.qtype, input_quantities=[quantity], operation=eval_moment)


def moments(quantity, moments_fn, mom_at_bottom=True):
    """Quantity of all moment functions (quantity_estimate.py:96-119).  Chunk ``[M, n, S] -> [M*R, n, S]`` with
    flat index ``m*R + r`` (``mom_at_bottom``) or ``r*M + m``."""
    def eval_moments(x):
        mom = moments_fn.eval_all(x)                                   # [M, n, S, R]
        mom = mom.permute(0, 3, 1, 2) if mom_at_bottom else mom.permute(3, 0, 1, 2)
        return mom.reshape((-1,) + tuple(mom.shape[-2:]))

    if mom_at_bottom:
        moments_array_type = qt.ArrayType(shape=(moments_fn.size,), qtype=qt.ScalarType())
        moments_qtype = quantity.qtype.replace_scalar(moments_array_type)
    else:
        moments_qtype = qt.ArrayType(shape=(moments_fn.size,), qtype=quantity.qtype)
    return _MomentsQuantity(moments_qtype, eval_moments, quantity, moments_fn, mom_at_bottom, "moments")


def covariance(quantity, moments_fn, cov_at_bottom=True):
    """Quantity of the moment outer products (quantity_estimate.py:122-156): ``[M, n, S] -> [M*R*R, n, S]``."""
    def eval_cov(x):
        mom = moments_fn.eval_all(x)                                   # [M, n, S, R]
        cov = mom.unsqueeze(-1) * mom.unsqueeze(-2)                    # [M, n, S, R, R]
        cov = cov.permute(0, 3, 4, 1, 2) if cov_at_bottom else cov.permute(3, 4, 0, 1, 2)
        return cov.reshape((-1,) + tuple(cov.shape[-2:]))

    r = moments_fn.size
    if cov_at_bottom:
        moments_qtype = quantity.qtype.replace_scalar(qt.ArrayType(shape=(r, r), qtype=qt.ScalarType()))
    else:
        moments_qtype = qt.ArrayType(shape=(r, r), qtype=quantity.qtype)
    return _MomentsQuantity(moments_qtype, eval_cov, quantity, moments_fn, cov_at_bottom, "covariance")


# ----------------------------------------------------------------------------------------------------------
class _Plan:
    """How one ``estimate_mean`` call is executed on the device."""

    def __init__(self, quantity, variance=True):
        self.quantity = quantity
        self.kind = "raw"
        self.ext_fn = self.table = None
        self.inner = quantity
        self.fn = None
        self.at_bottom = True
        if isinstance(quantity, _MomentsQuantity):
            fn = quantity._moments_fn
            inner = quantity._input_quantities[0]
            scalar = inner.size() == 1
            transformed = isinstance(fn, TransformedMoments)
            # accumulator columns of the fused kernel: 113 moments per thread, 226 with lane-pair columns (scalar
            # quantities only); larger bases take the generic path below (moments evaluated per row slice)
            base_ok = fn.base_moments().size <= (226 if scalar else 113)
            if quantity._fused_kind == "moments" and base_ok and not transformed:
                self.kind = "moments"
            elif quantity._fused_kind == "moments" and base_ok and transformed and fn.base_moments().size <= 104 \
                    and self._fits(inner.size() * fn.base_moments().size ** 2):
                self.kind = "transformed"
            elif quantity._fused_kind == "covariance" and not variance and self._linearise(fn, scalar):
                self.kind = "cov_linear"
            elif quantity._fused_kind == "covariance" and not transformed and fn.size <= 104 \
                    and self._fits(inner.size() * fn.size ** 2):
                self.kind = "covariance"
            if self.kind != "raw":
                self.inner, self.fn, self.at_bottom = inner, fn, quantity._at_bottom


    @staticmethod
    def _fits(n_sums, limit=1 << 27):
        """The fused covariance route keeps ``[L, 2 + 2 K]`` sums (K = M R^2) on the device: up to 1 GiB per level."""
        return n_sums <= limit

    def _linearise(self, fn, scalar):
        """Covariance MEANS from the moment sums of the extended basis (``Moments.product_table``) if that basis fits
        the fused moments kernel."""
        table = fn.product_table()
        if table is None or table[0].size > (226 if scalar else 113):
            return False
        self.ext_fn, self.table = table
        return True


def _product_table_on(fn, device):
    """``C_t [E, R*R]`` of ``fn.product_table()`` as a CUDA tensor, cached on the moments object."""
    hit = getattr(fn, "_product_table_dev", None)
    if hit is None or hit.device != device:
        hit = torch.from_numpy(fn.product_table()[1]).to(device)
        fn._product_table_dev = hit
    return hit


def _level_row_ranges(storage, level_ids):
    """Row range of every level handled by this rank (contiguous shards)."""
    n_collected = storage.get_n_collected()
    return {l: _dist.shard_range(int(n_collected[l])) for l in level_ids}


_RAW_SLICE_BYTES = 1 << 30


def _row_slices(chunks, plan):
    """The generic path materialises the whole quantity ([K, n, 2] doubles) per chunk: cut chunks into row slices of
    at most ~1 GB of evaluated values (a covariance of 130 moments is 270 kB per sample)."""
    if plan.kind != "raw":
        yield from chunks
        return
    per_row = 16 * max(1, plan.quantity.size())
    step = max(1, _RAW_SLICE_BYTES // per_row)
    for level_id, rows in chunks:
        if rows.shape[0] <= step:
            yield level_id, rows
        else:
            for start in range(0, rows.shape[0], step):
                yield level_id, rows[start:start + step]


def _basis_key(basis):
    return (basis.kind, basis.size, basis.is_log, basis.is_clip, basis.shift, basis.scale, basis.ref_lo, basis.ref_hi)


def _resident_fast_path(quantity, plan, storage, level_ids, n_levels, device):
    """Small-call latency: ``moments`` of a quantity that is a VIEW of storage rows (column selections of the root
    quantity), every level resident in HBM, single process.  The whole estimate -- zeroing, one fused launch per level,
    finalize, D2H, synchronise -- is then ONE native call on a cached ``_native.EstimatePlan``
    (``mlmcb200_estimate_moments_levels``); without it the ~10 host calls of the generic route cost several times the
    kernels of a small estimate (cfg1: 1e5 samples).  Returns the packed host result or None (generic route)."""
    if plan.kind != "moments" or _dist.world_size() > 1:
        return None
    basis = plan.fn.basis_struct()
    m = plan.inner.size()
    if m > 128 or (m > 1 and basis.size > 112) or basis.size > 226:
        return None
    rows_of = {}
    for level_id in level_ids:
        if storage.level_n_rows(level_id) == 0:
            continue
        rows = storage.device_rows(level_id, device, True)
        if rows is None:
            return None                                   # does not fit / must be streamed
        rows_of[level_id] = rows
    if not rows_of:
        return None
    cache = plan.inner.__dict__.setdefault("_estimate_plans", {})
    key = _basis_key(basis)
    hit = cache.get(key)
    if hit is None and key in cache:
        return None                                       # known not to be a view of the storage rows
    if hit is not None:
        ep, row_keys = hit
        if row_keys == {l: (r.data_ptr(), r.shape[0]) for l, r in rows_of.items()}:
            return ep, ep.run()
    views = {}
    for level_id, rows in rows_of.items():
        x = plan.inner.device_samples(q_mod.DeviceChunk(level_id, rows, 0))
        if x.dtype != torch.float64 or x.untyped_storage().data_ptr() != rows.untyped_storage().data_ptr():
            cache[key] = None                             # a computed quantity: its chunks are new tensors every time
            return None
        views[level_id] = x
    try:
        ep = _native.EstimatePlan(basis, views, n_levels)
    except _native.NativeError:
        return None
    if len(cache) >= 8:
        cache.clear()
    cache[key] = (ep, {l: (r.data_ptr(), r.shape[0]) for l, r in rows_of.items()})
    return ep, ep.run()


def estimate_mean(quantity, variance=True):
    """MLMC mean estimator (quantity_estimate.py:22-80) -> ``QuantityMean``.

    ``variance=False`` (extension): the caller needs ``mean`` / ``l_means`` only.  For a ``covariance`` quantity this
    selects the linearised path (table above) -- the reference's own ``construct_density`` reads nothing but the
    covariance means (estimator.py:311-312); ``l_vars`` / ``var`` of the result are then NaN."""
    cache_clear()
    storage_q = quantity.get_quantity_storage()
    storage = storage_q._storage
    level_ids = storage_q.level_ids()
    n_levels = int(np.max(level_ids)) + 1
    device = q_mod._device()
    plan = _Plan(quantity, variance)
    multi = _dist.world_size() > 1
    sharded = multi and not getattr(storage, "rows_are_local_shard", False)
    ranges = _level_row_ranges(storage, level_ids) if sharded else None

    fast = _resident_fast_path(quantity, plan, storage, level_ids, n_levels, device) \
        if getattr(storage, "resident_fraction", 0) > 0 else None
    if fast is not None:
        ep, packed = fast
        return _quantity_mean_from_packed(quantity, plan, np.array(packed), ep.L, ep.K)   # the plan's buffer is reused

    acc = None          # LevelAccumulator of the main statistics
    gram = None         # transformed moments: Gram of the base differences
    base_basis = None
    chunk_id = 0
    for level_id, rows in _row_slices(storage.device_chunks(level_ids, device, keep_resident=not sharded,
                                                            row_ranges=ranges if sharded else None), plan):
        if rows.shape[0] == 0:
            continue
        x = plan.inner.device_samples(q_mod.DeviceChunk(level_id, rows, chunk_id))
        chunk_id += 1
        if plan.kind == "moments":
            basis = plan.fn.basis_struct()
            if acc is None:
                acc = _native.LevelAccumulator(n_levels, x.shape[0] * basis.size, device)
            _native.moments_accumulate(basis, x, acc.level(level_id))
        elif plan.kind == "transformed":
            base_basis = plan.fn.basis_struct()
            r0 = base_basis.size
            if acc is None:
                acc = _native.LevelAccumulator(n_levels, x.shape[0] * r0, device)
                gram = _native.LevelAccumulator(n_levels, x.shape[0] * r0 * r0, device)
            valid = _native.sample_mask(base_basis, x) if x.shape[0] > 1 else None
            _native.moments_accumulate(base_basis, x, acc.level(level_id), valid=valid)
            _native.gram_accumulate(base_basis, x, gram.level(level_id), mode=1 if x.shape[2] == 2 else 0,
                                    want_var=False, valid=valid)
        elif plan.kind == "cov_linear":
            basis = plan.ext_fn.basis_struct()
            if acc is None:
                acc = _native.LevelAccumulator(n_levels, x.shape[0] * basis.size, device)
            _native.moments_accumulate(basis, x, acc.level(level_id), sums_only=True)
        elif plan.kind == "covariance":
            basis = plan.fn.basis_struct()
            if acc is None:
                acc = _native.LevelAccumulator(n_levels, x.shape[0] * basis.size * basis.size, device)
            _native.gram_accumulate(basis, x, acc.level(level_id), mode=0, want_var=True)
        else:
            if x.dtype == torch.bool:
                # the reference fails the same way: numpy refuses ``fine - coarse`` on boolean chunks
                raise TypeError("estimate_mean of a boolean quantity (use it in select(), or cast it)")
            x = x if x.dtype == torch.float64 else x.to(torch.float64)
            if acc is None:
                acc = _native.LevelAccumulator(n_levels, x.shape[0], device)
            _native.moments_accumulate(_native.RAW_BASIS, x, acc.level(level_id))

    if acc is None and multi:
        # A rank whose shard holds no rows still has to enter the all-reduce with zeros of the right width.
        if plan.kind == "moments":
            width = plan.inner.size() * plan.fn.size
        elif plan.kind == "transformed":
            r0 = plan.fn.base_moments().size
            width = plan.inner.size() * r0
            gram = _native.LevelAccumulator(n_levels, width * r0, device)
        elif plan.kind == "cov_linear":
            width = plan.inner.size() * plan.ext_fn.size
        elif plan.kind == "covariance":
            width = plan.inner.size() * plan.fn.size * plan.fn.size
        else:
            width = quantity.size()
        acc = _native.LevelAccumulator(n_levels, width, device)
    if acc is None:
        raise Exception("All samples were masked")
    peer = None
    if multi:
        # scalar-sized accumulators: the sum over the ranks rides in the finalize launch (NVLink peer memory);
        # anything else: one NCCL all-reduce
        peer = _dist.peer_state(acc.acc.numel()) if plan.kind not in ("transformed", "cov_linear") else None
        if peer is None:
            _dist.all_reduce_sum(acc.acc)
            if gram is not None:
                _dist.all_reduce_sum(gram.acc)

    if plan.kind == "transformed":
        acc = _transform_sums(acc, gram, plan.fn, device)
    elif plan.kind == "cov_linear":
        acc = _native.level_sums_transform(acc, plan.inner.size(), _product_table_on(plan.fn, device))
    out = acc.finalize(peer=peer)
    parts = [out["packed"].reshape(-1), acc.acc[:, :2].reshape(-1)]
    if peer is not None:
        parts.append(out["status"])                      # rides in the same copy
    packed = _to_host(torch.cat(parts))                                              # single D2H copy
    if peer is not None:
        if packed[-1] != 0.0:
            # The fused reduce gave up waiting for a peer (skewed ranks).  A rank that gives up poisons the epoch for
            # the others, so every rank lands here; the local sums are intact: reduce them with NCCL and finalize.
            _dist.note_peer_fallback()
            _dist.all_reduce_sum(acc.acc)
            out = acc.finalize()
            packed = _to_host(torch.cat([out["packed"].reshape(-1), acc.acc[:, :2].reshape(-1)]))
        else:
            packed = packed[:-1]
    return _quantity_mean_from_packed(quantity, plan, packed, acc.n_levels, acc.K)


def _quantity_mean_from_packed(quantity, plan, packed, L, K):
    """``QuantityMean`` from the packed host result [l_means | l_vars | mean | var | (n, n_rm) per level]."""
    l_means = packed[:L * K].reshape(L, K)
    l_vars = packed[L * K:2 * L * K].reshape(L, K)
    mean, var = packed[2 * L * K:(2 * L + 1) * K], packed[(2 * L + 1) * K:(2 * L + 2) * K]
    counts = packed[(2 * L + 2) * K:].reshape(L, 2)
    n_samples = [int(c) for c in counts[:, 0]]
    n_rm_samples = [int(c) for c in counts[:, 1]]
    if sum(n_samples) == 0:
        raise Exception("All samples were masked")
    if plan.kind in ("moments", "transformed") and not plan.at_bottom:
        r = plan.fn.size
        l_means = l_means.reshape(L, -1, r).transpose(0, 2, 1).reshape(L, K)
        l_vars = l_vars.reshape(L, -1, r).transpose(0, 2, 1).reshape(L, K)
        mean = mean.reshape(-1, r).T.reshape(K)
        var = var.reshape(-1, r).T.reshape(K)
    elif plan.kind in ("cov_linear", "covariance") and not plan.at_bottom:
        rr = plan.fn.size ** 2
        l_means = l_means.reshape(L, -1, rr).transpose(0, 2, 1).reshape(L, K)
        l_vars = l_vars.reshape(L, -1, rr).transpose(0, 2, 1).reshape(L, K)
        mean = mean.reshape(-1, rr).T.reshape(K)
        var = var.reshape(-1, rr).T.reshape(K)
    return q_mod.QuantityMean(quantity.qtype, l_means=l_means, l_vars=l_vars, n_samples=n_samples,
                              n_rm_samples=n_rm_samples, mean=mean, var=var)


_WEIGHTED_BLOCK_ROWS = 131072        # rows per row block of the multiplicity histogram (<= counts_block_rows())


def _weighted_bootstrap_applies(method, basis, x, sizes, n_chunk, expected_draws, n_replicates):
    """One-pass weighted sums (``mlmcb200_moments_accumulate_weighted``) or per-replicate gather
    (``mlmcb200_moments_accumulate_resampled``)?  The weighted pass costs ~ n_chunk x (replicates rounded up to 8), the
    gather ~ 3.6 x the draws (per moment): the pass wins when the replicates draw about as many rows as the chunk holds.
    The choice depends on the EXPECTED draws of a replicate in this chunk and on the total number of replicates only
    (not on this rank's share of them), so that a seed gives the same replicates for every world size.
    ``method``: None (choose), "weighted" (wherever the kernel applies), "gather"."""
    method = method or os.environ.get("MLMCB200_BOOTSTRAP") or None
    if method == "gather":
        return False
    if method not in (None, "weighted"):
        raise ValueError("bootstrap method must be None, 'weighted' or 'gather'")
    M, n, S = x.shape
    sn, ss = x.stride(1), x.stride(2)
    ok = (basis.kind in (_native.LEGENDRE, _native.MONOMIAL, _native.FOURIER) and basis.size <= _native.weighted_max_size() and M == 1 and n_chunk >= 1
          and (S == 1 or (sn == 2 and ss == 1 and x.data_ptr() % 16 == 0))
          and expected_draws <= 6 * n_chunk and 0 < int(sizes.max()) <= 8 * n_chunk)
    if not ok or method == "weighted":
        return ok
    return expected_draws * n_replicates >= 0.35 * n_chunk * 8 * (-(-n_replicates // 8))


def bootstrap_moments(quantity, moments_fn, sample_vector, n_subsamples, seed=None, mom_at_bottom=False,
                      return_indices=False, method=None):
    """All ``n_subsamples`` bootstrap replicates of ``estimate_mean(moments(quantity.subsample(sample_vector)))``
    (the loop body of ``Estimate.est_bootstrap``, mlmc/estimator.py:185-193) fused on the device.

    Replicate b draws, per level l, ``sample_vector[l]`` rows WITH replacement (``Quantity.pick_samples``,
    mlmc/quantity/quantity.py:307-322: a level held in several chunks is split chunk by chunk with hypergeometric
    sizes, rows are then chosen with replacement inside each chunk) and re-estimates the level sums.  The row
    numbers are drawn on the device (``mlmcb200_resample_indices``, counter-based Philox keyed by the seed); one
    launch per level handles all replicates (``mlmcb200_moments_accumulate_resampled``): samples are gathered through L2, never copied.

    Returns a dict of NumPy arrays ``mean, var [B, K]``, ``l_means, l_vars [B, L, K]``, ``n_samples [B, L]``
    (+ ``indices``: per level a list of (row offset, CUDA int32 tensor [B, draws]) when asked for -- the test hook
    that lets a CPU check repeat the estimate on exactly the same rows)."""
    import scipy.stats
    storage_q = quantity.get_quantity_storage()
    storage = storage_q._storage
    level_ids = storage_q.level_ids()
    n_levels = int(np.max(level_ids)) + 1
    device = q_mod._device()
    basis = moments_fn.basis_struct()
    n_collected = [int(n) for n in storage.get_n_collected()]
    # Multi-GPU: replicates are independent units -> rank r computes the replicates [r B / P, (r + 1) B / P) on the whole
    # levels and one all-gather of the small per-replicate results follows (no data-path collective).  Needs the same
    # seed on every rank.  (A storage of rank-local shards is bootstrapped locally, rank by rank.)
    n_total = int(n_subsamples)
    shard_replicates = _dist.world_size() > 1 and not getattr(storage, "rows_are_local_shard", False)
    if shard_replicates and seed is None:
        raise ValueError("multi-GPU bootstrap: pass the same `seed` on every rank")
    b_lo, b_hi = _dist.shard_range(n_total) if shard_replicates else (0, n_total)
    B = b_hi - b_lo
    seed = int(seed) if seed is not None else int(np.random.SeedSequence().entropy % (1 << 62))
    # Every random choice is keyed by (seed, GLOBAL replicate number): the host draws (hypergeometric chunk sizes,
    # multinomial block counts) by one generator per replicate, the row numbers by Philox with the replicate number in
    # the counter -- so the replicates do not depend on the world size or on how they are grouped into launches.
    host_rngs = [np.random.default_rng([seed, b_lo + b]) for b in range(B)]
    if sum(n_collected[l] for l in level_ids) == 0:                  # the same on every rank: no rank is left behind
        raise Exception("All samples were masked")

    width = 2 + 2 * quantity.size() * basis.size
    acc = torch.zeros((B, n_levels, width), dtype=torch.float64, device=device)
    seen_rows = False
    remaining = {l: (np.full(B, int(sample_vector[l]), dtype=np.int64), n_collected[l]) for l in level_ids}
    indices = {l: [] for l in level_ids}
    offsets = {l: 0 for l in level_ids}
    chunk_id = 0
    max_draws = 1 << 30                                               # row numbers per launch (4 GB of int32)
    for level_id, rows in (storage.device_chunks(level_ids, device, keep_resident=True) if B > 0 else ()):
        n_chunk = int(rows.shape[0])
        if n_chunk == 0:
            continue
        x = quantity.device_samples(q_mod.DeviceChunk(level_id, rows, chunk_id))
        chunk_id += 1
        seen_rows = True
        k_left, n_left = remaining[level_id]
        if n_chunk >= n_left:                                          # last (or only) chunk takes what is left
            sizes = k_left.copy()
        else:
            sizes = np.array([scipy.stats.hypergeom(n_left, int(k), n_chunk).rvs(random_state=host_rngs[b]) if k > 0
                              else 0 for b, k in enumerate(k_left)], dtype=np.int64)
        remaining[level_id] = (k_left - sizes, n_left - n_chunk)
        valid = _native.sample_mask(basis, x) if x.shape[0] > 1 else None
        level_acc = acc[:, level_id]
        stream_id = (level_id << 32) | chunk_id
        row_bytes = 8 * x.shape[0] * x.shape[2]
        # draws are listed by row block (<= 16 MB of rows each) so that the CTAs working on a replicate gather from
        # a window that stays in L2; the per-block counts are multinomial, which keeps the draws i.i.d. uniform
        n_blocks = int(min(64, max(1, -(-n_chunk * row_bytes // (16 << 20)))))

        def draw(b0, b1, k):
            cum = None
            if n_blocks > 1 and k >= 4096 * n_blocks:
                edges = (np.arange(n_blocks + 1, dtype=np.int64) * n_chunk) // n_blocks
                p_block = np.diff(edges) / n_chunk
                counts = np.stack([host_rngs[b].multinomial(k, p_block) for b in range(b0, b1)])
                cum_h = np.zeros((b1 - b0, n_blocks + 1), dtype=np.int64)
                np.cumsum(counts, axis=1, out=cum_h[:, 1:])
                cum = torch.from_numpy(cum_h).to(device)
            return _native.resample_indices(seed, stream_id, n_chunk, k, b1 - b0, device, block_cum=cum,
                                            rep_offset=b_lo + b0)

        expected = float(sample_vector[level_id]) * n_chunk / max(1, n_collected[level_id])
        if _weighted_bootstrap_applies(method, basis, x, sizes, n_chunk, expected, n_total):
            # ONE pass over the rows for all replicates: multiplicities (the same Philox draws, histogrammed) times the
            # moment differences on FP64 tensor tiles (csrc/bootstrap.cu)
            n_wblocks = max(1, -(-n_chunk // _WEIGHTED_BLOCK_ROWS))
            edges = (np.arange(n_wblocks + 1, dtype=np.int64) * n_chunk) // n_wblocks
            p_block = np.diff(edges) / n_chunk
            cum_h = np.zeros((B, n_wblocks + 1), dtype=np.int64)
            for b in range(B):
                if n_wblocks > 1:
                    np.cumsum(host_rngs[b].multinomial(int(sizes[b]), p_block), out=cum_h[b, 1:])
                else:
                    cum_h[b, 1] = int(sizes[b])
            stride = -(-n_chunk // 16) * 16
            # multiplicities of as many replicates at a time as a quarter of the free device memory holds (>= 2 GB), in
            # whole groups of the kernel (104 replicates) where possible: few replicates per pass waste the tensor tiles
            budget = 2 << 30
            if B * stride > budget:                                           # (the query synchronises: only when needed)
                budget = max(budget, torch.cuda.mem_get_info(device)[0] // 4)
            group = min(B, max(8, (budget // stride) // 8 * 8))
            if group < B and group > 104:
                group = group // 104 * 104
            for b0 in range(0, B, group):
                b1 = min(B, b0 + group)
                cum = torch.from_numpy(cum_h[b0:b1]).to(device)
                counts = _native.resample_counts(seed, stream_id, n_chunk, cum, int(sizes[b0:b1].max()), device,
                                                 rep_offset=b_lo + b0)
                _native.moments_accumulate_weighted(basis, x, counts, level_acc[b0:b1])
                del counts
            if return_indices:                                                 # test hook: the rows behind the counts
                for b in range(B):
                    if sizes[b] > 0:
                        cum = torch.from_numpy(cum_h[b:b + 1]).to(device) if n_wblocks > 1 else None
                        idx = _native.resample_indices(seed, stream_id, n_chunk, int(sizes[b]), 1, device,
                                                       block_cum=cum, rep_offset=b_lo + b)
                        indices[level_id].append((offsets[level_id], b_lo + b, idx))
        elif np.all(sizes == sizes[0]):
            k = int(sizes[0])
            if k > 0:
                group = max(1, min(B, max_draws // k))
                for b0 in range(0, B, group):
                    b1 = min(B, b0 + group)
                    idx = draw(b0, b1, k)
                    _native.moments_accumulate_resampled(basis, x, idx, level_acc[b0:b1], valid=valid)
                    if return_indices:
                        indices[level_id].append((offsets[level_id], b_lo + b0, idx))
        else:                                                          # ragged draws: one replicate per launch
            for b in range(B):
                k = int(sizes[b])
                if k == 0:
                    continue
                idx = draw(b, b + 1, k)
                _native.moments_accumulate_resampled(basis, x, idx, level_acc[b:b + 1], valid=valid)
                if return_indices:
                    indices[level_id].append((offsets[level_id], b_lo + b, idx))
        offsets[level_id] += n_chunk

    L, K = n_levels, (width - 2) // 2
    packed = torch.cat([_native.finalize_levels_batched(acc), acc[:, :, 0].reshape(B, L)], dim=1) if B > 0 else \
        torch.empty((0, 2 * L * K + 2 * K + L), dtype=torch.float64, device=device)
    if shard_replicates:
        import torch.distributed as td
        world = _dist.world_size()
        b_max = -(-n_total // world)                                   # equal-sized pieces for the all-gather
        piece = torch.zeros((b_max, packed.shape[1]), dtype=torch.float64, device=device)
        piece[:B] = packed
        gathered = torch.empty((world * b_max, packed.shape[1]), dtype=torch.float64, device=device)
        td.all_gather_into_tensor(gathered, piece, group=_dist._state["group"])
        keep = [r * b_max + i for r in range(world) for i in range(_dist.shard_range(n_total, r, world)[1]
                                                                  - _dist.shard_range(n_total, r, world)[0])]
        packed = gathered[torch.tensor(keep, device=device)]
        B = n_total
    host = _to_host(packed.reshape(-1)).reshape(B, -1)
    l_means = host[:, :L * K].reshape(B, L, K)
    l_vars = host[:, L * K:2 * L * K].reshape(B, L, K)
    mean, var = host[:, 2 * L * K:2 * L * K + K], host[:, 2 * L * K + K:2 * L * K + 2 * K]
    n_samples = host[:, 2 * L * K + 2 * K:].astype(np.int64)
    if not mom_at_bottom:
        r = moments_fn.size
        l_means = l_means.reshape(B, L, -1, r).transpose(0, 1, 3, 2).reshape(B, L, K)
        l_vars = l_vars.reshape(B, L, -1, r).transpose(0, 1, 3, 2).reshape(B, L, K)
        mean = mean.reshape(B, -1, r).transpose(0, 2, 1).reshape(B, K)
        var = var.reshape(B, -1, r).transpose(0, 2, 1).reshape(B, K)
    out = {"mean": np.array(mean), "var": np.array(var), "l_means": np.array(l_means), "l_vars": np.array(l_vars),
           "n_samples": n_samples}
    if return_indices:
        out["indices"] = indices
    return out


def can_fuse_bootstrap(quantity, moments_fn):
    """The fused path covers plain (non-transformed) bases up to the kernel's size limit."""
    limit = 226 if quantity.size() == 1 else 113
    return (not isinstance(moments_fn, TransformedMoments) and moments_fn.size <= limit
            and quantity.get_quantity_storage() is not None)


_host_buffers = {}


def _to_host(tensor):
    """Device -> host into pinned memory.  Small results are copied out of one cached staging buffer.  Large ones
    (vector quantities: tens of MB) are returned as views of pooled pinned buffers, so that no call pays for fresh
    pages; a pooled buffer is handed out again only when nothing references the arrays of the call that used it."""
    import sys
    n = tensor.numel()
    if n * 8 > (1 << 20):
        pool = _host_buffers.setdefault(("pool", n), [])
        for host, arr in pool:
            if sys.getrefcount(arr) <= 3:            # the pool tuple, the loop variable and getrefcount's argument
                break
        else:
            host = torch.empty(n, dtype=torch.float64).pin_memory()
            arr = host.numpy()
            if len(pool) < 4:
                pool.append((host, arr))
        host.copy_(tensor, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return arr
    buf = _host_buffers.get("f64")
    if buf is None or buf.numel() < n:
        buf = torch.empty(1 << 17, dtype=torch.float64).pin_memory()
        _host_buffers["f64"] = buf
    view = buf[:n]
    view.copy_(tensor, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return view.numpy().copy()


def _transform_sums(acc, gram, fn, device):
    """Level sums of ``L phi`` from the sums / Gram of the base differences, per component: ``sum d' = L sum d`` and
    ``sum d'_k^2 = L_k G L_k^T`` (moments.py:256-259 applied under the sums)."""
    l_mat = fn._matrix_on(device)                         # [R1, R0]
    r1, r0 = l_mat.shape
    n_comp = acc.K // r0
    out = _native.LevelAccumulator(acc.n_levels, n_comp * r1, device)
    out.acc[:, :2] = acc.acc[:, :2]
    sums = acc.acc[:, 2:2 + n_comp * r0].reshape(-1, n_comp, r0)
    out.acc[:, 2:2 + n_comp * r1] = (sums @ l_mat.T).reshape(-1, n_comp * r1)
    g = gram.acc[:, 2:2 + n_comp * r0 * r0].reshape(-1, n_comp, r0, r0)
    out.acc[:, 2 + n_comp * r1:] = torch.einsum("ki,lmij,kj->lmk", l_mat, g, l_mat).reshape(-1, n_comp * r1)
    return out
