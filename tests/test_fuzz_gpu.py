"""Seeded random sweep of shapes and options through the C ABI against the oracle: basis kind, number of moments,
number of components, level sizes, chunking, log / clipped domains, NaN and out-of-domain samples.  Every case must meet
the north_star tolerances (counts exact, level means / variances rel <= 1e-10, covariances <= 1e-8)."""
import numpy as np
import pytest
import torch

from oracle import mlmc_oracle as orc
from test_kernels_gpu import dev, native, rel_close, run_gram, run_moments, to_struct

pytestmark = pytest.mark.gpu


def _random_levels(rng, n_levels, sizes, n_comp, log, nan_frac, out_frac):
    steps = orc.level_steps(n_levels, (0.4, 0.02)) if n_levels > 1 else [0.3]
    levels = []
    for l, n in enumerate(sizes):
        x = rng.lognormal(0.0, 0.4, size=n) if log else rng.normal(size=n)
        rows = orc.synth_level_rows(x, steps[l], steps[l - 1] if l else None)          # [n, 2, 1]
        if n_comp > 1:
            rows = np.repeat(rows, n_comp, axis=2) * (1.0 + 0.01 * rng.normal(size=(1, 1, n_comp)))
        if log:
            rows = np.abs(rows) + 1e-6
        if l == 0:
            rows[:, 1, :] = 0
        if n > 4:
            k = rng.integers(0, n, size=max(1, int(nan_frac * n)))
            rows[k, 0, rng.integers(0, n_comp)] = np.nan
            k = rng.integers(0, n, size=max(1, int(out_frac * n)))
            rows[k, 1 if l else 0, rng.integers(0, n_comp)] = 1e7
        levels.append(rows)
    return levels


@pytest.mark.parametrize("seed", range(24))
def test_random_moment_cases(seed):
    rng = np.random.default_rng(1000 + seed)
    kind = ["legendre", "legendre", "monomial", "fourier"][seed % 4]
    size = int(rng.integers(1, {"legendre": 114, "monomial": 12, "fourier": 40}[kind]))
    n_comp = int(rng.choice([1, 1, 1, 2, 3, 7, 40, 129, 260]))
    if n_comp > 1:
        size = min(size, 24)
    if n_comp == 1 and kind == "legendre" and seed % 8 == 4:
        size = int(rng.integers(114, 227))                       # lane-pair accumulator columns
    n_levels = int(rng.integers(1, 5))
    sizes = [int(rng.choice([1, 2, 5, 31, 128, 129, 1000, 4097, 20000])) for _ in range(n_levels)]
    if n_comp > 16:
        sizes = [min(s, 1000) for s in sizes]
    log = bool(seed % 5 == 3) and kind != "fourier"
    safe = bool(seed % 3 != 2)
    dom = (0.2, 4.0) if log else (-2.5, 2.5) if safe else (-30.0, 30.0)
    levels = _random_levels(rng, n_levels, sizes, n_comp, log, 0.01, 0.01 if safe else 0.0)
    if not safe:                                                 # no clipping: keep the values inside the table's range
        for lv in levels:
            lv[np.abs(lv) > 1e6] = 1.0
    b = orc.Basis(kind, size, dom, log=log, safe_eval=safe)
    chunk = int(rng.choice([0, 64, 777, 5000])) or None
    with np.errstate(all="ignore"):
        want = orc.estimate_moments(levels, b)
    res = run_moments(to_struct(b), levels, chunk)
    tag = (seed, kind, size, n_comp, sizes, log, safe, chunk)
    assert np.array_equal(res["n"], want.n_samples), tag
    assert np.array_equal(res["n_rm"], want.n_rm_samples), tag
    fin = np.isfinite(want.l_means) & (want.n_samples[:, None] > 0)
    rel_close(np.where(fin, res["l_means"], 0.0), np.where(fin, want.l_means, 0.0), rtol=1e-10, atol_scale=1e-13,
              per_level=True)
    fin = np.isfinite(want.l_vars)
    assert np.array_equal(np.isinf(res["l_vars"]), np.isinf(want.l_vars)), tag
    # The reference's variance formula (sp - s^2 / n) / (n - 1) turns a rounding error of k eps in sp into k eps mean^2
    # in the variance, whatever the variance is (tests/test_conditioning_gpu.py measures it against an
    # extended-precision result): entries with |mean| >> std get that allowance (450 eps mean^2), nothing else does.
    cond = 1e-13 * np.where(np.isfinite(want.l_means), want.l_means, 0.0) ** 2
    rel_close(np.where(fin, res["l_vars"], 0.0), np.where(fin, want.l_vars, 0.0), rtol=1e-10, atol_scale=1e-13,
              per_level=True, extra_atol=cond)


@pytest.mark.parametrize("seed", range(10))
def test_random_covariance_cases(seed):
    rng = np.random.default_rng(2000 + seed)
    kind = ["legendre", "legendre", "monomial", "fourier", "legendre"][seed % 5]
    size = int(rng.integers(1, {"legendre": 105, "monomial": 10, "fourier": 33}[kind]))
    n_levels = int(rng.integers(1, 4))
    sizes = [int(rng.choice([1, 3, 63, 64, 65, 127, 128, 129, 1000, 3001])) for _ in range(n_levels)]
    levels = _random_levels(rng, n_levels, sizes, 1, False, 0.01, 0.01)
    b = orc.Basis(kind, size, (-2.5, 2.5))
    chunk = int(rng.choice([0, 100, 1000])) or None
    with np.errstate(all="ignore"):
        want = orc.estimate_covariance(levels, b)
    res = run_gram(to_struct(b), levels, want_var=True, chunk_rows=chunk)
    tag = (seed, kind, size, sizes, chunk)
    assert np.array_equal(res["n"], want.n_samples) and np.array_equal(res["n_rm"], want.n_rm_samples), tag
    ok = want.n_samples > 0
    rel_close(res["l_means"][ok], want.l_means[ok], rtol=1e-8, atol_scale=1e-12, per_level=True)
    fin = np.isfinite(want.l_vars)
    assert np.array_equal(np.isinf(res["l_vars"]), np.isinf(want.l_vars)), tag
    rel_close(np.where(fin, res["l_vars"], 0.0), np.where(fin, want.l_vars, 0.0), rtol=1e-8, atol_scale=1e-11,
              per_level=True)
    m = res["l_means"][ok].reshape(-1, size, size)
    assert np.array_equal(m, np.swapaxes(m, 1, 2))                # exactly symmetric
