"""Parity of the CUDA kernels (through the C ABI) with the oracle and the reference's golden outputs.

Tolerances (BASELINE.json north_star): moment means / variances rel <= 1e-10, covariance rel <= 1e-8,
max-ent pieces rel <= 1e-10 here (multipliers / pdf 1e-6 in test_api_gpu.py); sample counts exact.
"""
import numpy as np
import pytest
import torch

from oracle import mlmc_oracle as orc

pytestmark = pytest.mark.gpu

KINDS = {"raw": 0, "legendre": 1, "monomial": 2, "fourier": 3}


def dev():
    return torch.device("cuda:0")


def native():
    from mlmc_b200 import _native
    return _native


def to_struct(b: orc.Basis):
    return native().make_basis(KINDS[b.kind], b.size, b.domain, b.ref_domain, b.log, b.safe_eval)


def rel_close(got, want, rtol, atol_scale=1e-15, per_level=False, extra_atol=None):
    """|got - want| <= rtol |want| + atol_scale * scale.  ``scale`` is the largest finite |want| -- of the whole array,
    or, with ``per_level`` (arrays ``[L, K]``: one row per level), of each row separately: the variances of the fine
    levels are orders of magnitude below level 0's and must not hide behind its scale.  ``extra_atol`` (array like
    ``want``) adds an entry-wise absolute allowance."""
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    assert got.shape == want.shape, (got.shape, want.shape)
    finite = np.where(np.isfinite(want), np.abs(want), 0.0)
    if per_level and want.ndim >= 2:
        scale = finite.reshape(want.shape[0], -1).max(axis=1).reshape((-1,) + (1,) * (want.ndim - 1))
    else:
        scale = finite.max() if finite.size else 1.0
    tol = rtol * np.abs(want) + atol_scale * np.maximum(scale, 1e-300)
    if extra_atol is not None:
        tol = tol + np.where(np.isfinite(extra_atol), np.abs(extra_atol), 0.0)
    with np.errstate(invalid="ignore"):
        ok = (np.abs(got - want) <= tol) | (got == want) | (np.isnan(got) & np.isnan(want))
    assert ok.all(), "max abs err %.3e (scale %s) at %s" % (
        np.nanmax(np.abs(got - want)[~ok]), np.ravel(scale)[:8], np.argwhere(~ok)[:5].tolist())


def run_moments(basis, levels, chunk_rows=None):
    """levels: list of float64[N,2,M] -> dict of numpy arrays (l_means, l_vars, mean, var, n, n_rm)."""
    nat = native()
    M = levels[0].shape[2]
    acc = nat.LevelAccumulator(len(levels), M * basis.size, dev())
    for l, rows in enumerate(levels):
        d_rows = torch.from_numpy(np.ascontiguousarray(rows)).to(dev())
        step = chunk_rows or max(len(rows), 1)
        for start in range(0, len(rows), step):
            part = d_rows[start:start + step]
            x = part.permute(2, 0, 1)                 # [M, n, 2] view, storage strides
            if l == 0:
                x = x[:, :, :1]
            nat.moments_accumulate(basis, x, acc.level(l))
    res = {k: v.cpu().numpy() for k, v in acc.finalize().items()}
    a = acc.acc.cpu().numpy()
    res["n"], res["n_rm"] = a[:, 0].astype(np.int64), a[:, 1].astype(np.int64)
    return res


# ------------------------------------------------------------------------------------------------------
def test_library_loads_on_gpu():
    nat = native()
    assert nat.load().mlmcb200_abi_version() == nat.ABI_VERSION == 2
    assert nat.sm_count() > 0


def test_basis_tables_match_reference(golden):
    nat = native()
    g = golden("basis_tables")
    for i in range(int(g["n_cases"])):
        kind, size, dom, log, safe = g["case%02d_meta" % i]
        b = orc.Basis(str(kind), int(size), eval(str(dom)), log=bool(int(log)), safe_eval=bool(int(safe)))
        x = torch.from_numpy(g["case%02d_x" % i]).to(dev())
        got = nat.basis_eval(to_struct(b), x, b.size).cpu().numpy()
        want = g["case%02d_ref" % i]
        if b.kind == "fourier" or b.log:
            rel_close(got, want, rtol=1e-13, atol_scale=4e-15)
            assert np.array_equal(np.isnan(got), np.isnan(want))
        else:   # Legendre / Monomial follow numpy operation by operation: bit-identical
            assert np.array_equal(got, want, equal_nan=True), (kind, size)
    # TransformedMoments
    b = orc.Basis("legendre", 9, (-3.0, 5.0))
    mat = torch.from_numpy(g["transformed_matrix"]).to(dev())
    got = nat.basis_eval(to_struct(b), torch.from_numpy(g["transformed_x"]).to(dev()), mat.shape[0], mat).cpu().numpy()
    rel_close(got, g["transformed_ref"], rtol=1e-12, atol_scale=1e-14)
    # truncated evaluation (eval_all(value, size))
    got = nat.basis_eval(to_struct(b), torch.from_numpy(g["transformed_x"]).to(dev()), 4).cpu().numpy()
    assert np.array_equal(got, orc.basis_eval(b, g["transformed_x"], 4), equal_nan=True)


def test_reference_golden_means(golden):
    """test/test_sampling_pools.py:18,85-87: ref_means (atol 1e-5), means[0] == 1 and vars[0] == 0 exactly."""
    g = golden("sampling_pools")
    col = int(g["column"])
    levels = [g["rows%d" % l] for l in range(3)]
    basis = to_struct(orc.Basis("legendre", 5, tuple(g["domain"])))
    nat = native()
    acc = nat.LevelAccumulator(3, 5, dev())
    for l, rows in enumerate(levels):
        d_rows = torch.from_numpy(rows).to(dev())                       # [10, 2, 24]
        x = d_rows.permute(2, 0, 1)[col:col + 1]                        # strided column view
        if l == 0:
            x = x[:, :, :1]
        nat.moments_accumulate(basis, x, acc.level(l))
    out = acc.finalize()
    mean, var = out["mean"].cpu().numpy(), out["var"].cpu().numpy()
    assert mean[0] == 1.0 and var[0] == 0.0
    assert np.allclose(mean, g["ref_means"], atol=1e-5)
    rel_close(mean, g["means"], rtol=1e-10)
    rel_close(var, g["vars"], rtol=1e-10)


@pytest.mark.parametrize("tag,kind,size,dom,safe", [
    ("leg", "legendre", 12, None, True), ("mono", "monomial", 6, None, True),
    ("four", "fourier", 7, None, True), ("legraw", "legendre", 8, (-20.0, 20.0), False)])
@pytest.mark.parametrize("chunk_rows", [None, 512])
def test_moments_three_levels(golden, tag, kind, size, dom, safe, chunk_rows):
    g = golden("estimates")
    dom = tuple(g["A_domain"]) if dom is None else dom
    levels = [g["A_rows%d" % l] for l in range(3)]
    res = run_moments(to_struct(orc.Basis(kind, size, dom, safe_eval=safe)), levels, chunk_rows)
    assert np.array_equal(res["n"], g["A_%s_n" % tag]) and np.array_equal(res["n_rm"], g["A_%s_n_rm" % tag])
    rel_close(res["l_means"], g["A_%s_l_means" % tag], rtol=1e-10)
    rel_close(res["l_vars"], g["A_%s_l_vars" % tag], rtol=1e-10)
    rel_close(res["mean"], g["A_%s_mean" % tag], rtol=1e-10)
    rel_close(res["var"], g["A_%s_var" % tag], rtol=1e-10)
    assert res["mean"][0] == 1.0 and res["var"][0] == 0.0


def test_moments_log_domain(golden):
    g = golden("estimates")
    res = run_moments(to_struct(orc.Basis("legendre", 25, tuple(g["B_domain"]), log=True)), [g["B_rows0"]], 4096)
    assert np.array_equal(res["n"], g["B_n"]) and np.array_equal(res["n_rm"], g["B_n_rm"])
    rel_close(res["mean"], g["B_mean"], rtol=1e-10)
    rel_close(res["var"], g["B_var"], rtol=1e-10)


def test_moments_vector_quantity(golden):
    g = golden("estimates")
    levels = [g["C_rows%d" % l] for l in range(4)]
    res = run_moments(to_struct(orc.Basis("legendre", 5, tuple(g["C_domain"]))), levels, 256)
    assert np.array_equal(res["n"], g["C_bottom_n"]) and np.array_equal(res["n_rm"], g["C_bottom_n_rm"])
    rel_close(res["l_means"], g["C_bottom_l_means"].reshape(4, -1), rtol=1e-10)
    rel_close(res["l_vars"], g["C_bottom_l_vars"].reshape(4, -1), rtol=1e-10)
    # identity operation (plain estimate_mean of the quantity)
    raw = run_moments(native().RAW_BASIS, levels, 100)
    rel_close(raw["mean"], g["C_raw_mean"].reshape(-1), rtol=1e-10)
    rel_close(raw["var"], g["C_raw_var"].reshape(-1), rtol=1e-10)


@pytest.mark.parametrize("M", [130, 300])
def test_moments_wide_vector_vs_oracle(M):
    """More components than threads per CTA (component-tiled grid) against the oracle."""
    rng = np.random.default_rng(7)
    steps = orc.level_steps(2, (0.3, 0.03))
    levels = []
    for l, n in enumerate([257, 64]):
        rows = orc.synth_level_rows(rng.normal(size=n), steps[l], steps[l - 1] if l else None)
        rows = np.repeat(rows, M, axis=2) + np.arange(M)[None, None, :] * 1e-3
        if l == 0:
            rows[:, 1, :] = 0
        levels.append(rows)
    levels[1][3, 1, M - 1] = np.nan
    b = orc.Basis("fourier", 6, (-4.0, 4.5))
    want = orc.estimate_moments(levels, b)
    res = run_moments(to_struct(b), levels, 100)
    assert np.array_equal(res["n"], want.n_samples) and np.array_equal(res["n_rm"], want.n_rm_samples)
    rel_close(res["l_means"], want.l_means, rtol=1e-10)
    rel_close(res["l_vars"], want.l_vars, rtol=1e-10)


def run_gram(basis, levels, want_var=True, chunk_rows=None, mode=0):
    nat = native()
    R = basis.size
    acc = nat.LevelAccumulator(len(levels), R * R, dev())
    for l, rows in enumerate(levels):
        d_rows = torch.from_numpy(np.ascontiguousarray(rows)).to(dev())
        step = chunk_rows or max(len(rows), 1)
        for start in range(0, len(rows), step):
            x = d_rows[start:start + step].permute(2, 0, 1)
            if l == 0:
                x = x[:, :, :1]
            nat.gram_accumulate(basis, x, acc.level(l), mode=mode, want_var=want_var)
    res = {k: v.cpu().numpy() for k, v in acc.finalize().items()}
    a = acc.acc.cpu().numpy()
    res["n"], res["n_rm"] = a[:, 0].astype(np.int64), a[:, 1].astype(np.int64)
    return res


@pytest.mark.parametrize("chunk_rows", [None, 700])
def test_covariance_three_levels(golden, chunk_rows):
    g = golden("estimates")
    levels = [g["A_rows%d" % l] for l in range(3)]
    res = run_gram(to_struct(orc.Basis("legendre", 8, tuple(g["A_domain"]))), levels, True, chunk_rows)
    assert np.array_equal(res["n"], g["A_leg_n"])
    rel_close(res["mean"].reshape(8, 8), g["A_cov_mean"], rtol=1e-8)
    rel_close(res["var"].reshape(8, 8), g["A_cov_var"], rtol=1e-8)
    rel_close(res["l_means"].reshape(3, 8, 8), g["A_cov_l_means"], rtol=1e-8)
    rel_close(res["l_vars"].reshape(3, 8, 8), g["A_cov_l_vars"], rtol=1e-8)
    # structural identity of test/test_quantity_concept.py:613: first column of the covariance = moment means
    rel_close(res["mean"].reshape(8, 8)[:, 0], g["A_leg_mean"][:8], rtol=1e-10)


@pytest.mark.parametrize("kind,R", [("legendre", 25), ("legendre", 50), ("legendre", 100), ("monomial", 9),
                                    ("fourier", 32)])
def test_covariance_sizes_vs_oracle(kind, R):
    rng = np.random.default_rng(R)
    steps = orc.level_steps(2, (0.2, 0.02))
    levels = [orc.synth_level_rows(rng.normal(size=n), steps[l], steps[l - 1] if l else None)
              for l, n in enumerate([1500, 900])]
    b = orc.Basis(kind, R, (-3.0, 3.2))
    want = orc.estimate_covariance(levels, b, chunk_rows=256)
    res = run_gram(to_struct(b), levels, True, 1000)
    assert np.array_equal(res["n"], want.n_samples) and np.array_equal(res["n_rm"], want.n_rm_samples)
    rel_close(res["l_means"], want.l_means, rtol=1e-8, atol_scale=1e-13)
    rel_close(res["l_vars"], want.l_vars, rtol=1e-8, atol_scale=1e-13)
    mean_only = run_gram(to_struct(b), levels, False, 1000)       # pipelined variant: another summation order
    rel_close(mean_only["l_means"], res["l_means"], rtol=1e-11, atol_scale=1e-14)
    rel_close(mean_only["l_means"], want.l_means, rtol=1e-8, atol_scale=1e-13)


def test_difference_gram_vs_oracle():
    rng = np.random.default_rng(5)
    rows = orc.synth_level_rows(rng.normal(size=1200), 0.05, 0.3)
    b = orc.Basis("legendre", 14, (-3.0, 3.0))
    phi = orc.basis_eval(b, rows[:, :, 0])                      # [n, 2, R]
    good = ~np.isnan(phi).any(axis=(1, 2))
    d = phi[good, 0] - phi[good, 1]
    nat = native()
    acc = nat.LevelAccumulator(1, 14 * 14, dev())
    x = torch.from_numpy(rows).to(dev()).permute(2, 0, 1)
    nat.gram_accumulate(to_struct(b), x, acc.level(0), mode=1, want_var=False)
    a = acc.acc.cpu().numpy()[0]
    assert a[0] == good.sum() and a[1] == (~good).sum()
    rel_close(a[2:2 + 196].reshape(14, 14), d.T @ d, rtol=1e-9, atol_scale=1e-14)


@pytest.mark.parametrize("size", [5, 15, 25])
def test_maxent_pieces_match_reference(golden, size):
    g = golden("maxent")
    nat = native()
    domain = tuple(g["domain"])
    nodes, w = orc.gauss_panels(domain, int(g["n_panels"]))
    t = "R%d_" % size
    b = orc.Basis("legendre", size, domain, safe_eval=False)
    l_mat = torch.from_numpy(g[t + "L"]).to(dev())
    phi = nat.basis_eval(to_struct(b), torch.from_numpy(nodes).to(dev()), l_mat.shape[0], l_mat)
    lam = g[t + "lam"]
    out = nat.maxent_fgh(phi, torch.from_numpy(w).to(dev()), torch.from_numpy(lam).to(dev())).cpu().numpy()
    R = len(lam)
    mu = g[t + "mu"]
    f = np.sum(mu * lam) + out[0]
    grad = mu - out[1:1 + R]
    hess = out[1 + R:].reshape(R, R)
    rel_close(f, g[t + "F"], rtol=1e-10)
    rel_close(grad, g[t + "g"], rtol=1e-9, atol_scale=1e-12)
    rel_close(hess, g[t + "H"], rtol=1e-9, atol_scale=1e-13)
    assert np.array_equal(hess, hess.T)
    fg = nat.maxent_fgh(phi, torch.from_numpy(w).to(dev()), torch.from_numpy(lam).to(dev()), what=3).cpu().numpy()
    assert np.array_equal(fg[:1 + R], out[:1 + R])


def test_all_samples_masked_counts():
    nat = native()
    b = to_struct(orc.Basis("legendre", 4, (0.0, 1.0)))
    rows = np.full((40, 2, 1), 7.0)
    acc = nat.LevelAccumulator(1, 4, dev())
    nat.moments_accumulate(b, torch.from_numpy(rows).to(dev()).permute(2, 0, 1), acc.level(0))
    a = acc.acc.cpu().numpy()[0]
    assert a[0] == 0 and a[1] == 40 and np.all(a[2:] == 0)


def test_large_run_properties():
    """Size-independent properties at a size the oracle would not finish quickly: linearity in the sample set
    (two halves add up), determinism (bitwise repeatable) and moment 0 exactness."""
    nat = native()
    gen = torch.Generator(device=dev()).manual_seed(11)
    n = 3_000_000
    x = torch.randn(n, generator=gen, device=dev(), dtype=torch.float64)
    root = torch.sqrt(1e-4 + x.abs())
    rows = torch.stack([x + 0.05 * root, x + 0.5 * root], dim=1).unsqueeze(2).contiguous()     # [n, 2, 1]
    b = to_struct(orc.Basis("legendre", 50, (-3.7, 3.7)))

    def run(parts):
        acc = nat.LevelAccumulator(1, 50, dev())
        for p in parts:
            nat.moments_accumulate(b, p.permute(2, 0, 1), acc.level(0))
        return acc.acc.clone()

    whole = run([rows])
    again = run([rows])
    assert torch.equal(whole, again)
    halves = run([rows[: n // 2], rows[n // 2:]])
    assert torch.equal(whole[:, :2], halves[:, :2])
    assert torch.allclose(whole, halves, rtol=1e-11, atol=1e-9)
    a = whole.cpu().numpy()[0]
    assert a[0] + a[1] == n and a[2] == 0.0 and a[2 + 50] == 0.0
    # cross-check a 200k-sample slice against the oracle
    sl = rows[:200_000].cpu().numpy()
    want = orc.estimate_moments([np.zeros((1, 2, 1)), sl], orc.Basis("legendre", 50, (-3.7, 3.7)))
    acc = nat.LevelAccumulator(2, 50, dev())
    nat.moments_accumulate(b, rows[:200_000].permute(2, 0, 1), acc.level(1))
    nat.moments_accumulate(b, torch.zeros(1, 1, 1, dtype=torch.float64, device=dev()), acc.level(0))
    out = acc.finalize()
    rel_close(out["l_means"][1].cpu().numpy(), want.l_means[1], rtol=1e-10)
    rel_close(out["l_vars"][1].cpu().numpy(), want.l_vars[1], rtol=1e-10)


# ------------------------------------------------------------------------------------------------------
# edge cases: ragged / tiny / empty inputs, degenerate levels, size limits
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sizes", [[1, 3], [7, 1], [1025, 2, 1], [129, 128, 127]])
def test_ragged_and_tiny_levels_match_oracle(sizes):
    rng = np.random.default_rng(sum(sizes))
    steps = orc.level_steps(len(sizes), (0.4, 0.01)) if len(sizes) > 1 else [0.4]
    levels = [orc.synth_level_rows(rng.normal(size=n), steps[l], steps[l - 1] if l else None)
              for l, n in enumerate(sizes)]
    b = orc.Basis("legendre", 9, (-5.0, 5.0))
    with np.errstate(all="ignore"):
        want = orc.estimate_moments(levels, b)
    res = run_moments(to_struct(b), levels, 100)
    assert np.array_equal(res["n"], want.n_samples)
    rel_close(res["l_means"], want.l_means, rtol=1e-10)
    finite = np.isfinite(want.l_vars)
    assert np.array_equal(np.isinf(res["l_vars"]), np.isinf(want.l_vars))          # n == 1 -> +inf (quantity_estimate.py:77)
    rel_close(res["l_vars"][finite], want.l_vars[finite], rtol=1e-9, atol_scale=1e-13)


def test_fully_masked_level_behaves_like_reference():
    """One level loses every sample: the reference divides by n = 0 (NaN mean, inf variance) and goes on."""
    rng = np.random.default_rng(3)
    levels = [orc.synth_level_rows(rng.normal(size=50), 0.1, None), np.full((20, 2, 1), 99.0)]
    b = orc.Basis("monomial", 4, (-5.0, 5.0))
    with np.errstate(all="ignore"):
        want = orc.estimate_moments(levels, b)
    res = run_moments(to_struct(b), levels)
    assert np.array_equal(res["n"], [50, 0]) and np.array_equal(res["n_rm"], [0, 20])
    assert np.isnan(res["l_means"][1]).all() and np.isinf(res["l_vars"][1]).all()
    rel_close(res["l_means"][0], want.l_means[0], rtol=1e-10)


def test_empty_chunk_is_a_no_op():
    nat = native()
    b = to_struct(orc.Basis("legendre", 5, (-1.0, 1.0)))
    acc = nat.LevelAccumulator(1, 5, dev())
    nat.moments_accumulate(b, torch.empty(1, 0, 2, dtype=torch.float64, device=dev()), acc.level(0))
    acc2 = nat.LevelAccumulator(1, 25, dev())
    nat.gram_accumulate(b, torch.empty(1, 0, 2, dtype=torch.float64, device=dev()), acc2.level(0))
    assert float(acc.acc.abs().sum()) == 0.0 and float(acc2.acc.abs().sum()) == 0.0


def test_size_limits_and_error_text():
    nat = native()
    x = torch.zeros(1, 64, 2, dtype=torch.float64, device=dev())
    for size in (113, 226):                       # largest sizes with private / lane-pair accumulator columns
        b = to_struct(orc.Basis("legendre", size, (-1.0, 1.0)))
        acc = nat.LevelAccumulator(1, size, dev())
        nat.moments_accumulate(b, x, acc.level(0))
        a = acc.acc.cpu().numpy()[0]
        assert a[0] == 64 and a[2] == 0.0          # P_k(0) - P_k(0) = 0 for every k
    with pytest.raises(nat.NativeError, match="shared memory"):
        b = to_struct(orc.Basis("legendre", 240, (-1.0, 1.0)))
        nat.moments_accumulate(b, x, nat.LevelAccumulator(1, 240, dev()).level(0))
    with pytest.raises(nat.NativeError, match="block tasks|do not fit"):
        b = to_struct(orc.Basis("legendre", 130, (-1.0, 1.0)))
        nat.gram_accumulate(b, x, nat.LevelAccumulator(1, 130 * 130, dev()).level(0))
    with pytest.raises(nat.NativeError, match="basis size"):
        nat.basis_eval(nat.BasisStruct(1, 300, 0, 1, 0.0, 1.0, -1.0, 1.0), torch.zeros(4, dtype=torch.float64, device=dev()), 3)


def test_large_moment_count_values_match_oracle():
    """R = 200 (lane-pair columns, one CTA per SM): monic recurrence stays accurate over the whole range."""
    rng = np.random.default_rng(200)
    rows = orc.synth_level_rows(rng.uniform(-1.0, 1.0, size=3000), 0.02, 0.2)
    b = orc.Basis("legendre", 200, (-1.3, 1.3))
    want = orc.estimate_moments([np.zeros((1, 2, 1)), rows], b)
    nat = native()
    acc = nat.LevelAccumulator(2, 200, dev())
    nat.moments_accumulate(to_struct(b), torch.zeros(1, 1, 1, dtype=torch.float64, device=dev()), acc.level(0))
    nat.moments_accumulate(to_struct(b), torch.from_numpy(rows).to(dev()).permute(2, 0, 1), acc.level(1))
    out = acc.finalize()
    rel_close(out["l_means"][1].cpu().numpy(), want.l_means[1], rtol=1e-9, atol_scale=1e-13)
    rel_close(out["l_vars"][1].cpu().numpy(), want.l_vars[1], rtol=1e-9, atol_scale=1e-13)


def test_strided_scalar_column_of_wide_storage():
    """A scalar quantity that is one column of a 24-component storage row (strides 48 / 24 / 1): generic path."""
    rng = np.random.default_rng(11)
    wide = rng.normal(size=(700, 2, 24))
    b = orc.Basis("legendre", 7, (-3.0, 3.0))
    nat = native()
    for col in (0, 5, 23):
        want = orc.estimate_moments([np.zeros((1, 2, 1)), wide[:, :, col:col + 1]], b)
        acc = nat.LevelAccumulator(1, 7, dev())
        x = torch.from_numpy(wide).to(dev()).permute(2, 0, 1)[col:col + 1]
        nat.moments_accumulate(to_struct(b), x, acc.level(0))
        a = acc.acc.cpu().numpy()[0]
        n = a[0]
        assert n == want.n_samples[1]
        rel_close(a[2:9] / n, want.l_means[1], rtol=1e-10, atol_scale=1e-14)


@pytest.mark.parametrize("kind,R", [("legendre", 5), ("monomial", 6), ("legendre", 12), ("raw", 1)])
def test_streaming_variant_few_moments(kind, R):
    """Few moments + many samples take the TMA-ring (HBM-bound) variant: same results as the oracle, ragged tail
    tile included, level 0 in both layouts ([n, 1, 1] compact and [n, 2, 1] with the zero row)."""
    rng = np.random.default_rng(R)
    n1, n0 = 70_000 + 123, 50_000 + 7
    lvl1 = orc.synth_level_rows(rng.normal(size=n1), 0.05, 0.3)
    lvl0 = orc.synth_level_rows(rng.normal(size=n0), 0.3, None)
    lvl1[11, 0, 0] = np.nan
    lvl1[-1, 1, 0] = 99.0
    nat = native()
    if kind == "raw":
        basis, want = nat.RAW_BASIS, orc.estimate_mean([lvl0, lvl1], None)
    else:
        b = orc.Basis(kind, R, (-3.0, 3.1))
        basis, want = to_struct(b), orc.estimate_moments([lvl0, lvl1], b)
    for compact in (True, False):
        acc = nat.LevelAccumulator(2, R, dev())
        d0 = torch.from_numpy(np.ascontiguousarray(lvl0[:, :1, :] if compact else lvl0)).to(dev())
        x0 = d0.permute(2, 0, 1)[:, :, :1]
        nat.moments_accumulate(basis, x0, acc.level(0))
        nat.moments_accumulate(basis, torch.from_numpy(lvl1).to(dev()).permute(2, 0, 1), acc.level(1))
        out = acc.finalize()
        a = acc.acc.cpu().numpy()
        assert np.array_equal(a[:, 0], want.n_samples) and np.array_equal(a[:, 1], want.n_rm_samples)
        rel_close(out["l_means"].cpu().numpy(), want.l_means, rtol=1e-10, atol_scale=1e-14)
        rel_close(out["l_vars"].cpu().numpy(), want.l_vars, rtol=1e-10, atol_scale=1e-14)


# ------------------------------------------------------------------------------------------------------
# bootstrap re-sampling (estimator.py:171-218 over quantity.py:307-322): all replicates of a level in one launch
def _resampled_case(basis_o, rows, level0, n_rep, n_draws, seed):
    """rows float64[N,2,M]; returns (got [B, 2+2K] numpy, list of oracle LevelEstimate per replicate)."""
    nat = native()
    basis = to_struct(basis_o)
    M = rows.shape[2]
    d_rows = torch.from_numpy(np.ascontiguousarray(rows)).to(dev())
    x = d_rows.permute(2, 0, 1)
    if level0:
        x = x[:, :, :1]
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, len(rows), size=(n_rep, n_draws)).astype(np.int32)
    acc = torch.zeros((n_rep, 2 + 2 * M * basis.size), dtype=torch.float64, device=dev())
    nat.moments_accumulate_resampled(basis, x, torch.from_numpy(idx).to(dev()), acc)
    want = []
    for b in range(n_rep):
        picked = rows[idx[b]]
        want.append(orc.estimate_moments([picked] if level0 else [picked[:1], picked], basis_o))
    return acc.cpu().numpy(), want, idx


@pytest.mark.parametrize("tag,kind,size,level0,log,n_comp", [
    ("leg", "legendre", 12, False, False, 1),        # gather variant of the register-blocked scalar kernel
    ("leg0", "legendre", 9, True, False, 1),         # level 0 (16 samples per thread)
    ("legpair", "legendre", 60, False, False, 1),    # lane-pair accumulator columns
    ("leglog", "legendre", 7, False, True, 1),
    ("mono", "monomial", 5, False, False, 1),        # generic kernel with row indirection
    ("four", "fourier", 6, True, False, 1),
    ("vec", "legendre", 5, False, False, 3),         # vector quantity: mask indexed by storage row
])
def test_resampled_levels_match_oracle(tag, kind, size, level0, log, n_comp):
    rng = np.random.default_rng(42)
    n = 5000
    base = rng.lognormal(0.0, 0.5, size=(n, 1, n_comp)) if log else rng.normal(size=(n, 1, n_comp))
    rows = np.concatenate([base, base + 0.05 * rng.normal(size=(n, 1, n_comp))], axis=1)
    if log:
        rows = np.abs(rows) + 1e-3
    rows[rng.integers(0, n, 25), 0, 0] = np.nan                      # masked samples
    rows[rng.integers(0, n, 15), 1, n_comp - 1] = 1e6 if not log else 1e9   # outside the domain -> masked (safe_eval)
    dom = (0.05, 6.0) if log else (-3.5, 3.5)
    b = orc.Basis(kind, size, dom, log=log, safe_eval=True)
    n_rep, n_draws = 5, 3777                                          # ragged last tile
    got, want, idx = _resampled_case(b, rows, level0, n_rep, n_draws, seed=7)
    K = n_comp * size
    for r in range(n_rep):
        est = want[r]
        lvl = 0 if level0 else 1
        n_ok, n_rm = int(est.n_samples[lvl]), int(est.n_rm_samples[lvl])
        assert (int(got[r, 0]), int(got[r, 1])) == (n_ok, n_rm), (tag, r)
        assert n_ok + n_rm == n_draws
        rel_close(got[r, 2:2 + K] / n_ok, est.l_means[lvl], rtol=1e-10, atol_scale=1e-14)
        var = (got[r, 2 + K:] - got[r, 2:2 + K] ** 2 / n_ok) / (n_ok - 1)
        rel_close(var, est.l_vars[lvl], rtol=1e-9, atol_scale=1e-13)


def test_resampled_identity_draw_equals_plain_accumulate():
    """idx = 0..N-1 must reproduce the plain accumulate call (same tiles, same order of operations) bit for bit."""
    nat = native()
    rng = np.random.default_rng(3)
    n = 40000
    rows = rng.normal(size=(n, 2, 1))
    rows[:, 1, 0] = rows[:, 0, 0] + 0.01 * rng.normal(size=n)
    basis = to_struct(orc.Basis("legendre", 20, (-4.0, 4.0)))
    x = torch.from_numpy(rows).to(dev()).permute(2, 0, 1)
    idx = torch.arange(n, dtype=torch.int32, device=dev()).reshape(1, n)
    acc_r = torch.zeros((1, 42), dtype=torch.float64, device=dev())
    nat.moments_accumulate_resampled(basis, x, idx, acc_r)
    acc_p = torch.zeros(42, dtype=torch.float64, device=dev())
    nat.moments_accumulate(basis, x, acc_p)
    assert torch.equal(acc_r[0, :2], acc_p[:2])
    rel_close(acc_r[0].cpu().numpy(), acc_p.cpu().numpy(), rtol=1e-12)


def test_resampled_many_replicates_large():
    """Size-independent property at a larger size: with every replicate drawing the same rows, all replicates agree
    bit for bit, and their sums equal the plain accumulate of the gathered chunk."""
    nat = native()
    n, k, n_rep = 300000, 250000, 37
    g = torch.Generator(device=dev())
    g.manual_seed(5)
    fine = torch.randn(n, generator=g, device=dev(), dtype=torch.float64)
    rows = torch.stack([fine, fine + 0.02 * torch.randn(n, generator=g, device=dev(), dtype=torch.float64)], dim=1)
    x = rows.reshape(n, 2, 1).permute(2, 0, 1)
    basis = to_struct(orc.Basis("legendre", 30, (-4.0, 4.0)))
    one = torch.randint(0, n, (1, k), dtype=torch.int32, device=dev(), generator=g)
    idx = one.expand(n_rep, k).contiguous()
    acc = torch.zeros((n_rep, 62), dtype=torch.float64, device=dev())
    nat.moments_accumulate_resampled(basis, x, idx, acc)
    assert all(torch.equal(acc[0], acc[r]) for r in range(1, n_rep))
    gathered = rows[one[0].long()].reshape(k, 2, 1).permute(2, 0, 1)
    plain = torch.zeros(62, dtype=torch.float64, device=dev())
    nat.moments_accumulate(basis, gathered, plain)
    assert torch.equal(acc[0, :2], plain[:2])
    rel_close(acc[0].cpu().numpy(), plain.cpu().numpy(), rtol=1e-11)


def test_resample_indices_blocks_and_uniformity():
    nat = native()
    n_rows, k, n_rep, P = 1_000_003, 400_001, 3, 7
    plain = nat.resample_indices(5, 1, n_rows, k, n_rep, dev())
    again = nat.resample_indices(5, 1, n_rows, k, n_rep, dev())
    other = nat.resample_indices(6, 1, n_rows, k, n_rep, dev())
    assert torch.equal(plain, again) and not torch.equal(plain, other)
    assert not torch.equal(plain[0], plain[1])                         # replicates differ
    # a replicate's rows depend on (seed, its GLOBAL number) only, not on the grouping into calls
    tail = nat.resample_indices(5, 1, n_rows, k, 2, dev(), rep_offset=1)
    assert torch.equal(tail, plain[1:])
    assert not torch.equal(nat.resample_indices(6, 1, n_rows, k, 1, dev()), plain[1:2])     # seed+1 != replicate+1
    h = plain.cpu().numpy()
    assert h.min() >= 0 and h.max() < n_rows
    # uniform draws: mean n/2 +- 5 sigma, decile counts within 5 sigma of k/10
    assert abs(h.mean() - (n_rows - 1) / 2) < 5 * n_rows / np.sqrt(12 * h.size)
    dec = np.bincount((h.ravel().astype(np.int64) * 10) // n_rows, minlength=10)
    assert np.all(np.abs(dec - h.size / 10) < 5 * np.sqrt(h.size * 0.09))
    # block order with prescribed multinomial counts
    rng = np.random.default_rng(0)
    edges = (np.arange(P + 1, dtype=np.int64) * n_rows) // P
    counts = rng.multinomial(k, np.diff(edges) / n_rows, size=n_rep)
    cum = np.zeros((n_rep, P + 1), dtype=np.int64)
    np.cumsum(counts, axis=1, out=cum[:, 1:])
    blk = nat.resample_indices(5, 1, n_rows, k, n_rep, dev(), block_cum=torch.from_numpy(cum).to(dev())).cpu().numpy()
    for r in range(n_rep):
        block_of = np.searchsorted(edges, blk[r], side="right") - 1
        assert np.all(np.diff(block_of) >= 0)                          # listed in block order
        assert np.array_equal(np.bincount(block_of, minlength=P), counts[r])


def _block_cum(rng, n_rows, ks, P):
    edges = (np.arange(P + 1, dtype=np.int64) * n_rows) // P
    cum = np.zeros((len(ks), P + 1), dtype=np.int64)
    for b, k in enumerate(ks):
        np.cumsum(rng.multinomial(int(k), np.diff(edges) / n_rows), out=cum[b, 1:])
    return cum


@pytest.mark.parametrize("n_rows,P,ks", [(1_000_003, 7, [400_001, 400_001, 400_001]), (300_001, 3, [1, 0, 250_000, 999_999]),
                                         (5000, 1, [5000, 123]), (131, 1, [1000])])
def test_resample_counts_are_the_multiplicities_of_the_indices(n_rows, P, ks):
    """mlmcb200_resample_counts: counts[b, i] = how often mlmcb200_resample_indices lists row i for replicate b (same seed,
    stream, blocks): shared-memory histogram of the same Philox draws; ragged draws per replicate, unaligned block edges."""
    nat = native()
    rng = np.random.default_rng(1)
    cum = _block_cum(rng, n_rows, ks, P)
    cum_d = torch.from_numpy(cum).to(dev())
    counts = nat.resample_counts(5, 3, n_rows, cum_d, max(ks), dev(), rep_offset=2)
    assert counts.shape[0] == len(ks) and counts.shape[1] % 16 == 0 and counts.shape[1] >= n_rows
    for b, k in enumerate(ks):
        want = np.zeros(n_rows, dtype=np.int64)
        if k > 0:
            idx = nat.resample_indices(5, 3, n_rows, k, 1, dev(), block_cum=cum_d[b:b + 1].contiguous() if P > 1 else None,
                                       rep_offset=2 + b).cpu().numpy()[0]
            want = np.bincount(idx, minlength=n_rows)
        assert np.array_equal(counts[b, :n_rows].cpu().numpy().astype(np.int64), want)


@pytest.mark.parametrize("R,n_rep,level0,log,n_rows,kind", [
    (50, 100, False, False, 300_001, "legendre"), (51, 13, True, False, 70_000, "legendre"),
    (7, 300, False, True, 20_011, "legendre"), (1, 5, False, False, 999, "legendre"),
    (25, 16, True, True, 4096, "legendre"), (12, 40, False, False, 100, "legendre"),
    (100, 21, False, False, 50_001, "legendre"), (52, 9, True, False, 30_000, "legendre"),       # two column groups
    (160, 10, False, False, 9_000, "legendre"),                                                  # four
    (9, 24, False, False, 40_000, "monomial"), (60, 8, True, True, 20_000, "monomial"),
    (9, 24, False, False, 40_000, "fourier"), (32, 16, True, False, 20_000, "fourier"), (1, 5, False, False, 999, "fourier"),
    (60, 12, False, True, 10_000, "fourier"), (2, 9, False, False, 5_000, "fourier")])
def test_weighted_sums_equal_the_gather_kernel(R, n_rep, level0, log, n_rows, kind):
    """mlmcb200_moments_accumulate_weighted (all replicates in one pass: multiplicities x moment differences on DMMA tiles)
    adds what mlmcb200_moments_accumulate_resampled adds for the rows behind the multiplicities; samples outside the
    domain / NaN are counted as removed in both (mask_nan_samples, quantity_estimate.py:6-14)."""
    nat = native()
    rng = np.random.default_rng(R + n_rep)
    fine = rng.lognormal(size=n_rows) if log else rng.normal(size=n_rows)
    rows = np.stack([fine, fine * (1 + 0.03 * rng.normal(size=n_rows))], axis=1)
    bad = rng.choice(n_rows, size=max(1, n_rows // 50), replace=False)
    rows[bad[::3], 0] = np.nan
    rows[bad[1::3], 1] = 1e9
    rows[bad[2::3], 0] = -1e9 if not log else 1e-300
    if level0:
        rows = rows[:, :1]
    domain = (0.05, 12.0) if log else (-3.0, 3.0)
    basis = to_struct(orc.Basis(kind, R, domain, log=log))
    x = torch.from_numpy(np.ascontiguousarray(rows)).to(dev()).reshape(n_rows, rows.shape[1], 1).permute(2, 0, 1)
    P = max(1, -(-n_rows // 131072))
    ks = rng.integers(max(1, n_rows // 2), 2 * n_rows, size=n_rep)
    ks[0] = n_rows
    cum = torch.from_numpy(_block_cum(rng, n_rows, ks, P)).to(dev())
    counts = nat.resample_counts(9, 4, n_rows, cum, int(ks.max()), dev())
    width = 2 + 2 * R
    acc_w = torch.zeros((n_rep, width + 3), dtype=torch.float64, device=dev())[:, :width]      # strided replicate rows
    acc_w[:, :] = 0.5                                                                            # the call ADDS
    nat.moments_accumulate_weighted(basis, x, counts, acc_w)
    acc_g = torch.zeros((n_rep, width), dtype=torch.float64, device=dev())
    for b in range(n_rep):
        idx = nat.resample_indices(9, 4, n_rows, int(ks[b]), 1, dev(), block_cum=cum[b:b + 1].contiguous() if P > 1 else None,
                                   rep_offset=b)
        nat.moments_accumulate_resampled(basis, x, idx, acc_g[b:b + 1])
    got, want = acc_w.cpu().numpy() - 0.5, acc_g.cpu().numpy()
    assert np.array_equal(got[:, :2], want[:, :2]) and np.all(got[:, 0] + got[:, 1] == ks)
    assert want[:, 1].max() > 0 or (kind == "fourier" and R == 1)       # removed samples were drawn (a Fourier basis of
                                                                      # one function is the literal 1: nothing is dropped)
    rel_close(got[:, 2:2 + R], want[:, 2:2 + R], rtol=1e-10, atol_scale=1e-13)
    rel_close(got[:, 2 + R:], want[:, 2 + R:], rtol=1e-10, atol_scale=1e-13)


# ------------------------------------------------------------------------------------------------------
# order statistics for estimate_domain (estimator.py:299: np.percentile of the fine samples)
@pytest.mark.parametrize("case", ["normal", "dups", "tiny1", "tiny2", "tiny3", "signed_zero_inf", "strided", "big"])
def test_percentile_stats_equal_numpy(case):
    from mlmc_b200.estimator import _percentiles
    nat = native()
    rng = np.random.default_rng(11)
    stride = 1
    if case == "normal":
        x = rng.normal(size=100_003)
        x[rng.integers(0, x.size, 500)] = np.nan
    elif case == "dups":
        x = rng.integers(-5, 6, size=50_000).astype(float)
    elif case.startswith("tiny"):
        x = rng.normal(size=int(case[-1]))
    elif case == "signed_zero_inf":
        x = np.concatenate([rng.normal(size=1000), [0.0, -0.0, np.inf, -np.inf, 1e-310, -1e-310, np.nan]])
    elif case == "strided":
        x = rng.lognormal(size=(20_000, 3))
        stride = 3
    else:
        x = rng.standard_cauchy(size=3_000_000)
    t = torch.from_numpy(x).to(dev())
    values = t[:, 1] if stride == 3 else t
    host = x[:, 1] if stride == 3 else x
    host = host[~np.isnan(host)]
    fracs = [0.0, 0.001, 0.01, 0.25, 0.5, 0.99, 0.9999, 1.0]
    stats, n = nat.percentile_stats(values, fracs)
    assert n == host.size
    srt = np.sort(host)
    for f, (a, b) in zip(fracs, stats):
        pos = f * (n - 1)
        lo = int(np.floor(pos))
        hi = min(lo + 1, n - 1)
        assert a == srt[lo] and b == srt[hi], (case, f, a, srt[lo], b, srt[hi])
    percents = [100 * f for f in fracs]
    with np.errstate(invalid="ignore"):
        want = np.percentile(host, percents)
    got = _percentiles(values, percents)
    assert np.array_equal(got, want, equal_nan=True), (case, got, want)


@pytest.mark.parametrize("kind,size,n_coef", [("legendre", 25, 25), ("legendre", 50, 37), ("monomial", 8, 8),
                                              ("fourier", 9, 9)])
def test_density_eval_matches_oracle(kind, size, n_coef):
    """mlmcb200_density_eval = SimpleDistribution.density (simple_distribution.py:96-105): exp(clip(-phi . lambda/sigma)),
    NaN outside a clipped domain, the +-200 clip honoured."""
    nat = native()
    rng = np.random.default_rng(size)
    b = orc.Basis(kind, size, (-2.0, 3.0))
    x = np.concatenate([rng.uniform(-2.0, 3.0, 5000), [-2.0, 3.0, 3.5, -2.5]])
    coef = rng.normal(size=n_coef) * 0.7
    coef[0] = 1.3
    for scale in (1.0, 400.0):                                     # the second one drives |power| beyond 200
        want_pow = -np.sum(orc.basis_eval(b, x, n_coef) * coef * scale, axis=1)
        want = np.exp(np.minimum(np.maximum(want_pow, -200), 200))
        got = nat.density_eval(to_struct(b), torch.from_numpy(x).to(dev()), torch.from_numpy(coef * scale).to(dev()))
        got = got.cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(got[-2:]).all()
        inside = ~np.isnan(want)
        # d exp(p) = exp(p) dp: the dot product of n_coef terms of size |coef| differs by a few ulp of its largest term
        tol = 1e-13 * scale * np.sum(np.abs(coef)) + 1e-15
        assert np.max(np.abs(np.log(got[inside]) - np.log(want[inside]))) <= tol
        if scale > 1.0:
            assert (got[inside] == np.exp(200.0)).any() or (got[inside] == np.exp(-200.0)).any()


@pytest.mark.parametrize("R", [1, 2, 8, 9, 32, 33, 40, 41, 56, 57, 96, 104])
def test_covariance_block_boundaries_all_modes(R):
    """Every tile width (<= 4 / 7 / 13 moment blocks), group size and producer variant of the DMMA kernel at the sizes where
    they change over: sums, sums + squares, level 0, Gram of the differences; several ragged tiles per launch."""
    rng = np.random.default_rng(300 + R)
    n0, n1 = 1100, 777
    rows0 = orc.synth_level_rows(rng.normal(size=n0), 0.3, None)
    rows1 = orc.synth_level_rows(rng.normal(size=n1), 0.1, 0.3)
    rows1[5, 1, 0] = 40.0                                          # coarse value outside the domain: sample dropped
    b = orc.Basis("legendre", R, (-3.3, 3.1))
    want = orc.estimate_covariance([rows0, rows1], b, chunk_rows=512)
    res = run_gram(to_struct(b), [rows0, rows1], True)
    assert np.array_equal(res["n"], want.n_samples) and np.array_equal(res["n_rm"], want.n_rm_samples)
    rel_close(res["l_means"], want.l_means, rtol=1e-8, atol_scale=1e-13, per_level=True)
    rel_close(res["l_vars"], want.l_vars, rtol=1e-8, atol_scale=1e-12, per_level=True)
    sums_only = run_gram(to_struct(b), [rows0, rows1], False)
    rel_close(sums_only["l_means"], want.l_means, rtol=1e-8, atol_scale=1e-13, per_level=True)
    m = sums_only["l_means"].reshape(2, R, R)
    assert np.array_equal(m, np.swapaxes(m, 1, 2))
    # Gram of the differences (single-array tile)
    phi = orc.basis_eval(b, rows1[:, :, 0])
    good = ~np.isnan(phi).any(axis=(1, 2))
    d = phi[good, 0] - phi[good, 1]
    nat = native()
    acc = nat.LevelAccumulator(1, R * R, dev())
    nat.gram_accumulate(to_struct(b), torch.from_numpy(rows1).to(dev()).permute(2, 0, 1), acc.level(0), mode=1,
                        want_var=False)
    a = acc.acc.cpu().numpy()[0]
    assert a[0] == good.sum() and a[1] == (~good).sum()
    rel_close(a[2:2 + R * R].reshape(R, R), d.T @ d, rtol=1e-8, atol_scale=1e-13)


@pytest.mark.parametrize("n_comp,has_coarse,log", [(130, True, False), (257, False, False), (64, True, True),
                                                   (1000, True, False)])
def test_sample_mask_row_and_generic_kernels_agree(n_comp, has_coarse, log):
    """mlmcb200_sample_mask: the row-streaming kernel (storage layout, 16-byte loads, one warp per sample) and the generic
    strided kernel give the mask of mask_nan_samples applied to the moments of all components (quantity_estimate.py:6-14)."""
    nat = native()
    rng = np.random.default_rng(n_comp)
    n = 700
    rows = rng.uniform(0.5, 2.5, size=(n, 2 if has_coarse else 1, n_comp))
    bad = rng.choice(n, size=40, replace=False)
    rows[bad, rng.integers(0, rows.shape[1], 40), rng.integers(0, n_comp, 40)] = \
        rng.choice([7.0, 0.1, np.nan, np.inf], size=40)
    b = orc.Basis("legendre", 9, (0.4, 2.6), log=log)
    t = orc.to_ref_domain(b, rows)
    want = ~np.isnan(t).any(axis=(1, 2))
    d = torch.from_numpy(rows).to(dev())
    x = d.permute(2, 0, 1)                                          # storage layout: row-streaming kernel
    got_rows = nat.sample_mask(to_struct(b), x).cpu().numpy().astype(bool)
    wide = torch.zeros((n, rows.shape[1], n_comp + 3), dtype=torch.float64, device=dev())
    wide[:, :, 1:n_comp + 1] = d                                    # components at an odd offset of wider rows: generic kernel
    got_generic = nat.sample_mask(to_struct(b), wide[:, :, 1:n_comp + 1].permute(2, 0, 1)).cpu().numpy().astype(bool)
    assert np.array_equal(got_rows, want) and np.array_equal(got_generic, want)
    assert 0 < want.sum() < n
