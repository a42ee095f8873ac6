"""The quantity DAG on device tensors against NumPy evaluations of the same expressions on the raw sample arrays.

Storage shape and value pattern follow the reference's quantity tests (test/test_quantity_concept.py:498-535: three
quantities depth / length / width with shapes (2,2) / (2,3) / (2,4), times [1,2,3], two locations each -> 24 + 36 + 48
= 108 components, 3 levels x 150 samples, sample s holds integers from [5 + 5 s, 10 + 5 s)); the checks are this
repository's own: every derived quantity is compared with ``oracle.estimate_mean`` applied to the equivalent NumPy
chunk operation, not only with other routes through the same code.
"""
import numpy as np
import pytest

from oracle import mlmc_oracle as orc

pytestmark = pytest.mark.gpu
SIZES = (24, 36, 48)


def close(a, b, tol=1e-12):
    a, b = np.asarray(a, float).reshape(-1), np.asarray(b, float).reshape(-1)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.allclose(a, b, rtol=tol, atol=tol * max(1.0, float(np.max(np.abs(b))))), float(np.max(np.abs(a - b)))


@pytest.fixture(scope="module")
def world():
    from mlmc_b200.sample_storage import Memory
    from mlmc_b200.quantity.quantity import make_root_quantity
    from mlmc_b200.quantity.quantity_spec import QuantitySpec
    rng = np.random.default_rng(123)
    fmt = [QuantitySpec(name="depth", unit="mm", shape=(2, 2), times=[1, 2, 3], locations=["30", "40"]),
           QuantitySpec(name="length", unit="m", shape=(2, 3), times=[1, 2, 3], locations=["10", "20"]),
           QuantitySpec(name="width", unit="mm", shape=(2, 4), times=[1, 2, 3], locations=["30", "40"])]
    levels = []
    for l in range(3):
        lo = 5 + 5 * np.arange(150)[:, None, None]
        rows = rng.integers(0, 5, size=(150, 2, 108)).astype(float) + lo
        if l == 0:
            rows[:, 1, :] = 0.0
        levels.append(rows)
    storage = Memory.from_arrays(levels, level_parameters=[[1.0]] * 3, result_format=fmt)
    return levels, make_root_quantity(storage, fmt)


def ref_mean(levels, op=None):
    with np.errstate(all="ignore"):
        return orc.estimate_mean(levels, op)


def est(q):
    from mlmc_b200.quantity.quantity_estimate import estimate_mean
    return estimate_mean(q)


def test_structure_indexing_and_interpolation(world):
    levels, root = world
    m = est(root)
    assert len(m.mean) == sum(SIZES)
    want = ref_mean(levels)
    close(m.mean, want.mean)
    close(m.var, want.var)
    close(est(root + root).mean, 2 * want.mean)
    depth, length, width = root["depth"], root["length"], root["width"]
    close(est(depth).mean, want.mean[:24])
    close(est(length).mean, want.mean[24:60])
    close(est(width).mean, want.mean[60:])
    close(est((root + root)["width"]).mean, 2 * want.mean[60:])

    # linear interpolation in time, then location, then array entries
    def interp(x, lo, n_in, t0, t1, w):
        blk = x[lo:lo + 3 * n_in].reshape(3, n_in, *x.shape[1:])
        return blk[t0] + (blk[t1] - blk[t0]) * w
    at_25 = length.time_interpolation(2.5)
    w25 = ref_mean(levels, lambda x: interp(x, 24, 12, 1, 2, 0.5)).mean
    close(est(at_25).mean, w25)
    pos = at_25["10"]
    close(est(pos).mean, w25[:6])
    close(est(at_25["20"]).mean, w25[6:])
    pm = est(pos)
    assert pm[1:2].mean.shape == (1, 3) and pm[1].mean.shape == (3,)
    close(pm[1].mean, w25[3:6])
    close(est(pos[:, 2]).mean, w25[[2, 5]])
    close(est(pos[1, 2]).mean, w25[[5]])
    close(est(pos[:, :]).mean, w25[:6])
    close(est(pos[:1, 1:2]).mean, w25[[1]])
    assert est(pos[:2, ...]).mean.size == 6
    at_12 = width.time_interpolation(1.2)
    w12 = ref_mean(levels, lambda x: interp(x, 60, 16, 0, 1, 0.2)).mean
    close(est(at_12["30"]).mean, w12[:8], tol=1e-11)
    close(est(at_12["40"]).mean, w12[8:], tol=1e-11)
    with pytest.raises(ValueError):
        length.time_interpolation(3.5)
    close(est(np.add(5, pos[1, 2])).mean, w25[[5]] + 5)


def test_composition(world):
    from mlmc_b200.quantity.quantity import Quantity
    levels, root = world
    want = ref_mean(levels).mean
    depth, length = root["depth"], root["length"]
    qd = Quantity.QDict([("depth", depth), ("length", length)])
    close(est(qd).mean, want[:60])
    close(est(qd["length"]).mean, want[24:60])
    close(est(qd["depth"]).mean, want[:24])
    close(est(Quantity.QArray([[length, length], [length, length]])).mean, np.tile(want[24:60], 4))
    loc = length.time_interpolation(2.0)
    w2 = want[24 + 12:24 + 24]
    close(est(Quantity.QTimeSeries([(0, loc), (1, loc)])).mean, np.tile(w2, 2))
    close(est(Quantity.QField([("f1", length), ("f2", length)])).mean, np.tile(want[24:60], 2))


def test_arithmetic_with_constants(world):
    levels, root = world
    c = 5
    base = ref_mean(levels).mean
    close(est(root + c).mean, ref_mean(levels, lambda x: x + c).mean)
    close(est(c + root).mean, ref_mean(levels, lambda x: c + x).mean)
    close(est(root - c).mean, -est(c - root).mean)
    close(est(root * c).mean, c * base)
    close(est(c * root).mean, c * base)
    close(est(root / c).mean, base / c)
    close(est(c / root).mean, ref_mean(levels, lambda x: c / x).mean)
    close(est(root % c).mean, ref_mean(levels, lambda x: np.mod(x, c)).mean)
    assert len(est(c % root).mean) == 108
    close(est(root + root + root).mean, 3 * base)
    close(est(root + root * c).mean, base + base * c)
    close(est(root + root * root).mean, ref_mean(levels, lambda x: x + x * x).mean)
    close(est(root / root).mean, np.concatenate([[1.0] * 108]))        # level 0: 1, higher levels: 1 - 1 = 0


def test_conditions_select(world):
    levels, root = world
    base = est(root).mean
    close(est(root.select(np.logical_or(0 < root, root < 10))).mean, base)
    close(est(root.select(0 < root)).mean, base)
    close(est(root.select(root == root)).mean, base)
    close(est(root.select(-1 != root["length"])).mean, base)
    for nothing in (root < 0, root < root, root["length"] < 1, 10 ** 5 < root["length"], 10 ** 5 <= root["length"],
                    1 == root["length"]):
        with pytest.raises(Exception):
            est(root.select(nothing))

    def rows_where(cond):                                          # a sample stays if ALL its entries meet the condition
        out = []
        for l, lv in enumerate(levels):
            x = lv[:, :1, :] if l == 0 else lv
            out.append(lv[cond(x).all(axis=(1, 2))])
        return out
    bounded = est(root.select(0 < root, root < 10))
    close(bounded.mean, ref_mean(rows_where(lambda x: (0 < x) & (x < 10))).mean)
    assert list(bounded.n_samples) == [1, 1, 1]                            # only sample 0 lies in [5, 10)
    close(est(root.select(np.logical_and(0 < root, root < 10))).mean, bounded.mean)
    both = root + root
    close(est(both.select(0 < both, both < 20)).mean, 2 * bounded.mean)
    mid = est(root.select(10 < root, root < 20))
    close(mid.mean, ref_mean(rows_where(lambda x: (10 < x) & (x < 20))).mean)
    close(est(both.select(20 < both, both < 40)).mean, 2 * mid.mean)
    close(est(both.select(root < both)).mean, 2 * base)
    with pytest.raises(Exception):
        est(both.select(root > both))
    close(est(both.select(root < both, root["length"] < 10)).mean, 2 * bounded.mean)
    length = root["length"]
    close(est(length.select(length <= 9)).mean, bounded.mean[24:60])
    close(est(length.select(9 < length, length < 20)).mean,
          ref_mean(rows_where(lambda x: (9 < x[..., 24:60]) & (x[..., 24:60] < 20))).mean[24:60])


def test_numpy_functions(world):
    levels, root = world
    base = est(root).mean
    mx = est(np.max(root, axis=0, keepdims=True))
    assert len(mx.mean) == 1
    close(mx.mean, ref_mean(levels, lambda x: np.max(x, axis=0, keepdims=True)).mean)
    close(est(np.sum(root, axis=0, keepdims=True)).mean, ref_mean(levels, lambda x: np.sum(x, axis=0, keepdims=True)).mean)
    close(est(np.min(root["depth"], axis=0)).mean, ref_mean(levels, lambda x: np.min(x[:24], axis=0, keepdims=True)).mean)
    sin = est(np.sin(root))
    close(sin.mean, ref_mean(levels, np.sin).mean, tol=1e-11)
    close(est(np.sin(root["length"])).mean, sin.mean[24:60])
    close(est(np.add(root, root)).mean, 2 * base)
    ones = np.ones(108)
    close(est(np.add(ones, root)).mean, ref_mean(levels, lambda x: x + 1).mean)
    inv = est(np.divide(ones, root))
    assert np.all(inv.mean < 1)
    close(est(np.arctan2(ones, root)).mean, ref_mean(levels, lambda x: np.arctan2(1.0, x)).mean, tol=1e-11)
    close(est(np.maximum(root, root)).mean, base)
    close(est(np.exp(-root / 100)).mean, ref_mean(levels, lambda x: np.exp(-x / 100)).mean, tol=1e-11)
    close(est(abs(root - 300) ** 2).mean, ref_mean(levels, lambda x: np.abs(x - 300) ** 2).mean)
    with pytest.raises(TypeError):
        est(np.logical_and(True, root))
    with pytest.raises(ValueError):
        np.add(np.ones((108, 5, 2)), root)
    with pytest.raises(ValueError):
        np.divide(np.ones((108, 5, 2)), root)


def test_constants_stay_constants():
    from mlmc_b200.quantity.quantity import QuantityConst
    from mlmc_b200.quantity.quantity_types import ScalarType
    z = QuantityConst(ScalarType(), 5) + QuantityConst(ScalarType(), 10)
    assert isinstance(z, QuantityConst)
    assert float(z.samples(type("C", (), {"level_id": 0})())[0, 0, 0]) == 15.0
