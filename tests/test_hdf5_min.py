"""CPU: the minimal HDF5 reader behind ``SampleStorageHDF`` (no h5py in this image).

* a file written by the REAL HDF5 library (MATLAB v7.3 file from scipy's test data: user block, version 0 superblock,
  symbol-table group, version 1 object header, contiguous float64 dataset, string attribute) parses to its known
  content -- ``testdouble = linspace(0, 2 pi, 9)`` (the same variable as in scipy's ``testdouble_7.4_GLNX86.mat``);
* the reference's file structure (``mlmc/tool/hdf5.py:14-45, 311-320``) written by the fixture writer -- chunked
  ``collected_values`` of element type (2, M) float64, two-level chunk B-tree, partial last chunk -- reads back exactly,
  whole, by row ranges, and into caller-provided (staging) buffers;
* ``SampleStorageHDF`` keeps the storage contract of ``mlmc/sample_storage_hdf.py:144-184``;
* whenever h5py IS importable, a file written by h5py exactly as the reference writes it must read identically through
  both back ends.
"""
import os

import numpy as np
import pytest

from mlmc_b200.tool import hdf5_min


def _mat73():
    try:
        import scipy.io.matlab
        path = os.path.join(os.path.dirname(scipy.io.matlab.__file__), "tests", "data", "testhdf5_7.4_GLNX86.mat")
    except ImportError:
        return None
    return path if os.path.exists(path) else None


def test_reads_a_file_written_by_the_hdf5_library():
    path = _mat73()
    if path is None:
        pytest.skip("scipy's MATLAB v7.3 test file is not installed")
    with hdf5_min.File(path) as f:
        assert f._base == 512 and f.keys() == ["testdouble"]
        dset = f["testdouble"]
        assert dset.shape == (9, 1) and dset.dtype == np.dtype("<f8") and dset.chunks is None
        assert dset.attrs["MATLAB_class"].tobytes() == b"double"
        values = dset[()]
        assert np.array_equal(values[2:5], dset.read_rows(2, 5))
    assert np.allclose(values[:, 0], np.linspace(0, 2 * np.pi, 9), rtol=0, atol=1e-15)


def _levels(rng, sizes, m):
    out = []
    for l, n in enumerate(sizes):
        rows = rng.normal(size=(n, 2, m))
        if l == 0:
            rows[:, 1, :] = 0
        out.append(rows)
    return out


def test_round_trip_of_the_reference_layout(tmp_path):
    rng = np.random.default_rng(3)
    levels = _levels(rng, [1000, 333, 7, 0], 3)
    params = [[0.5], [0.1], [0.02], [0.004]]
    path = hdf5_min.write_mlmc_file(str(tmp_path / "mlmc.hdf5"), levels, params, n_ops=[1.0, 2.5, 7.0, 9.0], chunk_rows=7)
    with hdf5_min.File(path) as f:
        assert np.array_equal(f.attrs["level_parameters"], np.array(params))
        assert sorted(f["Levels"].keys()) == ["0", "1", "2", "3"]
        for l, rows in enumerate(levels):
            group = f["Levels"][str(l)]
            assert np.array_equal(group.attrs["n_ops_estimate"], [[1.0, 2.5, 7.0, 9.0][l], 1.0])
            dset = f["Levels/%d/collected_values" % l]
            assert dset.shape == (len(rows),) and dset.dtype == np.dtype(("<f8", (2, 3))) and dset.chunks == (7,)
            assert dset.maxshape == (None,)
            assert np.array_equal(dset[()], rows)
            if len(rows) > 20:
                assert np.array_equal(dset[5:19], rows[5:19]) and np.array_equal(dset.read_rows(13, 15), rows[13:15])
                out = np.full((10, 2, 3), np.nan)
                dset.read_rows(len(rows) - 10, len(rows), out=out)               # ends in the partial last chunk
                assert np.array_equal(out, rows[-10:])
            slices = [s[0] for s in dset.iter_chunks()]
            assert sum(s.stop - s.start for s in slices) == len(rows)
        # 1000 rows / 7 per chunk = 143 chunks > 64 entries per node: the chunk B-tree has two levels
        assert len(f["Levels/0/collected_values"]._index()) == 143


def test_sample_storage_hdf_contract(tmp_path):
    from mlmc_b200.sample_storage import SampleStorageHDF, RowSource
    from mlmc_b200.quantity.quantity_spec import ChunkSpec
    rng = np.random.default_rng(4)
    levels = _levels(rng, [50, 20, 9], 4)
    path = hdf5_min.write_mlmc_file(str(tmp_path / "s.hdf5"), levels, [[0.3], [0.1], [0.03]], n_ops=[2.0, 4.0, 8.0],
                                    chunk_rows=16)
    st = SampleStorageHDF(path, backend="min")
    assert st.backend == "min" and st.get_level_ids() == [0, 1, 2] and st.get_n_levels() == 3
    assert st.get_n_collected() == [50, 20, 9] and st.get_level_parameters() == [[0.3], [0.1], [0.03]]
    assert st.get_n_ops() == [2.0, 4.0, 8.0]
    c0 = st.sample_pairs_level(ChunkSpec(level_id=0))
    c1 = st.sample_pairs_level(ChunkSpec(level_id=1, chunk_slice=slice(3, 11)))
    assert c0.shape == (4, 50, 1) and c1.shape == (4, 8, 2)
    assert np.array_equal(c1, levels[1][3:11].transpose(2, 0, 1)) and np.array_equal(c0[:, :, 0], levels[0][:, 0, :].T)
    assert [s.level_id for s in st.chunks()] == [0, 1, 2]
    assert next(st.chunks(level_id=1, n_samples=5)).chunk_slice == slice(0, 5, 1)
    # the streaming feed's row source: level 0 without its zero coarse row, sliceable (multi-GPU row ranges)
    src = st._host_tensor(0)
    assert isinstance(src, RowSource) and src.shape == (50, 1, 4)
    buf = np.empty((10, 1, 4))
    src.slice_rows(20, 40).read_into(5, 15, buf)
    assert np.array_equal(buf, levels[0][25:35, :1, :])
    src1 = st._host_tensor(1)
    buf = np.empty((20, 2, 4))
    src1.read_into(0, 20, buf)
    assert src1.shape == (20, 2, 4) and np.array_equal(buf, levels[1])
    with pytest.raises(NotImplementedError):
        st.save_samples({}, {})


def test_npy_storage_row_counts_without_loading(tmp_path):
    from mlmc_b200.sample_storage import NpyStorage, RowSource
    rng = np.random.default_rng(5)
    levels = _levels(rng, [40, 10], 2)
    st = NpyStorage.write(str(tmp_path / "npy"), levels, [[0.5], [0.1]], [1.0, 3.0])
    assert st.get_n_collected() == [40, 10]
    assert isinstance(st.level_rows(0), np.memmap)                  # header + map only
    src = st._host_tensor(0)
    assert isinstance(src, RowSource) and src.shape == (40, 1, 2)


def test_unsupported_features_are_named(tmp_path):
    p = tmp_path / "x.bin"
    p.write_bytes(b"not an hdf5 file" * 100)
    with pytest.raises(ValueError, match="not an HDF5 file"):
        hdf5_min.File(str(p))


def test_h5py_written_file_reads_identically(tmp_path):
    """Runs wherever h5py is installed: the file is created the way the reference creates it (hdf5.py:83-106, 311-340:
    shape (0,), maxshape (None,), chunks=True, array dtype, resize + append in several batches)."""
    h5py = pytest.importorskip("h5py")
    from mlmc_b200.sample_storage import SampleStorageHDF
    rng = np.random.default_rng(6)
    levels = _levels(rng, [5000, 700], 5)
    path = str(tmp_path / "ref.hdf5")
    with h5py.File(path, "a") as f:
        f.attrs["version"] = "1.0.1"
        f.attrs["level_parameters"] = [[0.5], [0.05]]
        f.create_group("Levels")
        for l, rows in enumerate(levels):
            g = f["Levels"].create_group(str(l))
            g.attrs["level_id"] = str(l)
            g.attrs["n_ops_estimate"] = [3.0 * (l + 1), 2.0]
            d = g.create_dataset("collected_values", shape=(0,), dtype=np.dtype((np.float64, (2, 5))), maxshape=(None,),
                                 chunks=True)
            for lo in range(0, len(rows), 997):
                part = rows[lo:lo + 997]
                d.resize(d.shape[0] + len(part), axis=0)
                d[-len(part):] = part
    for backend in ("min", "h5py"):
        st = SampleStorageHDF(path, backend=backend)
        assert st.get_n_collected() == [5000, 700] and st.get_level_parameters() == [[0.5], [0.05]]
        assert st.get_n_ops() == [1.5, 3.0]
        for l, rows in enumerate(levels):
            assert np.array_equal(np.asarray(st.level_rows(l)), rows)
            assert np.array_equal(st.level_rows(l)[100:300], rows[100:300])


@pytest.mark.parametrize("chunk_rows", [1, 7, 64])
def test_row_source_reads_in_parallel_and_skips_the_coarse_row(tmp_path, monkeypatch, chunk_rows):
    """The staged feed's CPU stage (``RowSource.read_into``): a piece is cut into row ranges copied by a thread pool,
    runs of file-contiguous chunks are one copy each, and level 0 is read without its stored zero coarse row straight
    from the file mapping (``read_rows(keep_items=...)``) -- same bytes as a plain slice, for unaligned ranges too."""
    from mlmc_b200.sample_storage import RowSource, SampleStorageHDF
    from mlmc_b200.tool.hdf5_min import write_mlmc_file
    rng = np.random.default_rng(chunk_rows)
    m = 3000                                                       # 48 kB per row: pieces above the 8 MB threshold
    levels = [rng.normal(size=(400, 2, m)), rng.normal(size=(333, 2, m))]
    levels[0][:, 1] = 0.0
    path = write_mlmc_file(str(tmp_path / "s.hdf5"), levels, [[0.5], [0.1]], chunk_rows=chunk_rows)
    storage = SampleStorageHDF(path, backend="min")
    for threads in ("1", "5"):
        monkeypatch.setenv("MLMCB200_READ_THREADS", threads)
        for level, drop in ((0, True), (0, False), (1, False)):
            src = RowSource(storage.level_rows(level), drop_coarse=drop)
            want = levels[level][:, :1] if drop else levels[level]
            assert src.shape == want.shape
            for lo, hi in ((0, len(want)), (3, len(want) - 5), (17, 18), (len(want) - 1, len(want))):
                out = np.full((hi - lo,) + want.shape[1:], np.nan)
                src.read_into(lo, hi, out)
                assert np.array_equal(out, want[lo:hi])
            part = src.slice_rows(10, 300)
            out = np.empty((290,) + want.shape[1:])
            part.read_into(0, 290, out)
            assert np.array_equal(out, want[10:300])
    dset = storage.level_rows(1)._dset
    from mlmc_b200.tool.hdf5_min import Unsupported
    with pytest.raises(Unsupported):
        dset.read_rows(0, 4, out=np.empty((4, m)), keep_items=2 * m + 1)
    with pytest.raises(ValueError):
        dset.read_rows(0, 4, out=np.empty((3, m)), keep_items=m)
