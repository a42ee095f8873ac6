"""CPU stage of the staged feed alone (file mapping -> pinned staging buffers), cfg5-sized HDF5 file in the page cache:
GB/s of staged bytes for the NumPy thread pool and for the native multi-threaded copy, by thread count and piece size.
usage: feed_probe.py [n_locations]"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlmc_b200 import sample_storage as ss  # noqa: E402
from mlmc_b200.tool import hdf5_min  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
n_levels = [4096, 2048, 1024, 512, 256]
rng = np.random.default_rng(0)
levels = [rng.random((n, 2, M)) for n in n_levels]
tmp = tempfile.mkdtemp(prefix="mlmcb200_feed_")
path = hdf5_min.write_mlmc_file(os.path.join(tmp, "cfg5.hdf5"), levels, [[0.5 ** l] for l in range(5)])
del levels
storage = ss.SampleStorageHDF(path, backend="min")
pinned = torch.cuda.is_available()
print("cores", len(os.sched_getaffinity(0)), "pinned staging", pinned, "file GB %.2f" % (os.path.getsize(path) / 1e9))
native = ss._native_copy()


def run(piece_bytes):
    sources = [ss.RowSource(storage.level_rows(l), drop_coarse=(l == 0)) for l in range(5)]
    n_max = piece_bytes // 8
    bufs = [torch.empty(n_max, dtype=torch.float64, pin_memory=pinned) for _ in range(2)]
    best = 1e9
    staged = 0
    for rep in range(3):
        t0 = time.perf_counter()
        k = staged = 0
        for src in sources:
            rows = max(1, piece_bytes // (8 * src.shape[1] * src.shape[2]))
            for lo in range(0, len(src), rows):
                hi = min(len(src), lo + rows)
                n = (hi - lo) * src.shape[1] * src.shape[2]
                src.read_into(lo, hi, bufs[k % 2][:n].view(hi - lo, src.shape[1], src.shape[2]).numpy())
                staged += n * 8
                k += 1
        best = min(best, time.perf_counter() - t0)
    return staged / best / 1e9, best * 1e3


for mode in ("numpy-pool", "native"):
    for threads in (1, 4, 8, 16):
        for piece in (32 << 20, 128 << 20):
            os.environ["MLMCB200_READ_THREADS"] = str(threads)
            ss._pool.pop("copy", None)
            if mode == "numpy-pool":
                ss._pool["copy"] = None
                hdf5_min.parallel_copy = None
            else:
                hdf5_min.parallel_copy = ss._native_copy()
            for l in range(5):
                storage.level_rows(l).native_copy = hdf5_min.parallel_copy is not None
            gbs, ms = run(piece)
            print("%-10s threads %2d piece %3d MB: %6.1f GB/s  (%.1f ms for all levels)" % (mode, threads, piece >> 20, gbs, ms),
                  flush=True)

if pinned:
    # the whole pipeline (file -> staging -> device -> kernels) on the cfg5 estimate, by variant of the CPU stage
    from mlmc_b200.moments import Fourier
    from mlmc_b200.quantity import quantity_estimate as qe
    from mlmc_b200.quantity.quantity import make_root_quantity
    from mlmc_b200.quantity.quantity_spec import QuantitySpec
    spec = [QuantitySpec(name="field", unit="", shape=(1, 1), times=[0.0], locations=[str(i) for i in range(M)])]
    fn = Fourier(32, (-0.2, 1.2))
    for mode, threads, bg, piece in (("numpy-pool", 16, 0, 32), ("numpy-pool", 8, 0, 32), ("numpy-pool", 4, 0, 32),
                                     ("native", 16, 0, 32), ("native", 8, 0, 32), ("native", 4, 0, 32),
                                     ("numpy-pool", 16, 0, 128), ("native", 8, 0, 128), ("native", 8, 1, 32),
                                     ("numpy-pool", 1, 0, 32)):
        os.environ["MLMCB200_READ_THREADS"] = str(threads)
        os.environ["MLMCB200_FEED_THREAD"] = str(bg)
        ss._pool.pop("copy", None)
        if mode == "numpy-pool":
            ss._pool["copy"] = None
            hdf5_min.parallel_copy = None
        else:
            hdf5_min.parallel_copy = ss._native_copy()
        st = ss.SampleStorageHDF(path, backend="min")
        st.resident_fraction = 0.0
        st.device_chunk_bytes = piece << 20
        field = make_root_quantity(st, spec)["field"][0.0]
        qe.estimate_mean(qe.moments(field, fn))
        best = 1e9
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            qe.estimate_mean(qe.moments(field, fn))
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        print("pipeline %-10s threads %2d feed-thread %d piece %3d MB: %6.1f ms" % (mode, threads, bg, piece, best * 1e3), flush=True)
import shutil
shutil.rmtree(tmp, ignore_errors=True)
