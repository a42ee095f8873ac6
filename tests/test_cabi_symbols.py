"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol include/mlmcb200.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mlmcb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mlmcb200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from mlmc_b200 import _native
    lib = _native.load()
    names = declared_symbols()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(_native.EXPORTED_SYMBOLS)
    assert lib.mlmcb200_abi_version() == 2


def test_struct_layout_matches_header():
    from mlmc_b200 import _native
    assert ctypes.sizeof(_native.BasisStruct) == 4 * 4 + 4 * 8
    b = _native.make_basis(_native.LEGENDRE, 7, (2.0, 6.0), (-1.0, 1.0))
    assert (b.kind, b.size, b.is_log, b.is_clip) == (1, 7, 0, 1)
    assert b.shift == 2.0 and b.scale == 0.5 and (b.ref_lo, b.ref_hi) == (-1.0, 1.0)


def test_argument_errors_are_reported_without_a_gpu():
    from mlmc_b200 import _native
    lib = _native.load()
    bad = _native.BasisStruct(9, 3, 0, 1, 0.0, 1.0, -1.0, 1.0)
    rc = lib.mlmcb200_basis_eval(ctypes.byref(bad), None, 4, None, 0, 3, None, None)
    assert rc < 0 and b"unknown basis kind" in lib.mlmcb200_last_error()
    assert lib.mlmcb200_moments_workspace_bytes(50, 1) > 0
    assert lib.mlmcb200_moments_workspace_bytes(5000, 1) < 0       # does not fit shared memory
    assert lib.mlmcb200_gram_workspace_bytes(100) > 0
    assert lib.mlmcb200_gram_workspace_bytes_comp(25, 300) > lib.mlmcb200_gram_workspace_bytes_comp(25, 1) > 0
    good = _native.make_basis(_native.LEGENDRE, 7, (2.0, 6.0), (-1.0, 1.0))
    one = ctypes.c_void_p(8)                                       # never dereferenced: the argument checks come first
    rc = lib.mlmcb200_gram_accumulate_comp(ctypes.byref(good), one, 10, 3, 6, 3, 1, 1, None, 0, 1, one, one, 1 << 20, None)
    assert rc < 0 and b"sample mask" in lib.mlmcb200_last_error()
    rc = lib.mlmcb200_gram_accumulate_comp(ctypes.byref(good), one, 10, 1, 2, 1, 0, 0, None, 1, 0, one, one, 1 << 20, None)
    assert rc < 0 and b"needs a coarse side" in lib.mlmcb200_last_error()
    rc = lib.mlmcb200_density_eval(ctypes.byref(good), one, 5, one, 8, one, None)
    assert rc < 0 and b"n_coef" in lib.mlmcb200_last_error()
    # one-pass bootstrap (csrc/bootstrap.cu)
    assert lib.mlmcb200_moments_weighted_max_size() == 226 and 0 < lib.mlmcb200_resample_counts_block_rows() <= 196608
    assert lib.mlmcb200_moments_weighted_workspace_bytes(10_000_000, 300, 50) == \
        3 * lib.mlmcb200_moments_weighted_workspace_bytes(10_000_000, 100, 50) > 0
    assert lib.mlmcb200_moments_weighted_workspace_bytes(10_000_000, 100, 100) == \
        2 * lib.mlmcb200_moments_weighted_workspace_bytes(10_000_000, 100, 51)         # 202 columns: two groups of 104
    wide = _native.make_basis(_native.LEGENDRE, 227, (2.0, 6.0), (-1.0, 1.0))
    for b in (_native.RAW_BASIS, wide):
        rc = lib.mlmcb200_moments_accumulate_weighted(ctypes.byref(b), one, 10, 2, 1, one, 16, 3, one, 600, one, 1 << 30, None)
        assert rc < 0 and b"Fourier bases of at most 226" in lib.mlmcb200_last_error()
    rc = lib.mlmcb200_moments_accumulate_weighted(ctypes.byref(good), one, 10, 3, 1, ctypes.c_void_p(16), 16, 3, one, 200, one,
                                                  1 << 30, None)
    assert rc < 0 and b"storage order" in lib.mlmcb200_last_error()
    rc = lib.mlmcb200_moments_accumulate_weighted(ctypes.byref(good), ctypes.c_void_p(16), 10, 2, 1, ctypes.c_void_p(16), 10, 3,
                                                  one, 200, one, 1 << 30, None)
    assert rc < 0 and b"16-byte aligned" in lib.mlmcb200_last_error()
    rc = lib.mlmcb200_resample_counts(1, 1, 1_000_000, 10, 2, 0, 2, one, one, 1_000_000, None)
    assert rc < 0 and b"row blocks of more than" in lib.mlmcb200_last_error()
    rc = lib.mlmcb200_resample_counts(1, 1, 1000, 8001, 2, 0, 1, one, one, 1008, None)
    assert rc < 0 and b"more than 8 draws per row" in lib.mlmcb200_last_error()


def test_bootstrap_path_choice_is_world_size_independent():
    """quantity_estimate._weighted_bootstrap_applies: the one-pass weighted kernel for scalar Legendre bases of <= 51
    moments in storage order when the replicates draw about as many rows as the chunk holds; the rule reads the expected
    draws and the TOTAL replicate count, never this rank's share."""
    import numpy as np
    import torch
    from mlmc_b200 import _native
    from mlmc_b200.quantity import quantity_estimate as qe
    leg = _native.make_basis(_native.LEGENDRE, 50, (-3.0, 3.0), (-1.0, 1.0))
    rows = torch.zeros((1000, 2, 1), dtype=torch.float64)
    x = rows.permute(2, 0, 1)
    sizes = np.full(7, 1000)
    assert qe._weighted_bootstrap_applies(None, leg, x, sizes, 1000, 1000.0, 100)
    assert qe._weighted_bootstrap_applies(None, leg, x, sizes[:1], 1000, 1000.0, 100)      # a rank holding one replicate
    assert not qe._weighted_bootstrap_applies(None, leg, x, np.full(7, 100), 1000, 100.0, 100)        # few draws: gather
    assert qe._weighted_bootstrap_applies("weighted", leg, x, np.full(7, 100), 1000, 100.0, 100)
    assert not qe._weighted_bootstrap_applies("gather", leg, x, sizes, 1000, 1000.0, 100)
    assert not qe._weighted_bootstrap_applies("weighted", leg, x, np.full(7, 9000), 1000, 9000.0, 100)   # byte counters
    wide = _native.make_basis(_native.LEGENDRE, 60, (-3.0, 3.0), (-1.0, 1.0))
    huge = _native.make_basis(_native.LEGENDRE, 240, (-3.0, 3.0), (-1.0, 1.0))
    mono = _native.make_basis(_native.MONOMIAL, 9, (-3.0, 3.0), (0.0, 1.0))
    four = _native.make_basis(_native.FOURIER, 9, (-3.0, 3.0), (0.0, 6.283185307179586))
    assert qe._weighted_bootstrap_applies(None, wide, x, sizes, 1000, 1000.0, 100)           # two column groups
    assert qe._weighted_bootstrap_applies(None, mono, x, sizes, 1000, 1000.0, 100)
    assert qe._weighted_bootstrap_applies(None, four, x, sizes, 1000, 1000.0, 100)
    assert not qe._weighted_bootstrap_applies("weighted", huge, x, sizes, 1000, 1000.0, 100)
    vec = torch.zeros((1000, 2, 3), dtype=torch.float64).permute(2, 0, 1)
    assert not qe._weighted_bootstrap_applies("weighted", leg, vec, sizes, 1000, 1000.0, 100)
    assert not qe._weighted_bootstrap_applies("weighted", leg, vec[1:2], sizes, 1000, 1000.0, 100)   # one component of a field


def test_sass_contains_fp64_tensor_and_no_legacy_half_mma():
    """The covariance / max-ent kernels must carry DMMA (FP64 tensor op); built with -lineinfo for sm_100a."""
    import subprocess
    from mlmc_b200 import build
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(build.OUT_DIR, "gram.o")], capture_output=True, text=True)
    if out.returncode != 0:
        import pytest
        pytest.skip("cuobjdump not available")
    assert "DMMA" in out.stdout and "sm_100a" in out.stdout
