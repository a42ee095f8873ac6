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


def test_sass_contains_fp64_tensor_and_no_legacy_half_mma():
    """The covariance / max-ent kernels must carry DMMA (FP64 tensor op); built with -lineinfo for sm_100a."""
    import subprocess
    from mlmc_b200 import build
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(build.OUT_DIR, "gram.o")], capture_output=True, text=True)
    if out.returncode != 0:
        import pytest
        pytest.skip("cuobjdump not available")
    assert "DMMA" in out.stdout and "sm_100a" in out.stdout
