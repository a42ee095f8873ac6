"""In-tree build of libmlmcb200.so (hand-written CUDA for sm_100a behind the C ABI of include/mlmcb200.h).

``python -m mlmc_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles here without a GPU; the
resulting ``mlmc_b200/_lib/libmlmcb200.so`` is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
SRC_DIR = os.path.join(PKG_DIR, "csrc")
OUT_DIR = os.path.join(PKG_DIR, "_lib")
LIB_PATH = os.path.join(OUT_DIR, "libmlmcb200.so")
SOURCES = ["moments_k_legendre_coarse.cu", "moments_k_legendre_level0.cu", "moments_k_monomial.cu", "moments_k_raw_fourier.cu",
           "gram.cu", "api.cu", "moments.cu", "basis_eval.cu", "maxent.cu", "select.cu", "peer.cu", "bootstrap.cu"]
HEADERS = ["common.cuh", "legendre_tables.inc", "moments_types.cuh", "moments_kernel.cuh", "philox.cuh",
           os.path.join("..", "..", "include", "mlmcb200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a and link the shared library.  Returns the library path."""
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(SRC_DIR, h) for h in HEADERS]
    sources = [s for s in SOURCES if os.path.exists(os.path.join(SRC_DIR, s))]
    jobs = []
    objs = []
    for src in sources:
        src_path = os.path.join(SRC_DIR, src)
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src_path] + headers):
            jobs.append((src, [nvcc] + NVCC_FLAGS + ["-c", src_path, "-o", obj]))

    def run(job):
        src, cmd = job
        res = subprocess.run(cmd, capture_output=True, text=True)
        with open(os.path.join(OUT_DIR, src + ".ptxas.log"), "w") as f:
            f.write(res.stdout + res.stderr)
        return src, res

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        for src, res in pool.map(run, jobs):
            if res.returncode != 0:
                raise RuntimeError("nvcc failed on %s:\n%s" % (src, res.stderr))
            if verbose:
                sys.stderr.write(res.stderr)
    if force or jobs or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s" % res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
