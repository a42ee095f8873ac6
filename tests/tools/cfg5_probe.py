import os, sys, time, cProfile, pstats
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import mlmc_oracle as orc
from mlmc_b200.moments import Fourier
from mlmc_b200.sample_storage import Memory
from mlmc_b200.quantity.quantity import make_root_quantity
from mlmc_b200.quantity.quantity_spec import QuantitySpec
from mlmc_b200.quantity import quantity_estimate as qe
M = 10_000
n_levels = [4096, 2048, 1024, 512, 256]
steps = orc.level_steps(5, (0.5, 0.005))
rng = np.random.default_rng(5)
levels = []
for l, n in enumerate(n_levels):
    base_rows = orc.synth_level_rows(rng.normal(size=n), steps[l], steps[l - 1] if l else None)
    rows = np.repeat(base_rows, M, axis=2) + (np.arange(M) * 1e-4)[None, None, :]
    levels.append(rows)
spec = [QuantitySpec(name="field", unit="", shape=(1, 1), times=[0.0], locations=[str(i) for i in range(M)])]
storage = Memory.from_arrays(levels, level_parameters=[[h] for h in steps], result_format=spec)
field = make_root_quantity(storage, spec)["field"][0.0]
fn = Fourier(32, (-4.2, 5.4))
for _ in range(2):
    qe.estimate_mean(qe.moments(field, fn))
torch.cuda.synchronize()
t0 = time.perf_counter(); q = qe.moments(field, fn); t1 = time.perf_counter()
qm = qe.estimate_mean(q); torch.cuda.synchronize(); t2 = time.perf_counter()
m = qm.mean; t3 = time.perf_counter()
print("build quantity %.2f ms, estimate_mean %.2f ms, .mean %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
pr = cProfile.Profile(); pr.enable()
qm = qe.estimate_mean(qe.moments(field, fn)); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
ts = []
for _ in range(8):
    t0 = time.perf_counter(); qm = qe.estimate_mean(qe.moments(field, fn)); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("8 consecutive calls (ms):", [round(t, 2) for t in ts])
import gc
gc.disable()
ts = []
for _ in range(8):
    t0 = time.perf_counter(); qm = qe.estimate_mean(qe.moments(field, fn)); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("gc disabled           :", [round(t, 2) for t in ts])

# kernel-level breakdown (CUDA events around each launch group)
from mlmc_b200 import _native as nat
dev = torch.device("cuda:0")
basis = fn.basis_struct()
rows_dev = [torch.from_numpy(lv).to(dev) for lv in levels]
acc = nat.LevelAccumulator(5, M * 32, dev)
def ev():
    return torch.cuda.Event(enable_timing=True)
for rep in range(2):
    t_mask = t_acc = 0.0
    for l, rows in enumerate(rows_dev):
        x = rows.permute(2, 0, 1)
        if l == 0:
            x = x[:, :, :1]
        e0, e1, e2 = ev(), ev(), ev()
        e0.record(); valid = nat.sample_mask(basis, x); e1.record()
        nat.moments_accumulate(basis, x, acc.level(l), valid=valid); e2.record()
        torch.cuda.synchronize()
        t_mask += e0.elapsed_time(e1); t_acc += e1.elapsed_time(e2)
    e0, e1 = ev(), ev()
    e0.record(); out = acc.finalize(); e1.record(); torch.cuda.synchronize()
    print("sample_mask %.3f ms, moments kernels %.3f ms, finalize %.3f ms" % (t_mask, t_acc, e0.elapsed_time(e1)))
