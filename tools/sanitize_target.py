"""Small invocation of every kernel for compute-sanitizer (memcheck): odd sizes, tails, masks, strided views."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlmc_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
for kind, ref in ((nat.LEGENDRE, (-1.0, 1.0)), (nat.MONOMIAL, (0.0, 1.0)), (nat.FOURIER, (0.0, 2 * np.pi))):
    for R in (1, 2, 7, 50, 100):
        b = nat.make_basis(kind, R, (-2.5, 2.5), ref)
        for n, M in ((1, 1), (1027, 1), (333, 3), (77, 130)):
            rows = torch.from_numpy(rng.normal(size=(n, 2, M))).to(dev)
            x = rows.permute(2, 0, 1)
            acc = nat.LevelAccumulator(2, M * R, dev)
            nat.moments_accumulate(b, x, acc.level(1))
            nat.moments_accumulate(b, x[:, :, :1], acc.level(0))
            acc.finalize()
        rows = torch.from_numpy(rng.normal(size=(1027, 2, 1))).to(dev)
        acc = nat.LevelAccumulator(1, R * R, dev)
        for mode, var in ((0, True), (0, False), (1, False)):
            nat.gram_accumulate(b, rows.permute(2, 0, 1), acc.level(0), mode=mode, want_var=var)
        nat.gram_accumulate(b, rows.permute(2, 0, 1)[:, :, :1], acc.level(0), mode=0, want_var=True)
        phi = nat.basis_eval(b, rows[:, 0, 0].contiguous(), R)
        lam = torch.zeros(R, dtype=torch.float64, device=dev)
        nat.maxent_fgh(phi, torch.ones(1027, dtype=torch.float64, device=dev), lam)
wide = torch.from_numpy(rng.normal(size=(500, 2, 24))).to(dev)
b = nat.make_basis(nat.LEGENDRE, 9, (-2.5, 2.5), (-1.0, 1.0))
acc = nat.LevelAccumulator(1, 9, dev)
nat.moments_accumulate(b, wide.permute(2, 0, 1)[5:6], acc.level(0))
torch.cuda.synchronize()
print("sanitize target ok")
