"""One covariance launch per mode at R = 50 (1e6 samples) for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlmc_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
R = int(os.environ.get("PROBE_R", "50"))
n = 1_000_000
g = torch.Generator(device=dev).manual_seed(2)
x0 = torch.randn(n, generator=g, device=dev, dtype=torch.float64)
root = torch.sqrt(1e-4 + x0.abs())
rows = torch.stack([x0 + 0.05 * root, x0 + 0.5 * root], dim=1).unsqueeze(2).contiguous()
x = rows.permute(2, 0, 1)
basis = nat.make_basis(nat.LEGENDRE, R, (-3.72, 3.72), (-1.0, 1.0))
acc = nat.LevelAccumulator(1, R * R, dev)
for rep in range(2):
    nat.gram_accumulate(basis, x, acc.level(0), mode=0, want_var=False)
    nat.gram_accumulate(basis, x, acc.level(0), mode=0, want_var=True)
torch.cuda.synchronize()
print("ok")
