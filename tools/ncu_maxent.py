"""ncu target: a few max-ent F+g+H evaluations (Q nodes x R moments from argv)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlmc_b200 import _native as nat  # noqa: E402

Q, R = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
phi = torch.randn(Q, R, device=dev, dtype=torch.float64) * 0.1
w = torch.full((Q,), 1.0 / Q, device=dev, dtype=torch.float64)
lam = torch.randn(R, device=dev, dtype=torch.float64) * 0.1
res = torch.zeros(1 + R + R * R, device=dev, dtype=torch.float64)
for _ in range(4):
    nat.maxent_fgh(phi, w, lam, 7, res)
torch.cuda.synchronize()
print("ok", float(res[0]))
