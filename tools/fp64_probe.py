import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlmc_b200 import _native as nat
torch.cuda.init()
print("latency cycles/dependent DFMA:", nat.fp64_peak(3))
for base, name in ((0, "DFMA 2 invariant operands"), (2, "DFMA 3 distinct operands"), (1, "DMMA")):
    for w in (0, 4, 8, 12, 16, 32):
        print("%-28s warps/SM %2d : %.2f TFLOP/s" % (name, w or 64, nat.fp64_peak(base + 16 * w) / 1e12))
for n_acc in (1, 2, 4, 8):
    for w in (4, 8, 12, 16):
        print("DMMA, %d accumulator tile(s) per warp, fragments from smem, warps/SM %2d : %.2f TFLOP/s"
              % (n_acc, w, nat.fp64_peak(4 + 16 * w + 4096 * n_acc) / 1e12))
