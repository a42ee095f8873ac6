"""SASS evidence (no GPU needed): per object of libmlmcb200, the mnemonics that prove the Blackwell-side features --
DMMA.8x8x4 (FP64 tensor core; tcgen05 has no f64 kind, so no UTC*MMA is expected), UBLKCP (TMA bulk copy,
cp.async.bulk), SYNCS.* (mbarrier), DFMA / DMUL / DADD (FP64 pipe), LDG.E.*.128 (vectorised loads), SHFL, and the
register / spill figures from ptxas.  Writes profiles/r2_sass_counts.txt."""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mlmc_b200", "_lib")
WATCH = ["DMMA", "UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "DSETP", "LDG", "LDS", "STS", "SHFL", "BAR", "UTCHMMA", "HMMA",
         "ATOM", "RED", "MUFU"]


def main():
    out = ["SASS mnemonic counts per object (cuobjdump -sass, sm_100a), %s" % subprocess.run(
        ["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1]]
    for obj in sorted(f for f in os.listdir(LIB) if f.endswith(".o")):
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(LIB, obj)], capture_output=True, text=True).stdout
        n_fun = sass.count("Function :")
        ops = Counter()
        for m in re.finditer(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P[T0-9]+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", sass, flags=re.M):
            op, mods = m.group(1), m.group(2)
            ops[op] += 1
            if op == "DMMA":
                ops["DMMA" + mods] += 1
            if op == "LDG" and ".128" in mods:
                ops["LDG.128"] += 1
            if op == "SYNCS":
                ops["SYNCS" + mods] += 1
        line = "%-14s %3d kernels | " % (obj, n_fun) + "  ".join(
            "%s %d" % (k, v) for k, v in sorted(ops.items()) if any(k.startswith(w) for w in WATCH) and v)
        out.append(line)
        log = os.path.join(LIB, obj.replace(".o", ".cu.ptxas.log"))
        if os.path.exists(log):
            txt = open(log).read()
            regs = [int(v) for v in re.findall(r"Used (\d+) registers", txt)]
            spills = [int(v) for v in re.findall(r"(\d+) bytes spill stores", txt)]
            out.append("%-14s registers min/max %d/%d, kernels with spills %d (max %d B)" % (
                "", min(regs), max(regs), sum(1 for v in spills if v), max(spills) if spills else 0))
    text = "\n".join(out) + "\n"
    with open(os.path.join(ROOT, "profiles", "r2_sass_counts.txt"), "w") as f:
        f.write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
