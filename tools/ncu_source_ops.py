"""Aggregate the warp-state samples of an `ncu --page source --csv` export by SASS opcode and stall reason.
usage: python tools/ncu_source_ops.py gpurun_out/x_src.csv"""
import collections
import csv
import sys

STALLS = ("stall_barrier", "stall_math", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_not_selected",
          "stall_selected", "stall_mio", "stall_lg", "stall_branch_resolving", "stall_dispatch", "stall_no_inst")


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if "# Samples" in r)
    ci = {h: i for i, h in enumerate(hdr)}
    tot, by, exe, stall = 0, collections.Counter(), collections.Counter(), collections.Counter()
    per_op_stall = collections.defaultdict(collections.Counter)
    for r in rows:
        if len(r) < len(hdr) or not r[ci["# Samples"]].isdigit():
            continue
        parts = r[ci["Source"]].split()
        if not parts:
            continue
        op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
        n = int(r[ci["# Samples"]])
        tot += n
        by[op] += n
        exe[op] += int(r[ci["Instructions Executed"]] or 0)
        for k in STALLS:
            v = int(r[ci[k]] or 0)
            stall[k] += v
            per_op_stall[op][k] += v
    print(path, "total samples", tot)
    for op, n in by.most_common(14):
        top = ", ".join("%s %.1f" % (k[6:], 100.0 * v / tot) for k, v in per_op_stall[op].most_common(3) if v)
        print("  %-10s %7d %5.1f%%  executed %10d   [%s]" % (op, n, 100.0 * n / tot, exe[op], top))
    print("  stalls:", {k[6:]: round(100.0 * v / tot, 1) for k, v in stall.most_common()})


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
