"""Condense an `ncu --page raw --csv` export into the metrics the design notes quote: one row per captured launch.
usage: python tools/ncu_summary.py gpurun_out/x_raw.csv profiles/x_summary.csv"""
import csv
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg"]


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch %d" % k for k in range(len(data))])
        for k in KEEP:
            if k in ci:
                w.writerow([k, units[ci[k]]] + [r[ci[k]] for r in data])
        for h in stalls:
            vals = [r[ci[h]] for r in data]
            try:
                if max(float(v) for v in vals) < 0.05:
                    continue
            except ValueError:
                continue
            w.writerow(["stall_" + h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")] +
                        " (warps per issue)", ""] + vals)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
