"""Summarise an .ncu-rep: key raw metrics of the first kernel and the hottest source lines.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--lines 40]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, "--csv"] + list(args), capture_output=True, text=True).stdout


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 40
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw"))))
    hdr, units, first = rows[0], rows[1], rows[2]
    for i, h in enumerate(hdr):
        if h in KEYS or ("issue_stalled" in h and "per_issue_active" in h):
            print("%-90s %-10s %s" % (h, units[i], first[i]))
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--print-source", "cuda,sass"))))
    cur, agg = None, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 8 and r[0] not in ("", "Line No"):
            agg[(cur, int(r[0]))] = (r[1], num(r[7]), num(r[4]))
    tot = sum(v[1] for v in agg.values()) or 1
    ts = sum(v[2] for v in agg.values()) or 1
    print("total warp instructions (first kernel in the source page): %d, samples %d" % (tot, ts))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:nlines]:
        print("%s:%d  inst %5.1f%%  samples %5.1f%%  %s" % (k[0], k[1], 100 * v[1] / tot, 100 * v[2] / ts, v[0][:90]))


if __name__ == "__main__":
    main()
