#!/usr/bin/env python
"""Summarise an .ncu-rep (read here with `ncu -i`): headline metrics -> CSV for profiles/, plus the instructions
executed per source line (needs -lineinfo and --import-source on).  `python tools/ncu_summary.py rep.ncu-rep out.csv [units]`
where `units` = work items in the launch (pairs, sample x parameter-set ...) to print instructions per item."""
import collections
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores']


def main():
    rep, out = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, unit = rows[0], rows[1]
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "value", "unit"])
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            for k in WANT:
                if k in d:
                    w.writerow([d["Kernel Name"][:90], k, d[k], unit[hdr.index(k)]])
                    print(k, d[k], unit[hdr.index(k)])
            for k in hdr:
                if "issue_stalled" in k and "per_issue_active" in k:
                    w.writerow([d["Kernel Name"][:90], k, d[k], ""])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur, h, agg = None, None, collections.Counter()
    for r in csv.reader(src.split("\n")):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1]
        elif r[0] == "Line No":
            h = r
        elif h and len(r) == len(h):
            d = dict(zip(h, r))
            try:
                agg[(cur.split("/")[-1], int(d["Line No"]))] += int(d.get("Instructions Executed", "0"))
            except ValueError:
                pass
    tot = sum(agg.values())
    per = (units / 32.0) if units else None
    print("warp instructions by source line:", tot, ("= %.1f per unit" % (tot / per)) if per else "")
    for (f_, l), v in agg.most_common(45):
        try:
            text = open("/root/repo/bbm_b200/csrc/" + f_).read().split("\n")[l - 1].strip()[:100]
        except Exception:
            text = ""
        print("%-24s %5d %7.1f%% %s  %s" % (f_, l, 100.0 * v / tot, ("%6.1f" % (v / per)) if per else "", text))


if __name__ == "__main__":
    main()
