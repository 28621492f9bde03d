#!/usr/bin/env python
"""Pipe utilisation of every model's eval / sample / pdf / fused kernel, from ncu hardware counters (the north star's
"achieved FP32/SFU pipe utilisation for the transcendental-heavy models").  Runs `tools/model_throughput.py --once`
under `ncu --metrics ...` (one launch per (model, op), in model order) and writes one row per launch:
duration, issue-slot utilisation, FMA / ALU / XU (MUFU + conversions) / FP64 / LSU pipe utilisation, DRAM throughput and
bytes, registers, warp instructions per element.  Durations under ncu are cold-cache and serialised: use the ratios.
`python tools/pipe_scan.py --log2 22 --out gpurun_out/pipes.json`  (GPU box only)"""
import argparse
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "smsp__inst_executed.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active"]
SHORT = {"gpu__time_duration.sum": "ns", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
         "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pct",
         "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pct",
         "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
         "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "launch__registers_per_thread": "regs",
         "smsp__inst_executed.sum": "warp_inst", "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct"}
OPS = ["eval", "sample", "pdf", "sample_eval_pdf"]


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2", type=int, default=22)
    ap.add_argument("--out", default="gpurun_out/pipes.json")
    a = ap.parse_args()
    log = a.out + ".csv"
    cmd = ["ncu", "--metrics", ",".join(METRICS), "--clock-control", "none", "-k", "regex:k_foreach4", "--csv", "--log-file", log,
           sys.executable, os.path.join(ROOT, "tools", "model_throughput.py"), "--log2", str(a.log2), "--once"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:])
        sys.exit(r.returncode)
    sys.path.insert(0, ROOT)
    import bbm_b200 as bb
    models = [m for m in bb.model_names() if m != "Merl"]
    rows = [x for x in csv.reader(l for l in open(log) if not l.startswith("=="))]
    hdr = rows[0]
    launches = collections.OrderedDict()
    for x in rows[1:]:
        d = dict(zip(hdr, x))
        e = launches.setdefault(d["ID"], {"kernel": d["Kernel Name"]})
        k = SHORT.get(d["Metric Name"])
        if k:
            val = d["Metric Value"].replace(",", "")
            e[k] = to_bytes(val, d["Metric Unit"]) if k.startswith("dram_r") or k.startswith("dram_w") else float(val)
    ls = list(launches.values())
    assert len(ls) == len(models) * len(OPS), (len(ls), len(models))
    n = 1 << a.log2
    out = []
    for i, e in enumerate(ls):
        m, op = models[i // 4], OPS[i % 4]
        assert ("SampleEvalPdf" in e["kernel"]) == (op == "sample_eval_pdf"), (m, op, e["kernel"][:120])
        pipes = {k: e[k] for k in ("fma_pct", "alu_pct", "xu_pct", "fp64_pct", "lsu_pct")}
        top = max(pipes, key=pipes.get)
        bound = "hbm" if e["dram_pct"] >= max(e["issue_pct"], pipes[top]) else ("issue" if e["issue_pct"] >= 1.25 * pipes[top] else top[:-4])
        out.append({"model": m, "op": op, "ns": e["ns"], "issue_pct": e["issue_pct"], **pipes, "dram_pct": e["dram_pct"],
                    "dram_bytes_per_element": (e["dram_read"] + e["dram_write"]) / n, "warp_inst_per_element": e["warp_inst"] / n,
                    "regs": e["regs"], "occupancy_pct": e["occupancy_pct"], "binding": bound,
                    "binding_pct": max(e["dram_pct"], e["issue_pct"], pipes[top])})
        print("%-24s %-16s issue %5.1f  fma %5.1f  alu %5.1f  xu %5.1f  fp64 %5.1f  lsu %5.1f  dram %5.1f  -> %s %.0f%%" % (
            m, op, e["issue_pct"], e["fma_pct"], e["alu_pct"], e["xu_pct"], e["fp64_pct"], e["lsu_pct"], e["dram_pct"], bound, out[-1]["binding_pct"]))
    json.dump({"elements": n, "note": "ncu counters, one launch per row; durations are cold-cache", "rows": out}, open(a.out, "w"), indent=1)
    os.remove(log)


if __name__ == "__main__":
    main()
