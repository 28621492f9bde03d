#!/usr/bin/env python
"""Per-model kernel throughput on one GPU: eval (36 B/element), sample (40 B), pdf (28 B) and the fused
sample+eval+pdf pass (56 B) for all 34 analytic models at default parameters, device-resident SoA buffers, CUDA events
on the library's stream.  `python tools/model_throughput.py [--log2 24] [--out profiles/....json]`"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2", type=int, default=24)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=None)
    ap.add_argument("--once", action="store_true", help="one launch per (model, op), no warm-up: the form tools/pipe_scan.py profiles")
    a = ap.parse_args()
    import torch
    import bbm_b200 as bb
    peak = 6549.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    ctx = bb.Context(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    n = 1 << a.log2
    g = torch.Generator(device=dev).manual_seed(3)

    def hemi():
        z = torch.rand(n, device=dev, generator=g)
        ph = torch.rand(n, device=dev, generator=g) * (2 * np.pi)
        s = torch.sqrt(1 - z * z)
        return torch.stack([s * torch.cos(ph), s * torch.sin(ph), z]).contiguous()
    inn, out = hemi(), hemi()
    xi = torch.rand((2, n), device=dev, generator=g)
    rgb, pdf = torch.empty((3, n), device=dev), torch.empty(n, device=dev)
    d, sp, f = torch.empty((3, n), device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev, dtype=torch.int32)
    torch.cuda.synchronize()

    def timed(fn):
        if a.once:
            fn()
            ctx.synchronize()
            return 1.0
        fn(); fn()
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.reps):
            fn()
        e1.record(stream)
        ctx.synchronize()
        return e0.elapsed_time(e1) / a.reps * 1e-3

    rows = []
    for name in bb.model_names():
        if name == "Merl":
            continue
        b = bb.Bsdf(name + "()")
        t_e = timed(lambda: ctx.eval(b, inn, out, rgb=rgb))
        t_s = timed(lambda: ctx.sample(b, out, xi, outputs=(d, sp, f)))
        t_p = timed(lambda: ctx.pdf(b, inn, out, pdf=pdf))
        t_f = timed(lambda: ctx.sample_eval_pdf(b, out, xi, outputs=(d, sp, f, rgb, pdf)))
        row = {"model": name, "eval_G_per_s": n / t_e / 1e9, "eval_frac_hbm": 36 * n / t_e / 1e9 / peak,
               "sample_G_per_s": n / t_s / 1e9, "sample_frac_hbm": 40 * n / t_s / 1e9 / peak,
               "pdf_G_per_s": n / t_p / 1e9, "pdf_frac_hbm": 28 * n / t_p / 1e9 / peak,
               "sample_eval_pdf_G_per_s": n / t_f / 1e9, "sample_eval_pdf_frac_hbm": 56 * n / t_f / 1e9 / peak}
        rows.append(row)
        print("%-24s eval %7.2f G/s (%4.1f%%)  sample %7.2f (%4.1f%%)  pdf %7.2f (%4.1f%%)  fused %7.2f (%4.1f%%)" % (
            name, row["eval_G_per_s"], 100 * row["eval_frac_hbm"], row["sample_G_per_s"], 100 * row["sample_frac_hbm"],
            row["pdf_G_per_s"], 100 * row["pdf_frac_hbm"], row["sample_eval_pdf_G_per_s"], 100 * row["sample_eval_pdf_frac_hbm"]), flush=True)
    if a.out:
        json.dump({"elements": n, "hbm_peak_gbs": peak, "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
