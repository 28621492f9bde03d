#!/usr/bin/env python
"""Kernel throughput of run-time aggregates (the shape of every entry of the reference's fits/*.fit: Lambertian + one
specular lobe) next to the single-model kernels, device-resident buffers, CUDA events on the library's stream.
   python tools/aggregate_throughput.py [--log2 24] [--out f.json]"""
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
    ap.add_argument("--out", default=None)
    ap.add_argument("--models", default="CookTorrance,GGX,Ward,Phong,LowMicrofacet,AshikhminShirley")
    a = ap.parse_args()
    import torch
    import bbm_b200 as bb
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

    def timed(fn, reps=5):
        fn(); fn()
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        ctx.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3
    rows = []
    for m in a.models.split(","):
        for s in (m + "()", "Aggregate(Lambertian(), %s())" % m):
            b = bb.Bsdf(s)
            r = {"bsdf": s,
                 "eval_G_per_s": n / timed(lambda: ctx.eval(b, inn, out, rgb=rgb)) / 1e9,
                 "sample_G_per_s": n / timed(lambda: ctx.sample(b, out, xi, outputs=(d, sp, f))) / 1e9,
                 "pdf_G_per_s": n / timed(lambda: ctx.pdf(b, inn, out, pdf=pdf)) / 1e9,
                 "sample_eval_pdf_G_per_s": n / timed(lambda: ctx.sample_eval_pdf(b, out, xi, outputs=(d, sp, f, rgb, pdf))) / 1e9}
            rows.append(r)
            print("%-44s eval %7.2f  sample %7.2f  pdf %7.2f  fused %7.2f  G/s" % (s, r["eval_G_per_s"], r["sample_G_per_s"], r["pdf_G_per_s"], r["sample_eval_pdf_G_per_s"]), flush=True)
    if a.out:
        json.dump({"elements": n, "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
