#!/usr/bin/env python
"""Time the batched loss(+gradient) pass of BASELINE configs[2] with the compact pair kernel and with the generic tile
kernel (BBMCU_LOSS_NO_COMPACT=1) on the same loss object: K parameter sets per launch, CUDA events on the library's stream.
   python tools/loss_ab.py [--out file.json] [--div D]      (div: the shard one rank of a D-GPU run owns)"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bbm_b200 as bb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--div", type=int, default=1)
ap.add_argument("--models", default="CookTorrance,NganCookTorrance")
ap.add_argument("--ks", default="1,16,256", help="parameter sets per launch")
ap.add_argument("--reps", type=int, default=0, help="timed launches per measurement (0: 20, or 100 at K < 16)")
ap.add_argument("--single", action="store_true", help="the lobe alone instead of Aggregate(Lambertian(), lobe)")
ap.add_argument("--grid", default="merl", help="merl, or offhorizon: a 31 x 12 x 6 x 9 spherical grid that stays off the horizon (Ward lobes)")
args = ap.parse_args()
ctx = bb.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda:0"))
truth = bb.Bsdf("Aggregate(Lambertian([0.2,0.1,0.05]), CookTorrance([0.3,0.3,0.3], 0.2, 1.5))")
rows = []
for model in args.models.split(";" if (";" in args.models or "(" in args.models) else ","):
    spec = model if "(" in model else model + "()"
    fitted = bb.Bsdf(spec if args.single else "Aggregate(Lambertian(), %s)" % spec)
    p0 = fitted.parameter_values()
    for metric in ("nganL2", "standardLog"):
        if args.grid == "merl":
            L = ctx.loss(metric, truth, None, first=0, count=(bb.MERL_BINS + args.div - 1) // args.div)
        else:
            hp = float(np.float32(2) * np.float32(np.pi))
            L = ctx.loss(metric, truth, bb.spherical_grid((62, 24), (24, 18), start_in=(0, 0.05), start_out=(0, 0.05), end_in=(hp, 1.4), end_out=(hp, 1.4)))
        for K in [int(k) for k in args.ks.split(",")]:
            rng = np.random.default_rng(7)
            params = p0[None] * (1 + 0.1 * rng.random((K, len(p0))))
            res = torch.zeros((K, 1 + len(p0)), device="cuda:0", dtype=torch.float64)
            row = {"model": model.split("(")[0], "metric": metric, "K": K, "shard": "1/%d" % args.div}
            for name, env in (("compact", None), ("generic", "1")):
                if env:
                    os.environ["BBMCU_LOSS_NO_COMPACT"] = env
                for grad in (True, False):
                    reps = args.reps or (20 if K >= 16 else 100)
                    for _ in range(3):
                        L.eval_multi_device(fitted, params[None], res, grad=grad)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    for _ in range(reps):
                        L.eval_multi_device(fitted, params[None], res, grad=grad)
                    e1.record(stream)
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / reps
                    row["%s_%s_us" % (name, "grad" if grad else "value")] = round(ms * 1e3, 2)
                    row["%s_%s_passes_per_s" % (name, "grad" if grad else "value")] = round(K / (ms * 1e-3))
                if env:
                    del os.environ["BBMCU_LOSS_NO_COMPACT"]
            row["speedup_grad"] = round(row["generic_grad_us"] / row["compact_grad_us"], 3)
            rows.append(row)
            print(json.dumps(row), flush=True)
if args.out:
    json.dump(rows, open(args.out, "w"), indent=1)
