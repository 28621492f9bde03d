#!/usr/bin/env python
"""Launch-shape sweep of the batched loss kernel on ONE GPU: time per batch for shard sizes N, N/2, N/4, N/8 (what each rank
of a 1/2/4/8-GPU run owns) at several targets of resident blocks per SM (BBMCU_LOSS_BLOCKS_PER_SM, read once per process,
hence one subprocess per value).  Prints kernel-only passes/s and the strong-scaling efficiency the kernel alone allows.
   python tools/loss_shape_sweep.py [--fills 2,4,8,16,32] [--out f.json]"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(K, interleaved=False):
    import numpy as np
    import torch
    import bbm_b200 as bb
    ctx = bb.Context(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    fitted = bb.Bsdf("Aggregate(Lambertian(), CookTorrance())")
    truth = bb.Bsdf("Aggregate(Lambertian([0.2,0.1,0.05]), CookTorrance([0.3,0.3,0.3], 0.2, 1.5))")
    rng = np.random.default_rng(7)
    p0 = fitted.parameter_values()
    P = len(p0)
    out = {}
    for k in K:
        params = p0[None] * (1 + 0.1 * rng.random((k, P)))
        params[:, 7] = 1.2 + rng.random(k)
        res = torch.zeros((k, 1 + P), device=dev, dtype=torch.float64)
        for div in (1, 2, 4, 8):
            count = (bb.MERL_BINS + div - 1) // div
            # contiguous: the first 1/div of the grid (the most expensive contiguous shard: no pair below the horizon);
            # interleaved: blocks 0, div, 2 div, ... of 1024 samples (BBMCU_LOSS_SHARD_INTERLEAVED: what bench.py uses)
            L = ctx.loss("nganL2", truth, None, interleaved=(0, div)) if interleaved else ctx.loss("nganL2", truth, None, first=0, count=count)
            for _ in range(3):
                L.eval_device(fitted, params, res)
            ctx.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record(stream)
            for _ in range(reps):
                L.eval_device(fitted, params, res)
            e1.record(stream)
            ctx.synchronize()
            out["K%d_div%d" % (k, div)] = e0.elapsed_time(e1) / reps
            del L
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fills", default="2,4,8,16,32")
    ap.add_argument("--K", default="16,256")
    ap.add_argument("--out", default=None)
    ap.add_argument("--worker", action="store_true")
    ap.add_argument("--small-rounds", action="store_true", help="sweep BBMCU_LOSS_SMALLK_ROUNDS (rounds of resident blocks for launches with < 8 sets per block) with --fills values")
    ap.add_argument("--interleaved", action="store_true", help="shards dealt in blocks of 1024 samples instead of contiguous ranges")
    a = ap.parse_args()
    K = [int(x) for x in a.K.split(",")]
    if a.worker:
        return worker(K, a.interleaved)
    table = {}
    for f in a.fills.split(","):
        env = dict(os.environ, BBMCU_LOSS_BLOCKS_PER_SM=f)
        if a.small_rounds:
            env = dict(os.environ, BBMCU_LOSS_SMALLK_ROUNDS=f)      # --small-rounds: the swept value is the small-K round count instead
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", "--K", a.K] + (["--interleaved"] if a.interleaved else []), env=env, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stderr[-2000:])
            sys.exit(1)
        ms = json.loads(r.stdout.strip().splitlines()[-1])
        table[f] = ms
        for k in K:
            base = ms["K%d_div1" % k]
            print("fill %3s  K %4d  " % (f, k) + "  ".join("1/%d: %7.1f us (eff %3.0f%%)" % (d, 1e3 * ms["K%d_div%d" % (k, d)], 100 * base / d / ms["K%d_div%d" % (k, d)]) for d in (1, 2, 4, 8)), flush=True)
    if a.out:
        json.dump({"ms_per_batch": table, "note": "kernel + finish + parameter upload, one GPU, shard = N/div samples"}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
