#!/usr/bin/env python
"""BASELINE configs[4] in small: synthetic MERL-shaped materials x ALL 34 models x the 6 metrics, batched compass search
with a fixed step budget, on this rank's share (torchrun: one process per GPU, jobs partitioned by cost, no collective).
Prints fits/s and the time per model.  `python tools/sweep_bench.py --materials 2 --steps 20`"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--materials", type=int, default=2)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import bbm_b200 as bb
    from bbm_b200.fit import fit, fitted_string_for, sweep_jobs
    from bbm_b200.shard import partition_by_cost
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    ctx = bb.Context(int(os.environ.get("LOCAL_RANK", 0)))
    i, o = ctx.merl_dirs(0, bb.MERL_BINS)
    rng = np.random.default_rng(1)
    tables = {}
    for m in range(a.materials):
        d, s = rng.random(3) * 0.3, rng.random(3) * 0.3 + 0.05
        truth = f"Aggregate(Lambertian([{d[0]:.4f}, {d[1]:.4f}, {d[2]:.4f}]), LowCookTorrance([{s[0]:.4f}, {s[1]:.4f}, {s[2]:.4f}], {0.05 + 0.3*rng.random():.4f}, {1.2 + rng.random():.4f}))"
        tables[f"mat{m}"] = ctx.eval(bb.Bsdf(truth), i, o)
    models = [n for n in bb.model_names() if n != "Merl"]
    jobs, cost = sweep_jobs(sorted(tables), models)
    mine = partition_by_cost(cost, world)[rank]
    per_model = {}
    t0 = time.perf_counter()
    for j in mine:
        mat, mod, met = jobs[j]
        t = time.perf_counter()
        b, trace = fit(ctx, fitted_string_for(mod), tables[mat], met, None, a.steps)
        per_model.setdefault(mod, []).append(time.perf_counter() - t)
    dt = time.perf_counter() - t0
    rows = {m: float(np.mean(v)) for m, v in per_model.items()}
    res = {"rank": rank, "world": world, "fits": len(mine), "seconds": dt, "fits_per_s": len(mine) / dt, "compass_steps_per_fit": a.steps,
           "samples_per_loss": bb.MERL_BINS, "seconds_per_fit_by_model": rows}
    print(json.dumps(res))
    if a.out and rank == 0:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
