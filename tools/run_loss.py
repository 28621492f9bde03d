#!/usr/bin/env python
"""Run the batched loss + gradient of BASELINE configs[2] a few times (for `ncu -k regex:k_loss_tile`):
   python tools/run_loss.py [K] [metric] [div]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bbm_b200 as bb  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 256
metric = sys.argv[2] if len(sys.argv) > 2 else "nganL2"
div = int(sys.argv[3]) if len(sys.argv) > 3 else 1          # shard = first 1/div of the grid (what one rank of a div-GPU run owns)
ctx = bb.Context(0)
fitted = bb.Bsdf("Aggregate(Lambertian(), CookTorrance())")
truth = bb.Bsdf("Aggregate(Lambertian([0.2,0.1,0.05]), CookTorrance([0.3,0.3,0.3], 0.2, 1.5))")
L = ctx.loss(metric, truth, None, first=0, count=(bb.MERL_BINS + div - 1) // div)
rng = np.random.default_rng(7)
p0 = fitted.parameter_values()
params = p0[None] * (1 + 0.1 * rng.random((K, len(p0))))
params[:, 7] = 1.2 + rng.random(K)
res = torch.zeros((K, 1 + len(p0)), device="cuda:0", dtype=torch.float64)
for _ in range(3):
    L.eval_device(fitted, params, res)
ctx.synchronize()
print("ok", K, metric, float(res[0, 0].item()))
