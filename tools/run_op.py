#!/usr/bin/env python
"""Run one batched op of one BSDF a few times on device buffers (for `ncu -k regex:k_foreach4`):
   python tools/run_op.py "He()" eval 22"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bbm_b200 as bb  # noqa: E402

s, op, log2 = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 22
ctx = bb.Context(0)
dev = torch.device("cuda", 0)
n = 1 << log2
g = torch.Generator(device=dev).manual_seed(3)


def hemi():
    z = torch.rand(n, device=dev, generator=g)
    ph = torch.rand(n, device=dev, generator=g) * (2 * np.pi)
    r = torch.sqrt(1 - z * z)
    return torch.stack([r * torch.cos(ph), r * torch.sin(ph), z]).contiguous()


inn, out, xi = hemi(), hemi(), torch.rand((2, n), device=dev, generator=g)
b = bb.Bsdf(s)
torch.cuda.synchronize()
for _ in range(4):
    if op == "eval":
        ctx.eval(b, inn, out)
    elif op == "sample":
        ctx.sample(b, out, xi)
    elif op == "pdf":
        ctx.pdf(b, inn, out)
    else:
        ctx.sample_eval_pdf(b, out, xi)
ctx.synchronize()
print("ok", s, op, n)
