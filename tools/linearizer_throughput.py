#!/usr/bin/env python
"""Throughput of the linearizer kernels on one GPU (device-resident SoA buffers, CUDA events on the library's stream):
merl_index (direction pair -> bin, 24 B in + 4 B out), merl_dirs (bin -> direction pair, 24 B out) and spherical_dirs.
   python tools/linearizer_throughput.py [--log2 24] [--out f.json]"""
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
    torch.cuda.synchronize()

    def timed(fn, reps=10):
        fn(); fn()
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        ctx.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3
    res = {"elements": n}
    idx = torch.empty(n, device=dev, dtype=torch.int32)
    t = timed(lambda: ctx.merl_index(inn, out, index=idx))
    res["merl_index_G_per_s"] = n / t / 1e9
    res["merl_index_GBs_at_28B"] = 28 * n / t / 1e9
    m = bb.MERL_BINS
    gi, go = ctx.merl_dirs(0, m, like=out)
    t = timed(lambda: ctx.merl_dirs(0, m, outputs=(gi, go)), reps=50)
    res["merl_dirs_G_per_s"] = m / t / 1e9
    print(json.dumps(res))
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
