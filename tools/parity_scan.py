#!/usr/bin/env python
"""Large-sample parity scan on the GPU box: libbbmcu.so against the compiled unmodified reference (oracle/_ref).

For each BSDF string: N random (out, xi) pairs -> sample / eval / pdf on the GPU; the oracle samples the same pairs and
evaluates eval / pdf AT THE GPU's sampled directions ("input sensitivity of sharp lobes").
Reports the number of elements beyond tolerance instead of hiding rare ill-conditioned cases behind a loose bound
(report the residual rate).  Test infrastructure: the product never imports this.

  python tools/parity_scan.py --log2 22 "GGX()" "GGX([0.5,0.5,0.5], 0.01, 1.5)"
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import mismatch, pdf_floor  # noqa: E402


def synth(n, seed):
    rng = np.random.default_rng(seed)
    z = rng.random(n, dtype=np.float32)
    ph = rng.random(n, dtype=np.float32) * np.float32(2 * np.pi)
    s = np.sqrt(np.maximum(1 - z * z, 0)).astype(np.float32)
    out = np.stack([s * np.cos(ph), s * np.sin(ph), z], 1).astype(np.float32)
    nrm = np.sqrt((out.astype(np.float64) ** 2).sum(1))
    out = (out / nrm[:, None]).astype(np.float32)
    xi = rng.random((n, 2), dtype=np.float32)
    return out, xi


def scan(ctx, ref, s, n, seed, threads):
    import bbm_b200 as bb
    out, xi = synth(n, seed)
    b = bb.Bsdf(s)
    so, sx = np.ascontiguousarray(out.T), np.ascontiguousarray(xi.T)
    t = time.perf_counter()
    d, sp, f, rgb, p = ctx.sample_eval_pdf(b, so, sx)
    t_gpu = time.perf_counter() - t
    d_aos = np.ascontiguousarray(d.T)
    rd, rsp, rf = ref.sample(s, out, xi, threads=threads)
    ok = f != 0
    res = {"bsdf": s, "pairs": n, "seed": seed,
           "flag_mismatch": int((f != rf).sum()),
           "dir_beyond_1e-5": int(mismatch(d_aos, rd, 1e-5, 1e-5).any(1).sum()),
           "dir_beyond_1e-4": int(mismatch(d_aos, rd, 1e-4, 1e-4).any(1).sum())}
    want_e = ref.eval(s, d_aos, out, threads=threads)
    want_p = ref.pdf(s, d_aos, out, threads=threads)
    res["eval_at_gpu_dir_beyond_1e-5"] = int(mismatch(rgb.T[ok], want_e[ok], 1e-5, 1e-30).any(1).sum())
    res["pdf_at_gpu_dir_beyond_1e-5"] = int(mismatch(p[ok], want_p[ok], 1e-5, pdf_floor(s, want_p[ok])).sum())
    # sample.pdf == pdf(sample.direction, out) for every model except AshikhminShirleyFull, whose sample() mixes the pdfs
    # of two different directions (ashikhminshirleyfull.h:97-113): that one is judged against the reference's sample.pdf
    res["sample_pdf_vs_pdf_mismatch"] = int(mismatch(sp[ok], want_p[ok], 1e-5, pdf_floor(s, want_p[ok])).sum())
    res["sample_pdf_vs_ref_sample_pdf_beyond_1e-3"] = int(mismatch(sp[ok], rsp[ok], 1e-3, pdf_floor(s, rsp[ok])).sum())
    with np.errstate(invalid="ignore", divide="ignore"):
        err = np.abs(d_aos.astype(np.float64) - rd).max(1)
    res["dir_max_abs_err"] = float(np.nanmax(err))
    res["gpu_host_path_s"] = t_gpu
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("bsdfs", nargs="+")
    ap.add_argument("--log2", type=int, default=22)
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import bbm_b200 as bb
    from oracle import refbind
    ctx = bb.Context(0)
    ref = refbind.Ref("float")
    threads = len(os.sched_getaffinity(0))
    strings = []
    for s in a.bsdfs:
        if s == "@defaults":                                   # every analytic model at its default parameters
            strings += [m + "()" for m in bb.model_names() if m != "Merl"]
        elif s.startswith("@"):                                # the configurations of an earlier scan (its JSON file)
            strings += [r["bsdf"] for r in json.load(open(s[1:]))]
        else:
            strings.append(s)
    rows = []
    for s in strings:
        rows.append(scan(ctx, ref, s, 1 << a.log2, a.seed, threads))
        print(json.dumps(rows[-1]), flush=True)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
