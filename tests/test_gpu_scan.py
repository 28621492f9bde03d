"""GPU: the large-sample parity scan of tools/parity_scan.py as a driver-run test.

2^21 random (out, xi) pairs per configuration - all 34 models at their defaults and 17 fitted configurations with sharp
lobes (tests/golden/scan_configs.json) - sampled on the GPU and by the compiled unmodified reference (oracle/_ref); eval and
pdf are compared at the GPU's own sampled directions.  The contract is 1e-5; the budget of every count
is ZERO, except where a comment below says why not.  The golden fixtures hold 192 samples per case and cannot see residuals
at the 1e-6 level; this test does."""
import json
import os

import numpy as np
import pytest

from tools.parity_scan import scan

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOG2 = 21

# directions beyond 1e-5 allowed per configuration (of 2^21).  Everything not listed: 0.
# EPD: tan^2(theta) = beta^2 gamma_q_inv(1/p, xi)^(1/p) (ndf/epd.h:98) raises the float error of the inverse incomplete gamma
# function to the power 1/p = 5; the inverse runs on the device's logf / lgammaf / powf (1-2 ulp, not restated from the host
# libm), so about one pair in 2^21 lands just past the contract (1.3e-5 seen).  Bounded below: no more than 3, none beyond 3e-5.
DIR_BUDGET = {"EPD()": 3}
DIR_MAX_ERR = {"EPD()": 3e-5}
# ngan_lafortune.fit:alum-bronze evaluates to NaN weights in the reference itself (the Ngan normalisation underflows,
# SURVEY.md fact 10) and the reference's aggregate then returns an UNINITIALISED sample (fact 16): nothing to compare
SKIP_SUBSTR = ("NganLafortune(albedo = [0.322, 0.193, 0.105], Cxy = -0.579, Cz = 0.574, sharpness = 630)",)


def _configs():
    import bbm_b200 as bb
    cfg = [m + "()" for m in bb.model_names() if m != "Merl"]
    cfg += json.load(open(os.path.join(ROOT, "tests", "golden", "scan_configs.json")))["fitted"]
    return cfg


def test_parity_scan_all_models_and_fitted_configurations(ctx, ref):
    threads = len(os.sched_getaffinity(0))
    rows, failures = [], []
    for s in _configs():
        if any(k in s for k in SKIP_SUBSTR):
            continue
        r = scan(ctx, ref, s, 1 << LOG2, 11, threads)
        rows.append(r)
        for key, budget in (("flag_mismatch", 0), ("dir_beyond_1e-5", DIR_BUDGET.get(s, 0)), ("eval_at_gpu_dir_beyond_1e-5", 0), ("pdf_at_gpu_dir_beyond_1e-5", 0)):
            if r[key] > budget:
                failures.append((s, key, r[key], r.get("dir_max_abs_err")))
        if r["dir_beyond_1e-5"] and r["dir_max_abs_err"] > DIR_MAX_ERR.get(s, 0.0):
            failures.append((s, "dir_max_abs_err", r["dir_max_abs_err"]))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        json.dump(rows, open(os.path.join(out, "parity_scan_r02.json"), "w"), indent=1)
    assert len(rows) >= 50
    assert not failures, failures
