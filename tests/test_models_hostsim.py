"""CPU: the device model code (compiled for the host) against the golden vectors produced by the
unmodified reference - every model x {eval, pdf, reflectance, sample} x component."""
import numpy as np
import pytest

from tests.util import assert_parity, uses_only_implemented

COMPONENTS = (3, 1, 2)


def defined_samples(bsdf_string, reflectance, xi):
    """The reference's run-time aggregate returns an UNINITIALISED BsdfSample when the lobe weights sum
    to <= epsilon (include/bbm/aggregatebsdf.h:104,116-117); those elements of the
    golden vectors are stack garbage and are not compared (we return {0, 0, None} there).  The same
    happens at xi0 == 1 when (w0 + w1) - w0 > w1 in float: no lobe accepts the residual."""
    if not bsdf_string.startswith("Aggregate"):
        return np.ones(len(reflectance), bool)
    return (reflectance.astype(np.float32).sum(1) > np.float32(1.1920929e-07)) & (xi[:, 0] < 1)


def _cases(meta):
    return sorted(meta["cases"].items(), key=lambda kv: int(kv[0][4:]))


def test_all_cases_eval_pdf_reflectance(hostsim, golden_models):
    arr, meta = golden_models
    inn, out = arr["in"], arr["out"]
    checked = 0
    for key, rec in _cases(meta):
        s = rec["string"]
        if not uses_only_implemented(s):
            continue
        for c in COMPONENTS:
            assert_parity(hostsim.eval(s, inn, out, c), arr[f"{key}_eval_c{c}"], 1e-5, what=f"eval {s} comp {c}")
            assert_parity(hostsim.pdf(s, inn, out, c), arr[f"{key}_pdf_c{c}"], 1e-5, what=f"pdf {s} comp {c}")
            assert_parity(hostsim.reflectance(s, out, c), arr[f"{key}_refl_c{c}"], 1e-5, what=f"reflectance {s} comp {c}")
        checked += 1
    assert checked >= 60


def test_all_cases_sample(hostsim, golden_models):
    arr, meta = golden_models
    out, xi = arr["out"], arr["xi"]
    for key, rec in _cases(meta):
        s = rec["string"]
        if not uses_only_implemented(s):
            continue
        for c in COMPONENTS:
            d, p, f = hostsim.sample(s, out, xi, c)
            want_d, want_p, want_f = arr[f"{key}_sdir_c{c}"], arr[f"{key}_spdf_c{c}"], arr[f"{key}_sflag_c{c}"]
            ok = defined_samples(s, arr[f"{key}_refl_c{c}"], xi)
            d, p, f, want_d, want_p, want_f = d[ok], p[ok], f[ok], want_d[ok], want_p[ok], want_f[ok]
            assert np.array_equal(f, want_f.astype(np.int32)), f"sample flag {s} comp {c}"
            # sampled directions: 1e-5 relative to the unit vector
            assert_parity(d, want_d, 1e-5, floor=1e-5, what=f"sample dir {s} comp {c}")
            # the pdf of a sharp lobe amplifies last-bit differences of the direction; the strict pdf
            # parity is test_all_cases_eval_pdf_reflectance (same directions on both sides)
            assert_parity(p, want_p, 3e-2, what=f"sample pdf {s} comp {c}")


def test_studentt_factors_formed_once_equal_factors_per_evaluation(hostsim, golden_models):
    """NdfStudentT::pre / G1_pre (the eval kernels form the gamma-only factors of G1 once per thread, ndf/studentt.h:110-140)
    against the per-evaluation route (eval<float>, the run-time lobe list) bit for bit, and against the golden vectors; the
    fused pass with the factors formed once against sample + eval + pdf of the run-time lobe list"""
    arr, meta = golden_models
    inn, out, xi = arr["in"], arr["out"], arr["xi"]
    checked = 0
    for key, rec in _cases(meta):
        s = rec["string"]
        if not s.startswith("Ribardiere"):
            continue
        for c in COMPONENTS:
            e = hostsim.eval_pre(s, inn, out, c)
            assert np.array_equal(e.view(np.uint32), hostsim.eval(s, inn, out, c).view(np.uint32)), (s, c)
            assert_parity(e, arr[f"{key}_eval_c{c}"], 1e-5, what=f"eval (factors once) {s} comp {c}")
            d, sp, f, rgb, p = hostsim.sample_eval_pdf_pre(s, out, xi, c)
            d0, p0, f0 = hostsim.sample(s, out, xi, c)
            assert np.array_equal(f, f0) and np.array_equal(d.view(np.uint32), d0.view(np.uint32)), (s, c)
            assert np.array_equal(rgb.view(np.uint32), hostsim.eval(s, d, out, c).view(np.uint32)), (s, c)
            assert np.array_equal(p.view(np.uint32), hostsim.pdf(s, d, out, c).view(np.uint32)), (s, c)
            assert np.array_equal(sp[f != 0].view(np.uint32), p0[f != 0].view(np.uint32)), (s, c)
        checked += 1
    assert checked >= 2
