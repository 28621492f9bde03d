"""checkBsdf on the CUDA backbone: the six tests of the reference's bin/checkBsdf.cpp (:51 reflectance, :102 reciprocity,
:157 adjoint, :206 pdf, :272 pdfInt, :321 sample; dispatch :470-474) with the sample loops running as fused
generate -> evaluate -> reduce kernels (bbmcu_check_* of include/bbmcu.h).

Command line, same keywords and printed lines as the reference tool:

    python -m bbm_b200.check bsdfmodel="GGX()" test=pdf samples=100000 checkBelowHorizon

plus   rng=mt19937|philox   (default mt19937: the reference's own random stream, so the numbers can be compared with the
                             reference's output for the same command; philox draws on the device - use it for large runs)
       seed=N device=N

`bbm_info` (bin/bbm_info.cpp): python -m bbm_b200.check info"""
import ctypes as C
import math
import sys

import numpy as np

from . import Bsdf, Context, BbmError, _check, lib, model_names

RNG = {"philox": 0, "mt19937": 1}
_U64, _I = C.c_uint64, C.c_int


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


def reflectance(ctx, bsdf, samples=100000, theta=1, importance=False, rng="mt19937", seed=0):
    """testReflectance: -> (out[theta,3], estimate[theta,3], reflectance[theta,3])"""
    est, ref, out = np.zeros((theta, 3)), np.zeros((theta, 3), np.float32), np.zeros((theta, 3), np.float32)
    _check(lib().bbmcu_check_reflectance(ctx._h, bsdf._h, _U64(samples), _I(theta), _I(int(importance)), _I(RNG[rng]), _U64(seed), _dp(est), _dp(ref), _dp(out)), ctx)
    return out, est, ref


def reciprocity(ctx, bsdf, samples=1000000, rng="mt19937", seed=0):
    """testReciprocity / testAdjoint: -> (mean[3], max[3], (in[3], out[3]))"""
    mean, mx, pair = np.zeros(3), np.zeros(3, np.float32), np.zeros(6, np.float32)
    _check(lib().bbmcu_check_reciprocity(ctx._h, bsdf._h, _U64(samples), _I(RNG[rng]), _U64(seed), _dp(mean), _dp(mx), _dp(pair)), ctx)
    return mean, mx, (pair[:3], pair[3:])


def pdf(ctx, bsdf, samples=100000, sample_sphere=False, check_below_horizon=False, rng="mt19937", seed=0, max_offenders=0):
    """testPdf: -> dict(negative=(r, i), below=(r, i), mismatch=(r, i), offenders=[(kind, pdf, dir, view), ...])"""
    counts, mis = np.zeros(4, np.uint64), np.zeros(2)
    off, n_off = np.zeros((max(1, max_offenders), 8), np.float32), _I(0)
    _check(lib().bbmcu_check_pdf(ctx._h, bsdf._h, _U64(samples), _I(int(sample_sphere)), _I(int(check_below_horizon)), _I(RNG[rng]), _U64(seed),
                                 _dp(counts), _dp(mis), _dp(off) if max_offenders else None, _I(max_offenders), C.byref(n_off)), ctx)
    return {"negative": (int(counts[0]), int(counts[1])), "below": (int(counts[2]), int(counts[3])), "mismatch": (mis[0], mis[1]),
            "offenders": [(int(o[0]), float(o[1]), o[2:5].copy(), o[5:8].copy()) for o in off[:n_off.value]]}


def pdf_integral(ctx, bsdf, samples=1000000, trials=10, sample_sphere=False, rng="mt19937", seed=0):
    """testPdfInt: -> (integral[trials], dirs[trials,3])"""
    val, dirs = np.zeros(trials), np.zeros((trials, 3), np.float32)
    _check(lib().bbmcu_check_pdf_integral(ctx._h, bsdf._h, _U64(samples), _I(trials), _I(int(sample_sphere)), _I(RNG[rng]), _U64(seed), _dp(val), _dp(dirs)), ctx)
    return val, dirs


def sample(ctx, bsdf, pdf_samples=4096, samples=100000, theta=10, phi=20, trials=10, sample_sphere=False, include_zero_pdf_samples=False,
           rng="mt19937", seed=0, bins=False):
    """testSample (chi-square of the sampled directions against the binned pdf): -> dict(chi2, df, P, dirs[, bin_pdf, bin_count])"""
    chi2, df, P, dirs = np.zeros(trials), np.zeros(trials), np.zeros(trials), np.zeros((trials, 3), np.float32)
    bp = np.zeros((trials, theta, phi)) if bins else None
    bc = np.zeros((trials, theta, phi), np.uint64) if bins else None
    _check(lib().bbmcu_check_sample(ctx._h, bsdf._h, _U64(pdf_samples), _U64(samples), _I(theta), _I(phi), _I(trials), _I(int(sample_sphere)),
                                    _I(int(include_zero_pdf_samples)), _I(RNG[rng]), _U64(seed), _dp(chi2), _dp(df), _dp(P), _dp(dirs),
                                    _dp(bp) if bins else None, _dp(bc) if bins else None), ctx)
    r = {"chi2": chi2, "df": df, "P": P, "dirs": dirs}
    if bins:
        r["bin_pdf"], r["bin_count"] = bp, bc
    return r


# ---- the command-line tool ------------------------------------------------------------------------------------------
def _g(v):
    """operator<< of a float (6 significant digits)"""
    return "%g" % float(np.float32(v))


def _vec(v):
    return "[" + ", ".join(_g(x) for x in v) + "]"


def _parse(argv):
    """util/option.h:22-35: key=value, or a bare word = true"""
    opt = {}
    for a in argv:
        k, eq, v = a.partition("=")
        if eq:
            opt[k.strip()] = v.strip()
        else:
            opt[a.strip()] = "true"
    return opt


def _bool(s):
    return str(s).strip().lower() in ("true", "1", "yes")


USAGE = """Usage: python -m bbm_b200.check [bsdfmodel=<bsdf string>] [test=<test name> [test options] [rng=mt19937|philox] [seed=N] [device=N]
  + test=reflectance [samples=100000] [theta=1] [importanceSampling]: compare the approximated reflectance method with a MC integration of the BSDF.
  + test=reciprocity [samples=100000]: checks if the BSDF is symmetric for 'samples' random dirctions.
  + test=adjoint [samples=100000]: checks if the adjoint BSDF is equal to the BSDF with in/out swapped.
  + test=pdf [samples=100000] [maxError=10] [checkBelowHorizon] [sampleSphere]: checks if the PDF >= 0, and the PDF returned by the sampling method matches the pdf from the pdf-method.
  + test=pdfInt [samples=100000] [trials=10] [sampleSphere]: checks the integral (MC with 'samples' samples) of the PDF for 'trials' different directions.
  + test=sample [pdfSamples=4069] [samples=100000] [theta=10] [phi=20] [trials=10] [sampleSphere] [includeZeroPdfSamples]: perform Chi2 test on the sample vs the pdf method.
  + info: list the models of this backbone (bbm_info)."""

_KEYS = {"reflectance": {"samples", "theta", "importanceSampling"}, "reciprocity": {"samples"}, "adjoint": {"samples"},
         "pdf": {"samples", "maxError", "checkBelowHorizon", "sampleSphere"}, "pdfInt": {"samples", "trials", "sampleSphere"},
         "sample": {"pdfSamples", "samples", "theta", "phi", "trials", "sampleSphere", "includeZeroPdfSamples"}}
_COMMON = {"bsdfmodel", "test", "rng", "seed", "device"}


def info(out=sys.stderr):
    """bin/bbm_info.cpp"""
    names = model_names()
    analytic = [n for n in names if n != "Merl"]
    print("BBM_NAME = 'bbm_b200' using 'cuda' backbone and compiled with python support.", file=out)
    print(f"{len(analytic)} BSDF models supported:", file=out)
    for n in analytic:
        print(" + " + n, file=out)
    static = [n for n in names if n == "Merl"]
    print(f"{len(static)} Static BSDF models supported:", file=out)
    for n in static:
        print(" + " + n, file=out)


def main(argv=None, out=sys.stdout):
    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        print(USAGE, file=out)
        return -1
    opt = _parse(argv)
    if "info" in opt:
        info()
        return 0
    if "bsdfmodel" not in opt:
        raise BbmError("Missing required option bsdfmodel")
    if "test" not in opt:
        raise BbmError("Missing required option test")
    test = opt["test"]
    if test == "":
        print("ERROR: no test specified.", file=out)
        return -1
    if test not in _KEYS:
        print(f"Unrecognized test: '{test}'", file=out)
        return 0
    invalid = sorted(k for k in opt if k not in _KEYS[test] | _COMMON)
    if invalid:
        print("ERROR: invalid keywords: (" + ", ".join('"%s"' % k for k in invalid) + ").", file=out)
        return 0
    rng, seed = opt.get("rng", "mt19937"), int(opt.get("seed", 0))
    ctx = Context(int(opt.get("device", 0)))
    bsdf = Bsdf(opt["bsdfmodel"])
    geti = lambda k, d: int(opt.get(k, d))           # noqa: E731
    if test == "reflectance":
        samples, theta, imp = geti("samples", 100000), geti("theta", 1), _bool(opt.get("importanceSampling", "false"))
        print(f"Reflectance test with {theta} directions and {samples} samples.", file=out)
        o, est, ref = reflectance(ctx, bsdf, samples, theta, imp, rng, seed)
        for t in range(theta):
            print(f" out = {_vec(o[t])} => Estimate: {_vec(est[t])} vs. {_vec(ref[t])}", file=out)
    elif test in ("reciprocity", "adjoint"):
        samples = geti("samples", 1000000 if test == "reciprocity" else 100000)
        mean, mx, (a, b) = reciprocity(ctx, bsdf, samples, rng, seed)
        pair = f"({_vec(a)}, {_vec(b)})"
        if test == "reciprocity":
            print(f"Reciprocity test with {samples} samples.", file=out)
            print(f"Radiance   average = {_vec(mean)}, max = {_vec(mx)} at {pair}", file=out)
            print(f"Importance average = {_vec(mean)}, max = {_vec(mx)} at {pair}", file=out)
        else:
            print(f"Adjoint test with {samples} samples.", file=out)
            print(f"Adjoint difference average = {_vec(mean)}, max = {_vec(mx)} at {pair}", file=out)
    elif test == "pdf":
        samples, max_err = geti("samples", 100000), geti("maxError", 10)
        below = _bool(opt.get("checkBelowHorizon", "false"))
        print(f"Tesing PDF properties test with {samples} samples.", file=out)
        r = pdf(ctx, bsdf, samples, _bool(opt.get("sampleSphere", "false")), below, rng, seed, max_offenders=4 * max_err)
        shown = [0, 0, 0, 0]
        for kind, p, d, v in r["offenders"]:        # the reference prints while it loops and stops at maxError of one kind
            if shown[kind] >= max_err:
                continue
            shown[kind] += 1
            if kind < 2:
                print(f" Sampled direction {_vec(d)} below horizon for {_vec(v)}", file=out)
            else:
                print(f" Negative PDF ({_g(p)}) for ({_vec(d)}, {_vec(v)})", file=out)
        line = f"PDF has {r['negative'][0]}/{r['negative'][1]} negative PDF values, "
        if below:
            line += f"{r['below'][0]}/{r['below'][1]} sampled directions below the horizon, "
        line += f"and {_g(r['mismatch'][0])}/{_g(r['mismatch'][1])} average difference between the PDF from the sample method and the corresponding PDF from the pdf-method."
        print(line, file=out)
    elif test == "pdfInt":
        samples, trials, sph = geti("samples", 1000000), geti("trials", 10), _bool(opt.get("sampleSphere", "false"))
        print(f"Tesing PDF Integral with {samples} samples, for {trials} random directions sampled over the {'sphere' if sph else 'hemisphere'}", file=out)
        val, dirs = pdf_integral(ctx, bsdf, samples, trials, sph, rng, seed)
        for t in range(trials):
            print(f" Integral = {_g(val[t])}/{_g(val[t])} (radiance/importance) for {_vec(dirs[t])}", file=out)
    else:
        ps, samples, th, ph, trials = geti("pdfSamples", 4096), geti("samples", 100000), geti("theta", 10), geti("phi", 20), geti("trials", 10)
        zero = _bool(opt.get("includeZeroPdfSamples", "false"))
        print(f"Testing if sample and pdf match: {ps} PDF samples per bin, and {samples} direction samples, with ({ph} x {th}) bins over {trials} trials"
              + (", including zero pdf samples" if zero else "") + ".", file=out)
        r = sample(ctx, bsdf, ps, samples, th, ph, trials, _bool(opt.get("sampleSphere", "false")), zero, rng, seed)
        for t in range(trials):
            print(f" Chi2 for {_vec(r['dirs'][t])} = {_g(r['chi2'][t])} (with {_g(r['df'][t])} degrees of freedom).", file=out)
            if r["df"][t] > 1:
                print(f"  P = {_g(r['P'][t]) if not math.isnan(r['P'][t]) else 'nan'} (reject if lower than confidence).", file=out)
            else:
                print(" No degrees of freedom; need at least 1 to compute P.", file=out)
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
