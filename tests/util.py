"""Parity comparison used by every test (SURVEY.md section 8c "Parity definitions")."""
import numpy as np

# models whose CUDA kernels exist in this build; the golden fixtures cover all 34
IMPLEMENTED = None


def implemented_models():
    global IMPLEMENTED
    if IMPLEMENTED is None:
        import bbm_b200 as bb
        IMPLEMENTED = bb.model_names()          # all 34 analytic models + Merl
    return IMPLEMENTED


def uses_only_implemented(bsdf_string):
    import re
    names = re.findall(r"([A-Za-z]+)\(", bsdf_string)
    ok = set(implemented_models()) | {"Aggregate"}
    return all(n in ok for n in names)


def mismatch(got, want, rel=1e-5, floor=1e-30):
    """boolean array of elements violating: NaN matches NaN, +-Inf matches the same Inf, otherwise
    |got - want| <= rel * |want| where |want| > floor, else |got| <= floor-scale absolute."""
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    nan_ok = np.isnan(got) & np.isnan(want)
    inf_ok = np.isinf(got) & np.isinf(want) & (np.sign(got) == np.sign(want))
    with np.errstate(invalid="ignore"):
        close = np.abs(got - want) <= rel * np.abs(want) + floor
    return ~(nan_ok | inf_ok | (close & np.isfinite(got) & np.isfinite(want)))


def assert_parity(got, want, rel=1e-5, floor=1e-30, what="", max_bad=0):
    bad = mismatch(got, want, rel, floor)
    nbad = int(bad.sum())
    if nbad > max_bad:
        idx = np.argwhere(bad)[:5]
        g, w = np.asarray(got), np.asarray(want)
        lines = [f"  at {tuple(i)}: got {g[tuple(i)]!r} want {w[tuple(i)]!r}" for i in idx]
        raise AssertionError(f"{what}: {nbad} of {bad.size} elements differ beyond rel {rel}\n" + "\n".join(lines))


def soa(a):
    return np.ascontiguousarray(np.asarray(a, np.float32).T)


def pdf_floor(bsdf_string, want):
    """absolute floor for pdf comparisons.  The data-driven sampler of the He family (ndf/sampler.h:102-128) returns
    DIFFERENCES of neighbouring entries of a float CDF: in the tail (cdf ~ 1, bin mass < 1e-6) that difference
    cancels catastrophically in the reference itself, so a last-bit change of one CDF sample (device expf / erfcf
    versus glibc) moves those pdf values by far more than 1e-5 relative.  They are compared to 1e-5 of the peak pdf
    scaled by 1e-6 instead; every other model keeps the plain 1e-30 floor."""
    if "He" not in bsdf_string:
        return 1e-30
    w = np.asarray(want, np.float64)
    w = w[np.isfinite(w)]
    return 1e-6 * float(np.abs(w).max()) if w.size else 1e-30
