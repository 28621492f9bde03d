"""CPU: the compact pair loss (bbm_b200/csrc/bbmcu_losscompact.cuh, host-compiled) against the generic per-sample
loss + dual-number gradient of the same headers, for every model that has a compact kernel and all six metrics -
MERL-grid directions (below-horizon bins included) and measured values from another parameter set."""
import numpy as np
import pytest

METRICS = ["nganL2", "lowL2", "bieronL2", "lowLog", "bieronLog", "standardLog"]     # METRIC_* order of bbmcu_lossop.cuh

CASES = [
    ("Aggregate(Lambertian([0.3, 0.2, 0.1]), CookTorrance([0.4, 0.5, 0.6], 0.2, 1.6))", "Aggregate(Lambertian([0.25, 0.22, 0.12]), CookTorrance([0.5, 0.4, 0.7], 0.15, 1.4))"),
    ("Aggregate(Lambertian([0.1, 0.1, 0.3]), CookTorrance([1.4, 1.5, 0.9], 0.03, 1.3))", "Aggregate(Lambertian([0.12, 0.1, 0.25]), CookTorrance([1.0, 1.3, 1.1], 0.04, 1.25))"),
    ("Aggregate(Lambertian([0.3, 0.2, 0.1]), LowCookTorrance([0.4, 0.5, 0.6], 0.3, 1.8))", "Aggregate(Lambertian([0.2, 0.2, 0.2]), LowCookTorrance([0.3, 0.6, 0.5], 0.25, 1.5))"),
    ("Aggregate(Lambertian([0.3, 0.2, 0.1]), NganCookTorrance([0.4, 0.5, 0.6], 0.1, 0.2))", "Aggregate(Lambertian([0.2, 0.2, 0.2]), NganCookTorrance([0.3, 0.6, 0.5], 0.2, 0.1))"),
    ("Aggregate(Lambertian([0.3, 0.2, 0.1]), GGX([0.4, 0.5, 0.6], 0.2, 1.6))", "Aggregate(Lambertian([0.25, 0.22, 0.12]), GGX([0.5, 0.4, 0.7], 0.15, 1.4))"),
    ("Aggregate(Lambertian([0.1, 0.1, 0.3]), GGX([1.4, 1.5, 0.9], 0.02, 1.3))", "Aggregate(Lambertian([0.12, 0.1, 0.25]), GGX([1.0, 1.3, 1.1], 0.03, 1.25))"),
    ("Aggregate(Lambertian([0.3, 0.2, 0.1]), LowMicrofacet([0.4, 0.5, 0.6], 300.0, 1.2, 1.6))", "Aggregate(Lambertian([0.25, 0.22, 0.12]), LowMicrofacet([0.5, 0.4, 0.7], 200.0, 1.5, 1.4))"),
    ("Aggregate(Lambertian([0.1, 0.1, 0.3]), LowMicrofacetFit([51.7, 37.9, 27.4], 10482.1, 0.8167, 2.2365))", "Aggregate(Lambertian([0.12, 0.1, 0.25]), LowMicrofacetFit([40.0, 30.0, 30.0], 20000.0, 0.9, 1.8))"),
    ("Aggregate(Lambertian([0.3, 0.2, 0.1]), NganAshikhminShirley([0.4, 0.5, 0.6], 0.1, 80.0))", "Aggregate(Lambertian([0.25, 0.22, 0.12]), NganAshikhminShirley([0.5, 0.4, 0.7], 0.2, 120.0))"),
    ("Aggregate(Lambertian([0.3, 0.2, 0.1]), LowAshikhminShirley([0.4, 0.5, 0.6], 1.6, 2000.0))", "Aggregate(Lambertian([0.25, 0.22, 0.12]), LowAshikhminShirley([0.5, 0.4, 0.7], 1.4, 1500.0))"),
    ("Aggregate(Lambertian([0.3, 0.2, 0.1]), NganBlinnPhong([0.4, 0.5, 0.6], 60.0))", "Aggregate(Lambertian([0.25, 0.22, 0.12]), Phong([0.5, 0.4, 0.7], 90.0))"),
    ("Aggregate(Lambertian([0.3, 0.2, 0.1]), Phong([0.4, 0.5, 0.6], 900.0))", "Aggregate(Lambertian([0.25, 0.22, 0.12]), Phong([0.5, 0.4, 0.7], 700.0))"),
    ("Aggregate(Lambertian([0.3, 0.2, 0.1]), NganLafortune([0.4, 0.5, 0.6], -0.58, 0.57, 40.0))", "Aggregate(Lambertian([0.25, 0.22, 0.12]), NganLafortune([0.5, 0.4, 0.7], -0.6, 0.55, 60.0))"),
    ("Aggregate(Lambertian([0.3, 0.2, 0.1]), NganLafortune([0.4, 0.5, 0.6], -0.55, 0.6, 150.0))", "Aggregate(Lambertian([0.25, 0.22, 0.12]), NganLafortune([0.5, 0.4, 0.7], -0.6, 0.55, 100.0))"),
    ("Aggregate(Lambertian([0.3, 0.2, 0.1]), LowSmooth([40.0, 50.0, 60.0], 3000.0, 1.2, 1.6))", "Aggregate(Lambertian([0.25, 0.22, 0.12]), LowSmooth([50.0, 40.0, 70.0], 2000.0, 1.5, 1.4))"),
    # single lobes (no Lambertian lobe in front)
    ("CookTorrance([0.4, 0.5, 0.6], 0.2, 1.6)", "Aggregate(Lambertian([0.25, 0.22, 0.12]), CookTorrance([0.5, 0.4, 0.7], 0.15, 1.4))"),
    ("GGX([0.4, 0.5, 0.6], 0.2, 1.6)", "GGX([0.5, 0.4, 0.7], 0.15, 1.4)"),
    ("NganBlinnPhong([0.4, 0.5, 0.6], 60.0)", "Aggregate(Lambertian([0.25, 0.22, 0.12]), Phong([0.5, 0.4, 0.7], 90.0))"),
    # Ward lobes are Inf / NaN at the horizon in the reference: their samples stay off it (NO_HORIZON below)
    ("Aggregate(Lambertian([0.3, 0.2, 0.1]), NganWard([0.4, 0.5, 0.6], 0.2))", "Aggregate(Lambertian([0.25, 0.22, 0.12]), NganWard([0.5, 0.4, 0.7], 0.15))"),
    ("Aggregate(Lambertian([0.3, 0.2, 0.1]), NganWardDuer([0.4, 0.5, 0.6], 0.05))", "Aggregate(Lambertian([0.25, 0.22, 0.12]), NganWardDuer([0.5, 0.4, 0.7], 0.07))"),
    # total internal reflection inside the Fresnel term (eta < 1): g clamps to zero, F = 1
    ("Aggregate(Lambertian([0.3, 0.2, 0.1]), CookTorrance([0.4, 0.5, 0.6], 0.2, 0.9))", "Aggregate(Lambertian([0.25, 0.22, 0.12]), CookTorrance([0.5, 0.4, 0.7], 0.15, 1.4))"),
]


def _samples(hostsim, n, seed):
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, 1458000, n)
    i = np.empty((n, 3), np.float32); o = np.empty((n, 3), np.float32)
    for k, b in enumerate(idx):
        a, c = hostsim.merl_dirs(int(b), 1)
        i[k], o[k] = a[0], c[0]
    # the grid clamps to the horizon (z = 0: the diffuse lobe alone); random directions of the whole sphere go below it
    m = n // 4
    v = rng.normal(size=(2, m, 3)); v /= np.linalg.norm(v, axis=2, keepdims=True)
    i[:m], o[:m] = v[0].astype(np.float32), v[1].astype(np.float32)
    return i, o


@pytest.mark.parametrize("case", range(len(CASES)))
def test_compact_loss_equals_generic_on_host(hostsim, case):
    import bbm_b200 as bb
    fitted, truth = CASES[case]
    P = len(bb.Bsdf(fitted).parameter_values())
    i, o = _samples(hostsim, 6000, case)
    assert (i[:, 2] < 0).any() and (o[:, 2] < 0).any() and (i[:, 2] == 0).any()      # the below-horizon constant and the clamped bins are exercised
    if "Ward" in fitted:
        keep = (np.abs(i[:, 2]) > 1e-3) & (np.abs(o[:, 2]) > 1e-3)
        bad = hostsim.loss_compact(fitted, 0, i, o, hostsim.eval(truth, i, o), nparams=P)
        assert not np.isfinite(bad[0]) and not np.isfinite(hostsim.loss(fitted, 0, i, o, hostsim.eval(truth, i, o), nparams=P)[0])     # non-finite in both at the horizon
        i, o = i[keep], o[keep]
    ref = hostsim.eval(truth, i, o)
    for m, name in enumerate(METRICS):
        want, wg, _ = hostsim.loss(fitted, m, i, o, ref, nparams=P)
        got = hostsim.loss_compact(fitted, m, i, o, ref, nparams=P)
        assert got is not None, fitted
        assert abs(got[0] - want) <= 2e-5 * abs(want), (name, got[0], want)
        # eta < 1: d g / d eta = eta / g is singular where g = sqrt(eta^2 + c^2 - 1) -> 0, and the two kernels round g^2 differently
        rel = 2e-3 if "1.6, 0.9" in fitted or ", 0.9))" in fitted else 5e-5
        tol = rel * np.abs(wg) + 1e-6 * np.abs(wg).max()
        assert np.all(np.abs(got[1] - wg) <= tol), (name, got[1], wg)
        v = hostsim.loss_compact(fitted, m, i, o, ref, want_grad=False)
        assert abs(v[0] - want) <= 2e-5 * abs(want), name


def test_models_without_compact_kernel_report_so(hostsim):
    i, o = _samples(hostsim, 16, 1)
    ref = np.zeros((16, 3), np.float32)
    assert hostsim.loss_compact("Aggregate(Lambertian([0.3, 0.2, 0.1]), Ward([0.4, 0.5, 0.6], [0.2, 0.3]))", 0, i, o, ref) is None
    assert hostsim.loss_compact("CookTorranceWalter([0.4, 0.5, 0.6], 0.2, 1.6)", 0, i, o, ref) is None
