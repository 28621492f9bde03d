"""CPU: the analytic parameter gradient (new capability; the reference has none) of every
model, through the host-compiled device code, against central finite differences of the reference loss compiled
in doubleRGB (oracle/_ref) - the gradient oracle of SURVEY.md section 8(c).  Tolerance 1e-4 relative with an
absolute floor of 1e-4 of the largest gradient component (parameters the loss does not depend on)."""
import numpy as np
import pytest

from oracle.refbind import sph_desc

# (fitted, truth): single models, so the reference's parameter order is the forward order (fact 14 only
# affects run-time aggregates)
CASES = [
    ("Lambertian([0.4, 0.3, 0.2])", "Lambertian([0.5, 0.2, 0.1])"),
    ("OrenNayar([0.4, 0.3, 0.2], 0.3)", "OrenNayar([0.5, 0.2, 0.1], 0.5)"),
    ("Phong([0.4, 0.3, 0.2], 12)", "Phong([0.5, 0.2, 0.1], 20)"),
    ("Lafortune([0.4, 0.3, 0.2], [-0.6, -0.5], 0.6, 9)", "Lafortune([0.5, 0.2, 0.1], [-0.58, -0.58], 0.57, 12)"),
    ("NganLafortune([0.4, 0.3, 0.2], -0.62, 0.6, 9)", "NganLafortune([0.5, 0.2, 0.1], -0.58, 0.57, 12)"),
    ("Ward([0.4, 0.3, 0.2], [0.2, 0.3])", "Ward([0.5, 0.2, 0.1], [0.25, 0.25])"),
    ("WardDuer([0.4, 0.3, 0.2], [0.2, 0.3])", "WardDuer([0.5, 0.2, 0.1], [0.25, 0.25])"),
    ("WardDuerGeislerMoroder([0.4, 0.3, 0.2], [0.2, 0.3])", "WardDuerGeislerMoroder([0.5, 0.2, 0.1], [0.25, 0.25])"),
    ("NganWard([0.4, 0.3, 0.2], 0.2)", "NganWard([0.5, 0.2, 0.1], 0.25)"),
    ("NganWardDuer([0.4, 0.3, 0.2], 0.2)", "NganWardDuer([0.5, 0.2, 0.1], 0.25)"),
    ("AshikhminShirley([0.2, 0.3, 0.4], [20, 30])", "AshikhminShirley([0.3, 0.3, 0.3], [25, 25])"),
    ("AshikhminShirleyFull([0.3, 0.2, 0.1], [0.2, 0.3, 0.4], [20, 30])", "AshikhminShirleyFull([0.2, 0.2, 0.2], [0.3, 0.3, 0.3], [25, 25])"),
    ("NganAshikhminShirley([0.4, 0.3, 0.2], 0.2, 20)", "NganAshikhminShirley([0.5, 0.2, 0.1], 0.3, 25)"),
    ("LowAshikhminShirley([0.4, 0.3, 0.2], 1.4, 20)", "LowAshikhminShirley([0.5, 0.2, 0.1], 1.6, 25)"),
    ("CookTorrance([0.4, 0.3, 0.2], 0.25, 1.4)", "CookTorrance([0.5, 0.2, 0.1], 0.2, 1.6)"),
    ("NganCookTorrance([0.4, 0.3, 0.2], 0.25, 0.2)", "NganCookTorrance([0.5, 0.2, 0.1], 0.2, 0.3)"),
    ("CookTorranceWalter([0.4, 0.3, 0.2], 0.25, 1.4)", "CookTorranceWalter([0.5, 0.2, 0.1], 0.2, 1.6)"),
    ("CookTorranceHeitz([0.4, 0.3, 0.2], [0.25, 0.3], 1.4)", "CookTorranceHeitz([0.5, 0.2, 0.1], [0.2, 0.2], 1.6)"),
    ("GGX([0.4, 0.3, 0.2], 0.25, 1.4)", "GGX([0.5, 0.2, 0.1], 0.2, 1.6)"),
    ("GGXHeitz([0.4, 0.3, 0.2], [0.25, 0.3], 1.4)", "GGXHeitz([0.5, 0.2, 0.1], [0.2, 0.2], 1.6)"),
    ("PhongWalter([0.4, 0.3, 0.2], 20, 1.4)", "PhongWalter([0.5, 0.2, 0.1], 30, 1.6)"),
    ("LowMicrofacet([0.4, 0.3, 0.2], 30, 1.5, 1.4)", "LowMicrofacet([0.5, 0.2, 0.1], 40, 1.3, 1.6)"),
    ("LowSmooth([0.4, 0.3, 0.2], 30, 1.5, 1.4)", "LowSmooth([0.5, 0.2, 0.1], 40, 1.3, 1.6)"),
    ("Ribardiere([0.4, 0.3, 0.2], 0.25, 2.5, 1.4)", "Ribardiere([0.5, 0.2, 0.1], 0.2, 2.0, 1.6)"),
    ("RibardiereAnisotropic([0.4, 0.3, 0.2], [0.25, 0.3], 2.5, 1.4)", "RibardiereAnisotropic([0.5, 0.2, 0.1], [0.2, 0.2], 2.0, 1.6)"),
    ("Bagher([0.4, 0.3, 0.2], [7.5, 7.5, 7.5], [1, 1, 1], [1, 1, 1], [1.3, 1.3, 1.3], [1, 1, 1], [0.2, 0.25, 0.3], [0.6, 0.7, 0.8], [[0.9, 0.8, 0.7], [0.1, 0.2, 0.05]])",
     "Bagher([0.5, 0.2, 0.1], [7.5, 7.5, 7.5], [1, 1, 1], [1, 1, 1], [1.3, 1.3, 1.3], [1, 1, 1], [0.15, 0.15, 0.15], [0.64, 0.64, 0.64], [[1, 1, 1], [0, 0, 0]])"),
    ("EPD(0.3, 0.8, [1.5, 0.7])", "EPD(0.25, 1.0, [1.3, 0.2])"),
    ("He(0.25, 2.5, [[0.3, 1.0, 1.5], [2.5, 2.0, 1.5]])", "He(0.18, 3.0, [[0.2, 0.9, 1.4], [3.0, 2.4, 1.9]])"),
    ("HeWestin(0.25, 2.5, [[0.3, 1.0, 1.5], [2.5, 2.0, 1.5]])", "HeWestin(0.18, 3.0, [[0.2, 0.9, 1.4], [3.0, 2.4, 1.9]])"),
    ("HeHolzschuch(0.1, 2.5, [[0.3, 1.0, 1.5], [2.5, 2.0, 1.5]])", "HeHolzschuch(0.08, 3.0, [[0.2, 0.9, 1.4], [3.0, 2.4, 1.9]])"),
    ("NganHe([0.4, 0.3, 0.2], 0.25, 2.5, 1.4)", "NganHe([0.5, 0.2, 0.1], 0.18, 3.0, 1.6)"),
]


GRID = ((11, 6), (4, 5))
T0, TP = 0.05, 1.4            # stay off the zenith (He's geometrical factor is 0/0 at normal incidence) and off the horizon (Ward-type models are NaN/Inf at z == 0, fact 7)


def check_gradient_case(ref, refd, metric, fitted, truth, loss, grad):
    """loss / grad of `fitted` against `truth` over GRID (computed by the caller: host-compiled or device kernels):
    value against the floatRGB reference's per-sample terms accumulated in double, gradient against
    central differences of the doubleRGB reference loss.  Tolerance per component: 1e-4 relative plus 3e-6 of the largest
    component (central differences of a double loss with h = 1e-6 carry ~1e-10 / 1e-6 of absolute noise)."""
    import bbm_b200 as bb
    hp = float(np.float32(2) * np.float32(np.pi))
    df = sph_desc(*GRID, start_in=(0, T0), start_out=(0, T0), end_in=(hp, TP), end_out=(hp, TP))
    dd = sph_desc(*GRID, real=np.float64, start_in=(0, T0), start_out=(0, T0), end_in=(2 * np.pi, TP), end_out=(2 * np.pi, TP))
    p0 = bb.Bsdf(fitted).parameter_values()
    P = len(p0)

    def central(r, desc, h_rel):
        rows = []
        for j in range(P):
            h = h_rel * max(1.0, abs(p0[j]))
            pp, pm = p0.copy(), p0.copy()
            pp[j] += h
            pm[j] -= h
            rows += [pp, pm]
        l = r.loss_at(metric, desc, fitted, truth, np.stack(rows))
        return np.array([(l[2 * j] - l[2 * j + 1]) / (2 * h_rel * max(1.0, abs(p0[j]))) for j in range(P)])
    lf = ref.loss_at(metric, df, fitted, truth, p0[None])[0]
    assert abs(loss - lf) <= 1e-5 * abs(lf), (fitted, loss, lf)
    l0 = refd.loss_at(metric, dd, fitted, truth, p0[None])[0]
    if abs(lf - l0) <= 2e-5 * abs(l0):
        fd, rel, floor = central(refd, dd, 1e-6), 1e-4, 3e-6
    else:
        # the He family is a DIFFERENT function in floatRGB and doubleRGB: its adaptive Taylor series stops on
        # hmin(term) < Constants::Epsilon() (he.h:459), i.e. FLT_EPSILON vs DBL_EPSILON, and the channel with
        # the smallest term truncates the others (0.6 % in blue here).  The float function is the parity target,
        # so its gradient is checked against differences of the floatRGB reference.  That function also JUMPS
        # whenever the number of series terms changes with g(roughness) - a wide finite difference averages over the
        # jumps, a derivative does not - so the oracle is the MEDIAN slope of 40 short segments per parameter.
        assert "He" in fitted
        fd = np.empty(P)
        for j in range(P):
            d = 2.5e-5 * max(1.0, abs(p0[j]))
            pts = np.tile(p0, (41, 1))
            pts[:, j] += np.arange(-20, 21) * d
            fd[j] = np.median(np.diff(ref.loss_at(metric, df, fitted, truth, pts)) / d)
        rel, floor = 5e-3, 5e-3
    tol = rel * np.abs(fd) + floor * np.abs(fd).max()
    assert np.all(np.abs(grad - fd) <= tol), (fitted, metric, rel, grad, fd, np.abs(grad - fd) / tol)


@pytest.mark.parametrize("metric", ["nganL2", "standardLog"])
def test_gradient_vs_finite_differences_of_double_reference(hostsim, ref, refd, metric):
    import bbm_b200 as bb
    hp = float(np.float32(2) * np.float32(np.pi))
    N = 11 * 6 * 4 * 5
    i, o = hostsim.spherical_dirs([11, 6, 4, 5], [0, T0, hp, TP, 0, T0, hp, TP], 0, N)
    m = bb.METRICS.index(metric)
    for fitted, truth in CASES:
        P = len(bb.Bsdf(fitted).parameter_values())
        tv = hostsim.eval(truth, i, o)
        loss, grad, _ = hostsim.loss(fitted, m, i, o, tv, nparams=P)
        check_gradient_case(ref, refd, metric, fitted, truth, loss, grad)
