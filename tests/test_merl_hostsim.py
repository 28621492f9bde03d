"""CPU: the measured MERL model Merl("file") - nearest-bin lookup, data-driven sampling, pdf - through the host-compiled
device code against the compiled unmodified reference reading the SAME synthetic MERL binary (written here in the
reference's file format, include/staticmodel/merl.h:173-206)."""
import numpy as np
import pytest

from tests.util import mismatch, pdf_floor

N = 90 * 90 * 180


@pytest.fixture(scope="module")
def merl_file(tmp_path_factory, hostsim):
    """a synthetic MERL-shaped measurement: Lambertian + GGX evaluated at the grid's own directions"""
    path = str(tmp_path_factory.mktemp("merl") / "synthetic.binary")
    i, o = hostsim.merl_dirs(0, N)
    rgb = hostsim.eval("Aggregate(Lambertian([0.2, 0.1, 0.05]), GGX([0.3, 0.3, 0.3], 0.25, 1.5))", i, o).astype(np.float64)   # (N, 3)
    scale = np.array([1.0, 1.15, 1.66])
    with open(path, "wb") as f:
        np.array([90, 90, 180], np.uint32).tofile(f)
        (rgb.T * 1500.0 / scale[:, None]).astype(np.float64).tofile(f)
    return path


def _dirs(rng, n):
    z = rng.random(n)
    ph = rng.random(n) * 2 * np.pi
    s = np.sqrt(1 - z * z)
    return np.stack([s * np.cos(ph), s * np.sin(ph), z], 1).astype(np.float32)


def test_merl_eval_sample_pdf(hostsim, ref, merl_file):
    import bbm_b200 as bb
    s = f'Merl("{merl_file}")'
    assert bb.Bsdf(s).to_string() == ref.to_string(s) == s
    assert len(bb.Bsdf(s).parameter_values()) == 0
    rng = np.random.default_rng(9)
    n = 20000
    inn, out, xi = _dirs(rng, n), _dirs(rng, n), rng.random((n, 2)).astype(np.float32)
    inn[0] = out[0] = [0, 0, 1]
    for c in (3, 2, 1):
        assert not mismatch(hostsim.eval(s, inn, out, c), ref.eval(s, inn, out, c), 0.0, 0.0).any()          # a table lookup: exact
        assert not mismatch(hostsim.reflectance(s, out, c), ref.reflectance(s, out, c), 0.0, 0.0).any()
        want_p = ref.pdf(s, inn, out, c)
        assert not mismatch(hostsim.pdf(s, inn, out, c), want_p, 1e-5, pdf_floor("He", want_p)).any()
        d, p, f = hostsim.sample(s, out, xi, c)
        d2, p2, f2 = ref.sample(s, out, xi, c)
        assert np.array_equal(f, f2)
        assert not mismatch(d, d2, 1e-5, 1e-5).any()
        want = ref.pdf(s, d, out, c)
        assert not mismatch(p, want, 1e-5, pdf_floor("He", want)).any()
    # in an aggregate, and by name
    agg = f'Aggregate(Lambertian([0.1, 0.1, 0.1]), Merl(filename = "{merl_file}"))'
    assert not mismatch(hostsim.eval(agg, inn, out), ref.eval(agg, inn, out)).any()


def test_merl_errors():
    import bbm_b200 as bb
    with pytest.raises(bb.BbmError):
        bb.Bsdf('Merl("/nonexistent/file.binary")')
