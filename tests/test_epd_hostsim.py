"""CPU: EPD ("Holzschuch-Pacanowski") device code compiled for the host against the compiled unmodified
reference (oracle/_ref) over the (beta, p) range of its G1 table: a = 1/p on both sides of 1 exercises every
branch of the DiDonato-Morris inverse incomplete gamma estimate (include/util/invgamma.h)."""
import numpy as np
import pytest

from tests.util import mismatch


def _dirs(rng, n):
    z = rng.random(n)
    ph = rng.random(n) * 2 * np.pi
    s = np.sqrt(1 - z * z)
    return np.stack([s * np.cos(ph), s * np.sin(ph), z], 1).astype(np.float32)


@pytest.mark.parametrize("beta", [0.003, 0.05, 0.5])
def test_epd_eval_sample_pdf_range(hostsim, ref, beta):
    rng = np.random.default_rng(3)
    n = 4000
    inn, out, xi = _dirs(rng, n), _dirs(rng, n), rng.random((n, 2)).astype(np.float32)
    xi[:8] = [[0, 0], [1, 1], [0, 1], [1, 0], [0.5, 1e-7], [0.5, 0.9999999], [0.25, 0.5], [0.75, 0.5]]
    for p in (0.06, 0.1, 0.2, 0.35, 0.5, 0.77, 1.0, 1.5, 2.0, 3.3, 5.0):
        s = f"EPD({beta}, {p}, [1.5, 0.7])"
        assert not mismatch(hostsim.eval(s, inn, out), ref.eval(s, inn, out)).any(), s
        assert not mismatch(hostsim.pdf(s, inn, out), ref.pdf(s, inn, out)).any(), s
        d, pp, f = hostsim.sample(s, out, xi)
        d2, pp2, f2 = ref.sample(s, out, xi)
        assert np.array_equal(f, f2), s
        assert not mismatch(d, d2, 1e-5, 1e-5).any(), s
        assert not mismatch(pp, ref.pdf(s, d, out)).any(), s
