"""CPU: pins the plain-C oracle port (oracle/bbm_oracle.c) against the golden vectors produced by the unmodified
reference, and against the compiled reference itself when oracle/_ref is built."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.util import assert_parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def port():
    lib = os.path.join(ROOT, "oracle", "libbbm_oracle.so")
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(os.path.join(ROOT, "oracle", "bbm_oracle.c")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, capture_output=True)
    L = C.CDLL(lib)
    L.bbmo_ngan_l2_term.restype = C.c_float
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_port_merl_index_matches_golden(port, golden_lin):
    g = golden_lin
    for a, b, want in ((g["pairs_in"], g["pairs_out"], g["pairs_index"]), (g["grid_in"], g["grid_out"], g["grid_index_of_dirs"])):
        i, o = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
        idx = np.empty(len(i), np.uint32)
        port.bbmo_merl_index_n(C.c_size_t(len(i)), _p(i), _p(o), _p(idx))
        assert np.array_equal(idx, want)


def test_port_eval_matches_golden(port, golden_models):
    arr, meta = golden_models
    i, o = np.ascontiguousarray(arr["in"], np.float32), np.ascontiguousarray(arr["out"], np.float32)
    cases = {"Lambertian()": (0, [0.5, 0.5, 0.5]), "CookTorrance()": (1, [0.5, 0.5, 0.5, 0.1, 1.3]), "GGX()": (2, [0.5, 0.5, 0.5, 0.1, 1.3])}
    seen = 0
    for key, rec in meta["cases"].items():
        if rec["string"] in cases:
            model, a = cases[rec["string"]]
            a = np.array(a, np.float32)
            for c in (3, 1, 2):
                rgb = np.empty_like(i)
                port.bbmo_eval_n(model, _p(a), c, C.c_size_t(len(i)), _p(i), _p(o), _p(rgb))
                assert_parity(rgb, arr[f"{key}_eval_c{c}"], 1e-6, what=f"port eval {rec['string']} comp {c}")
            seen += 1
    assert seen == 3


def test_port_loss_term_matches_golden(port, golden_loss, hostsim):
    arr, meta = golden_loss
    hp, tp = float(np.float32(2) * np.float32(np.pi)), float(np.float32(0.5) * np.float32(np.pi))
    N = meta["metrics"]["nganL2"]["N"]
    i, o = hostsim.spherical_dirs([13, 8, 5, 6], [0, 0, hp, tp, 0, 0, hp, tp], 0, N)       # bit-exact with the reference (tested elsewhere)
    i, o = np.ascontiguousarray(i), np.ascontiguousarray(o)
    def ev(lam, ct):
        a0, a1 = np.array(lam, np.float32), np.array(ct, np.float32)
        r0, r1 = np.empty_like(i), np.empty_like(i)
        port.bbmo_eval_n(0, _p(a0), 3, C.c_size_t(N), _p(i), _p(o), _p(r0))
        port.bbmo_eval_n(1, _p(a1), 3, C.c_size_t(N), _p(i), _p(o), _p(r1))
        return (np.float32(0) + r0) + r1
    v = ev([0.5, 0.5, 0.5], [0.5, 0.5, 0.5, 0.1, 1.3])
    r = ev([0.2, 0.1, 0.05], [0.3, 0.3, 0.3, 0.2, 1.5])
    terms = np.array([port.bbmo_ngan_l2_term(_p(i[k]), _p(o[k]), _p(v[k]), _p(r[k])) for k in range(N)], np.float32)
    assert_parity(terms, arr["nganL2_terms"], 1e-5, floor=1e-12, what="port nganL2 terms")
