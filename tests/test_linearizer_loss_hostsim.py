"""CPU: linearizers (bit-exact), the glibc libm restatements, and the loss / gradient arithmetic of the
device headers (host-compiled) against the golden vectors of the unmodified reference."""
import numpy as np

from tests.util import assert_parity


def test_glibc_trig_ports_match_host_libm(hostsim):
    """the restated atan2f / sinf / cosf reproduce this host's glibc bit for bit"""
    rng = np.random.default_rng(0)
    n = 2_000_000
    y = (rng.random(n, dtype=np.float32)*2 - 1); x = (rng.random(n, dtype=np.float32)*2 - 1)
    y[::3] *= np.float32(1e-3); x[::7] *= np.float32(1e-4)
    assert hostsim.libm_mismatches(0, y, x) == 0
    a = (rng.random(n, dtype=np.float32)*2 - 1) * np.float32(7.0)
    a[::5] *= np.float32(1e-2)
    assert hostsim.libm_mismatches(1, a) == 0
    assert hostsim.libm_mismatches(2, a) == 0
    assert hostsim.libm_mismatches(3, a) == 0          # glibc_sincosf_both (shared reduction, signs applied last)
    assert hostsim.libm_mismatches(4, a) == 0
    e = (rng.random(n, dtype=np.float32) * 2 - 1) * np.float32(100.0)
    e[::2] *= rng.random(n // 2, dtype=np.float32)
    assert hostsim.libm_mismatches(5, e) == 0                          # glibc_expf (the host's FMA multiarch variant)
    l = rng.random(n, dtype=np.float32) * np.float32(10.0)
    l[::3] = np.float32(1.0) + (rng.random(len(l[::3]), dtype=np.float32) - np.float32(0.5)) * np.float32(1e-3)
    assert hostsim.libm_mismatches(6, l) == 0                          # glibc_logf
    assert hostsim.libm_mismatches(7, a) == 0                          # glibc_erff (a spans [-7, 7])
    z = np.float32(2 * np.pi) * rng.random(n, dtype=np.float32)       # the sampler's phi range
    assert hostsim.libm_mismatches(3, z) == 0 and hostsim.libm_mismatches(4, z) == 0
    t = (rng.random(n, dtype=np.float32) * 2 - 1) * np.float32(np.pi)  # LowSmooth: tan(pi xi)
    t[::4] = np.float32(np.pi / 2) + (rng.random(len(t[::4]), dtype=np.float32) - np.float32(0.5)) * np.float32(1e-3)
    assert hostsim.libm_mismatches(8, t) == 0                          # glibc_tanf (every |x| <= 100 swept once: tools/libm_sweep.cpp)
    c = (rng.random(n, dtype=np.float32) * 2 - 1) * np.float32(12.0)
    c[::3] *= rng.random(len(c[::3]), dtype=np.float32)
    assert hostsim.libm_mismatches(9, c) == 0                          # glibc_erfcf (all 2^32 floats swept once)
    assert hostsim.libm_mismatches(10, (e * np.float32(0.3))) == 0     # glibc_atanf


def test_fused_linearizer_tables_equal_direct_evaluation(hostsim):
    """merl_dirs_tab (what the eval-grid and loss kernels use) returns the bits of merl_dirs for every bin"""
    hostsim.lib.hostsim_merl_dirs_tab_mismatches.restype = __import__("ctypes").c_size_t
    assert hostsim.lib.hostsim_merl_dirs_tab_mismatches(__import__("ctypes").c_uint32(0), __import__("ctypes").c_size_t(1458000)) == 0


def test_merl_index_bit_exact(hostsim, golden_lin):
    g = golden_lin
    assert np.array_equal(hostsim.merl_index(g["pairs_in"], g["pairs_out"]), g["pairs_index"])
    assert np.array_equal(hostsim.merl_index(g["grid_in"], g["grid_out"]), g["grid_index_of_dirs"])


def test_merl_dirs_bit_exact(hostsim, golden_lin):
    g = golden_lin
    for k, idx in enumerate(g["grid_idx"][:512]):
        i, o = hostsim.merl_dirs(int(idx), 1)
        assert np.array_equal(i[0].view(np.uint32), g["grid_in"][k].view(np.uint32)), idx
        assert np.array_equal(o[0].view(np.uint32), g["grid_out"][k].view(np.uint32)), idx


def test_spherical_dirs_bit_exact(hostsim, golden_lin):
    g = golden_lin
    hp, tp = float(np.float32(2)*np.float32(np.pi)), float(np.float32(0.5)*np.float32(np.pi))
    i, o = hostsim.spherical_dirs([12, 7, 5, 6], [0, 0, hp, tp, 0, 0, hp, tp], 0, 12*7*5*6)
    assert np.array_equal(i.view(np.uint32), g["sph_in"].view(np.uint32)) and np.array_equal(o.view(np.uint32), g["sph_out"].view(np.uint32))
    i, o = hostsim.spherical_dirs([9, 4, 1, 5], [0.1, 0.05, 3.0, 1.4, 0.0, 0.2, 6.0, 1.5], 0, 9*4*5)
    assert np.array_equal(i.view(np.uint32), g["sph2_in"].view(np.uint32)) and np.array_equal(o.view(np.uint32), g["sph2_out"].view(np.uint32))


def test_reference_against_golden(ref, golden_lin, golden_loss):
    """pins the oracle build: the compiled reference reproduces the committed vectors"""
    g = golden_lin
    assert np.array_equal(ref.merl_index(g["pairs_in"], g["pairs_out"]).astype(np.uint32), g["pairs_index"])
    arr, meta = golden_loss
    from oracle.refbind import sph_desc
    t, _ = ref.loss("nganL2", sph_desc((13, 8), (5, 6)), meta["fitted"], meta["truth"], 0, meta["metrics"]["nganL2"]["N"])
    assert np.array_equal(t, arr["nganL2_terms"])


def test_loss_terms_total_and_gradient(hostsim, golden_loss):
    import bbm_b200 as bb
    arr, meta = golden_loss
    hp, tp = float(np.float32(2)*np.float32(np.pi)), float(np.float32(0.5)*np.float32(np.pi))
    P = len(bb.Bsdf(meta["fitted"]).parameter_values())
    for m, name in enumerate(bb.METRICS):
        rec = meta["metrics"][name]
        so = rec["samples_out"]
        i, o = hostsim.spherical_dirs([13, 8, so[0], so[1]], [0, 0, hp, tp, 0, 0, hp, tp], 0, rec["N"])
        truth = hostsim.eval(meta["truth"], i, o)
        loss, grad, terms = hostsim.loss(meta["fitted"], m, i, o, truth, nparams=P)
        assert_parity(terms, arr[name + "_terms"], 1e-5, floor=1e-12, what=f"{name} terms")
        assert abs(loss - rec["double_sum_of_float_terms"]) <= 1e-4*abs(rec["double_sum_of_float_terms"])
        assert abs(loss - rec["double_total"]) <= 1e-4*abs(rec["double_total"])
        fd = np.array(rec["fd_gradient"])
        assert np.all(np.abs(grad - fd) <= 1e-4*np.abs(fd) + 1e-9), (name, grad, fd)
