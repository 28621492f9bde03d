"""Holzschuch-Pacanowski renormalisation table (precompute/HolzschuchPacanowski/normalization.cpp):
the table is one of the reference's missing large blobs, so the oracle is the generator itself, compiled as it lies
(oracle/hp_driver.cpp -> oracle/_ref/libbbmref_hp.so) and asked for single entries.
CPU: the entry function of bbmcu_hpnorm.cuh compiled for the host gives the generator's bits (same libm).
GPU: the whole 100 x 100 x 100 table from one kernel against a sample of generator entries."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFLIB = os.path.join(ROOT, "oracle", "_ref", "libbbmref_hp.so")


def _ref():
    if not os.path.exists(REFLIB):
        pytest.skip("oracle/_ref/libbbmref_hp.so not built (make -C oracle ref)")
    L = C.CDLL(REFLIB)
    L.ref_hp_normalization_entry.restype = C.c_float
    L.ref_hp_normalization_entry.argtypes = [C.c_int] * 3
    return L


def _sample(n, seed):
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, 100, size=(n, 3))
    corners = [(0, 0, 0), (99, 99, 99), (0, 99, 0), (99, 0, 99), (0, 0, 99), (50, 50, 0), (99, 99, 1), (0, 0, 1)]
    return [tuple(int(v) for v in r) for r in idx] + corners


def test_entry_function_on_host_equals_the_generator(hostsim):
    L = _ref()
    f = hostsim.lib.hostsim_hp_normalization_entry
    f.restype = C.c_float
    f.argtypes = [C.c_int] * 3
    bad = []
    for b, c, s in _sample(400, 1):
        got, want = np.float32(f(b, c, s)), np.float32(L.ref_hp_normalization_entry(b, c, s))
        if got.view(np.uint32) != want.view(np.uint32):
            bad.append((b, c, s, float(got), float(want)))
    assert not bad, bad[:5]


@pytest.mark.gpu
def test_table_regenerated_on_gpu(ctx):
    import time
    L = _ref()
    t0 = time.perf_counter()
    table = ctx.hp_precompute_normalization()
    dt = time.perf_counter() - t0
    assert table.shape == (100, 100, 100) and np.isfinite(table).all()
    # device double pow / acos are not the host libm's: an entry is a float sum of up to 11 459 terms, each a double rounded once
    worst = 0.0
    for b, c, s in _sample(1500, 2):
        want = L.ref_hp_normalization_entry(b, c, s)
        err = abs(float(table[b, c, s]) - want)
        worst = max(worst, err / max(abs(want), 1e-6))
        assert err <= 2e-6 * abs(want) + 1e-7, (b, c, s, float(table[b, c, s]), want)
    # every entry is a pure function of its three indices: a second run returns the same bits
    assert np.array_equal(table.view(np.uint32), ctx.hp_precompute_normalization().view(np.uint32))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        import json
        json.dump({"seconds_whole_table_incl_copy": dt, "entries": 1000000, "quadrature_terms": float(sum(int(2 * (s / 100) / (0.01 * np.pi / 180)) + 1 for s in range(100)) * 10000),
                   "worst_relative_error_vs_generator_on_1508_entries": worst,
                   "table_min": float(table.min()), "table_max": float(table.max())}, open(os.path.join(out, "hp_normalization_r02.json"), "w"), indent=1)
