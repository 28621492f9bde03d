"""The C++ adapter headers (include/bbmcu/*.hpp): they compile as C++20 on CPU; on the GPU box a program written
against them the way docs/source/fitting.rst uses bbm is built, run and compared with the golden vectors of the
unmodified reference and with the reference's own compass search."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_adapters.cpp")
EXE = os.path.join(ROOT, "tests", "_build", "test_adapters")


def _build():
    import bbm_b200  # noqa: F401  (libbbmcu.so must exist)
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.join(ROOT, "bbm_b200")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")           # cudaMalloc for the device buffers of the peer-exchange block
    cmd = ["g++", "-std=c++20", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"), SRC, "-o", EXE,
           "-L" + libdir, "-l:libbbmcu.so", "-Wl,-rpath," + libdir, "-L" + os.path.join(cuda, "lib64"), "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_adapter_headers_compile_and_link():
    """no device needed: the headers are valid C++20 and every C-ABI symbol they use resolves"""
    _build()


@pytest.mark.gpu
def test_adapter_program_against_reference(golden_loss, ref):
    _build()
    steps = 30
    r = subprocess.run([EXE, str(steps)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    arr, meta = golden_loss
    m = meta["metrics"]["nganL2"]
    assert out["samples"] == m["N"]
    assert abs(out["loss0"] - m["double_total"]) <= 1e-4 * abs(m["double_total"])
    fd = np.array(m["fd_gradient"])
    assert np.all(np.abs(np.array(out["gradient"]) - fd) <= 1e-4 * np.abs(fd) + 1e-9)
    # two shards on two streams, combined over peer memory inside the library's kernels: identical on both, equal to the whole
    assert out["peer_identical"] is True
    assert abs(out["peer_loss"] - out["whole_loss"]) <= 1e-5 * abs(out["whole_loss"])
    assert abs(out["peer_grad0"] - out["whole_grad0"]) <= 1e-5 * abs(out["whole_grad0"]) + 1e-12
    # scalar concept calls against the compiled reference
    truth = meta["truth"]
    i = np.array([[0.3, 0.2, 0.9327379]], np.float32)
    o = np.array([[0.5, -0.1, 0.8602325]], np.float32)
    assert np.allclose(out["eval"], ref.eval(truth, i, o)[0], rtol=1e-5)
    assert np.allclose(out["pdf"], ref.pdf(truth, i, o)[0], rtol=1e-5)
    d, p, f = ref.sample(truth, o, np.array([[0.3, 0.6]], np.float32))
    assert np.allclose(out["sample"][:3], d[0], rtol=1e-5, atol=1e-5) and out["sample"][4] == f[0]
    assert np.allclose(out["reflectance"], ref.reflectance(truth, o)[0], rtol=1e-5)
    # compass: the reference's own optimizer on the same problem (float sums decide ties after a few steps, so the
    # traces are compared step by step only while they agree to 1e-4, and the final losses within 2 %)
    from oracle.refbind import sph_desc
    trace, _, _ = ref.compass("nganL2", sph_desc((13, 8), (5, 6)), "CookTorrance()", "CookTorrance([0.3, 0.3, 0.3], 0.2, 1.5)", steps)
    mine = np.array(out["compass_trace"])
    assert len(mine) == len(trace)
    assert abs(mine[0] - trace[0]) <= 1e-4 * trace[0]
    agree = np.abs(mine - trace) <= 1e-3 * trace
    assert agree[:10].all(), (mine, trace)                     # identical decisions for the first steps
    assert abs(mine[-1] - trace[-1]) <= 5e-2 * trace[-1], (mine, trace)
    assert np.all(np.diff(mine) <= 1e-12)                      # monotone
    batched = np.array(out["batched_trace"])
    assert abs(batched[-1] - trace[-1]) <= 5e-2 * trace[-1], (batched, trace)
    assert out["batched_launches"] <= 2 * steps + 2              # one loss launch (+ finish kernel) per step
    assert out["gd_last"] < 0.2 * out["gd_first"]
    assert out["invalid_argument"] is True
    assert out["toString"].startswith("Aggregate(Lambertian(")
