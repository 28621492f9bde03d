"""The drop-in boundary compiled against the REFERENCE's own headers.

tests/cpp/test_reference_boundary.cpp includes /root/reference/include (bbm.h, optimizer/compass.h, loss/*.h), the native
backbone's types through backbone/cuda/include/backbone.h, and backbone/cuda/include/bbm_cuda/loss.h; it static_asserts
concepts::config<floatRGB_cuda>, concepts::lossfunction / sampledlossfunction<bbm::cuda::loss<...>> and
concepts::optimization_algorithm<bbm::compass<bbm::cuda::loss<...>, ...>>, and runs the UNMODIFIED bbm::compass on the CUDA
loss for the example of docs/source/fitting.rst:16-60, next to the same search on the reference's own CPU loss.

The program can only be COMPILED where /root/reference exists (this container; __graft_entry__.build() does it) and only be
RUN where a GPU exists (the binary travels in tests/_build/, like the built libraries)."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_reference_boundary.cpp")
EXE = os.path.join(ROOT, "tests", "_build", "test_reference_boundary")
REF = "/root/reference"


def build(force=False):
    deps = [SRC, os.path.join(ROOT, "backbone", "cuda", "include", "backbone.h"), os.path.join(ROOT, "backbone", "cuda", "include", "bbm_cuda", "loss.h"),
            os.path.join(ROOT, "include", "bbmcu.h"), os.path.join(ROOT, "include", "bbmcu", "loss.hpp"), os.path.join(ROOT, "include", "bbmcu", "bsdf.hpp")]
    if not force and os.path.exists(EXE) and os.path.getmtime(EXE) >= max(os.path.getmtime(d) for d in deps):
        return EXE
    gen = os.path.join(ROOT, "oracle", "_ref", "gen")
    if not os.path.exists(os.path.join(gen, "bbm_bsdfmodels.h")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), os.path.join(gen, "bbm_bsdfmodels.h")], check=True, capture_output=True)
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.join(ROOT, "bbm_b200")
    cmd = ["g++", "-std=c++20", "-O2", "-Wno-attributes", "-DBBM_STRING_BSDF_IMPORTER", "-DBBM_BSDF_ENABLE_FORWARD", "-DBBM_NAME=bbm", "-DBBM_BACKBONE=cuda",
           "-I" + gen, "-I" + os.path.join(ROOT, "backbone", "cuda", "include"), "-I" + os.path.join(REF, "include"), "-I" + os.path.join(REF, "backbone", "native", "include"),
           "-I" + os.path.join(ROOT, "oracle", "stubs"), "-I" + os.path.join(ROOT, "include"), SRC, "-o", EXE,
           "-L" + libdir, "-l:libbbmcu.so", "-Wl,-rpath,$ORIGIN/../../bbm_b200"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-6000:])
    return EXE


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference's headers (/root/reference)")
def test_boundary_compiles_against_reference_concepts():
    """the static_asserts against bbm::concepts::{config, lossfunction, sampledlossfunction, optimization_algorithm} hold and
    the unmodified bbm::compass instantiates on bbm::cuda::loss"""
    import bbm_b200  # noqa: F401
    assert os.path.exists(build())


@pytest.mark.gpu
def test_unmodified_compass_on_cuda_loss_follows_the_reference():
    if not os.path.exists(EXE):
        if not os.path.isdir(REF):
            pytest.skip("tests/_build/test_reference_boundary was not built where the reference's headers are")
        build()
    steps = 30
    r = subprocess.run([EXE, str(steps)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        json.dump(out, open(os.path.join(out_dir, "reference_boundary_r02.json"), "w"), indent=1)
    assert out["samples"] == 90 * 30 * 1 * 9
    # the reference's own loss value is a sequential float sum (1e-3 .. 1e-2 away from the exact sum of its
    # own terms for the log metrics); per-sample terms are the tight comparison
    assert abs(out["term_12345"] - out["cpu_term_12345"]) <= 1e-5 * abs(out["cpu_term_12345"]) + 1e-12
    assert abs(out["initial_loss"] - out["cpu_initial_loss"]) <= 2e-2 * abs(out["cpu_initial_loss"])
    assert abs(out["gradient_loss"] - out["initial_loss"]) <= 1e-5 * abs(out["initial_loss"])
    cu, cp = np.array(out["cuda_trace"]), np.array(out["cpu_trace"])
    assert len(cu) == len(cp) == steps
    assert np.all(np.diff(cu) <= 1e-12)                                   # monotone: compass only accepts improvements
    # same decisions while float-sum noise does not decide ties; same place in the end
    assert np.all(np.abs(cu[:8] - cp[:8]) <= 2e-2 * cp[:8]), (cu, cp)
    assert abs(cu[-1] - cp[-1]) <= 5e-2 * cp[-1], (cu, cp)
    assert np.allclose(out["cuda_params"], out["cpu_params"], rtol=0.1, atol=0.02), (out["cuda_params"], out["cpu_params"])
    assert out["cuda_fitted"].startswith("Aggregate(Lambertian(albedo = [")
    assert out["cuda_seconds"] < out["cpu_seconds"]
