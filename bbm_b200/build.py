"""Build libbbmcu.so (sm_100a) in-tree with nvcc.  `python -m bbm_b200.build [--force] [-v]`.

Each .cu / .cpp under csrc/ is one object; objects are compiled in parallel and rebuilt only when a
source or header is newer.  Flags: -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false
(-fmad=false: the reference runs on x86-64 without FMA contraction; parity of sharp lobes and of the
MERL bin index depends on the float operation sequence)."""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libbbmcu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
DEFINES = [d for d in os.environ.get("BBMCU_DEFINES", "").split() if d]        # e.g. BBMCU_DEFINES="-DBBMCU_LOSS_MINB_EXPERIMENT" for tuning builds
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
              "-diag-suppress", "20012,20011,20014,177,550"] + DEFINES
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-ffp-contract=off"]


_INC = None


def _headers_of(path, seen=None):
    """the quoted includes of `path`, followed recursively through csrc/ and include/ (a translation unit is rebuilt only
    when one of ITS headers is newer: a full build is ~35 CPU-minutes)"""
    import re
    global _INC
    if _INC is None:
        _INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)
    seen = set() if seen is None else seen
    try:
        text = open(path).read()
    except OSError:
        return seen
    for name in _INC.findall(text):
        for base in (os.path.dirname(path), CSRC, os.path.join(HERE, "..", "include")):
            h = os.path.normpath(os.path.join(base, name))
            if os.path.exists(h):
                if h not in seen:
                    seen.add(h)
                    _headers_of(h, seen)
                break
    return seen


def _newest_header(src=None):
    if src is not None:
        hs = _headers_of(src)
        return max([os.path.getmtime(h) for h in hs] + [0.0])
    hs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.hpp")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.basename(src) + ".o")
    if src.endswith(".cu"):
        flags = list(NVCC_FLAGS)
        base = os.path.basename(src)
        if base.startswith(("bbmcu_loss_single_", "bbmcu_loss_pair_")):
            # the batched loss(+gradient) kernels are judged at 1e-5 per term / 1e-4 per total, not bit for bit: let the
            # compiler contract a*b+c there (no linearizer or sampler code lives in these translation units)
            flags[flags.index("-fmad=false")] = "-fmad=true"
        cmd = [NVCC] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    else:
        cmd = ["g++"] + CXX_FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, obj, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))
    todo, objs = [], []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s) + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), _newest_header(s)):
            todo.append(s)
    # model data linked into the library: data/epd_g1.f32 -> _binary_epd_g1_f32_start/_end
    data_obj = os.path.join(OBJ, "epd_g1.f32.o")
    data_src = os.path.join(HERE, "data", "epd_g1.f32")
    relink = False
    if force or not os.path.exists(data_obj) or os.path.getmtime(data_obj) < os.path.getmtime(data_src):
        r = subprocess.run(["ld", "-r", "-b", "binary", "-z", "noexecstack", "-o", data_obj, "epd_g1.f32"], cwd=os.path.join(HERE, "data"), capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("embedding data/epd_g1.f32 failed:\n" + r.stdout + r.stderr)
        relink = True
    objs.append(data_obj)
    log = []
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
            for src, obj, rc, out in ex.map(lambda s: _compile(s, verbose), todo):
                log.append((src, out))
                if rc != 0:
                    raise RuntimeError(f"compiling {src} failed:\n{out}")
    if todo or relink or not os.path.exists(LIB) or force:
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-lineinfo"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("linking libbbmcu.so failed:\n" + r.stdout + r.stderr)
    if verbose:
        for src, out in log:
            print("==", src)
            print(out)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
