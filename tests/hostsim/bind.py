"""TEST INFRASTRUCTURE ONLY: ctypes binding of tests/_build/libbbmcu_hostsim.so (the device headers
compiled for the host).  Builds it on demand with g++."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
LIB = os.path.join(ROOT, "tests", "_build", "libbbmcu_hostsim.so")


def build(force=False):
    srcs = [os.path.join(HERE, "hostsim.cpp"), os.path.join(ROOT, "bbm_b200", "csrc", "bbmcu_host.cpp")]
    deps = srcs + [os.path.join(ROOT, "bbm_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "bbm_b200", "csrc")) if f.endswith((".cuh", ".hpp"))]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cuda_inc = os.environ.get("CUDA_INC", "/usr/local/cuda/include")
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I" + cuda_inc,
           "-I" + os.path.join(ROOT, "bbm_b200", "csrc"), "-o", LIB] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-4000:])
    return LIB


def soa(a):
    """(n, k) AoS -> contiguous (k, n) SoA float32"""
    return np.ascontiguousarray(np.asarray(a, np.float32).T)


class HostSim:
    def __init__(self):
        self.lib = C.CDLL(build())
        self.lib.hostsim_last_error.restype = C.c_char_p
        self.lib.hostsim_libm_mismatches.restype = C.c_size_t
        self.lib.hostsim_gamma_q_inv.restype = C.c_float
        self.lib.hostsim_gamma_q_inv.argtypes = [C.c_float, C.c_float]
        self._epd = np.fromfile(os.path.join(ROOT, "bbm_b200", "data", "epd_g1.f32"), np.float32)
        assert self._epd.size == 100000
        self.lib.hostsim_set_epd_table(self._epd.ctypes.data_as(C.c_void_p))

    def _chk(self, rc):
        if rc:
            raise RuntimeError(self.lib.hostsim_last_error().decode())

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def eval(self, bsdf, inn, out, component=3):
        i, o = soa(inn), soa(out)
        n = i.shape[1]
        r = np.empty((3, n), np.float32)
        self._chk(self.lib.hostsim_eval(bsdf.encode(), component, self._p(i), self._p(o), C.c_size_t(n), self._p(r)))
        return r.T.copy()

    def pdf(self, bsdf, inn, out, component=3):
        i, o = soa(inn), soa(out)
        n = i.shape[1]
        r = np.empty(n, np.float32)
        self._chk(self.lib.hostsim_pdf(bsdf.encode(), component, self._p(i), self._p(o), C.c_size_t(n), self._p(r)))
        return r

    def reflectance(self, bsdf, out, component=3):
        o = soa(out)
        n = o.shape[1]
        r = np.empty((3, n), np.float32)
        self._chk(self.lib.hostsim_reflectance(bsdf.encode(), component, self._p(o), C.c_size_t(n), self._p(r)))
        return r.T.copy()

    def sample(self, bsdf, out, xi, component=3):
        o, x = soa(out), soa(xi)
        n = o.shape[1]
        d = np.empty((3, n), np.float32)
        p = np.empty(n, np.float32)
        f = np.empty(n, np.int32)
        self._chk(self.lib.hostsim_sample(bsdf.encode(), component, self._p(o), self._p(x), C.c_size_t(n), self._p(d), self._p(p), self._p(f)))
        return d.T.copy(), p, f

    def eval_pre(self, bsdf, inn, out, component=3):
        """eval of a single Student-t lobe with its parameter-only factors formed once (EvalOp::group_pre)"""
        i, o = soa(inn), soa(out)
        n = i.shape[1]
        r = np.empty((3, n), np.float32)
        self._chk(self.lib.hostsim_eval_pre(bsdf.encode(), component, self._p(i), self._p(o), C.c_size_t(n), self._p(r)))
        return r.T.copy()

    def sample_eval_pdf_pre(self, bsdf, out, xi, component=3):
        """the fused pass of a single Student-t lobe with its parameter-only factors formed once: (dir, sample_pdf, flag, rgb, pdf)"""
        o, x = soa(out), soa(xi)
        n = o.shape[1]
        d = np.empty((3, n), np.float32)
        sp = np.empty(n, np.float32)
        f = np.empty(n, np.int32)
        rgb = np.empty((3, n), np.float32)
        p = np.empty(n, np.float32)
        self._chk(self.lib.hostsim_sample_eval_pdf_pre(bsdf.encode(), component, self._p(o), self._p(x), C.c_size_t(n), self._p(d), self._p(sp), self._p(f),
                                                       self._p(rgb), self._p(p)))
        return d.T.copy(), sp, f, rgb.T.copy(), p

    def merl_index(self, inn, out):
        i, o = soa(inn), soa(out)
        n = i.shape[1]
        r = np.empty(n, np.uint32)
        self._chk(self.lib.hostsim_merl_index(self._p(i), self._p(o), C.c_size_t(n), self._p(r)))
        return r

    def merl_dirs(self, first, n):
        i = np.empty((3, n), np.float32)
        o = np.empty((3, n), np.float32)
        self._chk(self.lib.hostsim_merl_dirs(C.c_uint32(first), C.c_size_t(n), self._p(i), self._p(o)))
        return i.T.copy(), o.T.copy()

    def spherical_dirs(self, samples, ranges, first, n):
        s = np.asarray(samples, np.uint32)
        r = np.asarray(ranges, np.float32)
        i = np.empty((3, n), np.float32)
        o = np.empty((3, n), np.float32)
        self._chk(self.lib.hostsim_spherical_dirs(self._p(s), self._p(r), C.c_uint64(first), C.c_size_t(n), self._p(i), self._p(o)))
        return i.T.copy(), o.T.copy()

    def libm_mismatches(self, which, a, b=None):
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b if b is not None else a, np.float32)
        return self.lib.hostsim_libm_mismatches(which, self._p(a), self._p(b), C.c_size_t(len(a)))

    def loss(self, bsdf, metric, inn, out, ref, component=3, want_grad=True, nparams=0):
        i, o, r = soa(inn), soa(out), soa(ref)
        n = i.shape[1]
        loss = C.c_double(0)
        grad = np.zeros(max(nparams, 1), np.float64)
        terms = np.empty(n, np.float32)
        self._chk(self.lib.hostsim_loss(bsdf.encode(), metric, component, self._p(i), self._p(o), self._p(r), C.c_size_t(n), C.c_double(1.0 / n),
                                        C.byref(loss), self._p(grad) if want_grad else None, self._p(terms)))
        return loss.value, grad[:nparams], terms

    def loss_compact(self, bsdf, metric, inn, out, ref, want_grad=True, nparams=0):
        """the compact pair loss (bbmcu_losscompact.cuh) on the host; None if the BSDF has no compact kernel"""
        i, o, r = soa(inn), soa(out), soa(ref)
        n = i.shape[1]
        loss = C.c_double(0)
        grad = np.zeros(max(nparams, 1), np.float64)
        rc = self.lib.hostsim_loss_compact(bsdf.encode(), metric, self._p(i), self._p(o), self._p(r), C.c_size_t(n), C.c_double(1.0 / n),
                                           C.byref(loss), self._p(grad) if want_grad else None)
        if rc == 2:
            return None
        self._chk(rc)
        return loss.value, grad[:nparams]
