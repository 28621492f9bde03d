"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libbbmref_{float,double}.so, the
UNMODIFIED reference (bsdfbenchmark/bbm native backbone) built by oracle/Makefile.
Only tests/, __graft_entry__.smoke(), oracle/gen_golden.py and bench.py's reference /
cpu_baseline legs may import this module; the product package never does."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
METRICS = ["nganL2", "lowL2", "bieronL2", "lowLog", "bieronLog", "standardLog"]
NONE, DIFFUSE, SPECULAR, ALL = 0, 1, 2, 3
ATTR_ALL = 0x0F


class SphDesc(C.Structure):
    _fields_ = [("samplesIn", C.c_uint64 * 2), ("samplesOut", C.c_uint64 * 2),
                ("startIn", C.c_double * 2), ("endIn", C.c_double * 2),
                ("startOut", C.c_double * 2), ("endOut", C.c_double * 2)]


def sph_desc(samples_in, samples_out, start_in=(0.0, 0.0), end_in=None, start_out=(0.0, 0.0), end_out=None, real=np.float32):
    """spherical_linearizer constructor arguments (linearizer/spherical_linearizer.h:37-44);
    the defaults are Constants::Hemisphere() = (Pi(2), Pi(0.5)) in the config's Value type."""
    hemi = (float(real(2) * real(np.pi)), float(real(0.5) * real(np.pi)))
    d = SphDesc()
    d.samplesIn[:] = samples_in
    d.samplesOut[:] = samples_out
    d.startIn[:] = start_in
    d.endIn[:] = end_in or hemi
    d.startOut[:] = start_out
    d.endOut[:] = end_out or hemi
    return d


def available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libbbmref_float.so"))


class Ref:
    def __init__(self, precision="float"):
        self.real = np.float32 if precision == "float" else np.float64
        self.sfx = "_f" if precision == "float" else "_d"
        self.lib = C.CDLL(os.path.join(_HERE, "_ref", f"libbbmref_{precision}.so"))
        getattr(self.lib, "bbmref_last_error" + self.sfx).restype = C.c_char_p

    def _f(self, name):
        return getattr(self.lib, "bbmref_" + name + self.sfx)

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self._f("last_error")().decode())

    def _a(self, x, cols):
        x = np.ascontiguousarray(x, dtype=self.real)
        assert x.ndim == 2 and x.shape[1] == cols, x.shape
        return x

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def eval(self, bsdf, inn, out, component=ALL, unit=0, threads=1):
        inn, out = self._a(inn, 3), self._a(out, 3)
        rgb = np.empty_like(inn)
        self._chk(self._f("eval")(bsdf.encode(), component, unit, C.c_size_t(len(inn)), self._p(inn), self._p(out), self._p(rgb), threads))
        return rgb

    def sample(self, bsdf, out, xi, component=ALL, unit=0, threads=1):
        out, xi = self._a(out, 3), self._a(xi, 2)
        d = np.empty_like(out)
        pdf = np.empty(len(out), self.real)
        flag = np.empty(len(out), np.int32)
        self._chk(self._f("sample")(bsdf.encode(), component, unit, C.c_size_t(len(out)), self._p(out), self._p(xi), self._p(d), self._p(pdf), self._p(flag), threads))
        return d, pdf, flag

    def pdf(self, bsdf, inn, out, component=ALL, unit=0, threads=1):
        inn, out = self._a(inn, 3), self._a(out, 3)
        pdf = np.empty(len(inn), self.real)
        self._chk(self._f("pdf")(bsdf.encode(), component, unit, C.c_size_t(len(inn)), self._p(inn), self._p(out), self._p(pdf), threads))
        return pdf

    def reflectance(self, bsdf, out, component=ALL, unit=0, threads=1):
        out = self._a(out, 3)
        rgb = np.empty_like(out)
        self._chk(self._f("reflectance")(bsdf.encode(), component, unit, C.c_size_t(len(out)), self._p(out), self._p(rgb), threads))
        return rgb

    def sample_eval_pdf(self, bsdf, out, xi, threads=1):
        out, xi = self._a(out, 3), self._a(xi, 2)
        n = len(out)
        d = np.empty_like(out)
        spdf = np.empty(n, self.real)
        flag = np.empty(n, np.int32)
        rgb = np.empty_like(out)
        pdf = np.empty(n, self.real)
        self._chk(self._f("sample_eval_pdf")(bsdf.encode(), C.c_size_t(n), self._p(out), self._p(xi), self._p(d), self._p(spdf), self._p(flag), self._p(rgb), self._p(pdf), threads))
        return d, spdf, flag, rgb, pdf

    def params(self, bsdf, which=0, flag=ATTR_ALL):
        buf = np.empty(256, np.float64)
        n = C.c_int(0)
        self._chk(self._f("params")(bsdf.encode(), which, flag, self._p(buf), C.byref(n)))
        return buf[:n.value].copy()

    def to_string(self, bsdf):
        buf = C.create_string_buffer(1 << 14)
        self._chk(self._f("to_string")(bsdf.encode(), buf, C.c_size_t(len(buf))))
        return buf.value.decode()

    def merl_index(self, inn, out, threads=1):
        inn, out = self._a(inn, 3), self._a(out, 3)
        idx = np.empty(len(inn), np.uint64)
        self._chk(self._f("merl_index")(C.c_size_t(len(inn)), self._p(inn), self._p(out), self._p(idx), threads))
        return idx

    def merl_dirs(self, first, n, threads=1):
        inn = np.empty((n, 3), self.real)
        out = np.empty((n, 3), self.real)
        self._chk(self._f("merl_dirs")(C.c_size_t(first), C.c_size_t(n), self._p(inn), self._p(out), threads))
        return inn, out

    def spherical_dirs(self, desc, first, n):
        inn = np.empty((n, 3), self.real)
        out = np.empty((n, 3), self.real)
        self._chk(self._f("spherical_dirs")(C.byref(desc), C.c_size_t(first), C.c_size_t(n), self._p(inn), self._p(out)))
        return inn, out

    def spherical_index(self, desc, inn, out):
        inn, out = self._a(inn, 3), self._a(out, 3)
        idx = np.empty(len(inn), np.uint64)
        self._chk(self._f("spherical_index")(C.byref(desc), C.c_size_t(len(inn)), self._p(inn), self._p(out), self._p(idx)))
        return idx

    def loss(self, metric, desc, fitted, reference, first=0, n=0, want_total=True, threads=1):
        """returns (per-sample terms[first:first+n], the reference's own sequential total or None)"""
        terms = np.empty(n, self.real)
        total = self.real(0)
        tot = np.zeros(1, self.real)
        self._chk(self._f("loss")(METRICS.index(metric) if isinstance(metric, str) else metric,
                                  C.byref(desc) if desc is not None else None, fitted.encode(), reference.encode(),
                                  C.c_size_t(first), C.c_size_t(n), self._p(terms) if n else None,
                                  self._p(tot) if want_total else None, threads))
        return terms, (tot[0] if want_total else None)

    def loss_at(self, metric, desc, fitted, reference, params, accumulate_double=True, threads=1):
        params = np.ascontiguousarray(np.atleast_2d(params), np.float64)
        K, P = params.shape
        out = np.empty(K, np.float64)
        self._chk(self._f("loss_at")(METRICS.index(metric) if isinstance(metric, str) else metric,
                                     C.byref(desc) if desc is not None else None, fitted.encode(), reference.encode(),
                                     C.c_size_t(K), C.c_size_t(P), self._p(params), self._p(out), int(accumulate_double), threads))
        return out

    def compass(self, metric, desc, fitted, reference, max_steps):
        trace = np.zeros(max_steps, np.float64)
        steps = C.c_int(0)
        fp = np.zeros(256, np.float64)
        P = C.c_int(0)
        buf = C.create_string_buffer(1 << 14)
        self._chk(self._f("compass")(METRICS.index(metric) if isinstance(metric, str) else metric,
                                     C.byref(desc) if desc is not None else None, fitted.encode(), reference.encode(),
                                     max_steps, self._p(trace), C.byref(steps), self._p(fp), C.byref(P), buf, C.c_size_t(len(buf))))
        return trace[:steps.value], fp[:P.value].copy(), buf.value.decode()

    def import_fit(self, path):
        buf = C.create_string_buffer(1 << 22)
        n = C.c_int(0)
        self._chk(self._f("import_fit")(path.encode(), buf, C.c_size_t(len(buf)), C.byref(n)))
        out = {}
        for line in buf.value.decode().splitlines():
            k, v = line.split("\t", 1)
            out[k] = v
        return out
