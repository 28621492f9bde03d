"""The reference's Python module surface (include/python/py_core.h:44-190, py_bsdf.h:57-76) over the C ABI; shared by
bbm_b200.floatRGB and bbm_b200.doubleRGB (see floatRGB.py for the description)."""
import enum
import numbers

import numpy as np

from . import (ALL, ATTR_ALL, DIFFUSE, NONE, RADIANCE, SPECULAR, IMPORTANCE, Bsdf, BbmInvalidArgument, Context, model_names)

NAMES = ["BsdfPtr", "BsdfSample", "Aggregate", "bsdf_flag", "unit_t", "bsdf_attr", "parameter_values", "parameter_default_values",
         "parameter_lower_bound", "parameter_upper_bound", "RefValueVector", "ValueVector", "context"]


# include/bbm/bsdf_flag.h:21-27 (the member the reference calls None is reachable as getattr(bsdf_flag, "None") or bsdf_flag(0) in Python)
bsdf_flag = enum.IntFlag("bsdf_flag", [("None", NONE), ("Diffuse", DIFFUSE), ("Specular", SPECULAR), ("All", ALL)])


class unit_t(enum.IntEnum):                         # include/bbm/unit.h:20-24
    Radiance = RADIANCE
    Importance = IMPORTANCE


class bsdf_attr(enum.IntFlag):                      # include/bbm/bsdf_attr_flag.h:17-31
    DiffuseScale = 1
    DiffuseParameter = 2
    SpecularScale = 4
    SpecularParameter = 8
    Dependent = 16
    Diffuse = 3
    Specular = 12
    Scale = 5
    Parameter = 10
    All = 15


class BsdfSample:
    """include/bbm/bsdfsample.h: direction, pdf, flag"""
    __slots__ = ("direction", "pdf", "flag")

    def __init__(self, direction, pdf, flag):
        self.direction, self.pdf, self.flag = direction, pdf, flag

    def __str__(self):
        return "(direction = [%g, %g, %g], pdf = %g, flag = %s)" % (*self.direction, self.pdf, bsdf_flag(self.flag).name)


_ctx = None


def context():
    """the context single-direction calls run on (device 0, created on first use)"""
    global _ctx
    if _ctx is None:
        _ctx = Context(0)
    return _ctx


def _v(x, k):
    a = np.ascontiguousarray(np.asarray(x, np.float32).reshape(k, 1))
    return a


class BsdfPtr(Bsdf):
    """bbm::bsdf_ptr as the reference exports it to Python (include/python/py_core.h:104-112)"""
    CONFIG = "floatRGB"

    def __init__(self, other):
        super().__init__(other.to_string() if isinstance(other, Bsdf) else str(other), config=self.CONFIG)

    def eval(self, in_, out, component=bsdf_flag.All, unit=unit_t.Radiance, mask=True):
        if not mask:
            return np.zeros(3, np.float32)
        return context().eval(self, _v(in_, 3), _v(out, 3), int(component), int(unit))[:, 0].copy()

    def sample(self, out, xi, component=bsdf_flag.All, unit=unit_t.Radiance, mask=True):
        if not mask:
            return BsdfSample(np.zeros(3, np.float32), 0.0, bsdf_flag(0))
        d, p, f = context().sample(self, _v(out, 3), _v(xi, 2), int(component), int(unit))
        return BsdfSample(d[:, 0].copy(), float(p[0]), bsdf_flag(int(f[0])))

    def pdf(self, in_, out, component=bsdf_flag.All, unit=unit_t.Radiance, mask=True):
        if not mask:
            return 0.0
        return float(context().pdf(self, _v(in_, 3), _v(out, 3), int(component), int(unit))[0])

    def reflectance(self, out, component=bsdf_flag.All, unit=unit_t.Radiance, mask=True):
        if not mask:
            return np.zeros(3, np.float32)
        return context().reflectance(self, _v(out, 3), int(component), int(unit))[:, 0].copy()


def _fmt(v):
    """a Python value in the BSDF string grammar: numbers, (nested) lists -> [a, b, c], strings quoted, BSDFs as their toString"""
    if isinstance(v, Bsdf):
        return v.to_string()
    if isinstance(v, str):
        return '"%s"' % v
    if isinstance(v, bool):
        return "1" if v else "0"
    if isinstance(v, numbers.Real):
        return repr(float(v))
    if isinstance(v, (list, tuple, np.ndarray)):
        return "[" + ", ".join(_fmt(x) for x in v) + "]"
    raise BbmInvalidArgument("cannot convert %r to a BSDF constructor argument" % (v,))


def _factory(name, cls):
    def make(*args, **kwargs):
        parts = [_fmt(a) for a in args] + ["%s = %s" % (k, _fmt(v)) for k, v in kwargs.items()]
        return cls("%s(%s)" % (name, ", ".join(parts)))
    make.__name__ = name
    make.__doc__ = "Constructs: %s(...) - positional and/or named constructor arguments of the model (see bbm_b200.model_layout(%r))" % (name, name)
    return make


def _aggregate(cls):
    def Aggregate(*bsdfs):
        """AggregateBsdf(BsdfPtr...) combines as many BsdfPtrs as provided (include/python/py_core.h:115-122)"""
        for b in bsdfs:
            if not isinstance(b, Bsdf):
                raise TypeError("Aggregate() takes BsdfPtr arguments")
        return cls("Aggregate(%s)" % ", ".join(b.to_string() for b in bsdfs))
    return Aggregate


class RefValueVector:
    """bbm::vector<Value&> (include/python/py_core.h:127-150): a live view of a BSDF's parameters - item assignment writes
    through to the BSDF.  Arithmetic returns plain numpy arrays."""

    def __init__(self, bsdf, flags=ATTR_ALL):
        self._b, self._flags = bsdf, int(flags)

    def _get(self):
        return self._b.parameter_values(self._flags)

    def __len__(self):
        return len(self._get())

    def __getitem__(self, i):
        return float(np.float32(self._get()[i]))

    def __setitem__(self, i, val):
        v = self._get()
        v[i] = val
        self._b.set_parameter_values(v, self._flags)

    def __iter__(self):
        return iter(float(np.float32(x)) for x in self._get())

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self._get(), dtype or np.float64)

    def __str__(self):
        return "(" + ", ".join("%g" % x for x in self._get()) + ")"

    def _inplace(self, other, op):
        self._b.set_parameter_values(op(self._get(), np.asarray(other, np.float64)), self._flags)
        return self

    def __add__(self, o): return self._get() + np.asarray(o, np.float64)
    def __sub__(self, o): return self._get() - np.asarray(o, np.float64)
    def __mul__(self, o): return self._get() * np.asarray(o, np.float64)
    def __truediv__(self, o): return self._get() / np.asarray(o, np.float64)
    def __iadd__(self, o): return self._inplace(o, np.add)
    def __isub__(self, o): return self._inplace(o, np.subtract)
    def __imul__(self, o): return self._inplace(o, np.multiply)
    def __itruediv__(self, o): return self._inplace(o, np.divide)


class ValueVector(list):
    """bbm::vector<Value> (include/python/py_core.h:152-176)"""

    def __str__(self):
        return "(" + ", ".join("%g" % x for x in self) + ")"


def parameter_values(bsdf, flag=bsdf_attr.All):
    """List all parameter values of a given BSDF (a LIVE view: assignment writes into the BSDF)"""
    return RefValueVector(bsdf, flag)


def parameter_default_values(bsdf, flag=bsdf_attr.All):
    return ValueVector(float(np.float32(x)) for x in bsdf.parameter_default_values(int(flag)))


def parameter_lower_bound(bsdf, flag=bsdf_attr.All):
    return ValueVector(float(np.float32(x)) for x in bsdf.parameter_lower_bound(int(flag)))


def parameter_upper_bound(bsdf, flag=bsdf_attr.All):
    return ValueVector(float(np.float32(x)) for x in bsdf.parameter_upper_bound(int(flag)))


def populate(namespace, config):
    """fill a module namespace with the reference's Python surface for one configuration"""
    cls = BsdfPtr if config == "floatRGB" else type("BsdfPtr", (BsdfPtr,), {"CONFIG": config, "__doc__": BsdfPtr.__doc__})
    here = globals()
    exported = list(NAMES)
    for n in NAMES:
        if n in here:
            namespace[n] = here[n]
    namespace["BsdfPtr"] = cls
    namespace["Aggregate"] = _aggregate(cls)
    for name in model_names():
        namespace[name] = _factory(name, cls)
        exported.append(name)
    namespace["__all__"] = exported
