"""bbm_b200 - thin Python binding of libbbmcu.so (include/bbmcu.h), the B200 CUDA backbone for bbm's
batched BSDF eval / sample / pdf / reflectance, linearizers and fitting losses.

Everything here forwards to the C ABI; there is no Python or CPU implementation of the path.  If the
shared library is missing or no CUDA device is usable the calls raise - nothing falls back.

Arrays are struct-of-arrays: directions / spectra have shape (3, n), xi (2, n), float32, C-contiguous.
They may be numpy arrays (host memory, staged by the library) or torch tensors (CPU or CUDA).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbbmcu.so")

NONE, DIFFUSE, SPECULAR, ALL = 0, 1, 2, 3
RADIANCE, IMPORTANCE = 0, 1
ATTR_DIFFUSE_SCALE, ATTR_DIFFUSE_PARAMETER, ATTR_SPECULAR_SCALE, ATTR_SPECULAR_PARAMETER, ATTR_DEPENDENT, ATTR_ALL = 1, 2, 4, 8, 16, 15
METRICS = ["nganL2", "lowL2", "bieronL2", "lowLog", "bieronLog", "standardLog"]
PARAM_VALUE, PARAM_DEFAULT, PARAM_LOWER, PARAM_UPPER = 0, 1, 2, 3
MERL_BINS = 1458000
CONFIGS = ["floatRGB", "doubleRGB"]


class BbmError(RuntimeError):
    """std::runtime_error on the reference side (core/error.h:42-46)"""


class BbmInvalidArgument(BbmError, ValueError):
    """std::invalid_argument on the reference side (core/stringconvert.h:68,517,540,565)"""


class Attr(C.Structure):
    _fields_ = [("name", C.c_char_p), ("width", C.c_int), ("rows", C.c_int), ("flag", C.c_int), ("offset", C.c_int)]


class SphericalGrid(C.Structure):
    """spherical_linearizer constructor arguments (linearizer/spherical_linearizer.h:37-44)"""
    _fields_ = [("samples_in", C.c_uint32 * 2), ("samples_out", C.c_uint32 * 2),
                ("start_in", C.c_float * 2), ("end_in", C.c_float * 2),
                ("start_out", C.c_float * 2), ("end_out", C.c_float * 2)]

    def size(self):
        return self.samples_in[0] * self.samples_in[1] * self.samples_out[0] * self.samples_out[1]


_lib = None


def lib():
    """load libbbmcu.so (built in-tree by `python -m bbm_b200.build`); raises if it is missing"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BbmError(f"{LIB_PATH} not found: build it with `python -m bbm_b200.build` (there is no fallback path)")
        L = C.CDLL(LIB_PATH)
        L.bbmcu_last_error.restype = C.c_char_p
        L.bbmcu_last_error.argtypes = [C.c_void_p]
        L.bbmcu_model_name.restype = C.c_char_p
        L.bbmcu_fit_key.restype = C.c_char_p
        L.bbmcu_fit_key.argtypes = [C.c_void_p, C.c_int]
        L.bbmcu_stream.restype = C.c_void_p
        L.bbmcu_stream.argtypes = [C.c_void_p]
        L.bbmcu_launch_count.restype = C.c_uint64
        L.bbmcu_launch_count.argtypes = [C.c_void_p]
        L.bbmcu_loss_samples.restype = C.c_uint64
        L.bbmcu_loss_samples.argtypes = [C.c_void_p]
        L.bbmcu_loss_shard_count.restype = C.c_uint64
        L.bbmcu_loss_shard_count.argtypes = [C.c_void_p]
        L.bbmcu_loss_materials.argtypes = [C.c_void_p]
        L.bbmcu_set_plane_stride.argtypes = [C.c_void_p, C.c_size_t]
        L.bbmcu_host_register.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.bbmcu_host_unregister.argtypes = [C.c_void_p, C.c_void_p]
        for name in ("bbmcu_destroy", "bbmcu_bsdf_free", "bbmcu_loss_free", "bbmcu_fit_free"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _check(rc, ctx=None):
    if rc == 0:
        return
    msg = lib().bbmcu_last_error(ctx).decode(errors="replace")
    raise (BbmInvalidArgument if rc == 1 else BbmError)(msg)


def _ptr(a):
    """device or host address of a torch tensor / numpy array"""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        if not a.is_contiguous():
            raise BbmInvalidArgument("tensor must be contiguous")
        return C.c_void_p(a.data_ptr())
    a = np.asarray(a)
    if not a.flags["C_CONTIGUOUS"]:
        raise BbmInvalidArgument("array must be C-contiguous")
    return C.c_void_p(a.ctypes.data)


def _like(a, shape, dtype=np.float32):
    """output buffer living where `a` lives"""
    if hasattr(a, "data_ptr"):
        import torch
        td = {np.float32: torch.float32, np.int32: torch.int32, np.uint32: torch.int32}[dtype]
        return torch.empty(shape, dtype=td, device=a.device)
    return np.empty(shape, dtype)


def _n(a, rows):
    shape = tuple(a.shape)
    if len(shape) != 2 or shape[0] != rows:
        raise BbmInvalidArgument(f"expected a ({rows}, n) struct-of-arrays buffer, got {shape}")
    if str(a.dtype) not in ("float32", "torch.float32"):
        raise BbmInvalidArgument(f"expected float32, got {a.dtype}")
    return shape[1]


def _ld(a):
    """plane stride in elements of a (rows, n) buffer: n when contiguous, larger for a column slice x[:, :n] of a padded
    buffer (planes then start 16-byte aligned for any n, see bbmcu_set_plane_stride)"""
    if hasattr(a, "data_ptr"):
        st = a.stride()
    else:
        st = tuple(x // a.itemsize for x in a.strides)
    if len(st) != 2 or st[1] != 1:
        raise BbmInvalidArgument("planes of a struct-of-arrays buffer must be contiguous rows")
    return int(st[0]) if a.shape[0] > 1 else int(a.shape[1])


def _sptr(a):
    """address of a (rows, n) buffer whose rows may be strided (see _ld)"""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


def model_names():
    L = lib()
    return [L.bbmcu_model_name(i).decode() for i in range(L.bbmcu_model_count())]


def model_layout(name):
    """[(attribute name, width, rows, flag, offset)] in reflection order (util/reflection.h:141-148)"""
    L = lib()
    mid = C.c_int(0)
    _check(L.bbmcu_model_lookup(name.encode(), C.byref(mid)))
    n = C.c_int(0)
    _check(L.bbmcu_model_layout(mid, None, C.byref(n)))
    arr = (Attr * n.value)()
    _check(L.bbmcu_model_layout(mid, arr, C.byref(n)))
    return [(a.name.decode(), a.width, a.rows, a.flag, a.offset) for a in arr]


def spherical_grid(samples_in, samples_out, start_in=None, end_in=None, start_out=None, end_out=None):
    g = SphericalGrid()
    lib().bbmcu_spherical_grid_default(C.byref(g), samples_in[0], samples_in[1], samples_out[0], samples_out[1])
    for name, v in (("start_in", start_in), ("end_in", end_in), ("start_out", start_out), ("end_out", end_out)):
        if v is not None:
            getattr(g, name)[:] = [float(np.float32(x)) for x in v]
    return g


class Bsdf:
    """bbm::bsdf_ptr (include/bbm/bsdf_ptr.h:20-165): import from / export to the BSDF string grammar and
    the parameter enumeration of include/bbm/bsdf_enumerate.h:102-237 (forward order)."""

    def __init__(self, string, _handle=None, config="floatRGB"):
        """config: which of the reference's native configurations the host side mirrors - "floatRGB" parses with std::stof and
        prints floats, "doubleRGB" keeps doubles (needed e.g. for fits/bagher_sgd.fit); the kernels compute in FP32 either way"""
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = _handle
        else:
            _check(lib().bbmcu_bsdf_from_string_ex(None, string.encode(), C.c_int(CONFIGS.index(config)), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.bbmcu_bsdf_free(self._h)
            self._h = None

    def to_string(self):
        buf = C.create_string_buffer(1 << 14)
        _check(lib().bbmcu_bsdf_to_string(self._h, buf, C.c_size_t(len(buf))))
        return buf.value.decode()

    __str__ = to_string

    def _vec(self, which, flags):
        n = C.c_int(0)
        _check(lib().bbmcu_bsdf_get_params(self._h, which, flags, None, C.byref(n)))
        v = np.empty(n.value, np.float64)
        _check(lib().bbmcu_bsdf_get_params(self._h, which, flags, v.ctypes.data_as(C.c_void_p), C.byref(n)))
        return v

    def parameter_values(self, flags=ATTR_ALL):
        return self._vec(PARAM_VALUE, flags)

    def parameter_default_values(self, flags=ATTR_ALL):
        return self._vec(PARAM_DEFAULT, flags)

    def parameter_lower_bound(self, flags=ATTR_ALL):
        return self._vec(PARAM_LOWER, flags)

    def parameter_upper_bound(self, flags=ATTR_ALL):
        return self._vec(PARAM_UPPER, flags)

    def set_parameter_values(self, values, flags=ATTR_ALL):
        v = np.ascontiguousarray(values, np.float64)
        _check(lib().bbmcu_bsdf_set_params(self._h, flags, v.ctypes.data_as(C.c_void_p), len(v)))


class Context:
    """one CUDA device + stream (bbmcu_ctx)"""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _check(lib().bbmcu_init(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.bbmcu_destroy(self._h)
            self._h = None

    __del__ = close

    def synchronize(self):
        _check(lib().bbmcu_synchronize(self._h), self._h)

    @property
    def stream(self):
        return lib().bbmcu_stream(self._h)

    @property
    def launches(self):
        return int(lib().bbmcu_launch_count(self._h))

    # ---- the BSDF concept, batched (concepts/bsdfmodel.h:32-146) -------------------------------
    def eval(self, bsdf, in_xyz, out_xyz, component=ALL, unit=RADIANCE, rgb=None):
        n = _n(in_xyz, 3)
        if _n(out_xyz, 3) != n:
            raise BbmInvalidArgument("in/out batch sizes differ")
        rgb = _like(in_xyz, (3, n)) if rgb is None else rgb
        _check(lib().bbmcu_eval(self._h, bsdf._h, component, unit, _ptr(in_xyz), _ptr(out_xyz), C.c_size_t(n), _ptr(rgb)), self._h)
        return rgb

    def pdf(self, bsdf, in_xyz, out_xyz, component=ALL, unit=RADIANCE, pdf=None):
        n = _n(in_xyz, 3)
        if _n(out_xyz, 3) != n:
            raise BbmInvalidArgument("in/out batch sizes differ")
        pdf = _like(in_xyz, (n,)) if pdf is None else pdf
        _check(lib().bbmcu_pdf(self._h, bsdf._h, component, unit, _ptr(in_xyz), _ptr(out_xyz), C.c_size_t(n), _ptr(pdf)), self._h)
        return pdf

    def reflectance(self, bsdf, out_xyz, component=ALL, unit=RADIANCE, rgb=None):
        n = _n(out_xyz, 3)
        rgb = _like(out_xyz, (3, n)) if rgb is None else rgb
        _check(lib().bbmcu_reflectance(self._h, bsdf._h, component, unit, _ptr(out_xyz), C.c_size_t(n), _ptr(rgb)), self._h)
        return rgb

    def sample(self, bsdf, out_xyz, xi_uv, component=ALL, unit=RADIANCE, outputs=None):
        n = _n(out_xyz, 3)
        if _n(xi_uv, 2) != n:
            raise BbmInvalidArgument("out/xi batch sizes differ")
        d, p, f = outputs if outputs is not None else (_like(out_xyz, (3, n)), _like(out_xyz, (n,)), _like(out_xyz, (n,), np.int32))
        _check(lib().bbmcu_sample(self._h, bsdf._h, component, unit, _ptr(out_xyz), _ptr(xi_uv), C.c_size_t(n), _ptr(d), _ptr(p), _ptr(f)), self._h)
        return d, p, f

    OUTPUT_NAMES = ("dir", "sample_pdf", "flag", "rgb", "pdf")

    def _sep_outputs(self, proto, n, outputs, want):
        if outputs is None:
            want = self.OUTPUT_NAMES if want is None else want
            mk = {"dir": lambda: _like(proto, (3, n)), "sample_pdf": lambda: _like(proto, (n,)), "flag": lambda: _like(proto, (n,), np.int32),
                  "rgb": lambda: _like(proto, (3, n)), "pdf": lambda: _like(proto, (n,))}
            outputs = tuple(mk[k]() if k in want else None for k in self.OUTPUT_NAMES)
        if all(o is None for o in outputs):
            raise BbmInvalidArgument("at least one output is needed")
        return outputs

    def sample_eval_pdf(self, bsdf, out_xyz, xi_uv, component=ALL, unit=RADIANCE, outputs=None, want=None):
        """s = sample(out, xi); rgb = eval(s.direction, out); pdf = pdf(s.direction, out).
        Returns (dir, sample_pdf, flag, rgb, pdf); entries not named in `want` (or None in `outputs`) are neither stored
        nor copied back.  A column slice x[:, :n] of a padded buffer keeps 16-byte accesses for any n."""
        n = _n(out_xyz, 3)
        if _n(xi_uv, 2) != n:
            raise BbmInvalidArgument("out/xi batch sizes differ")
        outputs = self._sep_outputs(out_xyz, n, outputs, want)
        d, sp, f, rgb, p = outputs
        ld = _ld(out_xyz)
        if any(_ld(a) != ld for a in (xi_uv, d, rgb) if a is not None):
            raise BbmInvalidArgument("all struct-of-arrays arguments of one call must share one plane stride")
        if ld != n:
            _check(lib().bbmcu_set_plane_stride(self._h, ld), self._h)
        try:
            _check(lib().bbmcu_sample_eval_pdf(self._h, bsdf._h, component, unit, _sptr(out_xyz), _sptr(xi_uv), C.c_size_t(n),
                                               _sptr(d), _ptr(sp), _ptr(f), _sptr(rgb), _ptr(p)), self._h)
        finally:
            if ld != n:
                lib().bbmcu_set_plane_stride(self._h, 0)
        return outputs

    def sample_eval_pdf_generated(self, bsdf, seed, first, n, component=ALL, unit=RADIANCE, like=None, outputs=None, want=None, inputs=False):
        """the same pass over inputs drawn on the device from (seed, first + i) - 0 bytes in.  Returns
        (dir, sample_pdf, flag, rgb, pdf) and, with inputs=True, also the generated (out_xyz, xi_uv)."""
        proto = like if like is not None else np.empty(0, np.float32)
        outputs = self._sep_outputs(proto, n, outputs, want)
        d, sp, f, rgb, p = outputs
        go, gx = (_like(proto, (3, n)), _like(proto, (2, n))) if inputs is True else (inputs if inputs else (None, None))
        _check(lib().bbmcu_sample_eval_pdf_generated(self._h, bsdf._h, component, unit, C.c_uint64(seed), C.c_uint64(first), C.c_size_t(n),
                                                     _ptr(go), _ptr(gx), _ptr(d), _ptr(sp), _ptr(f), _ptr(rgb), _ptr(p)), self._h)
        return (outputs, (go, gx)) if inputs else outputs

    def eval_merl_grid(self, bsdf, first=0, n=MERL_BINS, component=ALL, unit=RADIANCE, like=None, rgb=None, dirs=False):
        """eval at (in, out) = merl_linearizer(idx), idx = first .. first+n-1, directions generated inside the kernel
        (0 B in, 12 B out per eval).  dirs=True also returns the generated (in, out) - bit-identical to merl_dirs()."""
        proto = like if like is not None else (rgb if rgb is not None else np.empty(0, np.float32))
        rgb = _like(proto, (3, n)) if rgb is None else rgb
        i, o = (_like(proto, (3, n)), _like(proto, (3, n))) if dirs is True else (dirs if dirs else (None, None))
        _check(lib().bbmcu_eval_merl_grid(self._h, bsdf._h, component, unit, C.c_uint32(first), C.c_size_t(n), _ptr(rgb), _ptr(i), _ptr(o)), self._h)
        return (rgb, i, o) if dirs else rgb

    def host_register(self, array):
        """page-lock a numpy array once so the host-pointer path DMAs it in place (bbmcu_host_register)"""
        _check(lib().bbmcu_host_register(self._h, C.c_void_p(array.ctypes.data), C.c_size_t(array.nbytes)), self._h)

    def host_unregister(self, array):
        _check(lib().bbmcu_host_unregister(self._h, C.c_void_p(array.ctypes.data)), self._h)

    # ---- linearizers ---------------------------------------------------------------------------
    def merl_index(self, in_xyz, out_xyz, index=None):
        n = _n(in_xyz, 3)
        index = _like(in_xyz, (n,), np.uint32) if index is None else index
        _check(lib().bbmcu_merl_index(self._h, _ptr(in_xyz), _ptr(out_xyz), C.c_size_t(n), _ptr(index)), self._h)
        return index

    def merl_dirs(self, first, n, like=None, outputs=None):
        proto = like if like is not None else np.empty(0, np.float32)
        i, o = outputs if outputs is not None else (_like(proto, (3, n)), _like(proto, (3, n)))
        _check(lib().bbmcu_merl_dirs(self._h, C.c_uint32(first), C.c_size_t(n), _ptr(i), _ptr(o)), self._h)
        return i, o

    def spherical_dirs(self, grid, first, n, like=None):
        proto = like if like is not None else np.empty(0, np.float32)
        i, o = _like(proto, (3, n)), _like(proto, (3, n))
        _check(lib().bbmcu_spherical_dirs(self._h, C.byref(grid), C.c_uint64(first), C.c_size_t(n), _ptr(i), _ptr(o)), self._h)
        return i, o

    # ---- measured data ---------------------------------------------------------------------------
    def merl_read(self, filename):
        rgb = np.empty((3, MERL_BINS), np.float32)
        _check(lib().bbmcu_merl_read(self._h, filename.encode(), _ptr(rgb)), self._h)
        return rgb

    def merl_write(self, filename, rgb):
        rgb = np.ascontiguousarray(rgb, np.float32)
        if rgb.shape != (3, MERL_BINS):
            raise BbmInvalidArgument("expected a (3, 1458000) table")
        _check(lib().bbmcu_merl_write(self._h, filename.encode(), _ptr(rgb)), self._h)

    def hp_precompute_g1(self):
        """the (100, 1000) Holzschuch-Pacanowski G1 table, recomputed on the GPU (precompute/HolzschuchPacanowski/G1.cpp)"""
        t = np.empty((100, 1000), np.float32)
        _check(lib().bbmcu_hp_precompute_g1(self._h, _ptr(t)), self._h)
        return t

    def hp_precompute_normalization(self):
        """the (100, 100, 100) renormalisation table sigma_rel^2 / sigma_s^2 over (b, c, sin theta_i), recomputed on the GPU
        (precompute/HolzschuchPacanowski/normalization.cpp; the reference ships no copy of it)"""
        t = np.empty((100, 100, 100), np.float32)
        _check(lib().bbmcu_hp_precompute_normalization(self._h, _ptr(t)), self._h)
        return t

    # ---- losses ------------------------------------------------------------------------------------
    def loss(self, metric, reference, grid=None, component=ALL, unit=RADIANCE, first=0, count=0, materialise=False, interleaved=None):
        return Loss(self, metric, reference, grid, component, unit, first, count, materialise, interleaved)


class Loss:
    """a bbm::sampledlossfunction (include/bbm/sampledlossfunction.h:26-95) with one of the six error
    functors of include/loss/*.h.  `reference` is a Bsdf, a (3, 1458000) measured MERL table, or a LIST of such tables
    (a batch of materials evaluated in one launch, see eval_multi).  By default the kernels generate the linearizer's
    directions from the sample index; materialise=True keeps them as planes in device memory instead (same bits).
    The shard this object owns: samples [first, first + count) (count 0 = all), or - interleaved=(rank, world) - every
    world-th block of 1024 samples starting at block `rank`: shards of equal cost (BBMCU_LOSS_SHARD_INTERLEAVED)."""

    def __init__(self, ctx, metric, reference, grid=None, component=ALL, unit=RADIANCE, first=0, count=0, materialise=False, interleaved=None):
        self.ctx = ctx
        self._h = C.c_void_p()
        m = METRICS.index(metric) if isinstance(metric, str) else int(metric)
        self._keep = reference
        flags = 1 if materialise else 0
        if interleaved is not None:
            first, count = int(interleaved[0]), int(interleaved[1])
            flags |= 2
        if isinstance(reference, Bsdf):
            ref_b, tabs, M = reference._h, None, 1
        else:
            tables = list(reference) if isinstance(reference, (list, tuple)) else [reference]
            M = len(tables)
            tabs = (C.c_void_p * M)(*[_ptr(t) for t in tables])
            ref_b = None
        _check(lib().bbmcu_loss_create_ex(ctx._h, m, C.byref(grid) if grid is not None else None, component, unit, ref_b, tabs, C.c_int(M),
                                          C.c_uint64(first), C.c_uint64(count), C.c_uint(flags), C.byref(self._h)), ctx._h)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.bbmcu_loss_free(self._h)
            self._h = None

    def samples(self):
        return int(lib().bbmcu_loss_samples(self._h))

    def __call__(self, bsdf, params=None, grad=False):
        """loss (K,) and, if grad, gradient (K, P) at K parameter vectors (None: the bsdf's current ones)"""
        P = len(bsdf.parameter_values())
        if params is None:
            K, pp = 1, None
        else:
            params = np.ascontiguousarray(np.atleast_2d(params), np.float64)
            K, pp = params.shape[0], params.ctypes.data_as(C.c_void_p)
            if params.shape[1] != P:
                raise BbmInvalidArgument(f"expected {P} parameters per row, got {params.shape[1]}")
        loss = np.empty(K, np.float64)
        g = np.empty((K, P), np.float64) if grad else None
        _check(lib().bbmcu_loss_eval(self._h, bsdf._h, pp, C.c_size_t(K), loss.ctypes.data_as(C.c_void_p),
                                     g.ctypes.data_as(C.c_void_p) if grad else None, None), self.ctx._h)
        return (loss, g) if grad else loss

    def eval_device(self, bsdf, params, device_out):
        """same, results left in a CUDA buffer of K*(1+P) doubles on the context's stream (for NCCL)"""
        params = np.ascontiguousarray(np.atleast_2d(params), np.float64)
        _check(lib().bbmcu_loss_eval(self._h, bsdf._h, params.ctypes.data_as(C.c_void_p), C.c_size_t(params.shape[0]), None, None,
                                     _ptr(device_out)), self.ctx._h)

    # ---- multi-GPU combine over NVLink peer memory (bbmcu_loss_peer_*): after connect_peers every call of this loss is a
    # collective over the shards and returns the sum over them - no NCCL call, bit-identical on every rank
    def peer_init(self, rank, world, max_values):
        """allocate this shard's exchange window for batches of up to max_values = K*(1+P) doubles;
        returns (64-byte cudaIpc handle for other processes, device address for contexts of this process)"""
        h = (C.c_ubyte * 64)()
        w = C.c_void_p()
        _check(lib().bbmcu_loss_peer_init(self._h, C.c_int(rank), C.c_int(world), C.c_size_t(max_values), h, C.byref(w)), self.ctx._h)
        return bytes(h), int(w.value)

    def peer_connect(self, handles):
        """handles: the `world` 64-byte handles in rank order (other processes' shards)"""
        blob = b"".join(handles)
        _check(lib().bbmcu_loss_peer_connect(self._h, C.c_char_p(blob)), self.ctx._h)

    def peer_connect_ptrs(self, windows):
        """windows: the `world` device addresses in rank order (shards living in this process)"""
        arr = (C.c_void_p * len(windows))(*[C.c_void_p(int(w)) for w in windows])
        _check(lib().bbmcu_loss_peer_connect_ptrs(self._h, arr), self.ctx._h)

    def connect_peers(self, max_values, group=None):
        """one process per GPU under torch.distributed: exchange the window handles and connect.  Collective: every rank
        runs the same sequence of gathers whatever fails locally, and all ranks raise if any of them could not connect."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        err, h = None, b""
        try:
            h, _ = self.peer_init(rank, world, max_values)
        except BbmError as e:
            err = e
        handles = [None] * world
        dist.all_gather_object(handles, h, group=group)
        if err is None and all(len(x) == 64 for x in handles):
            try:
                self.peer_connect(handles)
            except BbmError as e:
                err = e
        oks = [None] * world
        dist.all_gather_object(oks, err is None, group=group)
        if not all(oks):
            raise err if err is not None else BbmError("a peer rank could not map the exchange windows")

    def shard_count(self):
        return int(lib().bbmcu_loss_shard_count(self._h))

    def materials(self):
        return int(lib().bbmcu_loss_materials(self._h))

    def set_metric(self, metric):
        """switch the per-sample functor (the tabulated reference is metric independent): free"""
        m = METRICS.index(metric) if isinstance(metric, str) else int(metric)
        _check(lib().bbmcu_loss_set_metric(self._h, m), self.ctx._h)

    def eval_multi(self, bsdf, params, grad=False):
        """params (M, K, P): K parameter vectors for each of the M materials, ONE launch -> loss (M, K) [, gradient (M, K, P)]"""
        P = len(bsdf.parameter_values())
        params = np.ascontiguousarray(params, np.float64)
        M = self.materials()
        if params.ndim != 3 or params.shape[0] != M or params.shape[2] != P:
            raise BbmInvalidArgument(f"expected parameters of shape ({M}, K, {P}), got {params.shape}")
        K = params.shape[1]
        loss = np.empty((M, K), np.float64)
        g = np.empty((M, K, P), np.float64) if grad else None
        _check(lib().bbmcu_loss_eval_multi(self._h, bsdf._h, params.ctypes.data_as(C.c_void_p), C.c_size_t(K), loss.ctypes.data_as(C.c_void_p),
                                           g.ctypes.data_as(C.c_void_p) if grad else None, None), self.ctx._h)
        return (loss, g) if grad else loss

    def eval_multi_device(self, bsdf, params, device_out, grad=True):
        """same, rows [loss, gradient...] left in a CUDA buffer of (M, K, 1+P) doubles on the context's stream; grad=False
        runs the value-only kernel (the gradient columns are then zero)"""
        params = np.ascontiguousarray(params, np.float64)
        _check(lib().bbmcu_loss_eval_multi_ex(self._h, bsdf._h, params.ctypes.data_as(C.c_void_p), C.c_size_t(params.shape[1]), None, None, _ptr(device_out),
                                              C.c_int(1 if grad else 0)), self.ctx._h)

    def terms(self, bsdf, count=None, material=0):
        """per-sample terms l(idx) of this shard (sampledlossfunction::operator()(idx)); `count` is ignored (kept for
        callers of the first version: the buffer is always sized from the shard)"""
        t = np.empty(self.shard_count(), np.float32)
        _check(lib().bbmcu_loss_terms_at(self._h, bsdf._h, C.c_int(material), _ptr(t)), self.ctx._h)
        return t


def import_fit(filename, config="floatRGB"):
    """io::importFIT (include/io/fit.h:34-52): {key: Bsdf}, keys sorted like the reference's std::map.  config as in Bsdf()"""
    L = lib()
    h = C.c_void_p()
    _check(L.bbmcu_fit_import_ex(None, filename.encode(), C.c_int(CONFIGS.index(config)), C.byref(h)))
    try:
        out = {}
        for i in range(L.bbmcu_fit_count(h)):
            b = C.c_void_p()
            _check(L.bbmcu_fit_bsdf(h, i, C.byref(b)))
            out[L.bbmcu_fit_key(h, i).decode()] = Bsdf(None, _handle=b)
        return out
    finally:
        L.bbmcu_fit_free(h)


def export_fit(filename, data, comment=""):
    """io::exportFIT (include/io/fit.h:62-77)"""
    L = lib()
    h = C.c_void_p()
    _check(L.bbmcu_fit_create(C.byref(h)))
    try:
        for k, b in data.items():
            _check(L.bbmcu_fit_add(h, k.encode(), b._h))
        _check(L.bbmcu_fit_export(None, h, filename.encode(), comment.encode()))
    finally:
        L.bbmcu_fit_free(h)
