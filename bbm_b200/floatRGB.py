"""bbm_b200.floatRGB - the module surface of the reference's Python binding (`bbm_floatRGB`, built by src/python/bbm_python.cpp
from include/python/py_core.h:44-190 and include/python/py_bsdf.h:57-76) on the CUDA backbone:

    import bbm_b200.floatRGB as bbm
    b = bbm.Aggregate(bbm.Lambertian([0.2, 0.1, 0.05]), bbm.CookTorrance([0.3, 0.3, 0.3], roughness = 0.2, eta = 1.5))
    b.eval([0.3, 0.2, 0.93], [0.5, -0.1, 0.86]);  s = b.sample([0.5, -0.1, 0.86], [0.3, 0.6]);  str(b)
    bbm.parameter_values(b)[6] = 0.25

so that the reference's own `fits/import_fits.py` - which evaluates every line of a .fit file as Python against the module's
names - runs UNCHANGED:  `import_fits("ngan_ward.fit", bbm_b200.floatRGB)`.

One factory per model (the 34 analytic models + Merl) taking the model's constructor arguments positionally or by name
(include/core/args.h; a scalar where an RGB is expected is broadcast, core/stringconvert.h:333-360), `Aggregate(*bsdfs)`,
`BsdfPtr` with eval / sample / pdf / reflectance / __str__, `BsdfSample`, the enums bsdf_flag / unit_t / bsdf_attr, and the four
parameter enumerations.  Arguments are rendered into the BSDF string grammar and parsed by the library (bbmcu_bsdf_from_string),
so defaults, names, broadcasting and error messages are the C ABI's.  Single-direction calls are n = 1 launches on a lazily
created context of device 0 - convenient, ~20 us each; batches go through bbm_b200.Context (numpy / torch arrays, (3, n))."""
from . import _surface

_surface.populate(globals(), "floatRGB")
