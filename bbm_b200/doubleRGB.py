"""bbm_b200.doubleRGB - the same module surface as bbm_b200.floatRGB (the reference's `bbm_doubleRGB`), with the HOST side in the
reference's doubleRGB configuration: constructor arguments and .fit values are parsed with std::stod and kept as doubles in
the parameter vectors and strings (fits/bagher_sgd.fit, whose c = 1.3e49 overflows std::stof, imports here as it does in the
reference's doubleRGB build).  The kernels compute in FP32 either way."""
from . import _surface

_surface.populate(globals(), "doubleRGB")
