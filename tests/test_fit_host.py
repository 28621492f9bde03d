"""CPU: the host logic of the batched compass search (bbm_b200/fit.py) against the reference's own compass
(include/optimizer/compass.h) - the loss is the compiled reference (oracle/_ref), so only the search logic is ours."""
import numpy as np

from oracle.refbind import sph_desc


class RefLoss:
    """callable with the signature of bbm_b200.Loss: K parameter rows -> K losses, evaluated by the float reference
    with its own sequential float sum (what its compass sees)"""

    def __init__(self, ref, metric, desc, fitted, truth):
        self.ref, self.metric, self.desc, self.fitted, self.truth = ref, metric, desc, fitted, truth

    def __call__(self, bsdf, params):
        return self.ref.loss_at(self.metric, self.desc, self.fitted, self.truth, params, accumulate_double=False, threads=8)


def test_batched_compass_follows_reference_compass(ref):
    import bbm_b200 as bb
    from bbm_b200.fit import CompassBatched
    fitted, truth = "CookTorrance()", "CookTorrance([0.3, 0.3, 0.3], 0.2, 1.5)"
    desc = sph_desc((13, 8), (5, 6))
    steps = 25
    want, want_p, _ = ref.compass("nganL2", desc, fitted, truth, steps)
    b = bb.Bsdf(fitted)
    opt = CompassBatched(RefLoss(ref, "nganL2", desc, fitted, truth), b, b.parameter_lower_bound(), b.parameter_upper_bound())
    got = np.array([opt.step() for _ in range(steps)])
    # same loss values, same decisions: the only difference is the reference's "p + s - s" float drift of the base point
    assert np.allclose(got, want, rtol=2e-3), (got, want)
    assert np.allclose(opt.param, want_p, rtol=1e-5, atol=1e-6), (opt.param, want_p)


def test_probe_order_and_box():
    import bbm_b200 as bb
    from bbm_b200.fit import CompassBatched
    b = bb.Bsdf("CookTorrance()")
    calls = []

    def loss(bsdf, params):
        calls.append(np.array(params))
        return np.full(len(params), 1.0)
    opt = CompassBatched(loss, b, b.parameter_lower_bound(), b.parameter_upper_bound())
    pr, ok = opt.probes()
    P = 5
    assert pr.shape == (2 * P, P)
    # +1, -1, +2, -2, ...; albedo 0.5 +- 1 and roughness 0.1 +- 1 leave the box, eta 1.3 + 1 stays inside [1, 5]
    assert list(ok) == [False] * 8 + [True, False]
    assert pr[8, 4] == np.float32(1.3) + np.float32(1.0)
    assert opt.step() == 1.0 and opt.step_size == np.float32(0.5)          # nothing better: contraction
