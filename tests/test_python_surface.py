"""The reference's Python module surface (include/python/py_core.h:44-190, py_bsdf.h:57-76) on the CUDA backbone."""
import numpy as np
import pytest


def test_factories_aggregate_and_parameter_views_on_cpu():
    """object construction, strings and parameter enumeration need no device"""
    import bbm_b200 as bb
    import bbm_b200.floatRGB as bbm
    for name in bb.model_names():
        if name != "Merl":
            assert str(getattr(bbm, name)()) == bb.Bsdf(name + "()").to_string()
    b = bbm.Aggregate(bbm.Lambertian([0.2, 0.1, 0.05]), bbm.CookTorrance([0.3, 0.3, 0.3], eta=1.5, roughness=0.2))
    assert str(b) == "Aggregate(Lambertian(albedo = [0.2, 0.1, 0.05]), CookTorrance(albedo = [0.3, 0.3, 0.3], roughness = 0.2, eta = 1.5))"
    assert str(bbm.Lambertian(0.25)) == "Lambertian(albedo = [0.25, 0.25, 0.25])"            # scalar broadcast
    assert "eta = [[1.5, 1.4, 1.3], [3, 2, 1]]" in str(bbm.He(0.2, 3, [[1.5, 1.4, 1.3], [3, 2, 1]]))
    p = bbm.parameter_values(b)
    assert len(p) == 8 and p[6] == pytest.approx(0.2)
    p[6] = 0.25                                                                             # a LIVE view (bbm::vector<Value&>)
    assert "roughness = 0.25" in str(b)
    p *= 2.0
    assert "roughness = 0.5" in str(b) and "albedo = [0.4, 0.2, 0.1]" in str(b)
    assert list(bbm.parameter_upper_bound(b)) == [1, 1, 1, 1, 1, 1, 1, 5]
    assert len(bbm.parameter_values(b, bbm.bsdf_attr.DiffuseScale)) == 3
    assert str(bbm.BsdfPtr(b)) == str(b)
    assert bbm.bsdf_flag.All == 3 and getattr(bbm.bsdf_flag, "None") == 0 and bbm.unit_t.Importance == 1
    with pytest.raises(bb.BbmInvalidArgument):
        bbm.CookTorrance(nonsense=1.0)
    with pytest.raises(TypeError):
        bbm.Aggregate("Lambertian()")


@pytest.mark.gpu
def test_bsdfptr_methods_against_reference(ref):
    import bbm_b200.floatRGB as bbm
    s = "Aggregate(Lambertian([0.2, 0.1, 0.05]), CookTorrance([0.3, 0.3, 0.3], 0.2, 1.5))"
    b = bbm.Aggregate(bbm.Lambertian([0.2, 0.1, 0.05]), bbm.CookTorrance([0.3, 0.3, 0.3], 0.2, 1.5))
    i = np.array([[0.3, 0.2, 0.9327379]], np.float32)
    o = np.array([[0.5, -0.1, 0.8602325]], np.float32)
    assert np.allclose(b.eval(i[0], o[0]), ref.eval(s, i, o)[0], rtol=1e-5)
    assert np.allclose(b.eval(i[0], o[0], bbm.bsdf_flag.Diffuse), ref.eval(s, i, o, component=1)[0], rtol=1e-5)
    assert b.pdf(i[0], o[0]) == pytest.approx(float(ref.pdf(s, i, o)[0]), rel=1e-5)
    d, p, f = ref.sample(s, o, np.array([[0.3, 0.6]], np.float32))
    smp = b.sample(o[0], [0.3, 0.6])
    assert np.allclose(smp.direction, d[0], rtol=1e-5, atol=1e-5) and int(smp.flag) == int(f[0]) and "flag = " in str(smp)
    assert np.allclose(b.reflectance(o[0]), ref.reflectance(s, o)[0], rtol=1e-5)
    assert b.pdf(i[0], o[0], mask=False) == 0.0
