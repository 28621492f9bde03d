"""GPU tests of the round-2 paths: fused-linearizer eval / loss kernels, the multi-material batched loss, optional outputs,
device-generated inputs, plane strides, the pageable-host ring, measured lobes inside a fitted BSDF, and the analytic
gradient against central differences of the doubleRGB reference (oracle/_ref)."""
import os

import numpy as np
import pytest

from tests.util import assert_parity, soa

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRUTH = "Aggregate(Lambertian([0.2, 0.1, 0.05]), CookTorrance([0.3, 0.3, 0.3], 0.2, 1.5))"


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


# ---- fused linearizer ----------------------------------------------------------------------------------------------------------
def test_eval_merl_grid_fused_is_bit_identical_to_materialised(ctx, ref):
    """EvalGridOp generates merl_linearizer(idx) in registers: same directions, same values as eval over bbmcu_merl_dirs, bit
    for bit (the SURVEY section 7 protocol, with 0 ulp instead of 2), and within 1e-5 of the reference at those directions"""
    import bbm_b200 as bb
    N = bb.MERL_BINS
    i, o = ctx.merl_dirs(0, N)
    for s in ("CookTorrance()", "GGX([0.5, 0.4, 0.3], 0.05, 1.6)", "Aggregate(Lambertian([0.2, 0.1, 0.05]), LowCookTorrance([0.3, 0.3, 0.3], 0.01, 1.5))", "He()",
              "Aggregate(Lambertian(), Phong(), Ward())"):
        b = bb.Bsdf(s)
        rgb, gi, go = ctx.eval_merl_grid(b, dirs=True)
        assert np.array_equal(bits(gi), bits(i)) and np.array_equal(bits(go), bits(o)), s
        want = ctx.eval(b, i, o)
        assert np.array_equal(bits(rgb), bits(want)), s
    # a shard that starts in the middle, odd length, without the direction outputs
    b = bb.Bsdf("CookTorrance()")
    part = ctx.eval_merl_grid(b, first=700001, n=4099)
    assert np.array_equal(bits(part), bits(ctx.eval(b, np.ascontiguousarray(i[:, 700001:704100]), np.ascontiguousarray(o[:, 700001:704100]))))
    sub = slice(123456, 123456 + 50000)
    assert_parity(ctx.eval_merl_grid(b)[:, sub].T, ref.eval("CookTorrance()", i.T[sub].copy(), o.T[sub].copy(), threads=8), 1e-5, what="fused grid eval vs reference")
    with pytest.raises(bb.BbmError):
        ctx.eval_merl_grid(b, first=N - 10, n=11)


@pytest.mark.parametrize("grid_kind", ["merl", "spherical"])
def test_loss_fused_linearizer_equals_materialised(ctx, grid_kind):
    """the default loss objects keep no direction planes (12 B per sample per pass); results are bit-identical to the
    materialised mode for the static single / pair kernels and the run-time lobe list, with and without gradient, sharded"""
    import bbm_b200 as bb
    truth = bb.Bsdf(TRUTH)
    grid = None if grid_kind == "merl" else bb.spherical_grid((31, 16), (5, 9))
    N = bb.MERL_BINS if grid is None else grid.size()
    rng = np.random.default_rng(5)
    for fitted, metric in (("Aggregate(Lambertian(), CookTorrance())", "nganL2"), ("GGX()", "bieronLog"), ("Aggregate(Lambertian(), Phong(), GGX())", "lowL2"),
                           ("Aggregate(Lambertian(), NganHe())", "standardLog")):
        fb = bb.Bsdf(fitted)
        p0 = fb.parameter_values()
        params = p0[None] * (1 + 0.05 * rng.random((5, len(p0))))
        for first, count in ((0, 0), (N // 3 + 1, N // 5 + 3)):
            Lf = ctx.loss(metric, truth, grid, first=first, count=count)
            Lm = ctx.loss(metric, truth, grid, first=first, count=count, materialise=True)
            lf, gf = Lf(fb, params, grad=True)
            lm, gm = Lm(fb, params, grad=True)
            assert np.array_equal(lf.view(np.uint64), lm.view(np.uint64)) and np.array_equal(gf.view(np.uint64), gm.view(np.uint64)), (fitted, metric, first)
            assert np.array_equal(Lf(fb, params).view(np.uint64), Lm(fb, params).view(np.uint64))
            assert np.array_equal(bits(Lf.terms(fb)), bits(Lm.terms(fb)))
            assert Lf.shard_count() == (count or N) and len(Lf.terms(fb)) == Lf.shard_count()


# ---- multi-material batched loss ---------------------------------------------------------------------------------------------
def _synthetic_tables(ctx, strings):
    import bbm_b200 as bb
    return [np.ascontiguousarray(ctx.eval_merl_grid(bb.Bsdf(s))) for s in strings]


def test_loss_multi_material_equals_single_material_losses(ctx):
    """M measured tables x K parameter sets in ONE launch (block z = material) return the bits of M separate losses"""
    import bbm_b200 as bb
    mats = ["Aggregate(Lambertian([0.2, 0.1, 0.05]), CookTorrance([0.3, 0.3, 0.3], 0.2, 1.5))",
            "Aggregate(Lambertian([0.05, 0.1, 0.2]), CookTorrance([0.1, 0.2, 0.3], 0.05, 1.3))",
            "Aggregate(Lambertian([0.3, 0.3, 0.3]), GGX([0.2, 0.2, 0.2], 0.3, 1.8))"]
    tables = _synthetic_tables(ctx, mats)
    rng = np.random.default_rng(9)
    for fitted in ("Aggregate(Lambertian(), CookTorrance())", "Aggregate(Lambertian(), Phong(), GGX())"):
        fb = bb.Bsdf(fitted)
        p0 = fb.parameter_values()
        K = 6
        params = p0[None, None] * (1 + 0.1 * rng.random((len(mats), K, len(p0))))
        Lm = ctx.loss("nganL2", tables, None, first=1000, count=300001)
        assert Lm.materials() == 3
        loss, grad = Lm.eval_multi(fb, params, grad=True)
        plain = Lm.eval_multi(fb, params)
        for m, t in enumerate(tables):
            L1 = ctx.loss("nganL2", t, None, first=1000, count=300001)
            l1, g1 = L1(fb, params[m], grad=True)
            assert np.array_equal(loss[m].view(np.uint64), l1.view(np.uint64)) and np.array_equal(grad[m].view(np.uint64), g1.view(np.uint64)), (fitted, m)
            assert np.array_equal(plain[m].view(np.uint64), L1(fb, params[m]).view(np.uint64))
            assert np.array_equal(bits(Lm.terms(fb, material=m)), bits(L1.terms(fb)))
        # the metric is a per-call switch: the tabulated reference does not depend on it
        Lm.set_metric("standardLog")
        L1 = ctx.loss("standardLog", tables[1], None, first=1000, count=300001)
        assert np.array_equal(Lm.eval_multi(fb, params)[1].view(np.uint64), L1(fb, params[1]).view(np.uint64))
    with pytest.raises(bb.BbmError):
        Lm(fb, params[0])                                  # single-material call on a batched loss


def test_batched_compass_over_materials_matches_per_material_fits(ctx):
    """fit.CompassMulti advances one compass search per material with one launch per step; each trajectory equals the
    single-material CompassBatched trajectory (same probes, same losses)"""
    import bbm_b200 as bb
    from bbm_b200 import fit
    mats = ["Aggregate(Lambertian([0.2, 0.1, 0.05]), CookTorrance([0.3, 0.3, 0.3], 0.2, 1.5))",
            "Aggregate(Lambertian([0.05, 0.1, 0.2]), CookTorrance([0.1, 0.2, 0.3], 0.1, 1.3))"]
    tables = _synthetic_tables(ctx, mats)
    grid = bb.spherical_grid((31, 16), (1, 9))
    fitted = "Aggregate(Lambertian(), CookTorrance())"
    Lm = ctx.loss("nganL2", tables, grid)
    b = bb.Bsdf(fitted)
    multi = fit.CompassMulti(Lm, b, b.parameter_lower_bound(), b.parameter_upper_bound())
    traces = [multi.step().copy() for _ in range(12)]
    for m, t in enumerate(tables):
        b1 = bb.Bsdf(fitted)
        single = fit.CompassBatched(ctx.loss("nganL2", t, grid), b1, b1.parameter_lower_bound(), b1.parameter_upper_bound())
        for k in range(12):
            assert single.step() == pytest.approx(float(traces[k][m]), rel=1e-6, abs=1e-12), (m, k)
        assert np.allclose(single.param, multi.param[m], rtol=1e-6)


# ---- optional outputs, generated inputs, strides, pageable memory ----------------------------------------------------------------
def test_optional_outputs_and_generated_inputs(ctx, hostsim, ref):
    import ctypes as C
    import torch
    import bbm_b200 as bb
    n = 300001
    b = bb.Bsdf("GGX([0.5, 0.4, 0.3], 0.2, 1.5)")
    dev = torch.device("cuda", 0)
    like = torch.empty(0, device=dev)
    full, (go, gx) = ctx.sample_eval_pdf_generated(b, seed=1234, first=77, n=n, like=like, inputs=True)
    ctx.synchronize()
    # (1) the generator is a pure function of (seed, index): the host-compiled copy of the same device function agrees bit for bit
    ho, hx = np.empty((3, n), np.float32), np.empty((2, n), np.float32)
    hostsim.lib.hostsim_generate_inputs(C.c_uint64(1234), C.c_uint64(77), C.c_size_t(n), ho.ctypes.data_as(C.c_void_p), hx.ctypes.data_as(C.c_void_p))
    assert np.array_equal(bits(go.cpu().numpy()), bits(ho)) and np.array_equal(bits(gx.cpu().numpy()), bits(hx))
    assert 0.0 <= float(hx.min()) and float(hx.max()) < 1.0 and abs(float(np.linalg.norm(ho, axis=0).mean()) - 1.0) < 1e-6
    assert abs(float(ho[2].mean()) - 0.5) < 5e-3                      # uniform on the hemisphere: E[cos theta] = 1/2
    # (2) same outputs as the pass that READS those inputs
    want = ctx.sample_eval_pdf(b, go, gx)
    ctx.synchronize()
    for a, w in zip(full, want):
        assert np.array_equal(bits(a.cpu().numpy()), bits(w.cpu().numpy()))
    # (3) outputs that are not wanted are not written; the others are unchanged (device and host paths)
    sub = ctx.sample_eval_pdf(b, go, gx, want=("rgb", "pdf"))
    assert sub[0] is None and sub[1] is None and sub[2] is None
    assert np.array_equal(bits(sub[3].cpu().numpy()), bits(want[3].cpu().numpy())) and np.array_equal(bits(sub[4].cpu().numpy()), bits(want[4].cpu().numpy()))
    hsub = ctx.sample_eval_pdf(b, ho, hx, want=("dir", "pdf"))
    assert hsub[3] is None and np.array_equal(bits(hsub[0]), bits(want[0].cpu().numpy())) and np.array_equal(bits(hsub[4]), bits(want[4].cpu().numpy()))
    gsub = ctx.sample_eval_pdf_generated(b, seed=1234, first=77, n=n, want=("rgb",))       # host outputs, chunked: offsets carried across chunks
    assert np.array_equal(bits(gsub[3]), bits(want[3].cpu().numpy()))
    big = ctx.sample_eval_pdf_generated(b, seed=5, first=0, n=(1 << 21) + 5, want=("pdf",))
    tail = ctx.sample_eval_pdf_generated(b, seed=5, first=1 << 21, n=5, want=("pdf",))
    assert np.array_equal(bits(big[4][-5:]), bits(tail[4]))
    # (4) and the reference agrees at the generated inputs
    k = 100000
    rd, rp, rf = ref.sample("GGX([0.5, 0.4, 0.3], 0.2, 1.5)", ho.T[:k].copy(), hx.T[:k].copy(), threads=8)
    assert_parity(full[0].cpu().numpy().T[:k], rd, 1e-5, floor=1e-5, what="generated-input sample direction vs reference")
    with pytest.raises(bb.BbmError):
        ctx.sample_eval_pdf(b, ho, hx, outputs=(None, None, None, None, None))


def test_plane_stride_keeps_vector_access_for_any_n(ctx):
    """n = 2^20 + 1: contiguous planes are misaligned (scalar path); column slices of padded buffers run vectorised - same bits"""
    import torch
    import bbm_b200 as bb
    n, ld = (1 << 20) + 1, (1 << 20) + 4
    rng = np.random.default_rng(2)
    out = rng.standard_normal((3, n)).astype(np.float32); out[2] = np.abs(out[2]); out /= np.linalg.norm(out, axis=0)
    xi = rng.random((2, n), dtype=np.float32)
    b = bb.Bsdf("Aggregate(Lambertian([0.2, 0.1, 0.05]), GGX([0.3, 0.3, 0.3], 0.2, 1.5))")
    want = ctx.sample_eval_pdf(b, torch.from_numpy(out).cuda(), torch.from_numpy(xi).cuda())
    po, px = torch.zeros((3, ld), device="cuda"), torch.zeros((2, ld), device="cuda")
    po[:, :n] = torch.from_numpy(out).cuda(); px[:, :n] = torch.from_numpy(xi).cuda()
    outs = (torch.zeros((3, ld), device="cuda")[:, :n], torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda", dtype=torch.int32),
            torch.zeros((3, ld), device="cuda")[:, :n], torch.zeros(n, device="cuda"))
    got = ctx.sample_eval_pdf(b, po[:, :n], px[:, :n], outputs=outs)
    ctx.synchronize()
    for a, w in zip(got, want):
        assert np.array_equal(bits(a.cpu().numpy()), bits(w.cpu().numpy()))
    # host arrays with a stride take the same route (2-D copies with the caller's pitch)
    ho = np.zeros((3, ld), np.float32); ho[:, :n] = out
    hx = np.zeros((2, ld), np.float32); hx[:, :n] = xi
    hres = ctx.sample_eval_pdf(b, ho[:, :n], hx[:, :n], want=("rgb", "pdf"),
                               outputs=(None, None, None, np.zeros((3, ld), np.float32)[:, :n], np.zeros(n, np.float32)))
    assert np.array_equal(bits(hres[3]), bits(want[3].cpu().numpy())) and np.array_equal(bits(hres[4]), bits(want[4].cpu().numpy()))


def test_pageable_and_pinned_host_memory_agree(ctx):
    """pageable numpy arrays go through the internal pinned ring, registered ones are DMA'd in place: same results"""
    import bbm_b200 as bb
    n = (1 << 22) + 12345                                   # three chunks, ragged tail
    rng = np.random.default_rng(4)
    out = rng.standard_normal((3, n)).astype(np.float32); out[2] = np.abs(out[2]); out /= np.linalg.norm(out, axis=0)
    xi = rng.random((2, n), dtype=np.float32)
    b = bb.Bsdf("GGX()")
    pageable = ctx.sample_eval_pdf(b, out, xi)
    ctx.host_register(out); ctx.host_register(xi)
    try:
        pinned = ctx.sample_eval_pdf(b, out, xi)
    finally:
        ctx.host_unregister(out); ctx.host_unregister(xi)
    for a, w in zip(pageable, pinned):
        assert np.array_equal(bits(a), bits(w))


# ---- measured lobes inside a fitted BSDF (attribute blocks must carry the table's device address) ----------------------
def test_loss_of_bsdf_with_measured_lobe(ctx, tmp_path):
    import bbm_b200 as bb
    table = ctx.eval_merl_grid(bb.Bsdf(TRUTH))
    path = str(tmp_path / "m.binary")
    ctx.merl_write(path, table)
    grid = bb.spherical_grid((13, 8), (5, 6))
    L = ctx.loss("nganL2", bb.Bsdf("Aggregate(Lambertian([0.1, 0.1, 0.1]), CookTorrance())"), grid)
    # a fitted BSDF holding a Merl(...) lobe: run-time lobe list; and the measured model alone (P = 0: static kernel)
    agg = bb.Bsdf(f'Aggregate(Lambertian([0.2, 0.2, 0.2]), Merl("{path}"))')
    l, g = L(agg, grad=True)
    terms = L.terms(agg)
    assert np.isfinite(l[0]) and l[0] > 0 and abs(l[0] - terms.astype(np.float64).sum() / L.samples()) <= 1e-4 * l[0]
    l2 = L(agg, agg.parameter_values()[None] * np.array([[1.0, 1.0, 1.0], [0.5, 0.5, 0.5]]))
    assert l2[0] == l[0] and l2[1] != l[0]
    alone = bb.Bsdf(f'Merl("{path}")')
    la = L(alone)
    ta = L.terms(alone)
    assert np.isfinite(la[0]) and abs(la[0] - ta.astype(np.float64).sum() / L.samples()) <= 1e-4 * abs(la[0])
    ctx.synchronize()                                       # the context survived (no illegal address)
    assert np.isfinite(L(bb.Bsdf("Lambertian()"))[0])


# ---- gradient on the GPU against the doubleRGB reference, one case per model family ------------------------------------------------
GRAD_FAMILIES = ("Lambertian(", "Phong(", "NganLafortune(", "Ward(", "NganWardDuer(", "AshikhminShirleyFull(", "CookTorrance(", "GGXHeitz(", "LowMicrofacet(", "LowSmooth(",
                 "Ribardiere(", "Bagher(", "EPD(", "HeHolzschuch(")


@pytest.mark.parametrize("metric", ["nganL2", "standardLog"])
def test_gpu_gradient_vs_finite_differences_of_double_reference(ctx, ref, refd, metric):
    """the device loss kernels (not the host-compiled copy) against central differences of the reference compiled in
    doubleRGB - the gradient oracle of SURVEY.md section 8(c); same protocol as tests/test_gradients_hostsim.py"""
    import bbm_b200 as bb
    from tests.test_gradients_hostsim import CASES, check_gradient_case
    hp = float(np.float32(2) * np.float32(np.pi))
    grid = bb.spherical_grid((11, 6), (4, 5), start_in=(0, 0.05), start_out=(0, 0.05), end_in=(hp, 1.4), end_out=(hp, 1.4))
    done = 0
    for fitted, truth in CASES:
        if not fitted.startswith(GRAD_FAMILIES):
            continue
        L = ctx.loss(metric, bb.Bsdf(truth), grid)
        loss, grad = L(bb.Bsdf(fitted), grad=True)
        check_gradient_case(ref, refd, metric, fitted, truth, float(loss[0]), grad[0])
        done += 1
    assert done == len(GRAD_FAMILIES)


@pytest.mark.gpu
@pytest.mark.parametrize("materialise", [False, True])
def test_interleaved_loss_shards_sum_to_the_whole(ctx, materialise):
    """BBMCU_LOSS_SHARD_INTERLEAVED: blocks of 1024 samples dealt round-robin to the shards.  The shards' sums add up to the
    unsharded loss and gradient, their per-sample terms are the unsharded terms at bbm_b200.shard.interleaved_indices, and
    the fused and materialised modes agree bit for bit."""
    import bbm_b200 as bb
    from bbm_b200.shard import interleaved_indices
    truth = bb.Bsdf("Aggregate(Lambertian([0.2, 0.1, 0.05]), CookTorrance([0.3, 0.3, 0.3], 0.2, 1.5))")
    fitted = bb.Bsdf("Aggregate(Lambertian(), CookTorrance())")
    rng = np.random.default_rng(3)
    params = fitted.parameter_values()[None] * (1 + 0.1 * rng.random((5, 8)))
    whole = ctx.loss("nganL2", truth, None, materialise=materialise)
    lw, gw = whole(fitted, params, grad=True)
    tw = whole.terms(fitted)
    for world in (3, 8):
        ls, gs, total = 0.0, 0.0, 0
        for r in range(world):
            L = ctx.loss("nganL2", truth, None, materialise=materialise, interleaved=(r, world))
            idx = interleaved_indices(bb.MERL_BINS, r, world)
            assert L.shard_count() == len(idx)
            l, g = L(fitted, params, grad=True)
            ls, gs, total = ls + l, gs + g, total + len(idx)
            assert np.array_equal(L.terms(fitted).view(np.uint32), tw[idx].view(np.uint32))
        assert total == bb.MERL_BINS
        assert np.allclose(ls, lw, rtol=1e-13, atol=0) and np.allclose(gs, gw, rtol=1e-11, atol=1e-14)
    with pytest.raises(bb.BbmError):
        ctx.loss("nganL2", truth, None, interleaved=(3, 3))


@pytest.mark.gpu
@pytest.mark.parametrize("s", ["GGX()", "GGX([0.3, 0.6, 0.9], 0.15, 1.7)", "Phong()", "CookTorrance()", "Lambertian()", "Aggregate(Lambertian(), GGX())", "Aggregate(Lambertian(), Phong(), GGX())", "He()"])
def test_host_path_with_narrowed_and_derived_outputs_equals_device_path(ctx, s):
    """host pointers: the flag plane travels as one byte per element and - for models whose sample.pdf is pdf(sample.direction,
    out) - sample.pdf does not travel at all (host threads write both planes); the eval of the hand-merged GGX lobe travels
    as one plane before its RGB scale and host threads form the three products.  Every output equals the device-pointer
    call's, bit for bit, for pinned and for pageable caller memory, also when some outputs are not asked for."""
    import torch
    import bbm_b200 as bb
    n = (1 << 22) + 4097                                     # three chunks, the last one ragged and not a multiple of 4
    rng = np.random.default_rng(11)
    z = rng.random(n, dtype=np.float32); ph = rng.random(n, dtype=np.float32) * np.float32(2 * np.pi)
    out = np.stack([np.sqrt(1 - z * z) * np.cos(ph), np.sqrt(1 - z * z) * np.sin(ph), z]).astype(np.float32)
    xi = rng.random((2, n), dtype=np.float32)
    xi[0, :5] = [-0.5, 1.5, 0.0, 1.0, 0.5]                   # invalid random numbers: flag None, sample.pdf 0
    b = bb.Bsdf(s)
    dev = torch.device("cuda", 0)
    res = ctx.sample_eval_pdf(b, torch.from_numpy(out).to(dev), torch.from_numpy(xi).to(dev))
    ctx.synchronize()                                        # the library's own stream: finish before torch reads the tensors
    want = [t.cpu().numpy() for t in res]
    got = ctx.sample_eval_pdf(b, out, xi)                    # pageable numpy memory
    for name, g, w in zip(("dir", "sample_pdf", "flag", "rgb", "pdf"), got, want):
        assert np.array_equal(g.view(np.uint32), w.view(np.uint32)), (s, name, "pageable")
    pin = lambda a: torch.from_numpy(a).pin_memory()        # noqa: E731
    outs = (torch.empty((3, n)).pin_memory(), torch.empty(n).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty((3, n)).pin_memory(), torch.empty(n).pin_memory())
    ctx.sample_eval_pdf(b, pin(out), pin(xi), outputs=outs)
    for name, g, w in zip(("dir", "sample_pdf", "flag", "rgb", "pdf"), outs, want):
        assert np.array_equal(g.numpy().view(np.uint32), w.view(np.uint32)), (s, name, "pinned")
    # sample.pdf without pdf (nothing to derive it from) and without the flag
    d2, sp2, f2, rgb2, p2 = ctx.sample_eval_pdf(b, out, xi, want=("sample_pdf",))
    assert np.array_equal(sp2.view(np.uint32), want[1].view(np.uint32)) and d2 is None and p2 is None
    d3, sp3, f3, rgb3, p3 = ctx.sample_eval_pdf(b, out, xi, want=("sample_pdf", "pdf"))
    assert np.array_equal(sp3.view(np.uint32), want[1].view(np.uint32)) and np.array_equal(p3.view(np.uint32), want[4].view(np.uint32))


# ---- compact pair loss kernel (bbmcu_losscompact.cuh) -----------------------------------------------------------------------------
@pytest.mark.parametrize("metric", ["nganL2", "lowL2", "bieronL2", "lowLog", "bieronLog", "standardLog"])
def test_compact_loss_equals_generic_tile_kernel(ctx, metric):
    """Aggregate(Lambertian, Cook-Torrance family / GGX / LowMicrofacet / isotropic Ashikhmin-Shirley / Phong / NganLafortune) runs the compact kernel (per-sample invariants, per-set constants, closed-form
    jacobian); BBMCU_LOSS_NO_COMPACT=1 keeps the generic dual-number tile kernel.  Both on the full MERL grid, a shard of it
    and a spherical grid, with and without gradient, K = 1 / 7 / 40 parameter sets: equal to float rounding of the terms."""
    import bbm_b200 as bb
    truth = bb.Bsdf(TRUTH)
    rng = np.random.default_rng(11)
    for fitted in ("Aggregate(Lambertian([0.3, 0.2, 0.1]), CookTorrance([0.4, 0.5, 0.6], 0.1, 1.6))",
                   "Aggregate(Lambertian(), LowCookTorrance())", "Aggregate(Lambertian([0.3, 0.2, 0.1]), NganCookTorrance([0.4, 0.5, 0.6], 0.1, 0.2))",
                   "Aggregate(Lambertian([0.3, 0.2, 0.1]), GGX([0.4, 0.5, 0.6], 0.1, 1.6))",
                   "Aggregate(Lambertian([0.1, 0.1, 0.3]), LowMicrofacetFit([51.7, 37.9, 27.4], 10482.1, 0.8167, 2.2365))",
                   "Aggregate(Lambertian([0.3, 0.2, 0.1]), NganAshikhminShirley([0.4, 0.5, 0.6], 0.1, 80.0))",
                   "Aggregate(Lambertian([0.3, 0.2, 0.1]), LowAshikhminShirley([0.4, 0.5, 0.6], 1.6, 2000.0))",
                   "Aggregate(Lambertian([0.3, 0.2, 0.1]), NganBlinnPhong([0.4, 0.5, 0.6], 60.0))",
                   "Aggregate(Lambertian([0.3, 0.2, 0.1]), NganLafortune([0.4, 0.5, 0.6], -0.58, 0.57, 40.0))",
                   "Aggregate(Lambertian([0.3, 0.2, 0.1]), LowSmooth([40.0, 50.0, 60.0], 3000.0, 1.2, 1.6))",
                   "CookTorrance([0.4, 0.5, 0.6], 0.1, 1.6)", "GGX([0.4, 0.5, 0.6], 0.1, 1.6)", "NganLafortune([0.4, 0.5, 0.6], -0.58, 0.57, 40.0)"):      # the lobe alone
        fb = bb.Bsdf(fitted)
        p0 = fb.parameter_values()
        for grid, first, count in ((None, 0, 0), (None, 400_001, 300_007), (bb.spherical_grid((31, 16), (5, 9)), 0, 0)):
            L = ctx.loss(metric, truth, grid, first=first, count=count)
            for K in (1, 7, 40):
                params = p0[None] * (1 + 0.05 * rng.random((K, len(p0))))
                lc, gc = L(fb, params, grad=True)
                vc = L(fb, params)
                os.environ["BBMCU_LOSS_NO_COMPACT"] = "1"
                try:
                    lg, gg = L(fb, params, grad=True)
                    vg = L(fb, params)
                finally:
                    del os.environ["BBMCU_LOSS_NO_COMPACT"]
                assert np.all(np.abs(lc - lg) <= 2e-5 * np.abs(lg)), (fitted, metric, K, lc, lg)
                assert np.all(np.abs(vc - lg) <= 2e-5 * np.abs(lg)) and np.all(np.abs(vg - lg) <= 2e-5 * np.abs(lg))
                tol = 1e-4 * np.abs(gg) + 1e-5 * np.abs(gg).max(axis=1, keepdims=True)      # (components that nearly cancel over 1.4 M float terms)
                assert np.all(np.abs(gc - gg) <= tol), (fitted, metric, K, np.abs(gc - gg).max(), gc[0], gg[0])
    # the isotropic Ward lobes: Inf / NaN at the horizon like the reference (non-finite loss over the MERL grid in both kernels);
    # compared on a grid that stays off it
    hp = float(np.float32(2) * np.float32(np.pi))
    off_horizon = bb.spherical_grid((31, 12), (6, 9), start_in=(0, 0.05), start_out=(0, 0.05), end_in=(hp, 1.4), end_out=(hp, 1.4))
    for fitted in ("Aggregate(Lambertian([0.3, 0.2, 0.1]), NganWard([0.4, 0.5, 0.6], 0.2))", "Aggregate(Lambertian([0.3, 0.2, 0.1]), NganWardDuer([0.4, 0.5, 0.6], 0.05))"):
        fb = bb.Bsdf(fitted)
        p0 = fb.parameter_values()
        params = p0[None] * (1 + 0.05 * rng.random((9, len(p0))))
        assert not np.isfinite(ctx.loss(metric, truth, None)(fb, params)).any()
        L = ctx.loss(metric, truth, off_horizon)
        lc, gc = L(fb, params, grad=True)
        os.environ["BBMCU_LOSS_NO_COMPACT"] = "1"
        try:
            lg, gg = L(fb, params, grad=True)
            assert not np.isfinite(ctx.loss(metric, truth, None)(fb, params)).any()
        finally:
            del os.environ["BBMCU_LOSS_NO_COMPACT"]
        assert np.all(np.abs(lc - lg) <= 2e-5 * np.abs(lg)), (fitted, metric, lc, lg)
        assert np.all(np.abs(gc - gg) <= 1e-4 * np.abs(gg) + 2e-6 * np.abs(gg).max(axis=1, keepdims=True)), (fitted, metric)
    # a parameter set's result does not depend on what else shares the launch (block partials per (set, tile), fixed order)
    L = ctx.loss(metric, truth, None)
    fb = bb.Bsdf("Aggregate(Lambertian(), CookTorrance())")
    p0 = fb.parameter_values()
    params = p0[None] * (1 + 0.05 * rng.random((33, len(p0))))
    la, ga = L(fb, params, grad=True)
    lb, gb = L(fb, params[5:6], grad=True)
    assert np.array_equal(la[5:6].view(np.uint64), lb.view(np.uint64)) and np.array_equal(ga[5:6].view(np.uint64), gb.view(np.uint64))
    # several materials per block (the direction-only part of a tile serves all of them): against the generic kernel
    if metric in ("nganL2", "bieronLog"):
        mats = ["Aggregate(Lambertian([%g, 0.1, 0.05]), CookTorrance([0.3, 0.3, %g], %g, 1.5))" % (0.1 + 0.05 * m, 0.1 + 0.1 * m, 0.05 + 0.04 * m) for m in range(5)]
        tables = _synthetic_tables(ctx, mats)
        Lm = ctx.loss(metric, tables, None, first=200_000, count=150_001)
        for K in (1, 3):
            pm = p0[None, None] * (1 + 0.05 * rng.random((len(mats), K, len(p0))))
            lc, gc = Lm.eval_multi(fb, pm, grad=True)
            vc = Lm.eval_multi(fb, pm)
            os.environ["BBMCU_LOSS_NO_COMPACT"] = "1"
            try:
                lg, gg = Lm.eval_multi(fb, pm, grad=True)
            finally:
                del os.environ["BBMCU_LOSS_NO_COMPACT"]
            assert np.all(np.abs(lc - lg) <= 2e-5 * np.abs(lg)) and np.all(np.abs(vc - lg) <= 2e-5 * np.abs(lg)), (metric, K)
            assert np.all(np.abs(gc - gg) <= 1e-4 * np.abs(gg) + 2e-6 * np.abs(gg).max(axis=2, keepdims=True)), (metric, K)


# ---- parameter-only factors formed once per thread (Student-t G1) --------------------------------------------------------------
@pytest.mark.parametrize("s", ["Ribardiere()", "RibardiereAnisotropic()", "Ribardiere([0.3, 0.5, 0.7], 0.12, 2.2, 1.6)",
                               "RibardiereAnisotropic([0.3, 0.5, 0.7], [0.1, 0.4], 3.5, 1.4)"])
def test_studentt_factors_per_thread_equal_factors_per_evaluation(ctx, ref, s):
    """The eval kernel, the fused sample -> eval -> pdf pass and the fused-linearizer grid eval form the gamma-only factors of
    the Student-t G1 (ndf/studentt.h:110-140) once per thread (NdfStudentT::pre); the Aggregate(Lambertian, M) pair kernel
    forms them per evaluation through the same functions: all routes agree bit for bit, and with the reference to 1e-5."""
    import bbm_b200 as bb
    rng = np.random.default_rng(5)
    n = (1 << 18) + 3
    z = rng.random(n, dtype=np.float32)
    ph = rng.random(n, dtype=np.float32) * np.float32(2 * np.pi)
    r = np.sqrt(1 - z * z)
    out = np.ascontiguousarray(np.stack([r * np.cos(ph), r * np.sin(ph), z]).astype(np.float32))
    xi = np.ascontiguousarray(rng.random((2, n), dtype=np.float32))
    b = bb.Bsdf(s)
    agg = bb.Bsdf("Aggregate(Lambertian([0, 0, 0]), " + s + ")")          # 0 + x = x: the lobe's value through the pair kernel
    d, sp, f, rgb, p = ctx.sample_eval_pdf(b, out, xi)
    e = ctx.eval(b, d, out)
    assert np.count_nonzero(e[0]) > n // 2
    assert np.array_equal(bits(rgb), bits(e))
    assert np.array_equal(bits(ctx.eval(agg, d, out)), bits(e))
    assert np.array_equal(bits(ctx.pdf(b, d, out))[f != 0], bits(p)[f != 0])
    g_rgb, g_in, g_out = ctx.eval_merl_grid(b, first=1000, n=(1 << 16) + 1, dirs=True)
    assert np.array_equal(bits(g_rgb), bits(ctx.eval(b, np.ascontiguousarray(g_in), np.ascontiguousarray(g_out))))
    if s.endswith("()"):       # the default configurations are the ones the large-sample scan (tests/test_gpu_scan.py) holds at 0 beyond 1e-5
        sub = slice(0, 20000)
        assert_parity(e.T[sub], ref.eval(s, d.T[sub].copy(), out.T[sub].copy()), 1e-5, what="Student-t eval (factors per thread) vs reference: " + s)
