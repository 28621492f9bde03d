"""GPU parity tests: the CUDA path through the C ABI (bbm_b200 -> libbbmcu.so) against
(1) the committed golden vectors of the unmodified reference and (2) the compiled reference itself
(oracle/_ref, when present) on seeded inputs.  Tolerances are SURVEY.md section 8(c)'s:
1e-5 relative for eval / pdf / reflectance / sampled directions, exact flags and bin indices,
1e-4 for loss and gradient."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from tests.test_models_hostsim import COMPONENTS, _cases, defined_samples
from tests.util import assert_parity, pdf_floor, soa, uses_only_implemented

pytestmark = pytest.mark.gpu


def hemisphere(rng, n):
    z = rng.random(n)
    ph = rng.random(n) * 2 * np.pi
    s = np.sqrt(1 - z * z)
    return np.stack([s * np.cos(ph), s * np.sin(ph), z], 1).astype(np.float32)


def test_golden_eval_pdf_reflectance(ctx, golden_models):
    import bbm_b200 as bb
    arr, meta = golden_models
    inn, out = soa(arr["in"]), soa(arr["out"])
    n_cases = 0
    for key, rec in _cases(meta):
        s = rec["string"]
        if not uses_only_implemented(s):
            continue
        b = bb.Bsdf(s)
        for c in COMPONENTS:
            assert_parity(ctx.eval(b, inn, out, c).T, arr[f"{key}_eval_c{c}"], 1e-5, what=f"eval {s} comp {c}")
            assert_parity(ctx.pdf(b, inn, out, c), arr[f"{key}_pdf_c{c}"], 1e-5, floor=pdf_floor(s, arr[f"{key}_pdf_c{c}"]), what=f"pdf {s} comp {c}")
            assert_parity(ctx.reflectance(b, out, c).T, arr[f"{key}_refl_c{c}"], 1e-5, what=f"reflectance {s} comp {c}")
        n_cases += 1
    assert n_cases >= 60


def test_golden_sample(ctx, golden_models):
    import bbm_b200 as bb
    arr, meta = golden_models
    out, xi = soa(arr["out"]), soa(arr["xi"])
    for key, rec in _cases(meta):
        s = rec["string"]
        if not uses_only_implemented(s):
            continue
        b = bb.Bsdf(s)
        for c in COMPONENTS:
            d, p, f = ctx.sample(b, out, xi, c)
            ok = defined_samples(s, arr[f"{key}_refl_c{c}"], arr["xi"])
            assert np.array_equal(f[ok], arr[f"{key}_sflag_c{c}"].astype(np.int32)[ok]), f"sample flag {s} comp {c}"
            assert_parity(d.T[ok], arr[f"{key}_sdir_c{c}"][ok], 1e-5, floor=1e-5, what=f"sample dir {s} comp {c}")
            # loose: the golden pdf was evaluated at the reference's own direction; a last-bit difference of the
            # direction moves the pdf of a sharp lobe (B = 6e4 fits) by up to 1e-2.  The strict check is
            # test_sample_pdf_consistency_on_gpu_direction (oracle pdf at the GPU's direction, 1e-5)
            assert_parity(p[ok], arr[f"{key}_spdf_c{c}"][ok], 3e-2, what=f"sample pdf {s} comp {c}")


def test_sample_pdf_consistency_on_gpu_direction(ctx, ref, golden_models):
    """sample.pdf must equal the ORACLE's pdf evaluated at the GPU's own sampled direction (SURVEY.md
    'Input sensitivity of sharp lobes')"""
    import bbm_b200 as bb
    arr, meta = golden_models
    out, xi = arr["out"], arr["xi"]
    cases = ["GGX([0.1, 0.2, 0.3], 0.01, 1.5)", "CookTorrance([0.1, 0.2, 0.3], 0.02, 2.5)", "Phong([0.2, 0.3, 0.4], 800)", "Ward([0.3, 0.2, 0.1], [0.05, 0.3])"]
    cases += [v for v in meta["fits"].values() if uses_only_implemented(v)]
    for s in cases:
        d, p, f = ctx.sample(bb.Bsdf(s), soa(out), soa(xi))
        want = ref.pdf(s, d.T.copy(), out)
        ok = f != 0                                  # masked samples return {0, 0, None}
        assert_parity(p[ok], want[ok], 1e-5, floor=pdf_floor(s, want[ok]), what=f"sample pdf at gpu direction {s}")


def test_device_and_host_pointers_agree(ctx):
    import torch
    import bbm_b200 as bb
    rng = np.random.default_rng(3)
    n = 100003                      # odd size: exercises the unaligned tail
    inn, out = soa(hemisphere(rng, n)), soa(hemisphere(rng, n))
    b = bb.Bsdf("Aggregate(Lambertian([0.2,0.1,0.05]), GGX([0.3,0.3,0.3], 0.2, 1.5))")
    host = ctx.eval(b, inn, out)
    dev = ctx.eval(b, torch.from_numpy(inn).cuda(), torch.from_numpy(out).cuda())
    ctx.synchronize()
    assert np.array_equal(host.view(np.uint32), dev.cpu().numpy().view(np.uint32))


def test_eval_against_reference_random(ctx, ref):
    import bbm_b200 as bb
    from tests.util import implemented_models
    rng = np.random.default_rng(11)
    n = 20000
    a, b = hemisphere(rng, n), hemisphere(rng, n)
    for name in implemented_models():
        if name == "Merl":
            continue                                   # needs a file: test_merl_model_on_gpu
        s = name + "()"
        got = ctx.eval(bb.Bsdf(s), soa(a), soa(b)).T
        assert_parity(got, ref.eval(s, a, b, threads=8), 1e-5, what=f"eval {s}")
        got = ctx.pdf(bb.Bsdf(s), soa(a), soa(b))
        want = ref.pdf(s, a, b, threads=8)
        assert_parity(got, want, 1e-5, floor=pdf_floor(s, want), what=f"pdf {s}")


def test_merl_index_golden_bit_exact(ctx, golden_lin):
    g = golden_lin
    idx = ctx.merl_index(soa(g["pairs_in"]), soa(g["pairs_out"]))
    assert np.array_equal(idx, g["pairs_index"])
    idx = ctx.merl_index(soa(g["grid_in"]), soa(g["grid_out"]))
    assert np.array_equal(idx, g["grid_index_of_dirs"])


def test_merl_dirs_golden_bit_exact(ctx, golden_lin):
    g = golden_lin
    i, o = ctx.merl_dirs(0, 1458000)
    assert np.array_equal(i.T[g["grid_idx"]].view(np.uint32), g["grid_in"].view(np.uint32))
    assert np.array_equal(o.T[g["grid_idx"]].view(np.uint32), g["grid_out"].view(np.uint32))


def test_spherical_dirs_golden_bit_exact(ctx, golden_lin):
    import bbm_b200 as bb
    g = golden_lin
    i, o = ctx.spherical_dirs(bb.spherical_grid((12, 7), (5, 6)), 0, 12*7*5*6)
    assert np.array_equal(i.T.view(np.uint32), g["sph_in"].view(np.uint32)) and np.array_equal(o.T.view(np.uint32), g["sph_out"].view(np.uint32))
    i, o = ctx.spherical_dirs(bb.spherical_grid((9, 4), (1, 5), (0.1, 0.05), (3.0, 1.4), (0.0, 0.2), (6.0, 1.5)), 0, 9*4*5)
    assert np.array_equal(i.T.view(np.uint32), g["sph2_in"].view(np.uint32)) and np.array_equal(o.T.view(np.uint32), g["sph2_out"].view(np.uint32))


def test_merl_index_full_grid_and_random_against_reference(ctx, ref):
    """all 1 458 000 grid directions and 2^22 random pairs: the bin index must be bit-exact"""
    i, o = ctx.merl_dirs(0, 1458000)
    ri, ro = ref.merl_dirs(0, 1458000, threads=8)
    assert np.array_equal(i.T.view(np.uint32), ri.view(np.uint32)) and np.array_equal(o.T.view(np.uint32), ro.view(np.uint32))
    assert np.array_equal(ctx.merl_index(i, o), ref.merl_index(ri, ro, threads=8).astype(np.uint32))
    rng = np.random.default_rng(12)
    n = 1 << 22
    a, b = hemisphere(rng, n), hemisphere(rng, n)
    got = ctx.merl_index(soa(a), soa(b))
    want = ref.merl_index(a, b, threads=8).astype(np.uint32)
    assert int((got != want).sum()) == 0


def test_merl_index_below_horizon_and_nan(ctx):
    inn = soa(np.array([[0, 0, 1], [0.6, 0, -0.8], [-1, 0, 0], [0, 0, 1]], np.float32))
    out = soa(np.array([[0, 0, 1], [0, 0, 1], [1, 0, 0], [0.6, 0, -0.8]], np.float32))
    idx = ctx.merl_index(inn, out)
    assert idx[1] == 1458000 and idx[3] == 1458000          # size() for masked pairs
    assert idx[2] == 0xFFFFFFFF                              # antipodal grazing pair: NaN half vector
    assert idx[0] < 1458000


def test_loss_terms_and_totals_golden(ctx, golden_loss):
    import bbm_b200 as bb
    arr, meta = golden_loss
    fitted, truth = bb.Bsdf(meta["fitted"]), bb.Bsdf(meta["truth"])
    for name, m in meta["metrics"].items():
        low = name in ("lowL2", "lowLog")
        grid = bb.spherical_grid((13, 8), m["samples_out"], end_out=(2*np.pi, 0.5*np.pi) if low else None)
        L = ctx.loss(name, truth, grid)
        assert L.samples() == m["N"]
        assert_parity(L.terms(fitted, m["N"]), arr[name + "_terms"], 1e-5, floor=1e-12, what=f"{name} per-sample terms")
        loss, grad = L(fitted, grad=True)
        assert abs(loss[0] - m["double_sum_of_float_terms"]) <= 1e-4*abs(m["double_sum_of_float_terms"]), name
        assert abs(loss[0] - m["double_total"]) <= 1e-4*abs(m["double_total"]), name
        fd = np.array(m["fd_gradient"])
        assert np.all(np.abs(grad[0] - fd) <= 1e-4*np.abs(fd) + 1e-9), (name, grad[0], fd)
        # loss without gradient takes the float-only kernel path: same value
        assert abs(L(fitted)[0] - loss[0]) <= 1e-6*abs(loss[0])
    L = ctx.loss("nganL2", truth, None, first=700000, count=4096)
    assert_parity(L.terms(fitted, 4096), arr["merl_nganL2_terms_700000"], 1e-5, floor=1e-12, what="MERL-grid nganL2 terms")


def test_loss_batched_and_sharded(ctx):
    """K parameter sets in one launch equal K single launches; two shards add up to the whole"""
    import bbm_b200 as bb
    fitted = bb.Bsdf("Aggregate(Lambertian(), CookTorrance())")
    truth = bb.Bsdf("Aggregate(Lambertian([0.2,0.1,0.05]), CookTorrance([0.3,0.3,0.3], 0.2, 1.5))")
    L = ctx.loss("standardLog", truth, None)
    rng = np.random.default_rng(5)
    p0 = fitted.parameter_values()
    params = p0[None] * (1 + 0.2*rng.random((7, len(p0))))
    params[:, 7] = 1.2 + rng.random(7)
    lk, gk = L(fitted, params, grad=True)
    for k in range(7):
        l1, g1 = L(fitted, params[k], grad=True)
        # the tile kernel reduces every (parameter set, sample tile) in the same order whatever K is: bit-identical
        assert l1[0] == lk[k] and np.array_equal(g1[0], gk[k])
    A = ctx.loss("standardLog", truth, None, first=0, count=700001)
    B = ctx.loss("standardLog", truth, None, first=700001, count=1458000 - 700001)
    la, ga = A(fitted, params, grad=True)
    lb, gb = B(fitted, params, grad=True)
    # a shard boundary moves the 1024-sample tiles, i.e. which samples share a float partial sum: equal to ~1e-6
    assert np.allclose(la + lb, lk, rtol=2e-6) and np.allclose(ga + gb, gk, rtol=2e-5, atol=1e-7*np.abs(gk).max())


def test_loss_peer_exchange_two_shards_one_device(ctx):
    """the finish kernel fused with the exchange over peer memory (bbmcu_loss_peer_*): two shards of one loss, each on its
    own context (stream) of this device, combine through each other's windows; both get the same bits, equal to the
    unsharded loss, over three batches (both window parities and the first one reused)"""
    import torch
    import bbm_b200 as bb
    ctx_b = bb.Context(0)
    fitted = bb.Bsdf("Aggregate(Lambertian(), CookTorrance())")
    truth = bb.Bsdf("Aggregate(Lambertian([0.2,0.1,0.05]), CookTorrance([0.3,0.3,0.3], 0.2, 1.5))")
    whole = ctx.loss("nganL2", truth, None)
    A = ctx.loss("nganL2", truth, None, first=0, count=700001)
    B = ctx_b.loss("nganL2", truth, None, first=700001, count=1458000 - 700001)
    rng = np.random.default_rng(9)
    p0 = fitted.parameter_values()
    K, cols = 5, 1 + len(p0)
    params = p0[None] * (1 + 0.2*rng.random((K, len(p0))))
    params[:, 7] = 1.2 + rng.random(K)
    _, wa = A.peer_init(0, 2, 8*cols)
    _, wb = B.peer_init(1, 2, 8*cols)
    A.peer_connect_ptrs([wa, wb])
    B.peer_connect_ptrs([wa, wb])
    dev = torch.device("cuda", 0)
    for batch in range(3):
        p = params * (1 + 0.01*batch)
        ra = torch.zeros((K, cols), dtype=torch.float64, device=dev)
        rb = torch.zeros((K, cols), dtype=torch.float64, device=dev)
        A.eval_device(fitted, p, ra)            # asynchronous: A's exchange block waits on the device for B's rows
        B.eval_device(fitted, p, rb)
        ctx.synchronize(); ctx_b.synchronize()
        assert torch.equal(ra, rb)
        lw, gw = whole(fitted, p, grad=True)
        got = ra.cpu().numpy()
        assert np.allclose(got[:, 0], lw, rtol=2e-6) and np.allclose(got[:, 1:], gw, rtol=2e-5, atol=1e-7*np.abs(gw).max())
    with pytest.raises(bb.BbmInvalidArgument):
        A.eval_device(fitted, np.repeat(params, 2, 0), torch.zeros((2*K, cols), dtype=torch.float64, device=dev))     # 10 rows > the 8-row window
    with pytest.raises(bb.BbmInvalidArgument):
        A.peer_init(0, 2, 8*cols)                                       # one window per loss
    with pytest.raises(bb.BbmInvalidArgument):
        whole.peer_init(2, 2, 8*cols)                                   # rank outside [0, world)
    with pytest.raises(bb.BbmInvalidArgument):
        whole.peer_connect_ptrs([wa, wb])                               # connect before init


def test_pair_kernels_equal_runtime_lobe_list():
    """Aggregate(Lambertian, M) runs on the compile-time pair kernels; the run-time lobe list (BBMCU_DISABLE_PAIR_KERNELS=1,
    read once per process, hence a subprocess) does the same arithmetic in the same order: every output bit-identical"""
    import subprocess
    import sys
    code = r'''
import sys, hashlib
import numpy as np
sys.path.insert(0, %r)
import bbm_b200 as bb
ctx = bb.Context(0)
rng = np.random.default_rng(21)
n = 1 << 16
def hemi():
    z = rng.random(n); ph = rng.random(n) * 2 * np.pi; s = np.sqrt(1 - z * z)
    return np.ascontiguousarray(np.stack([s * np.cos(ph), s * np.sin(ph), z]).astype(np.float32))
i, o, xi = hemi(), hemi(), np.ascontiguousarray(rng.random((2, n)).astype(np.float32))
h = hashlib.sha256()
for m in ("CookTorrance([0.3,0.3,0.3], 0.2, 1.5)", "GGX([0.5,0.4,0.3], 0.15, 1.7)", "Ward()", "NganLafortune()", "LowMicrofacet()", "AshikhminShirley()"):
    b = bb.Bsdf("Aggregate(Lambertian([0.2,0.1,0.05]), %%s)" %% m)
    for a in (ctx.eval(b, i, o), ctx.pdf(b, i, o)) + tuple(ctx.sample(b, o, xi)) + tuple(ctx.sample_eval_pdf(b, o, xi)):
        h.update(np.ascontiguousarray(a).tobytes())
print(h.hexdigest())
''' % ROOT
    digests = []
    for flag in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, BBMCU_DISABLE_PAIR_KERNELS=flag), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        digests.append(r.stdout.strip().splitlines()[-1])
    assert digests[0] == digests[1]


def test_loss_against_measured_table(ctx, ref, tmp_path):
    """reference operand = a MERL binary written by us and read by the unmodified reference's merl<> loader"""
    import bbm_b200 as bb
    truth = bb.Bsdf("Aggregate(Lambertian([0.2,0.1,0.05]), CookTorrance([0.3,0.3,0.3], 0.2, 1.5))")
    i, o = ctx.merl_dirs(0, 1458000)
    table = ctx.eval(truth, i, o)
    path = str(tmp_path / "synthetic.binary")
    ctx.merl_write(path, table)
    back = ctx.merl_read(path)
    assert_parity(back, table, 1e-6, what="MERL binary round trip")
    fitted = "Aggregate(Lambertian(), CookTorrance())"
    grid = bb.spherical_grid((31, 16), (1, 9))
    L = ctx.loss("standardLog", back, grid)
    from oracle.refbind import sph_desc
    terms, _ = ref.loss("standardLog", sph_desc((31, 16), (1, 9)), fitted, f'Merl("{path}")', 0, 31*16*9, want_total=False)
    assert_parity(L.terms(bb.Bsdf(fitted), 31*16*9), terms, 1e-5, floor=1e-12, what="standardLog vs merl<> reference")


def test_grid_census_against_reference(ctx, ref):
    """whole-grid properties of the reference that a restatement has to reproduce exactly:
    which bins do NOT round-trip through index -> directions -> index, and where Ward / He evaluate to NaN / Inf"""
    import bbm_b200 as bb
    N = bb.MERL_BINS
    i, o = ctx.merl_dirs(0, N)
    back = ctx.merl_index(i, o)
    no_round_trip = int((back != np.arange(N, dtype=np.uint32)).sum())
    ri, ro = ref.merl_dirs(0, N, threads=8)
    want = int((ref.merl_index(ri, ro, threads=8) != np.arange(N)).sum())
    assert no_round_trip == want == 378883                       # samples sit on bin lower edges (merl_linearizer.h:72-73)
    assert int((i[2] == 0).sum()) == 174592 and int((o[2] == 0).sum()) == 171978
    for name, nan_inf in (("Ward", (133744, 212826)), ("WardDuer", (133744, 212826)), ("NganWard", (133744, 212826)), ("He", (1, 1)), ("CookTorrance", (0, 0))):
        s = name + "()"
        got = ctx.eval(bb.Bsdf(s), i, o).T
        assert_parity(got, ref.eval(s, ri, ro, threads=8), 1e-5, what=f"grid eval {s}")      # NaN == NaN, Inf == Inf with sign
        bad = ~np.isfinite(got).all(1)
        nan = int(np.isnan(got).any(1).sum())
        assert (nan, int(bad.sum()) - nan) == nan_inf, (name, nan, int(bad.sum()) - nan)


def test_full_size_properties(ctx):
    """2^26 pairs (the BASELINE config-2 workload at full size): sample.pdf == pdf(sample.direction), flags consistent,
    unit sampled directions, reciprocity of eval"""
    import torch
    import bbm_b200 as bb
    n = 1 << 26
    g = torch.Generator(device="cuda").manual_seed(1)
    z = torch.rand(n, device="cuda", generator=g)
    ph = torch.rand(n, device="cuda", generator=g) * (2*np.pi)
    s = torch.sqrt(1 - z*z)
    out = torch.stack([s*torch.cos(ph), s*torch.sin(ph), z]).contiguous()
    xi = torch.rand((2, n), device="cuda", generator=g)
    b = bb.Bsdf("GGX()")
    d, sp, f, rgb, p = ctx.sample_eval_pdf(b, out, xi)
    ctx.synchronize()
    assert bool(((f == 2) | (f == 0)).all())
    assert torch.equal(sp, p)                                  # sample.pdf is pdf(sample.direction, out)
    ok = (f == 2) & (d[2] > 0)
    rev = ctx.eval(b, out, d)                                  # reciprocity: eval(in, out) == eval(out, in)
    ctx.synchronize()
    rel = ((rev - rgb).abs() / rgb.abs().clamp_min(1e-20))[:, ok]
    assert float(rel.max()) < 1e-4
    nrm = (d*d).sum(0)[f == 2]
    assert float((nrm - 1).abs().max()) < 1e-5


def test_fit_sweep_small(ctx):
    """BASELINE configs[4] in miniature: synthetic MERL-shaped materials x models x metrics, batched compass per fit;
    two ranks' shares are disjoint and cover every job (no collective)."""
    import bbm_b200 as bb
    from bbm_b200.fit import run_sweep, sweep_jobs, fit
    i, o = ctx.merl_dirs(0, bb.MERL_BINS)
    mats = {"matA": "Aggregate(Lambertian([0.2,0.1,0.05]), GGX([0.3,0.3,0.3], 0.2, 1.5))",
            "matB": "Aggregate(Lambertian([0.05,0.1,0.2]), Phong([0.2,0.2,0.2], 40))"}
    tables = {k: ctx.eval(bb.Bsdf(v), i, o) for k, v in mats.items()}
    models, metrics = ["GGX", "Phong"], ["nganL2", "standardLog"]
    r0 = run_sweep(ctx, tables, models, metrics, rank=0, world=2, max_steps=8)
    r1 = run_sweep(ctx, tables, models, metrics, rank=1, world=2, max_steps=8)
    jobs, _ = sweep_jobs(sorted(tables), models, metrics)
    assert set(r0) | set(r1) == set(jobs) and not (set(r0) & set(r1))
    for key, (s, loss, steps) in {**r0, **r1}.items():
        assert np.isfinite(loss) and steps == 8
        bb.Bsdf(s)                                            # the result is a valid BSDF string
    # the matching model gets (much) closer than the wrong one
    assert r0.get(("matA", "GGX", "nganL2"), r1.get(("matA", "GGX", "nganL2")))[1] < {**r0, **r1}[("matA", "Phong", "nganL2")][1]
    # a longer fit of the right model recovers the material
    b, trace = fit(ctx, "Aggregate(Lambertian(), GGX())", tables["matA"], "nganL2", max_steps=60)
    assert trace[-1] < 1e-2 * trace[0]


def test_merl_model_on_gpu(ctx, ref, tmp_path):
    """Merl("file") as a first-class BSDF: eval / sample / pdf kernels against the unmodified reference reading the
    same file, and as the reference operand of a loss (== the raw-table path)"""
    import bbm_b200 as bb
    from tests.util import mismatch
    truth = bb.Bsdf("Aggregate(Lambertian([0.2,0.1,0.05]), GGX([0.3,0.3,0.3], 0.25, 1.5))")
    i, o = ctx.merl_dirs(0, bb.MERL_BINS)
    table = ctx.eval(truth, i, o)
    path = str(tmp_path / "synthetic.binary")
    ctx.merl_write(path, table)
    s = f'Merl("{path}")'
    b = bb.Bsdf(s)
    assert b.to_string() == s and len(b.parameter_values()) == 0
    rng = np.random.default_rng(4)
    n = 1 << 16
    z = rng.random(n); ph = rng.random(n)*2*np.pi; r = np.sqrt(1 - z*z)
    out = np.stack([r*np.cos(ph), r*np.sin(ph), z], 1).astype(np.float32)
    z = rng.random(n); ph = rng.random(n)*2*np.pi; r = np.sqrt(1 - z*z)
    inn = np.stack([r*np.cos(ph), r*np.sin(ph), z], 1).astype(np.float32)
    xi = rng.random((n, 2)).astype(np.float32)
    assert np.array_equal(ctx.eval(b, soa(inn), soa(out)).T, ref.eval(s, inn, out))
    want_p = ref.pdf(s, inn, out)
    assert_parity(ctx.pdf(b, soa(inn), soa(out)), want_p, 1e-5, floor=pdf_floor("He", want_p), what="Merl pdf")
    d, p, f = ctx.sample(b, soa(out), soa(xi))
    d2, p2, f2 = ref.sample(s, out, xi)
    assert np.array_equal(f, f2)
    assert_parity(d.T, d2, 1e-5, floor=1e-5, what="Merl sampled direction", max_bad=2)
    want = ref.pdf(s, d.T.copy(), out)
    assert_parity(p, want, 1e-5, floor=pdf_floor("He", want), what="Merl sample pdf at gpu direction")
    # as a loss reference: BSDF object and raw table agree bit for bit
    fitted = bb.Bsdf("Aggregate(Lambertian(), GGX())")
    la = ctx.loss("nganL2", b, None)(fitted)
    lb = ctx.loss("nganL2", ctx.merl_read(path), None)(fitted)
    assert la[0] == lb[0]


def test_hp_g1_table_regenerated_on_gpu(ctx):
    """the reference's G1 generator (hours on one core) as one kernel; the result is the shipped
    table (printed with 6 significant digits by the generator) to 5e-6 relative / 1e-6 absolute, every one of 100 000 entries"""
    shipped = np.fromfile(os.path.join(ROOT, "bbm_b200", "data", "epd_g1.f32"), np.float32).reshape(100, 1000)
    got = ctx.hp_precompute_g1()
    err = np.abs(got.astype(np.float64) - shipped)
    tol = 5e-6 * np.abs(shipped) + 1e-6
    assert np.isfinite(got).all()
    assert int((err > tol).sum()) == 0, (int((err > tol).sum()), float(err.max()), np.argwhere(err > tol)[:5])


@pytest.mark.parametrize("metric", ["nganL2", "bieronLog"])
def test_loss_and_gradient_kernels_of_every_model(ctx, hostsim, metric):
    """every compile-time loss kernel (one per model) on the device against the same device headers compiled for the host
    (tests/hostsim), whose values and gradients are pinned against the reference in tests/test_gradients_hostsim.py"""
    import bbm_b200 as bb
    from tests.test_gradients_hostsim import CASES
    hp = float(np.float32(2) * np.float32(np.pi))
    tp, t0 = 1.4, 0.05
    grid = bb.spherical_grid((11, 6), (4, 5), start_in=(0, t0), start_out=(0, t0), end_in=(hp, tp), end_out=(hp, tp))
    N = 11 * 6 * 4 * 5
    i, o = hostsim.spherical_dirs([11, 6, 4, 5], [0, t0, hp, tp, 0, t0, hp, tp], 0, N)
    m = bb.METRICS.index(metric)
    for fitted, truth in CASES:
        fb = bb.Bsdf(fitted)
        P = len(fb.parameter_values())
        L = ctx.loss(metric, bb.Bsdf(truth), grid)
        loss, grad = L(fb, grad=True)
        plain = L(fb)
        tv = hostsim.eval(truth, i, o)
        want_l, want_g, _ = hostsim.loss(fitted, m, i, o, tv, nparams=P)
        assert abs(loss[0] - want_l) <= 1e-4 * abs(want_l), (fitted, loss[0], want_l)
        assert abs(plain[0] - want_l) <= 1e-4 * abs(want_l), (fitted, plain[0], want_l)
        tol = 2e-4 * np.abs(want_g) + 2e-4 * np.abs(want_g).max()
        assert np.all(np.abs(grad[0] - want_g) <= tol), (fitted, metric, grad[0], want_g)
