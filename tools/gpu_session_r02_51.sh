#!/bin/bash
# round 2, GPU session 51: Student-t gamma-only factors once per thread also in the fused sample -> eval -> pdf pass and the
# fused-linearizer grid eval: bits against the per-evaluation route (the Aggregate(Lambertian, M) pair kernel), the whole
# GPU suite, smoke, model throughput, ncu of the Ribardiere eval kernel (after-state of r02_s4_ncu_ribardiere_eval), bench arms
mkdir -p gpurun_out
python - > gpurun_out/r02_s51_studentt_bits.log 2>&1 <<'P'
import numpy as np, bbm_b200 as bb
ctx = bb.Context(0)
rng = np.random.default_rng(5)
n = 1 << 20
z = rng.random(n, dtype=np.float32); ph = rng.random(n, dtype=np.float32) * np.float32(2*np.pi)
out = np.ascontiguousarray(np.stack([np.sqrt(1 - z*z)*np.cos(ph), np.sqrt(1 - z*z)*np.sin(ph), z]).astype(np.float32))
xi = np.ascontiguousarray(rng.random((2, n), dtype=np.float32))
def rel(a, b):
    ok = np.isfinite(a) & np.isfinite(b)
    return float(np.max(np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1e-30))) if ok.any() else 0.0
for s in ["Ribardiere()", "RibardiereAnisotropic()", "Ribardiere([0.3, 0.5, 0.7], 0.12, 2.2, 1.6)", "RibardiereAnisotropic([0.3, 0.5, 0.7], [0.1, 0.4], 3.5, 1.4)"]:
    b = bb.Bsdf(s)
    agg = bb.Bsdf("Aggregate(Lambertian([0, 0, 0]), " + s + ")")
    d, sp, f, rgb, p = ctx.sample_eval_pdf(b, out, xi)          # fused pass: factors once per thread
    e = ctx.eval(b, d, out)                                     # eval kernel: factors once per thread
    ea = ctx.eval(agg, d, out)                                  # pair kernel: eval<float>, factors per evaluation (0 + x = x)
    pp = ctx.pdf(b, d, out)
    same_f = np.array_equal(rgb.view(np.uint32), e.view(np.uint32))
    same_a = np.array_equal(ea.view(np.uint32), e.view(np.uint32))
    same_p = np.array_equal(pp.view(np.uint32)[f != 0], p.view(np.uint32)[f != 0])
    g_rgb, g_in, g_out = ctx.eval_merl_grid(b, first=1000, n=1 << 18, dirs=True)
    eg = ctx.eval(b, np.ascontiguousarray(g_in), np.ascontiguousarray(g_out))
    same_g = np.array_equal(np.asarray(g_rgb).view(np.uint32), eg.view(np.uint32))
    print(s, "| fused == eval:", same_f, rel(rgb, e), "| pair == eval:", same_a, rel(ea, e), "| fused pdf == pdf:", same_p, "| grid == eval:", same_g, rel(np.asarray(g_rgb), eg), "| nonzero:", int(np.count_nonzero(e[0])))
    assert (same_f or rel(rgb, e) < 1e-6) and (same_a or rel(ea, e) < 1e-6) and (same_g or rel(np.asarray(g_rgb), eg) < 1e-6)
print("studentt bits ok")
P
echo "studentt bits rc=$?"; tail -6 gpurun_out/r02_s51_studentt_bits.log
python -m pytest tests -m gpu -q > gpurun_out/r02_s51_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_s51_pytest.log
tail -6 gpurun_out/r02_s51_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_s51_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_s51_smoke.log
python tools/model_throughput.py --out gpurun_out/r02_s51_model_throughput.json > gpurun_out/r02_s51_model_throughput.log 2>&1; echo "throughput rc=$?"
grep -i "ribard" gpurun_out/r02_s51_model_throughput.log | head
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_s51_bench_reference.json 2> gpurun_out/r02_s51_bench_reference.err; echo "reference arm rc=$?"
( time python bench.py ) > gpurun_out/r02_s51_bench.json 2> gpurun_out/r02_s51_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r02_s51_bench.err
python -c "
import json
d = json.loads(open('gpurun_out/r02_s51_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['loss_grad']['value'])
"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_foreach4 -s 2 -c 1 -f -o /tmp/r02_s51_rib python tools/run_op.py "Ribardiere()" eval 22 > gpurun_out/r02_s51_ncu_ribardiere_eval.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py /tmp/r02_s51_rib.ncu-rep gpurun_out/r02_s51_ncu_ribardiere_eval_after.csv $((1 << 22)) > gpurun_out/r02_s51_ncu_ribardiere_eval_after.txt 2>&1
head -22 gpurun_out/r02_s51_ncu_ribardiere_eval_after.txt
