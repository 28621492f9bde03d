#!/bin/bash
# round 2, GPU session 50: Student-t G1 with its gamma-only factors formed once per thread (eval kernels of Ribardiere /
# RibardiereAnisotropic): bit-equality of the two eval routes, the whole GPU suite, smoke, model throughput, both bench arms
mkdir -p gpurun_out
python - > gpurun_out/r02_s50_studentt_bits.log 2>&1 <<'P'
import numpy as np, bbm_b200 as bb
ctx = bb.Context(0)
rng = np.random.default_rng(5)
n = 1 << 20
z = rng.random(n, dtype=np.float32); ph = rng.random(n, dtype=np.float32) * np.float32(2*np.pi)
out = np.ascontiguousarray(np.stack([np.sqrt(1 - z*z)*np.cos(ph), np.sqrt(1 - z*z)*np.sin(ph), z]).astype(np.float32))
xi = np.ascontiguousarray(rng.random((2, n), dtype=np.float32))
for s in ["Ribardiere()", "RibardiereAnisotropic()", "Ribardiere([0.3, 0.5, 0.7], 0.12, 2.2, 1.6)", "RibardiereAnisotropic([0.3, 0.5, 0.7], [0.1, 0.4], 3.5, 1.4)"]:
    b = bb.Bsdf(s)
    d, sp, f, rgb, p = ctx.sample_eval_pdf(b, out, xi)          # fused pass: eval<float> (factors per evaluation)
    e = ctx.eval(b, d, out)                                     # eval kernel: factors once per thread
    same = np.array_equal(rgb.view(np.uint32), e.view(np.uint32))
    ok = np.isfinite(e) & np.isfinite(rgb)
    rel = float(np.max(np.abs(e[ok] - rgb[ok]) / np.maximum(np.abs(rgb[ok]), 1e-30))) if ok.any() else 0.0
    print(s, "bit-identical:", same, "max rel diff:", rel, "nonzero:", int(np.count_nonzero(e[0])), "max:", float(np.nanmax(e)))
    assert same or rel < 1e-6
print("studentt bits ok")
P
echo "studentt bits rc=$?"; tail -5 gpurun_out/r02_s50_studentt_bits.log
python -m pytest tests -m gpu -q > gpurun_out/r02_s50_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_s50_pytest.log
tail -6 gpurun_out/r02_s50_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_s50_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_s50_smoke.log
python tools/model_throughput.py --out gpurun_out/r02_s50_model_throughput.json > gpurun_out/r02_s50_model_throughput.log 2>&1; echo "throughput rc=$?"
grep -i "ribard\|bagher" gpurun_out/r02_s50_model_throughput.log | head
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_s50_bench_reference.json 2> gpurun_out/r02_s50_bench_reference.err; echo "reference arm rc=$?"
( time python bench.py ) > gpurun_out/r02_s50_bench.json 2> gpurun_out/r02_s50_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r02_s50_bench.err
python -c "
import json
d = json.loads(open('gpurun_out/r02_s50_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['loss_grad']['value'])
"
