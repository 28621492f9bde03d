#!/bin/bash
# round 2, GPU session 34: headline kernel fast path (full groups without per-plane tests), GGX compact loss kernel
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_s34_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_s34_pytest.log
tail -5 gpurun_out/r02_s34_pytest.log
python tools/loss_ab.py --models GGX,LowMicrofacet --out gpurun_out/r02_s34_loss_ab_ggx.json > gpurun_out/r02_s34_loss_ab_ggx.log 2>&1; echo "ab rc=$?"; tail -6 gpurun_out/r02_s34_loss_ab_ggx.log | cut -c1-400
( time python bench.py --no-extras ) > gpurun_out/r02_s34_bench.json 2> gpurun_out/r02_s34_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r02_s34_bench.err
python -c "
import json
d = json.loads(open('gpurun_out/r02_s34_bench.json').read().strip().splitlines()[-1])
print(json.dumps({k: d.get(k) for k in ('value', 'ms_per_step', 'roofline', 'clocks')}), d['e2e']['value'])
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_foreach4 -s 3 -c 1 -f -o gpurun_out/r02_s34_ggx_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --no-loss > gpurun_out/r02_s34_ncu.log 2>&1; echo "ncu rc=$?"
