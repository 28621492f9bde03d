#!/bin/bash
# round 2, GPU session 29: compact kernel with the material loop - loss tests, A/B timing, the whole bench line
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -k "loss or compact or gradient or compass or sweep or shard or multi" > gpurun_out/r02_s29_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_s29_pytest.log
tail -15 gpurun_out/r02_s29_pytest.log
python tools/loss_ab.py --models CookTorrance --out gpurun_out/r02_s29_loss_ab.json > gpurun_out/r02_s29_loss_ab.log 2>&1; echo "ab rc=$?"; tail -6 gpurun_out/r02_s29_loss_ab.log
python tools/loss_ab.py --div 8 --models CookTorrance --out gpurun_out/r02_s29_loss_ab_div8.json > gpurun_out/r02_s29_loss_ab_div8.log 2>&1; echo "ab div8 rc=$?"
( time python bench.py ) > gpurun_out/r02_s29_bench.json 2> gpurun_out/r02_s29_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r02_s29_bench.err
python -c "
import json
d = json.loads(open('gpurun_out/r02_s29_bench.json').read().strip().splitlines()[-1])
print(json.dumps({k: d.get(k) for k in ('value', 'ms_per_step', 'e2e', 'roofline')}))
print(json.dumps(d.get('loss_grad'), indent=1))
print(json.dumps(d.get('loss_multi'), indent=1))
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_loss_tile_compact -s 1 -c 1 -f -o gpurun_out/r02_s29_loss_compact_full python tools/run_loss.py 256 > gpurun_out/r02_s29_ncu_loss.log 2>&1; echo "ncu loss rc=$?"
