#!/bin/bash
# round 2, GPU session 45: value-only batched losses of the He family through the all-float templates (Dual<0>)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -k "loss or compact or gradient or compass or sweep or shard or multi" > gpurun_out/r02_s45_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_s45_pytest.log
tail -5 gpurun_out/r02_s45_pytest.log
python tools/loss_ab.py --ks 16 --reps 5 --models "He;HeWestin;HeHolzschuch;NganHe" --out gpurun_out/r02_s45_loss_he.json > gpurun_out/r02_s45_loss_he.log 2>&1; echo "ab rc=$?"
python -c "
import json
for r in json.load(open('gpurun_out/r02_s45_loss_he.json')):
    print(r['model'], r['metric'], r['K'], 'grad', r['compact_grad_us'], 'value', r['compact_value_us'])
"
