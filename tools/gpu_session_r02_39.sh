#!/bin/bash
# round 2, GPU session 39: compact loss kernels of the isotropic Ashikhmin-Shirley, Phong / Blinn-Phong and NganLafortune pairs
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -k "loss or compact or gradient or compass or sweep or shard or multi" > gpurun_out/r02_s39_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_s39_pytest.log
tail -8 gpurun_out/r02_s39_pytest.log
python tools/loss_ab.py --models "NganWard;NganWardDuer" --grid offhorizon --out gpurun_out/r02_s39_loss_ab.json > gpurun_out/r02_s39_loss_ab.log 2>&1; echo "ab rc=$?"; tail -3 gpurun_out/r02_s39_loss_ab.log | cut -c1-300
python -c "
import json
for r in json.load(open('gpurun_out/r02_s39_loss_ab.json')):
    print(r['model'], r['metric'], r['K'], r['compact_grad_us'], r['generic_grad_us'], r['speedup_grad'])
"
