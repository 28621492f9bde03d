#!/bin/bash
# round 2, GPU session 40: single-lobe compact loss kernels - the whole GPU suite (every model's loss + gradient kernel is compared
# with the host-compiled generic code and with finite differences of the doubleRGB reference there), bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_s40_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_s40_pytest.log
tail -12 gpurun_out/r02_s40_pytest.log
python tools/loss_ab.py --single --models "CookTorrance([0.4,0.5,0.6],0.1,1.6)" --out gpurun_out/r02_s40_loss_ab_single.json > gpurun_out/r02_s40_loss_ab.log 2>&1; echo "ab rc=$?"
python -c "
import json
for r in json.load(open('gpurun_out/r02_s40_loss_ab_single.json')):
    print(r['model'], r['metric'], r['K'], r['compact_grad_us'], r['generic_grad_us'], r['speedup_grad'])
"
( time python bench.py ) > gpurun_out/r02_s40_bench.json 2> gpurun_out/r02_s40_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_s40_bench.err
