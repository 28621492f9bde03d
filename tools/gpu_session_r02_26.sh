#!/bin/bash
# round 2, GPU session 26: the compact pair loss kernel - parity against the generic tile kernel, A/B timing, full ncu capture
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -k "loss or compact or gradient or compass or sweep or shard or multi" > gpurun_out/r02_s26_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_s26_pytest.log
tail -15 gpurun_out/r02_s26_pytest.log
python tools/loss_ab.py --out gpurun_out/r02_s26_loss_ab.json > gpurun_out/r02_s26_loss_ab.log 2>&1; echo "ab rc=$?"; tail -12 gpurun_out/r02_s26_loss_ab.log
python tools/loss_ab.py --div 8 --models CookTorrance --out gpurun_out/r02_s26_loss_ab_div8.json > gpurun_out/r02_s26_loss_ab_div8.log 2>&1; echo "ab div8 rc=$?"; tail -6 gpurun_out/r02_s26_loss_ab_div8.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_loss_tile_compact -s 1 -c 1 -f -o gpurun_out/r02_s26_loss_compact_full python tools/run_loss.py 256 > gpurun_out/r02_s26_ncu_loss.log 2>&1; echo "ncu loss rc=$?"
ls -la gpurun_out | tail -8
