#!/bin/bash
# round 2, GPU session 3: tests, bench, launch list, per-model throughput table, full captures (headline, loss, multi-material loss)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_s3_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_s3_pytest.log
tail -30 gpurun_out/r02_s3_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_s3_bench.json 2> gpurun_out/r02_s3_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02_s3_bench.err
python tools/model_throughput.py --log2 24 --out gpurun_out/r02_s3_model_throughput.json > gpurun_out/r02_s3_model_throughput.log 2>&1; echo "model throughput rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_s3_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02_s3_ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_loss_tile -s 1 -c 1 -f -o gpurun_out/r02_s3_loss_full python tools/run_loss.py 256 > gpurun_out/r02_s3_ncu_loss.log 2>&1; echo "ncu loss rc=$?"
ls -la gpurun_out | tail -20
