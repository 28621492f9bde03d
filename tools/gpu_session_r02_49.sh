#!/bin/bash
# round 2, GPU session 49: the whole GPU suite, smoke, both bench arms, launch list of the bench command
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_s49_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_s49_pytest.log
tail -6 gpurun_out/r02_s49_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_s49_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_s49_smoke.log
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_s49_bench_reference.json 2> gpurun_out/r02_s49_bench_reference.err; echo "reference arm rc=$?"; tail -3 gpurun_out/r02_s49_bench_reference.err
( time python bench.py ) > gpurun_out/r02_s49_bench.json 2> gpurun_out/r02_s49_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r02_s49_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_s49_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02_s49_ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
ls -la gpurun_out | tail -8
python tools/loss_ab.py --ks 16 --reps 5 --models "He;HeWestin;NganHe" --out gpurun_out/r02_s49_loss_he.json > gpurun_out/r02_s49_loss_he.log 2>&1; echo "he rc=$?"
python -c "
import json
for r in json.load(open('gpurun_out/r02_s49_loss_he.json')):
    print(r['model'], r['metric'], r['K'], 'grad', r['compact_grad_us'], 'value', r['compact_value_us'])
"
