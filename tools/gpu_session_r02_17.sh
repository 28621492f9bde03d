#!/bin/bash
# round 2, GPU session 17: bench (default flags) + launch list of the bench command
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/r02_s17_bench.json 2> gpurun_out/r02_s17_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r02_s17_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_s17_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02_s17_ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
