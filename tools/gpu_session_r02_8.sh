#!/bin/bash
# round 2, 8-GPU session: the bench line at N = 8 (loss strong scaling through the peer exchange, eval / sample weak scaling),
# then the FULL configs[4] sweep: 100 materials x 34 models x 6 metrics, split by material, no collective
mkdir -p gpurun_out
N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 420 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_s8_bench_${N}gpu.json 2> gpurun_out/r02_s8_bench_${N}gpu.err; echo "bench N=$N rc=$?"
tail -c 1500 gpurun_out/r02_s8_bench_${N}gpu.err
timeout 600 $RUN bench.py --gpus $N --steps 10 --warmup 3 --sweep --sweep-steps ${2:-20} --no-cpu-baseline --no-e2e > gpurun_out/r02_s8_bench_${N}gpu_sweep.json 2> gpurun_out/r02_s8_bench_${N}gpu_sweep.err; echo "sweep N=$N rc=$?"
tail -c 1500 gpurun_out/r02_s8_bench_${N}gpu_sweep.err
cp gpurun_out/sweep_full.json gpurun_out/r02_s8_sweep_full_${N}gpu.json 2>/dev/null
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv | head -12
ls -la gpurun_out | tail -6
