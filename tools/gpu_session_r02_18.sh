#!/bin/bash
# round 2, multi-GPU session 18: the bench line at N GPUs with interleaved loss shards
mkdir -p gpurun_out
N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 420 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_s18_bench_${N}gpu.json 2> gpurun_out/r02_s18_bench_${N}gpu.err; echo "bench N=$N rc=$?"
tail -c 600 gpurun_out/r02_s18_bench_${N}gpu.err
