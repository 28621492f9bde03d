#!/bin/bash
# round 2, GPU session 41: host-thread count of the host-pointer path (end-to-end step), same box
mkdir -p gpurun_out
nproc
for t in 2 4 8 12 16; do
  BBMCU_HOST_THREADS=$t python bench.py --no-extras --no-loss --no-cpu-baseline > gpurun_out/r02_s41_bench_t$t.json 2> gpurun_out/r02_s41_bench_t$t.err
  python -c "
import json
d = json.loads(open('gpurun_out/r02_s41_bench_t$t.json').read().strip().splitlines()[-1])
print('threads', $t, 'e2e', round(d['e2e']['value'], 3), [round(v['value'], 3) for v in d.get('e2e_variants', [])])
"
done
