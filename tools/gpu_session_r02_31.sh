#!/bin/bash
# round 2, GPU session 31: eval of the hand-merged GGX lobe crosses the link before its RGB scale (21 B per pair down instead of 29)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "host_path or optional or pageable or stride" > gpurun_out/r02_s31_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_s31_pytest.log
tail -5 gpurun_out/r02_s31_pytest.log
( time python bench.py --no-extras ) > gpurun_out/r02_s31_bench.json 2> gpurun_out/r02_s31_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r02_s31_bench.err
BBMCU_HOST_TRANSFER_PLAIN=1 python bench.py --no-extras --no-loss --no-cpu-baseline > gpurun_out/r02_s31_bench_plain.json 2> gpurun_out/r02_s31_bench_plain.err; echo "bench plain rc=$?"
python -c "
import json
for f in ('gpurun_out/r02_s31_bench.json', 'gpurun_out/r02_s31_bench_plain.json'):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, json.dumps({k: d.get(k) for k in ('value', 'ms_per_step')}), json.dumps(d['e2e']['value']), d['e2e']['ms_per_step'], [(v['contract'][:40], round(v['value'], 3)) for v in d.get('e2e_variants', [])])
"
