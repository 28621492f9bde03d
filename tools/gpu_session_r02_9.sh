#!/bin/bash
# round 2, GPU session 9: EPD G1 rows staged in shared memory - whole GPU suite, throughput table, capture of the EPD eval kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_s9_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_s9_pytest.log
tail -8 gpurun_out/r02_s9_pytest.log
python tools/model_throughput.py --log2 24 --out gpurun_out/r02_s9_model_throughput.json > gpurun_out/r02_s9_model_throughput.log 2>&1; echo "model throughput rc=$?"
grep -i "EPD\|Bagher\|Ribard" gpurun_out/r02_s9_model_throughput.log
cap() { # name bsdf op log2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_foreach4 -s 2 -c 1 -f -o /tmp/r02_s9_$1 python tools/run_op.py "$2" $3 $4 > gpurun_out/r02_s9_$1.log 2>&1; echo "ncu $1 rc=$?"
  python tools/ncu_summary.py /tmp/r02_s9_$1.ncu-rep gpurun_out/r02_s9_ncu_$1.csv $((1 << $4)) > gpurun_out/r02_s9_ncu_$1.txt 2>&1
  ncu -i /tmp/r02_s9_$1.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
for k in ('l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','smsp__inst_executed_op_shared_ld.sum','launch__shared_mem_per_block_static'):
    print(k, r[h.index(k)] if k in h else 'n/a')
" >> gpurun_out/r02_s9_ncu_$1.txt
  rm -f /tmp/r02_s9_$1.ncu-rep
}
cap epd_eval "EPD(0.05, 0.5, [1.5, 0.5])" eval 22
