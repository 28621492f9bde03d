#!/bin/bash
# round 2, GPU session 5: after the table-driven double log (Beckmann sampler) and the EPD hoists: tests incl. the parity scan, throughput, captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_s5_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_s5_pytest.log
tail -15 gpurun_out/r02_s5_pytest.log
python tools/model_throughput.py --log2 24 --out gpurun_out/r02_s5_model_throughput.json > gpurun_out/r02_s5_model_throughput.log 2>&1; echo "model throughput rc=$?"
cap() { # name bsdf op log2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_foreach4 -s 2 -c 1 -f -o /tmp/r02_s5_$1 python tools/run_op.py "$2" $3 $4 > gpurun_out/r02_s5_$1.log 2>&1; echo "ncu $1 rc=$?"
  python tools/ncu_summary.py /tmp/r02_s5_$1.ncu-rep gpurun_out/r02_s5_ncu_$1.csv $((1 << $4)) > gpurun_out/r02_s5_ncu_$1.txt 2>&1
  rm -f /tmp/r02_s5_$1.ncu-rep
}
cap ct_sample "CookTorrance()" sample 22
cap epd_sample "EPD()" sample 22
ls -la gpurun_out | tail -8
