#!/bin/bash
# round 2, GPU session 4: source-level captures of the slow kernel families (before-state of VERDICT item 6);
# the reports are summarised on the box (tools/ncu_summary.py) and deleted: only the summaries travel back
mkdir -p gpurun_out
cap() { # name bsdf op log2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_foreach4 -s 2 -c 1 -f -o /tmp/r02_s4_$1 python tools/run_op.py "$2" $3 $4 > gpurun_out/r02_s4_$1.log 2>&1; echo "ncu $1 rc=$?"
  python tools/ncu_summary.py /tmp/r02_s4_$1.ncu-rep gpurun_out/r02_s4_ncu_$1.csv $((1 << $4)) > gpurun_out/r02_s4_ncu_$1.txt 2>&1
  rm -f /tmp/r02_s4_$1.ncu-rep
}
cap ct_sample "CookTorrance()" sample 22
cap epd_sample "EPD()" sample 22
cap bagher_eval "Bagher()" eval 22
cap ribardiere_eval "Ribardiere()" eval 22
cap lowmf_eval "LowMicrofacet()" eval 22
cap asfull_sample "AshikhminShirleyFull()" sample 22
ls -la gpurun_out | tail -14
