#!/bin/bash
# development tool: compile the stand-alone kernel for a matrix of (model, op, block, min blocks) -> _bin/mb_<model>_<op>_<block>_<minb>
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false --expt-relaxed-constexpr -diag-suppress 20012,20011,20014,177,550 -I../../bbm_b200/csrc -I../../include"
MODELS=${MODELS:-"M_Bagher M_LowMicrofacet M_Ribardiere M_CookTorrance M_AshikhminShirley M_Ward M_Phong"}
SHAPES=${SHAPES:-"256,1 256,2 256,3 128,4 128,6"}
mkdir -p _bin
for m in $MODELS; do for op in 0 1 2; do for s in $SHAPES; do
  b=${s%,*}; k=${s#*,}
  echo "nvcc $FLAGS -DMODEL=$m -DOP=$op -DBLOCK=$b -DMINB=$k ggx_bench.cu -o _bin/mb_${m}_${op}_${b}_${k} 2>/dev/null"
done; done; done | xargs -P 8 -I{} bash -c "{}"
ls _bin/mb_* | wc -l
