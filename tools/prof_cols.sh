#!/bin/bash
# ncu --set full of k_backward_cols (4096 instances, second Newton iteration) summarised per warp role: gpurun -- bash tools/prof_cols.sh [tag]
T=/tmp/acoc_prof; mkdir -p $T; tag=${1:-cols}
ncu --set full --clock-control none --import-source on -f -k regex:"k_backward_cols" -s 1 -c 1 -o $T/cols python tools/prof_small.py newton > $T/sm.log 2>&1
python profiles/summarize_ncu.py $T/cols.ncu-rep > gpurun_out/${tag}_summary.txt 2>&1
ncu -i $T/cols.ncu-rep --page source --csv > $T/cols_source.csv 2>/dev/null
python tools/ncu_roles.py $T/cols_source.csv > gpurun_out/${tag}_roles.txt 2>&1
python tools/ncu_mix.py $T/cols.ncu-rep 0 127872 > gpurun_out/${tag}_mix.txt 2>&1
tail -2 $T/sm.log
