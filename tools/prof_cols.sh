T=/tmp/acoc_prof; mkdir -p $T
ncu --set full --clock-control none --import-source on -f -k regex:"k_backward_cols" -s 12 -c 1 -o $T/cols python bench.py --workload single-step --no-cpu --steps 10 --warmup 3 > $T/sm.log 2>&1
python profiles/summarize_ncu.py $T/cols.ncu-rep > gpurun_out/c3_cols_summary.txt 2>&1
python tools/ncu_hot.py $T/cols.ncu-rep 0 90 > gpurun_out/c3_cols_hot.txt 2>&1
python tools/ncu_mix.py $T/cols.ncu-rep 0 999 > gpurun_out/c3_cols_mix.txt 2>&1
ncu -i $T/cols.ncu-rep --page source --csv > gpurun_out/c3_cols_source.csv 2>/dev/null
ls -la $T
