#!/bin/bash
# ncu evidence of the round-2 kernels, captured and summarised ON the GPU box (the .ncu-rep files are too big to travel back):
#   gpurun -- tools/profile_r02.sh        -> gpurun_out/r02_*  (launch list, per-kernel summaries, kernel counters, hottest SASS lines)
set -u
O=gpurun_out
B="python bench.py --no-split --no-e2e --no-cpu --no-roofline --warmup 5 --steps 14"
N="ncu --set full --clock-control none -f"
T=/tmp/acoc_prof; mkdir -p $T
python bench.py --no-e2e --no-cpu --no-roofline --steps 8 --warmup 3 > $O/r02_plain_run.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r02_launches_bench_steps8_warmup3.csv python bench.py --no-e2e --no-cpu --no-roofline --steps 8 --warmup 3 > $T/l.log 2>&1
$N -k regex:k_candidates_list -s 25 -c 1 -o $T/cand $B > $T/c.log 2>&1
$N -k regex:k_rollout_write_tma -s 16 -c 1 -o $T/upd $B > $T/u.log 2>&1
$N -k regex:"k_backward_tma|k_forward_cand0_tma" -s 32 -c 2 -o $T/sweeps $B > $T/s.log 2>&1
$N -k regex:"k_backward_cols|k_search_fused" -s 30 -c 2 -o $T/small python bench.py --workload single-step --no-cpu --steps 10 --warmup 3 > $T/sm.log 2>&1
python profiles/kernel_counters.py $T/sweeps.ncu-rep $T/cand.ncu-rep $T/upd.ncu-rep $T/small.ncu-rep > $O/r02_kernel_counters.json
for r in sweeps cand upd small; do python profiles/summarize_ncu.py $T/$r.ncu-rep > $O/r02_${r}_ncu_summary.txt; done
{ for r in sweeps cand upd small; do n=$(ncu -i $T/$r.ncu-rep --page raw --csv 2>/dev/null | tail -n +3 | wc -l); for i in $(seq 0 $((n-1))); do python tools/ncu_hot.py $T/$r.ncu-rep $i 14; python tools/ncu_mix.py $T/$r.ncu-rep $i 1 | head -16; echo; done; done; } > $O/r02_hot_instructions.txt 2>&1
# the warp-role pipeline of small batches, per role (4096 instances, second Newton iteration)
$N --import-source on -k regex:"k_backward_cols" -s 1 -c 1 -o $T/cols python tools/prof_small.py newton > $T/pc.log 2>&1
ncu -i $T/cols.ncu-rep --page source --csv > $T/cols_source.csv 2>/dev/null
{ python profiles/summarize_ncu.py $T/cols.ncu-rep; python tools/ncu_roles.py $T/cols_source.csv 127872; } > $O/r02_backward_cols_roles.txt 2>&1
ls -la $O/r02_*
