#!/bin/bash
# Round-2 evidence of the final build in one GPU call: gpurun -- bash tools/final_r02.sh
O=gpurun_out
timeout 900 bash tools/profile_r02.sh > $O/f_profile.log 2>&1
cp $O/r02_kernel_counters.json profiles/r02_kernel_counters.json   # bench.py reads the counters of THIS build from profiles/
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r02_bench.json 2> $O/f_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_reference_arm.json 2> $O/f_ref.err
timeout 300 python bench.py --precision f32 --steps 20 --warmup 5 --no-cpu > $O/r02_bench_fp32.json 2> $O/f_fp32.err
timeout 300 python bench.py --workload acro --steps 20 --warmup 5 --no-cpu > $O/r02_bench_acro.json 2> $O/f_acro.err
timeout 200 python bench.py --workload track > $O/r02_bench_track.json 2> $O/f_track.err
timeout 200 python bench.py --workload single-step --steps 10 --warmup 3 > $O/r02_bench_single_step.json 2> $O/f_ss.err
timeout 200 python bench.py --workload single-acro --steps 10 --warmup 3 > $O/r02_bench_single_acro.json 2> $O/f_sa.err
timeout 100 python tools/time_small_batch.py 1 2200 4096 > $O/r02_small_batch.json 2>/dev/null
echo final done
