timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_search or two_range or newton_configs or tma_rings" 2>&1 | tail -12
run() { python bench.py --no-e2e --no-cpu --no-roofline --steps 8 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(' value', round(d['value']), 'ms/iter', round(d['ms_per_step'],3), 'whole solve ms', round(d['whole_solve']['device_ms'],1))"; }
echo fused; run
echo separate; ACOC_NO_FWD_CAND0=1 run
echo fused; run
echo separate; ACOC_NO_FWD_CAND0=1 run
echo "fused, no split"; python bench.py --no-split --no-e2e --no-cpu --no-roofline --steps 8 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(' value', round(d['value']), 'ms/iter', round(d['ms_per_step'],3), 'whole solve ms', round(d['whole_solve']['device_ms'],1))"
