run() { python bench.py --no-e2e --no-cpu --no-roofline --steps 8 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(' value', round(d['value']), 'ms/iter', round(d['ms_per_step'],3), 'whole solve ms', round(d['whole_solve']['device_ms'],1))"; }
echo base; run
for f in variants/*.so; do echo $f; ACOC_LIB=$PWD/$f run; done
echo base again; run
