run() { python bench.py --no-cpu --no-roofline --steps 8 --warmup 3 $1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(' e2e wall', round(d['e2e']['wall_s'],4), 'e2e it/s', round(d['e2e']['value']))"; }
for c in 4 2 3 5 4; do echo "chunks=$c"; run "--chunks $c"; done
