#!/bin/bash
# Run the short device-resident bench on every tuning/*.so build variant (ACOC_LIB override) and on the default build.
for f in tuning/*.so; do echo "$f"; ACOC_LIB=$PWD/$f python bench.py --no-e2e --no-cpu --steps 6 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), {k:round(v/6,3) for k,v in d['phase_ms'].items()})"; done
echo base; python bench.py --no-e2e --no-cpu --steps 6 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), {k:round(v/6,3) for k,v in d['phase_ms'].items()})"
