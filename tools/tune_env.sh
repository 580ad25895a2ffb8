#!/bin/bash
# Device-resident bench and full-solve time for several values of an environment tuning knob:  tools/tune_env.sh ACOC_SUB 1 2 3 4
var=$1; shift
for v in "$@"; do echo "$var=$v"; env $var=$v python bench.py --no-e2e --no-cpu --no-roofline --steps 8 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(' value', round(d['value']), 'ms/iter', round(d['ms_per_step'],3), 'whole solve ms', round(d['whole_solve']['device_ms'],1))"; done
