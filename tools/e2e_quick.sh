#!/bin/bash
# headline end-to-end leg only: tools/e2e_quick.sh "name|ENV=..|bench args" ...
for spec in "$@"; do
  name=${spec%%|*}; rest=${spec#*|}; envs=${rest%%|*}; extra=${rest#*|}
  [ "$extra" = "$rest" ] && extra=""
  env $(echo $envs | tr ';' ' ') python bench.py --no-cpu --no-roofline --no-e2e-host --steps 4 --warmup 3 $extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('%-16s e2e %.3fM it/s wall %.4f s chunks %d  solve %.1f ms' % ('$name', e['value']/1e6, e['wall_s'], e['chunks'], d['whole_solve']['device_ms']))"
done
