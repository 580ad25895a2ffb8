#!/bin/bash
# end-to-end wall of the pipelined solve for different sub-batch counts: tools/e2e_chunks.sh OUT chunks...
out=gpurun_out/$1.txt; shift
: > $out
for c in "$@"; do
  line=$(python bench.py --no-cpu --no-roofline --steps 4 --warmup 3 --chunks $c 2>/dev/null)
  python - "$c" >> $out <<PY
import json,sys
d=json.loads('''$line''')
e=d["e2e"]; h=d.get("e2e_host_refs") or {}
print("chunks %s: e2e generated %.4f s (%.3f M it/s)  host-refs %.4f s  whole solve %.1f ms" % (sys.argv[1], e["wall_s"], e["value"]/1e6, h.get("wall_s", float("nan")), d["whole_solve"]["device_ms"]))
PY
done
cat $out
