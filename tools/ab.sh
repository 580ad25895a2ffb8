#!/bin/bash
# A/B runs of bench.py (device-resident legs only): tools/ab.sh OUT "name|ENV=..;ENV2=..|extra bench args" ...
# appends the full JSON lines to gpurun_out/OUT.jsonl and prints one compact line per configuration
out=gpurun_out/$1.jsonl; shift
: > $out
for spec in "$@"; do
  name=${spec%%|*}; rest=${spec#*|}; envs=${rest%%|*}; extra=${rest#*|}
  [ "$extra" = "$rest" ] && extra=""
  line=$(env $(echo $envs | tr ';' ' ') python bench.py --no-e2e --no-cpu --steps 20 --warmup 5 $extra 2>gpurun_out/ab_$name.err)
  [ -z "$line" ] && line=null
  echo "{\"name\": \"$name\", \"line\": $line}" >> $out
done
python - $out <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    o = json.loads(ln); d = o["line"]
    if d is None:
        print("%-14s FAILED" % o["name"]); continue
    p = d.get("phase_ms") or {}
    print("%-14s value %.3fM  ms/it %.3f  solve %.1f ms  phases b/f/c/u %s" % (o["name"], d["value"] / 1e6, d["ms_per_step"], d["whole_solve"]["device_ms"],
          "/".join("%.2f" % (p.get(k, 0) / d["steps"]) for k in ("backward", "forward", "candidates", "update"))))
PY
