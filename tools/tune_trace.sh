#!/bin/bash
# Full-solve trace (tools/solve_trace.py) for every tuning/*.so build variant: solve time and the candidates phase of the first iterations.
for f in tuning/*.so; do echo "$f"; ACOC_LIB=$PWD/$f python tools/solve_trace.py 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(' solve_ms', round(d['e2e_rep1']['solve_device_ms'],1), 'iters', d['e2e_rep1']['iters']); print(' cand phase it0,1,14,15,16:', [d['phases_bwd_fwd_cand_upd'][k][2] for k in (0,1,14,15,16)])"; done
