#!/bin/bash
# ONE GPU: grazing-ray kernel (and prologue) on the high-priority stream; run-to-run scatter of the sustained rate.
set -u
mkdir -p gpurun_out
show='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print(d["config"], d["rows"], "lanes", d["lanes"], d.get("debug"), "ms/view", d["ms_per_view"], "G/s", d["Gsteps_per_s"], "reps", d.get("reps_ms"))
    else: print(l.rstrip())
'
rm -f gpurun_out/exp_prio.jsonl
B3="0,0;0,505;505,660;786,905"
for p in 0 2 3; do
  timeout 600 python scripts/exp_lanes.py C3 --rows "$B3" --lanes 4,4,2 --views 32 --debug prep_priority=$p 2>&1 | tee -a gpurun_out/exp_prio.jsonl | python -c "$show"
done
exit 0
