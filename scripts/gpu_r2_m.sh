#!/bin/bash
# HISTORICAL: the walk_smem_kb knob this run used was removed after the measurement (profiles/r02_exp_walk_block_cap_by_smem.jsonl);
# it ran at commit 'Negative result recorded: capping the pixel kernel...'^ and is kept as the record of the command.
# ONE GPU: does the pixel kernel filling every SM keep the other lanes' prologues out? Cap its blocks per SM.
set -u
mkdir -p gpurun_out
show='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print(d["config"], d["rows"], "lanes", d["lanes"], d.get("debug"), "ms/view", d["ms_per_view"], "G/s", d["Gsteps_per_s"], "reps", d.get("reps_ms"))
    else: print(l.rstrip())
'
rm -f gpurun_out/exp_walk_smem.jsonl
for kb in 0 33 40; do
  timeout 600 python scripts/exp_lanes.py C3 --rows "0,0;0,505;505,660;786,905" --lanes 4,2 --views 32 --debug walk_smem_kb=$kb 2>&1 | tee -a gpurun_out/exp_walk_smem.jsonl | python -c "$show"
done
exit 0
