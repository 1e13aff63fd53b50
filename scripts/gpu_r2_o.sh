#!/bin/bash
# ONE GPU: the software-pipelined step loop in 1/8 bands with four views in flight (a band is one wave: does the
# shorter chain per ray pay when occupancy is not the limit?)
set -u
mkdir -p gpurun_out
show='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print(d["config"], d["rows"], "lanes", d["lanes"], "ms/view", d["ms_per_view"], "G/s", d["Gsteps_per_s"], "reps", d.get("reps_ms"))
    else: print(l.rstrip())
'
rm -f gpurun_out/exp_sp_bands.jsonl
for var in default sp5 sp6; do
  echo "== variant $var"
  C5_WALK_VARIANT=$var C5GPU_LIBRARY=build/exp/libc5gpu_exp.so timeout 600 python scripts/exp_lanes.py C3 --rows "0,505;505,660;649,782;786,905" --lanes 4 --views 32 2>&1 | sed "s/^{/{\"variant\": \"$var\", /" | tee -a gpurun_out/exp_sp_bands.jsonl | python -c "$show"
done
exit 0
