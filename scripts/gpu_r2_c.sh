#!/bin/bash
# ONE GPU: grazing-ray kernel grid, e2e in place against copy engine, lanes.
set -u
mkdir -p gpurun_out
for gb in 4 8 12 16; do
  timeout 300 python scripts/exp_lanes.py C3 --rows "0,0;430,555" --lanes 1,4 --views 24 --debug graze_blocks=$gb 2>&1 | tee -a gpurun_out/exp_graze_blocks.jsonl | cut -c1-300
done
for mode in inplace copy; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-mode $mode > gpurun_out/bench_e2e_$mode.json 2> gpurun_out/bench_e2e_$mode.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_e2e_$mode.json').read().strip().splitlines()[-1])
print('$mode', 'value ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3))" || tail -5 gpurun_out/bench_e2e_$mode.err
done
exit 0
