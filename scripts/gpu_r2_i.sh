#!/bin/bash
# ONE GPU: software-pipelined step loop (sp5/sp6/sp7) against the product kernel; does the NVML poller cost anything;
# ncu of the cooperative mask kernels.
set -u
mkdir -p gpurun_out
show='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print(d["config"], d["view"], d["rows"], d["variant"], "walk", d["ms_walk"], "graze", d["ms_graze"], "mask", d["ms_mask"], "total", d["ms_total"], "G/s", d["walk_Gsteps_per_s"])
    else: print(l.rstrip())
'
rm -f gpurun_out/exp_sp.jsonl
echo "== C3 README view: whole, central band"
C5GPU_LIBRARY=build/exp/libc5gpu_exp.so timeout 900 python scripts/exp_configs.py C3 --reps 5 --rows "0,1800;828,911" --variants default,sp5,sp6,sp7 2>&1 | tee -a gpurun_out/exp_sp.jsonl | python -c "$show"
echo "== C3 oblique"
C5GPU_LIBRARY=build/exp/libc5gpu_exp.so timeout 900 python scripts/exp_configs.py C3 --reps 5 --view 0.4,0.3 --variants default,sp5,sp6,sp7 2>&1 | tee -a gpurun_out/exp_sp.jsonl | python -c "$show"
echo "== C5t (50M tets, 2400x1800)"
C5GPU_LIBRARY=build/exp/libc5gpu_exp.so timeout 900 python scripts/exp_configs.py C5t --reps 3 --variants default,sp5,sp6 2>&1 | tee -a gpurun_out/exp_sp.jsonl | python -c "$show"
for mode in "" "--no-clock-sampler"; do
  echo "== bench N=1 40 steps $mode"
  timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline $mode > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; rc=$?
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_ab.json') if l.startswith('{')][0]); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['clocks'])"
done
echo "== ncu full: mask kernels (cooperative), whole view"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"solid_mask|mask_tile" -s 6 -c 6 -f -o gpurun_out/mask_r02_coop python scripts/exp_configs.py C3 --reps 1 > gpurun_out/ncu_mask.log 2>&1; echo rc=$?
ncu -i gpurun_out/mask_r02_coop.ncu-rep --page raw --csv > gpurun_out/mask_r02_coop_raw.csv 2>/dev/null
exit 0
