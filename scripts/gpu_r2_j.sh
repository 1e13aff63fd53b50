#!/bin/bash
# ONE GPU: what each of the 8 bands of the N = 8 run sustains alone with 2..8 views in flight (is a rank bound by
# throughput or by the length of one view's chain?), with and without the high-priority prologue stream; the same for
# C5t; bench at N = 1 for C5t; the C4 sweep on one GPU; mask kernels after the write-only spans / face-id order.
set -u
mkdir -p gpurun_out
show='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print(d["config"], d["rows"], "lanes", d["lanes"], d.get("debug"), "ms/view", d["ms_per_view"], "G/s", d["Gsteps_per_s"], "alone", d["one_view_ms_total"], "graze", d["one_view_ms_graze"])
    else: print(l.rstrip())
'
rm -f gpurun_out/exp_lanes_bands.jsonl
B3="0,505;505,660;660,786;786,905;1163,1301;1301,1800"
timeout 600 python scripts/exp_lanes.py C3 --rows "$B3" --lanes 2,4,6,8 --views 32 2>&1 | tee -a gpurun_out/exp_lanes_bands.jsonl | python -c "$show"
timeout 600 python scripts/exp_lanes.py C3 --rows "0,505;505,660;786,905" --lanes 4,8 --views 32 --debug prep_priority=1 2>&1 | tee -a gpurun_out/exp_lanes_bands.jsonl | python -c "$show"
B5="0,0;0,535;535,665;806,967;1117,1226;1329,1800"
timeout 600 python scripts/exp_lanes.py C5t --rows "$B5" --lanes 2,4,8 --views 24 2>&1 | tee -a gpurun_out/exp_lanes_bands.jsonl | python -c "$show"
echo "== mask after write-only spans + face-id order"
timeout 600 python scripts/exp_configs.py C3 --reps 5 --rows "0,1800;828,911;702,828;534,702" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['config'], d['rows'], 'mask', d['ms_mask'], 'walk', d['ms_walk'], 'total', d['ms_total'])
"
echo "== pytest mask"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mask or golden or c1 or c2 or static" 2>&1 | tail -2
echo "== bench N=1 C5t"
timeout 900 python bench.py --steps 20 --warmup 5 --workload C5t --no-cpu-baseline > gpurun_out/bench_c5t_n1.json 2> gpurun_out/bench_c5t_n1.err; echo rc=$?
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_c5t_n1.json') if l.startswith('{')][0]); print('C5t N=1 ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['phases_ms'], 'frac', d['roofline']['frac'])"
echo "== C4 sweep 360 views, one GPU"
timeout 600 python -m course5_b200.sweep --config C4 --frames 360 > gpurun_out/sweep_c4_n1.json 2> gpurun_out/sweep_c4_n1.err; echo rc=$?; cut -c1-400 gpurun_out/sweep_c4_n1.json
exit 0
