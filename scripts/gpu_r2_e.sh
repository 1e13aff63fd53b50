#!/bin/bash
# HISTORICAL: the mask_lane_shift knob belonged to the per-face mask kernels, which now live only in the experiments
# build (c5_debug_set "mask_per_face"); kept as the record of the command behind profiles/r02_exp_mask_tile_skip_lanes.jsonl.
# ONE GPU: solid mask passes — lanes per tall face, per-kernel times (ncu launch list).
set -u
mkdir -p gpurun_out
show='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print(d["config"], d["view"], d["rows"], "walk", d["ms_walk"], "graze", d["ms_graze"], "mask", d["ms_mask"], "total", d["ms_total"])
    else: print(l.rstrip())
'
echo "== pytest mask" && timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mask or golden or c1 or c2" > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -3 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && { tail -60 gpurun_out/pytest_gpu.log | cut -c1-300; exit $rc; }
rm -f gpurun_out/exp_mask_lanes.jsonl
for ls in 0 2 4 6; do
  echo "== mask_lane_shift=$ls"
  timeout 600 python scripts/exp_configs.py C3 --reps 5 --rows "0,1800;828,911;702,828" --debug mask_lane_shift=$ls 2>&1 | tee -a gpurun_out/exp_mask_lanes.jsonl | python -c "$show"
done
timeout 600 python scripts/exp_configs.py C3 --reps 5 --view 0.4,0.3 2>&1 | tee -a gpurun_out/exp_mask_lanes.jsonl | python -c "$show"
echo "== per-kernel times (ncu launch list; cold-cache, serialised)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_mask.csv python scripts/exp_configs.py C3 --reps 1 --rows "0,1800;828,911" > gpurun_out/ncu_list.log 2>&1; echo rc=$?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_mask.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
seq=[(r[ki][:40], float(r[vi].replace(",",""))/1000.0) for r in rows[1:]]
for name,us in seq[-48:]: print("%-42s %9.1f us" % (name, us))
PY
exit 0
