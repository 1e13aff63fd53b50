#!/bin/bash
# ONE GPU: tile size of the solid mask's flags; ncu of all mask kernels.
set -u
mkdir -p gpurun_out
show='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print(d["config"], d["view"], d["rows"], "mask", d["ms_mask"], "total", d["ms_total"])
    else: print(l.rstrip())
'
echo "== pytest mask" && timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mask or golden or c1 or c2" > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -3 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && { tail -60 gpurun_out/pytest_gpu.log | cut -c1-300; exit $rc; }
rm -f gpurun_out/exp_mask_tile.jsonl
for t in 1608 1604 1602 802 801 1601 3201; do
  echo "== mask_tile=$t"
  timeout 600 python scripts/exp_configs.py C3 --reps 5 --rows "0,1800;828,911;702,828" --debug mask_tile=$t 2>&1 | tee -a gpurun_out/exp_mask_tile.jsonl | python -c "$show"
done
echo "== ncu full: all mask kernels, whole view (default tile)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"solid_mask|mask_tile" -s 15 -c 5 -f -o gpurun_out/mask_r02 python scripts/exp_configs.py C3 --reps 1 > gpurun_out/ncu_mask.log 2>&1; echo rc=$?
ncu -i gpurun_out/mask_r02.ncu-rep --page raw --csv > gpurun_out/mask_r02_raw.csv 2>/dev/null
exit 0
