#!/bin/bash
# Where does a 125-row band's time go? Launch list + full captures of the pixel kernel and the
# grazing-ray kernel on band [430,555) of the C3 README view, then the same for the whole image.
set -u
mkdir -p gpurun_out
echo "== pytest (mask + grazing only)" && timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c1 or grazing or golden or solid" > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -3 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && exit $rc
show='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print(d["config"], d["view"], d["precision"], d["variant"], "pf", d["prefetch"], d["rows"], "walk", d["ms_walk"], "mask", d["ms_mask"], "total", d["ms_total"], "G/s", d["walk_Gsteps_per_s"], "graze", d["grazing_rays"])
    else: print(l.rstrip())
'
timeout 900 python scripts/exp_configs.py C3 --top 0 --reps 5 --prefetch 0 --variants default,b64,r80 --rows "0,1800;430,555;0,400;430,680" 2>&1 | tee gpurun_out/exp_bands.jsonl | python -c "$show"
BAND="python scripts/exp_configs.py C3 --top 0 --reps 1 --prefetch 0 --rows 430,555"
FULL="python scripts/exp_configs.py C3 --top 0 --reps 1 --prefetch 0"
echo "== launch list (band)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_band.csv $BAND > gpurun_out/ncu_list_band.log 2>&1; echo rc=$?
grep -E "tet_walk|grazing|solid_mask|classify|rotate|refit" gpurun_out/launches_band.csv | tail -12 | cut -d, -f5,13- | cut -c1-160
echo "== full capture (band): pixel kernel, grazing kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tet_walk|grazing" -s 4 -c 2 -f -o gpurun_out/walk_band2 $BAND > gpurun_out/ncu_band2.log 2>&1; echo rc=$?
echo "== launch list (whole image)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_full.csv $FULL > gpurun_out/ncu_list_full.log 2>&1; echo rc=$?
grep -E "tet_walk|grazing|solid_mask|classify|rotate|refit" gpurun_out/launches_full.csv | tail -12 | cut -d, -f5,13- | cut -c1-160
echo "== full capture (whole image)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tet_walk|grazing" -s 4 -c 2 -f -o gpurun_out/walk_full2 $FULL > gpurun_out/ncu_full2.log 2>&1; echo rc=$?
exit 0
