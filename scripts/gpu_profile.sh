#!/bin/bash
# Where does the time go? Launch list + `ncu --set full` of the pixel kernel and the grazing-ray kernel,
# on the whole C3 view and on band [430,555) (one wave of rays: one GPU's share out of eight), the
# solid-mask kernels, and a block timeline of the band. Results -> gpurun_out/ (summaries are copied
# to profiles/ by hand, see profiles/README.md).
set -u
mkdir -p gpurun_out
if [ "${1:-}" != "profile-only" ]; then
  echo "== pytest (quick)" && timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c1 or grazing or golden" > gpurun_out/pytest_gpu.log 2>&1
  rc=$?; tail -3 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && exit $rc
fi
BAND="python scripts/exp_configs.py C3 --top 0 --reps 1 --rows 430,555"
FULL="python scripts/exp_configs.py C3 --top 0 --reps 1"
for W in band full; do
  CMD=$FULL; [ $W = band ] && CMD=$BAND
  echo "== launch list ($W)"
  timeout 600 $CMD > gpurun_out/plain_$W.log 2>&1 || { tail -5 gpurun_out/plain_$W.log; exit 1; }
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$W.csv $CMD > gpurun_out/ncu_list_$W.log 2>&1; echo rc=$?
  grep -E "tet_walk|grazing|solid_mask|fill_background|rotate|refit" gpurun_out/launches_$W.csv | tail -8 | cut -d, -f5,13- | cut -c1-160
  echo "== full capture ($W): pixel kernel, grazing-ray kernel"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tet_walk|grazing" -s 4 -c 2 -f -o gpurun_out/walk_$W $CMD > gpurun_out/ncu_$W.log 2>&1; echo rc=$?
done
echo "== full capture: solid_mask (whole view)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solid_mask -s 4 -c 2 -f -o gpurun_out/mask_full $FULL > gpurun_out/ncu_mask.log 2>&1; echo rc=$?
echo "== block timeline of the band, 1 and 2 views in flight"
rm -f gpurun_out/trace_band.txt
C5_TRACE_FILE=gpurun_out/trace_band.txt timeout 600 python scripts/exp_lanes.py C3 --rows "430,555" --lanes 1,2 --views 6 2>&1 | cut -c1-200
python scripts/trace_blocks.py gpurun_out/trace_band.txt 2>/dev/null | head -12 | cut -c1-400
gzip -f gpurun_out/trace_band.txt
exit 0
