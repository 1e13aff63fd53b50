#!/bin/bash
# Why is the axis-aligned C3 view 2x slower per tet-step than an oblique one? Two full ncu captures.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu" && timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -5 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && exit $rc
echo "== fp64 vs fp32, two views"
timeout 900 python scripts/exp_configs.py C3 --precision 64,32 2>&1 | tee gpurun_out/exp_prec.jsonl
timeout 900 python scripts/exp_configs.py C3 --view 0.4,0.3 --precision 64,32 2>&1 | tee -a gpurun_out/exp_prec.jsonl
A="python scripts/exp_configs.py C3 --reps 1 --no-solids"
B="python scripts/exp_configs.py C3 --view 0.4,0.3 --reps 1 --no-solids"
echo "== ncu aligned view"
timeout 600 $A > gpurun_out/plainA.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tet_walk_fp64 -s 2 -c 1 -f -o gpurun_out/walk_aligned $A > gpurun_out/ncuA.log 2>&1
echo "rc=$?"
echo "== ncu oblique view"
timeout 600 $B > gpurun_out/plainB.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tet_walk_fp64 -s 2 -c 1 -f -o gpurun_out/walk_oblique $B > gpurun_out/ncuB.log 2>&1
echo "rc=$?"
exit 0
