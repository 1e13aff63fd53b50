#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest parity (1 GPU)" && timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -5 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && { tail -60 gpurun_out/pytest_gpu.log | cut -c1-300; exit $rc; }
echo "== pytest dist (2 GPUs)" && timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/pytest_gpu_dist.log 2>&1
rc=$?; tail -5 gpurun_out/pytest_gpu_dist.log | cut -c1-1500; [ $rc -ne 0 ] && { grep -n "AssertionError\|rows that differ" gpurun_out/pytest_gpu_dist.log | cut -c1-1500; exit $rc; }
bash scripts/gpu_r2d.sh skiptests
