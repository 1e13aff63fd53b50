#!/bin/bash
# One gpurun call on ONE GPU: parity tests -> smoke -> bench -> (with "ncu") launch list + full captures.
# Each stage runs only if the previous one exited 0 (ncu must never see a faulting program).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
echo "== pytest -m gpu" && timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -5 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && { tail -60 gpurun_out/pytest_gpu.log | cut -c1-300; exit $rc; }
echo "== smoke" && timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; tail -3 gpurun_out/smoke.log; [ $rc -ne 0 ] && exit $rc
echo "== bench" && timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
rc=$?; cut -c1-2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err; [ $rc -ne 0 ] && exit $rc
[ "${1:-}" = "ncu" ] && bash scripts/gpu_profile.sh profile-only
exit 0
