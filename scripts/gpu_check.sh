#!/bin/bash
# One gpurun call on ONE GPU: parity tests -> smoke -> bench (with the cpu_baseline + parity leg).
# Each stage runs only if the previous one exited 0.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
nproc > gpurun_out/host_cores.txt; free -g >> gpurun_out/host_cores.txt
echo "== pytest -m gpu" && timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -14 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && { tail -80 gpurun_out/pytest_gpu.log | cut -c1-300; exit $rc; }
echo "== smoke" && timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; tail -3 gpurun_out/smoke.log; [ $rc -ne 0 ] && exit $rc
echo "== bench" && timeout 1200 python bench.py --steps 20 --warmup 5 --timeline gpurun_out/timeline_n1.json > gpurun_out/bench.json 2> gpurun_out/bench.err
rc=$?; cut -c1-6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err; [ $rc -ne 0 ] && exit $rc
exit 0
