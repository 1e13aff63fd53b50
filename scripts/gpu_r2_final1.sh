#!/bin/bash
# ONE GPU, final state of round 2: whole GPU test suite, smoke, bench the way the driver runs it (both arms),
# launch list of that bench command, ncu --set full of the walk kernels.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu" && timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -9 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && { tail -80 gpurun_out/pytest_gpu.log | cut -c1-300; }
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (ours)"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo rc=$?
cut -c1-600 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
echo "== bench (reference arm)"
timeout 1500 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?
cut -c1-500 gpurun_out/bench_ref.json
echo "== launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench.csv python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo rc=$?
echo "== ncu full: walk kernels, whole view"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tet_walk|grazing" -s 4 -c 2 -f -o gpurun_out/walk_r02 python scripts/exp_configs.py C3 --reps 1 > gpurun_out/ncu_walk.log 2>&1; echo rc=$?
ncu -i gpurun_out/walk_r02.ncu-rep --page raw --csv > gpurun_out/walk_r02_raw.csv 2>/dev/null
ls -la gpurun_out/walk_r02*
exit 0
