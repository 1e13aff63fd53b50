#!/bin/bash
# One gpurun call: GPU parity tests -> smoke -> bench -> ncu launch list -> ncu full capture of the walk.
# Each stage runs only if the previous one exited 0 (ncu must never see a faulting program).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
echo "== pytest -m gpu" && timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -15 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && exit $rc
echo "== smoke" && timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; tail -3 gpurun_out/smoke.log; [ $rc -ne 0 ] && exit $rc
echo "== bench" && timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
rc=$?; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err; [ $rc -ne 0 ] && exit $rc
if [ "${1:-}" = "ncu" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
  echo "== ncu launch list"
  timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "ncu list rc=$?"
  echo "== ncu full (walk kernel)"
  timeout 600 $CMD > gpurun_out/plain2.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tet_walk -s 3 -c 1 \
      -f -o gpurun_out/walk_full $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
fi
NG=$(nvidia-smi -L | wc -l)
for N in 2 4 8; do
  if [ "$NG" -ge "$N" ]; then
    echo "== bench N=$N"
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
        bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
    grep '^{' gpurun_out/bench_n$N.json | cut -c1-400; tail -3 gpurun_out/bench_n$N.err
  fi
done
exit 0
