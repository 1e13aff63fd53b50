#!/bin/bash
# GPU experiments: parity tests first, then config/variant sweeps, then (if N>1 visible) a 2-GPU bench.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu" && timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -8 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && exit $rc
echo "== entry lists; software pipelining; register caps (aligned README view, then oblique)"
timeout 900 python scripts/exp_configs.py C3 --variants default,r72,swp,r96 --top 0 --precision 64 2>&1 | tee gpurun_out/exp_c3.jsonl
timeout 900 python scripts/exp_configs.py C3 --variants default --top 0 --precision 32 2>&1 | tee -a gpurun_out/exp_c3.jsonl
timeout 900 python scripts/exp_configs.py C3 --view 0.4,0.3 --variants default,r72,swp --top 0 2>&1 | tee -a gpurun_out/exp_c3.jsonl
echo "== other configs"
timeout 1500 python scripts/exp_configs.py C1 C2 C5 --top 0 2>&1 | tee gpurun_out/exp_configs.jsonl
echo "== bench N=1"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  echo "== bench N=2"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
  cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
fi
exit 0
