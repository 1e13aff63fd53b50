#!/bin/bash
# Scaling run: bench.py at N = 1, 2, 4, 8 on one box (what the driver's SCALE step does), plus the
# in-process multi-device context through the course CLI.
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
grep '^{' gpurun_out/scale_n1.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1', round(d['value']/1e9,2),'G steps/s', round(d['ms_per_step'],3),'ms', 'e2e', round(d['e2e']['value']/1e9,2), d['phases_ms'])"
for N in 2 4 8; do
  if [ "$NG" -ge "$N" ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N \
        bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
    grep '^{' gpurun_out/scale_n$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=$N', round(d['value']/1e9,2),'G steps/s', round(d['ms_per_step'],3),'ms', 'e2e', round(d['e2e']['value']/1e9,2), d['phases_ms'])" || tail -5 gpurun_out/scale_n$N.err
  fi
done
if [ "$NG" -ge 2 ]; then
  echo "== sweep (C4, 48 frames) on $NG GPUs"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 \
      -m course5_b200.sweep --config C4 --frames 48 2> gpurun_out/sweep.err | grep '^{' | tee gpurun_out/sweep.json
  timeout 600 python -m course5_b200.sweep --config C4 --frames 24 2>> gpurun_out/sweep.err | grep '^{' | tee -a gpurun_out/sweep.json
fi
exit 0
