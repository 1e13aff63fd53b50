#!/bin/bash
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
sumline='import sys,json; d=json.loads(sys.stdin.read()); print("N=%d" % d["n_gpus"], "steps", d["steps"], round(d["value"]/1e9,2),"G steps/s", round(d["ms_per_step"],3),"ms", "e2e", round(d["e2e"]["ms_per_step"],3), d["clocks"]); print("   last calibration round:", d["calibration"][-1]["sustained_ms"])'
run() { tag=$1; steps=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $NG --steps $steps --warmup 5 "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  echo "-- $tag rc=$?"; grep '^{' gpurun_out/bench_$tag.json | python -c "$sumline" || tail -5 gpurun_out/bench_$tag.err; }
run n${NG}_nosampler 40 --no-clock-sampler
run n${NG}_sampler 40
run n${NG}_nosampler2 40 --no-clock-sampler
exit 0
