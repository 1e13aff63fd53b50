#!/bin/bash
# Scaling run: bench.py at N = 1, 2, 4, 8 on one box, the way the driver launches it (default flags:
# two views in flight, ncclSend/Recv gather, shared pinned host image for e2e; each N > 1 run falls
# back to one view in flight + gather-and-copy e2e by itself if the first attempt stalls).
# Cost: N x the box time — at 8 GPUs a minute of wall clock is eight GPU-minutes.
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
sumline='import sys,json; d=json.loads(sys.stdin.read()); print("N=%d" % d["n_gpus"], "lanes", d["config"]["views_in_flight"], round(d["value"]/1e9,2),"G steps/s", round(d["ms_per_step"],3),"ms", "e2e", round(d["e2e"]["value"]/1e9,2), round(d["e2e"]["ms_per_step"],3), "ms", {k: round(v,3) for k,v in d["phases_ms"].items()}, "frac", round(d["roofline"]["frac"],3), "bands", d.get("bands"), "attempts", d.get("attempts"))'
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
grep '^{' gpurun_out/scale_n1.json | python -c "$sumline" || tail -5 gpurun_out/scale_n1.err
for N in 2 4 8; do
  if [ "$NG" -ge "$N" ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N \
        bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
    grep '^{' gpurun_out/scale_n$N.json | python -c "$sumline" || grep -v "OMP_NUM\|\*\*\*" gpurun_out/scale_n$N.err | tail -12
  fi
done
exit 0
