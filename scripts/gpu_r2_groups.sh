#!/bin/bash
# N GPUs (charged Nx): view groups. usage: gpu_r2_groups.sh "<bench args of run 1>" "<bench args of run 2>" ...
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG  host cores: $(nproc)"
sumline='import sys,json; d=json.loads(sys.stdin.read()); print("N=%d" % d["n_gpus"], d["config"]["workload"][:4], "groups", d["execution"]["view_groups"], "x", d["execution"]["bands_per_view"], "bands, lanes", d["execution"]["views_in_flight"], "steps", d["steps"], "|", round(d["value"]/1e9,2),"G steps/s", round(d["ms_per_step"],3),"ms | e2e", d["e2e"]["mode"], round(d["e2e"]["value"]/1e9,2), round(d["e2e"]["ms_per_step"],3), "ms | parity", d.get("parity",{}).get("ok"), "rows differ", d.get("parity",{}).get("rows_that_differ"), "bands", d.get("bands"), d["clocks"]["samples"], d["clocks"]["reasons"]); print("   last calibration round:", d["calibration"][-1]["sustained_ms"])'
i=0
for args in "$@"; do
  i=$((i+1)); tag=n${NG}_g$i
  C5_BENCH_VERBOSE=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 2957$i \
      bench.py --gpus $NG --warmup 5 $args > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  echo "-- $tag [$args] rc=$?"; grep '^{' gpurun_out/bench_$tag.json | python -c "$sumline" || { grep -v "OMP_NUM\|\*\*\*" gpurun_out/bench_$tag.err | tail -25 | cut -c1-300; }
done
exit 0
