#!/bin/bash
# N GPUs (charged Nx): bench.py the way the driver launches it.
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG  host cores: $(nproc)"
sumline='import sys,json; d=json.loads(sys.stdin.read()); print("N=%d" % d["n_gpus"], d["config"]["workload"][:4], "steps", d["steps"], "lanes", d["execution"]["views_in_flight"], round(d["value"]/1e9,2),"G steps/s", round(d["ms_per_step"],3),"ms", "e2e", d["e2e"]["mode"], round(d["e2e"]["value"]/1e9,2), round(d["e2e"]["ms_per_step"],3), "ms", "host enqueue", round(d["execution"]["host_enqueue_ms_per_view"],3), "bands", d.get("bands"), "parity", d.get("parity",{}).get("ok"), "clocks", d["clocks"], "attempts", d.get("attempts")); print("   per rank (Msteps, alone ms, walk, graze, mask):", [(round(p["tet_steps"]/1e6,1), round(p["view_ms_alone"],3), round(p["walk_ms"],3), round(p["graze_ms"],3), round(p["mask_ms"],3)) for p in d["per_rank"]]); print("   calibration:", d.get("calibration"))'
run() { # tag steps extra...
  tag=$1; steps=$2; shift 2
  C5_BENCH_VERBOSE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29577 \
      bench.py --gpus $NG --steps $steps --warmup 5 "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  echo "-- $tag rc=$?"; grep '^{' gpurun_out/bench_$tag.json | python -c "$sumline" || grep "bench rank\|Error\|error" gpurun_out/bench_$tag.err | tail -20
}
run n${NG}_c3_s20 20 --timeline gpurun_out/timeline_n${NG}_c3_s20.json
exit 0
