#!/bin/bash
# 8 GPUs (charged 8x): bench.py at N = 8 the way the driver launches it (20 steps) and with 40 steps + timeline,
# the north_star target workload (C5t: 50M tets at 2400x1800), and the C4 sweep (360 views) dealt over 8 GPUs.
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG  host cores: $(nproc)"
sumline='import sys,json; d=json.loads(sys.stdin.read()); print("N=%d" % d["n_gpus"], d["config"]["workload"][:4], "steps", d["steps"], "lanes", d["execution"]["views_in_flight"], round(d["value"]/1e9,2),"G steps/s", round(d["ms_per_step"],3),"ms", "e2e", round(d["e2e"]["value"]/1e9,2), round(d["e2e"]["ms_per_step"],3), "ms", "host enqueue", round(d["execution"]["host_enqueue_ms_per_view"],3), "bands", d.get("bands"), "parity", d.get("parity",{}).get("ok"), "clocks", d["clocks"], "attempts", d.get("attempts")); print("   per rank (Msteps, alone ms, walk, graze, mask):", [(round(p["tet_steps"]/1e6,1), round(p["view_ms_alone"],3), round(p["walk_ms"],3), round(p["graze_ms"],3), round(p["mask_ms"],3)) for p in d["per_rank"]])'
run() { # tag steps extra...
  tag=$1; steps=$2; shift 2
  C5_BENCH_VERBOSE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29577 \
      bench.py --gpus $NG --steps $steps --warmup 5 "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  echo "-- $tag rc=$?"; grep '^{' gpurun_out/bench_$tag.json | python -c "$sumline" || grep "bench rank\|Error\|error" gpurun_out/bench_$tag.err | tail -20
}
run n${NG}_c3_s40 40 --timeline gpurun_out/timeline_n${NG}_c3_s40.json
run n${NG}_c3_s20 20
run n${NG}_c5t_s20 20 --workload C5t --timeline gpurun_out/timeline_n${NG}_c5t_s20.json
echo "== C4 sweep, 360 views over $NG GPUs"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29578 -m course5_b200.sweep --config C4 --frames 360 > gpurun_out/sweep_c4_n${NG}.json 2> gpurun_out/sweep_c4_n${NG}.err
echo "rc=$?"; grep '^{' gpurun_out/sweep_c4_n${NG}.json | cut -c1-600
exit 0
