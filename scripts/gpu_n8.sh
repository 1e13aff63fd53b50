#!/bin/bash
# 8 GPUs (charged 8x): bench.py at N = 8, p2p (walk kernels store into rank 0's image over NVLink) with a
# per-rank timeline, then the NCCL send/recv baseline. Each run has its own timeout and watchdog.
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
sumline='import sys,json; d=json.loads(sys.stdin.read()); print("N=%d" % d["n_gpus"], "lanes", d["execution"]["views_in_flight"], round(d["value"]/1e9,2),"G steps/s", round(d["ms_per_step"],3),"ms", "e2e", round(d["e2e"]["value"]/1e9,2), round(d["e2e"]["ms_per_step"],3), "ms", "host enqueue", round(d["execution"]["host_enqueue_ms_per_view"],3), "bands", d.get("bands"), "parity", d.get("parity",{}).get("ok"), "attempts", d.get("attempts")); print("   per rank (Msteps, alone ms, walk, graze, mask):", [(round(p["tet_steps"]/1e6,1), round(p["view_ms_alone"],3), round(p["walk_ms"],3), round(p["graze_ms"],3), round(p["mask_ms"],3)) for p in d["per_rank"]])'
run() { # N gather lanes
  tag=n$1_$2_l$3
  C5_BENCH_VERBOSE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 295$1$3 \
      bench.py --gpus $1 --steps 40 --warmup 5 --gather $2 --lanes $3 --timeline gpurun_out/timeline_$tag.json > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  echo "-- N=$1 gather=$2 lanes=$3 rc=$?"; grep '^{' gpurun_out/bench_$tag.json | python -c "$sumline" || grep "bench rank\|Error\|error" gpurun_out/bench_$tag.err | tail -20
}
run $NG p2p 4
run $NG sendrecv 4
exit 0
