#!/bin/bash
# N GPUs visible (use --gpus 2 first: N-GPU box time is charged N times): the two-process image test,
# then bench.py at N = all visible GPUs: p2p (walk kernels store into rank 0's image) with a timeline,
# then the NCCL send/recv baseline. Every bench call has its own timeout; bench.py's watchdog ends a
# stalled rank after 90 s.
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
echo "== pytest dist" && timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/pytest_gpu_dist.log 2>&1
rc=$?; tail -3 gpurun_out/pytest_gpu_dist.log | cut -c1-1500; [ $rc -ne 0 ] && { grep -n "AssertionError\|rows that differ\|Error" gpurun_out/pytest_gpu_dist.log | cut -c1-1500 | head -20; exit $rc; }
sumline='import sys,json; d=json.loads(sys.stdin.read()); print("N=%d" % d["n_gpus"], "lanes", d["execution"]["views_in_flight"], round(d["value"]/1e9,2),"G steps/s", round(d["ms_per_step"],3),"ms", "e2e", round(d["e2e"]["value"]/1e9,2), round(d["e2e"]["ms_per_step"],3), "ms", "host enqueue", round(d["execution"]["host_enqueue_ms_per_view"],3), {k: round(v,3) for k,v in d["phases_ms"].items()}, "frac", round(d["roofline"]["frac"],3), "bands", d.get("bands"), "parity", d.get("parity",{}).get("ok"), "attempts", d.get("attempts")); print("   per rank:", [(round(p["tet_steps"]/1e6,1), round(p["view_ms_alone"],3)) for p in d["per_rank"]])'
run() { # N gather lanes [extra flags]
  tag=n$1_$2_l$3
  C5_BENCH_VERBOSE=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 295$1$3 \
      bench.py --gpus $1 --steps 40 --warmup 8 --gather $2 --lanes $3 --timeline gpurun_out/timeline_$tag.json ${4:-} > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  echo "-- N=$1 gather=$2 lanes=$3 rc=$?"; grep '^{' gpurun_out/bench_$tag.json | python -c "$sumline" || grep "bench rank\|Error\|error" gpurun_out/bench_$tag.err | tail -12
}
run $NG p2p 4
run $NG p2p 2
run $NG sendrecv 4
exit 0
