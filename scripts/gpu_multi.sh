#!/bin/bash
# N GPUs visible (use --gpus 2 first: N-GPU box time is charged N times): the two-process image test,
# then bench.py at N = all visible GPUs with the three image-assembly variants. Every bench call has
# its own timeout; bench.py's watchdog ends a stalled rank after 90 s.
set -u
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
echo "== pytest dist" && timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/pytest_gpu_dist.log 2>&1
rc=$?; tail -3 gpurun_out/pytest_gpu_dist.log | cut -c1-1500; [ $rc -ne 0 ] && { grep -n "AssertionError\|rows that differ\|Error" gpurun_out/pytest_gpu_dist.log | cut -c1-1500 | head -20; exit $rc; }
sumline='import sys,json; d=json.loads(sys.stdin.read()); print("N=%d" % d["n_gpus"], "lanes", d["config"]["views_in_flight"], round(d["value"]/1e9,2),"G steps/s", round(d["ms_per_step"],3),"ms", "e2e", round(d["e2e"]["value"]/1e9,2), round(d["e2e"]["ms_per_step"],3), "ms", {k: round(v,3) for k,v in d["phases_ms"].items()}, "frac", round(d["roofline"]["frac"],3), "bands", d.get("bands"), "attempts", d.get("attempts"))'
run() { # N gather lanes
  C5_BENCH_VERBOSE=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 295$1$3 \
      bench.py --gpus $1 --steps 20 --warmup 3 --gather $2 --lanes $3 > gpurun_out/bench_n$1_$2_l$3.json 2> gpurun_out/bench_n$1_$2_l$3.err
  echo "-- N=$1 gather=$2 lanes=$3 rc=$?"; grep '^{' gpurun_out/bench_n$1_$2_l$3.json | python -c "$sumline" || grep "bench rank" gpurun_out/bench_n$1_$2_l$3.err | tail -12
}
run $NG sendrecv 2
run $NG sendrecv 1
run $NG p2p 2
exit 0
