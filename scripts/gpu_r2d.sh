#!/bin/bash
# 2 GPUs: parity (incl. the two-process image tests), bands with the rect grid + concurrent grazing
# kernel, bench at N = 1 and N = 2 (stores over NVLink vs ncclSend/Recv gather).
set -u
mkdir -p gpurun_out
if [ "${1:-}" != "skiptests" ]; then timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; else rc=0; fi
tail -3 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && exit $rc
show='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print(d["config"], d["view"], d["precision"], d["variant"], d["rows"], "walk", d["ms_walk"], "mask", d["ms_mask"], "total", d["ms_total"], "G/s", d["walk_Gsteps_per_s"], "graze", d["grazing_rays"])
    else: print(l.rstrip())
'
timeout 900 python scripts/exp_configs.py C3 --top 0 --reps 5 --rows "0,1800;430,555;0,400;430,680;800,925" 2>&1 | tee gpurun_out/exp_bands.jsonl | python -c "$show"
timeout 900 python scripts/exp_configs.py C3 --top 0 --reps 5 --view 0.4,0.3 --precision 64,32 2>&1 | tee gpurun_out/exp_c3.jsonl | python -c "$show"
timeout 900 python scripts/exp_configs.py C1 C2 C5 --top 0 2>&1 | tee gpurun_out/exp_configs.jsonl | python -c "$show"
sumline='import sys,json; d=json.loads(sys.stdin.read()); print("N=%d" % d["n_gpus"], round(d["value"]/1e9,2),"G steps/s", round(d["ms_per_step"],3),"ms", "e2e", round(d["e2e"]["value"]/1e9,2), round(d["e2e"]["ms_per_step"],3), "ms", d["phases_ms"], "frac", round(d["roofline"]["frac"],3))'
echo "== bench N=1"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; grep '^{' gpurun_out/bench_n1.json | python -c "$sumline"; tail -3 gpurun_out/bench_n1.err
for G in p2p sendrecv; do
  echo "== bench N=2 gather=$G"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 10 --warmup 3 --gather $G > gpurun_out/bench_n2_$G.json 2> gpurun_out/bench_n2_$G.err
  grep '^{' gpurun_out/bench_n2_$G.json | python -c "$sumline"; tail -5 gpurun_out/bench_n2_$G.err
done
exit 0
