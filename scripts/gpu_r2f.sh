#!/bin/bash
# Two views in flight (context + sibling on two streams): parity, then bench N=1 with 1 and 2 lanes.
set -u
mkdir -p gpurun_out
echo "== pytest parity" && timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -3 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && { tail -60 gpurun_out/pytest_gpu.log | cut -c1-300; exit $rc; }
sumline='import sys,json; d=json.loads(sys.stdin.read()); print("N=%d" % d["n_gpus"], "lanes", d["config"]["views_in_flight"], round(d["value"]/1e9,2),"G steps/s", round(d["ms_per_step"],3),"ms", "e2e", round(d["e2e"]["value"]/1e9,2), round(d["e2e"]["ms_per_step"],3), "ms", {k: round(v,3) for k,v in d["phases_ms"].items()}, "frac", round(d["roofline"]["frac"],3), "launches", d["gpu_launches"])'
for L in 1 2; do
  echo "== bench N=1 lanes=$L"
  timeout 900 python bench.py --steps 20 --warmup 3 --lanes $L --no-cpu-baseline > gpurun_out/bench_n1_l$L.json 2> gpurun_out/bench_n1_l$L.err; grep '^{' gpurun_out/bench_n1_l$L.json | python -c "$sumline"; tail -3 gpurun_out/bench_n1_l$L.err
done
show='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print(d["config"], d["view"], d["precision"], d["variant"], d["rows"], "walk", d["ms_walk"], "mask", d["ms_mask"], "total", d["ms_total"], "G/s", d["walk_Gsteps_per_s"], "graze", d["grazing_rays"])
    else: print(l.rstrip())
'
timeout 900 python scripts/exp_configs.py C3 --top 0 --reps 5 --rows "0,1800;430,555" 2>&1 | tee gpurun_out/exp_bands.jsonl | python -c "$show"
echo "== bench N=1 default (with cpu baseline)"
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; grep '^{' gpurun_out/bench_n1.json | python -c "$sumline"; tail -3 gpurun_out/bench_n1.err
exit 0
