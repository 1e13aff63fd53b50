#!/bin/bash
# ONE GPU, final tree: whole GPU test suite, smoke, bench both arms the way the driver runs them.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu" && timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -3 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && { tail -80 gpurun_out/pytest_gpu.log | cut -c1-300; }
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (reference arm, 3 steps)"
timeout 1500 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?
echo "== bench (ours)"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo rc=$?
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench.json') if l.startswith('{')][0]); r=json.loads([l for l in open('gpurun_out/bench_ref.json') if l.startswith('{')][0])
print('ours value', round(d['value']/1e9,2), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']/1e9,2), round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'], 'frac', round(d['roofline']['frac'],3), 'parity', d['parity']['ok'], 'clocks', d['clocks'])
print('reference', round(r['value']/1e6,2), 'M/s', round(r['ms_per_step'],1), 'ms; e2e ratio', round(d['e2e']['value']/r['value'],1), 'same config', d['config']==r['config'])"
exit 0
