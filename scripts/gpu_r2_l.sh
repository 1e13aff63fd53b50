#!/bin/bash
# ONE GPU: views in flight at N = 1 (device-timed value and e2e), 2 / 3 / 4.
set -u
mkdir -p gpurun_out
for L in 2 3 4 2 4; do
  timeout 600 python bench.py --steps 20 --warmup 5 --lanes $L --no-cpu-baseline > gpurun_out/bench_l$L.json 2> gpurun_out/bench_l$L.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_l$L.json') if l.startswith('{')][0]); print('lanes', $L, 'ms_per_step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['e2e']['mode'], d['clocks']['samples'])"
done
timeout 600 python bench.py --steps 20 --warmup 5 --lanes 2 --no-cpu-baseline --e2e-mode copy > gpurun_out/bench_l2c.json 2> gpurun_out/bench_l2c.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_l2c.json') if l.startswith('{')][0]); print('lanes 2 copy', 'ms_per_step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['e2e']['mode'])"
exit 0
