#!/bin/bash
# Grazing-ray kernel: parity first (fail fast), sanitizer on a small grazing case, then row bands.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu" && timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
rc=$?; tail -8 gpurun_out/pytest_gpu.log; [ $rc -ne 0 ] && exit $rc
echo "== compute-sanitizer memcheck + racecheck on a grazing view"
cat > /tmp/graze_small.py <<'PY'
import sys; sys.path.insert(0, '.')
from course5_b200 import api, synth
mesh = synth.kuhn_cube(16, seed=7)
with api.Context(devices=(0,)) as ctx:
    ctx.upload_mesh(mesh.points, mesh.tets, mesh.alpha, mesh.q)
    img = ctx.render_raw(api.make_view(320, 240, X=0.5, Y=0.0))
    print(img.stats)
PY
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/graze_small.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python /tmp/graze_small.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/racecheck.log
echo "== bands (C3 README view)"
timeout 900 python scripts/exp_configs.py C3 --top 0 --reps 5 --variants default,r72 --rows "0,1800;0,400;430,555;800,925;395,405" 2>&1 | tee gpurun_out/exp_bands.jsonl | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['variant'], d['rows'], 'walk', d['ms_walk'], 'mask', d['ms_mask'], 'steps M', round(d['tet_steps']/1e6,1), 'G/s', d['walk_Gsteps_per_s'])
    else: print(l.rstrip())
"
echo "== oblique + fp32"
timeout 900 python scripts/exp_configs.py C3 --top 0 --reps 5 --variants default,r72 --view 0.4,0.3 --precision 64,32 2>&1 | tee gpurun_out/exp_c3.jsonl | cut -c1-420
echo "== other configs"
timeout 1500 python scripts/exp_configs.py C1 C2 C5 --top 0 2>&1 | tee gpurun_out/exp_configs.jsonl | cut -c1-420
echo "== bench N=1"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-1500 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
exit 0
