#!/bin/bash
# 2 GPUs: the two-GPU tests (skipped by a one-GPU box) and bench at N = 2.
set -u
mkdir -p gpurun_out
echo "== pytest tests/test_gpu_dist.py + multi-device context"; timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus or view_groups or multi_device" 2>&1 | tail -4
sumline='import sys,json; d=json.loads(sys.stdin.read()); print("N=%d" % d["n_gpus"], "groups", d["execution"]["view_groups"], "x", d["execution"]["bands_per_view"], "steps", d["steps"], "|", round(d["value"]/1e9,2),"G steps/s", round(d["ms_per_step"],3),"ms | e2e", round(d["e2e"]["value"]/1e9,2), round(d["e2e"]["ms_per_step"],3), "ms | parity", d.get("parity",{}).get("ok"), "bands", d.get("bands"), "numa", d["execution"]["numa"]); print("   last calibration round:", d["calibration"][-1]["sustained_ms"])'
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "rc=$?"; grep '^{' gpurun_out/bench_n2.json | python -c "$sumline" || tail -20 gpurun_out/bench_n2.err
exit 0
