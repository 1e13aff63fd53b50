#!/bin/bash
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 scripts/debug_p2p.py base 2>&1 | grep "it=" 
C5_GRAZE_SERIAL=1 timeout 300 $TR --master-port 29522 scripts/debug_p2p.py serial 2>&1 | grep "it="
timeout 300 $TR --master-port 29523 scripts/debug_p2p.py wait 2>&1 | grep "it="
C5_STORE_FENCE=1 timeout 300 $TR --master-port 29524 scripts/debug_p2p.py fence 2>&1 | grep "it="
exit 0
