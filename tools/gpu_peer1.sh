#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=1 --master-addr 127.0.0.1 --master-port 29641 tools/sharded_check.py > gpurun_out/peer_check_1.log 2>&1; echo "check1 rc=$?"; grep -v "^\s*$" gpurun_out/peer_check_1.log | tail -25 | cut -c1-300
