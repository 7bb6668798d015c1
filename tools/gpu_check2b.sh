#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "order_free or sharded_engines or full_size" > gpurun_out/pytest_chk2b.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_chk2b.log | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29641 tools/sharded_check.py > gpurun_out/peer_check_2.log 2>&1; echo "check2 rc=$?"; grep -a "duplicated\|SHARDED\|Error\|error" gpurun_out/peer_check_2.log | cut -c1-260 | head
