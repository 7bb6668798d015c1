#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "full_size or sharded" > gpurun_out/pytest_chk2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_chk2.log | cut -c1-600
bash tools/gpu_peer_multi.sh 2
