#!/usr/bin/env bash
# full GPU suite + smoke + default bench line (1 GPU)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2_pytest3.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/r2_smoke.log | cut -c1-300
