#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s -k "blocked_lu or lu_refactorisation or generated_dual" > gpurun_out/pytest_k4.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_k4.log | cut -c1-300
