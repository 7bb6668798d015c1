#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s -k "batch" > gpurun_out/pytest_batch.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_batch.log | cut -c1-300
