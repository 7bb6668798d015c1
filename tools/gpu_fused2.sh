#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_fused2.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_fused2.log | cut -c1-600
timeout 300 python tools/phase_timing.py > gpurun_out/phase_timing3.jsonl 2>&1; cut -c1-420 gpurun_out/phase_timing3.jsonl
