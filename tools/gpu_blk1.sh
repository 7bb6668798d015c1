#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "rankk or blocked" > gpurun_out/pytest_blk1.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_blk1.log | cut -c1-300
timeout 900 python tools/blk_sweep.py 16384 16384 32768 32768 > gpurun_out/blk_sweep.jsonl 2> gpurun_out/blk_sweep.err; echo "sweep rc=$?"; tail -3 gpurun_out/blk_sweep.err
cat gpurun_out/blk_sweep.jsonl | cut -c1-330
