#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "rankk or blocked" > gpurun_out/pytest_blk3.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_blk3.log | cut -c1-300
timeout 900 python tools/blk_sweep.py 32768 32768 16384 16384 > gpurun_out/blk_sweep2.jsonl 2> gpurun_out/blk_sweep2.err; echo "sweep rc=$?"; tail -3 gpurun_out/blk_sweep2.err
cat gpurun_out/blk_sweep2.jsonl | cut -c1-330
