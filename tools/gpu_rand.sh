#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "random_lps" > gpurun_out/pytest_rand.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_rand.log | cut -c1-700
