#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "more_rows_than_threads" > gpurun_out/pytest_rows.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_rows.log | cut -c1-700
