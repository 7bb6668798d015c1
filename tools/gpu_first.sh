#!/usr/bin/env bash
# first GPU call: environment probe, smoke, parity tests, K3 sweep
mkdir -p gpurun_out
{ nvidia-smi; nproc; free -g; cat /sys/fs/cgroup/memory.max 2>/dev/null; ulimit -l; } > gpurun_out/env.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/k3_sweep.py 16384 32768 4096 4096 > gpurun_out/k3_sweep.jsonl 2>&1
tail -5 gpurun_out/smoke.log; tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/k3_sweep.jsonl | tail -30
