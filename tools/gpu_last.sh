#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_last.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_last.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_last_g2.json 2> gpurun_out/bench_last_g2.err; echo "bench g2 rc=$?"; cut -c1-200 gpurun_out/bench_last_g2.json
timeout 300 python bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | cut -c1-200
