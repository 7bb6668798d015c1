#!/usr/bin/env bash
# Round-2 GPU call: bit-identity test of the rank-k row-reduction kernels (incl. the new k_blk_flush5), then the sweep against k_blk_flush4.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rankk_update" > gpurun_out/r2_flush5_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_flush5_pytest.log | cut -c1-300
timeout 400 python tools/flush5_sweep.py > gpurun_out/r02_flush5_sweep.jsonl 2> gpurun_out/r02_flush5_sweep.err; echo "sweep rc=$?"
python -c "
import json
for l in open('gpurun_out/r02_flush5_sweep.jsonl'):
    d = json.loads(l); print({k: d[k] for k in d if k in ('kind', 'C', 'k', 'col_steps', 'ms', 'TFLOPs', 'flush_kernel', 'block_k', 'pivots_per_s', 'row_reduction_ms', 'obj')})
"
tail -3 gpurun_out/r02_flush5_sweep.err
