timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pipelined or residual or complete_solve or devex or batch" > gpurun_out/r2_pytest10.log 2>&1; echo "pytest subset rc=$?"; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest10.log | head; grep -n "^E  " gpurun_out/r2_pytest10.log | head -12 | cut -c1-600
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --workload batch_small_lps_65536x64x128 --steps 5 --warmup 3 > gpurun_out/r02_bench_batch_small_lps_65536x64x128_g1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_batch_small_lps_65536x64x128_g1.json').read().strip().splitlines()[-1]); print('batch g1', round(d['value']), d['e2e'])"
bash tools/gpu_r2_multi.sh 2 bench batch
