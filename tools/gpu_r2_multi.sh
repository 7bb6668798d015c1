#!/usr/bin/env bash
# Round 2, N GPUs (torchrun): parity of the sharded engines (primal + dual), then the bench lines of the sharded workloads.
#   gpurun --gpus N -- 'bash tools/gpu_r2_multi.sh N [check] [bench] [dual] [batch] [16k]'
N=${1:-2}; shift || true
WHAT="${*:-check bench dual batch}"
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
if [[ " $WHAT " == *" check "* ]]; then
  run tools/sharded_check.py > gpurun_out/r02_sharded_check_g$N.log 2>&1; echo "sharded_check rc=$?"; grep -c "ok=True\|_ok=True\|identical=True" gpurun_out/r02_sharded_check_g$N.log; grep "False\|SHARDED_CHECK_OK\|Error\|error" gpurun_out/r02_sharded_check_g$N.log | head -20
fi
if [[ " $WHAT " == *" bench "* ]]; then
  run bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_default_g$N.json 2> gpurun_out/r02_bench_default_g$N.err; echo "bench default rc=$?"; cut -c1-400 gpurun_out/r02_bench_default_g$N.json; tail -3 gpurun_out/r02_bench_default_g$N.err
fi
if [[ " $WHAT " == *" dual "* ]]; then
  for wl in dense_tableau_dual_32768x65536 dense_tableau_dual_4096x12288; do
    run bench.py --gpus $N --workload $wl --steps 10 --warmup 3 > gpurun_out/r02_bench_${wl}_g$N.json 2> gpurun_out/r02_bench_${wl}_g$N.err; echo "bench $wl rc=$?"; cut -c1-300 gpurun_out/r02_bench_${wl}_g$N.json; tail -3 gpurun_out/r02_bench_${wl}_g$N.err
  done
fi
if [[ " $WHAT " == *" 16k "* ]]; then
  run bench.py --gpus $N --workload dense_tableau_16384x32768 --steps 10 --warmup 3 > gpurun_out/r02_bench_16k_g$N.json 2> gpurun_out/r02_bench_16k_g$N.err; echo "bench 16k rc=$?"; cut -c1-300 gpurun_out/r02_bench_16k_g$N.json
fi
if [[ " $WHAT " == *" batch "* ]]; then
  run bench.py --gpus $N --workload batch_small_lps_65536x64x128 --steps 5 --warmup 3 > gpurun_out/r02_bench_batch_g$N.json 2> gpurun_out/r02_bench_batch_g$N.err; echo "bench batch rc=$?"; cut -c1-400 gpurun_out/r02_bench_batch_g$N.json
fi
if [[ " $WHAT " == *" owner "* ]]; then
  for orr in 0 1; do
    run bench.py --gpus $N --steps 10 --warmup 3 --owner-ratio $orr --no-e2e > gpurun_out/r02_bench_default_owner${orr}_g$N.json 2> gpurun_out/r02_bench_default_owner${orr}_g$N.err; echo "bench default owner_ratio=$orr rc=$?"; cut -c1-160 gpurun_out/r02_bench_default_owner${orr}_g$N.json
  done
fi
