#!/usr/bin/env bash
# usage: gpu_peer_multi.sh NGPU  -- sharded parity check + sharded bench lines on NGPU GPUs
G=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29641 tools/sharded_check.py > gpurun_out/peer_check_$G.log 2>&1; echo "check$G rc=$?"; grep -v "^\s*$" gpurun_out/peer_check_$G.log | grep -v "rank [1-9]" | tail -14 | cut -c1-300
for w in dense_tableau_32768x65536 dense_tableau_16384x32768; do
  timeout 900 $TR --master-port 29643 bench.py --gpus $G --workload $w > gpurun_out/bench_peer_${w}_g$G.json 2> gpurun_out/bench_peer_${w}_g$G.err; echo "$w peer g$G rc=$?"; tail -2 gpurun_out/bench_peer_${w}_g$G.err | cut -c1-300; cut -c1-2200 gpurun_out/bench_peer_${w}_g$G.json
done
timeout 900 $TR --master-port 29645 bench.py --gpus $G --block-k 0 --pivots 40 --no-e2e > gpurun_out/bench_nccl_g$G.json 2> gpurun_out/bench_nccl_g$G.err; echo "nccl path g$G rc=$?"; cut -c1-400 gpurun_out/bench_nccl_g$G.json
