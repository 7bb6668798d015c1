#!/usr/bin/env bash
# usage: gpu_peer_multi.sh NGPU [nccl] -- sharded parity check + sharded bench lines on NGPU GPUs
G=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29641 tools/sharded_check.py > gpurun_out/peer_check_$G.log 2>&1; echo "check$G rc=$?"; grep -a "rank 0\|SHARDED" gpurun_out/peer_check_$G.log | cut -c1-260
for w in dense_tableau_32768x65536 dense_tableau_16384x32768; do
  timeout 900 $TR --master-port 29643 bench.py --gpus $G --workload $w > gpurun_out/bench_peer_${w}_g$G.json 2> gpurun_out/bench_peer_${w}_g$G.err; echo "$w peer g$G rc=$?"; tail -2 gpurun_out/bench_peer_${w}_g$G.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_peer_${w}_g$G.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$w g$G value",round(d["value"]),"e2e",d["e2e"] and round(d["e2e"]["value"]),"ms/step",round(d["ms_per_step"],2),"flush ms",round(r["ms_per_launch"],3),"bound",r["bound"],"frac",round(r["frac"],3),"clk",d["clocks"])
except Exception as e: print("ERR",e)
PY
done
if [ "$2" = "nccl" ]; then
timeout 900 $TR --master-port 29645 bench.py --gpus $G --block-k 0 --pivots 40 --no-e2e > gpurun_out/bench_nccl_g$G.json 2> gpurun_out/bench_nccl_g$G.err; echo "nccl path g$G rc=$?"; cut -c1-200 gpurun_out/bench_nccl_g$G.json
fi
