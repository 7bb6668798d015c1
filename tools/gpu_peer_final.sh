#!/usr/bin/env bash
# usage: gpu_peer_final.sh NGPU [check] [16k] [batch]
G=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1"
if [[ " $* " == *" check "* ]]; then
timeout 600 $TR --master-port 29641 tools/sharded_check.py > gpurun_out/peer_check_$G.log 2>&1; echo "check$G rc=$?"; grep -a "rank 0.*duplicated\|rank 0: peer vs\|SHARDED" gpurun_out/peer_check_$G.log | cut -c1-220
fi
WL="dense_tableau_32768x65536"
if [[ " $* " == *" 16k "* ]]; then WL="$WL dense_tableau_16384x32768"; fi
for w in $WL; do
  timeout 900 $TR --master-port 29643 bench.py --gpus $G --workload $w > gpurun_out/bench_peer_${w}_g$G.json 2> gpurun_out/bench_peer_${w}_g$G.err; echo "$w peer g$G rc=$?"; tail -2 gpurun_out/bench_peer_${w}_g$G.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_peer_${w}_g$G.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$w g$G value",round(d["value"]),"e2e",d["e2e"] and round(d["e2e"]["value"]),"ms/step",round(d["ms_per_step"],2),"flush ms",round(r["ms_per_launch"],3),"bound",r["bound"],"frac",round(r["frac"],3),"clk",d["clocks"])
except Exception as e: print("ERR",e)
PY
done
if [[ " $* " == *" batch "* ]]; then
timeout 900 $TR --master-port 29647 bench.py --gpus $G --workload batch_small_lps_65536x64x128 --steps 3 > gpurun_out/bench_batch_g$G.json 2> gpurun_out/bench_batch_g$G.err; echo "batch g$G rc=$?"; tail -2 gpurun_out/bench_batch_g$G.err | cut -c1-300; cut -c1-500 gpurun_out/bench_batch_g$G.json
fi
