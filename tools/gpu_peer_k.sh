#!/usr/bin/env bash
G=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$G --master-addr 127.0.0.1"
for k in 32 40; do
  timeout 600 $TR --master-port 29643 bench.py --gpus $G --block-k $k --pivots 640 --no-e2e > gpurun_out/bench_peer_k${k}_g$G.json 2> gpurun_out/bench_peer_k${k}_g$G.err; echo "k=$k g$G rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_peer_k${k}_g$G.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("k=$k g$G value",round(d["value"]),"ms/step",round(d["ms_per_step"],2),"flush ms",round(r["ms_per_launch"],3),"bound",r["bound"],"frac",round(r["frac"],3))
except Exception as e: print("ERR",e)
PY
done
