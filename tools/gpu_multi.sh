#!/usr/bin/env bash
# usage: bash tools/gpu_multi.sh NGPUS
NG=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR tools/sharded_check.py > gpurun_out/sharded_check_$NG.log 2>&1; echo "sharded_check rc=$?"; grep -E "rank 0|SHARDED|Error|error" gpurun_out/sharded_check_$NG.log | tail -8
timeout 900 $TR bench.py --gpus $NG --workload dense_tableau_16384x32768 > gpurun_out/bench_16k_g$NG.json 2> gpurun_out/bench_16k_g$NG.err; echo "bench16k rc=$?"; tail -3 gpurun_out/bench_16k_g$NG.err
timeout 900 $TR bench.py --gpus $NG > gpurun_out/bench_default_g$NG.json 2> gpurun_out/bench_default_g$NG.err; echo "benchdefault rc=$?"; tail -3 gpurun_out/bench_default_g$NG.err
python - <<PY
import json
for f in ("gpurun_out/bench_16k_g$NG.json","gpurun_out/bench_default_g$NG.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1),"ms/step",round(d["ms_per_step"],2),"dev",round(d["device_ms_per_step"],2),"roof",round(d["roofline"]["achieved"],1),"k3ms",round(d["roofline"]["ms_per_launch"],4),"e2e",d["e2e"] and round(d["e2e"]["value"],1), "launches", d["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
