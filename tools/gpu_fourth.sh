#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "generated" > gpurun_out/pytest_gen.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gen.log
timeout 900 python bench.py --workload dense_revised_dual_4096x12288 > gpurun_out/bench_dual4k.json 2> gpurun_out/bench_dual4k.err; echo "dual4k rc=$?"; tail -3 gpurun_out/bench_dual4k.err
CMD="python bench.py --workload dense_tableau_16384x32768 --steps 1 --warmup 1 --pivots 4 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_16k_v2.csv $CMD > gpurun_out/ncu_list.log 2>&1
CMD2="python bench.py --workload dense_revised_dual_4096x12288 --steps 1 --warmup 1 --pivots 4 --no-e2e --no-cpu"
$CMD2 > gpurun_out/plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_dual4k.csv $CMD2 > gpurun_out/ncu_list2.log 2>&1
python - <<'PY'
import json,csv,collections
try:
    d=json.loads(open("gpurun_out/bench_dual4k.json").read().strip().splitlines()[-1])
    print("dual4k value",round(d["value"],1),"ms/step",round(d["ms_per_step"],2),"dev",round(d["device_ms_per_step"],2),"roof",round(d["roofline"]["achieved"],1),"k3ms",round(d["roofline"]["ms_per_launch"],4),"share",round(d["roofline"]["share_of_step_device_time"],3),"e2e",d["e2e"] and round(d["e2e"]["value"],1),"cpu",d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2), "launches", d["gpu_launches"])
except Exception as e: print("ERR",e)
for f in ("gpurun_out/launches_16k_v2.csv","gpurun_out/launches_dual4k.csv"):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
    agg=collections.OrderedDict()
    for r in rows[1:]:
        try: v=float(r[vi].replace(',',''))
        except: continue
        agg.setdefault(r[ki][:44],[]).append(v)
    tot=sum(sum(v) for v in agg.values())
    print(f)
    for k,v in agg.items(): print(f"  {k:44s} n={len(v):4d} avg={sum(v)/len(v)/1e3:9.1f} us share={sum(v)/tot:.3f}")
PY
