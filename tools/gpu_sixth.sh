#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu6.log | cut -c1-250
for w in dense_revised_dual_4096x12288 dense_revised_dual_dse_4096x12288; do
timeout 600 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"; tail -2 gpurun_out/bench_$w.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
    print("$w value",round(d["value"],1),"pivots/step",d["config"]["pivots_per_step"],"dev ms/step",round(d["device_ms_per_step"],2),"k3 ms",round(d["roofline"]["ms_per_launch"],4),"share",round(d["roofline"]["share_of_step_device_time"],3),"e2e",d["e2e"] and round(d["e2e"]["value"],1),"cpu",d["cpu_baseline"] and round(d["cpu_baseline"]["value"],2))
except Exception as e: print("ERR", e)
PY
done
CMD2="python bench.py --workload dense_revised_dual_4096x12288 --steps 1 --warmup 1 --pivots 16 --no-e2e --no-cpu"
$CMD2 > gpurun_out/plain3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_dual4k.csv $CMD2 > gpurun_out/ncu_list2.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_dual4k.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg.setdefault(r[ki][:44],[]).append(v)
tot=sum(sum(v) for v in agg.values())
for k,v in agg.items(): print(f"  {k:44s} n={len(v):4d} avg={sum(v)/len(v)/1e3:9.1f} us share={sum(v)/tot:.3f}")
PY
