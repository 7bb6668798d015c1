#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/k4_bench.py 1024 4096 8192 > gpurun_out/k4_bench.jsonl 2>&1; echo "k4 rc=$?"; cat gpurun_out/k4_bench.jsonl
CMD="python tools/k4_bench.py 4096"
$CMD > gpurun_out/plain_k4.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file gpurun_out/launches_k4.csv $CMD > gpurun_out/ncu_k4_list.log 2>&1
$CMD > gpurun_out/plain_k4b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_dgemm_sub_dmma -s 40 -c 2 -o gpurun_out/prof_k4_dgemm $CMD > gpurun_out/ncu_k4_full.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_k4.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg.setdefault(r[ki][:40],[]).append(v)
tot=sum(sum(v) for v in agg.values())
for k,v in agg.items(): print(f"  {k:40s} n={len(v):5d} sum={sum(v)/1e6:8.2f} ms avg={sum(v)/len(v)/1e3:8.1f} us share={sum(v)/tot:.3f}")
PY
