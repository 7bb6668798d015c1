#!/usr/bin/env bash
mkdir -p gpurun_out
python tools/blk_probe.py 32768 32768 48 96 2 > gpurun_out/blk_probe_plain.log 2>&1; echo "plain rc=$?"; cat gpurun_out/blk_probe_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_blk48_32k.csv python tools/blk_probe.py 32768 32768 48 96 2 > gpurun_out/ncu_blk_list.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_blk48_32k.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:60]].append(float(r[vi].replace(",","")))
    except Exception: pass
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print(f"{k:62s} n={len(v):5d} avg={sum(v)/len(v)/1e3:9.2f} us  share={sum(v)/tot:6.3f}")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_blk_flush -c 1 -o gpurun_out/prof_blk_flush_k48 python tools/blk_probe.py 32768 32768 48 96 1 > gpurun_out/ncu_blk_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_blk_full.log
