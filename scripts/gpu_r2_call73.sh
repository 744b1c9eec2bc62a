#!/bin/bash
# In-step DRAM traffic of the headline kernels: ncu with --cache-control none (caches are NOT flushed between launches, one pass per kernel:
# the two dram byte counters fit one pass), against the cold-cache figures of the --set full captures
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --batch 296 --no-index --no-cpu --no-extra"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/r2c76_traffic_warm.csv $CMD > gpurun_out/r2c76_bench.log 2>&1
python - <<'PY'
import csv, collections, re
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2c76_traffic_warm.csv') if l.startswith('"'))]
hdr=rows[0]; col={h:i for i,h in enumerate(hdr)}
agg=collections.OrderedDict()
for r in rows[1:]:
    name=re.sub(r"\(.*","",r[col["Kernel Name"]]).replace("void ","").replace("kj::","")
    m=r[col["Metric Name"]]; v=float(r[col["Metric Value"]].replace(",","")); u=r[col["Metric Unit"]]
    if u=="Kbyte": v*=1e3
    elif u=="Mbyte": v*=1e6
    elif u=="Gbyte": v*=1e9
    elif u in ("usecond","us"): v*=1e-6
    elif u=="msecond": v*=1e-3
    elif u in ("nsecond","ns"): v*=1e-9
    a=agg.setdefault(name,collections.defaultdict(float)); a[m]+=v; a["n_"+m]+=1
print("| kernel | launches | dram read MB per launch | dram write MB per launch | us per launch |")
print("|---|---|---|---|---|")
for k,a in agg.items():
    n=a["n_dram__bytes_read.sum"] or 1
    print(f"| `{k}` | {int(n)} | {a['dram__bytes_read.sum']/n/1e6:.1f} | {a['dram__bytes_write.sum']/n/1e6:.1f} | {a['gpu__time_duration.sum']/n*1e6:.1f} |")
PY
