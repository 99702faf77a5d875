timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/km_launches.csv python benchmarks/bench_kmeans.py > gpurun_out/km_bench.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/km_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[vi].replace(",",""));
    u=r[ui]
    if u=="us": v/=1e3
    elif u=="ns": v/=1e6
    elif u=="s" or u=="second": v*=1e3
    a=agg.setdefault(r[ki][:70],[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1]): print("%-72s n=%6d total %9.3f ms mean %8.4f" % (k,c,t,t/c))
PY
cat gpurun_out/km_bench.log | tail -3
