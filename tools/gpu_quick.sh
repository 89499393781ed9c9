#!/bin/bash
TAG=${1:-rXX}; O=gpurun_out; mkdir -p $O
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > $O/${TAG}_bench_weak_short.json 2> $O/${TAG}.err
timeout 900 python bench.py --scaling strong --steps 40 --warmup 5 --no-cpu > $O/${TAG}_bench_strong_n1.json 2>> $O/${TAG}.err
for f in weak_short strong_n1; do python - $O/${TAG}_bench_$f.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); b=d["detail"]["breakdown"]
print(d["config"]["workload"][:60], "%.3e"%d["value"], "ms/step %.3f"%d["ms_per_step"], "e2e", d["e2e"] and "%.3e"%d["e2e"]["value"], {k:round(v["ms_total"]/v["intervals"],3) for k,v in b.items()})
PY
done
tail -3 $O/${TAG}.err
