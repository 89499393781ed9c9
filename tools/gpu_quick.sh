#!/bin/bash
TAG=${1:-rXX}; O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_tests.log
tail -3 $O/${TAG}_tests.log
timeout 300 python tools/side_bench.py lj 100 10 0 >> $O/${TAG}_side.json 2>> $O/${TAG}.err
timeout 300 python tools/side_bench.py lj2m 40 5 0 >> $O/${TAG}_side.json 2>> $O/${TAG}.err
timeout 600 python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu > $O/${TAG}_bench_short.json 2>> $O/${TAG}.err
timeout 600 python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu --separate-integrator > $O/${TAG}_bench_short_sep.json 2>> $O/${TAG}.err
python - $O/${TAG}_side.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d["workload"][:30], d["dtype"][:4], "%.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], d["breakdown_ms_per_call"], d.get("rebuilds"))
PY
for f in bench_short bench_short_sep; do python - $O/${TAG}_$f.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); b=d["detail"]["breakdown"]
print("bench %.3e"%d["value"], "ms/step %.3f"%d["ms_per_step"], "launches", d["gpu_launches"], {k:round(v["ms_total"]/v["intervals"],3) for k,v in b.items()})
PY
done
tail -5 $O/${TAG}.err
