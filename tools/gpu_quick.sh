#!/bin/bash
TAG=${1:-rXX}; O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_tests.log
tail -12 $O/${TAG}_tests.log | grep -E "passed|failed|rc=|Error|assert" | head
timeout 600 python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu > $O/${TAG}_bench_short.json 2>> $O/${TAG}.err
python - $O/${TAG}_bench_short.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); b=d["detail"]["breakdown"]
print("bench %.3e"%d["value"], "ms/step %.3f"%d["ms_per_step"], {k:round(v["ms_total"]/v["intervals"],3) for k,v in b.items()})
print("mixed", d["mixed_precision"])
PY
tail -5 $O/${TAG}.err
