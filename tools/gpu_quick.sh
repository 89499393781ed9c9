#!/bin/bash
TAG=${1:-rXX}; O=gpurun_out; mkdir -p $O
timeout 300 python tools/side_bench.py c2j 40 5 0 >> $O/${TAG}_side.json 2>> $O/${TAG}.err
XSB_NO_PAIR_CACHE=1 timeout 300 python tools/side_bench.py c2j 40 5 0 >> $O/${TAG}_side.json 2>> $O/${TAG}.err
timeout 300 python tools/side_bench.py snap 20 3 0 >> $O/${TAG}_side.json 2>> $O/${TAG}.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tile_pass_kernel' --launch-skip 4 -c 2 -f -o $O/${TAG}_c2j python tools/side_bench.py c2j 3 1 0 > $O/${TAG}_ncu.out 2>&1
python - $O/${TAG}_side.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d["workload"][:40], d["dtype"][:4], "%.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], d["breakdown_ms_per_call"], d.get("rebuilds"))
PY
tail -5 $O/${TAG}.err
