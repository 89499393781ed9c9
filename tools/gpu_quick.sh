#!/bin/bash
# quick gpurun call: GPU parity tests + LJ tile-width sweep.  usage: tools/gpu_quick.sh TAG
TAG=${1:-rXX}; O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_tests.log
tail -3 $O/${TAG}_tests.log
for tx in 0 1 2 3 4; do
  if [ $tx = 0 ]; then unset XSB_TILE_TX; else export XSB_TILE_TX=$tx; fi
  timeout 300 python tools/side_bench.py lj2m 40 5 0 >> $O/${TAG}_lj2m.json 2>> $O/${TAG}.err
  timeout 300 python tools/side_bench.py lj 100 10 0 >> $O/${TAG}_lj2m.json 2>> $O/${TAG}.err
done
unset XSB_TILE_TX
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/${TAG}_lj_launches.csv python tools/side_bench.py lj2m 2 1 0 > $O/${TAG}_lj_launches.out 2>&1
python - $O/${TAG}_lj2m.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); print(d["workload"][:30], d["dtype"][:4], "%.3e"%d["value"], "ms/step %.4f"%d["ms_per_step"], d["breakdown_ms_per_call"])
PY
tail -3 $O/${TAG}.err
