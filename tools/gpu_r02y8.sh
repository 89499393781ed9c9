#!/bin/bash
TAG=${1:-r02y8}
O=gpurun_out; mkdir -p $O
for s in 0.16 0.2 0.28; do
timeout 900 python bench.py --workload c5 --steps 20 --warmup 5 --no-cpu --no-e2e --no-mixed --inner-skin $s > $O/${TAG}_bench_c5_skin$s.json 2> $O/${TAG}_bench_c5.err
done
for f in $O/${TAG}_bench_c5_skin*.json; do [ -f $f ] && (echo "== $f"; python -c "
import json; d=json.loads(open('$f').read()); print(d['value'], d['ms_per_step'], {k: round(v['ms_total']/v['intervals'],3) for k,v in d['detail']['breakdown'].items()}, d['detail']['eam_sublist'])"); done
exit 0
