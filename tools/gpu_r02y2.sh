#!/bin/bash
TAG=${1:-r02y2}
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q -s -k "analytic or johnson" > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_tests.log
timeout 900 python bench.py --workload c2j --steps 40 --warmup 5 --no-cpu > $O/${TAG}_bench_c2j.json 2> $O/${TAG}_bench_c2j.err
grep "mixed:" $O/${TAG}_tests.log; tail -4 $O/${TAG}_tests.log
for f in $O/${TAG}_bench_c2j.json; do [ -f $f ] && (echo "== $f"; python -c "
import json; d=json.loads(open('$f').read()); print(d['value'], d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('value'), 'mixed', (d.get('mixed_precision') or {}), {k: round(v['ms_total']/v['intervals'],3) for k,v in d['detail']['breakdown'].items()})"); done
tail -n 3 $O/${TAG}_bench_c2j.err
exit 0
