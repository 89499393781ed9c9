#!/bin/bash
TAG=${1:-r02w}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -s -k "sublist or c5 or pair or inner_skin or johnson or eam_alloy" > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_tests.log
timeout 900 python bench.py --workload c5 --steps 20 --warmup 5 --no-cpu > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err
XSB_PAIR_NO_SUBLIST=1 timeout 900 python bench.py --workload c5 --steps 20 --warmup 5 --no-cpu --no-e2e --no-mixed > $O/${TAG}_bench_c5_nosub.json 2>> $O/${TAG}_bench_c5.err
tail -5 $O/${TAG}_tests.log
for f in $O/${TAG}_bench_c5.json $O/${TAG}_bench_c5_nosub.json; do [ -f $f ] && (echo "== $f"; python -c "
import json; d=json.loads(open('$f').read()); print(d['value'], d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('value'), {k: round(v['ms_total']/v['intervals'],3) for k,v in d['detail']['breakdown'].items()})"); done
tail -5 $O/${TAG}_bench_c5.err
exit 0
