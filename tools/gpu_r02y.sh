#!/bin/bash
TAG=${1:-r02y}
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_tests.log
timeout 900 python bench.py --workload c5 --steps 20 --warmup 5 --no-cpu > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err
timeout 900 python bench.py --workload c2j --steps 40 --warmup 5 > $O/${TAG}_bench_c2j.json 2> $O/${TAG}_bench_c2j.err
timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
tail -5 $O/${TAG}_tests.log
for f in $O/${TAG}_bench_c5.json $O/${TAG}_bench_c2j.json $O/${TAG}_bench.json; do [ -f $f ] && (echo "== $f"; python -c "
import json; d=json.loads(open('$f').read()); print(d['value'], d['ms_per_step'], 'e2e', (d.get('e2e') or {}).get('value'), 'mixed', (d.get('mixed_precision') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), {k: round(v['ms_total']/v['intervals'],3) for k,v in d['detail']['breakdown'].items()}, d['detail']['force_checksum_sum_abs'], d['detail'].get('pair_operators_fused_into_eam_force_pass'))"); done
tail -3 $O/${TAG}_bench_c5.err $O/${TAG}_bench_c2j.err $O/${TAG}_bench.err
exit 0
