#!/bin/bash
TAG=${1:-r02r}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -s -k "inner_skin or recorded_step or transfers or snap_force_parity" > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu --no-e2e > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err
tail -8 $O/${TAG}_tests.log
for f in $O/${TAG}_bench.json $O/${TAG}_bench_c3.json; do [ -f $f ] && (echo "== $f"; cut -c1-330 $f; python -c "
import json,sys; d=json.loads(open('$f').read()); print('e2e', d.get('e2e')); print('mixed', d.get('mixed_precision'))"); done
tail -5 $O/${TAG}_bench*.err 2>/dev/null
exit 0
