#!/bin/bash
TAG=${1:-r02y7}
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_tests.log
timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu --no-e2e --no-mixed > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
XSB_NBR_EXPAND_DIRECT=1 timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu --no-e2e --no-mixed > $O/${TAG}_bench_direct.json 2>> $O/${TAG}_bench.err
timeout 900 python bench.py --workload c5 --steps 20 --warmup 5 --no-cpu --no-e2e --no-mixed > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tile_pass_kernel|nbr_count|nbr_expand' -c 12 -f -o $O/${TAG}_c5 python bench.py --workload c5 --cells 60 --steps 2 --warmup 1 --no-e2e --no-cpu --no-mixed > $O/${TAG}_ncu_c5.out 2>&1
tail -4 $O/${TAG}_tests.log
for f in $O/${TAG}_bench.json $O/${TAG}_bench_direct.json $O/${TAG}_bench_c5.json; do [ -f $f ] && (echo "== $f"; python -c "
import json; d=json.loads(open('$f').read()); print(d['value'], d['ms_per_step'], {k: round(v['ms_total']/v['intervals'],3) for k,v in d['detail']['breakdown'].items()}, d['detail']['force_checksum_sum_abs'])"); done
tail -n 3 $O/${TAG}_bench.err; tail -n 2 $O/${TAG}_ncu_c5.out; ls -la $O/${TAG}_c5.ncu-rep
exit 0
