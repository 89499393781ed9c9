#!/bin/bash
# tools/gpu_r02p.sh TAG: full GPU tests + c1 (recorded steps on / off) + c3 + SNAP ncu
TAG=${1:-r02p}
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_tests.log
timeout 600 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu > $O/${TAG}_bench_c1.json 2> $O/${TAG}_bench_c1.err
timeout 600 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu --graph off --no-e2e --no-mixed > $O/${TAG}_bench_c1_nograph.json 2>> $O/${TAG}_bench_c1.err
timeout 600 python bench.py --workload c1 --steps 100 --warmup 5 --no-cpu --flush-l2 off --no-e2e --no-mixed > $O/${TAG}_bench_c1_noflush.json 2>> $O/${TAG}_bench_c1.err
tools/gpu_snap.sh $TAG bench ncu > /dev/null 2>&1
tail -6 $O/${TAG}_tests.log
for f in $O/${TAG}_bench_c1.json $O/${TAG}_bench_c1_nograph.json $O/${TAG}_bench_c1_noflush.json $O/${TAG}_bench_c3.json; do [ -f $f ] && (echo "== $f"; cut -c1-330 $f); done
tail -5 $O/${TAG}_bench*.err 2>/dev/null
grep -E "^==|gpu__time_duration|fp64.avg|l1tex__throughput|warps_active|registers_per" $O/${TAG}_snap_ncu_full.txt
exit 0
