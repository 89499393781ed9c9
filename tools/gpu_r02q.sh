#!/bin/bash
# tools/gpu_r02q.sh TAG: SNAP tests + c3 + SNAP ncu (with details page) + c1 recorded-step A/B
TAG=${1:-r02q}
O=gpurun_out; mkdir -p $O
tools/gpu_snap.sh $TAG tests bench ncu > /dev/null 2>&1
timeout 600 python bench.py --workload c1 --steps 20 --warmup 5 --no-cpu > $O/${TAG}_bench_c1.json 2> $O/${TAG}_bench_c1.err
timeout 600 python bench.py --workload c1 --steps 100 --warmup 5 --no-cpu --flush-l2 off --no-e2e --no-mixed > $O/${TAG}_bench_c1_noflush.json 2>> $O/${TAG}_bench_c1.err
timeout 600 python bench.py --workload c1 --steps 100 --warmup 5 --no-cpu --flush-l2 off --graph off --no-e2e --no-mixed > $O/${TAG}_bench_c1_noflush_nograph.json 2>> $O/${TAG}_bench_c1.err
tail -4 $O/${TAG}_snaptests.log
for f in $O/${TAG}_bench_c1.json $O/${TAG}_bench_c1_noflush.json $O/${TAG}_bench_c1_noflush_nograph.json $O/${TAG}_bench_c3.json; do [ -f $f ] && (echo "== $f"; cut -c1-330 $f); done
tail -5 $O/${TAG}_bench*.err 2>/dev/null
grep -E "^==|gpu__time_duration|fp64.avg|l1tex__throughput|warps_active|registers_per" $O/${TAG}_snap_ncu_full.txt
exit 0
