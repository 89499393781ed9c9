#!/bin/bash
# SNAP-only gpurun call: tools/gpu_snap.sh TAG [tests] [bench] [ab] [ncu]
TAG=${1:-r02s}; shift
O=gpurun_out; mkdir -p $O
for what in "$@"; do
case $what in
tests) timeout 900 python -m pytest tests -m gpu -x -q -s -k "snap or Ta06A or WBe or C3" > $O/${TAG}_snaptests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_snaptests.log ;;
bench) timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err ;;
ab) XSB_SNAP_FKERNEL=1 XSB_SNAP_YKERNEL=1 XSB_SNAP_UKERNEL=1 timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu --no-e2e --no-mixed > $O/${TAG}_bench_c3_old.json 2> $O/${TAG}_bench_c3_old.err ;;
ncu) timeout 900 ncu --set full --clock-control none -k regex:'snap_' -c 3 -f -o $O/${TAG}_snap python bench.py --workload c3 --steps 1 --warmup 1 --no-e2e --no-cpu --no-mixed > $O/${TAG}_ncu_snap.out 2>&1; python tools/ncu_summary.py full $O/${TAG}_snap.ncu-rep > $O/${TAG}_snap_ncu_full.txt 2>&1; ncu -i $O/${TAG}_snap.ncu-rep --page details > $O/${TAG}_snap_ncu_details.txt 2>&1; [ $(stat -c %s $O/${TAG}_snap.ncu-rep) -gt 40000000 ] && rm -f $O/${TAG}_snap.ncu-rep ;;
esac
done
for f in $O/${TAG}_snaptests.log; do [ -f $f ] && tail -15 $f; done
for f in $O/${TAG}_bench_c3.json $O/${TAG}_bench_c3_old.json; do [ -f $f ] && (echo "== $f"; cut -c1-1200 $f); done
tail -5 $O/${TAG}_bench*.err 2>/dev/null
exit 0
