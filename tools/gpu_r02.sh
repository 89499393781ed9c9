#!/bin/bash
# one gpurun call of round 2.  usage: tools/gpu_r02.sh TAG [tests] [bench] [ref] [ncu] [micro] [side]
TAG=${1:-r02x}; shift
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/${TAG}_gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket|NUMA" >> $O/${TAG}_gpu.txt 2>&1
for what in "$@"; do
case $what in
micro) timeout 120 ./tools/microbench > $O/${TAG}_microbench.txt 2>&1 ;;
tests) timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_tests.log ;;
newtests) timeout 900 python -m pytest tests/test_reference_files.py -m gpu -q -s > $O/${TAG}_newtests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_newtests.log ;;
smoke) timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/${TAG}_smoke.log 2>&1 ;;
bench) timeout 900 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err ;;
bench100) timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu > $O/${TAG}_bench100.json 2> $O/${TAG}_bench100.err ;;
ref) timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err ;;
launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu --no-mixed > $O/${TAG}_launches.out 2>&1 ;;
ncu) timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tile_pass_kernel|nbr_' -c 8 -f -o $O/${TAG}_eam python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-mixed > $O/${TAG}_ncu_eam.out 2>&1 ;;
side) for w in c1 c3 c5; do timeout 900 python bench.py --workload $w --steps 20 --warmup 5 > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; done ;;
c4) timeout 900 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err ;;
esac
done
for f in $O/${TAG}_tests.log $O/${TAG}_newtests.log $O/${TAG}_smoke.log; do [ -f $f ] && tail -4 $f; done
for f in $O/${TAG}_microbench.txt; do [ -f $f ] && cat $f; done
for f in $O/${TAG}_bench_reference.json $O/${TAG}_bench.json $O/${TAG}_bench100.json $O/${TAG}_bench_c1.json $O/${TAG}_bench_c3.json $O/${TAG}_bench_c5.json $O/${TAG}_bench_c4.json; do [ -f $f ] && (echo "== $f"; cut -c1-2500 $f); done
tail -5 $O/${TAG}_bench*.err 2>/dev/null
exit 0
