#!/bin/bash
# final single-GPU set of round 2 (after the chain fusion / mixed analytic EAM / two-atom count sweep)
TAG=${1:-r02zz}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/${TAG}_gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket|NUMA" >> $O/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_tests.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_smoke.log
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err
for w in c1 c3 c5 c2j; do timeout 900 python bench.py --workload $w --steps 20 --warmup 5 > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu --no-mixed > $O/${TAG}_launches.out 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chained or stream_bit_exact or analytic_models_mixed or empty_and_ragged or c5_two_species" > $O/${TAG}_sanitizer.log 2>&1; echo "sanitizer rc=$?" >> $O/${TAG}_sanitizer.log
tail -3 $O/${TAG}_tests.log; tail -3 $O/${TAG}_smoke.log; tail -6 $O/${TAG}_sanitizer.log
for f in $O/${TAG}_bench.json $O/${TAG}_bench_reference.json $O/${TAG}_bench_c1.json $O/${TAG}_bench_c3.json $O/${TAG}_bench_c5.json $O/${TAG}_bench_c2j.json; do [ -f $f ] && (echo "== $f"; python -c "
import json; d=json.loads(open('$f').read()); print(d.get('value'), d.get('ms_per_step'), 'e2e', (d.get('e2e') or {}).get('value'), 'mixed', (d.get('mixed_precision') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'roof', (d.get('roofline') or {}).get('frac'), d.get('clocks'))"); done
tail -n 4 $O/${TAG}_bench.err
exit 0
