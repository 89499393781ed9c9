#!/bin/bash
# one gpurun call: GPU parity tests, smoke, full bench (both arms), ncu launch list.  usage: tools/gpu_round2.sh TAG [full]
# "full" adds the side benches and the ncu --set full captures (EAM + list build, LJ).
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_tests.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/${TAG}_smoke.log 2>&1
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu --no-mixed > $O/${TAG}_launches.out 2>&1
if [ "$2" = "full" ]; then
for w in "lj 100 10 0" "lj 100 10 1" "lj2m 40 5 0" "c2j 40 5 0" "snap 20 3 0" "c5 30 5 0" "c5 30 5 1"; do timeout 600 python tools/side_bench.py $w >> $O/${TAG}_side.json 2>> $O/${TAG}_side.err; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tile_pass_kernel|nbr_' -c 8 -f -o $O/${TAG}_eam python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-mixed > $O/${TAG}_ncu_eam.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tile_pass_kernel' --launch-skip 3 -c 1 -f -o $O/${TAG}_lj python tools/side_bench.py lj2m 3 1 0 > $O/${TAG}_ncu_lj.out 2>&1
fi
tail -3 $O/${TAG}_tests.log; tail -2 $O/${TAG}_smoke.log
cat $O/${TAG}_bench_reference.json | cut -c1-300
cat $O/${TAG}_bench.json
