#!/bin/bash
# one gpurun call: GPU parity tests, smoke, bench, snap timing, ncu launch list, ncu full captures.  usage: tools/gpu_round.sh TAG
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_tests.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/${TAG}_smoke.log 2>&1
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 300 python tools/snap_bench.py > $O/${TAG}_snap.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > $O/${TAG}_launches.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tile_pass_kernel' --launch-skip 6 -c 3 -f -o $O/${TAG}_eam python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/${TAG}_ncu_eam.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'snap_force_kernel' --launch-skip 1 -c 1 -f -o $O/${TAG}_snap python tools/snap_bench.py 32 8 1 > $O/${TAG}_ncu_snap.out 2>&1
tail -3 $O/${TAG}_tests.log; cat $O/${TAG}_smoke.log | tail -2; cat $O/${TAG}_bench.json; tail -2 $O/${TAG}_snap.log
