#!/bin/bash
TAG=${1:-r02y6}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_host_decks.py -m gpu -x -q > $O/${TAG}_decktests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_decktests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tile_pass_kernel|nbr_count|nbr_expand' -c 12 -f -o $O/${TAG}_c5 python bench.py --workload c5 --cells 60 --steps 2 --warmup 1 --no-e2e --no-cpu --no-mixed > $O/${TAG}_ncu_c5.out 2>&1
tail -4 $O/${TAG}_decktests.log; tail -3 $O/${TAG}_ncu_c5.out; ls -la $O/${TAG}_c5.ncu-rep
exit 0
