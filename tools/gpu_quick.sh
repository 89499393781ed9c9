#!/bin/bash
TAG=${1:-rXX}; O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q -k "snap or deck" > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_tests.log
tail -3 $O/${TAG}_tests.log
timeout 300 python tools/snap_bench.py 63 8 3 > $O/${TAG}_snap.out 2>&1; tail -1 $O/${TAG}_snap.out
XSB_SNAP_FKERNEL=1 timeout 300 python tools/snap_bench.py 63 8 3 > $O/${TAG}_snap_old.out 2>&1; tail -1 $O/${TAG}_snap_old.out
