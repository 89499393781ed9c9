#!/bin/bash
# usage: tools/mgpu_round.sh TAG N  -- multi-GPU parity check + weak-scaling bench at N GPUs
TAG=${1:-rXX}; N=${2:-2}; O=gpurun_out; mkdir -p $O
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py > $O/${TAG}_mgpu_check_n$N.log 2>&1; echo "rc=$?" >> $O/${TAG}_mgpu_check_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err
tail -25 $O/${TAG}_mgpu_check_n$N.log; tail -5 $O/${TAG}_bench_n$N.err; cat $O/${TAG}_bench_n$N.json
