#!/bin/bash
# usage: tools/mgpu_r02.sh TAG N [check] [bench] [bench3] [ref] [strong] -- multi-GPU parity check + scaling benches at N GPUs
TAG=${1:-r02x}; N=${2:-2}; shift; shift; O=gpurun_out; mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for what in "$@"; do
case $what in
topo) nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1 ;;
check) timeout 300 $RUN --master-port 29511 tests/mgpu_check.py > $O/${TAG}_mgpu_check_n$N.log 2>&1; echo "rc=$?" >> $O/${TAG}_mgpu_check_n$N.log ;;
bench) timeout 900 $RUN --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err ;;
bench3) for i in 1 2 3; do timeout 900 $RUN --master-port 2952$i bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-mixed --no-e2e >> $O/${TAG}_bench3_n$N.json 2>> $O/${TAG}_bench_n$N.err; done ;;
sync) timeout 900 $RUN --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-mixed --no-e2e --sync-displ > $O/${TAG}_bench_sync_n$N.json 2>> $O/${TAG}_bench_n$N.err ;;
ref) timeout 900 $RUN --master-port 29514 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/${TAG}_bench_ref_n$N.json 2>> $O/${TAG}_bench_n$N.err ;;
strong) timeout 900 $RUN --master-port 29515 bench.py --workload c4 --gpus $N --steps 20 --warmup 5 --no-cpu --no-mixed > $O/${TAG}_bench_c4_n$N.json 2>> $O/${TAG}_bench_n$N.err ;;
esac
done
[ -f $O/${TAG}_mgpu_check_n$N.log ] && tail -12 $O/${TAG}_mgpu_check_n$N.log
tail -5 $O/${TAG}_bench_n$N.err 2>/dev/null
for f in $O/${TAG}_bench_n$N.json $O/${TAG}_bench3_n$N.json $O/${TAG}_bench_sync_n$N.json $O/${TAG}_bench_ref_n$N.json $O/${TAG}_bench_c4_n$N.json; do [ -f $f ] && (echo "== $f"; python - "$f" <<'PY'
import json,sys
for line in open(sys.argv[1]):
    line=line.strip()
    if not line.startswith("{"): continue
    b=json.loads(line)
    d=b.get("detail",{})
    print(json.dumps({k:b.get(k) for k in ("value","ms_per_step","n_gpus")}), "e2e", (b.get("e2e") or {}).get("value"), "ranks", (d.get("ranks") or {}).get("timed_region_ms"), "rebuild", (d.get("ranks") or {}).get("rebuild_wall_s"), "ghost", (d.get("ranks") or {}).get("ghost_ms"), "move", (d.get("ranks") or {}).get("move_ms"))
PY
); done
exit 0
