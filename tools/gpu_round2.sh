#!/bin/bash
# one gpurun call: GPU parity tests, smoke, A/B bench runs of the kernel switches, full bench, ncu captures.  usage: tools/gpu_round2.sh TAG [quick]
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?" >> $O/${TAG}_tests.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/${TAG}_smoke.log 2>&1
S="--steps 40 --warmup 5 --no-e2e --no-cpu"
timeout 600 python bench.py $S > $O/${TAG}_ab_default.json 2> $O/${TAG}_ab.err
XSB_TILE_DEAL=1 timeout 600 python bench.py $S > $O/${TAG}_ab_deal.json 2>> $O/${TAG}_ab.err
XSB_NO_PAIR_CACHE=1 timeout 600 python bench.py $S > $O/${TAG}_ab_nocache.json 2>> $O/${TAG}_ab.err
for m in 0 1; do
timeout 300 python tools/side_bench.py lj2m 40 5 $m >> $O/${TAG}_lj2m.json 2>> $O/${TAG}_ab.err
XSB_TILE_DEAL=1 timeout 300 python tools/side_bench.py lj2m 40 5 $m >> $O/${TAG}_lj2m.json 2>> $O/${TAG}_ab.err
done
if [ "$2" != "quick" ]; then
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > $O/${TAG}_launches.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tile_pass_kernel' --launch-skip 6 -c 4 -f -o $O/${TAG}_eam python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/${TAG}_ncu_eam.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tile_pass_kernel' --launch-skip 3 -c 1 -f -o $O/${TAG}_lj python tools/side_bench.py lj2m 3 1 0 > $O/${TAG}_ncu_lj.out 2>&1
fi
tail -3 $O/${TAG}_tests.log; tail -2 $O/${TAG}_smoke.log
for f in default deal nocache; do python - $O/${TAG}_ab_$f.json $f <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); b=d["detail"]["breakdown"]
    print(sys.argv[2], "%.3e"%d["value"], "ms/step %.3f"%d["ms_per_step"], {k:round(v["ms_total"]/v["intervals"],3) for k,v in b.items()})
except Exception as e: print(sys.argv[2], "failed", e)
PY
done
cat $O/${TAG}_lj2m.json
cat $O/${TAG}_bench.json 2>/dev/null
