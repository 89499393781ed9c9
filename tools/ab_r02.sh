#!/bin/bash
# A/B switches of round 2 on one GPU: same bench, one environment switch each.  usage: tools/ab_r02.sh TAG
TAG=${1:-r02x}; O=gpurun_out; mkdir -p $O
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-mixed"
run() { name=$1; shift; env "$@" $B > $O/${TAG}_ab_$name.json 2> $O/${TAG}_ab_$name.err; python - $O/${TAG}_ab_$name.json $name <<'PY'
import json,sys
try:
    b=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); d=b["detail"]["breakdown"]
    print("%-14s %.4f ms/step  %.4e  " % (sys.argv[2], b["ms_per_step"], b["value"]) + " ".join("%s %.4f" % (k, v["ms_total"]/v["intervals"]) for k,v in d.items()))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run base XSB_DUMMY=1
run subcell1 XSB_SUBCELL_SORT=1
run subcell2 XSB_SUBCELL_SORT=2
run tpa8 XSB_TPA=8
run sub2_tpa8 XSB_SUBCELL_SORT=2 XSB_TPA=8
for w in c1; do for s in 0 2; do XSB_SUBCELL_SORT=$s python bench.py --workload $w --steps 20 --warmup 5 --no-cpu > $O/${TAG}_ab_${w}_sub$s.json 2>/dev/null; python - $O/${TAG}_ab_${w}_sub$s.json <<'PY'
import json,sys
b=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); print(sys.argv[1], b["ms_per_step"], b["value"], b["e2e"]["value"], (b.get("mixed_precision") or {}).get("value"))
PY
done; done
exit 0
