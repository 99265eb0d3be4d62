#!/bin/bash
# usage (GPU box): scripts/bench_variants2.sh <workload> v1 v2 ... ; prints graph-replayed ms/step per variant
wl=$1; shift
for v in base "$@"; do
  if [ "$v" = base ]; then unset RTDA_LIB_PATH; else export RTDA_LIB_PATH=$PWD/variants/$v.so; fi
  python bench.py --workload $wl --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/var_${wl}_$v.json 2> gpurun_out/var_${wl}_$v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/var_${wl}_$v.json").read().strip().splitlines()[-1])
    ks={k["name"]:k["ms_per_launch"] for k in d["roofline"]["kernels"]}
    print("$wl $v ms %.4f value %.1f fb %s split %.3f part %.3f local %.3f" % (d["ms_per_step"], d["value"], d["config"]["segments_sorted_by_lsd_fallback"], ks.get("k_split",0), ks.get("k_part",0), ks.get("k_local",0)))
except Exception as e:
    print("$wl $v failed", e)
PY
done
