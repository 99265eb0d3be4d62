#!/bin/bash
# usage: scripts/bench_variants.sh <bench args...> -- v1 v2 ...   (runs on the GPU box)
args=(); while [ "$1" != "--" ]; do args+=("$1"); shift; done; shift
mkdir -p gpurun_out
for v in "$@"; do
  export RTDA_LIB_PATH=$PWD/variants/$v.so
  timeout 300 python -m pytest tests/test_lovasz_gpu.py -m gpu -q -x -k "golden or random_shapes or ties or full_size" > gpurun_out/test_$v.log 2>&1
  echo "$v tests rc=$? $(tail -1 gpurun_out/test_$v.log)"
  timeout 300 python bench.py "${args[@]}" --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err || tail -3 gpurun_out/bench_$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/bench_{v}.json"))
    ks={k["name"]:k["ms_per_launch"] for k in d["roofline"]["kernels"]}
    print(f"{v}: {d['value']:.0f} Mpix/s {d['ms_per_step']:.3f} ms/step e2e {d['e2e']['value']:.0f} clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']} | "+" ".join(f"{k.replace('k_sort_pass','p').replace('k_','')}={ms:.3f}" for k,ms in ks.items()))
except Exception as e: print(v,"bench failed",e)
PY
done
