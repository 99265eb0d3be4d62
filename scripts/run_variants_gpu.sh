#!/bin/bash
# usage (GPU box): scripts/run_variants_gpu.sh <workload> v1 v2@KS ...  -> gpurun_out/variants_<workload>.log
# "name@K" runs variants/name.so with RTDA_SPLIT_KS=K (k_split blocks per segment)
wl=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=${v%%@*}; ks=""; [ "$lib" != "$v" ] && ks=${v#*@}
  out=$(RTDA_SPLIT_KS=$ks RTDA_LIB_PATH=$PWD/variants/$lib.so timeout 120 python scripts/gpu_variant_time.py $wl 60 2>&1 | tail -1)
  echo "$out ${ks:+[ks=$ks]}" | tee -a gpurun_out/variants_$wl.log
done
