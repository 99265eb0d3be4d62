#!/bin/bash
# GPU box: one `ncu --set full` capture of the kernels matching $1 (regex) in one fwd+bwd step of c2 (eager),
# after the plain run exited 0.  usage: scripts/ncu_one.sh <kernel-regex> <tag> [launch-skip] [launch-count] [extra bench args]
set -u
re=$1; tag=$2; skip=${3:-3}; count=${4:-1}; shift 4 2>/dev/null
out=gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-graph "$@" > $out/ncu_plain_$tag.json 2> $out/ncu_plain_$tag.err || { tail -5 $out/ncu_plain_$tag.err; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name regex:"$re" --launch-skip $skip --launch-count $count \
    -o $out/prof_$tag -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-graph "$@" > $out/ncu_full_$tag.log 2>&1
tail -2 $out/ncu_full_$tag.log
