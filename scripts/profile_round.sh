#!/bin/bash
# Runs on the GPU box: (1) ncu launch list of a short bench.py run, (2) one ncu --set full capture of
# every kernel of one fwd+bwd step (workload c2), (3) raw CSV export.  usage: scripts/profile_round.sh <tag>
set -u
tag=${1:-r1}
out=gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $out/profile_plain_$tag.json 2> $out/profile_plain_$tag.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $out/ncu_launches_$tag.log 2>&1
# 14 kernels of this library per step (13 forward + k_bwd); skip the 3 warm-up steps
ncu --set full --clock-control none --import-source on --kernel-name regex:"k_" --launch-skip 42 --launch-count 14 \
    -o $out/prof_$tag -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > $out/ncu_full_$tag.log 2>&1
ncu -i $out/prof_$tag.ncu-rep --page raw --csv > $out/prof_${tag}_raw.csv 2>/dev/null
tail -2 $out/ncu_full_$tag.log
