#!/bin/bash
# Runs on the GPU box (round 2, final kernels): (1) ncu launch list of a short default bench.py run (eager, no secondary configs),
# (2) one `ncu --set full` capture of every kernel of one fwd+bwd step of c2 (plain Lovasz, final-gradient mode:
# the forward writes the finished gradient) and of one fused CE+Lovasz step, (3) the confusion-matrix kernels, (4) raw CSV exports.
# usage: scripts/profile_round2.sh <tag>
set -u
tag=${1:-r2}
out=gpurun_out
B="--steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-graph"
python bench.py $B > $out/profile_plain_$tag.json 2> $out/profile_plain_$tag.err || { tail -5 $out/profile_plain_$tag.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv \
    python bench.py $B > $out/ncu_launches_$tag.log 2>&1
# 15 kernels of this library per plain step (14 forward + the in-place rescale of the backward, a no-op for grad_out = 1); skip the warm-up steps
ncu --set full --clock-control none --import-source on --kernel-name regex:"k_" --launch-skip 75 --launch-count 15 \
    -o $out/prof_$tag -f python bench.py $B > $out/ncu_full_$tag.log 2>&1
ncu -i $out/prof_$tag.ncu-rep --page raw --csv > $out/prof_${tag}_raw.csv 2>/dev/null
tail -2 $out/ncu_full_$tag.log
# fused CE + Lovasz step: the two kernels that differ (k_prep_logits, k_hybrid_bwd_reg)
python bench.py $B --hybrid fused > $out/profile_hybrid_$tag.json 2> $out/profile_hybrid_$tag.err || { tail -5 $out/profile_hybrid_$tag.err; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name regex:"k_prep_logits|k_hybrid_bwd" --launch-skip 10 --launch-count 2 \
    -o $out/prof_hybrid_$tag -f python bench.py $B --hybrid fused > $out/ncu_full_hybrid_$tag.log 2>&1
ncu -i $out/prof_hybrid_$tag.ncu-rep --page raw --csv > $out/prof_hybrid_${tag}_raw.csv 2>/dev/null
tail -2 $out/ncu_full_hybrid_$tag.log
# confusion matrix: 1 map and 8 maps per call
python scripts/gpu_cm_probe.py > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name regex:"k_confusion" --launch-skip 30 --launch-count 12 \
    -o $out/prof_cm_$tag -f python scripts/gpu_cm_probe.py > $out/ncu_full_cm_$tag.log 2>&1
ncu -i $out/prof_cm_$tag.ncu-rep --page raw --csv > $out/prof_cm_${tag}_raw.csv 2>/dev/null
tail -2 $out/ncu_full_cm_$tag.log
