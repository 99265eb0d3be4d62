#!/bin/bash
# GPU box: kernel time of the confusion matrix for 1 and 8 maps per call (bench.py --workload c4)
for m in 1 8; do python bench.py --workload c4 --steps 40 --no-cpu-baseline --maps-per-call $m 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['maps_per_step'], 'maps: graph ms/step', round(d['ms_per_step'],5), 'kernel ms', round(d['roofline']['ms_per_launch'],5), 'frac', round(d['roofline']['frac'],4), 'value', round(d['value']))"; done
