#!/bin/bash
# GPU box: graph-replayed c2 step and its top kernels on the confident and the diffuse synthetic distribution
for d in conf diffuse; do python bench.py --distribution $d --no-secondary --no-cpu-baseline --steps 20 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$d', 'graph ms', round(d['ms_per_step'],4), 'loss', d['config']['loss'], [ (k['name'],k['ms_per_launch']) for k in d['roofline']['kernels'][:5]])"; done
