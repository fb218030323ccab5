#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/cufft_compare.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['kind'], d['shape'], d['b200_ms_min'], d['speedup_vs_cufft'], d['rel_l2_vs_cufft'], [p.split(' | ')[0] for p in d['passes']])
"
python bench.py --no-cpu-baseline | cut -c1-200
