#!/bin/bash
for alt in 0 1 2; do echo "== alt=$alt"; FFTB200_TILE_ALT=$alt python bench.py --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], [(p['kernel'][17:70],p['ms']) for p in d['roofline']['passes']])"; done
python -c "import __graft_entry__ as g; g.smoke()"
