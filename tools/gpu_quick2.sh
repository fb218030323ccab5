#!/bin/bash
for pf in -1 1 0; do echo "== FFTB200_PREFETCH=$pf"; 
if [ $pf = -1 ]; then unset FFTB200_PREFETCH; else export FFTB200_PREFETCH=$pf; fi
python bench.py --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], [(p['ms']) for p in d['roofline']['passes']])"
python bench.py --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], [(p['ms']) for p in d['roofline']['passes']])"
done
