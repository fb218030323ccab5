#!/bin/bash
for rep in 1 2; do
for lib in prev cur; do
if [ $lib = prev ]; then export FFTB200_LIB_PATH=$PWD/tools/altlibs/lib_prev.so; else unset FFTB200_LIB_PATH; fi
echo "== $lib"; python bench.py --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], [(p['ms']) for p in d['roofline']['passes']])"
done; done
