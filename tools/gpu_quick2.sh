#!/bin/bash
for alt in 0 1 2 3; do echo "== alt=$alt"; FFTB200_TILE_ALT=$alt python tools/cufft_compare.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        if d['shape']==[1024,1024,1024] and d['kind']=='z2z': print(d['kind'], d['shape'], d['b200_ms_min'], [p for p in d['passes']][:1])
"; done
