#!/bin/bash
python tools/ab_probe.py $PWD/regent-fft-arjun_b200/libfft_b200.so $PWD/tools/altlibs/lib_noalloc.so
