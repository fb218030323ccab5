#!/bin/bash
mkdir -p gpurun_out
python tools/stride_probe.py | tee gpurun_out/stride_probe.log
