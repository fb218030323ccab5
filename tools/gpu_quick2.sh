#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "inverse_real or golden or random or properties" 2>&1 | tail -15
