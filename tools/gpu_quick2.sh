#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_plans or concurrent or host_memory" 2>&1 | tail -15
