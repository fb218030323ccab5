#!/bin/bash
timeout 900 python -m pytest tests/test_slab.py -m gpu -x -q -k "2d" 2>&1 | tail -8
python - <<'PY'
import os, sys, json, torch, numpy as np
sys.path.insert(0, os.getcwd())
PY
