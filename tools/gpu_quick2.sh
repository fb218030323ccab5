#!/bin/bash
for cap in 0 296 444; do echo "== cap=$cap"; FFTB200_GRID_CAP=$cap python bench.py --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], [(p['kernel'][17:50],p['ms']) for p in d['roofline']['passes']])"; done
timeout 600 python -m pytest tests/test_slab.py -m gpu -x -q 2>&1 | tail -3
for cap in 148 96 64 222; do
FFTB200_SLAB_P2_CTAS=$cap python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/slab_probe.py 512 z2z cap 2>&1 | grep "^{" | sed "s/^/cap=$cap /"
done
