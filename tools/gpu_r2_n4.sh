#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 tools/slab_probe2.py z2z:512:FFTB200_SLAB_FUSED=1 z2z:512:FFTB200_SLAB_FUSED=0 z2z:512:FFTB200_SLAB_FUSED=1,FFTB200_SLAB_PLANE_CHUNKS=2 z2z:512:FFTB200_SLAB_FUSED=0,FFTB200_SLAB_PLANE_CHUNKS=2 z2z:1024:FFTB200_SLAB_FUSED=0 d2z:1024 c2c:512 > gpurun_out/r02_slab_probe_n${N}.jsonl 2> gpurun_out/r02_slab_probe_n${N}.err; echo "probe rc=$?"
cat gpurun_out/r02_slab_probe_n${N}.jsonl; tail -n 3 gpurun_out/r02_slab_probe_n${N}.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n${N}.json").read().strip().split("\n")[-1])
print("N=$N 512^3 ms", round(d["ms_per_step"],4), "GF", round(d["value"]), "parity", d["parity"]["ok"], "1024^3", {k:d["scaling_1024"].get(k) for k in ("ms","GFLOP/s")}, d["scaling_1024"]["parity"]["ok"], "e2e", {k:d["e2e"].get(k) for k in ("value","ms_per_step","host_link_GB/s_each_way")})
PY
