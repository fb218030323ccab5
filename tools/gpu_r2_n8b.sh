#!/bin/bash
# eight GPUs: host-link ceiling at 1/2/4/8 ranks, final bench line at N=8
mkdir -p gpurun_out
O=gpurun_out/r02_host_link_probe.jsonl; : > $O
python tools/host_link_probe.py 256 >> $O 2>/dev/null
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2965$n tools/host_link_probe.py 256 2>/dev/null | grep "^{" >> $O
done
cat $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n8.json").read().strip().split("\n")[-1])
print("N=8 512^3 ms", round(d["ms_per_step"],4), "GF", round(d["value"]), "parity", d["parity"]["ok"], "1024^3", {k:d["scaling_1024"].get(k) for k in ("ms","GFLOP/s")}, d["scaling_1024"]["parity"]["ok"], "e2e", {k:d["e2e"].get(k) for k in ("value","ms_per_step","host_link_GB/s_each_way")})
PY
