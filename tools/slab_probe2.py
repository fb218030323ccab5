"""torchrun worker (tools only): slab transform timings with phase breakdown and a sampled-bin parity check.
   args: kind:n[:ENV=VAL,...] ...      e.g.  z2z:1024  z2z:1024:FFTB200_SLAB_FUSED=1  d2z:1024"""
import os, sys, json, math
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
import bench
fft = load_package()
from regent_fft_arjun_b200 import distributed as D
L = fft._lib
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
for spec in sys.argv[1:]:
    parts = spec.split(":")
    kind, n = parts[0], int(parts[1])
    env = dict(kv.split("=") for kv in parts[2].split(",")) if len(parts) > 2 else {}
    for k, v in env.items():
        os.environ[k] = v
    shape = (n, n, n)
    real = kind == "d2z"
    dt = {"z2z": fft.complex64, "d2z": fft.double, "c2c": fft.complex32}[kind]
    flops = (2.5 if real else 5.0) * n ** 3 * math.log2(float(n) ** 3)
    plan = D.SlabFFT3D(shape, dt, rank=rank, world=world, device=dev, mode="p2p")
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    x = torch.rand(*plan.local_in_shape, *(() if real else (2,)), dtype=torch.float32 if kind == "c2c" else torch.float64, device=dev, generator=g).sub_(0.5)
    if not real:
        x = torch.view_as_complex(x)
    for _ in range(3): plan.execute(x)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    L.slab_set_timing(plan.engine.h, True)
    K = 10
    e0.record()
    for _ in range(K): plan.execute(x)
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ph = L.slab_phase_ms(plan.engine.h)
    # sampled bins against fp64 direct sums
    n0l, n1l = n // world, n // world
    n2c = n // 2 + 1 if real else n
    bins = bench.pick_bins((n, n, n2c), 64, 5)
    want = bench.direct_bins(x.to(torch.complex128), rank * n0l, shape, bins)
    got = torch.zeros(len(bins), dtype=torch.complex128, device=dev)
    for i, (k0, k1, k2) in enumerate(bins):
        if k1 // n1l == rank:
            got[i] = plan.out[k1 - rank * n1l, k0, k2]
    dist.all_reduce(torch.view_as_real(want)); dist.all_reduce(torch.view_as_real(got))
    err = float((torch.linalg.vector_norm(got - want) / torch.linalg.vector_norm(want)).item())
    if rank == 0:
        print(json.dumps({"n": n, "kind": kind, "world": world, "env": env, "ms": round(float(t.item()), 4), "launches": L.launch_count(plan.engine.h),
                          "GFLOP/s": round(flops / float(t.item()) / 1e6, 1), "phase_ms_rank0": [round(v, 4) for v in ph],
                          "rel_l2_64_bins": err}), flush=True)
    plan.destroy(); del plan, x
    for k in env:
        os.environ.pop(k, None)
    torch.cuda.empty_cache()
dist.destroy_process_group()
