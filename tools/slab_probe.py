"""torchrun worker: timing of the slab paths for a shape at several chunk counts (p2p) and NCCL mode."""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
fft = load_package()
from regent_fft_arjun_b200 import distributed as D
L = fft._lib
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
kind = sys.argv[2] if len(sys.argv) > 2 else "z2z"
shape = (n, n, n)
dt = {"z2z": fft.complex64, "d2z": fft.double, "c2c": fft.complex32}[kind]
flops = (2.5 if kind == "d2z" else 5.0) * n ** 3 * np.log2(float(n) ** 3)
quick = len(sys.argv) > 3 and sys.argv[3] == "quick"
one = len(sys.argv) > 3 and sys.argv[3] == "one"
capmode = len(sys.argv) > 3 and sys.argv[3] == "cap"
cfgs = [("p2p", 0, int(os.environ.get("FFTB200_SLAB_P2_CTAS", "148")))] if capmode else [("p2p", 0, 148), ("p2p", 0, 296), ("p2p", 0, 0)] if one else [("p2p", 0, 148), ("p2p", 1, 0), ("p2p", 2, 0), ("nccl", 1, 0)] if quick else [("p2p", 1, 0)] + [("p2p", c, cap) for c in (2, 4, 8) for cap in (148, 296, 444, 0)] + [("nccl", 1, 0)]
for mode, chunks, cap in cfgs:
    os.environ["FFTB200_SLAB_P2_CTAS"] = str(cap)
    plan = D.SlabFFT3D(shape, dt, rank=rank, world=world, device=dev, mode=mode, chunks=chunks)
    x = torch.zeros(plan.local_in_shape, dtype=dt.torch, device=dev)
    (torch.view_as_real(x) if x.is_complex() else x).uniform_(-0.5, 0.5)
    for _ in range(3): plan.execute(x)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if mode == "p2p": L.slab_set_timing(plan.engine.h, True)
    K = 10
    e0.record()
    for _ in range(K): plan.execute(x)
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ph = L.slab_phase_ms(plan.engine.h) if mode == "p2p" else None
    if rank == 0:
        print(json.dumps({"n": n, "kind": kind, "world": world, "mode": mode, "chunks": chunks, "p2_ctas": cap, "ms": round(float(t.item()), 4),
                          "GFLOP/s": round(flops / float(t.item()) / 1e6, 1), "phase_ms_rank0": ph}), flush=True)
    plan.destroy(); del plan, x
    torch.cuda.empty_cache()
dist.destroy_process_group()
