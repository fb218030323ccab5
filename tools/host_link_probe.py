"""torchrun worker (tools only): what the host link gives N ranks at once.  Each rank copies `mb` MiB pinned -> device
and device -> pinned concurrently on two streams (plain cudaMemcpyAsync through torch), K times; prints per-rank GB/s
each way (min over ranks) and the aggregate.  The e2e numbers of bench.py are bounded by this, not by the library."""
import os, sys, json, torch, torch.distributed as dist
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = mb * 1024 * 1024
hx = torch.empty(n, dtype=torch.uint8, pin_memory=True); hy = torch.empty(n, dtype=torch.uint8, pin_memory=True)
dx = torch.empty(n, dtype=torch.uint8, device=dev); dy = torch.zeros(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(k, both=True):
    for _ in range(k):
        with torch.cuda.stream(s1): dx.copy_(hx, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): hy.copy_(dy, non_blocking=True)
for both in (False, True):
    run(2, both); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    e0.record(); torch.cuda.current_stream().wait_event(e0)
    s1.wait_event(e0); s2.wait_event(e0)
    run(K, both)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    gbs = n * K / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([gbs], device=dev, dtype=torch.float64)
    tmin, tsum = t.clone(), t.clone()
    if world > 1:
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({"ranks": world, "MiB": mb, "mode": "H2D + D2H concurrently" if both else "H2D only",
                          "GB/s_each_way_per_rank_min": round(float(tmin.item()), 2), "GB/s_each_way_aggregate": round(float(tsum.item()), 1)}), flush=True)
if world > 1: dist.destroy_process_group()
