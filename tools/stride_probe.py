"""How does the strided-axis pass depend on the line stride?  Same 512-point fp64 column kernel, same 2 GiB of
data, axis-0 stride from 64 KiB to 4 MiB (tools only; not product code)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
fft = load_package(); L = fft._lib
for n1 in (8, 32, 128, 512):
    batch = 512 // n1
    shape = (512, n1, 512)
    x = torch.zeros((batch,) + shape, dtype=torch.complex128, device="cuda")
    torch.view_as_real(x).uniform_(-0.5, 0.5)
    y = torch.empty_like(x)
    h = L.plan_many(3, list(shape), None, 0, 0, None, 0, 0, L.Z2Z, batch)
    for _ in range(3): L.execute(h, L.Z2Z, x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    L.set_profiling(h, True)
    for _ in range(10): L.execute(h, L.Z2Z, x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    nl = L.launch_count(h)
    desc = L.describe(h).strip().split("\n")
    out = []
    for i in range(nl):
        ms = L.launch_ms(h, i); b = L.launch_bytes(h, i)
        out.append(f"{ms:.3f} ms {b / ms / 1e6:.0f} GB/s ({desc[i].split('(')[-1][:-1]})")
    print(json.dumps({"shape": shape, "batch": batch, "axis0_stride_KiB": n1 * 512 * 16 // 1024, "passes": out}), flush=True)
    L.destroy(h); del x, y
