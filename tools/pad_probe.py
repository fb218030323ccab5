"""Does the z-axis pass suffer from the power-of-two line stride (channel/bank aliasing)?  Same 512^3 transform with
the rows padded (advanced layout).  tools only."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
fft = load_package(); L = fft._lib
n = 512
for pitch_x, pitch_y in [(512, 512), (520, 512), (528, 512), (512, 513), (512, 516), (576, 512)]:
    ie = [n, pitch_y, pitch_x]
    dist = n * pitch_y * pitch_x
    x = torch.zeros(dist, dtype=torch.complex128, device="cuda"); torch.view_as_real(x).uniform_(-0.5, 0.5)
    y = torch.zeros_like(x)
    h = L.plan_many(3, [n, n, n], ie, 1, dist, ie, 1, dist, L.Z2Z, 1)
    for _ in range(3): L.execute(h, L.Z2Z, x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    L.set_profiling(h, True)
    for _ in range(10): L.execute(h, L.Z2Z, x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    nl = L.launch_count(h)
    per = [round(L.launch_ms(h, i), 4) for i in range(nl)]
    print(json.dumps({"pitch_x": pitch_x, "pitch_y": pitch_y, "z_stride_KiB": pitch_x * pitch_y * 16 / 1024, "pass_ms": per, "total": round(sum(per), 4),
                      "desc": [d.split(" threads=")[0][5:] for d in L.describe(h).strip().split("\n")]}), flush=True)
    L.destroy(h); del x, y
