"""Incumbent check (NOT product code): cuFFT through torch.fft on the same shapes, same timing method.
The reference's GPU branch calls cuFFT (src/fft.rg:571-580); this shows where libfft_b200 stands against it."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
fft = load_package(); L = fft._lib

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))

cases = [("z2z", (512, 512, 512)), ("z2z", (256, 256, 256)), ("d2z", (4096, 4096)), ("c2c", (1 << 27,)), ("z2z", (1024, 1024, 1024)),
         ("d2z", (1024, 1024, 1024)), ("c2c", (512, 512, 512)), ("z2z", (4096, 4096)), ("z2z", (1 << 24,)), ("c2c", (1024, 1024, 1024))]
for kind, shape in cases:
    real = kind in ("d2z", "r2c")
    dt = {"z2z": torch.complex128, "c2c": torch.complex64, "d2z": torch.float64, "r2c": torch.float32}[kind]
    ftype = {"z2z": L.Z2Z, "c2c": L.C2C, "d2z": L.D2Z, "r2c": L.R2C}[kind]
    x = torch.zeros(shape, dtype=dt, device="cuda")
    (torch.view_as_real(x) if x.is_complex() else x).uniform_(-0.5, 0.5)
    oshape = list(shape[:-1]) + [shape[-1] // 2 + 1] if real else list(shape)
    y = torch.empty(oshape, dtype=torch.complex128 if kind in ("z2z", "d2z") else torch.complex64, device="cuda")
    h = L.plan_many(len(shape), list(shape), None, 0, 0, None, 0, 0, ftype, 1)
    L.set_stream(h, torch.cuda.current_stream().cuda_stream)
    ours = timeit(lambda: L.execute(h, ftype, x.data_ptr(), y.data_ptr()))
    desc = L.describe(h).strip().split("\n")
    nl = L.launch_count(h)
    L.set_profiling(h, True)
    for _ in range(5): L.execute(h, ftype, x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    per = [(L.launch_ms(h, i), L.launch_bytes(h, i)) for i in range(nl)]
    L.set_profiling(h, False)
    if real:
        cu = timeit(lambda: torch.fft.rfftn(x, out=y))
        ref = torch.fft.rfftn(x)
    else:
        cu = timeit(lambda: torch.fft.fftn(x, out=y))
        ref = torch.fft.fftn(x)
    L.execute(h, ftype, x.data_ptr(), y.data_ptr()); torch.cuda.synchronize()
    diff = float((torch.linalg.vector_norm((y - ref).to(torch.complex128)) / torch.linalg.vector_norm(ref.to(torch.complex128))).item())
    del ref
    n = float(np.prod(shape)); flops = (2.5 if real else 5.0) * n * np.log2(n)
    print(json.dumps({"kind": kind, "shape": shape, "b200_ms_min": round(ours[0], 4), "cufft_ms_min": round(cu[0], 4),
                      "speedup_vs_cufft": round(cu[0] / ours[0], 3), "b200_GFLOP/s": round(flops / ours[0] / 1e6, 1),
                      "rel_l2_vs_cufft": diff,
                      "passes": [f"{ms:.3f} ms {b / ms / 1e6:.0f} GB/s | {d.split(' lines=')[0]}" for (ms, b), d in zip(per, desc)]}), flush=True)
    L.destroy(h); del x, y
    torch.cuda.empty_cache()
