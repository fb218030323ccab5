"""Small transforms of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
fft = load_package(); L = fft._lib
cases = [("z2z", (16, 32, 64)), ("c2c", (8, 64, 32)), ("d2z", (16, 16, 64)), ("r2c", (4, 32, 128)), ("z2z", (1 << 15,)),
         ("c2c", (1 << 16,)), ("z2z", (1024, 16)), ("c2c", (2048, 16)), ("z2z", (6, 10)), ("d2z", (5, 6, 7))]
for kind, shape in cases:
    ftype = {"z2z": L.Z2Z, "c2c": L.C2C, "d2z": L.D2Z, "r2c": L.R2C}[kind]
    real = kind in ("d2z", "r2c")
    dt = {"z2z": torch.complex128, "c2c": torch.complex64, "d2z": torch.float64, "r2c": torch.float32}[kind]
    x = torch.zeros(shape, dtype=dt, device="cuda")
    (torch.view_as_real(x) if x.is_complex() else x).uniform_(-0.5, 0.5)
    oshape = list(shape[:-1]) + [shape[-1] // 2 + 1] if real else list(shape)
    y = torch.empty(oshape, dtype=torch.complex128 if kind in ("z2z", "d2z") else torch.complex64, device="cuda")
    h = L.plan_many(len(shape), list(shape), None, 0, 0, None, 0, 0, ftype, 1)
    L.execute(h, ftype, x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    ref = torch.fft.rfftn(x) if real else torch.fft.fftn(x)
    err = float((torch.linalg.vector_norm((y - ref).to(torch.complex128)) / torch.linalg.vector_norm(ref.to(torch.complex128))).item())
    print(kind, shape, "err", err, L.describe(h).count("\n"), "launches")
    L.destroy(h)
# slab plan, single rank, chunked (two streams)
from regent_fft_arjun_b200 import distributed as D
p = D.SlabFFT3D((16, 32, 64), fft.complex64, rank=0, world=1, device="cuda:0", mode="p2p", chunks=2)
x = torch.zeros(p.local_in_shape, dtype=torch.complex128, device="cuda"); torch.view_as_real(x).uniform_(-0.5, 0.5)
p.execute(x); torch.cuda.synchronize(); p.destroy()
print("sanitize cases done")
