"""Tiny driver for ncu: plan + a few execs of one transform.  python tools/prof_case.py z2z 512 512 512 [reps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
fft = load_package()
L = fft._lib
kind = sys.argv[1]
args = [int(a) for a in sys.argv[2:]]
reps = 2
shape = args
ftype = {"z2z": L.Z2Z, "c2c": L.C2C, "d2z": L.D2Z, "r2c": L.R2C}[kind]
real = kind in ("d2z", "r2c")
dt_in = {"z2z": torch.complex128, "c2c": torch.complex64, "d2z": torch.float64, "r2c": torch.float32}[kind]
dt_out = torch.complex128 if kind in ("z2z", "d2z") else torch.complex64
x = torch.zeros(shape, dtype=dt_in, device="cuda")
x.view(torch.float64 if kind in ("z2z", "d2z") else torch.float32).uniform_(-0.5, 0.5)
oshape = shape[:-1] + [shape[-1] // 2 + 1] if real else shape
y = torch.empty(oshape, dtype=dt_out, device="cuda")
h = L.plan_many(len(shape), shape, None, 0, 0, None, 0, 0, ftype, 1)
for _ in range(reps):
    L.execute(h, ftype, x.data_ptr(), y.data_ptr())
torch.cuda.synchronize()
print(L.describe(h))
L.destroy(h)
