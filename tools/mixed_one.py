"""One non-power-of-two transform, a few executions (for ncu captures of the mixed-radix kernel); tools only.
usage: mixed_one.py z2z|c2c batch n0 [n1 [n2]]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
fft = load_package(); L = fft._lib
kind, batch, shape = sys.argv[1], int(sys.argv[2]), [int(a) for a in sys.argv[3:]]
dt = torch.complex128 if kind == "z2z" else torch.complex64
ftype = L.Z2Z if kind == "z2z" else L.C2C
x = torch.zeros(([batch] if batch > 1 else []) + shape, dtype=dt, device="cuda")
torch.view_as_real(x).uniform_(-0.5, 0.5)
y = torch.empty_like(x)
h = L.plan_many(len(shape), shape, None, 0, 0, None, 0, 0, ftype, batch)
L.set_stream(h, torch.cuda.current_stream().cuda_stream)
for _ in range(int(os.environ.get("REPS", "3"))):
    L.execute(h, ftype, x.data_ptr(), y.data_ptr())
torch.cuda.synchronize()
print(L.describe(h))
L.destroy(h)
