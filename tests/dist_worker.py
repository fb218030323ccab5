"""Worker for the multi-rank slab tests; launched by torchrun (tests/test_slab.py).

  gloo + --engine oracle : CPU ranks, local FFTs by the oracle (test infrastructure) — checks the
                           decomposition, block layout and exchange bookkeeping of SlabFFT3D
  nccl + --engine cuda   : one GPU per rank, the product path (mode p2p or nccl)
Every rank checks the gathered natural-order result against the oracle's full transform.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402


class OracleSlabEngine:
    """CPU stand-in with the library's pre/post contract (include/fft_b200.h, slab section)."""

    def __init__(self, shape, world, real):
        self.n0, self.n1, self.n2 = shape
        self.G, self.real = world, real

    def pre(self, x, send):
        xs = x.numpy()
        n0l = xs.shape[0]
        if self.real:
            y = np.stack([oracle.port_r2c(p.astype(np.float64)) for p in xs])
        else:
            y = np.stack([oracle.port_dft(p.astype(np.complex128)) for p in xs])          # x and y axes of each plane
        n2c = y.shape[2]
        blocks = y.reshape(n0l, self.G, self.n1 // self.G, n2c).transpose(1, 0, 2, 3)        # [d][n0l][n1l][n2c]
        send.copy_(torch.from_numpy(np.ascontiguousarray(blocks)).to(send.dtype))

    def post(self, recv, out):
        r = recv.numpy().astype(np.complex128)                                                # [s][n0l][n1l][n2c]
        G, n0l, n1l, n2c = r.shape
        full = r.reshape(G * n0l, n1l * n2c).T.copy()                                         # lines along n0, contiguous
        z = np.stack([oracle.port_dft(line) for line in full]).T.reshape(G * n0l, n1l, n2c)
        out.copy_(torch.from_numpy(np.ascontiguousarray(z.transpose(1, 0, 2))).to(out.dtype))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="gloo")
    ap.add_argument("--engine", default="oracle")
    ap.add_argument("--mode", default="nccl")
    ap.add_argument("--shape", default="8,4,8")
    ap.add_argument("--kind", default="z2z")
    ap.add_argument("--chunks", type=int, default=2)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    shape = tuple(int(v) for v in a.shape.split(","))
    fft = load_package()
    from regent_fft_arjun_b200 import distributed as D
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if a.backend == "nccl":
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo")
    dt_in = {"z2z": fft.complex64, "c2c": fft.complex32, "d2z": fft.double, "r2c": fft.float32}[a.kind]
    np_in = {"z2z": np.complex128, "c2c": np.complex64, "d2z": np.float64, "r2c": np.float32}[a.kind]
    real = a.kind in ("d2z", "r2c")
    single = a.kind in ("c2c", "r2c")
    full = oracle.synth(shape, np_in, seed=77)
    engine = OracleSlabEngine(shape, world, real) if a.engine == "oracle" else None
    if len(shape) == 2:
        plan = D.SlabFFT2D(shape, dt_in, rank=rank, world=world, device=dev)
    else:
        plan = D.SlabFFT3D(shape, dt_in, rank=rank, world=world, device=dev, mode=a.mode, chunks=a.chunks, engine=engine)
    n0l = shape[0] // world
    x = torch.from_numpy(np.ascontiguousarray(full[rank * n0l:(rank + 1) * n0l])).to(dev)
    x_keep = x.clone()
    for _ in range(a.reps):                       # repeated collective calls exercise the buffer hand-shake
        plan.execute(x)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    got = plan.gather_natural()
    x64 = full.astype(np.float64 if real else np.complex128)
    if oracle.have_fftw("ref"):                  # the reference's FFTW when it travelled, else the C port
        F = oracle.FFTW.get("ref")
        want = F.r2c(x64) if real else F.dft(x64)
    else:
        want = oracle.port_r2c(x64) if real else oracle.port_dft(x64)
    err = oracle.rel_l2(got, want)
    tol = oracle.tolerance(int(np.prod(shape)), single)
    assert err <= tol, f"rank {rank}: rel-L2 {err:.3e} > {tol:.3e}"
    assert torch.equal(x, x_keep), "input slab was modified"
    assert got.shape == (tuple(shape) if len(shape) == 2 else (shape[0], shape[1], shape[2] // 2 + 1 if real else shape[2]))
    plan.destroy()
    dist.barrier()
    if rank == 0:
        print(f"SLAB_OK world={world} backend={a.backend} mode={a.mode} kind={a.kind} shape={shape} err={err:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
