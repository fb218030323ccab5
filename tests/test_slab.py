"""Slab decomposition.

CPU (no GPU): shape bookkeeping, and the N>1 host path under gloo with world_size 2 and 4 (local FFTs
by the oracle engine of tests/dist_worker.py).
GPU, one device: G ranks emulated in one process through the library's staged pre/post halves with the
all-to-all done by torch indexing; the fused p2p path with G = 1 (chunk pipeline, no peers).
GPU, >= 2 devices (skipped otherwise): the real thing under torchrun + NCCL, both exchange modes.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def _torchrun(nproc, args, port, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER] + args
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert res.returncode == 0 and "SLAB_OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
    return res.stdout


def test_slab_shapes(fft):
    from regent_fft_arjun_b200 import distributed as D
    assert D.slab_shapes((512, 512, 512), 8, False) == ((64, 512, 512), (64, 512, 512), (8, 64, 64, 512))
    assert D.slab_shapes((1024, 1024, 1024), 8, True) == ((128, 1024, 1024), (128, 1024, 513), (8, 128, 128, 513))
    with pytest.raises(AssertionError):
        D.slab_shapes((6, 8, 8), 4, False)
    assert D.slab_chunks(512, 4) == 4 and D.slab_chunks(513, 4) == 4 and D.slab_chunks(8, 4) == 1
    parts = [np.arange(2 * 4 * 3).reshape(2, 4, 3) + 100 * r for r in range(2)]     # [n1l=2][n0=4][n2c=3] per rank
    nat = D.assemble_transposed(parts)
    assert nat.shape == (4, 4, 3) and nat[1, 2, 0] == parts[1][0, 1, 0]


@pytest.mark.parametrize("world,kind,shape", [(2, "z2z", "8,4,8"), (2, "d2z", "4,8,16"), (4, "z2z", "8,8,4")])
def test_slab_host_path_gloo(built, world, kind, shape):
    _torchrun(world, ["--backend", "gloo", "--engine", "oracle", "--mode", "nccl", "--kind", kind, "--shape", shape],
              port=29700 + world * 10 + len(kind) + len(shape))


# ------------------------------------------------------------------------------------------
def _emulate(L, oracle, kind, shape, G):
    ftype = {"z2z": L.Z2Z, "c2c": L.C2C, "d2z": L.D2Z, "r2c": L.R2C}[kind]
    np_in = {"z2z": np.complex128, "c2c": np.complex64, "d2z": np.float64, "r2c": np.float32}[kind]
    real, single = kind in ("d2z", "r2c"), kind in ("c2c", "r2c")
    cdt = torch.complex64 if single else torch.complex128
    n0, n1, n2 = shape
    n2c = n2 // 2 + 1 if real else n2
    n0l, n1l = n0 // G, n1 // G
    full = oracle.synth(shape, np_in, seed=88)
    plans = [L.slab_plan(list(shape), ftype, r, G, 1) for r in range(G)]
    sends = []
    for r in range(G):
        x = torch.from_numpy(np.ascontiguousarray(full[r * n0l:(r + 1) * n0l])).cuda()
        send = torch.zeros(G, n0l, n1l, n2c, dtype=cdt, device="cuda")
        L.slab_exec_pre(plans[r], x.data_ptr(), send.data_ptr())
        sends.append(send)
    torch.cuda.synchronize()
    outs = []
    for d in range(G):
        recv = torch.stack([sends[s][d] for s in range(G)]).contiguous()       # the all-to-all
        out = torch.zeros(n1l, n0, n2c, dtype=cdt, device="cuda")
        L.slab_exec_post(plans[d], recv.data_ptr(), out.data_ptr())
        outs.append(out)
    torch.cuda.synchronize()
    for h in plans:
        L.destroy(h)
    got = np.concatenate([o.cpu().numpy() for o in outs], axis=0).transpose(1, 0, 2)
    x64 = full.astype(np.float64 if real else np.complex128)
    want = oracle.port_r2c(x64) if real else oracle.port_dft(x64)
    return oracle.rel_l2(got, want), oracle.tolerance(int(np.prod(shape)), single)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["z2z", "c2c", "d2z", "r2c"])
def test_slab_staged_emulated_ranks(fft, oracle, kind):
    L = fft._lib
    for shape, G in [((16, 16, 16), 2), ((32, 64, 16), 4), ((64, 32, 128), 8), ((16, 16, 8), 16)]:
        err, tol = _emulate(L, oracle, kind, shape, G)
        assert err <= tol, (kind, shape, G, err)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,chunks", [("z2z", 1), ("z2z", 4), ("d2z", 3), ("c2c", 2)])
def test_slab_fused_single_rank_chunks(fft, oracle, kind, chunks):
    """G = 1: the fused path degenerates to 3 local passes, pipelined over chunks on two streams."""
    L = fft._lib
    from regent_fft_arjun_b200 import distributed as D
    dt = {"z2z": fft.complex64, "c2c": fft.complex32, "d2z": fft.double}[kind]
    np_in = {"z2z": np.complex128, "c2c": np.complex64, "d2z": np.float64}[kind]
    shape = (32, 64, 128)
    x = oracle.synth(shape, np_in, seed=89)
    plan = D.SlabFFT3D(shape, dt, rank=0, world=1, device="cuda:0", mode="p2p", chunks=chunks)
    for _ in range(3):
        plan.execute(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    got = plan.gather_natural()
    plan.destroy()
    real = kind == "d2z"
    x64 = x.astype(np.float64 if real else np.complex128)
    want = oracle.port_r2c(x64) if real else oracle.port_dft(x64)
    assert oracle.rel_l2(got, want) <= oracle.tolerance(int(np.prod(shape)), kind == "c2c")


@pytest.mark.gpu
@pytest.mark.parametrize("kind,planes", [("z2z", 4), ("c2c", 2), ("z2z", 8)])
def test_slab_fused_single_rank_plane_chunks(fft, oracle, kind, planes, monkeypatch):
    """complex transforms run y (+exchange) -> x -> z with the x pass pipelined over plane chunks"""
    from regent_fft_arjun_b200 import distributed as D
    monkeypatch.setenv("FFTB200_SLAB_PLANE_CHUNKS", str(planes))
    dt = {"z2z": fft.complex64, "c2c": fft.complex32}[kind]
    np_in = {"z2z": np.complex128, "c2c": np.complex64}[kind]
    shape = (32, 64, 128)
    x = oracle.synth(shape, np_in, seed=90)
    plan = D.SlabFFT3D(shape, dt, rank=0, world=1, device="cuda:0", mode="p2p")
    assert fft._lib.launch_count(plan.engine.h) == 1 + 2 * planes
    xd = torch.from_numpy(x).cuda()
    for _ in range(3):
        plan.execute(xd)
    torch.cuda.synchronize()
    got = plan.gather_natural()
    plan.destroy()
    assert np.array_equal(xd.cpu().numpy(), x)
    assert oracle.rel_l2(got, oracle.port_dft(x.astype(np.complex128))) <= oracle.tolerance(int(np.prod(shape)), kind == "c2c")


@pytest.mark.gpu
@pytest.mark.parametrize("kind,side,planes", [("z2z", 128, 4), ("z2z", 128, 1), ("c2c", 128, 8), ("z2z", 256, 4), ("c2c", 256, 2)])
def test_slab_single_kernel_cubes_single_rank(fft, oracle, kind, side, planes, monkeypatch):
    """complex cubes run the whole slab transform as ONE persistent kernel (tickets: y chunks, x chunks as they
    'arrive', z); with one rank the hand-shake flags are the rank's own.  Also in place-safe reuse and the multi-launch
    path (FFTB200_SLAB_FUSED=0) giving the same bits."""
    from regent_fft_arjun_b200 import distributed as D
    monkeypatch.setenv("FFTB200_SLAB_PLANE_CHUNKS", str(planes))
    dt = {"z2z": fft.complex64, "c2c": fft.complex32}[kind]
    np_in = {"z2z": np.complex128, "c2c": np.complex64}[kind]
    shape = (side, side, side)
    x = oracle.synth(shape, np_in, seed=93)
    xd = torch.from_numpy(x).cuda()
    outs = []
    for fused in ("1", "0"):
        monkeypatch.setenv("FFTB200_SLAB_FUSED", fused)
        plan = D.SlabFFT3D(shape, dt, rank=0, world=1, device="cuda:0", mode="p2p")
        assert (fft._lib.launch_count(plan.engine.h) == 1) == (fused == "1"), fft._lib.describe(plan.engine.h)
        for _ in range(3):
            plan.execute(xd)
        torch.cuda.synchronize()
        outs.append(plan.out.clone())
        got = plan.gather_natural()
        plan.destroy()
        assert np.array_equal(xd.cpu().numpy(), x)
        want = oracle.FFTW.get("ref").dft(x.astype(np.complex128)) if oracle.have_fftw("ref") else oracle.port_dft(x.astype(np.complex128))
        assert oracle.rel_l2(got, want) <= oracle.tolerance(int(np.prod(shape)), kind == "c2c"), (kind, side, fused)
    assert torch.equal(outs[0], outs[1]), "single-kernel and multi-launch paths differ"


@pytest.mark.gpu
def test_slab_2d_single_rank(fft, oracle):
    """2-D slab plan with G = 1: row pass into the (local) receive slab transposed, row pass back"""
    from regent_fft_arjun_b200 import distributed as D
    for kind, shape in [("z2z", (64, 128)), ("c2c", (256, 32)), ("z2z", (1024, 16)), ("z2z", (8, 4096))]:
        dt = {"z2z": fft.complex64, "c2c": fft.complex32}[kind]
        np_in = {"z2z": np.complex128, "c2c": np.complex64}[kind]
        x = oracle.synth(shape, np_in, seed=91)
        plan = D.SlabFFT2D(shape, dt, rank=0, world=1, device="cuda:0")
        xd = torch.from_numpy(x).cuda()
        for _ in range(2):
            plan.execute(xd)
        torch.cuda.synchronize()
        got = plan.gather_natural()
        plan.destroy()
        assert np.array_equal(xd.cpu().numpy(), x)
        assert oracle.rel_l2(got, oracle.port_dft(x.astype(np.complex128))) <= oracle.tolerance(int(np.prod(shape)), kind == "c2c")


@pytest.mark.gpu
def test_slab_2d_multi_gpu(built):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    for kind, shape in [("z2z", "256,512"), ("c2c", "1024,128"), ("z2z", "2048,4096")]:
        _torchrun(world, ["--backend", "nccl", "--engine", "cuda", "--mode", "p2p", "--kind", kind, "--shape", shape,
                          "--reps", "3"], port=29850 + len(shape))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["nccl", "p2p"])
def test_slab_multi_gpu(built, mode):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    # (the 128^3 and 256^3 complex cubes take the single-kernel path in p2p mode: slab_fused_kernel.cuh)
    for kind, shape in [("z2z", "64,64,64"), ("d2z", "32,64,128"), ("c2c", "128,128,64"), ("z2z", "128,128,128"),
                        ("c2c", "256,256,256")]:
        _torchrun(world, ["--backend", "nccl", "--engine", "cuda", "--mode", mode, "--kind", kind, "--shape", shape,
                          "--chunks", "4", "--reps", "3"], port=29800 + (7 if mode == "p2p" else 0) + len(shape))


@pytest.mark.gpu
@pytest.mark.parametrize("kind,shape", [("z2z", (128, 128, 128)), ("d2z", (64, 128, 256)), ("z2z", (64, 32, 48 // 3 * 4))])
def test_slab_single_process_two_gpus(fft, oracle, kind, shape):
    """The Legion way: ONE process, one plan per GPU created with that GPU current, the exchange areas
    connected with fftb200_slab_get_area / fftb200_slab_connect_ptrs (peer access, no IPC), both execs issued from the same
    host thread (they are asynchronous; each GPU's kernels wait for the other's flags on the device)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    L = fft._lib
    G = 2
    ftype = {"z2z": L.Z2Z, "d2z": L.D2Z}[kind]
    np_in = {"z2z": np.complex128, "d2z": np.float64}[kind]
    real = kind == "d2z"
    n0, n1, n2 = shape
    n2c = n2 // 2 + 1 if real else n2
    full = oracle.synth(shape, np_in, seed=97)
    plans, xs, outs = [], [], []
    for r in range(G):
        with torch.cuda.device(r):
            plans.append(L.slab_plan(list(shape), ftype, r, G, 0))
            xs.append(torch.from_numpy(np.ascontiguousarray(full[r * (n0 // G):(r + 1) * (n0 // G)])).cuda(r))
            outs.append(torch.zeros((n1 // G, n0, n2c), dtype=torch.complex128, device=f"cuda:{r}"))
    areas = [L.slab_area(h)[0] for h in plans]
    for r in range(G):
        with torch.cuda.device(r):
            L.slab_connect_ptrs(plans[r], areas)
    for rep in range(3):
        for r in range(G):
            with torch.cuda.device(r):
                L.set_stream(plans[r], torch.cuda.current_stream(r).cuda_stream)
                L.slab_exec(plans[r], xs[r].data_ptr(), outs[r].data_ptr())
    for r in range(G):
        torch.cuda.synchronize(r)
    from regent_fft_arjun_b200 import distributed as D
    got = D.assemble_transposed([o.cpu().numpy() for o in outs])
    for r in range(G):      # no rank frees its exchange area while the peer could still store into it
        torch.cuda.synchronize(r)
    for r in range(G):
        with torch.cuda.device(r):
            L.destroy(plans[r])
    F = oracle.FFTW.get("ref") if oracle.have_fftw("ref") else None
    x64 = full.astype(np.float64 if real else np.complex128)
    want = (F.r2c(x64) if real else F.dft(x64)) if F else (oracle.port_r2c(x64) if real else oracle.port_dft(x64))
    assert oracle.rel_l2(got, want) <= oracle.tolerance(int(np.prod(shape)), False), (kind, shape)
    for r in range(G):
        assert np.array_equal(xs[r].cpu().numpy(), full[r * (n0 // G):(r + 1) * (n0 // G)]), "input slab modified"
