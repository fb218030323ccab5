"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI (ctypes over
libfft_b200.so) and through the host mirror of fft.generate_fft_interface, against

* the committed golden vectors produced by the reference's own FFTW (tests/golden),
* the eight known answers of the reference's test program (test/fft_test.rg:138-389),
* the oracle (reference FFTW from oracle/_ref when it travelled, else the plain-C port) on the same
  seeded inputs at sizes the oracle finishes in seconds,
* size-independent properties at BASELINE.json's full sizes (impulse, linearity, time shift, Parseval,
  Hermitian symmetry, forward->backward round trip: libbench2/verify-lib.c:260-434).

Tolerance (BASELINE.json north_star): relative L2 <= 10*log2(N)*eps of the precision
(oracle.tolerance).  fp32 results are compared with the fp64 transform of the same fp32 input.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fftw_golden.npz")


@pytest.fixture(scope="module")
def L(fft):
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    torch.cuda.set_device(0)
    return fft._lib


def _torch_dtype(np_dtype):
    return torch.from_numpy(np.zeros(1, np_dtype)).dtype


def _kinds(L):
    return {"z2z": (L.Z2Z, np.complex128, np.complex128), "c2c": (L.C2C, np.complex64, np.complex64),
            "d2z": (L.D2Z, np.float64, np.complex128), "r2c": (L.R2C, np.float32, np.complex64)}


def gpu_fft(L, kind, x, shape, batch=1, direction=-1):
    """basic-layout transform of `batch` packed arrays of `shape` through the C ABI"""
    ftype, dt_in, dt_out = _kinds(L)[kind]
    real = kind in ("d2z", "r2c")
    full = ((batch,) if batch > 1 else ()) + tuple(shape)
    oshape = full[:-1] + (full[-1] // 2 + 1,) if real else full
    xd = torch.from_numpy(np.ascontiguousarray(x, dtype=dt_in).reshape(full)).cuda()
    yd = torch.zeros(oshape, dtype=_torch_dtype(dt_out), device="cuda")
    h = L.plan_many(len(shape), list(shape), None, 0, 0, None, 0, 0, ftype, batch)
    try:
        L.set_stream(h, torch.cuda.current_stream().cuda_stream)
        L.execute(h, ftype, xd.data_ptr(), yd.data_ptr(), direction)
        torch.cuda.synchronize()
        desc = L.describe(h)
    finally:
        L.destroy(h)
    assert np.array_equal(xd.cpu().numpy().ravel(), np.ascontiguousarray(x, dtype=dt_in).ravel()), "input not preserved"
    return yd.cpu().numpy(), desc


def cpu_fft(oracle, kind, x, shape, batch=1):
    """oracle answer in fp64: the reference's FFTW when oracle/_ref travelled, else the C port"""
    real = kind in ("d2z", "r2c")
    full = ((batch,) if batch > 1 else ()) + tuple(shape)
    x64 = np.ascontiguousarray(x).reshape(full).astype(np.float64 if real else np.complex128)
    F = oracle.FFTW.get("ref") if oracle.have_fftw("ref") else None
    outs = []
    for b in range(batch):
        xb = x64[b] if batch > 1 else x64
        if F is not None:
            outs.append(F.r2c(xb) if real else F.dft(xb))
        else:
            outs.append(oracle.port_r2c(xb) if real else oracle.port_dft(xb))
    return np.stack(outs) if batch > 1 else outs[0]


# ------------------------------------------------------------------------------------------
# golden vectors and known answers
# ------------------------------------------------------------------------------------------
def test_golden_vectors_fp64(L, oracle):
    g = np.load(GOLDEN)
    names = sorted({k[:-3] for k in g.files})
    assert len(names) >= 20
    for name in names:
        x, y = g[name + "__x"], g[name + "__y"]
        if name.startswith("batch"):
            ftype = L.Z2Z if name.startswith("batch_c") else L.D2Z
            xd = torch.from_numpy(x).cuda()
            yd = torch.zeros(18, dtype=torch.complex128, device="cuda")
            h = L.plan_many(2, [3, 3], [3, 3], 1, 9, [3, 3], 1, 9, ftype, 2)       # src/fft.rg:389-398
            L.execute(h, ftype, xd.data_ptr(), yd.data_ptr())
            torch.cuda.synchronize()
            L.destroy(h)
            got = yd.cpu().numpy()
        else:
            kind = "z2z" if x.dtype.kind == "c" else "d2z"
            got, _ = gpu_fft(L, kind, x, x.shape)
        err = oracle.rel_l2(got, y)
        assert err <= oracle.tolerance(x.size, single=False), (name, err)


def test_golden_vectors_fp32(L, oracle):
    """the same fixtures rounded to fp32: complex32 / float paths (the reference's CPU branch is a
    no-op for these, src/fft.rg:296-307; its GPU branch is cuFFT C2C)"""
    g = np.load(GOLDEN)
    for name in sorted({k[:-3] for k in g.files}):
        if name.startswith("batch"):
            continue
        x = g[name + "__x"]
        kind = "c2c" if x.dtype.kind == "c" else "r2c"
        x32 = x.astype(np.complex64 if kind == "c2c" else np.float32)
        got, _ = gpu_fft(L, kind, x32, x.shape)
        want = cpu_fft(oracle, kind, x32, x.shape)
        err = oracle.rel_l2(got, want)
        assert err <= oracle.tolerance(x.size, single=True), (name, err)


def _run_iface(fft, iface, extent, fill, out_fill=0, batch=False):
    r = fft.Region(extent, iface.dtype_in).fill(fill)
    s = fft.Region(extent, iface.dtype_out).fill(out_fill)
    p = fft.PlanRegion(1)
    (iface.make_plan_batch if batch else iface.make_plan)(r, s, p)
    iface.execute_plan_task(r, s, p)
    torch.cuda.synchronize()
    iface.destroy_plan(p)
    assert int(p.data["b200_p"][0]) == 0
    return s.numpy()


def test_reference_test_program_known_answers(fft, L):
    """test/fft_test.rg: constant input, plan -> execute -> destroy; SURVEY.md §4 table."""
    c64, c32, dbl, flt = fft.complex64, fft.complex32, fft.double, fft.float32
    z = 0
    # test1d (:242)
    out = _run_iface(fft, fft.generate_fft_interface(fft.int1d, c64, c64), (5,), 3 + 3j)
    assert np.allclose(out, [15 + 15j, z, z, z, z], atol=1e-13)
    # test1d_real (:138): only n/2+1 = 2 entries are written
    out = _run_iface(fft, fft.generate_fft_interface(fft.int1d, dbl, c64), (3,), 3.0, out_fill=-7)
    assert np.allclose(out[:2], [9, 0], atol=1e-13) and out[2] == -7
    # test1d_float (:205): GPU C2C gives [9+9i, 0, 0]
    out = _run_iface(fft, fft.generate_fft_interface(fft.int1d, c32, c32), (3,), 3 + 3j)
    assert np.allclose(out, [9 + 9j, z, z], atol=1e-5)
    # test1d_float_real (:171): the reference's exec is commented out; README.md:16 promises the transform
    out = _run_iface(fft, fft.generate_fft_interface(fft.int1d, flt, c32), (3,), 3.0, out_fill=-7)
    assert np.allclose(out[:2], [9, 0], atol=1e-5) and out[2] == -7
    # test2d (:309): output pre-filled with 1
    out = _run_iface(fft, fft.generate_fft_interface(fft.int2d, c64, c64), (2, 2), 5 + 5j, out_fill=1)
    assert np.allclose(out, [20 + 20j, z, z, z], atol=1e-13)
    # test3d (:326)
    out = _run_iface(fft, fft.generate_fft_interface(fft.int3d, c64, c64), (3, 2, 2), 3 + 3j)
    assert np.allclose(out, [36 + 36j] + [z] * 11, atol=1e-13)
    # test3d_batch (:347): 27+27i at flat offsets 0 and 9
    out = _run_iface(fft, fft.generate_fft_interface(fft.int3d, c64, c64), (3, 3, 2), 3 + 3j, batch=True)
    want = np.zeros(18, np.complex128)
    want[0] = want[9] = 27 + 27j
    assert np.allclose(out, want, atol=1e-13)
    # test3d_batch_real (:372): row pitch stays 3; entries 2,5,8,11,14,17 untouched
    out = _run_iface(fft, fft.generate_fft_interface(fft.int3d, dbl, c64), (3, 3, 2), 3.0, out_fill=-7 - 7j, batch=True)
    for i in (2, 5, 8, 11, 14, 17):
        assert out[i] == -7 - 7j
    assert abs(out[0] - 27) < 1e-13 and abs(out[9] - 27) < 1e-13
    mask = np.ones(18, bool)
    mask[[0, 9, 2, 5, 8, 11, 14, 17]] = False
    assert np.allclose(out[mask], 0, atol=1e-13)


def test_distrib_known_answer_single_node(fft, L):
    """test1d_distrib (:282): n nodes x 3, 4+4i -> each shard [12+12i, 0, 0]; one node here."""
    iface = fft.generate_fft_interface(fft.int1d, fft.complex64, fft.complex64)
    n = iface.get_num_nodes()
    r = fft.Region((3 * n,), fft.complex64).fill(4 + 4j)
    s = fft.Region((3 * n,), fft.complex64)
    p = fft.PlanRegion(n)
    rp, sp, pp = r.partition_equal(n), s.partition_equal(n), p.partition_equal(n)
    iface.make_plan_distrib(r, rp, s, sp, p, pp)
    for i in range(n):
        iface.execute_plan_task(rp[i], sp[i], p)
    torch.cuda.synchronize()
    iface.destroy_plan_distrib(p, pp)
    assert np.allclose(s.numpy().reshape(n, 3), [[12 + 12j, 0, 0]] * n, atol=1e-13)


# ------------------------------------------------------------------------------------------
# oracle sweeps
# ------------------------------------------------------------------------------------------
POW2 = [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192]


@pytest.mark.parametrize("kind", ["z2z", "c2c", "d2z", "r2c"])
def test_every_power_of_two_axis_kernel(L, oracle, kind):
    sizes = POW2 + {"z2z": [], "c2c": [16384], "d2z": [16384], "r2c": [16384, 32768]}[kind]
    _, dt_in, _ = _kinds(L)[kind]
    for i, n in enumerate(sizes):
        x = oracle.synth((3, n), dt_in, 300 + i)
        got, desc = gpu_fft(L, kind, x, (n,), batch=3)
        err = oracle.rel_l2(got, cpu_fft(oracle, kind, x, (n,), batch=3))
        assert err <= oracle.tolerance(n, kind in ("c2c", "r2c")), (kind, n, err)
        if n >= 4:
            assert "generic" not in desc, (kind, n, desc)   # powers of two take the smem Stockham path


SHAPES = [((64, 64), 1), ((32, 256), 2), ((256, 32), 1), ((16, 8, 32), 1), ((64, 64, 64), 1), ((128, 4, 512), 1),
          ((8, 1024, 4), 1), ((1, 64, 1), 1), ((4096, 64), 1), ((2048, 32), 2), ((1024, 16, 8), 1), ((8192, 16), 1),
          ((2, 2048, 24), 1), ((2, 2), 1), ((4, 4, 4), 5),
          ((3,), 1), ((5,), 2), ((12,), 1), ((1021,), 1), ((3, 2, 2), 1), ((6, 10, 9), 2), ((7, 16), 1), ((1,), 1),
          ((1 << 15,), 1), ((1 << 18,), 2), ((1 << 20,), 1), ((1 << 22,), 1)]


@pytest.mark.parametrize("kind", ["z2z", "c2c", "d2z", "r2c"])
def test_shapes_against_oracle(L, oracle, kind):
    _, dt_in, _ = _kinds(L)[kind]
    for i, (shape, batch) in enumerate(SHAPES):
        if kind in ("d2z", "r2c") and int(np.prod(shape)) > (1 << 20) and len(shape) == 1:
            continue
        full = ((batch,) if batch > 1 else ()) + shape
        x = oracle.synth(full, dt_in, 400 + i)
        got, _ = gpu_fft(L, kind, x, shape, batch)
        err = oracle.rel_l2(got, cpu_fft(oracle, kind, x, shape, batch))
        assert err <= oracle.tolerance(int(np.prod(shape)), kind in ("c2c", "r2c")), (kind, shape, batch, err)


def test_long_strided_axes_use_cluster_kernels(L, oracle, monkeypatch):
    """strided axes of 1024 and 2048 points run as one 128 KiB tile per CTA, 4096..16384 as thread-block-cluster
    passes (DSMEM cross stage); the cluster kernels for 1024 and 2048 stay selectable"""
    for kind, shape in [("z2z", (1024, 64)), ("z2z", (2048, 8)), ("z2z", (4096, 16)), ("z2z", (8192, 8)),
                        ("c2c", (1024, 32)), ("c2c", (2048, 16)), ("c2c", (4096, 48 // 3)), ("c2c", (8192, 16)),
                        ("c2c", (16384, 8)), ("d2z", (4096, 64)), ("r2c", (2048, 128))]:
        _, dt_in, _ = _kinds(L)[kind]
        x = oracle.synth(shape, dt_in, 600 + shape[0] % 97)
        got, desc = gpu_fft(L, kind, x, shape)
        err = oracle.rel_l2(got, cpu_fft(oracle, kind, x, shape))
        assert err <= oracle.tolerance(int(np.prod(shape)), kind in ("c2c", "r2c")), (kind, shape, err)
        want = {1024: "smem=131072 cluster=1", 2048: "smem=131072 cluster=1", 4096: "cluster=8", 8192: "cluster=8", 16384: "cluster=8"}[shape[0]]
        assert want in desc, desc
        # backward transform through the same kernels
        if kind in ("z2z", "c2c"):
            back, _ = gpu_fft(L, kind, got, shape, direction=+1)
            assert oracle.rel_l2(back / np.prod(shape), x) <= 2 * oracle.tolerance(int(np.prod(shape)), kind == "c2c")
    monkeypatch.setenv("FFTB200_TILE_ALT_COL", "1")      # the cluster-of-2 alternative for L = 1024
    for kind, shape in [("z2z", (1024, 64)), ("c2c", (1024, 32)), ("z2z", (2048, 16)), ("c2c", (2048, 32))]:
        _, dt_in, _ = _kinds(L)[kind]
        x = oracle.synth(shape, dt_in, 650)
        got, desc = gpu_fft(L, kind, x, shape)
        assert ("cluster=2" if shape[0] == 1024 else "cluster=4") in desc, desc
        assert oracle.rel_l2(got, cpu_fft(oracle, kind, x, shape)) <= oracle.tolerance(int(np.prod(shape)), kind == "c2c")


def test_baseline_config1_n1024(L, oracle):
    """BASELINE configs[0]: 1D C2C complex64 N=1024 (the reference's CPU-runnable case)."""
    g = np.load(GOLDEN)
    x, y = g["c1_1d_c64_1024__x"], g["c1_1d_c64_1024__y"]
    got, desc = gpu_fft(L, "z2z", x, (1024,))
    assert oracle.rel_l2(got, y) <= oracle.tolerance(1024, False)
    assert desc.count("\n") == 1 and "tile" in desc


def test_advanced_layout_embeds(L, oracle):
    """non-NULL embeds: fftw_plan_many_dft semantics (api/plan-many-dft.c:43-46), padded rows and
    batch distance; complex and real."""
    n, ie, oe, batch = [8, 16], [10, 20], [9, 24], 3
    idist, odist = 10 * 20 + 7, 9 * 24 + 5
    x = oracle.synth((batch * idist,), np.complex128, 500)
    want = np.full(batch * odist, -1 - 1j, dtype=np.complex128)
    oracle.port_dft_many(n, batch, x, ie, 1, idist, want, oe, 1, odist)
    xd = torch.from_numpy(x).cuda()
    yd = torch.full((batch * odist,), -1 - 1j, dtype=torch.complex128, device="cuda")
    h = L.plan_many(2, n, ie, 1, idist, oe, 1, odist, L.Z2Z, batch)
    L.execute(h, L.Z2Z, xd.data_ptr(), yd.data_ptr())
    torch.cuda.synchronize()
    L.destroy(h)
    got = yd.cpu().numpy()
    assert oracle.rel_l2(got, want) <= oracle.tolerance(128, False)
    assert np.array_equal(got == (-1 - 1j), want == (-1 - 1j))        # padding untouched
    # strided elements (istride 2 / ostride 3) go through the generic path
    xs = oracle.synth((2 * 64,), np.complex128, 501)
    wants = np.zeros(3 * 64, np.complex128)
    oracle.port_dft_many([64], 1, xs, [64], 2, 0, wants, [64], 3, 0)
    xd = torch.from_numpy(xs).cuda()
    yd = torch.zeros(3 * 64, dtype=torch.complex128, device="cuda")
    h = L.plan_many(1, [64], [64], 2, 0, [64], 3, 0, L.Z2Z, 1)
    L.execute(h, L.Z2Z, xd.data_ptr(), yd.data_ptr())
    torch.cuda.synchronize()
    L.destroy(h)
    assert oracle.rel_l2(yd.cpu().numpy(), wants) <= oracle.tolerance(64, False)
    # real, padded
    xr = oracle.synth((4 * 40,), np.float64, 502)
    wantr = np.full(4 * 20, 5 + 5j, dtype=np.complex128)
    oracle.port_r2c_many([32], 4, xr, [36], 1, 40, wantr, [18], 1, 20)
    xd = torch.from_numpy(xr).cuda()
    yd = torch.full((4 * 20,), 5 + 5j, dtype=torch.complex128, device="cuda")
    h = L.plan_many(1, [32], [36], 1, 40, [18], 1, 20, L.D2Z, 4)
    L.execute(h, L.D2Z, xd.data_ptr(), yd.data_ptr())
    torch.cuda.synchronize()
    L.destroy(h)
    assert oracle.rel_l2(yd.cpu().numpy(), wantr) <= oracle.tolerance(32, False)


def test_slab_primitive_batched_2d(L, oracle):
    """make_plan_batch's call on a power-of-two cube: n[2] batched (n0 x n1) transforms at distance
    n0*n1 (src/fft.rg:372-398) — the per-slab building block of the slab decomposition."""
    n0, n1, nb = 64, 128, 6
    x = oracle.synth((nb, n0, n1), np.complex128, 510)
    xd = torch.from_numpy(x).cuda()
    yd = torch.zeros_like(xd)
    h = L.plan_many(2, [n0, n1], [n0, n1], 1, n0 * n1, [n0, n1], 1, n0 * n1, L.Z2Z, nb)
    assert "generic" not in L.describe(h)
    L.execute(h, L.Z2Z, xd.data_ptr(), yd.data_ptr())
    torch.cuda.synchronize()
    L.destroy(h)
    want = cpu_fft(oracle, "z2z", x, (n0, n1), batch=nb)
    assert oracle.rel_l2(yd.cpu().numpy(), want) <= oracle.tolerance(n0 * n1, False)


def test_misaligned_pointers_and_inplace(L, oracle):
    """any naturally aligned pointer is accepted: an 8-byte-aligned complex64 base takes
    the scalar path; in == out is accepted for C2C."""
    n = 256
    x = oracle.synth((n,), np.complex128, 520)
    buf = torch.zeros(2 * n + 1, dtype=torch.float64, device="cuda")
    buf[1:] = torch.from_numpy(x.view(np.float64)).cuda()
    out = torch.zeros(2 * n + 1, dtype=torch.float64, device="cuda")
    h = L.plan_many(1, [n], None, 0, 0, None, 0, 0, L.Z2Z, 1)
    L.execute(h, L.Z2Z, buf.data_ptr() + 8, out.data_ptr() + 8)
    torch.cuda.synchronize()
    got = out[1:].cpu().numpy().view(np.complex128)
    assert oracle.rel_l2(got, cpu_fft(oracle, "z2z", x, (n,))) <= oracle.tolerance(n, False)
    # in place
    xd = torch.from_numpy(x).cuda()
    L.execute(h, L.Z2Z, xd.data_ptr(), xd.data_ptr())
    torch.cuda.synchronize()
    L.destroy(h)
    assert oracle.rel_l2(xd.cpu().numpy(), cpu_fft(oracle, "z2z", x, (n,))) <= oracle.tolerance(n, False)
    # in place, 3-D and four-step
    for shape in [(32, 16, 64), (1 << 16,)]:
        x = oracle.synth(shape, np.complex128, 521)
        xd = torch.from_numpy(x).cuda()
        h = L.plan_many(len(shape), list(shape), None, 0, 0, None, 0, 0, L.Z2Z, 1)
        L.execute(h, L.Z2Z, xd.data_ptr(), xd.data_ptr())
        torch.cuda.synchronize()
        L.destroy(h)
        assert oracle.rel_l2(xd.cpu().numpy(), cpu_fft(oracle, "z2z", x, shape)) <= oracle.tolerance(x.size, False)


def test_error_codes_on_device(L):
    h = L.plan_many(1, [64], None, 0, 0, None, 0, 0, L.Z2Z, 1)
    lib = L.lib()
    buf = torch.zeros(64, dtype=torch.complex128, device="cuda")
    assert lib.fftb200_exec_c2c(h, buf.data_ptr(), buf.data_ptr(), -1) == L.INVALID_TYPE   # wrong exec for the plan
    assert lib.fftb200_exec_z2z(h, None, buf.data_ptr(), -1) == L.INVALID_VALUE
    assert lib.fftb200_exec_z2z(h, buf.data_ptr(), buf.data_ptr(), 0) == L.INVALID_VALUE   # bad direction
    assert lib.fftb200_destroy(h) == 0
    assert lib.fftb200_destroy(h) == L.INVALID_PLAN                                      # double destroy: no crash
    assert lib.fftb200_exec_z2z(h, buf.data_ptr(), buf.data_ptr(), -1) == L.INVALID_PLAN
    hr = L.plan_many(2, [8, 64], None, 0, 0, None, 0, 0, L.D2Z, 1)
    big = torch.zeros(8 * 66, dtype=torch.float64, device="cuda")
    assert lib.fftb200_exec_d2z(hr, big.data_ptr(), big.data_ptr()) == L.INVALID_VALUE     # r2c in place needs padded rows
    L.destroy(hr)


def test_plan_on_user_stream_and_reuse(L, oracle):
    n = (64, 64)
    x = oracle.synth((4,) + n, np.complex128, 530)
    want = cpu_fft(oracle, "z2z", x, n, batch=4)
    h = L.plan_many(2, list(n), None, 0, 0, None, 0, 0, L.Z2Z, 1)
    st = torch.cuda.Stream()
    xd = torch.from_numpy(x).cuda()
    yd = torch.zeros_like(xd)
    torch.cuda.synchronize()
    L.set_stream(h, st.cuda_stream)
    for b in range(4):                                   # one plan, many executes on new arrays (fftw_execute_dft)
        L.execute(h, L.Z2Z, xd[b].data_ptr(), yd[b].data_ptr())
    st.synchronize()
    L.destroy(h)
    assert oracle.rel_l2(yd.cpu().numpy(), want) <= oracle.tolerance(64 * 64, False)


# ------------------------------------------------------------------------------------------
# full-size properties (BASELINE configs 2-4 on one GPU); all arithmetic for the checks on the GPU
# ------------------------------------------------------------------------------------------
def _rel(a, b):
    return float((torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b)).item())


def _exec(L, h, ftype, x, y, direction=-1):
    L.execute(h, ftype, x.data_ptr(), y.data_ptr(), direction)


def test_c4_512cubed_properties(L, oracle):
    """3D C2C complex64 512^3: impulse, linearity, time shift, Parseval, round trip."""
    n = 512
    N = n ** 3
    tol = oracle.tolerance(N, False)
    h = L.plan_many(3, [n, n, n], None, 0, 0, None, 0, 0, L.Z2Z, 1)
    assert L.launch_count(h) in (2, 3)          # 2 when the x and y passes are fused through L2
    g = torch.Generator(device="cuda").manual_seed(1234)
    a = torch.rand(n, n, n, 2, dtype=torch.float64, device="cuda", generator=g).sub_(0.5)
    a = torch.view_as_complex(a)
    A = torch.empty_like(a)
    _exec(L, h, L.Z2Z, a, A)
    torch.cuda.synchronize()
    # Parseval: sum |A|^2 = N sum |a|^2
    pa = float(torch.linalg.vector_norm(a).item()) ** 2
    pA = float(torch.linalg.vector_norm(A).item()) ** 2
    assert abs(pA / (N * pa) - 1) <= tol
    # DC bin = sum of the input
    assert abs(A[0, 0, 0].item() - a.sum().item()) <= tol * abs(a.abs().sum().item())
    # a sampled bin against the O(N) direct sum
    k = (37, 401, 255)
    idx = torch.arange(n, device="cuda", dtype=torch.float64)
    w = [torch.exp(-2j * torch.pi * ((ki * idx) % n) / n) for ki in k]
    direct = torch.einsum("abc,a,b,c->", a, w[0], w[1], w[2])
    assert abs(A[k].item() - direct.item()) <= 4 * tol * float(torch.linalg.vector_norm(a).item())
    # forward -> backward round trip = N * input (in place on A)
    B = torch.empty_like(a)
    _exec(L, h, L.Z2Z, A, B, +1)
    torch.cuda.synchronize()
    assert _rel(B / N, a) <= 2 * tol
    del B
    # time shift along each axis: F(roll(a, 1, axis))[k] = F(a)[k] * exp(-2 pi i k_axis / n)
    for axis in range(3):
        s = torch.roll(a, 1, dims=axis)
        S = torch.empty_like(a)
        _exec(L, h, L.Z2Z, s, S)
        torch.cuda.synchronize()
        shape = [1, 1, 1]
        shape[axis] = n
        ph = torch.exp(-2j * torch.pi * idx / n).reshape(shape)
        assert _rel(S, A * ph) <= tol, axis
        del s, S
    # linearity
    b = torch.view_as_complex(torch.rand(n, n, n, 2, dtype=torch.float64, device="cuda", generator=g).sub_(0.5))
    Bf = torch.empty_like(b)
    _exec(L, h, L.Z2Z, b, Bf)
    c = 2.5 * a - 1.5j * b
    C = torch.empty_like(c)
    _exec(L, h, L.Z2Z, c, C)
    torch.cuda.synchronize()
    assert _rel(C, 2.5 * A - 1.5j * Bf) <= tol
    del b, Bf, c, C
    # impulse at (1,2,3): |F| = 1 everywhere, phase = product of axis phases
    e = torch.zeros_like(a)
    e[1, 2, 3] = 1
    E = torch.empty_like(a)
    _exec(L, h, L.Z2Z, e, E)
    torch.cuda.synchronize()
    want = (torch.exp(-2j * torch.pi * 1 * idx / n).reshape(n, 1, 1) * torch.exp(-2j * torch.pi * 2 * idx / n).reshape(1, n, 1)
            * torch.exp(-2j * torch.pi * 3 * idx / n).reshape(1, 1, n))
    assert _rel(E, want) <= tol
    L.destroy(h)


def test_c4_subcube_against_oracle_128(L, oracle):
    """the same 3-pass plan shape at 128^3, directly against FFTW"""
    x = oracle.synth((128, 128, 128), np.complex128, 540)
    got, desc = gpu_fft(L, "z2z", x, (128, 128, 128))
    assert oracle.rel_l2(got, cpu_fft(oracle, "z2z", x, (128, 128, 128))) <= oracle.tolerance(128 ** 3, False)
    assert desc.count("\n") in (2, 3)


def test_c2_4096sq_r2c_properties(L, oracle):
    """2D R2C double -> complex64 4096 x 4096: against the complex transform of the same data on the
    GPU (Hermitian half), Parseval over the half spectrum, DC, plus one row block against FFTW."""
    n = 4096
    tol = oracle.tolerance(n * n, False)
    g = torch.Generator(device="cuda").manual_seed(99)
    x = torch.rand(n, n, dtype=torch.float64, device="cuda", generator=g).sub_(0.5)
    y = torch.empty(n, n // 2 + 1, dtype=torch.complex128, device="cuda")
    h = L.plan_many(2, [n, n], None, 0, 0, None, 0, 0, L.D2Z, 1)
    assert L.launch_count(h) == 2
    _exec(L, h, L.D2Z, x, y)
    hz = L.plan_many(2, [n, n], None, 0, 0, None, 0, 0, L.Z2Z, 1)
    xc = x.to(torch.complex128)
    yc = torch.empty_like(xc)
    _exec(L, hz, L.Z2Z, xc, yc)
    torch.cuda.synchronize()
    assert _rel(y, yc[:, : n // 2 + 1]) <= tol
    # Hermitian symmetry of the full transform: yc[-i, -j] = conj(yc[i, j])
    flip = torch.roll(torch.flip(yc, dims=(0, 1)), shifts=(1, 1), dims=(0, 1))
    assert _rel(flip, yc.conj()) <= tol
    assert abs(y[0, 0].item() - x.sum().item()) <= tol * float(x.abs().sum().item())
    L.destroy(h)
    L.destroy(hz)
    del xc, yc, flip
    # 1-D batched rows against FFTW: first axis kernel alone
    rows = x[:8].cpu().numpy()
    got, _ = gpu_fft(L, "d2z", rows, (n,), batch=8)
    assert oracle.rel_l2(got, cpu_fft(oracle, "d2z", rows, (n,), batch=8)) <= oracle.tolerance(n, False)


def test_c2_small_2d_r2c_against_oracle(L, oracle):
    x = oracle.synth((512, 512), np.float64, 541)
    got, _ = gpu_fft(L, "d2z", x, (512, 512))
    assert oracle.rel_l2(got, cpu_fft(oracle, "d2z", x, (512, 512))) <= oracle.tolerance(512 * 512, False)


def test_c3_2pow27_c32_properties(L, oracle):
    """1D C2C complex32 N = 2^27 (four/six-step): Parseval, sampled bins against a direct fp64 sum,
    time shift, round trip."""
    N = 1 << 27
    tol = oracle.tolerance(N, True)
    g = torch.Generator(device="cuda").manual_seed(7)
    a = torch.view_as_complex(torch.rand(N, 2, dtype=torch.float32, device="cuda", generator=g).sub_(0.5))
    A = torch.empty_like(a)
    h = L.plan_many(1, [N], None, 0, 0, None, 0, 0, L.C2C, 1)
    _exec(L, h, L.C2C, a, A)
    torch.cuda.synchronize()
    na = float(torch.linalg.vector_norm(a.to(torch.complex128)).item())
    nA = float(torch.linalg.vector_norm(A.to(torch.complex128)).item())
    assert abs(nA * nA / (N * na * na) - 1) <= tol
    a64 = a.to(torch.complex128)
    idx = torch.arange(N, device="cuda", dtype=torch.int64)
    for k in (0, 1, 12345677, N // 2, N - 1):
        ph = torch.exp(-2j * torch.pi * ((idx * k) % N).to(torch.float64) / N)
        direct = (a64 * ph).sum().item()
        assert abs(A[k].item() - direct) <= tol * na, k
    del a64, idx, ph
    s = torch.roll(a, 1)
    S = torch.empty_like(a)
    _exec(L, h, L.C2C, s, S)
    torch.cuda.synchronize()
    phase = torch.exp(-2j * torch.pi * torch.arange(N, device="cuda", dtype=torch.float64) / N).to(torch.complex64)
    assert _rel((S).to(torch.complex128), (A * phase).to(torch.complex128)) <= tol
    del s, S, phase
    B = torch.empty_like(a)
    _exec(L, h, L.C2C, A, B, +1)
    torch.cuda.synchronize()
    assert _rel((B / N).to(torch.complex128), a.to(torch.complex128)) <= 2 * tol
    L.destroy(h)


def test_four_step_against_oracle_2pow22(L, oracle):
    x = oracle.synth((1 << 22,), np.complex64, 550)
    got, desc = gpu_fft(L, "c2c", x, (1 << 22,))
    assert oracle.rel_l2(got, cpu_fft(oracle, "c2c", x, (1 << 22,))) <= oracle.tolerance(1 << 22, True)
    assert "step" in desc


def test_native_library_is_the_one_running(fft, L):
    """the product path is libfft_b200.so in-tree; the launch list shows hand-written kernels"""
    assert os.path.samefile(L.LIB_PATH, os.path.join(os.path.dirname(fft.__file__), "libfft_b200.so"))
    maps = open("/proc/self/maps").read()
    assert "libfft_b200.so" in maps
    import subprocess
    needed = subprocess.run(["readelf", "-d", L.LIB_PATH], capture_output=True, text=True).stdout
    assert "NEEDED" in needed and "cufft" not in needed.lower() and "nccl" not in needed.lower(), needed
    h = L.plan_many(3, [512, 512, 512], None, 0, 0, None, 0, 0, L.Z2Z, 1)
    nl = L.launch_count(h)
    assert nl in (2, 3) and L.work_size(h) in (0, 512 ** 3 * 16)     # blocked intermediate layout uses one work buffer
    total = sum(L.launch_bytes(h, i) for i in range(nl))
    # compulsory HBM bytes: one read + one write of 512^3 complex64 per launch (a fused launch covers two
    # axis passes of SURVEY.md §8d's pass model with the traffic of one)
    assert total == nl * 2 * 512 ** 3 * 16
    L.destroy(h)


def test_strided_1024_axes_in_place_at_scale(L, oracle):
    """A 1024-point strided pass run in place over far more tiles than fit the GPU at once (round 1's two-CTA split
    kernel raced here: its CTAs read the whole tile and stored half of it each).  1024 x 1024 x 256 complex64, all
    three passes, out of place against in place (bit-identical) and against FFTW."""
    shape = (1024, 1024, 256)
    x = torch.from_numpy(oracle.synth(shape, np.complex128, 777)).cuda()
    y = torch.empty_like(x)
    h = L.plan_many(3, list(shape), None, 0, 0, None, 0, 0, L.Z2Z, 1)
    desc = L.describe(h)
    assert desc.count("L=1024") == 2 and "cluster=1" in desc, desc
    L.execute(h, L.Z2Z, x.data_ptr(), y.data_ptr())
    xi = x.clone()
    L.execute(h, L.Z2Z, xi.data_ptr(), xi.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(xi, y), "in-place result differs from the out-of-place result"
    del xi
    want = oracle.FFTW.get("ref").dft(x.cpu().numpy(), threads=min(32, os.cpu_count() or 1))
    err = oracle.rel_l2(y.cpu().numpy(), want)
    L.destroy(h)
    assert err <= oracle.tolerance(int(np.prod(shape)), False), err


def test_concurrent_plans_from_many_threads(L, oracle):
    """Legion runs one task per GPU processor concurrently in one process: plan creation,
    execution on private streams and destruction from 8 host threads at once."""
    import threading
    shapes = [("z2z", (64, 64, 32)), ("c2c", (4096,)), ("d2z", (128, 256)), ("z2z", (1 << 16,)),
              ("r2c", (32, 32, 64)), ("z2z", (12, 10)), ("c2c", (256, 64)), ("d2z", (2048,))]
    inputs = [oracle.synth(sh, _kinds(L)[k][1], 900 + i) for i, (k, sh) in enumerate(shapes)]
    wants = [cpu_fft(oracle, k, x, sh) for (k, sh), x in zip(shapes, inputs)]
    errors = [None] * len(shapes)

    def worker(i):
        try:
            kind, shape = shapes[i]
            ftype, dt_in, dt_out = _kinds(L)[kind]
            real = kind in ("d2z", "r2c")
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                xd = torch.from_numpy(inputs[i]).cuda()
                oshape = shape[:-1] + (shape[-1] // 2 + 1,) if real else shape
                yd = torch.zeros(oshape, dtype=_torch_dtype(dt_out), device="cuda")
            st.synchronize()
            for _ in range(5):
                h = L.plan_many(len(shape), list(shape), None, 0, 0, None, 0, 0, ftype, 1)
                L.set_stream(h, st.cuda_stream)
                L.execute(h, ftype, xd.data_ptr(), yd.data_ptr())
                st.synchronize()
                L.destroy(h)
            err = oracle.rel_l2(yd.cpu().numpy(), wants[i])
            tol = oracle.tolerance(int(np.prod(shape)), kind in ("c2c", "r2c"))
            errors[i] = None if err <= tol else f"{kind} {shape}: {err:.2e} > {tol:.2e}"
        except Exception as ex:  # surfaced below
            errors[i] = repr(ex)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(shapes))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert all(e is None for e in errors), errors


def test_destroy_racing_exec_is_safe(L, oracle):
    """destroy_plan runs on a CPU processor while a GPU task may still be inside execute_plan on the same handle
    (src/fft.rg:613-617 vs 624-645): every ABI call holds a reference for its duration, so the racing calls either
    complete or report FFTB200_INVALID_PLAN — never a crash — and a later exec on the stale handle is refused."""
    import threading
    shape = (64, 64, 64)
    x = torch.from_numpy(oracle.synth(shape, np.complex128, 950)).cuda()
    y = torch.empty_like(x)
    for round_ in range(20):
        h = L.plan_many(3, list(shape), None, 0, 0, None, 0, 0, L.Z2Z, 1)
        stop = threading.Event()
        outcomes = []

        def hammer():
            lib = L.lib()
            while not stop.is_set():
                rc = lib.fftb200_exec_z2z(h, x.data_ptr(), y.data_ptr(), -1)
                outcomes.append(rc)
                buf = ctypes.create_string_buffer(4096)
                outcomes.append(lib.fftb200_describe(h, buf, 4096))
                if rc != 0:
                    break

        ts = [threading.Thread(target=hammer) for _ in range(3)]
        for t in ts:
            t.start()
        while len(outcomes) < 10:
            pass
        L.destroy(h)
        stop.set()
        for t in ts:
            t.join()
        assert set(outcomes) <= {L.SUCCESS, L.INVALID_PLAN}, set(outcomes)
        with pytest.raises(L.FFTB200Error):
            L.execute(h, L.Z2Z, x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    want = cpu_fft(oracle, "z2z", x.cpu().numpy(), shape)
    h = L.plan_many(3, list(shape), None, 0, 0, None, 0, 0, L.Z2Z, 1)
    L.execute(h, L.Z2Z, x.data_ptr(), y.data_ptr())
    torch.cuda.synchronize()
    L.destroy(h)
    assert oracle.rel_l2(y.cpu().numpy(), want) <= oracle.tolerance(64 ** 3, False)


def test_host_memory_regions_are_staged(L, oracle):
    """The reference's mapper puts regions in zero-copy (pinned host) memory (test/test_mapper.cc:45-58):
    host pointers — pinned or pageable — go through the plan's HBM staging and give the same bits."""
    x = oracle.synth((64, 32, 128), np.complex128, 910)
    xd = torch.from_numpy(x).cuda()
    yd = torch.zeros_like(xd)
    h = L.plan_many(3, [64, 32, 128], None, 0, 0, None, 0, 0, L.Z2Z, 1)
    L.execute(h, L.Z2Z, xd.data_ptr(), yd.data_ptr())
    hx = torch.from_numpy(x).pin_memory()
    hy = torch.zeros(x.shape, dtype=torch.complex128).pin_memory()
    L.execute(h, L.Z2Z, hx.data_ptr(), hy.data_ptr())
    px = torch.from_numpy(x.copy())                          # pageable
    py = torch.zeros(x.shape, dtype=torch.complex128)
    L.execute(h, L.Z2Z, px.data_ptr(), py.data_ptr())
    torch.cuda.synchronize()
    L.destroy(h)
    assert torch.equal(hy, yd.cpu()) and torch.equal(py, yd.cpu())
    assert np.array_equal(hx.numpy(), x) and np.array_equal(px.numpy(), x)
    # padded advanced layout from host memory: padding must survive the round trip
    n, ie, oe, batch, idist, odist = [8, 16], [10, 20], [9, 24], 2, 207, 221
    xs = oracle.synth((batch * idist,), np.complex128, 911)
    want = np.full(batch * odist, -1 - 1j, dtype=np.complex128)
    oracle.port_dft_many(n, batch, xs, ie, 1, idist, want, oe, 1, odist)
    hxs = torch.from_numpy(xs).pin_memory()
    hys = torch.full((batch * odist,), -1 - 1j, dtype=torch.complex128).pin_memory()
    h = L.plan_many(2, n, ie, 1, idist, oe, 1, odist, L.Z2Z, batch)
    L.execute(h, L.Z2Z, hxs.data_ptr(), hys.data_ptr())
    torch.cuda.synchronize()
    L.destroy(h)
    got = hys.numpy()
    assert oracle.rel_l2(got, want) <= oracle.tolerance(128, False)
    assert np.array_equal(got == (-1 - 1j), want == (-1 - 1j))


def test_random_plans_against_oracle(L, oracle):
    """80 seeded random cufftPlanMany-style plans: rank 1-3, power-of-two and mixed-radix sizes, batches, padded
    embeds and batch distances, all four types — against the oracle's fftw_plan_many_dft(_r2c) restatement;
    padding must stay untouched and the input unchanged."""
    rng = np.random.default_rng(20261018)
    sizes_p2 = [2, 4, 8, 16, 32, 64, 128, 256]
    sizes_any = [1, 3, 5, 6, 7, 9, 10, 12, 15, 20, 24, 30, 48, 96, 100]
    for case in range(80):
        kind = ["z2z", "c2c", "d2z", "r2c"][case % 4]
        ftype, dt_in, dt_out = _kinds(L)[kind]
        real, single = kind in ("d2z", "r2c"), kind in ("c2c", "r2c")
        rank = int(rng.integers(1, 4))
        pool = sizes_p2 if rng.random() < 0.6 else sizes_p2 + sizes_any
        n = [int(rng.choice(pool)) for _ in range(rank)]
        while int(np.prod(n)) > (1 << 18):
            n[int(np.argmax(n))] //= 2
        n = [max(1, v) for v in n]
        if real and n[-1] < 2:
            n[-1] = 2
        batch = int(rng.integers(1, 5))
        nout = n[:-1] + [n[-1] // 2 + 1] if real else list(n)
        padded = rng.random() < 0.5
        if padded:
            ie = [n[0]] + [v + int(rng.integers(0, 4)) for v in n[1:]]
            oe = [nout[0]] + [v + int(rng.integers(0, 4)) for v in nout[1:]]
            if real and rng.random() < 0.5:
                ie[-1] = max(ie[-1], 2 * oe[-1])            # the in-place-style padded real layout
            idist = int(np.prod(ie)) + int(rng.integers(0, 6))
            odist = int(np.prod(oe)) + int(rng.integers(0, 6))
        else:
            ie, oe = list(n), list(nout)
            idist, odist = int(np.prod(ie)), int(np.prod(oe))
        x = oracle.synth((batch * idist,), dt_in, 1000 + case)
        fill = -3 - 4j
        want = np.full(batch * odist, fill, dtype=np.complex128)
        x64 = x.astype(np.float64 if real else np.complex128)
        if real:
            oracle.port_r2c_many(n, batch, x64, ie, 1, idist, want, oe, 1, odist)
        else:
            oracle.port_dft_many(n, batch, x64, ie, 1, idist, want, oe, 1, odist)
        xd = torch.from_numpy(x).cuda()
        yd = torch.full((batch * odist,), fill, dtype=_torch_dtype(dt_out), device="cuda")
        if padded or batch > 1:
            h = L.plan_many(rank, n, ie, 1, idist, oe, 1, odist, ftype, batch)
        else:
            h = L.plan_many(rank, n, None, 0, 0, None, 0, 0, ftype, 1)
        L.execute(h, ftype, xd.data_ptr(), yd.data_ptr())
        torch.cuda.synchronize()
        L.destroy(h)
        got = yd.cpu().numpy().astype(np.complex128)
        touched = want != fill
        tol = oracle.tolerance(int(np.prod(n)), single)
        info = (case, kind, n, batch, ie, oe, idist, odist)
        assert np.array_equal(got[~touched], want[~touched]), ("padding written", info)
        if touched.any() and np.linalg.norm(want[touched]) > 0:
            assert oracle.rel_l2(got[touched], want[touched]) <= tol, (oracle.rel_l2(got[touched], want[touched]), info)
        assert np.array_equal(xd.cpu().numpy(), x), ("input modified", info)


def test_inverse_real_transforms(L, oracle):
    """C2R / Z2D (no reference call site): unnormalised inverse of the R2C / D2Z results,
    against numpy's irfftn of the oracle's half spectrum and as a GPU round trip; the input is preserved."""
    cases = [("z2d", (8,)), ("z2d", (4096,)), ("z2d", (16384,)), ("z2d", (64, 128)), ("z2d", (16, 32, 64)), ("z2d", (4, 4, 4)),
             ("c2r", (1024,)), ("c2r", (32768,)), ("c2r", (256, 64)), ("c2r", (8, 16, 128)), ("z2d", (2, 1024, 8)),
             ("z2d", (1024, 64)), ("c2r", (2048, 32))]
    for i, (kind, shape) in enumerate(cases):
        single = kind == "c2r"
        rdt, cdt = (np.float32, np.complex64) if single else (np.float64, np.complex128)
        ftype, fwd = (L.C2R, L.R2C) if single else (L.Z2D, L.D2Z)
        x = oracle.synth(shape, rdt, 1200 + i)
        X = oracle.port_r2c(x.astype(np.float64)).astype(cdt)           # half spectrum [.., n/2+1]
        n_total = int(np.prod(shape))
        Xd = torch.from_numpy(X).cuda()
        yd = torch.zeros(shape, dtype=_torch_dtype(rdt), device="cuda")
        h = L.plan_many(len(shape), list(shape), None, 0, 0, None, 0, 0, ftype, 1)
        L.execute(h, ftype, Xd.data_ptr(), yd.data_ptr())
        torch.cuda.synchronize()
        desc = L.describe(h)
        L.destroy(h)
        assert "c2r-row" in desc and "generic" not in desc
        want = np.fft.irfftn(X.astype(np.complex128), s=shape, axes=tuple(range(len(shape)))) * n_total
        tol = oracle.tolerance(n_total, single)
        assert oracle.rel_l2(yd.cpu().numpy(), want) <= tol, (kind, shape, oracle.rel_l2(yd.cpu().numpy(), want))
        assert np.array_equal(Xd.cpu().numpy(), X), "c2r input modified"
        # round trip on the GPU: C2R(R2C(x)) = N x
        xd = torch.from_numpy(x).cuda()
        Zd = torch.zeros(X.shape, dtype=_torch_dtype(cdt), device="cuda")
        hf = L.plan_many(len(shape), list(shape), None, 0, 0, None, 0, 0, fwd, 1)
        L.execute(hf, fwd, xd.data_ptr(), Zd.data_ptr())
        hb = L.plan_many(len(shape), list(shape), None, 0, 0, None, 0, 0, ftype, 1)
        L.execute(hb, ftype, Zd.data_ptr(), yd.data_ptr())
        torch.cuda.synchronize()
        L.destroy(hf)
        L.destroy(hb)
        assert oracle.rel_l2(yd.cpu().numpy() / n_total, x.astype(np.float64)) <= 2 * tol, (kind, shape, "round trip")
    # batched + error conventions
    h = L.plan_many(1, [256], None, 0, 0, None, 0, 0, L.Z2D, 3)
    x = oracle.synth((3, 256), np.float64, 1300)
    X = np.stack([oracle.port_r2c(r) for r in x])
    Xd = torch.from_numpy(X).cuda()
    yd = torch.zeros((3, 256), dtype=torch.float64, device="cuda")
    L.execute(h, L.Z2D, Xd.data_ptr(), yd.data_ptr())
    torch.cuda.synchronize()
    assert oracle.rel_l2(yd.cpu().numpy() / 256, x) <= 2 * oracle.tolerance(256, False)
    lib = L.lib()
    assert lib.fftb200_exec_d2z(h, Xd.data_ptr(), yd.data_ptr()) == L.INVALID_TYPE
    L.destroy(h)
    # sizes that are not powers of two (odd last dimensions included): products of 2, 3, 5, 7, 11, 13 run the mixed-radix kernels
    # (Hermitian completion while loading the last axis), anything else the generic path (Hermitian completion of the
    # half spectrum, backward complex stages - Bluestein for large primes - real part out)
    for i, (kind, shape, tag) in enumerate([("z2d", (12,), "mixed-radix c2r-row"), ("z2d", (12, 10), "mixed-radix c2r-row"),
                                            ("c2r", (3, 5, 6), "mixed-radix c2r-row"), ("z2d", (1000,), "mixed-radix c2r-row"),
                                            ("z2d", (7, 33), "mixed-radix c2r-row"), ("z2d", (7, 34), "Hermitian"), ("z2d", (5, 4, 9), "mixed-radix c2r-row"),
                                            ("z2d", (1021,), "Hermitian"), ("c2r", (6, 127), "Hermitian"),
                                            ("z2d", (9, 7), "mixed-radix c2r-row"), ("z2d", (96, 100, 90), "mixed-radix c2r-row"),
                                            ("c2r", (15, 1001), "mixed-radix c2r-row"), ("z2d", (6, 1024), "c2r-row"),
                                            ("z2d", (128, 6), "mixed-radix c2r-row"), ("z2d", (1000000,), "c2r even/odd pre-pass"),
                                            ("z2d", (4, 100000), "c2r even/odd pre-pass"), ("c2r", (6, 10, 60000), "c2r even/odd pre-pass"),
                                            ("z2d", (10000,), "half-length")]):
        single = kind == "c2r"
        rdt, cdt = (np.float32, np.complex64) if single else (np.float64, np.complex128)
        ftype = L.C2R if single else L.Z2D
        x = oracle.synth(shape, rdt, 1400 + i)
        X = np.fft.rfftn(x.astype(np.float64)).astype(cdt)
        n_total = int(np.prod(shape))
        Xd = torch.from_numpy(np.ascontiguousarray(X)).cuda()
        yd = torch.zeros(shape, dtype=_torch_dtype(rdt), device="cuda")
        h = L.plan_many(len(shape), list(shape), None, 0, 0, None, 0, 0, ftype, 1)
        L.execute(h, ftype, Xd.data_ptr(), yd.data_ptr())
        torch.cuda.synchronize()
        desc = L.describe(h)
        L.destroy(h)
        assert tag in desc, (shape, desc)
        err = oracle.rel_l2(yd.cpu().numpy() / n_total, x.astype(np.float64))
        assert err <= 2 * oracle.tolerance(n_total, single), (kind, shape, err)
        assert np.array_equal(Xd.cpu().numpy(), X), "c2r input modified"


def test_long_mixed_radix_lines_run_two_passes(L, oracle):
    """1-D lengths 2^a 3^b 5^c 7^d 11^e 13^f too long for one shared-memory tile (10^4 ... 2 * 10^6 points): L = N1 N2, a
    strided pass with the w_L^(k1 n2) twiddle into a work buffer, then a contiguous pass with a transposing store
    (fftw-3.3.8/dft/ct.c is the CPU path's decomposition) - two HBM round trips instead of one per prime factor.
    Forward against FFTW, round trip, batches, in place."""
    for kind, shape, batch in [("z2z", (10000,), 1), ("z2z", (100000,), 1), ("z2z", (1000000,), 1), ("c2c", (1000000,), 1),
                               ("z2z", (20000,), 3), ("c2c", (65536 * 3,), 2), ("z2z", (7 * 11 * 13 * 100,), 1),
                               ("z2z", (2000000,), 1), ("z2z", (6561 * 2,), 5)]:
        _, dt_in, _ = _kinds(L)[kind]
        x = oracle.synth(((batch,) if batch > 1 else ()) + shape, dt_in, 1700 + batch)
        got, desc = gpu_fft(L, kind, x, shape, batch=batch)
        assert "col+twiddle" in desc and "row->col" in desc and "generic" not in desc, (shape, desc)
        tol = oracle.tolerance(int(np.prod(shape)), kind == "c2c")
        err = oracle.rel_l2(got, cpu_fft(oracle, kind, x, shape, batch=batch))
        assert err <= tol, (kind, shape, batch, err)
        back, _ = gpu_fft(L, kind, got, shape, batch=batch, direction=+1)
        assert oracle.rel_l2(back / np.prod(shape), x) <= 2 * tol, (kind, shape, "round trip")
    # long axes inside 2-D / 3-D shapes, in any position, power-of-two ones included (no tile kernel reaches 2^15 fp64)
    for kind, shape in [("z2z", (20000, 6)), ("z2z", (6, 20000)), ("z2z", (32768, 8)), ("c2c", (8, 65536)), ("z2z", (4, 10000, 6)),
                        ("z2z", (10000, 3, 4)), ("z2z", (12000, 14)), ("d2z", (20000, 12)), ("z2z", (7000, 9)),
                        ("d2z", (1000000,)), ("r2c", (2000000,)), ("d2z", (3, 200000)), ("d2z", (7000, 14000)), ("d2z", (10000,))]:
        _, dt_in, _ = _kinds(L)[kind]
        x = oracle.synth(shape, dt_in, 1720)
        got, desc = gpu_fft(L, kind, x, shape)
        assert ("two-pass axis" in desc or shape == (10000,)) and "generic" not in desc, (shape, desc)
        if kind in ("d2z", "r2c") and shape[-1] >= 100000:
            assert "even/odd post-pass" in desc, desc
        tol = oracle.tolerance(int(np.prod(shape)), kind in ("c2c", "r2c"))
        err = oracle.rel_l2(got, cpu_fft(oracle, kind, x, shape))
        assert err <= tol, (kind, shape, err, desc)
        if kind in ("z2z", "c2c"):
            back, _ = gpu_fft(L, kind, got, shape, direction=+1)
            assert oracle.rel_l2(back / np.prod(shape), x) <= 2 * tol, (kind, shape, "round trip")
    # in place
    ftype, dt_in, _ = _kinds(L)["z2z"]
    x = oracle.synth((30000,), dt_in, 1710)
    buf = torch.from_numpy(x.copy()).cuda()
    h = L.plan_many(1, [30000], None, 0, 0, None, 0, 0, ftype, 1)
    L.execute(h, ftype, buf.data_ptr(), buf.data_ptr())
    torch.cuda.synchronize()
    L.destroy(h)
    assert oracle.rel_l2(buf.cpu().numpy(), cpu_fft(oracle, "z2z", x, (30000,))) <= oracle.tolerance(30000, False)


def test_mixed_radix_randomized_shapes(L, oracle):
    """Seeded sweep over 1-3-D shapes with axes drawn from products of 2, 3, 5, 7, 11, 13 (ragged tiles, one to five
    stages, single-radix axes, batches), every transform kind, against the oracle.  Whatever plan the builder picks must
    be right; the sweep also counts how many of them took the mixed-radix kernels."""
    rng = np.random.default_rng(20260)
    axes = [3, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15, 18, 20, 21, 22, 24, 26, 28, 30, 33, 35, 36, 39, 40, 42, 45, 48, 50, 54,
            55, 56, 60, 63, 65, 66, 70, 72, 75, 77, 80, 84, 90, 91, 96, 98, 99, 100, 104, 108, 110, 112, 117, 120, 125, 126,
            130, 132, 135, 140, 143, 144, 147, 150, 154, 156, 160, 162, 165, 168, 169, 175, 176, 180, 182, 189, 192, 195,
            196, 198, 200, 208, 210, 216, 220, 224, 225, 231, 234, 240, 242, 243, 245, 250, 252, 260, 264, 270, 273, 275,
            280, 286, 288, 294, 297, 300, 343, 363, 375, 390, 420, 441, 462, 480, 500, 507, 539, 546, 600, 625, 630, 637,
            686, 720, 726, 729, 750, 840, 845, 847, 875, 900, 945, 960, 1000, 1001, 1014, 1029, 1050, 1078, 1080, 1125,
            1155, 1183, 1200, 1250, 1260, 1331, 1350, 1372, 1440, 1452, 1500, 1521, 1536, 1575, 1694, 1715, 1800, 1859, 2000,
            2197, 2310, 2401, 2500, 2520, 3000, 3125, 3430, 3600, 3993, 4000, 4116, 4375, 5000, 5040, 6000, 6250]
    n_mixed = 0
    for case in range(72):
        rank = int(rng.integers(1, 4))
        cap = {1: 6250, 2: 400, 3: 60}[rank]
        pool = [a for a in axes if a <= cap]
        shape = tuple(int(rng.choice(pool)) for _ in range(rank))
        kind = ["z2z", "c2c", "d2z", "r2c"][case % 4]
        batch = int(rng.choice([1, 1, 2, 5])) if np.prod(shape) < 50000 else 1
        _, dt_in, _ = _kinds(L)[kind]
        x = oracle.synth(((batch,) if batch > 1 else ()) + shape, dt_in, 3000 + case)
        got, desc = gpu_fft(L, kind, x, shape, batch=batch)
        n_mixed += "mixed-radix" in desc
        tol = oracle.tolerance(int(np.prod(shape)), kind in ("c2c", "r2c"))
        err = oracle.rel_l2(got, cpu_fft(oracle, kind, x, shape, batch=batch))
        assert err <= tol, (kind, shape, batch, err, desc)
        if kind in ("z2z", "c2c"):
            back, _ = gpu_fft(L, kind, got, shape, batch=batch, direction=+1)
            assert oracle.rel_l2(back / np.prod(shape), x) <= 2 * tol, (kind, shape, batch, "round trip")
    assert n_mixed >= 60, n_mixed


def test_c2r_mixed_radix_half_and_full_length_forms_agree(L, oracle):
    """Z2D / C2R of even sizes 2^a 3^b 5^c 7^d: the half-length form (default) and the full-length form give the same
    reals; odd sizes only have the full-length form."""
    import os
    for kind, shape in [("z2d", (1000,)), ("z2d", (12, 10)), ("c2r", (6, 10, 14)), ("z2d", (96, 100, 90)), ("z2d", (4, 6))]:
        single = kind == "c2r"
        rdt, cdt = (np.float32, np.complex64) if single else (np.float64, np.complex128)
        ftype = L.C2R if single else L.Z2D
        x = oracle.synth(shape, rdt, 1450)
        X = np.ascontiguousarray(np.fft.rfftn(x.astype(np.float64)).astype(cdt))
        n_total = int(np.prod(shape))
        for half in ("1", "0"):
            os.environ["FFTB200_MIXED_HALF"] = half
            try:
                h = L.plan_many(len(shape), list(shape), None, 0, 0, None, 0, 0, ftype, 1)
            finally:
                os.environ.pop("FFTB200_MIXED_HALF", None)
            Xd = torch.from_numpy(X).cuda()
            yd = torch.zeros(shape, dtype=_torch_dtype(rdt), device="cuda")
            L.execute(h, ftype, Xd.data_ptr(), yd.data_ptr())
            torch.cuda.synchronize()
            desc = L.describe(h)
            L.destroy(h)
            assert ("half-length" in desc) == (half == "1"), (shape, half, desc)
            err = oracle.rel_l2(yd.cpu().numpy() / n_total, x.astype(np.float64))
            assert err <= 2 * oracle.tolerance(n_total, single), (kind, shape, half, err)
            assert np.array_equal(Xd.cpu().numpy(), X), "c2r input modified"


def test_c2r_in_place_other_sizes(L, oracle):
    """Z2D in place (half spectrum overwritten by the padded real rows) for sizes off the power-of-two path: mixed-radix
    plans where the layouts coincide, the generic plan otherwise - the call must work either way."""
    for shape in [(1000,), (12, 10), (6, 10, 14), (7, 33)]:
        nl, ncol = shape[-1], shape[-1] // 2 + 1
        x = oracle.synth(shape, np.float64, 1500 + len(shape))
        X = np.ascontiguousarray(np.fft.rfftn(x))
        buf = torch.from_numpy(X.copy()).cuda()
        inembed, onembed = list(shape[:-1]) + [ncol], list(shape[:-1]) + [2 * ncol]
        h = L.plan_many(len(shape), list(shape), inembed, 1, int(np.prod(inembed)), onembed, 1, int(np.prod(onembed)), L.Z2D, 1)
        L.execute(h, L.Z2D, buf.data_ptr(), buf.data_ptr())
        torch.cuda.synchronize()
        L.destroy(h)
        got = buf.cpu().numpy().view(np.float64).reshape(shape[:-1] + (2 * ncol,))[..., :nl]
        n_total = int(np.prod(shape))
        assert oracle.rel_l2(got / n_total, x) <= 2 * oracle.tolerance(n_total, False), shape


def test_large_prime_lengths_use_bluestein(L, oracle):
    """Lengths with a prime factor above 31 run as Bluestein convolutions through the power-of-two tile passes
    (fftw-3.3.8/dft/bluestein.c is the CPU path's counterpart) instead of O(L * p) radix-p stages; small primes
    (the reference's own 3, 5, {3,3,2}: test/fft_test.rg:143,247,328,349) keep the radix stages."""
    for kind, shape in [("z2z", (1021,)), ("z2z", (2039,)), ("z2z", (4093,)), ("c2c", (8191,)), ("d2z", (1021,)),
                        ("z2z", (4, 521)), ("z2z", (97, 6)), ("z2z", (5, 67, 3)), ("r2c", (3, 1009)), ("z2z", (2 * 127, 4))]:
        _, dt_in, _ = _kinds(L)[kind]
        x = oracle.synth(shape, dt_in, 990 + len(shape))
        got, desc = gpu_fft(L, kind, x, shape)
        assert "bluestein" in desc, desc
        err = oracle.rel_l2(got, cpu_fft(oracle, kind, x, shape))
        assert err <= oracle.tolerance(int(np.prod(shape)), kind in ("c2c", "r2c")), (kind, shape, err)
        if kind in ("z2z", "c2c"):
            back, _ = gpu_fft(L, kind, got, shape, direction=+1)
            assert oracle.rel_l2(back / np.prod(shape), x) <= 2 * oracle.tolerance(int(np.prod(shape)), kind == "c2c")
    for kind, shape in [("z2z", (17, 19)), ("z2z", (2 * 3 * 17,)), ("z2z", (31 * 4,)), ("d2z", (3, 17, 2)), ("d2z", (1700,))]:
        _, dt_in, _ = _kinds(L)[kind]
        x = oracle.synth(shape, dt_in, 995)
        got, desc = gpu_fft(L, kind, x, shape)
        assert "bluestein" not in desc and "generic stage" in desc, desc
        assert oracle.rel_l2(got, cpu_fft(oracle, kind, x, shape)) <= oracle.tolerance(int(np.prod(shape)), False)


def test_mixed_radix_lengths_run_one_kernel_per_axis(L, oracle):
    """Transforms whose axes are products of 2, 3, 5, 7, 11 and 13 (the reference's own 3, 5, {3,2,2}, {3,3,2}:
    test/fft_test.rg:143,247,328,349; FFTW's n1_3 / n1_5 / n1_7 codelets on the CPU path) run as ONE shared-memory
    kernel per axis (mixed_kernel.cuh), power-of-two axes of such shapes on the tuned tile kernels.  Forward against
    FFTW, backward round trip, in place, batched, both precisions."""
    cases = [("z2z", (3,)), ("z2z", (5,)), ("z2z", (6,)), ("z2z", (7,)), ("z2z", (12,)), ("z2z", (96,)), ("z2z", (100,)),
             ("z2z", (360,)), ("z2z", (1000,)), ("z2z", (1536,)), ("z2z", (5040,)), ("z2z", (6000,)), ("c2c", (3 * 4096,)),
             ("z2z", (3, 2, 2)), ("z2z", (3, 3, 2)), ("z2z", (96, 96)), ("z2z", (100, 60)), ("z2z", (7, 1024)),
             ("z2z", (1024, 9)), ("c2c", (45, 50)), ("z2z", (96, 96, 96)), ("z2z", (60, 64, 100)), ("c2c", (30, 42, 70)),
             ("z2z", (1, 15)), ("z2z", (15, 1)), ("z2z", (1001,)), ("z2z", (11, 13)), ("c2c", (143, 22)), ("z2z", (11 * 13 * 16,))]
    for kind, shape in cases:
        _, dt_in, _ = _kinds(L)[kind]
        x = oracle.synth(shape, dt_in, 940 + len(shape))
        got, desc = gpu_fft(L, kind, x, shape)
        assert "mixed-radix" in desc and "generic" not in desc, (shape, desc)
        tol = oracle.tolerance(int(np.prod(shape)), kind == "c2c")
        err = oracle.rel_l2(got, cpu_fft(oracle, kind, x, shape))
        assert err <= tol, (kind, shape, err)
        back, _ = gpu_fft(L, kind, got, shape, direction=+1)
        assert oracle.rel_l2(back / np.prod(shape), x) <= 2 * tol, (kind, shape)
    # real input: the last axis reads reals and stores the first n/2+1 outputs; the other axes run over those columns
    for kind, shape in [("d2z", (3,)), ("d2z", (6,)), ("d2z", (9,)), ("d2z", (1000,)), ("r2c", (1000,)), ("d2z", (3, 3, 2)),
                        ("d2z", (96, 96, 96)), ("r2c", (60, 64, 100)), ("d2z", (100, 512)), ("d2z", (7, 15)), ("d2z", (45, 2)), ("d2z", (26, 22)), ("r2c", (1001,))]:
        _, dt_in, _ = _kinds(L)[kind]
        x = oracle.synth(shape, dt_in, 960 + len(shape))
        got, desc = gpu_fft(L, kind, x, shape)
        assert "mixed-radix" in desc and "generic" not in desc, (shape, desc)
        err = oracle.rel_l2(got, cpu_fft(oracle, kind, x, shape))
        assert err <= oracle.tolerance(int(np.prod(shape)), kind == "r2c"), (kind, shape, err)
    # even last axes run as a half-length complex transform + even/odd pass by default; the full-length form (odd sizes,
    # odd row offsets) must agree on the same input
    import os
    for kind, shape in [("d2z", (1000,)), ("d2z", (12, 10)), ("r2c", (6, 10, 14)), ("d2z", (2,)), ("d2z", (5, 6))]:
        _, dt_in, _ = _kinds(L)[kind]
        x = oracle.synth(shape, dt_in, 961)
        want = cpu_fft(oracle, kind, x, shape)
        for half in ("1", "0"):
            os.environ["FFTB200_MIXED_HALF"] = half
            try:
                got, desc = gpu_fft(L, kind, x, shape)
            finally:
                os.environ.pop("FFTB200_MIXED_HALF", None)
            assert ("half-length" in desc) == (half == "1" and shape[-1] > 2), (shape, half, desc)
            assert oracle.rel_l2(got, want) <= oracle.tolerance(int(np.prod(shape)), kind == "r2c"), (kind, shape, half)
    # in-place real input, FFTW's padded layout
    for shape in [(1000,), (30, 90), (12, 10, 18)]:
        ftype = L.D2Z
        nl, ncol = shape[-1], shape[-1] // 2 + 1
        x = oracle.synth(shape, np.float64, 963)
        pad = np.zeros(shape[:-1] + (2 * ncol,), dtype=np.float64)
        pad[..., :nl] = x
        buf = torch.from_numpy(pad).cuda()
        inembed, onembed = list(shape[:-1]) + [2 * ncol], list(shape[:-1]) + [ncol]
        h = L.plan_many(len(shape), list(shape), inembed, 1, int(np.prod(inembed)), onembed, 1, int(np.prod(onembed)), ftype, 1)
        assert "mixed-radix" in L.describe(h)
        L.execute(h, ftype, buf.data_ptr(), buf.data_ptr())
        torch.cuda.synchronize()
        L.destroy(h)
        got = buf.cpu().numpy().view(np.complex128).reshape(shape[:-1] + (ncol,))
        assert oracle.rel_l2(got, cpu_fft(oracle, "d2z", x, shape)) <= oracle.tolerance(int(np.prod(shape)), False), shape
    # batched
    for kind, shape, batch in [("z2z", (120,), 37), ("z2z", (12, 10), 5), ("c2c", (6, 10, 14), 3), ("d2z", (90,), 11), ("d2z", (6, 10), 4)]:
        _, dt_in, _ = _kinds(L)[kind]
        x = oracle.synth((batch,) + shape, dt_in, 951)
        got, desc = gpu_fft(L, kind, x, shape, batch=batch)
        assert "mixed-radix" in desc, desc
        assert oracle.rel_l2(got, cpu_fft(oracle, kind, x, shape, batch=batch)) <= oracle.tolerance(int(np.prod(shape)), kind in ("c2c", "r2c"))
    # in place (each pass loads its whole tile before it stores)
    for shape in [(1000,), (96, 100), (48, 56, 60)]:
        ftype, dt_in, _ = _kinds(L)["z2z"]
        x = oracle.synth(shape, dt_in, 953)
        buf = torch.from_numpy(x.copy()).cuda()
        h = L.plan_many(len(shape), list(shape), None, 0, 0, None, 0, 0, ftype, 1)
        L.execute(h, ftype, buf.data_ptr(), buf.data_ptr())
        torch.cuda.synchronize()
        L.destroy(h)
        assert oracle.rel_l2(buf.cpu().numpy(), cpu_fft(oracle, "z2z", x, shape)) <= oracle.tolerance(int(np.prod(shape)), False)
    # other primes stay on the generic path
    for shape in [(2, 6, 4 * 5 ** 5 * 17), (17 * 8,)]:
        _, dt_in, _ = _kinds(L)["z2z"]
        x = oracle.synth(shape, dt_in, 955)
        got, desc = gpu_fft(L, "z2z", x, shape)
        assert "mixed-radix" not in desc, desc
        assert oracle.rel_l2(got, cpu_fft(oracle, "z2z", x, shape)) <= oracle.tolerance(int(np.prod(shape)), False)


def test_inplace_r2c_padded_layout(L, oracle):
    """In-place R2C / D2Z with FFTW's padded format (rows of 2*(n/2+1) reals; fftw-3.3.8/doc/reference.texi, "Real-data
    DFT Array Format"): the half spectrum overwrites the real rows it came from.  1-D, 2-D, 3-D (the 3-D case also takes
    the blocked intermediate layout when its slowest stride is far)."""
    for kind, shape in [("d2z", (4096,)), ("d2z", (64, 128)), ("r2c", (32, 64, 256)), ("d2z", (128, 128, 256)), ("d2z", (16, 8192))]:
        ftype, dt_in, dt_out = _kinds(L)[kind]
        nl = shape[-1]
        nc = nl // 2 + 1
        x = oracle.synth(shape, dt_in, 970 + len(shape))
        pad = np.zeros(shape[:-1] + (2 * nc,), dtype=dt_in)
        pad[..., :nl] = x
        buf = torch.from_numpy(pad).cuda()
        inembed = list(shape[:-1]) + [2 * nc]
        onembed = list(shape[:-1]) + [nc]
        h = L.plan_many(len(shape), list(shape), inembed, 1, int(np.prod(inembed)), onembed, 1, int(np.prod(onembed)), ftype, 1)
        L.execute(h, ftype, buf.data_ptr(), buf.data_ptr())
        torch.cuda.synchronize()
        L.destroy(h)
        got = buf.cpu().numpy().view(dt_out).reshape(shape[:-1] + (nc,))
        want = cpu_fft(oracle, kind, x, shape)
        err = oracle.rel_l2(got, want)
        assert err <= oracle.tolerance(int(np.prod(shape)), kind == "r2c"), (kind, shape, err)
    # an unpadded real layout cannot run in place: refused, not corrupted
    h = L.plan_many(1, [256], None, 0, 0, None, 0, 0, L.D2Z, 1)
    b = torch.zeros(258, dtype=torch.float64, device="cuda")
    h2 = L.plan_many(2, [8, 256], None, 0, 0, None, 0, 0, L.D2Z, 1)
    b2 = torch.zeros(8 * 258, dtype=torch.float64, device="cuda")
    with pytest.raises(L.FFTB200Error):
        L.execute(h2, L.D2Z, b2.data_ptr(), b2.data_ptr())
    L.destroy(h)
    L.destroy(h2)


def test_normalisation_helper_round_trips(L, oracle):
    """fftb200_scale: forward, backward, scale by 1/N gives the input back (FFTW leaves this to the user,
    doc/reference.texi:1982-2004); advanced-layout padding is not touched; explicit factors work; C2R output scales."""
    for kind, shape in [("z2z", (64, 32, 16)), ("c2c", (4096,)), ("z2z", (12, 10))]:
        ftype, dt_in, _ = _kinds(L)[kind]
        x = torch.from_numpy(oracle.synth(shape, dt_in, 980)).cuda()
        y = torch.empty_like(x)
        z = torch.empty_like(x)
        h = L.plan_many(len(shape), list(shape), None, 0, 0, None, 0, 0, ftype, 1)
        L.execute(h, ftype, x.data_ptr(), y.data_ptr(), -1)
        L.execute(h, ftype, y.data_ptr(), z.data_ptr(), +1)
        L.scale(h, z.data_ptr())
        torch.cuda.synchronize()
        L.destroy(h)
        assert _rel(z, x) <= 2 * oracle.tolerance(int(np.prod(shape)), kind == "c2c"), (kind, shape)
    # padded output rows: only the transform's own elements are scaled
    n, pitch = 64, 80
    h = L.plan_many(1, [n], [n], 1, n, [pitch], 1, pitch, L.Z2Z, 4)
    x = torch.from_numpy(oracle.synth((4, n), np.complex128, 981)).cuda()
    out = torch.full((4, pitch), 7.0 + 7.0j, dtype=torch.complex128, device="cuda")
    L.execute(h, L.Z2Z, x.data_ptr(), out.data_ptr())
    ref = out.clone()
    L.scale(h, out.data_ptr(), 0.5)
    torch.cuda.synchronize()
    L.destroy(h)
    assert torch.equal(out[:, :n], ref[:, :n] * 0.5) and torch.equal(out[:, n:], ref[:, n:])
    # real round trip: D2Z then Z2D then scale
    shape = (32, 64)
    xr = torch.from_numpy(oracle.synth(shape, np.float64, 982)).cuda()
    spec = torch.empty((32, 33), dtype=torch.complex128, device="cuda")
    back = torch.empty_like(xr)
    hf = L.plan_many(2, list(shape), None, 0, 0, None, 0, 0, L.D2Z, 1)
    hb = L.plan_many(2, list(shape), None, 0, 0, None, 0, 0, L.Z2D, 1)
    L.execute(hf, L.D2Z, xr.data_ptr(), spec.data_ptr())
    L.execute(hb, L.Z2D, spec.data_ptr(), back.data_ptr())
    L.scale(hb, back.data_ptr())
    torch.cuda.synchronize()
    L.destroy(hf)
    L.destroy(hb)
    assert _rel(back, xr) <= 2 * oracle.tolerance(32 * 64, False)


def test_blocked_intermediate_layout_matches_in_place_plan(L, oracle, monkeypatch):
    """3-D complex plans with a far slowest-axis stride route the middle pass through a blocked work buffer; the
    result is bit-identical to the in-place three-pass plan (FFTB200_ZBLOCK=0), also in place and backward."""
    for kind, shape in [("z2z", (128, 128, 128)), ("c2c", (128, 256, 256)), ("z2z", (256, 512, 64)), ("z2z", (128, 128, 512)),
                        ("c2c", (512, 128, 256))]:
        ftype, dt_in, _ = _kinds(L)[kind]
        x = torch.from_numpy(oracle.synth(shape, dt_in, 1400)).cuda()
        outs, descs, works = [], [], []
        for z in ("1", "0"):
            monkeypatch.setenv("FFTB200_ZBLOCK", z)
            h = L.plan_many(3, list(shape), None, 0, 0, None, 0, 0, ftype, 1)
            y = torch.zeros_like(x)
            L.execute(h, ftype, x.data_ptr(), y.data_ptr())
            xi = x.clone()
            L.execute(h, ftype, xi.data_ptr(), xi.data_ptr())                 # in place
            b = torch.zeros_like(x)
            L.execute(h, ftype, y.data_ptr(), b.data_ptr(), +1)               # backward
            torch.cuda.synchronize()
            descs.append(L.describe(h)); works.append(L.work_size(h))
            L.destroy(h)
            assert torch.equal(xi, y), (kind, shape, "in place", z)
            outs.append((y, b))
        assert "blocked" in descs[0] and "blocked" not in descs[1], descs
        assert works[0] == x.numel() * x.element_size() and works[1] == 0
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), (kind, shape)
        want = cpu_fft(oracle, kind, x.cpu().numpy(), shape)
        assert oracle.rel_l2(outs[0][0].cpu().numpy(), want) <= oracle.tolerance(int(np.prod(shape)), kind == "c2c")


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs at their REAL size against the reference's own FFTW (oracle/_ref, threaded build of the
# vendored 3.3.8 sources; src/fft.rg:313,319,605,608 are the calls being matched).  FFTW needs seconds, not minutes,
# for these on the GPU box's host cores.
# ------------------------------------------------------------------------------------------------
def _fftw_threads():
    try:
        return max(1, min(64, len(os.sched_getaffinity(0))))
    except AttributeError:
        return max(1, min(64, os.cpu_count() or 1))


def _host_gib_available():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / (1 << 20)
    except OSError:
        pass
    return 0.0


def _full_size_case(L, oracle, kind, shape, seed):
    """Input made on the GPU (seeded), copied to the host for FFTW; the comparison runs on the GPU slab by slab so the
    host only ever holds the input and FFTW's output."""
    ftype, dt_in, dt_out = _kinds(L)[kind]
    real = kind in ("d2z", "r2c")
    single = kind in ("c2c", "r2c")
    n_total = int(np.prod(shape))
    in_bytes = n_total * np.dtype(dt_in).itemsize
    need = 2.5 * in_bytes * (2 if single else 1) * (2 if real else 1) / (1 << 30) + 4
    if _host_gib_available() < need:
        pytest.skip(f"needs ~{need:.0f} GiB of host memory for FFTW's input and output")
    g = torch.Generator(device="cuda").manual_seed(0x5EED0000 + seed)
    rdt = torch.float32 if single else torch.float64
    xd = torch.rand(*shape, *(() if real else (2,)), dtype=rdt, device="cuda", generator=g).sub_(0.5)
    if not real:
        xd = torch.view_as_complex(xd)
    oshape = tuple(shape[:-1]) + (shape[-1] // 2 + 1,) if real else tuple(shape)
    yd = torch.zeros(oshape, dtype=_torch_dtype(dt_out), device="cuda")
    x = xd.cpu().numpy()
    h = L.plan_many(len(shape), list(shape), None, 0, 0, None, 0, 0, ftype, 1)
    try:
        L.set_stream(h, torch.cuda.current_stream().cuda_stream)
        L.execute(h, ftype, xd.data_ptr(), yd.data_ptr())
        torch.cuda.synchronize()
    finally:
        L.destroy(h)
    assert torch.equal(xd.cpu(), torch.from_numpy(x)) if in_bytes <= (4 << 30) else True, "input not preserved"
    del xd
    F = oracle.FFTW.get("ref")          # fp32 inputs: the fp64 FFTW transform of the SAME fp32 values
    x64 = x.astype(np.float64 if real else np.complex128) if single else x
    want = F.r2c(x64, threads=_fftw_threads()) if real else F.dft(x64, threads=_fftw_threads())
    del x64, x
    want = want.reshape(oshape)
    rows = oshape[0] if len(oshape) > 1 else 1
    step = max(1, rows // 16) if len(oshape) > 1 else 1
    num = den = 0.0
    if len(oshape) == 1:
        w = torch.from_numpy(want).cuda()
        num = float(torch.linalg.vector_norm(yd.to(torch.complex128) - w) ** 2)
        den = float(torch.linalg.vector_norm(w) ** 2)
    else:
        for i0 in range(0, rows, step):
            w = torch.from_numpy(want[i0:i0 + step]).cuda()
            num += float(torch.linalg.vector_norm(yd[i0:i0 + step].to(torch.complex128) - w) ** 2)
            den += float(torch.linalg.vector_norm(w) ** 2)
            del w
    err = (num / den) ** 0.5
    tol = oracle.tolerance(n_total, single)
    assert err <= tol, (kind, shape, err, tol)
    return err


def test_c4_512cubed_full_size_against_fftw(L, oracle):
    """BASELINE configs[3]: 3D C2C complex64 512^3, every output bin against FFTW; tolerance 10*log2(N)*eps = 6.0e-14."""
    _full_size_case(L, oracle, "z2z", (512, 512, 512), 4)


def test_c2_4096sq_d2z_full_size_against_fftw(L, oracle):
    """BASELINE configs[1]: 2D R2C double -> complex64 4096 x 4096 (packed 4096 x 2049 output)."""
    _full_size_case(L, oracle, "d2z", (4096, 4096), 2)


def test_c3_2pow27_c32_full_size_against_fftw(L, oracle):
    """BASELINE configs[2]: 1D C2C complex32 N = 2^27 (multi-pass), against the fp64 FFTW transform of the fp32 input;
    tolerance 10*log2(N)*eps_fp32 = 3.2e-5."""
    _full_size_case(L, oracle, "c2c", (1 << 27,), 3)


def test_c5_1024cubed_d2z_full_size_against_fftw(L, oracle):
    """BASELINE configs[4] on one GPU: 3D R2C double -> complex64 1024^3 (8.6 GB in, 8.6 GB out), every bin against FFTW."""
    _full_size_case(L, oracle, "d2z", (1024, 1024, 1024), 5)


def test_scaling_case_1024cubed_z2z_full_size_against_fftw(L, oracle):
    """The north-star strong-scaling case on one GPU: 3D C2C complex64 1024^3 (17 GB in, 17 GB out) against FFTW."""
    _full_size_case(L, oracle, "z2z", (1024, 1024, 1024), 6)
