"""Pins the oracle (CPU; no GPU needed).

1. the plain-C restatement (oracle/fft_oracle.c) against the committed golden vectors that were
   produced by the reference's own FFTW 3.3.8 (tests/golden/make_golden.py);
2. the eight known answers implied by the reference's test program (test/fft_test.rg:138-389,
   SURVEY.md §4): constant input -> DC bin = sum, everything else 0;
3. the restatement against the reference FFTW itself (oracle/_ref) on fresh seeded inputs, when
   that build is present (dev container; it also travels to the GPU box);
4. Ergun-style properties FFTW's own self-test uses (libbench2/verify-lib.c:260-414).
"""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fftw_golden.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def _names(g):
    return sorted({k[:-3] for k in g.files})


def test_port_matches_fftw_golden(oracle, golden):
    assert len(_names(golden)) >= 20
    for name in _names(golden):
        x, y = golden[name + "__x"], golden[name + "__y"]
        if name.startswith("batch_c"):
            got = np.zeros(18, np.complex128)
            oracle.port_dft_many([3, 3], 2, x, [3, 3], 1, 9, got, [3, 3], 1, 9)
        elif name.startswith("batch_r"):
            got = np.zeros(18, np.complex128)
            oracle.port_r2c_many([3, 3], 2, x, [3, 3], 1, 9, got, [3, 3], 1, 9)
        elif x.dtype.kind == "c":
            got = oracle.port_dft(x)
        else:
            got = oracle.port_r2c(x)
        err = oracle.rel_l2(got, y)
        assert err < 2e-15, (name, err)


def test_golden_matches_numpy(oracle, golden):
    """independent third opinion on the fixtures themselves"""
    for name in _names(golden):
        if name.startswith("batch"):
            continue
        x, y = golden[name + "__x"], golden[name + "__y"]
        want = np.fft.fftn(x) if x.dtype.kind == "c" else np.fft.rfftn(x)
        assert oracle.rel_l2(y, want) < 2e-15, name


# (test task, extent, input value, expected flat output buffer) — SURVEY.md §4 table
def _kat_cases():
    z = 0
    return [
        ("test1d", (5,), 3 + 3j, [15 + 15j, z, z, z, z]),
        ("test1d_distrib_shard", (3,), 4 + 4j, [12 + 12j, z, z]),
        ("test2d", (2, 2), 5 + 5j, [20 + 20j, z, z, z]),
        ("test3d", (3, 2, 2), 3 + 3j, [36 + 36j] + [z] * 11),
    ]


def test_known_answers_complex(oracle):
    for name, extent, val, want in _kat_cases():
        x = np.full(extent, val, dtype=np.complex128)
        got = oracle.port_dft(x).ravel()
        assert np.allclose(got, np.array(want, dtype=np.complex128), atol=1e-13), name


def test_known_answers_real_and_batch(oracle):
    # test1d_real: int1d 3, 3.0 -> [9, 0] (only n/2+1 = 2 entries written)
    got = oracle.port_r2c(np.full((3,), 3.0))
    assert np.allclose(got, [9, 0], atol=1e-13)
    # test3d_batch: {3,3,2}, 3+3i: 2 batches of 3x3 at distance 9 -> 27+27i at flat 0 and 9
    x = np.full(18, 3 + 3j, dtype=np.complex128)
    out = np.zeros(18, np.complex128)
    oracle.port_dft_many([3, 3], 2, x, [3, 3], 1, 9, out, [3, 3], 1, 9)
    want = np.zeros(18, np.complex128)
    want[0] = want[9] = 27 + 27j
    assert np.allclose(out, want, atol=1e-13)
    # test3d_batch_real: 27 at flat 0 and 9; row pitch stays 3, entries 2,5,8,11,14,17 untouched
    xr = np.full(18, 3.0)
    out = np.full(18, -7 - 7j, dtype=np.complex128)
    oracle.port_r2c_many([3, 3], 2, xr, [3, 3], 1, 9, out, [3, 3], 1, 9)
    for i in (2, 5, 8, 11, 14, 17):
        assert out[i] == -7 - 7j
    assert abs(out[0] - 27) < 1e-13 and abs(out[9] - 27) < 1e-13
    mask = np.ones(18, bool)
    mask[[0, 9, 2, 5, 8, 11, 14, 17]] = False
    assert np.allclose(out[mask], 0, atol=1e-13)


def test_twiddle_matches_definition(oracle):
    for n in (8, 12, 1024, 1000003):
        for m in (0, 1, n // 3, n - 1):
            w = oracle.port_twiddle(m, n, -1)
            mm = m if 2 * m <= n else m - n      # reduce the argument so numpy's exp is itself accurate
            assert abs(w - np.exp(-2j * np.pi * mm / n)) < 3e-16


@pytest.mark.parametrize("which", ["ref", "prebuilt"])
def test_port_against_reference_fftw(oracle, which):
    if not oracle.have_fftw(which):
        pytest.skip(f"oracle/_ref FFTW build {which!r} not present")
    F = oracle.FFTW.get(which)
    for i, shape in enumerate([(7,), (1024,), (30, 7), (16, 16, 16), (9, 5, 4), (2048,), (257,)]):
        x = oracle.synth(shape, np.complex128, 100 + i)
        assert oracle.rel_l2(oracle.port_dft(x), F.dft(x)) < 2e-15, shape
        xr = oracle.synth(shape, np.float64, 200 + i)
        assert oracle.rel_l2(oracle.port_r2c(xr), F.r2c(xr)) < 2e-15, shape


def test_float_fftw_against_double(oracle):
    if not oracle.have_fftw("float"):
        pytest.skip("float FFTW build not present")
    F = oracle.FFTW.get("float")
    x = oracle.synth((4096,), np.complex64, 7)
    err = oracle.rel_l2(F.dft(x), oracle.port_dft(x.astype(np.complex128)))
    assert err < oracle.tolerance(4096, single=True)


def test_threaded_fftw_matches_single(oracle):
    if not oracle.have_fftw("ref"):
        pytest.skip("reference FFTW build not present")
    F = oracle.FFTW.get("ref")
    x = oracle.synth((64, 64, 64), np.complex128, 9)
    assert oracle.rel_l2(F.dft(x, threads=4), F.dft(x, threads=1)) < 1e-15


def test_ergun_properties(oracle):
    """impulse, linearity, time shift (libbench2/verify-lib.c:260-414)."""
    n = 96
    e = np.zeros(n, np.complex128)
    e[0] = 1
    assert np.allclose(oracle.port_dft(e), 1)
    a = oracle.synth((n,), np.complex128, 1)
    b = oracle.synth((n,), np.complex128, 2)
    lhs = oracle.port_dft(2.5 * a - 1.5j * b)
    rhs = 2.5 * oracle.port_dft(a) - 1.5j * oracle.port_dft(b)
    assert oracle.rel_l2(lhs, rhs) < 1e-15
    shifted = oracle.port_dft(np.roll(a, 1))
    assert oracle.rel_l2(shifted, oracle.port_dft(a) * np.exp(-2j * np.pi * np.arange(n) / n)) < 2e-15


def test_tolerance_formula(oracle):
    assert abs(oracle.tolerance(512 ** 3, False) - 10 * 27 * 2.220446049250313e-16) < 1e-20
    assert abs(oracle.tolerance(2 ** 27, True) - 10 * 27 * 1.1920929e-07) < 1e-9
