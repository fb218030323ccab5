"""oracle — TEST INFRASTRUCTURE ONLY.

CPU checkers for the Regent-FFT hot path (reference: /root/reference/src/fft.rg,
CPU branch `fftw_plan_dft*` :313,319,483,500 and `fftw_execute_dft*` :605,608).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package, and only as the checker.  The product
(regent-fft-arjun_b200/, libfft_b200.so) never imports, links or calls anything here.

Two independent checkers, both driven through ctypes:

* ``port_*``  — oracle/fft_oracle.c, our plain-C restatement (kind "port").
* ``fftw_*``  — the reference's own FFTW 3.3.8, compiled by oracle/Makefile from the
  vendored sources where they lie (oracle/_ref/libfftw3_ref.so: threads + AVX2;
  libfftw3f_ref.so: float; libfftw3_prebuilt.so: the author's scalar binary)
  (kind "reference").  Built in the dev container; on the GPU box the prebuilt
  files travel with the snapshot (/root/reference does not exist there).

Parity of the port is PINNED: tests/test_oracle.py checks it against the reference's
FFTW on seeded inputs, against the eight known answers of test/fft_test.rg
 and against tests/golden/*.npz produced by the real FFTW.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")

FFTW_FORWARD = -1   # src/fft.rg:22
FFTW_ESTIMATE = 1 << 6  # src/fft.rg:25


def build(verbose: bool = False) -> None:
    """Compile the restatement and (when /root/reference is present) the reference FFTW."""
    cmd = ["make", "-C", _HERE, "-j", str(os.cpu_count() or 4), "all"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:], res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("oracle build failed")


def _load(name: str) -> ctypes.CDLL:
    path = os.path.join(_REF, name)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `make -C oracle` (dev container) first")
    return ctypes.CDLL(path)


# --------------------------------------------------------------------------- port
_port = None


def _port_lib() -> ctypes.CDLL:
    global _port
    if _port is None:
        L = _load("libfft_oracle.so")
        ip = ctypes.POINTER(ctypes.c_int)
        L.oracle_dft_many.argtypes = [ctypes.c_int, ip, ctypes.c_int, ctypes.c_void_p, ip, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p, ip, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.oracle_dft_r2c_many.argtypes = [ctypes.c_int, ip, ctypes.c_int, ctypes.c_void_p, ip, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p, ip, ctypes.c_int, ctypes.c_int]
        L.oracle_twiddle.argtypes = [ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_void_p]
        _port = L
    return _port


def _ints(v):
    return None if v is None else (ctypes.c_int * len(v))(*[int(x) for x in v])


def port_dft_many(n, howmany, x, inembed, istride, idist, out, onembed, ostride, odist, sign=-1):
    """api/plan-many-dft.c:26-51 semantics on flat complex128 buffers (in place on `out`)."""
    assert x.dtype == np.complex128 and out.dtype == np.complex128
    rc = _port_lib().oracle_dft_many(len(n), _ints(n), howmany, x.ctypes.data, _ints(inembed), istride, idist,
                                     out.ctypes.data, _ints(onembed), ostride, odist, sign)
    if rc:
        raise RuntimeError(f"oracle_dft_many rc={rc}")
    return out


def port_r2c_many(n, howmany, x, inembed, istride, idist, out, onembed, ostride, odist):
    assert x.dtype == np.float64 and out.dtype == np.complex128
    rc = _port_lib().oracle_dft_r2c_many(len(n), _ints(n), howmany, x.ctypes.data, _ints(inembed), istride, idist,
                                         out.ctypes.data, _ints(onembed), ostride, odist)
    if rc:
        raise RuntimeError(f"oracle_dft_r2c_many rc={rc}")
    return out


def port_dft(x: np.ndarray, sign: int = -1) -> np.ndarray:
    """Forward (sign=-1) unnormalised DFT over all axes of a C-contiguous complex array."""
    x = np.ascontiguousarray(x, dtype=np.complex128)
    out = np.empty_like(x)
    return port_dft_many(x.shape, 1, x, None, 1, 0, out, None, 1, 0, sign)


def port_r2c(x: np.ndarray) -> np.ndarray:
    """r2c over all axes; output last dim n/2+1 (packed)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(x.shape[:-1] + (x.shape[-1] // 2 + 1,), dtype=np.complex128)
    return port_r2c_many(x.shape, 1, x, None, 1, 0, out, None, 1, 0)


def port_twiddle(m: int, n: int, sign: int = -1) -> complex:
    buf = (ctypes.c_double * 2)()
    _port_lib().oracle_twiddle(m, n, sign, buf)
    return complex(buf[0], buf[1])


# --------------------------------------------------------------------------- reference FFTW
class FFTW:
    """ctypes view of one FFTW build, exposing exactly the seven calls fft.rg makes (+ threads)."""

    _cache: dict = {}

    def __init__(self, which: str = "ref"):
        fname, prefix, self.cdtype, self.rdtype = {
            "ref": ("libfftw3_ref.so", "fftw_", np.complex128, np.float64),
            "prebuilt": ("libfftw3_prebuilt.so", "fftw_", np.complex128, np.float64),
            "float": ("libfftw3f_ref.so", "fftwf_", np.complex64, np.float32),
        }[which]
        self.which = which
        L = _load(fname)
        ip = ctypes.POINTER(ctypes.c_int)
        vp = ctypes.c_void_p
        f = lambda s: getattr(L, prefix + s)
        self._plan_dft = f("plan_dft"); self._plan_dft.restype = vp
        self._plan_dft.argtypes = [ctypes.c_int, ip, vp, vp, ctypes.c_int, ctypes.c_uint]
        self._plan_dft_r2c = f("plan_dft_r2c"); self._plan_dft_r2c.restype = vp
        self._plan_dft_r2c.argtypes = [ctypes.c_int, ip, vp, vp, ctypes.c_uint]
        self._plan_many_dft = f("plan_many_dft"); self._plan_many_dft.restype = vp
        self._plan_many_dft.argtypes = [ctypes.c_int, ip, ctypes.c_int, vp, ip, ctypes.c_int, ctypes.c_int,
                                        vp, ip, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint]
        self._plan_many_dft_r2c = f("plan_many_dft_r2c"); self._plan_many_dft_r2c.restype = vp
        self._plan_many_dft_r2c.argtypes = [ctypes.c_int, ip, ctypes.c_int, vp, ip, ctypes.c_int, ctypes.c_int,
                                            vp, ip, ctypes.c_int, ctypes.c_int, ctypes.c_uint]
        self._execute_dft = f("execute_dft"); self._execute_dft.argtypes = [vp, vp, vp]
        self._execute_dft_r2c = f("execute_dft_r2c"); self._execute_dft_r2c.argtypes = [vp, vp, vp]
        self._destroy_plan = f("destroy_plan"); self._destroy_plan.argtypes = [vp]
        self.has_threads = hasattr(L, prefix + "init_threads")
        if self.has_threads:
            f("init_threads").restype = ctypes.c_int
            f("init_threads")()
            self._plan_with_nthreads = f("plan_with_nthreads")
            self._plan_with_nthreads.argtypes = [ctypes.c_int]
        self.L = L

    @classmethod
    def get(cls, which: str = "ref") -> "FFTW":
        if which not in cls._cache:
            cls._cache[which] = cls(which)
        return cls._cache[which]

    def set_threads(self, nthreads: int) -> None:
        if self.has_threads:
            self._plan_with_nthreads(int(nthreads))
        elif nthreads != 1:
            raise RuntimeError(f"FFTW build {self.which!r} has no threads")

    # -- the calls of src/fft.rg:313,319 (basic) ------------------------------------------
    def dft(self, x: np.ndarray, threads: int = 1) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=self.cdtype)
        out = np.empty_like(x)
        self.set_threads(threads)
        p = self._plan_dft(x.ndim, _ints(x.shape), x.ctypes.data, out.ctypes.data, FFTW_FORWARD, FFTW_ESTIMATE)
        assert p, "fftw_plan_dft returned NULL"
        self._execute_dft(p, x.ctypes.data, out.ctypes.data)
        self._destroy_plan(p)
        return out

    def r2c(self, x: np.ndarray, threads: int = 1) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=self.rdtype)
        out = np.empty(x.shape[:-1] + (x.shape[-1] // 2 + 1,), dtype=self.cdtype)
        self.set_threads(threads)
        p = self._plan_dft_r2c(x.ndim, _ints(x.shape), x.ctypes.data, out.ctypes.data, FFTW_ESTIMATE)
        assert p, "fftw_plan_dft_r2c returned NULL"
        self._execute_dft_r2c(p, x.ctypes.data, out.ctypes.data)
        self._destroy_plan(p)
        return out

    # -- the calls of src/fft.rg:483,500 (advanced; flat buffers, `out` written in place) --
    def dft_many(self, n, howmany, x, inembed, istride, idist, out, onembed, ostride, odist, threads: int = 1):
        assert x.dtype == self.cdtype and out.dtype == self.cdtype
        self.set_threads(threads)
        p = self._plan_many_dft(len(n), _ints(n), howmany, x.ctypes.data, _ints(inembed), istride, idist,
                                out.ctypes.data, _ints(onembed), ostride, odist, FFTW_FORWARD, FFTW_ESTIMATE)
        assert p, "fftw_plan_many_dft returned NULL"
        self._execute_dft(p, x.ctypes.data, out.ctypes.data)
        self._destroy_plan(p)
        return out

    def r2c_many(self, n, howmany, x, inembed, istride, idist, out, onembed, ostride, odist, threads: int = 1):
        assert x.dtype == self.rdtype and out.dtype == self.cdtype
        self.set_threads(threads)
        p = self._plan_many_dft_r2c(len(n), _ints(n), howmany, x.ctypes.data, _ints(inembed), istride, idist,
                                    out.ctypes.data, _ints(onembed), ostride, odist, FFTW_ESTIMATE)
        assert p, "fftw_plan_many_dft_r2c returned NULL"
        self._execute_dft_r2c(p, x.ctypes.data, out.ctypes.data)
        self._destroy_plan(p)
        return out

    # -- timing (bench.py cpu_baseline / --impl reference) ---------------------------------
    def time_transform(self, x: np.ndarray, real: bool, threads: int, reps: int, warmup: int = 1):
        """Plan once (ESTIMATE, like fft.rg), then time `reps` new-array executes.  Returns (times_s, out)."""
        self.set_threads(threads)
        if real:
            x = np.ascontiguousarray(x, dtype=self.rdtype)
            out = np.empty(x.shape[:-1] + (x.shape[-1] // 2 + 1,), dtype=self.cdtype)
            p = self._plan_dft_r2c(x.ndim, _ints(x.shape), x.ctypes.data, out.ctypes.data, FFTW_ESTIMATE)
            run = lambda: self._execute_dft_r2c(p, x.ctypes.data, out.ctypes.data)
        else:
            x = np.ascontiguousarray(x, dtype=self.cdtype)
            out = np.empty_like(x)
            p = self._plan_dft(x.ndim, _ints(x.shape), x.ctypes.data, out.ctypes.data, FFTW_FORWARD, FFTW_ESTIMATE)
            run = lambda: self._execute_dft(p, x.ctypes.data, out.ctypes.data)
        assert p
        for _ in range(warmup):
            run()
        times = []
        for _ in range(reps):
            t0 = time.perf_counter()
            run()
            times.append(time.perf_counter() - t0)
        self._destroy_plan(p)
        return times, out


def have_fftw(which: str = "ref") -> bool:
    try:
        FFTW.get(which)
        return True
    except (FileNotFoundError, OSError):
        return False


# --------------------------------------------------------------------------- inputs & metric
def synth(shape, dtype, seed: int) -> np.ndarray:
    """Seeded i.i.d. uniform[-0.5,0.5) inputs (libbench2/verify-lib.c:64-67 distribution), generated
    in fp64 and rounded once to the working precision."""
    rng = np.random.default_rng(0x5EED0000 + seed)
    dt = np.dtype(dtype)
    if dt.kind == "c":
        re = rng.random(shape) - 0.5
        im = rng.random(shape) - 0.5
        return (re + 1j * im).astype(dt)
    return (rng.random(shape) - 0.5).astype(dt)


def rel_l2(got: np.ndarray, want: np.ndarray) -> float:
    got = np.asarray(got).astype(np.complex128, copy=False).ravel()
    want = np.asarray(want).astype(np.complex128, copy=False).ravel()
    den = float(np.linalg.norm(want))
    num = float(np.linalg.norm(got - want))
    return num / den if den > 0 else num


def tolerance(n_total: int, single: bool) -> float:
    """BASELINE.json north_star: relative L2 <= 10*log2(N)*eps of the precision."""
    eps = float(np.finfo(np.float32 if single else np.float64).eps)
    return 10.0 * max(1.0, float(np.log2(max(2, n_total)))) * eps
