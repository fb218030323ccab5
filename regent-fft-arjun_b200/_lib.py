"""ctypes binding of libfft_b200.so (include/fft_b200.h).

This is the Python twin of what `terralib.includec("fft_b200.h")` +
`terralib.linklibrary("libfft_b200.so")` give Regent (reference src/fft.rg:15-20 does the same
for cufftXt.h / libcufft.so).  There is NO fallback: if the shared library is missing or does
not export a declared symbol, import fails loudly.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FFTB200_LIB_PATH: tuning experiments only (A/B of differently compiled builds); the default is the in-tree build
LIB_PATH = os.environ.get("FFTB200_LIB_PATH") or os.path.join(_HERE, "libfft_b200.so")

# enums of include/fft_b200.h
R2C, C2C, D2Z, Z2Z = 0x2A, 0x29, 0x6A, 0x69
C2R, Z2D = 0x2C, 0x6C
SUCCESS, INVALID_PLAN, ALLOC_FAILED, INVALID_TYPE, INVALID_VALUE = 0, 1, 2, 3, 4
INTERNAL_ERROR, EXEC_FAILED, SETUP_FAILED, INVALID_SIZE, UNSUPPORTED = 5, 6, 7, 8, 16
FORWARD, INVERSE = -1, 1

_handle = ctypes.c_ulonglong
_ip = ctypes.POINTER(ctypes.c_int)
_vp = ctypes.c_void_p

# every symbol the header declares: name -> (restype, argtypes)
SYMBOLS = {
    "fftb200_plan_many": (ctypes.c_int, [ctypes.POINTER(_handle), ctypes.c_int, _ip, _ip, ctypes.c_int, ctypes.c_int,
                                         _ip, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "fftb200_set_stream": (ctypes.c_int, [_handle, _vp]),
    "fftb200_exec_c2c": (ctypes.c_int, [_handle, _vp, _vp, ctypes.c_int]),
    "fftb200_exec_z2z": (ctypes.c_int, [_handle, _vp, _vp, ctypes.c_int]),
    "fftb200_exec_r2c": (ctypes.c_int, [_handle, _vp, _vp]),
    "fftb200_exec_d2z": (ctypes.c_int, [_handle, _vp, _vp]),
    "fftb200_exec_c2r": (ctypes.c_int, [_handle, _vp, _vp]),
    "fftb200_exec_z2d": (ctypes.c_int, [_handle, _vp, _vp]),
    "fftb200_scale": (ctypes.c_int, [_handle, _vp, ctypes.c_double]),
    "fftb200_destroy": (ctypes.c_int, [_handle]),
    "fftb200_get_work_size": (ctypes.c_int, [_handle, ctypes.POINTER(ctypes.c_ulonglong)]),
    "fftb200_get_launch_count": (ctypes.c_int, [_handle, _ip]),
    "fftb200_describe": (ctypes.c_int, [_handle, ctypes.c_char_p, ctypes.c_int]),
    "fftb200_get_launch_bytes": (ctypes.c_int, [_handle, ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong)]),
    "fftb200_set_profiling": (ctypes.c_int, [_handle, ctypes.c_int]),
    "fftb200_get_launch_ms": (ctypes.c_int, [_handle, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]),
    "fftb200_slab_plan": (ctypes.c_int, [ctypes.POINTER(_handle), _ip, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "fftb200_slab_plan_2d": (ctypes.c_int, [ctypes.POINTER(_handle), _ip, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "fftb200_slab_get_ipc_handle": (ctypes.c_int, [_handle, _vp]),
    "fftb200_slab_connect_ipc": (ctypes.c_int, [_handle, _vp]),
    "fftb200_slab_get_area": (ctypes.c_int, [_handle, ctypes.POINTER(_vp), ctypes.POINTER(ctypes.c_ulonglong)]),
    "fftb200_slab_connect_ptrs": (ctypes.c_int, [_handle, ctypes.POINTER(_vp)]),
    "fftb200_slab_exec": (ctypes.c_int, [_handle, _vp, _vp, ctypes.c_int]),
    "fftb200_slab_exec_pre": (ctypes.c_int, [_handle, _vp, _vp, ctypes.c_int]),
    "fftb200_slab_exec_post": (ctypes.c_int, [_handle, _vp, _vp, ctypes.c_int]),
    "fftb200_slab_set_timing": (ctypes.c_int, [_handle, ctypes.c_int]),
    "fftb200_slab_get_phase_ms": (ctypes.c_int, [_handle, ctypes.POINTER(ctypes.c_float)]),
    "fftb200_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "fftb200_version": (ctypes.c_int, []),
}


class FFTB200Error(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        msg = lib().fftb200_strerror(code).decode() if _lib is not None else "?"
        super().__init__(f"{where} failed: {msg} (code {code})")


_lib = None


def lib() -> ctypes.CDLL:
    """Load libfft_b200.so once.  Raises if it is absent: the product has no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `make -C regent-fft-arjun_b200` "
                "(or __graft_entry__.build()).  There is no CPU or library fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code: int, where: str) -> None:
    if code != SUCCESS:
        raise FFTB200Error(code, where)


def _ints(v):
    return None if v is None else (ctypes.c_int * len(v))(*[int(x) for x in v])


def plan_many(rank, n, inembed, istride, idist, onembed, ostride, odist, ftype, batch) -> int:
    h = _handle(0)
    rc = lib().fftb200_plan_many(ctypes.byref(h), rank, _ints(n), _ints(inembed), istride, idist,
                                 _ints(onembed), ostride, odist, ftype, batch)
    check(rc, "fftb200_plan_many")
    return int(h.value)


def set_stream(h: int, stream_ptr: int) -> None:
    check(lib().fftb200_set_stream(h, stream_ptr), "fftb200_set_stream")


def execute(h: int, ftype: int, in_ptr: int, out_ptr: int, direction: int = FORWARD) -> None:
    L = lib()
    if ftype == Z2Z:
        rc = L.fftb200_exec_z2z(h, in_ptr, out_ptr, direction)
    elif ftype == C2C:
        rc = L.fftb200_exec_c2c(h, in_ptr, out_ptr, direction)
    elif ftype == D2Z:
        rc = L.fftb200_exec_d2z(h, in_ptr, out_ptr)
    elif ftype == R2C:
        rc = L.fftb200_exec_r2c(h, in_ptr, out_ptr)
    elif ftype == C2R:
        rc = L.fftb200_exec_c2r(h, in_ptr, out_ptr)
    elif ftype == Z2D:
        rc = L.fftb200_exec_z2d(h, in_ptr, out_ptr)
    else:
        raise ValueError("bad transform type")
    check(rc, "fftb200_exec")


def scale(h: int, data_ptr: int, factor: float = 0.0) -> None:
    """normalisation helper: scale the plan's output array in place (factor 0 = 1 / n_total)"""
    check(lib().fftb200_scale(h, data_ptr, factor), "fftb200_scale")


def destroy(h: int) -> None:
    check(lib().fftb200_destroy(h), "fftb200_destroy")


def describe(h: int) -> str:
    buf = ctypes.create_string_buffer(8192)
    check(lib().fftb200_describe(h, buf, len(buf)), "fftb200_describe")
    return buf.value.decode()


def launch_count(h: int) -> int:
    n = ctypes.c_int(0)
    check(lib().fftb200_get_launch_count(h, ctypes.byref(n)), "fftb200_get_launch_count")
    return n.value


def launch_bytes(h: int, i: int) -> int:
    b = ctypes.c_ulonglong(0)
    check(lib().fftb200_get_launch_bytes(h, i, ctypes.byref(b)), "fftb200_get_launch_bytes")
    return int(b.value)


def work_size(h: int) -> int:
    b = ctypes.c_ulonglong(0)
    check(lib().fftb200_get_work_size(h, ctypes.byref(b)), "fftb200_get_work_size")
    return int(b.value)


def set_profiling(h: int, on: bool) -> None:
    check(lib().fftb200_set_profiling(h, 1 if on else 0), "fftb200_set_profiling")


def launch_ms(h: int, i: int) -> float:
    ms = ctypes.c_float(0)
    check(lib().fftb200_get_launch_ms(h, i, ctypes.byref(ms)), "fftb200_get_launch_ms")
    return float(ms.value)


# ---- multi-GPU slab transforms ---------------------------------------------------------------
def slab_plan(n, ftype, rank, nranks, chunks=1) -> int:
    h = _handle(0)
    check(lib().fftb200_slab_plan(ctypes.byref(h), _ints(n), ftype, rank, nranks, chunks), "fftb200_slab_plan")
    return int(h.value)


def slab_plan_2d(n, ftype, rank, nranks) -> int:
    h = _handle(0)
    check(lib().fftb200_slab_plan_2d(ctypes.byref(h), _ints(n), ftype, rank, nranks), "fftb200_slab_plan_2d")
    return int(h.value)


def slab_ipc_handle(h: int) -> bytes:
    buf = ctypes.create_string_buffer(64)
    check(lib().fftb200_slab_get_ipc_handle(h, buf), "fftb200_slab_get_ipc_handle")
    return buf.raw


def slab_connect_ipc(h: int, handles: bytes) -> None:
    buf = ctypes.create_string_buffer(handles, len(handles))
    check(lib().fftb200_slab_connect_ipc(h, buf), "fftb200_slab_connect_ipc")


def slab_area(h: int):
    p, b = _vp(0), ctypes.c_ulonglong(0)
    check(lib().fftb200_slab_get_area(h, ctypes.byref(p), ctypes.byref(b)), "fftb200_slab_get_area")
    return int(p.value or 0), int(b.value)


def slab_connect_ptrs(h: int, areas) -> None:
    arr = (_vp * len(areas))(*[_vp(a) for a in areas])
    check(lib().fftb200_slab_connect_ptrs(h, arr), "fftb200_slab_connect_ptrs")


def slab_exec(h: int, in_ptr: int, out_ptr: int, direction: int = FORWARD) -> None:
    check(lib().fftb200_slab_exec(h, in_ptr, out_ptr, direction), "fftb200_slab_exec")


def slab_exec_pre(h: int, in_ptr: int, send_ptr: int, direction: int = FORWARD) -> None:
    check(lib().fftb200_slab_exec_pre(h, in_ptr, send_ptr, direction), "fftb200_slab_exec_pre")


def slab_exec_post(h: int, recv_ptr: int, out_ptr: int, direction: int = FORWARD) -> None:
    check(lib().fftb200_slab_exec_post(h, recv_ptr, out_ptr, direction), "fftb200_slab_exec_post")


def slab_set_timing(h: int, on: bool) -> None:
    check(lib().fftb200_slab_set_timing(h, 1 if on else 0), "fftb200_slab_set_timing")


def slab_phase_ms(h: int):
    ms = (ctypes.c_float * 3)()
    check(lib().fftb200_slab_get_phase_ms(h, ms), "fftb200_slab_get_phase_ms")
    return [float(v) for v in ms]
