"""A/B of two library builds on the same GPU (tools only): per-pass times of z2z 512^3 through ctypes directly."""
import ctypes, sys, torch
libs = sys.argv[1:]
x = torch.zeros((512, 512, 512), dtype=torch.complex128, device="cuda"); torch.view_as_real(x).uniform_(-0.5, 0.5)
y = torch.empty_like(x)
for rep in range(4):
    for path in libs:
        L = ctypes.CDLL(path)
        h = ctypes.c_ulonglong(0)
        n = (ctypes.c_int * 3)(512, 512, 512)
        L.fftb200_plan_many.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        assert L.fftb200_plan_many(ctypes.byref(h), 3, n, None, 0, 0, None, 0, 0, 0x69, 1) == 0
        L.fftb200_exec_z2z.argtypes = [ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.fftb200_set_profiling.argtypes = [ctypes.c_ulonglong, ctypes.c_int]
        L.fftb200_get_launch_ms.argtypes = [ctypes.c_ulonglong, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
        L.fftb200_destroy.argtypes = [ctypes.c_ulonglong]
        for _ in range(5): L.fftb200_exec_z2z(h, x.data_ptr(), y.data_ptr(), -1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): L.fftb200_exec_z2z(h, x.data_ptr(), y.data_ptr(), -1)
        e1.record(); torch.cuda.synchronize()
        tot = e0.elapsed_time(e1) / 20
        per = []
        for i in range(3):
            ts = []
            for _ in range(5):
                L.fftb200_set_profiling(h, 1)
                L.fftb200_exec_z2z(h, x.data_ptr(), y.data_ptr(), -1); torch.cuda.synchronize()
                ms = ctypes.c_float(0); L.fftb200_get_launch_ms(h, i, ctypes.byref(ms)); ts.append(ms.value)
            per.append(round(min(ts), 4))
        print(path.split("/")[-1], round(tot, 4), per, flush=True)
        L.fftb200_destroy(h)
