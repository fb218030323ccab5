"""Development battery: parity of many shapes against the oracle + per-launch timings.
Usage: python tools/gpu_check.py [quick|full] (writes gpurun_out/gpu_check.log when run via gpurun)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from __graft_entry__ import load_package
fft = load_package()
L = fft._lib

TYPES = {"z2z": (L.Z2Z, np.complex128, np.complex128), "c2c": (L.C2C, np.complex64, np.complex64),
         "d2z": (L.D2Z, np.float64, np.complex128), "r2c": (L.R2C, np.float32, np.complex64)}

def run_case(kind, shape, batch=1, seed=1, check=True, reps=0):
    ftype, dt_in, dt_out = TYPES[kind]
    real = kind in ("d2z", "r2c")
    full = (batch,) + tuple(shape) if batch > 1 else tuple(shape)
    x = oracle.synth(full, dt_in, seed)
    oshape = full[:-1] + (full[-1] // 2 + 1,) if real else full
    xd = torch.from_numpy(x).cuda()
    yd = torch.zeros(oshape, dtype=torch.from_numpy(np.zeros(1, dt_out)).dtype, device="cuda")
    h = L.plan_many(len(shape), list(shape), None, 0, 0, None, 0, 0, ftype, batch)
    L.execute(h, ftype, xd.data_ptr(), yd.data_ptr())
    torch.cuda.synchronize()
    err = None
    if check:
        x64 = x.astype(np.float64 if real else np.complex128)
        axes = tuple(range(-len(shape), 0))
        want = np.fft.rfftn(x64, axes=axes) if real else np.fft.fftn(x64, axes=axes)
        err = oracle.rel_l2(yd.cpu().numpy(), want)
    ms = None
    if reps:
        for _ in range(3): L.execute(h, ftype, xd.data_ptr(), yd.data_ptr())
        torch.cuda.synchronize()
        L.set_profiling(h, True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        nl = L.launch_count(h)
        per = np.zeros(nl)
        for _ in range(reps):
            e0.record(); L.execute(h, ftype, xd.data_ptr(), yd.data_ptr()); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            per += np.array([L.launch_ms(h, i) for i in range(nl)])
        per /= reps
        ms = (min(ts), float(np.median(ts)), per, [L.launch_bytes(h, i) for i in range(nl)])
    desc = L.describe(h)
    L.destroy(h)
    return err, ms, desc

def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count, "SMs")
    bad = 0
    cases = []
    for k in ("z2z", "c2c"):
        for n in [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192] + ([16384] if k == "c2c" else []):
            cases.append((k, (n,), 3))
        cases += [(k, (3,), 1), (k, (5,), 2), (k, (12,), 1), (k, (1021,), 1), (k, (2, 2), 1), (k, (3, 2, 2), 1),
                  (k, (64, 64), 1), (k, (32, 256), 2), (k, (16, 8, 32), 1), (k, (64, 64, 64), 1), (k, (128, 4, 512), 1),
                  (k, (6, 10, 9), 2), (k, (1 << 15,), 1), (k, (1 << 18,), 2), (k, (1 << 20,), 1), (k, (1 << 22,), 1)]
    for k in ("d2z", "r2c"):
        for n in [2, 4, 8, 16, 64, 256, 1024, 4096, 8192, 16384]:
            cases.append((k, (n,), 2))
        cases += [(k, (3,), 1), (k, (9,), 1), (k, (10,), 3), (k, (64, 64), 1), (k, (16, 8, 32), 1), (k, (64, 64, 64), 1),
                  (k, (3, 3, 2), 1), (k, (5, 6, 7), 2), (k, (256, 512), 1), (k, (32, 4, 8), 1)]
    for kind, shape, batch in cases:
        try:
            err, _, desc = run_case(kind, shape, batch)
        except Exception as ex:
            print(f"FAIL {kind} {shape} b={batch}: {ex}"); bad += 1; continue
        ntot = int(np.prod(shape))
        tol = oracle.tolerance(ntot, kind in ("c2c", "r2c"))
        ok = err <= tol
        bad += (not ok)
        path = "generic" if "generic" in desc else "tile"
        print(f"{'ok  ' if ok else 'BAD '} {kind} {str(shape):>18} b={batch} err={err:.2e} tol={tol:.1e} [{path}, {desc.count(chr(10))} launches]")
    print("parity failures:", bad)
    # timings
    tcases = [("z2z", (256, 256, 256)), ("z2z", (512, 512, 512)), ("c2c", (512, 512, 512)), ("d2z", (4096, 4096)),
              ("c2c", (1 << 27,)), ("z2z", (1 << 24,)), ("z2z", (4096, 4096)), ("z2z", (1024, 1024, 256))]
    if mode == "quick": tcases = tcases[:2]
    for kind, shape in tcases:
        try:
            err, ms, desc = run_case(kind, shape, 1, check=(np.prod(shape) <= (1 << 25)), reps=10)
        except Exception as ex:
            print(f"FAIL timing {kind} {shape}: {ex}"); continue
        tmin, tmed, per, bytes_ = ms
        n = float(np.prod(shape)); flops = (2.5 if kind in ("d2z", "r2c") else 5.0) * n * np.log2(n)
        print(f"\n== {kind} {shape}: min {tmin:.3f} ms med {tmed:.3f} ms  {flops / tmin / 1e6:.1f} GFLOP/s  err={err}")
        for i, line in enumerate(desc.strip().split("\n")):
            print(f"   [{i}] {per[i]:.3f} ms  {bytes_[i] / per[i] / 1e6:.0f} GB/s  {line}")
    return bad

if __name__ == "__main__":
    sys.exit(1 if main() else 0)
