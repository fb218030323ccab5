"""Non-power-of-two complex transforms: the mixed-radix shared-memory path against the generic multi-pass path
(FFTB200_MIXED=0) and cuFFT (torch.fft) on the same arrays; tools only.  One JSON line per shape."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
fft = load_package(); L = fft._lib


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def ours(x, y, shape, batch, ftype, mixed, env=None):
    os.environ["FFTB200_MIXED"] = "1" if mixed else "0"
    for k, v in (env or {}).items():
        os.environ[k] = str(v)
    h = L.plan_many(len(shape), list(shape), None, 0, 0, None, 0, 0, ftype, batch)
    L.set_stream(h, torch.cuda.current_stream().cuda_stream)
    ms = timed(lambda: L.execute(h, ftype, x.data_ptr(), y.data_ptr()))
    n = L.launch_count(h)
    desc = L.describe(h).strip().split("\n")
    L.destroy(h)
    os.environ.pop("FFTB200_MIXED", None)
    for k in (env or {}):
        os.environ.pop(k, None)
    return ms, n, desc


for prec, shape, batch in [("z2z", (96, 96, 96), 1), ("z2z", (192, 192, 192), 1), ("z2z", (384, 384, 384), 1), ("z2z", (100, 100, 100), 1),
                           ("z2z", (360, 360), 64), ("z2z", (1000,), 16384), ("z2z", (1536, 1536), 4), ("z2z", (6000,), 2048),
                           ("c2c", (384, 384, 384), 1), ("c2c", (1000,), 32768), ("z2z", (720, 1280), 8),
                           ("z2z", (1000000,), 8), ("z2z", (100000,), 64), ("c2c", (3000000,), 4), ("d2z", (1000000,), 16), ("z2z", (16384, 16384), 1), ("z2z", (10000, 10000), 1),
                           ("d2z", (384, 384, 384), 1), ("d2z", (1000,), 32768), ("r2c", (360, 360), 128), ("d2z", (1080, 1920), 4)]:
    real = prec in ("d2z", "r2c")
    dt = {"z2z": torch.complex128, "c2c": torch.complex64, "d2z": torch.float64, "r2c": torch.float32}[prec]
    ftype = {"z2z": L.Z2Z, "c2c": L.C2C, "d2z": L.D2Z, "r2c": L.R2C}[prec]
    full = ((batch,) if batch > 1 else ()) + shape
    x = torch.zeros(full, dtype=dt, device="cuda")
    (torch.view_as_real(x) if not real else x).uniform_(-0.5, 0.5)
    y = torch.empty_like(x) if not real else torch.empty(full[:-1] + (full[-1] // 2 + 1,), dtype=torch.complex128 if prec == "d2z" else torch.complex64, device="cuda")
    dims = tuple(range(len(full) - len(shape), len(full)))
    ms_m, n_m, desc = ours(x, y, shape, batch, ftype, True)
    ym = y.clone()
    ms_g, n_g, _ = ours(x, y, shape, batch, ftype, False)
    variants = {}
    if os.environ.get("MIXED_PROBE_VARIANTS"):
        for name, env in [("maxr10", {"FFTB200_MIXED_MAXR": 10}), ("maxr8", {"FFTB200_MIXED_MAXR": 8}),
                          ("kb32", {"FFTB200_MIXED_TILE_KB": 32}), ("kb64", {"FFTB200_MIXED_TILE_KB": 64})]:
            try:
                v_ms, _, v_desc = ours(x, y, shape, batch, ftype, True, env)
                variants[name] = round(v_ms, 4)
            except Exception as ex:
                variants[name] = str(ex)
    cufft = (lambda: torch.fft.rfftn(x, dim=dims)) if real else (lambda: torch.fft.fftn(x, dim=dims))
    ref = cufft()
    ms_c = timed(cufft)
    rel = lambda a: float((torch.linalg.vector_norm((a - ref).to(torch.complex128)) / torch.linalg.vector_norm(ref.to(torch.complex128))).item())
    nbytes = (x.numel() * x.element_size() + y.numel() * y.element_size()) / 2  # mean of input and output array
    print(json.dumps({"kind": prec, "shape": shape, "batch": batch, "MiB": round(nbytes / 2**20, 1),
                      "mixed_ms": round(ms_m, 4), "mixed_launches": n_m, "mixed_GB/s_per_pass": round(2 * nbytes * n_m / ms_m / 1e6),
                      "generic_ms": round(ms_g, 4), "generic_launches": n_g, "cufft_ms": round(ms_c, 4),
                      "mixed_vs_cufft": round(ms_c / ms_m, 3), "mixed_vs_generic": round(ms_g / ms_m, 2),
                      "rel_l2_mixed_vs_cufft": rel(ym), "rel_l2_generic_vs_cufft": rel(y), "variants": variants, "plan": desc}), flush=True)
    del x, y, ym, ref
