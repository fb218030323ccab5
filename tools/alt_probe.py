"""Per-pass timings of one transform for each alternative tile-table row (FFTB200_TILE_ALT=k); tools only."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
fft = load_package(); L = fft._lib

def run(kind, shape, alts, env=None):
    real = kind in ("d2z", "r2c")
    dt = {"z2z": torch.complex128, "c2c": torch.complex64, "d2z": torch.float64, "r2c": torch.float32}[kind]
    ftype = {"z2z": L.Z2Z, "c2c": L.C2C, "d2z": L.D2Z, "r2c": L.R2C}[kind]
    x = torch.zeros(shape, dtype=dt, device="cuda")
    (torch.view_as_real(x) if x.is_complex() else x).uniform_(-0.5, 0.5)
    oshape = list(shape[:-1]) + [shape[-1] // 2 + 1] if real else list(shape)
    y = torch.empty(oshape, dtype=torch.complex128 if kind in ("z2z", "d2z") else torch.complex64, device="cuda")
    ref = None
    for alt in alts:
        if isinstance(alt, tuple):
            os.environ["FFTB200_TILE_ALT_ROW"], os.environ["FFTB200_TILE_ALT_COL"] = str(alt[0]), str(alt[1])
        else:
            os.environ["FFTB200_TILE_ALT"] = str(alt)
        for k, v in (env or {}).items():
            os.environ[k] = str(v)
        try:
            h = L.plan_many(len(shape), list(shape), None, 0, 0, None, 0, 0, ftype, 1)
        except Exception as ex:
            print(json.dumps({"kind": kind, "shape": shape, "alt": alt, "error": str(ex)}), flush=True)
            continue
        L.set_stream(h, torch.cuda.current_stream().cuda_stream)
        nl = L.launch_count(h)
        desc = L.describe(h).strip().split("\n")
        for _ in range(3):
            L.execute(h, ftype, x.data_ptr(), y.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            L.execute(h, ftype, x.data_ptr(), y.data_ptr())
        e1.record(); torch.cuda.synchronize()
        tot = e0.elapsed_time(e1) / 5
        L.set_profiling(h, True)
        for _ in range(5):
            L.execute(h, ftype, x.data_ptr(), y.data_ptr())
        torch.cuda.synchronize()
        per = [(L.launch_ms(h, i), L.launch_bytes(h, i)) for i in range(nl)]
        L.set_profiling(h, False)
        if ref is None:
            ref = y.clone()
            diff = 0.0
        else:
            diff = float((torch.linalg.vector_norm((y - ref).to(torch.complex128)) / torch.linalg.vector_norm(ref.to(torch.complex128))).item())
        print(json.dumps({"kind": kind, "shape": shape, "alt": alt, "env": env, "ms": round(tot, 4), "rel_l2_vs_alt0": diff,
                          "passes": [f"{ms:.3f} ms {b / ms / 1e6:.0f} GB/s | {d.split(' lines=')[0]}" for (ms, b), d in zip(per, desc)]}), flush=True)
        L.destroy(h)
    for k in ("FFTB200_TILE_ALT", "FFTB200_TILE_ALT_ROW", "FFTB200_TILE_ALT_COL"):
        os.environ.pop(k, None)
    for k in (env or {}):
        os.environ.pop(k, None)
    del x, y, ref
    torch.cuda.empty_cache()

if __name__ == "__main__":
    kind = sys.argv[1]
    shape = tuple(int(v) for v in sys.argv[2].split(","))
    # "0,1,2" = the same alternative for every pass; "0:1,2:0" = (row alternative : column alternative) pairs
    alts = [tuple(int(q) for q in v.split(":")) if ":" in v else int(v) for v in sys.argv[3].split(",")]
    env = dict(kv.split("=") for kv in sys.argv[4:])
    run(kind, shape, alts, env)
