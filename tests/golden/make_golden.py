"""Generates tests/golden/fftw_golden.npz with the REFERENCE's own FFTW (oracle/_ref, compiled
from /root/reference/fftw-3.3.8): seeded inputs -> outputs of the exact calls fft.rg makes
(fftw_plan_dft / fftw_plan_dft_r2c / fftw_plan_many_dft(_r2c), FFTW_FORWARD, FFTW_ESTIMATE).
Run in the dev container:  python tests/golden/make_golden.py
The fixtures travel to the GPU box; /root/reference does not."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

CASES = [  # (name, kind, shape, seed)
    ("c1_1d_c64_1024", "c2c", (1024,), 1),        # BASELINE config 1
    ("t1d_5", "c2c", (5,), 2), ("t1d_3", "c2c", (3,), 3), ("t2d_2x2", "c2c", (2, 2), 4),
    ("t3d_3x2x2", "c2c", (3, 2, 2), 5), ("t3d_3x3x2", "c2c", (3, 3, 2), 6),
    ("p2_64", "c2c", (64,), 7), ("p2_4096", "c2c", (4096,), 8), ("p2_8192", "c2c", (8192,), 9),
    ("p2_2d_32x64", "c2c", (32, 64), 10), ("p2_3d_16", "c2c", (16, 16, 16), 11),
    ("mix_12x10", "c2c", (12, 10), 12), ("prime_97", "c2c", (97,), 13),
    ("r_1d_3", "r2c", (3,), 14), ("r_1d_4096", "r2c", (4096,), 15), ("r_2d_64x64", "r2c", (64, 64), 16),
    ("r_3d_16x8x32", "r2c", (16, 8, 32), 17), ("r_3d_3x3x2", "r2c", (3, 3, 2), 18), ("r_1d_10", "r2c", (10,), 19),
]


def main():
    F = oracle.FFTW.get("ref")
    P = oracle.FFTW.get("prebuilt")
    out = {}
    for name, kind, shape, seed in CASES:
        if kind == "c2c":
            x = oracle.synth(shape, np.complex128, seed)
            y = F.dft(x)
            yp = P.dft(x)
        else:
            x = oracle.synth(shape, np.float64, seed)
            y = F.r2c(x)
            yp = P.r2c(x)
        assert oracle.rel_l2(yp, y) < 5e-16, name  # the author's binary agrees with our in-place build
        out[name + "__x"] = x
        out[name + "__y"] = y
    # the batched calls of make_plan_batch on {3,3,2} (src/fft.rg:483,500), flat buffers
    n = [3, 3]
    x = oracle.synth((18,), np.complex128, 20)
    y = np.zeros(18, np.complex128)
    F.dft_many(n, 2, x, n, 1, 9, y, n, 1, 9)
    out["batch_c_3x3x2__x"], out["batch_c_3x3x2__y"] = x, y
    xr = oracle.synth((18,), np.float64, 21)
    yr = np.zeros(18, np.complex128)
    F.r2c_many(n, 2, xr, n, 1, 9, yr, n, 1, 9)
    out["batch_r_3x3x2__x"], out["batch_r_3x3x2__y"] = xr, yr
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fftw_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out) // 2, "cases")


if __name__ == "__main__":
    main()
