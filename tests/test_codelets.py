"""In-register DFTs of radix_dft.cuh (the radices of the mixed-radix kernel) against the O(N^2) definition in long
double, compiled for the host: no GPU needed."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_radix_dfts_match_definition(tmp_path):
    exe = tmp_path / "codelet_check"
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", str(exe),
                    os.path.join(ROOT, "tests", "src", "codelet_check.cu")], check=True, capture_output=True, timeout=300)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    rows = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert sorted(row["n"] for row in rows) == [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]
    for row in rows:
        assert row["rel_l2_fp64"] < 4e-16 and row["rel_l2_fp32"] < 3e-7, row
    assert r.returncode == 0


def test_root_table_is_current():
    """dft_roots.inc is generated (tools/gen_dft_roots.py); its exact entries and a sample of the others are checked here"""
    import math
    import re
    txt = open(os.path.join(ROOT, "regent-fft-arjun_b200", "csrc", "dft_roots.inc")).read()
    tabs = re.findall(r"root_(cos|sin)<(\d+)>\(int e\) \{\s*constexpr double t\[\d+\] = \{([^}]*)\}", txt)
    assert len(tabs) == 26
    for fn, n, body in tabs:
        n = int(n)
        vals = [float(v) for v in body.split(",")]
        assert len(vals) == n
        for e, v in enumerate(vals):
            want = math.cos(2 * math.pi * e / n) if fn == "cos" else math.sin(2 * math.pi * e / n)
            assert abs(v - want) < 4e-15, (fn, n, e)  # (math.cos of a rounded angle is the looser side)
