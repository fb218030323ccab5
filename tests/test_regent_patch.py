"""regent/fft_rg.patch applies to the reference's src/fft.rg and test/fft_test.rg (dev container only: the reference
tree does not travel to the GPU box) and every libfft_b200 name it introduces is declared in include/fft_b200.h."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PATCH = os.path.join(ROOT, "regent", "fft_rg.patch")

needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "src", "fft.rg")), reason="reference tree absent")


def _declared_names():
    text = open(os.path.join(ROOT, "include", "fft_b200.h")).read()
    return set(re.findall(r"\b(fftb200_\w+|FFTB200_\w+)\b", text))


def test_patch_only_uses_declared_names():
    added = [l[1:] for l in open(PATCH) if l.startswith("+") and not l.startswith("+++")]
    used = set(re.findall(r"b200_c\.(\w+)", "".join(added)))
    assert used, "the patch binds nothing?"
    missing = used - _declared_names()
    assert not missing, f"names used by the patch but not declared in fft_b200.h: {sorted(missing)}"
    # result codes fft.rg compares against keep cuFFT's numeric values (src/fft.rg:246-250, 584-591)
    hdr = open(os.path.join(ROOT, "include", "fft_b200.h")).read()
    for name, val in (("FFTB200_SUCCESS", 0), ("FFTB200_INVALID_PLAN", 1), ("FFTB200_INVALID_VALUE", 4)):
        assert re.search(rf"{name}\s*=\s*{val}\b", hdr), name


@needs_ref
def test_patch_applies_to_reference(tmp_path):
    for rel in ("src/fft.rg", "test/fft_test.rg"):
        dst = tmp_path / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copy(os.path.join(REF, rel), dst)
    dry = subprocess.run(["patch", "--dry-run", "-p1", "-d", str(tmp_path), "-i", PATCH], capture_output=True, text=True)
    assert dry.returncode == 0, dry.stdout + dry.stderr
    assert "fuzz" not in dry.stdout and "offset" not in dry.stdout, dry.stdout   # applies exactly, not approximately
    res = subprocess.run(["patch", "-p1", "-d", str(tmp_path), "-i", PATCH], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    new = (tmp_path / "src" / "fft.rg").read_text()
    assert "cufft_c" not in new and "cufft_p" not in new and "libcufft" not in new
    assert 'terralib.includec("fft_b200.h")' in new and 'terralib.linklibrary("libfft_b200.so")' in new
    # the four plan types and the four exec calls of src/fft.rg:231-243, 387-399, 569-581
    for ty in ("R2C", "C2C", "D2Z", "Z2Z"):
        assert new.count(f"b200_c.FFTB200_{ty}") >= 2
        assert f"b200_c.fftb200_exec_{ty.lower()}(p.b200_p" in new
    assert "b200_c.fftb200_destroy(p.b200_p)" in new
    assert "libcufft" not in (tmp_path / "test" / "fft_test.rg").read_text()


@needs_ref
def test_patch_is_what_the_generator_makes(tmp_path):
    """the committed patch is reproducible from the reference with regent/make_patch.py"""
    work = tmp_path / "regent"
    work.mkdir()
    shutil.copy(os.path.join(ROOT, "regent", "make_patch.py"), work / "make_patch.py")
    res = subprocess.run([sys.executable, str(work / "make_patch.py"), REF], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert (work / "fft_rg.patch").read_text() == open(PATCH).read()
