#!/usr/bin/env python
"""Regenerates regent/fft_rg.patch from the reference tree (run in the dev container, where /root/reference exists):

    python regent/make_patch.py [/root/reference]

The patch swaps the cuFFT binding of src/fft.rg for libfft_b200 (include/fft_b200.h) and fixes the plan-lifecycle
debts next to it.  It is produced by exact textual substitutions so that a changed reference
fails loudly here instead of yielding a patch that no longer applies.  tests/test_regent_patch.py applies it to a
scratch copy of the reference with `patch -p1` and checks the result.
"""
import difflib
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def sub(text, old, new, count=1):
    assert text.count(old) == count, f"expected {count} occurrence(s), found {text.count(old)}: {old[:70]!r}"
    return text.replace(old, new)


def patch_fft_rg(t):
    # ---- binding (src/fft.rg:15-20)
    t = sub(t, '''--Import cuFFT API
local cufft_c
if default_foreign then
  cufft_c = terralib.includec("cufftXt.h")
  terralib.linklibrary("libcufft.so")
end''', '''--Import the B200 FFT engine (include/fft_b200.h; same call shapes and result codes as cuFFT)
local b200_c
if default_foreign then
  b200_c = terralib.includec("fft_b200.h")
  terralib.linklibrary("libfft_b200.so")
end''')
    # ---- plan field space (src/fft.rg:49-54): a 64-bit POD handle, 0 = null, safe to copy with the instance
    t = sub(t, "      cufft_p : cufft_c.cufftHandle,", "      b200_p : b200_c.fftb200_handle,")
    # ---- plan creation (src/fft.rg:231-243 and :387-399)
    for ty in ("R2C", "C2C", "D2Z", "Z2Z"):
        t = sub(t, f"ok = cufft_c.cufftPlanMany(&p.cufft_p, dim, &n[0], [&int](0), 0, 0, [&int](0), 0, 0, cufft_c.CUFFT_{ty}, 1)",
                f"ok = b200_c.fftb200_plan_many(&p.b200_p, dim, &n[0], [&int](0), 0, 0, [&int](0), 0, 0, b200_c.FFTB200_{ty}, 1)")
        # (the Z2Z line also survives as a comment inside make_plan_batch, src/fft.rg:498: both spellings change)
        t = sub(t, f"ok = cufft_c.cufftPlanMany(&p.cufft_p, dim-1, &n_batch[0], &n_batch[0], 1, i_dist, &n_batch[0], 1, i_dist, cufft_c.CUFFT_{ty}, n[dim-1])",
                f"ok = b200_c.fftb200_plan_many(&p.b200_p, dim-1, &n_batch[0], &n_batch[0], 1, i_dist, &n_batch[0], 1, i_dist, b200_c.FFTB200_{ty}, n[dim-1])",
                count=2 if ty == "Z2Z" else 1)
    t = sub(t, '''        if ok == cufft_c.CUFFT_INVALID_VALUE then
          format.println("Invalid value in cufftPlanMany")
        end

        regentlib.assert(ok == cufft_c.CUFFT_SUCCESS, "cufftPlanMany failed")''', '''        if ok == b200_c.FFTB200_INVALID_VALUE then
          format.println("Invalid value in fftb200_plan_many")
        end

        regentlib.assert(ok == b200_c.FFTB200_SUCCESS, "fftb200_plan_many failed")''', count=2)
    # the batch task copies dim entries into an array of dim-1 (src/fft.rg:367-370): stop at dim-1
    t = sub(t, '''        var n_batch : int[dim-1]
        for i = 0, dim do
          n_batch[i] = n[i]
        end''', '''        var n_batch : int[dim-1]
        for i = 0, dim-1 do
          n_batch[i] = n[i]
        end''')
    t = sub(t, '''      var n_batch : int[dim-1]
      for i = 0, dim do
        n_batch[i] = n[i]
      end''', '''      var n_batch : int[dim-1]
      for i = 0, dim-1 do
        n_batch[i] = n[i]
      end''', count=2)
    # make_plan_batch hands its GPU child p.address_space before assigning it (src/fft.rg:455 vs :504)
    t = sub(t, "        make_plan_gpu_batch(input, output, plan, p.address_space)",
            "        make_plan_gpu_batch(input, output, plan, address_space)")
    # ---- make_plan_distrib zero-fills the plan region (src/fft.rg:523-531)
    t = sub(t, "      p.cufft_p = 0", "      p.b200_p = 0")
    t = sub(t, "Calls cufftPlanMany and stores plan in cufft_p", "Calls fftb200_plan_many and stores plan in b200_p", count=2)
    # ---- plan lifecycle: every handle starts null, so destroy can tell what was created (src/fft.rg:268, 642)
    t = sub(t, '''    var p = iface.get_plan(plan, false)

    --Get_executing process''', '''    var p = iface.get_plan(plan, false)
    p.p = [fftw_c.fftw_plan](0)
    p.float_p = [fftw_c.fftwf_plan](0)
    ;[default_foreign and rquote p.b200_p = 0 end or rquote end]

    --Get_executing process''', count=t.count('''    var p = iface.get_plan(plan, false)

    --Get_executing process'''))
    t = sub(t, '''    var p = iface.get_plan(plan, false)

    var address_space = c.legion_processor_address_space''', '''    var p = iface.get_plan(plan, false)
    p.p = [fftw_c.fftw_plan](0)
    p.float_p = [fftw_c.fftwf_plan](0)
    ;[default_foreign and rquote p.b200_p = 0 end or rquote end]

    var address_space = c.legion_processor_address_space''')
    # ---- execute (src/fft.rg:569-591); the float R2C call the reference left commented out is live
    t = sub(t, "        --ok = cufft_c.cufftExecR2C(p.cufft_p, [&cufft_c.cufftReal](input_base), [&cufft_c.cufftComplex](output_base))",
            "        ok = b200_c.fftb200_exec_r2c(p.b200_p, [&opaque](input_base), [&opaque](output_base))")
    t = sub(t, "        ok = cufft_c.cufftExecC2C(p.cufft_p, [&cufft_c.cufftComplex](input_base), [&cufft_c.cufftComplex](output_base), cufft_c.CUFFT_FORWARD)",
            "        ok = b200_c.fftb200_exec_c2c(p.b200_p, [&opaque](input_base), [&opaque](output_base), b200_c.FFTB200_FORWARD)")
    t = sub(t, "        ok = cufft_c.cufftExecD2Z(p.cufft_p, [&cufft_c.cufftDoubleReal](input_base), [&cufft_c.cufftDoubleComplex](output_base))",
            "        ok = b200_c.fftb200_exec_d2z(p.b200_p, [&opaque](input_base), [&opaque](output_base))")
    t = sub(t, "        ok = cufft_c.cufftExecZ2Z(p.cufft_p, [&cufft_c.cufftDoubleComplex](input_base), [&cufft_c.cufftDoubleComplex](output_base), cufft_c.CUFFT_FORWARD)",
            "        ok = b200_c.fftb200_exec_z2z(p.b200_p, [&opaque](input_base), [&opaque](output_base), b200_c.FFTB200_FORWARD)")
    t = sub(t, '''      if ok == cufft_c.CUFFT_INVALID_VALUE then
          format.println("Invalid value in cufftExecZ2Z")
      elseif ok == cufft_c.CUFFT_INVALID_PLAN then
          format.println("Invalid plan passed to cufftExecZ2Z")
      end''', '''      if ok == b200_c.FFTB200_INVALID_VALUE then
          format.println("Invalid value in fftb200_exec")
      elseif ok == b200_c.FFTB200_INVALID_PLAN then
          format.println("Invalid plan passed to fftb200_exec")
      end''')
    t = sub(t, '''      regentlib.assert(ok == cufft_c.CUFFT_SUCCESS, "cufftExecZ2Z failed")''',
            '''      regentlib.assert(ok == b200_c.FFTB200_SUCCESS, "fftb200_exec failed")''')
    # ---- destroy (src/fft.rg:634-642): destroy_plan is inlined into CPU tasks, so the TOC branch never ran and the GPU
    # plan leaked; fftb200_destroy needs no current device and ignores the null handle, fftw_destroy_plan must not see an
    # unset pointer (float mode never creates p.p)
    t = sub(t, '''    -- If using GPUs, call cufftDestroy
    if c.legion_processor_kind(proc) == c.TOC_PROC then
      c.printf("Destroy plan via cuFFT\\n")

      --Function: cufftResult cufftDestroy(cufftHandle plan)
      cufft_c.cufftDestroy(p.cufft_p) 
    else
      -- Else, call fftw_destroy
      c.printf("Destroy plan via FFTW\\n")
      fftw_c.fftw_destroy_plan(p.p)
      --fftw_c.fftwf_destroy_plan(p.float_p)
    end''', '''    -- The GPU plan (if one was made) goes first: callable from any processor kind, null handle is a no-op
    ;[default_foreign and rquote
      if p.b200_p ~= 0 then
        c.printf("Destroy plan via libfft_b200\\n")
        b200_c.fftb200_destroy(p.b200_p)
        p.b200_p = 0
      end
    end or rquote end]
    -- Then the FFTW plan, when make_plan created one (double precision only)
    if p.p ~= [fftw_c.fftw_plan](0) then
      c.printf("Destroy plan via FFTW\\n")
      fftw_c.fftw_destroy_plan(p.p)
      p.p = [fftw_c.fftw_plan](0)
    end''')
    assert "cufft_c" not in t and "cufft_p" not in t, "a cuFFT binding survived"
    return t


def patch_fft_test_rg(t):
    # test/fft_test.rg:9-10 binds cuFFT unconditionally and never uses it
    return sub(t, '''local cufft_c = terralib.includec("cufftXt.h")
terralib.linklibrary("libcufft.so")
''', "")


def main():
    out = []
    for rel, fn in (("src/fft.rg", patch_fft_rg), ("test/fft_test.rg", patch_fft_test_rg)):
        with open(os.path.join(REF, rel)) as f:
            old = f.read()
        new = fn(old)
        out += difflib.unified_diff(old.splitlines(True), new.splitlines(True), "a/" + rel, "b/" + rel, n=3)
    with open(os.path.join(HERE, "fft_rg.patch"), "w") as f:
        f.writelines(out)
    print(f"wrote {os.path.join(HERE, 'fft_rg.patch')}: {sum(1 for l in out if l.startswith('@@'))} hunks")


if __name__ == "__main__":
    main()
