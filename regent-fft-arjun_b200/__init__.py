"""regent-fft-arjun_b200 — host-side mirror of Regent-FFT's interface over libfft_b200.

The reference exposes one Lua-time factory, `fft.generate_fft_interface(itype, dtype_in,
dtype_out)` (reference src/fft.rg:31-41, 663), returning a table of tasks:

    iface.plan, make_plan, make_plan_batch, make_plan_task, make_plan_distrib,
    execute_plan, execute_plan_task, destroy_plan, destroy_plan_task, destroy_plan_distrib,
    get_plan, get_tunable, get_num_nodes, get_num_local_gpus

Regent/Terra/Legion are not installable in this image, so this module restates that table in
Python over the same C ABI the Regent patch binds: same names, same argument
meaning, same assertions, same buffer conventions.  Every execute goes to the CUDA library
(`_lib.py`); there is no CPU branch here (the reference's CPU branch is FFTW, which lives only
in oracle/ as the checker).

Regions: a `Region` is what `make_get_base` (src/fft.rg:68-108) sees of a Legion physical
region — a base pointer, bounds and per-dimension byte offsets.  Data lives in one flat device
buffer; default byte offsets follow Legion's default instance layout (dimension 0 fastest),
which is what makes `i_dist = offsets[2]/offsets[0]` of make_plan_batch come out as n0*n1
(src/fft.rg:372-377).  As in the reference, `n[i]` = extent of dimension i is handed to the
engine as row-major sizes and the flat buffer is transformed "as if" it were `[n0][n1][n2]`
row-major (dimension-order contract) — the library does not reorder.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import FFTB200Error  # noqa: F401

__all__ = ["generate_fft_interface", "Region", "PlanRegion", "int1d", "int2d", "int3d",
           "double", "float32", "complex64", "complex32", "FFTB200Error"]


# ------------------------------------------------------------------------------------------
# Regent's types, by their Regent names (complex64 = 2 x fp64, complex32 = 2 x fp32)
# ------------------------------------------------------------------------------------------
class _IndexType:
    def __init__(self, dim):
        self.dim = dim

    def __repr__(self):
        return f"int{self.dim}d"


int1d, int2d, int3d = _IndexType(1), _IndexType(2), _IndexType(3)


class _DType:
    def __init__(self, name, torch_dtype, size, is_real):
        self.name, self.torch, self.size, self.is_real = name, torch_dtype, size, is_real

    def __repr__(self):
        return self.name


double = _DType("double", torch.float64, 8, True)
float32 = _DType("float", torch.float32, 4, True)
complex64 = _DType("complex64", torch.complex128, 16, False)
complex32 = _DType("complex32", torch.complex64, 8, False)
_BY_NAME = {"double": double, "float": float32, "complex64": complex64, "complex32": complex32}


def _dtype(x) -> _DType:
    if isinstance(x, _DType):
        return x
    if isinstance(x, str) and x in _BY_NAME:
        return _BY_NAME[x]
    raise TypeError(f"unknown Regent element type {x!r}")


# ------------------------------------------------------------------------------------------
# regions
# ------------------------------------------------------------------------------------------
class Region:
    """region(ispace(intNd, extent), dtype) held in device memory (GPU framebuffer)."""

    def __init__(self, extent, dtype, device="cuda", flat: torch.Tensor | None = None, lo=None, offsets=None):
        self.dtype = _dtype(dtype)
        extent = (extent,) if isinstance(extent, int) else tuple(int(e) for e in extent)
        self.dim = len(extent)
        self.lo = tuple(lo) if lo is not None else (0,) * self.dim
        self.hi = tuple(l + e - 1 for l, e in zip(self.lo, extent))
        n = int(np.prod(extent))
        if flat is None:
            flat = torch.zeros(n, dtype=self.dtype.torch, device=device)
        assert flat.dtype == self.dtype.torch and flat.is_contiguous() and flat.numel() >= n
        self.flat = flat
        if offsets is None:  # Legion default: dimension 0 fastest
            offsets, step = [], self.dtype.size
            for e in extent:
                offsets.append(step)
                step *= e
        self.offsets = tuple(offsets)

    # ispace.bounds
    @property
    def bounds(self):
        return (self.lo, self.hi)

    @property
    def extent(self):
        return tuple(h - l + 1 for l, h in zip(self.lo, self.hi))

    @property
    def volume(self):
        return int(np.prod(self.extent))

    def fill(self, value):
        self.flat.fill_(value)
        return self

    # make_get_base: raw pointer of the physical instance (src/fft.rg:76-93)
    def base_pointer(self) -> int:
        return self.flat.data_ptr()

    def numpy(self) -> np.ndarray:
        return self.flat.detach().cpu().numpy()

    @classmethod
    def from_numpy(cls, array: np.ndarray, dtype, device="cuda") -> "Region":
        """Region whose flat buffer is `array` in C order and whose extents are array.shape:
        exactly what the engine transforms (row-major n[])."""
        dt = _dtype(dtype)
        flat = torch.from_numpy(np.ascontiguousarray(array).ravel()).to(device=device, dtype=dt.torch)
        return cls(array.shape, dt, device=device, flat=flat)

    def partition_equal(self, ncolors: int):
        """partition(equal, r, ispace(int1d, n)) for 1-D regions (test/fft_test.rg:286-288)."""
        assert self.dim == 1 and self.volume % ncolors == 0
        m = self.volume // ncolors
        return [Region((m,), self.dtype, flat=self.flat[c * m:(c + 1) * m], lo=(self.lo[0] + c * m,))
                for c in range(ncolors)]


# the plan fieldspace (src/fft.rg:48-65): stored BY VALUE in a 1-element (or n_nodes-element)
# int1d region that must live in host-visible memory (src/fft.rg:165-169)
PLAN_FSPACE = np.dtype([("p", np.uint64), ("float_p", np.uint64), ("b200_p", np.uint64),
                        ("address_space", np.uint32), ("ftype", np.uint32)])


class PlanRegion:
    """region(ispace(int1d, n), iface.plan)"""

    def __init__(self, n: int = 1):
        self.data = np.zeros(n, dtype=PLAN_FSPACE)

    @property
    def volume(self):
        return len(self.data)

    def partition_equal(self, ncolors: int):
        assert self.volume % ncolors == 0
        m = self.volume // ncolors
        parts = []
        for c in range(ncolors):
            pr = PlanRegion.__new__(PlanRegion)
            pr.data = self.data[c * m:(c + 1) * m]
            parts.append(pr)
        return parts


def _address_space() -> int:
    """legion_processor_address_space: one address space per process (rank)."""
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


class _Interface:
    """The table `fft.generate_fft_interface` returns."""

    def __init__(self, itype, dtype_in, dtype_out):
        self.itype = itype
        self.dim = itype.dim
        self.dtype_in, self.dtype_out = _dtype(dtype_in), _dtype(dtype_out)
        self.dtype_size = self.dtype_out.size                    # src/fft.rg:34
        self.real_flag = self.dtype_in.is_real                    # src/fft.rg:36-39
        self.plan = PLAN_FSPACE                                   # iface.plan
        # cufftType selection of src/fft.rg:231-243
        if self.dtype_size == 8 and self.real_flag:
            self.ftype = _lib.R2C
        elif self.dtype_size == 8:
            self.ftype = _lib.C2C
        elif self.real_flag and self.dtype_size == 16:
            self.ftype = _lib.D2Z
        else:
            self.ftype = _lib.Z2Z

    # ---- tunables (src/fft.rg:124-153) -----------------------------------------------------
    DEFAULT_TUNABLE_NODE_COUNT, DEFAULT_TUNABLE_LOCAL_GPUS = 0, 2

    def get_tunable(self, tunable_id: int) -> int:
        import torch.distributed as dist
        if tunable_id == self.DEFAULT_TUNABLE_NODE_COUNT:
            return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if tunable_id == self.DEFAULT_TUNABLE_LOCAL_GPUS:
            return torch.cuda.device_count() if torch.cuda.is_available() else 0
        raise ValueError("unknown tunable")

    def get_num_nodes(self) -> int:
        return self.get_tunable(self.DEFAULT_TUNABLE_NODE_COUNT)

    def get_num_local_gpus(self) -> int:
        return self.get_tunable(self.DEFAULT_TUNABLE_LOCAL_GPUS)

    # ---- get_plan (src/fft.rg:156-189) -----------------------------------------------------
    def get_plan(self, plan: PlanRegion, check: bool):
        assert isinstance(plan, PlanRegion), "plan must be a region of iface.plan"
        i = _address_space() if plan.volume > 1 else 0
        assert i < plan.volume, "plan region too small for this address space"
        p = plan.data[i:i + 1]
        if check:
            assert int(p["address_space"][0]) == _address_space(), \
                "plans can only be used on the node where they are originally created"
        return p

    # ---- make_plan (src/fft.rg:261-333 + make_plan_gpu :195-258) ----------------------------
    def _check_regions(self, input: Region, output: Region):
        assert input.dim == self.dim and output.dim == self.dim, "region dimension does not match the interface"
        assert input.dtype is self.dtype_in and output.dtype is self.dtype_out, "region element types do not match"
        assert input.bounds == output.bounds, "input and output regions must be identical in size"  # :276
        assert input.flat.is_cuda and output.flat.is_cuda, \
            "regions must be GPU instances: libfft_b200 has no CPU path (no CUDA device => no transform)"

    def make_plan(self, input: Region, output: Region, plan: PlanRegion) -> None:
        self._check_regions(input, output)
        p = self.get_plan(plan, False)
        n = list(input.extent)                                    # n[i] = hi.x[i] - lo.x[i] + 1  (:220-223)
        with torch.cuda.device(input.flat.device):
            h = _lib.plan_many(self.dim, n, None, 0, 0, None, 0, 0, self.ftype, 1)   # :233-242
        p["b200_p"][0] = h
        p["ftype"][0] = self.ftype
        p["address_space"][0] = _address_space()                  # :322

    def make_plan_task(self, input, output, plan):                # src/fft.rg:506-511
        self.make_plan(input, output, plan)

    # ---- make_plan_batch (src/fft.rg:416-504 + make_plan_gpu_batch :336-414) ----------------
    def make_plan_batch(self, input: Region, output: Region, plan: PlanRegion) -> None:
        self._check_regions(input, output)
        assert self.dim == 3, "make_plan_batch reads offsets[2]: 3-D regions only (src/fft.rg:372-377)"
        p = self.get_plan(plan, False)
        n = list(input.extent)
        n_batch = n[:self.dim - 1]                                # :367-370
        i_dist = input.offsets[2] // input.offsets[0]             # :374-377
        with torch.cuda.device(input.flat.device):
            h = _lib.plan_many(self.dim - 1, n_batch, n_batch, 1, i_dist, n_batch, 1, i_dist,
                               self.ftype, n[self.dim - 1])       # :389-398
        p["b200_p"][0] = h
        p["ftype"][0] = self.ftype
        p["address_space"][0] = _address_space()

    # ---- make_plan_distrib (src/fft.rg:513-537) ---------------------------------------------
    def make_plan_distrib(self, input, input_part, output, output_part, plan: PlanRegion, plan_part) -> None:
        n = self.get_num_nodes()
        assert len(input_part) == n and len(output_part) == n and len(plan_part) == n, \
            "number of colors must match the number of nodes"      # :518-521
        plan.data[:] = 0                                          # null handles (:523-531)
        me = _address_space()
        for c in range(n):                                        # index launch; color c runs on node c
            if c == me:
                self.make_plan_task(input_part[c], output_part[c], plan_part[c] if plan.volume == 1 else plan)

    # ---- execute_plan (src/fft.rg:543-611), GPU branch only ----------------------------------
    def execute_plan(self, input: Region, output: Region, plan: PlanRegion, direction: int = _lib.FORWARD) -> None:
        p = self.get_plan(plan, True)
        h = int(p["b200_p"][0])
        stream = torch.cuda.current_stream(input.flat.device)
        _lib.set_stream(h, stream.cuda_stream)
        _lib.execute(h, self.ftype, input.base_pointer(), output.base_pointer(), direction)

    def execute_plan_task(self, input, output, plan):             # src/fft.rg:613-617
        self.execute_plan(input, output, plan)

    # ---- destroy (src/fft.rg:624-661) --------------------------------------------------------
    def destroy_plan(self, plan: PlanRegion) -> None:
        p = self.get_plan(plan, True)
        _lib.destroy(int(p["b200_p"][0]))
        p["b200_p"][0] = 0

    def destroy_plan_task(self, plan):
        self.destroy_plan(plan)

    def destroy_plan_distrib(self, plan: PlanRegion, plan_part) -> None:
        me = _address_space()
        for c, part in enumerate(plan_part):
            if c == me:
                self.destroy_plan_task(part if plan.volume == 1 else plan)

    # ---- helpers that have no reference counterpart -------------------------------------------
    def packed_output_shape(self, extent):
        """Shape of the data execute_plan writes from the output base pointer: R2C rows are packed to
        n_last/2+1 although the output region has full extent (Appendix A.2 of SURVEY.md)."""
        extent = tuple(extent)
        return extent[:-1] + (extent[-1] // 2 + 1,) if self.real_flag else extent


def generate_fft_interface(itype, dtype_in, dtype_out) -> _Interface:
    """fft.generate_fft_interface (reference src/fft.rg:31-41)."""
    assert isinstance(itype, _IndexType), "requires an index type as the first argument"
    assert 1 <= itype.dim <= 3, "currently only 1 <= dim <= 3 is supported"
    return _Interface(itype, dtype_in, dtype_out)
