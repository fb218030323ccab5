"""Slab-decomposed 3-D transforms over the GPUs of one box: one process (rank) per GPU.

The reference has no distributed transform (README.md:117-119 "Future Developments"; its
`make_plan_distrib`, src/fft.rg:513-537, runs independent shard FFTs).  This module is the host side
of libfft_b200's slab plans (include/fft_b200.h, csrc/slab_plan.cu), laid out like the vendored
FFTW-MPI (fftw-3.3.8/mpi/dft-rank-geq2.c:40-59, doc/mpi.texi:259-270, 443-466):

    rank r holds  in  [n0/G][n1][n2]     (slab r of dimension 0)
    and gets      out [n1/G][n0][n2c]    (slab r of dimension 1: FFTW_MPI_TRANSPOSED_OUT)

`torch.distributed` is plumbing only: it carries the 64-byte IPC handles once (mode "p2p", after which
the exchange is the FFT kernel's own stores into peer memory over NVLink) or runs the all-to-all
(mode "nccl": `all_to_all_single` between the library's pre and post halves).

The local FFT work is done by an *engine*.  The product engine is the CUDA library; tests inject a
CPU engine (built on the oracle) to check the decomposition and exchange bookkeeping under gloo.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def slab_shapes(shape, world: int, real: bool):
    """(local input shape, local output shape, blocks shape of the all-to-all buffers)."""
    n0, n1, n2 = (int(v) for v in shape)
    assert n0 % world == 0 and n1 % world == 0, "n0 and n1 must be divisible by the number of ranks"
    n2c = n2 // 2 + 1 if real else n2
    return (n0 // world, n1, n2), (n1 // world, n0, n2c), (world, n0 // world, n1 // world, n2c)


def slab_chunks(n2c: int, chunks: int) -> int:
    """number of pipeline chunks the library makes of the contiguous index (csrc/slab_plan.cu)"""
    cw = -(-n2c // max(1, chunks))
    cw = -(-cw // 16) * 16
    return -(-n2c // cw)


def assemble_transposed(parts):
    """Natural-order [n0][n1][n2c] array from the ranks' transposed-out slabs [n1/G][n0][n2c]."""
    return np.concatenate([np.asarray(p) for p in parts], axis=0).transpose(1, 0, 2)


class CudaSlabEngine:
    """The product path: libfft_b200 slab plan on this rank's GPU."""

    def __init__(self, shape, ftype, rank, world, device, chunks):
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            if len(shape) == 2:
                self.h = _lib.slab_plan_2d(list(shape), ftype, rank, world)
            else:
                self.h = _lib.slab_plan(list(shape), ftype, rank, world, chunks)
        self.ftype = ftype

    def bind_stream(self):
        _lib.set_stream(self.h, torch.cuda.current_stream(self.device).cuda_stream)

    def pre(self, x, send):
        self.bind_stream()
        _lib.slab_exec_pre(self.h, x.data_ptr(), send.data_ptr())

    def post(self, recv, out):
        self.bind_stream()
        _lib.slab_exec_post(self.h, recv.data_ptr(), out.data_ptr())

    def fused(self, x, out):
        self.bind_stream()
        _lib.slab_exec(self.h, x.data_ptr(), out.data_ptr())

    def connect(self, group):
        """all-gather the exchange areas' IPC handles (64 bytes per rank) and map the peers"""
        mine = torch.frombuffer(bytearray(_lib.slab_ipc_handle(self.h)), dtype=torch.uint8).to(self.device)
        world = dist.get_world_size(group)
        allh = torch.empty(64 * world, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(allh, mine, group=group)
        _lib.slab_connect_ipc(self.h, bytes(allh.cpu().numpy().tobytes()))

    def destroy(self):
        if self.h:
            _lib.destroy(self.h)
            self.h = 0


class SlabFFT3D:
    """One rank's view of a slab-decomposed forward transform of `shape` = (n0, n1, n2).

    dtype_in/out follow generate_fft_interface: complex64->complex64 (Z2Z), complex32->complex32 (C2C),
    double->complex64 (D2Z), float->complex32 (R2C).  execute() is collective.
    """

    def __init__(self, shape, dtype_in, dtype_out=None, rank=None, world=None, device=None, mode="p2p", chunks=0,
                 group=None, engine=None):
        from . import _dtype, complex32, complex64
        self.shape = tuple(int(v) for v in shape)
        assert len(self.shape) == 3, "slab decomposition is for 3-D transforms"
        self.dtype_in = _dtype(dtype_in)
        self.dtype_out = _dtype(dtype_out) if dtype_out is not None else (
            self.dtype_in if not self.dtype_in.is_real else (complex64 if self.dtype_in.size == 8 else complex32))
        self.real = self.dtype_in.is_real
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.mode = mode
        assert mode in ("p2p", "nccl")
        if self.dtype_out.size == 8:
            self.ftype = _lib.R2C if self.real else _lib.C2C
        else:
            self.ftype = _lib.D2Z if self.real else _lib.Z2Z
        self.local_in_shape, self.local_out_shape, self.blocks_shape = slab_shapes(self.shape, self.world, self.real)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.chunks = chunks if mode == "p2p" else 1
        self.engine = engine if engine is not None else CudaSlabEngine(self.shape, self.ftype, self.rank, self.world,
                                                                       self.device, self.chunks)
        self.out = torch.empty(self.local_out_shape, dtype=self.dtype_out.torch, device=self.device)
        self.x = None
        self.send = self.recv = None
        if mode == "nccl" or engine is not None:
            self.send = torch.empty(self.blocks_shape, dtype=self.dtype_out.torch, device=self.device)
            self.recv = torch.empty(self.blocks_shape, dtype=self.dtype_out.torch, device=self.device)
        if mode == "p2p" and engine is None and self.world > 1:
            # mapping the peers' exchange areas can fail on one rank only (IPC disabled, no P2P path): agree on
            # the outcome collectively, otherwise the ranks that succeeded would wait for flags that never come
            ok = 1
            try:
                self.engine.connect(group)
            except _lib.FFTB200Error as ex:
                ok = 0
                self.connect_error = str(ex)
            flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                self.mode = "nccl"       # staged exchange through all_to_all_single; still the CUDA library's passes
                self.send = torch.empty(self.blocks_shape, dtype=self.dtype_out.torch, device=self.device)
                self.recv = torch.empty(self.blocks_shape, dtype=self.dtype_out.torch, device=self.device)
        n0, n1, n2 = self.shape
        n2c = self.local_out_shape[2]
        ce = self.dtype_out.size
        # algorithmic bytes per rank and step
        self.local_pass_bytes = (self.dtype_in.size * n0 * n1 * n2 + ce * n0 * n1 * n2c * 5) // self.world
        self.exchange_bytes_out = ce * n0 * n1 * n2c // self.world * (self.world - 1) // self.world

    # ---- execution -----------------------------------------------------------------------------
    def set_input(self, x: torch.Tensor):
        assert tuple(x.shape) == self.local_in_shape and x.dtype == self.dtype_in.torch and x.is_contiguous()
        self.x = x

    def execute(self, x: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
        if x is not None:
            self.set_input(x)
        out = self.out if out is None else out
        if self.mode == "p2p" and self.send is None:
            self.engine.fused(self.x, out)
            return out
        self.engine.pre(self.x, self.send)
        if self.world > 1:
            dist.all_to_all_single(torch.view_as_real(self.recv), torch.view_as_real(self.send), group=self.group)
            recv = self.recv
        else:
            recv = self.send
        self.engine.post(recv, out)
        return out

    @property
    def launches_per_step(self) -> int:
        if self.mode == "p2p" and hasattr(self.engine, "h"):
            return _lib.launch_count(self.engine.h)
        return 3

    def describe(self) -> str:
        ex = ("y-axis FFT pass stores into peer HBM over NVLink (fused exchange, chunks=%s)" % (self.chunks or "auto")
              if self.mode == "p2p" else "NCCL all_to_all_single between the y- and z-axis passes")
        return f"slab x{self.world}: dim0 -> dim1 (transposed-out), {ex}"

    def roofline(self, ms_step: float, hbm_peak: float, peak_src: str, nvlink_peak: float = 770.0):
        """Step-level roofline for N>1: the step is bound by max(local HBM passes, NVLink exchange)."""
        t = ms_step * 1e-3
        hbm_t = self.local_pass_bytes / (hbm_peak * 1e9)
        nvl_t = self.exchange_bytes_out / (nvlink_peak * 1e9)
        bound = "nvlink" if nvl_t > hbm_t else "hbm"
        ach = (self.exchange_bytes_out if bound == "nvlink" else self.local_pass_bytes) / t / 1e9
        peak = nvlink_peak if bound == "nvlink" else hbm_peak
        return {"bound": bound, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "peak_source": peak_src + "; NVLink 770 GB/s per direction measured peer copy (B200_PROFILING.md)",
                # what kernels reach when all 8 GPUs exchange at once (tools/peer_probe.cu, profiles/r02_peer_probe_n8.jsonl)
                "nvlink_all_to_all_ceiling_GB/s": 645.0,
                "frac_of_all_to_all_ceiling": (ach / 645.0) if bound == "nvlink" else None,
                "per_rank": {"hbm_pass_model_bytes": self.local_pass_bytes, "nvlink_bytes_out": self.exchange_bytes_out,
                             "hbm_GB/s": self.local_pass_bytes / t / 1e9, "nvlink_GB/s_out": self.exchange_bytes_out / t / 1e9,
                             "ideal_ms_overlapped": max(hbm_t, nvl_t) * 1e3, "ideal_ms_serial": (hbm_t + nvl_t) * 1e3}}

    def make_host_pipeline(self, x_dev: torch.Tensor, slots: int = 2):
        """e2e steps: this rank's slab starts in pinned host memory and the result ends there.

        `slots` transforms are in flight: slot s owns a pinned input, a device input, a device output and a pinned
        output; host->device copies, the slab transform (collective, on the current stream) and device->host copies
        run on three streams chained by events, so step k's D2H and step k+1's H2D overlap step k's/k+1's
        transform on the full-duplex host link.  Returns (run(nsteps), h2d_bytes, d2h_bytes, host_outputs).
        run() records its end on the current stream (the caller times with events there).
        """
        dev = self.device
        cur = torch.cuda.current_stream(dev)
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        hx = [torch.empty(self.local_in_shape, dtype=self.dtype_in.torch, pin_memory=True) for _ in range(slots)]
        hy = [torch.empty(self.local_out_shape, dtype=self.dtype_out.torch, pin_memory=True) for _ in range(slots)]
        xin = [torch.empty_like(x_dev) for _ in range(slots)]
        outs = [torch.empty_like(self.out) for _ in range(slots)]
        for h in hx:
            h.copy_(x_dev)
        ev_in = [torch.cuda.Event() for _ in range(slots)]      # H2D of the slot finished
        ev_fft = [torch.cuda.Event() for _ in range(slots)]     # transform of the slot finished (xin free, out ready)
        ev_out = [torch.cuda.Event() for _ in range(slots)]     # D2H of the slot finished (out free)
        torch.cuda.synchronize(dev)

        def run(nsteps: int):
            s_in.wait_stream(cur)
            s_out.wait_stream(cur)
            for k in range(nsteps):
                s = k % slots
                with torch.cuda.stream(s_in):
                    if k >= slots:
                        s_in.wait_event(ev_fft[s])          # the previous transform of this slot has read xin[s]
                    xin[s].copy_(hx[s], non_blocking=True)
                    ev_in[s].record(s_in)
                cur.wait_event(ev_in[s])
                if k >= slots:
                    cur.wait_event(ev_out[s])               # the previous result of this slot has left outs[s]
                self.execute(xin[s], outs[s])
                ev_fft[s].record(cur)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_fft[s])
                    hy[s].copy_(outs[s], non_blocking=True)
                    ev_out[s].record(s_out)
            cur.wait_stream(s_out)
            cur.wait_stream(s_in)

        return run, hx[0].numel() * hx[0].element_size(), hy[0].numel() * hy[0].element_size(), hy

    def gather_natural(self, out: torch.Tensor | None = None) -> np.ndarray:
        """all ranks: the full result in natural order [n0][n1][n2c] (tests; not a hot path)"""
        out = self.out if out is None else out
        if self.world == 1:
            return assemble_transposed([out.cpu().numpy()])
        parts = [torch.empty_like(out) for _ in range(self.world)]
        dist.all_gather([torch.view_as_real(p) for p in parts], torch.view_as_real(out.contiguous()), group=self.group)
        return assemble_transposed([p.cpu().numpy() for p in parts])

    def destroy(self):
        """collective: no rank frees its exchange area while a peer may still store into it"""
        if self.world > 1 and dist.is_initialized():
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
        if hasattr(self.engine, "destroy"):
            self.engine.destroy()


class SlabFFT2D:
    """One rank's view of a slab-decomposed forward 2-D complex transform of `shape` = (n0, n1):
    rank r holds in [n0/G][n1] and gets out [n1/G][n0] (FFTW-MPI's transposed-out layout, doc/mpi.texi:443-466).
    The row FFT's store is the global transpose into the peers' receive slabs over NVLink; execute() is collective."""

    def __init__(self, shape, dtype, rank=None, world=None, device=None, group=None):
        from . import _dtype
        self.shape = tuple(int(v) for v in shape)
        assert len(self.shape) == 2
        self.dtype = _dtype(dtype)
        assert not self.dtype.is_real, "2-D slabs are complex-to-complex"
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        n0, n1 = self.shape
        assert n0 % self.world == 0 and n1 % self.world == 0
        self.local_in_shape, self.local_out_shape = (n0 // self.world, n1), (n1 // self.world, n0)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        ftype = _lib.C2C if self.dtype.size == 8 else _lib.Z2Z
        self.engine = CudaSlabEngine(self.shape, ftype, self.rank, self.world, self.device, 1)
        self.out = torch.empty(self.local_out_shape, dtype=self.dtype.torch, device=self.device)
        if self.world > 1:
            ok = 1
            try:
                self.engine.connect(group)
            except _lib.FFTB200Error:
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                raise RuntimeError("2-D slab transforms need peer access between the GPUs (fused exchange only)")

    def execute(self, x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        assert tuple(x.shape) == self.local_in_shape and x.dtype == self.dtype.torch and x.is_contiguous()
        out = self.out if out is None else out
        self.engine.fused(x, out)
        return out

    def gather_natural(self, out: torch.Tensor | None = None) -> np.ndarray:
        out = self.out if out is None else out
        if self.world == 1:
            return out.cpu().numpy().T
        parts = [torch.empty_like(out) for _ in range(self.world)]
        dist.all_gather([torch.view_as_real(p) for p in parts], torch.view_as_real(out.contiguous()), group=self.group)
        return np.concatenate([p.cpu().numpy() for p in parts], axis=0).T

    def destroy(self):
        if self.world > 1 and dist.is_initialized():
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
        self.engine.destroy()
