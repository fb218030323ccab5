#!/usr/bin/env python
"""bench.py — the Regent-FFT hot path on B200: 3D C2C complex64 (fp64 parts) 512^3, forward, out of place.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
         bench.py --gpus N --steps K --warmup W

One "step" = one execute_plan over one 512^3 batch of synthetic input (BASELINE.json configs[3]; metric
5*N*log2(N)/t, fftw-3.3.8/libbench2/mflops.c:22-23).  N=1: the whole transform on one GPU through
libfft_b200's C ABI.  N>1: the same 512^3 transform slab-decomposed over N ranks (strong scaling), the
global transpose fused into the middle FFT pass as peer-to-peer stores over NVLink
(regent-fft-arjun_b200/distributed.py).

Printed JSON (rank 0, one line):
  value      whole-job GFLOP/s with the input resident in HBM, CUDA events, max over ranks
  e2e        same metric through the C ABI with HOST (pinned) in/out buffers: H2D + transform + D2H timed
  roofline   dominant kernel: algorithmic bytes per launch / mean launch time (events inside the timed
             region), against the measured HBM copy peak (MEASURED_PEAKS.json, else the recipe's fallback)
  cpu_baseline  the reference's FFTW 3.3.8 (oracle/_ref, compiled from the vendored sources) on this
             box's host cores, rank 0, N=1 only, bounded sample
  parity     after the timed loop, outside it: the step's output (gathered over the ranks for N>1) against
             FFTW 3.3.8 on the same input, rel-L2 <= 10*log2(N)*eps; the run exits non-zero when it fails
  scaling_1024  the north-star strong-scaling case, 3D C2C fp64 1024^3, timed the same way at this N
             (one GPU: the ordinary plan; N>1: slabs), with 128 sampled output bins checked against fp64
             direct sums
  --impl reference   times only that FFTW path (all host threads) and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SIDE = 512
SHAPE = (N_SIDE, N_SIDE, N_SIDE)
N_TOTAL = N_SIDE ** 3
FLOPS = 5.0 * N_TOTAL * math.log2(N_TOTAL)            # libbench2/mflops.c:22-23
ELT = 16                                              # complex64 = 2 x fp64
PASS_MODEL_BYTES = 3 * 2 * N_TOTAL * ELT              # three axis passes, each one read + one write
METRIC = "3D C2C fp64 512^3 GFLOP/s (5NlogN/t)"
WORKLOAD = "3D C2C complex64 (fp64) 512^3 forward out-of-place, BASELINE configs[3]"
FALLBACK_HBM_GBS = 6650.0                             # /opt/skills/guides/B200_PROFILING.md


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (NVML; same fields as the recipe's nvidia-smi line)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int, period_s: float = 0.005):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as ex:  # no NVML: report it rather than invent clocks
            self.err = str(ex)
        self.period = period_s

    def _names(self, mask):
        nv = self.nv
        table = [("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap"),
                 ("hw_power_brake", "nvmlClocksEventReasonHwPowerBrakeSlowdown", "nvmlClocksThrottleReasonHwPowerBrakeSlowdown")]
        out = []
        for name, a, b in table:
            bit = getattr(nv, a, None) or getattr(nv, b, None)
            if bit and (mask & bit):
                out.append(name)
        return out

    def _loop(self):
        nv = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                if get_reasons:
                    self.reasons.update(self._names(int(get_reasons(self.h))))
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if self.ok:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own FFTW on the host cores
# ------------------------------------------------------------------------------------------------
def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def pick_threads(F, all_threads: int) -> int:
    """FFTW spawns its workers per call; on a many-core host fewer threads than cores can be faster.
    Probe a 128^3 transform (tens of ms each) at a few counts and keep the fastest."""
    import numpy as np
    import oracle
    x = oracle.synth((128, 128, 128), np.complex128, seed=5)
    best, best_t = all_threads, None
    for t in sorted({all_threads, min(all_threads, 64), min(all_threads, 32), min(all_threads, 16)}, reverse=True):
        times, _ = F.time_transform(x, real=False, threads=t, reps=3, warmup=1)
        if best_t is None or min(times) < best_t:
            best, best_t = t, min(times)
    return best


def fftw_engine():
    """(engine, kind): the reference's FFTW compiled from its vendored sources (kind "reference")."""
    import oracle
    if oracle.have_fftw("ref") and not os.environ.get("FFTB200_BENCH_NO_FFTW"):   # (the variable: exercises the fall-backs)
        return oracle.FFTW.get("ref"), "reference"
    return None, "port"


def time_fftw(steps: int, warmup: int, budget_s: float, threads: int):
    """Times fftw_plan_dft(3, {512,512,512}, FORWARD, ESTIMATE) + fftw_execute_dft exactly as
    src/fft.rg:319,608 issue them, with `threads` host threads.  Returns (mean_s, reps_done, sample)."""
    import numpy as np
    import oracle
    F, kind = fftw_engine()
    x = oracle.synth(SHAPE, np.complex128, seed=4)
    if F is None:
        # plain-C port: far slower than FFTW; time one 2-D plane batch and say so
        sub = x[:4]
        t0 = time.perf_counter()
        oracle.port_dft(sub)
        dt = time.perf_counter() - t0
        est = dt * (N_SIDE / 4) * (27.0 / 18.0)
        return est, 1, "oracle port, 4 of 512 planes (2-D part), scaled by planes and log2 ratio", kind, 1
    threads = pick_threads(F, threads)
    t_begin = time.perf_counter()
    # one untimed plan + warm-up execute, then as many timed executes as fit in the budget (<= steps)
    w_times, _ = F.time_transform(x, real=False, threads=threads, reps=1, warmup=0)
    t_first = w_times[0]
    reps = max(1, min(steps, int((budget_s - (time.perf_counter() - t_begin)) / max(t_first, 1e-3))))
    extra_warm = max(0, min(warmup - 1, int(0.25 * budget_s / max(t_first, 1e-3))))
    times, _ = F.time_transform(x, real=False, threads=threads, reps=reps, warmup=extra_warm)
    mean = float(sum(times) / len(times))
    sample = (f"full 512^3 transform, FFTW 3.3.8 (vendored sources, threads+AVX2), ESTIMATE, {threads} of {host_threads()} host threads (fastest of a 128^3 probe), "
              f"{1 + extra_warm} warm-up + {reps} timed executes (mean; min {min(times):.3f} s)")
    return mean, reps, sample, kind, threads


def time_fftw_faithful():
    """What src/fft.rg's CPU branch really runs: the author's prebuilt scalar libfftw3.so (src/fft.rg:13), ESTIMATE,
    ONE thread (fft.rg never calls fftw_init_threads).  512^3 takes ~14 s that way, so the sample is 256^3."""
    import numpy as np
    import oracle
    try:
        if not oracle.have_fftw("prebuilt"):
            return {"value": None, "sample": "prebuilt libfftw3.so not present"}
        F = oracle.FFTW.get("prebuilt")
        n = 256
        x = oracle.synth((n, n, n), np.complex128, seed=6)
        times, _ = F.time_transform(x, real=False, threads=1, reps=2, warmup=0)
        t = min(times)
        fl = 5.0 * n ** 3 * math.log2(n ** 3)
        return {"value": fl / t / 1e9, "unit": "GFLOP/s", "cores": 1, "ms_per_transform": t * 1e3,
                "sample": "3D C2C fp64 256^3 (1/8 of the workload), reference's prebuilt scalar libfftw3.so.3.5.8, 1 thread, min of 2"}
    except Exception as ex:
        return {"value": None, "sample": f"failed: {ex}"}


def run_reference(args, rank: int) -> int:
    if rank != 0:
        return 0
    threads = host_threads()
    mean_s, reps, sample, kind, used = time_fftw(args.steps, args.warmup, budget_s=150.0, threads=threads)
    gf = FLOPS / mean_s / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gf, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean_s * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "plan": {"engine": "FFTW 3.3.8 CPU path of src/fft.rg:319,608", "timed_executes": reps},
        "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": used, "kind": kind, "sample": sample},
        "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0



# ------------------------------------------------------------------------------------------------
# host placement: run each rank (and first-touch its pinned buffers) on the NUMA node of its GPU
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local_rank: int):
    """Restrict this process to the CPUs local to its GPU's PCIe root (sysfs local_cpulist), so that pinned host
    buffers are first-touched on that node and the copies do not cross the socket interconnect.  Returns a note."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:       # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        base = "/sys/bus/pci/devices/" + bus
        with open(base + "/local_cpulist") as f:
            cpulist = f.read().strip()
        with open(base + "/numa_node") as f:
            node = int(f.read().strip())
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return f"numa node {node}: affinity unchanged ({len(allowed)} cpus)"
        os.sched_setaffinity(0, cpus)
        return f"bound to numa node {node} ({len(cpus)} cpus local to GPU {bus})"
    except Exception as ex:
        return f"not bound ({type(ex).__name__}: {ex})"


# ------------------------------------------------------------------------------------------------
# parity checks run after the timed loops (never inside them)
# ------------------------------------------------------------------------------------------------
def parity_vs_fftw(x_full, y_full, n_total: int):
    """rel-L2 of the transform's full output against the reference's FFTW on the same input (src/fft.rg:319,608).
    x_full, y_full: device tensors in natural order.  The comparison runs on the GPU slab by slab."""
    import numpy as np
    import torch
    import oracle
    F, kind = fftw_engine()
    if F is None:
        return None
    x = x_full.cpu().numpy()
    want = F.dft(x, threads=min(64, host_threads()))
    del x
    num = den = 0.0
    step = max(1, want.shape[0] // 8)
    for i0 in range(0, want.shape[0], step):
        w = torch.from_numpy(want[i0:i0 + step]).to(y_full.device)
        num += float(torch.linalg.vector_norm(y_full[i0:i0 + step] - w) ** 2)
        den += float(torch.linalg.vector_norm(w) ** 2)
        del w
    err = (num / den) ** 0.5
    tol = oracle.tolerance(n_total, single=False)
    return {"rel_l2": err, "tol": tol, "ok": bool(err <= tol),
            "against": "FFTW 3.3.8 (oracle/_ref, the reference's CPU path) on the same input, every output bin"}


def direct_bins(x_local, z0: int, shape, bins):
    """fp64 direct sums  X[k] = sum_n x[n] exp(-2 pi i (k0 n0/N0 + k1 n1/N1 + k2 n2/N2))  restricted to the planes
    n0 in [z0, z0 + x_local.shape[0]) this rank holds, for a list of bins.  Returns a complex128 device tensor [len(bins)]
    (partial sums: add them over the ranks)."""
    import torch
    n0, n1, n2 = shape
    dev = x_local.device
    kb = torch.tensor(bins, dtype=torch.int64, device=dev)                       # [B, 3]

    def phases(k, n, lo, cnt):                                                    # [cnt, B], exponent reduced mod n exactly
        j = torch.arange(lo, lo + cnt, dtype=torch.int64, device=dev)[:, None]
        m = (j * k[None, :]) % n
        ang = m.to(torch.float64) * (-2.0 * math.pi / n)
        return torch.complex(torch.cos(ang), torch.sin(ang))

    w2 = phases(kb[:, 2], n2, 0, n2)                                             # [n2, B]
    w1 = phases(kb[:, 1], n1, 0, n1)                                             # [n1, B]
    w0 = phases(kb[:, 0], n0, z0, x_local.shape[0])                              # [n0l, B]
    acc = torch.zeros(len(bins), dtype=torch.complex128, device=dev)
    planes = max(1, (1 << 26) // (n1 * len(bins)))                               # bound the [p, n1, B] intermediate
    for p0 in range(0, x_local.shape[0], planes):
        xs = x_local[p0:p0 + planes]                                             # [p, n1, n2]
        t = xs.reshape(-1, n2) @ w2                                              # [p*n1, B]
        t = (t.reshape(xs.shape[0], n1, -1) * w1[None]).sum(dim=1)               # [p, B]
        acc += (t * w0[p0:p0 + xs.shape[0]]).sum(dim=0)
    return acc


def sampled_bins_parity(x_local, out_local, rank: int, world: int, shape, count: int, seed: int):
    """rel-L2 over `count` sampled output bins between the transform's output and fp64 direct sums over the same input.
    x_local: this rank's planes [n0/world][n1][n2]; out_local: the whole output [n0][n1][n2] when world == 1, else this
    rank's transposed-out slab [n1/world][n0][n2].  Collective for world > 1."""
    import torch
    import torch.distributed as dist
    n0, n1, n2 = shape
    bins = pick_bins(shape, count, seed)
    want = direct_bins(x_local, rank * (n0 // world), shape, bins)
    got = torch.zeros(len(bins), dtype=torch.complex128, device=x_local.device)
    n1l = n1 // world
    for i, (k0, k1, k2) in enumerate(bins):
        if world == 1:
            got[i] = out_local[k0, k1, k2]
        elif k1 // n1l == rank:
            got[i] = out_local[k1 - rank * n1l, k0, k2]
    if world > 1:
        dist.all_reduce(torch.view_as_real(want))
        dist.all_reduce(torch.view_as_real(got))
    err = float((torch.linalg.vector_norm(got - want) / torch.linalg.vector_norm(want)).item())
    tol = 10.0 * math.log2(float(n0) * n1 * n2) * 2.220446049250313e-16
    return {"rel_l2": err, "tol": tol, "ok": bool(err <= tol), "bins": len(bins),
            "against": "fp64 direct sums over the same input at %d sampled output bins" % len(bins)}


def pick_bins(shape, count: int, seed: int):
    import random
    r = random.Random(seed)
    bins = [(0, 0, 0), (shape[0] - 1, shape[1] - 1, shape[2] - 1), (shape[0] // 2, shape[1] // 2, shape[2] // 2)]
    while len(bins) < count:
        bins.append((r.randrange(shape[0]), r.randrange(shape[1]), r.randrange(shape[2])))
    return bins


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def plan_fingerprint(desc: str) -> str:
    """Identifies the kernel configuration of a plan (the text fftb200_describe returns)."""
    import hashlib
    return hashlib.sha1(desc.strip().encode()).hexdigest()[:16]


def ncu_traffic(kernel_key: str, fingerprint: str):
    """(dram bytes per launch of the dominant kernel from the committed ncu capture, note).  The capture is only
    valid for the plan it was taken on: profiles/traffic.json records that plan's fingerprint and a different
    plan gets None rather than a stale number."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
    except Exception:
        return None, "no profiles/traffic.json"
    if d.get("_plan_fingerprint") != fingerprint:
        return None, "profiles/traffic.json was captured on another plan (%s != %s): not reported" % (
            d.get("_plan_fingerprint"), fingerprint)
    return d.get(kernel_key), d.get("_source")


def run_1024(args, L, fft, rank, world, dev, barrier, stream):
    """The north-star strong-scaling case beside the headline: 3D C2C fp64 1024^3 at this N (one GPU: the ordinary
    plan; N>1: slabs), timed like the headline, plus 128 sampled output bins against fp64 direct sums."""
    import torch
    import torch.distributed as dist
    n = 1024
    shape = (n, n, n)
    flops = 5.0 * n ** 3 * math.log2(n ** 3)
    steps, warmup = 5, 3
    g = torch.Generator(device=dev).manual_seed(0x5EED0000 + 1024 + rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0l = n // world
    x = torch.view_as_complex(torch.rand(n0l, n, n, 2, dtype=torch.float64, device=dev, generator=g).sub_(0.5))
    dplan = h = None
    if world == 1:
        y = torch.empty_like(x)
        h = L.plan_many(3, list(shape), None, 0, 0, None, 0, 0, L.Z2Z, 1)
        L.set_stream(h, stream.cuda_stream)
        step = lambda: L.execute(h, L.Z2Z, x.data_ptr(), y.data_ptr())
        desc = "single GPU: " + "; ".join(d.split(" threads=")[0].strip() for d in L.describe(h).strip().split("\n"))
    else:
        from regent_fft_arjun_b200 import distributed as D
        dplan = D.SlabFFT3D(shape, fft.complex64, rank=rank, world=world, device=dev, mode=args.exchange)
        dplan.set_input(x)
        step = dplan.execute
        desc = dplan.describe()
    for _ in range(warmup):
        step()
    barrier()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # ---- sampled bins against fp64 direct sums over the same input
    par = sampled_bins_parity(x, y if world == 1 else dplan.out, rank, world, shape, 128, 1024)
    if h is not None:
        L.destroy(h)
    if dplan is not None:
        dplan.destroy()
    pass_model = 3 * 2 * n ** 3 * ELT
    return {"workload": "3D C2C complex64 (fp64) 1024^3 forward out-of-place, strong-scaled over %d GPU(s)" % world,
            "ms": ms, "GFLOP/s": flops / (ms * 1e-3) / 1e9, "steps": steps, "warmup": warmup, "plan": desc,
            "pass_model_GB/s_per_gpu": pass_model / world / ms / 1e6,
            "parity": par}


def run_b200(args, rank: int, world: int, local_rank: int) -> int:
    import numpy as np
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libfft_b200 has no CPU path (use --impl reference for FFTW)")
    all_cpus = os.sched_getaffinity(0)
    numa_note = bind_to_gpu_numa(local_rank)      # before any pinned allocation: first touch decides the node
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    fft = load_package()
    L = fft._lib
    L.lib()
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps, warmup = args.steps, max(3, args.warmup)
    stream = torch.cuda.current_stream()
    g = torch.Generator(device=dev).manual_seed(0x5EED0000 + 4 + rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    if world == 1:
        x = torch.view_as_complex(torch.rand(*SHAPE, 2, dtype=torch.float64, device=dev, generator=g).sub_(0.5))
        y = torch.empty_like(x)
        h = L.plan_many(3, list(SHAPE), None, 0, 0, None, 0, 0, L.Z2Z, 1)
        L.set_stream(h, stream.cuda_stream)
        nl = L.launch_count(h)
        step = lambda: L.execute(h, L.Z2Z, x.data_ptr(), y.data_ptr())
        launches_per_step = nl
        parallelism = "single GPU, %d launches: %s" % (nl, "; ".join(d.split(" threads=")[0].strip() for d in L.describe(h).strip().split("\n")))
        dplan = None
    else:
        from regent_fft_arjun_b200 import distributed as D
        dplan = D.SlabFFT3D(SHAPE, fft.complex64, rank=rank, world=world, device=dev, mode=args.exchange)
        x = torch.view_as_complex(torch.rand(*dplan.local_in_shape, 2, dtype=torch.float64, device=dev, generator=g).sub_(0.5))
        dplan.set_input(x)
        step = dplan.execute
        launches_per_step = dplan.launches_per_step
        parallelism = dplan.describe()
        h = None

    for _ in range(warmup):
        step()
    barrier()
    if h is not None:
        L.set_profiling(h, True)
    clocks = ClockSampler(local_rank).start()
    barrier()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    barrier()
    clk = clocks.stop()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / steps
    value = FLOPS / (ms_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (events recorded inside the timed region above) -------------
    peak, peak_src = measured_peak()
    roof = None
    passes = []
    if h is not None:
        per = [L.launch_ms(h, i) for i in range(nl)]
        L.set_profiling(h, False)
        desc = L.describe(h).strip().split("\n")
        byts = [L.launch_bytes(h, i) for i in range(nl)]
        top = max(range(nl), key=lambda i: per[i])
        for i in range(nl):
            passes.append({"kernel": desc[i].split(" lines=")[0], "ms": round(per[i], 4),
                           "GB/s": round(byts[i] / per[i] / 1e6, 1), "frac_of_peak": round(byts[i] / per[i] / 1e6 / peak, 4)})
        ach = byts[top] / per[top] / 1e6
        fingerprint = plan_fingerprint(L.describe(h))
        traffic, traffic_note = ncu_traffic("z2z_512_launch%d" % top, fingerprint)
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": traffic_note, "plan_fingerprint": fingerprint,
                "kernel": desc[top].split(" lines=")[0],
                "algorithmic_bytes_per_launch": byts[top], "ms_per_launch": per[top], "peak_source": peak_src,
                "whole_transform": {"pass_model_bytes": PASS_MODEL_BYTES, "GB/s": PASS_MODEL_BYTES / ms_step / 1e6,
                                    "frac_of_measured_peak": PASS_MODEL_BYTES / ms_step / 1e6 / peak,
                                    "frac_of_nominal_8TBs": PASS_MODEL_BYTES / ms_step / 1e6 / 8000.0,
                                    "launch_hbm_bytes": sum(byts),
                                    "strict_min_bytes": 2 * N_TOTAL * ELT,
                                    "strict_min_GB/s": 2 * N_TOTAL * ELT / ms_step / 1e6},
                "passes": passes}
    else:
        roof = dplan.roofline(ms_step, peak, peak_src)

    # ---- parity of the timed step's output (outside the timed region): every bin against FFTW on the same input ----
    parity = None
    if not args.no_parity:
        os.sched_setaffinity(0, all_cpus)          # FFTW gets every host thread
        if world == 1:
            parity = parity_vs_fftw(x, y, N_TOTAL)
        else:
            # natural-order input and output on every rank's GPU (2 GiB each), compared on rank 0
            x_full = torch.empty(SHAPE, dtype=torch.complex128, device=dev)
            dist.all_gather_into_tensor(torch.view_as_real(x_full), torch.view_as_real(x))
            o = dplan.out                                                        # [n1/G][n0][n2]
            o_full = torch.empty((SHAPE[1], SHAPE[0], SHAPE[2]), dtype=torch.complex128, device=dev)
            dist.all_gather_into_tensor(torch.view_as_real(o_full), torch.view_as_real(o.contiguous()))
            ok = torch.ones(1, dtype=torch.int32, device=dev)
            if rank == 0:
                y_full = o_full.permute(1, 0, 2).contiguous()
                del o_full
                parity = parity_vs_fftw(x_full, y_full, N_TOTAL)
                if parity is not None:
                    parity["gathered_from_ranks"] = world
                    ok[0] = 1 if parity["ok"] else 0
                del y_full
            else:
                del o_full
            del x_full
            dist.broadcast(ok, src=0)
            if rank != 0:
                parity = {"ok": bool(ok.item())}
        have_fftw = torch.tensor([0 if (rank == 0 and parity is None) else 1], dtype=torch.int32, device=dev)
        if world > 1:
            dist.broadcast(have_fftw, src=0)
        if int(have_fftw.item()) == 0:
            # oracle/_ref did not travel: check 256 sampled bins against fp64 direct sums instead of skipping the check
            parity = sampled_bins_parity(x, y if world == 1 else dplan.out, rank, world, SHAPE, 256, 512)
            parity["note"] = "oracle/_ref/libfftw3_ref.so missing: sampled bins instead of the full FFTW comparison"
        torch.cuda.empty_cache()

    # ---- e2e: same metric through the C ABI with HOST buffers (H2D + transform + D2H in the timed region) ----
    # A caller that streams batches keeps two transforms in flight: step k's device->host copy then overlaps
    # step k+1's host->device copy on the full-duplex host link.  Every step still copies its own input in and
    # its own result out.  N=1: two plans on two streams, each call stream-ordered like cufftExec (the library stages
    # host pointers itself).  N>1: two slots per rank around the collective slab transform (distributed.py).
    os.sched_setaffinity(0, all_cpus)
    numa_note = bind_to_gpu_numa(local_rank)
    e_steps = max(4, min(steps, 8))
    if world == 1:
        npipe = 2
        try:
            hx = [torch.empty(SHAPE, dtype=torch.complex128, pin_memory=True) for _ in range(npipe)]
            hy = [torch.empty(SHAPE, dtype=torch.complex128, pin_memory=True) for _ in range(npipe)]
        except RuntimeError:  # not enough lockable host memory for 4 x 2 GiB: one transform in flight
            npipe = 1
            hx = [torch.empty(SHAPE, dtype=torch.complex128, pin_memory=True)]
            hy = [torch.empty(SHAPE, dtype=torch.complex128, pin_memory=True)]
        for b in hx:
            b.copy_(x)
        pipes = [torch.cuda.Stream(device=dev) for _ in range(npipe)]
        eplans = [L.plan_many(3, list(SHAPE), None, 0, 0, None, 0, 0, L.Z2Z, 1) for _ in range(npipe)]
        for hp, st in zip(eplans, pipes):
            L.set_stream(hp, st.cuda_stream)
        done = [torch.cuda.Event() for _ in range(npipe)]

        def e2e_run(nsteps):
            for st in pipes:
                st.wait_stream(stream)
            for k in range(nsteps):
                i = k % npipe
                L.execute(eplans[i], L.Z2Z, hx[i].data_ptr(), hy[i].data_ptr())
            for st, ev in zip(pipes, done):
                ev.record(st)
                stream.wait_event(ev)

        h2d = d2h = N_TOTAL * ELT
        e2e_path = ("fftb200_exec_z2z(host pinned in, host pinned out), %d plan(s) on %d stream(s) alternating: "
                    "staged H2D, passes on HBM, D2H; consecutive steps' copies overlap" % (npipe, npipe))
    else:
        e2e_run, h2d, d2h, hy = dplan.make_host_pipeline(x, slots=2)
        e2e_path = ("per-rank slab, 2 slots in flight: pinned host -> HBM (copy stream), collective slab transform, "
                    "HBM -> pinned host (copy stream)")
    e2e_run(2)
    barrier()
    e0.record(stream)
    e2e_run(e_steps)
    e1.record(stream)
    barrier()
    e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_ms = float(t.item())
    e_ms /= e_steps
    e2e = {"value": FLOPS / (e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": h2d * world if world > 1 else h2d,
           "d2h_bytes_per_step": d2h * world if world > 1 else d2h, "ms_per_step": e_ms, "steps": e_steps,
           "host_link_GB/s_each_way": h2d / (e_ms * 1e-3) / 1e9, "path": e2e_path, "host_placement": numa_note}
    # the host-buffer calls give the same bits as the resident-input call on the same data
    y_ref = y if world == 1 else dplan.execute(x).clone()
    torch.cuda.synchronize()
    for b in hy:
        assert torch.equal(b.to(dev), y_ref), "e2e result differs from the resident-input result"
    if world == 1:
        for hp in eplans:
            L.destroy(hp)
        del hx
    del hy, y_ref

    # ---- cpu baseline: the reference's FFTW on this box's host cores (rank 0, N=1 only) -----------------
    os.sched_setaffinity(0, all_cpus)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            mean_s, reps, sample, kind, used = time_fftw(3, 1, budget_s=25.0, threads=host_threads())
            cpu = {"value": FLOPS / mean_s / 1e9, "unit": "GFLOP/s", "cores": used, "kind": kind, "sample": sample,
                   "ms_per_transform": mean_s * 1e3}
            cpu["reference_faithful_1thread"] = time_fftw_faithful()
        except Exception as ex:  # the baseline is a report, never a dependency of the product number
            cpu = {"value": None, "unit": "GFLOP/s", "cores": 0, "kind": "reference", "sample": f"failed: {ex}"}

    if h is not None:
        L.destroy(h)
    if dplan is not None:
        dplan.destroy()
    del x
    if world == 1:
        del y
    torch.cuda.empty_cache()

    # ---- the north-star strong-scaling case at this N -----------------------------------------------------
    s1024 = None
    if not args.no_1024:
        try:
            s1024 = run_1024(args, L, fft, rank, world, dev, barrier, stream)
        except Exception as ex:
            s1024 = {"error": f"{type(ex).__name__}: {ex}", "parity": {"ok": False}}

    ok_all = (parity is None or parity.get("ok", False)) and (s1024 is None or s1024["parity"]["ok"])
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "plan": {"parallelism": parallelism,
                     "l2": "input+output 4.3 GB per step >> 126 MB L2 (no flush needed)",
                     "tolerance": "rel-L2 <= 10*log2(N)*eps vs FFTW (parity block; tests/test_gpu_parity.py)"},
            "clocks": clk, "e2e": e2e, "gpu_launches": launches_per_step * steps,
            "roofline": roof, "cpu_baseline": cpu, "parity": parity, "scaling_1024": s1024,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not ok_all:
        sys.stderr.write("bench.py: PARITY FAILED: %s / %s\n" % (parity, s1024 and s1024.get("parity")))
        return 3
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="N>1: fused peer stores or NCCL all-to-all")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the FFTW comparison of the step's output")
    ap.add_argument("--no-1024", action="store_true", help="skip the 1024^3 strong-scaling block")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world == 1 and args.gpus > 1 and args.impl == "b200":
        # launched without torchrun: re-launch ourselves one rank per GPU
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    # stdout carries exactly one JSON line: anything libraries print there (NCCL's version banner, for one) goes to
    # stderr instead, and the line is written to the real stdout at the end
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    try:
        if args.impl == "reference":
            return run_reference(args, rank)
        return run_b200(args, rank, world, local_rank)
    finally:
        real_stdout.flush()


if __name__ == "__main__":
    sys.exit(main())
