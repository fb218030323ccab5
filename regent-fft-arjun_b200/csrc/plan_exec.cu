// plan_exec.cu — runs a plan's launches (host-memory staging, per-launch profiling events), owns the handle table.
// No CPU fallback exists: if a kernel cannot be launched the call returns an error code.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "plan_internal.h"

namespace fftb200 {

// one launch of a tile pass; cluster kernels go through cudaLaunchKernelEx with the cluster dimension
// FFTB200_DEBUG=1: say which CUDA call failed (stderr); the ABI itself only returns codes
static cudaError_t report(cudaError_t e, const TileKernelInfo *ki, unsigned grid) {
    if (e != cudaSuccess && getenv("FFTB200_DEBUG"))
        fprintf(stderr, "libfft_b200: launch of tile kernel L=%d R=%d W=%d cluster=%d grid=%u threads=%d smem=%d failed: %s\n", ki->L, ki->R,
                ki->W, ki->cluster, grid, ki->threads, ki->smem_bytes, cudaGetErrorString(e));
    return e;
}

cudaError_t launch_tile(const TileKernelInfo *ki, unsigned grid, cudaStream_t st, const TileParams &tp) {
    if (ki->cluster <= 1) {
        ki->fn<<<grid, ki->threads, ki->smem_bytes, st>>>(tp);
        return report(cudaGetLastError(), ki, grid);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid / ki->cluster * ki->cluster, 1, 1);
    cfg.blockDim = dim3(ki->threads, 1, 1);
    cfg.dynamicSmemBytes = (size_t)ki->smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)ki->cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, ki->fn, tp);
    return report(e != cudaSuccess ? e : cudaGetLastError(), ki, grid);
}

void free_plan_resources(Plan *p) {
    DeviceGuard g(p->device);
    slab_free(p);
    for (void *d : p->dev_allocs) cudaFree(d);
    p->dev_allocs.clear();
    for (auto &row : p->prof_rows)
        for (cudaEvent_t e : row) cudaEventDestroy(e);
    p->prof_rows.clear();
    p->prof_used = 0;
    if (p->stage_in) cudaFree(p->stage_in);
    if (p->stage_out) cudaFree(p->stage_out);
    p->stage_in = p->stage_out = nullptr;
    p->fallback.reset();
}

// ------------------------------------------------------------------------------------------
// handle table: handle = (generation << 32) | (slot + 1); never a raw pointer, so a stale or
// zero-filled plan region (src/fft.rg:523-531) cannot crash the library
// ------------------------------------------------------------------------------------------
// Every ABI call holds a shared_ptr for its duration, so fftb200_destroy racing an exec / describe / timing
// query on another thread (Legion runs destroy_plan from a CPU processor, src/fft.rg:624-645) cannot free the
// plan under it: the resources go when the last holder returns (~Plan).
static std::mutex g_mu;
static std::vector<std::pair<unsigned, std::shared_ptr<Plan>>> g_slots;  // (generation, plan)

Plan::~Plan() { free_plan_resources(this); }

fftb200_handle register_plan(Plan *p) {
    std::shared_ptr<Plan> sp(p);
    std::lock_guard<std::mutex> lk(g_mu);
    for (size_t i = 0; i < g_slots.size(); ++i)
        if (!g_slots[i].second) {
            g_slots[i].first++;
            g_slots[i].second = std::move(sp);
            return ((fftb200_handle)g_slots[i].first << 32) | (fftb200_handle)(i + 1);
        }
    g_slots.push_back({1u, std::move(sp)});
    return ((fftb200_handle)1 << 32) | (fftb200_handle)g_slots.size();
}

std::shared_ptr<Plan> lookup_plan(fftb200_handle h) {
    std::lock_guard<std::mutex> lk(g_mu);
    const size_t slot = (size_t)(h & 0xffffffffull);
    const unsigned gen = (unsigned)(h >> 32);
    if (slot == 0 || slot > g_slots.size()) return nullptr;
    if (g_slots[slot - 1].first != gen) return nullptr;
    return g_slots[slot - 1].second;
}

std::shared_ptr<Plan> unregister_plan(fftb200_handle h) {
    std::lock_guard<std::mutex> lk(g_mu);
    const size_t slot = (size_t)(h & 0xffffffffull);
    const unsigned gen = (unsigned)(h >> 32);
    if (slot == 0 || slot > g_slots.size()) return nullptr;
    if (g_slots[slot - 1].first != gen) return nullptr;
    std::shared_ptr<Plan> p = std::move(g_slots[slot - 1].second);
    g_slots[slot - 1].second.reset();
    return p;
}

// ------------------------------------------------------------------------------------------
// execution
// ------------------------------------------------------------------------------------------
template <typename T> static cudaError_t launch_generic(const Launch &ln, const void *src, void *dst, int inverse,
                                                        cudaStream_t st) {
    using C = cplx<T>;
    switch (ln.kind) {
        case Launch::GEN_GATHER:
            if (ln.real_in)
                gen_gather_kernel<T, true><<<ln.grid, 256, 0, st>>>(src, (C *)dst, ln.lay, ln.total, 0);
            else
                gen_gather_kernel<T, false><<<ln.grid, 256, 0, st>>>(src, (C *)dst, ln.lay, ln.total, inverse);
            break;
        case Launch::GEN_STAGE:
            gen_stage_kernel<T><<<ln.grid, 256, 0, st>>>((const C *)src, (C *)dst, ln.gtw, ln.outer, ln.L, ln.inner,
                                                         ln.p, ln.Ns);
            break;
        case Launch::GEN_TRUNC:
            gen_truncate_kernel<T><<<ln.grid, 256, 0, st>>>((const C *)src, (C *)dst, ln.outer, ln.L, ln.Lc);
            break;
        case Launch::GEN_SCATTER:
            gen_scatter_kernel<T><<<ln.grid, 256, 0, st>>>((const C *)src, (C *)dst, ln.lay, ln.total, inverse);
            break;
        case Launch::GEN_GATHER_HERM:
            gen_gather_herm_kernel<T><<<ln.grid, 256, 0, st>>>((const C *)src, (C *)dst, ln.lay, ln.total);
            break;
        case Launch::GEN_SCATTER_REAL:
            gen_scatter_real_kernel<T><<<ln.grid, 256, 0, st>>>((const C *)src, (T *)dst, ln.lay, ln.total);
            break;
        case Launch::R2C_POST:
            gen_r2c_post_kernel<T><<<ln.grid, 256, 0, st>>>((C *)dst, ln.lay, ln.total, ln.L, (const C *)ln.bhat);
            break;
        case Launch::C2R_PRE:
            gen_c2r_pre_kernel<T><<<ln.grid, 256, 0, st>>>((const C *)src, (C *)dst, ln.lay2, ln.lay, ln.total, ln.L, (const C *)ln.bhat);
            break;
        case Launch::BLU_PRE:
            gen_blu_pre_kernel<T><<<ln.grid, 256, 0, st>>>((const C *)src, (C *)dst, ln.chirp, ln.outer, ln.L, ln.M, ln.inner);
            break;
        case Launch::BLU_MUL:
            gen_blu_mul_kernel<T><<<ln.grid, 256, 0, st>>>((C *)dst, (const C *)ln.bhat, ln.outer, ln.M, ln.inner);
            break;
        case Launch::BLU_POST:
            gen_blu_post_kernel<T><<<ln.grid, 256, 0, st>>>((const C *)src, (C *)dst, ln.chirp, ln.outer, ln.L, ln.M, ln.inner,
                                                            1.0 / (double)ln.M);
            break;
        default: break;
    }
    return cudaGetLastError();
}


static int exec_fallback(Plan *P, const void *in, void *out, int direction);

// true when the kernels cannot (or should not) read the pointer directly: pageable host memory is
// not device-accessible at all, pinned / zero-copy host memory would be re-streamed over PCIe by
// every pass
static bool is_host_memory(const void *ptr) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeUnregistered;
}

static const size_t PROF_MAX_EXECS = 256;

static int run_launches(Plan *P, const void *in, void *out, int inverse) {
    const size_t nl = P->launches.size();
    std::vector<cudaEvent_t> *row = nullptr;
    if (P->profiling && P->prof_used < PROF_MAX_EXECS) {
        if (P->prof_rows.size() <= P->prof_used) {
            std::vector<cudaEvent_t> r(nl + 1, nullptr);
            for (size_t i = 0; i <= nl; ++i)
                if (cudaEventCreate(&r[i]) != cudaSuccess) return FFTB200_INTERNAL_ERROR;
            P->prof_rows.push_back(r);
        }
        row = &P->prof_rows[P->prof_used++];
        cudaEventRecord((*row)[0], P->stream);
    }
    for (size_t i = 0; i < nl; ++i) {
        const Launch &ln = P->launches[i];
        const void *bufs_src[5] = {in, out, P->work[0], P->work[1], P->blu};
        void *bufs_dst[5] = {nullptr, out, P->work[0], P->work[1], P->blu};
        const void *src = bufs_src[ln.src];
        void *dst = bufs_dst[ln.dst];
        cudaError_t ce;
        if (ln.kind == Launch::TILE) {
            TileParams tp = ln.tp;
            const size_t elt = P->prec ? 16 : 8;  // launch offsets are in complex elements
            tp.in = (const char *)src + (size_t)ln.in_off * elt;
            tp.out = (char *)dst + (size_t)ln.out_off * elt;
            tp.inverse = ln.dir_override ? (ln.dir_override == 2) : inverse;
            ce = launch_tile(ln.ki, ln.grid, P->stream, tp);
        } else if (ln.kind == Launch::MIXED) {
            TileParams tp = ln.tp;
            const size_t elt = P->prec ? 16 : 8;
            tp.in = (const char *)src + (size_t)ln.in_off * elt;
            tp.out = (char *)dst + (size_t)ln.out_off * elt;
            tp.inverse = inverse;
            reinterpret_cast<MixedKernelFn>(ln.ki->fn)<<<ln.grid, ln.ki->threads, ln.ki->smem_bytes, P->stream>>>(tp, ln.mixed);
            ce = report(cudaGetLastError(), ln.ki, ln.grid);
        } else {
            ce = P->prec ? launch_generic<double>(ln, src, dst, inverse, P->stream)
                         : launch_generic<float>(ln, src, dst, inverse, P->stream);
        }
        if (ce != cudaSuccess) return FFTB200_EXEC_FAILED;
        if (row) cudaEventRecord((*row)[i + 1], P->stream);
    }
    return FFTB200_SUCCESS;
}

int exec_plan(Plan *P, const void *in, void *out, int direction) {
    if (!in || !out) return FFTB200_INVALID_VALUE;
    if (direction != FFTB200_FORWARD && direction != FFTB200_INVERSE) return FFTB200_INVALID_VALUE;
    if (P->real && direction != FFTB200_FORWARD) return FFTB200_INVALID_VALUE;
    if (P->c2r && direction != FFTB200_INVERSE) return FFTB200_INVALID_VALUE;
    if (in == out && !P->inplace_ok) {
        // (mixed-radix plans replaced the generic path for these sizes, which worked in place for any layout: keep that)
        if (!P->mixed_infos.empty()) return exec_fallback(P, in, out, direction);
        return FFTB200_INVALID_VALUE;
    }
    const bool host_in = is_host_memory(in), host_out = is_host_memory(out);
    if (!P->generic && !(host_in && host_out)) {
        const size_t a_in = P->real ? 2 * P->elt_in() : P->elt_in();
        const size_t a_out = P->c2r ? 2 * P->elt_out() : P->elt_out();
        if ((!host_in && ((uintptr_t)in % a_in)) || (!host_out && ((uintptr_t)out % a_out))) {
            return exec_fallback(P, in, out, direction);
        }
    }
    DeviceGuard g(P->device);
    std::lock_guard<std::mutex> lk(P->mu);
    const int inverse = (direction == FFTB200_INVERSE) ? 1 : 0;
    const void *din = in;
    void *dout = out;
    if (host_in) {
        if (!P->stage_in && cudaMalloc(&P->stage_in, P->span_in) != cudaSuccess) { cudaGetLastError(); return FFTB200_ALLOC_FAILED; }
        if (cudaMemcpyAsync(P->stage_in, in, P->span_in, cudaMemcpyHostToDevice, P->stream) != cudaSuccess) {
            cudaGetLastError();
            return FFTB200_EXEC_FAILED;
        }
        din = P->stage_in;
    }
    if (host_out) {
        if (!P->stage_out && cudaMalloc(&P->stage_out, P->span_out) != cudaSuccess) { cudaGetLastError(); return FFTB200_ALLOC_FAILED; }
        // padding between rows / batches must survive the round trip
        if (!P->out_dense &&
            cudaMemcpyAsync(P->stage_out, out, P->span_out, cudaMemcpyHostToDevice, P->stream) != cudaSuccess) {
            cudaGetLastError();
            return FFTB200_EXEC_FAILED;
        }
        dout = P->stage_out;
    }
    const int rc = run_launches(P, din, dout, inverse);
    if (rc != FFTB200_SUCCESS) return rc;
    if (host_out && cudaMemcpyAsync(out, P->stage_out, P->span_out, cudaMemcpyDeviceToHost, P->stream) != cudaSuccess) {
        cudaGetLastError();
        return FFTB200_EXEC_FAILED;
    }
    return FFTB200_SUCCESS;
}

// normalisation helper: scales the plan's output layout in place (include/fft_b200.h: fftb200_scale)
int scale_plan(Plan *P, void *data, double factor) {
    if (!data) return FFTB200_INVALID_VALUE;
    if (P->slab) return FFTB200_UNSUPPORTED;
    if (is_host_memory(data)) return FFTB200_INVALID_VALUE;  // device arrays only
    DeviceGuard g(P->device);
    std::lock_guard<std::mutex> lk(P->mu);
    GenLayout lay{};
    lay.nd = P->rank + 1;
    lay.n[0] = P->batch;
    lay.stride[0] = P->out_stride[0];
    long long total = P->batch, n_total = 1;
    for (int d = 0; d < P->rank; ++d) {
        const long long no = (P->real && d == P->rank - 1) ? P->n[d] / 2 + 1 : P->n[d];
        lay.n[d + 1] = no;
        lay.stride[d + 1] = P->out_stride[d + 1];
        total *= no;
        n_total *= P->n[d];
    }
    const double f = factor != 0.0 ? factor : 1.0 / (double)n_total;
    long long gl = (total + 255) / 256;
    if (gl > 148ll * 64) gl = 148ll * 64;
    const unsigned grid = (unsigned)(gl < 1 ? 1 : gl);
    if (P->prec) {
        if (P->c2r) gen_scale_kernel<double, 1><<<grid, 256, 0, P->stream>>>((double *)data, lay, total, f);
        else gen_scale_kernel<double, 2><<<grid, 256, 0, P->stream>>>((double *)data, lay, total, f);
    } else {
        if (P->c2r) gen_scale_kernel<float, 1><<<grid, 256, 0, P->stream>>>((float *)data, lay, total, (float)f);
        else gen_scale_kernel<float, 2><<<grid, 256, 0, P->stream>>>((float *)data, lay, total, (float)f);
    }
    return cudaGetLastError() == cudaSuccess ? FFTB200_SUCCESS : FFTB200_EXEC_FAILED;
}

static int exec_fallback(Plan *P, const void *in, void *out, int direction) {
    Plan *fb = nullptr;
    cudaStream_t st = nullptr;
    {
        std::lock_guard<std::mutex> lk(P->mu);
        if (!P->fallback) {
            DeviceGuard g(P->device);
            Plan *made = nullptr;
            const int rc = create_plan(&made, P->rank, P->n, P->batch, P->in_stride, P->out_stride, P->type, true);
            if (rc != FFTB200_SUCCESS) return rc;
            P->fallback.reset(made);
        }
        fb = P->fallback.get();
        st = P->stream;
    }
    {
        std::lock_guard<std::mutex> lk(fb->mu);  // the fallback's stream is read under its own lock (exec_plan)
        fb->stream = st;
    }
    return exec_plan(fb, in, out, direction);
}

}  // namespace fftb200
