// abi.cu — the C ABI of libfft_b200 (include/fft_b200.h): the layer that replaces cuFFT behind Regent-FFT's GPU
// branch (reference src/fft.rg:233-242, 389-398, 571-580, 638).  Nothing throws or aborts across this boundary.
#include <cstdio>
#include <cstring>

#include "plan_internal.h"

using namespace fftb200;

extern "C" {

int fftb200_plan_many(fftb200_handle *plan, int rank, const int *n, const int *inembed, int istride, int idist,
                      const int *onembed, int ostride, int odist, fftb200_type type, int batch) {
    if (!plan) return FFTB200_INVALID_VALUE;
    *plan = 0;
    if (!n || rank < 1 || rank > 3 || batch < 1) return FFTB200_INVALID_VALUE;
    if (type != FFTB200_R2C && type != FFTB200_C2C && type != FFTB200_D2Z && type != FFTB200_Z2Z && type != FFTB200_C2R &&
        type != FFTB200_Z2D)
        return FFTB200_INVALID_TYPE;
    const bool real = (type == FFTB200_R2C || type == FFTB200_D2Z);
    const bool c2r = (type == FFTB200_C2R || type == FFTB200_Z2D);
    long long nn[3] = {1, 1, 1}, ie[3], oe[3];
    for (int d = 0; d < rank; ++d) {
        if (n[d] < 1) return FFTB200_INVALID_SIZE;
        nn[d] = n[d];
    }
    long long is = 1, os = 1, id, od;
    if (!inembed || !onembed) {
        // cuFFT basic layout: packed, strides/dists ignored
        for (int d = 0; d < rank; ++d) { ie[d] = nn[d]; oe[d] = nn[d]; }
        if (real) oe[rank - 1] = nn[rank - 1] / 2 + 1;
        if (c2r) ie[rank - 1] = nn[rank - 1] / 2 + 1;
        id = od = 1;
        for (int d = 0; d < rank; ++d) { id *= ie[d]; od *= oe[d]; }
    } else {
        if (istride < 1 || ostride < 1) return FFTB200_INVALID_VALUE;
        for (int d = 0; d < rank; ++d) {
            ie[d] = inembed[d];
            oe[d] = onembed[d];
            const long long need_o = (real && d == rank - 1) ? nn[d] / 2 + 1 : nn[d];
            const long long need_i = (c2r && d == rank - 1) ? nn[d] / 2 + 1 : nn[d];
            if (d > 0 && (ie[d] < need_i || oe[d] < need_o)) return FFTB200_INVALID_VALUE;
        }
        is = istride; os = ostride; id = idist; od = odist;
        if (batch > 1 && (id < 1 || od < 1)) return FFTB200_INVALID_VALUE;
    }
    long long in_stride[4], out_stride[4];  // [batch, d0.., d_last]
    in_stride[rank] = is;
    out_stride[rank] = os;
    for (int d = rank - 1; d >= 1; --d) {
        in_stride[d] = in_stride[d + 1] * ie[d];
        out_stride[d] = out_stride[d + 1] * oe[d];
    }
    in_stride[0] = id;
    out_stride[0] = od;
    Plan *P = nullptr;
    const int rc = create_plan(&P, rank, nn, batch, in_stride, out_stride, type, false);
    if (rc != FFTB200_SUCCESS) return rc;
    *plan = register_plan(P);
    return FFTB200_SUCCESS;
}

int fftb200_set_stream(fftb200_handle plan, void *cuda_stream) {
    const std::shared_ptr<Plan> hold = lookup_plan(plan);
    Plan *P = hold.get();
    if (!P) return FFTB200_INVALID_PLAN;
    std::lock_guard<std::mutex> lk(P->mu);
    P->stream = (cudaStream_t)cuda_stream;
    return FFTB200_SUCCESS;
}

static int exec_typed(fftb200_handle plan, const void *in, void *out, int direction, fftb200_type want) {
    const std::shared_ptr<Plan> hold = lookup_plan(plan);
    Plan *P = hold.get();
    if (!P) return FFTB200_INVALID_PLAN;
    if (P->type != want) return FFTB200_INVALID_TYPE;
    if (P->slab) return FFTB200_INVALID_PLAN;  // slab plans run through fftb200_slab_exec*
    return exec_plan(P, in, out, direction);
}

int fftb200_exec_c2c(fftb200_handle plan, const void *in, void *out, int direction) {
    return exec_typed(plan, in, out, direction, FFTB200_C2C);
}
int fftb200_exec_z2z(fftb200_handle plan, const void *in, void *out, int direction) {
    return exec_typed(plan, in, out, direction, FFTB200_Z2Z);
}
int fftb200_exec_r2c(fftb200_handle plan, const void *in, void *out) {
    return exec_typed(plan, in, out, FFTB200_FORWARD, FFTB200_R2C);
}
int fftb200_exec_d2z(fftb200_handle plan, const void *in, void *out) {
    return exec_typed(plan, in, out, FFTB200_FORWARD, FFTB200_D2Z);
}
int fftb200_exec_c2r(fftb200_handle plan, const void *in, void *out) {
    return exec_typed(plan, in, out, FFTB200_INVERSE, FFTB200_C2R);
}
int fftb200_exec_z2d(fftb200_handle plan, const void *in, void *out) {
    return exec_typed(plan, in, out, FFTB200_INVERSE, FFTB200_Z2D);
}

int fftb200_scale(fftb200_handle plan, void *data, double factor) {
    const std::shared_ptr<Plan> hold = lookup_plan(plan);
    Plan *P = hold.get();
    if (!P) return FFTB200_INVALID_PLAN;
    return scale_plan(P, data, factor);
}

int fftb200_destroy(fftb200_handle plan) {
    if (plan == 0) return FFTB200_SUCCESS;  // zero-filled plan regions (src/fft.rg:523-531)
    const std::shared_ptr<Plan> hold = unregister_plan(plan);
    if (!hold) return FFTB200_INVALID_PLAN;
    return FFTB200_SUCCESS;  // resources are released when the last call still using the plan returns (~Plan)
}

int fftb200_get_work_size(fftb200_handle plan, unsigned long long *bytes) {
    const std::shared_ptr<Plan> hold = lookup_plan(plan);
    Plan *P = hold.get();
    if (!P || !bytes) return P ? FFTB200_INVALID_VALUE : FFTB200_INVALID_PLAN;
    *bytes = (unsigned long long)P->work_bytes * ((P->work[0] ? 1 : 0) + (P->work[1] ? 1 : 0)) + P->blu_bytes;
    return FFTB200_SUCCESS;
}

int fftb200_get_launch_count(fftb200_handle plan, int *launches) {
    const std::shared_ptr<Plan> hold = lookup_plan(plan);
    Plan *P = hold.get();
    if (!P || !launches) return P ? FFTB200_INVALID_VALUE : FFTB200_INVALID_PLAN;
    *launches = (int)P->launches.size();
    if (P->slab) *launches = slab_launches_per_exec(P);
    return FFTB200_SUCCESS;
}

int fftb200_describe(fftb200_handle plan, char *buf, int buflen) {
    const std::shared_ptr<Plan> hold = lookup_plan(plan);
    Plan *P = hold.get();
    if (!P || !buf || buflen < 1) return P ? FFTB200_INVALID_VALUE : FFTB200_INVALID_PLAN;
    std::string s;
    for (const Launch &l : P->launches) { s += l.desc; s += "\n"; }
    snprintf(buf, (size_t)buflen, "%s", s.c_str());
    return FFTB200_SUCCESS;
}

int fftb200_get_launch_bytes(fftb200_handle plan, int i, unsigned long long *bytes) {
    const std::shared_ptr<Plan> hold = lookup_plan(plan);
    Plan *P = hold.get();
    if (!P || !bytes) return P ? FFTB200_INVALID_VALUE : FFTB200_INVALID_PLAN;
    if (i < 0 || i >= (int)P->launches.size()) return FFTB200_INVALID_VALUE;
    *bytes = P->launches[i].algo_bytes;
    return FFTB200_SUCCESS;
}

int fftb200_set_profiling(fftb200_handle plan, int on) {
    const std::shared_ptr<Plan> hold = lookup_plan(plan);
    Plan *P = hold.get();
    if (!P) return FFTB200_INVALID_PLAN;
    std::lock_guard<std::mutex> lk(P->mu);
    P->profiling = on != 0;
    if (on) P->prof_used = 0;  // start a new series; event rows are reused
    return FFTB200_SUCCESS;
}

// mean duration of launch i over the execs recorded since profiling was switched on
int fftb200_get_launch_ms(fftb200_handle plan, int i, float *ms) {
    const std::shared_ptr<Plan> hold = lookup_plan(plan);
    Plan *P = hold.get();
    if (!P || !ms) return P ? FFTB200_INVALID_VALUE : FFTB200_INVALID_PLAN;
    std::lock_guard<std::mutex> lk(P->mu);
    if (i < 0 || i >= (int)P->launches.size() || P->prof_used == 0) return FFTB200_INVALID_VALUE;
    DeviceGuard g(P->device);
    double sum = 0;
    for (size_t r = 0; r < P->prof_used; ++r) {
        float t = 0.f;
        if (cudaEventSynchronize(P->prof_rows[r][i + 1]) != cudaSuccess) { cudaGetLastError(); return FFTB200_EXEC_FAILED; }
        if (cudaEventElapsedTime(&t, P->prof_rows[r][i], P->prof_rows[r][i + 1]) != cudaSuccess) { cudaGetLastError(); return FFTB200_EXEC_FAILED; }
        sum += t;
    }
    *ms = (float)(sum / (double)P->prof_used);
    return FFTB200_SUCCESS;
}

// ---- multi-GPU slab transforms ------------------------------------------------------------
static std::shared_ptr<Plan> lookup_slab(fftb200_handle plan) {
    std::shared_ptr<Plan> p = lookup_plan(plan);
    return (p && p->slab) ? p : nullptr;
}

static int slab_direction(Plan *P, int direction, int *inverse) {
    if (direction != FFTB200_FORWARD && direction != FFTB200_INVERSE) return FFTB200_INVALID_VALUE;
    if (P->real && direction != FFTB200_FORWARD) return FFTB200_INVALID_VALUE;
    *inverse = direction == FFTB200_INVERSE;
    return FFTB200_SUCCESS;
}

int fftb200_slab_plan(fftb200_handle *plan, const int *n, fftb200_type type, int rank, int nranks, int chunks) {
    if (!plan) return FFTB200_INVALID_VALUE;
    *plan = 0;
    if (!n) return FFTB200_INVALID_VALUE;
    if (type != FFTB200_R2C && type != FFTB200_C2C && type != FFTB200_D2Z && type != FFTB200_Z2Z)
        return FFTB200_INVALID_TYPE;
    Plan *P = nullptr;
    const int rc = slab_create(&P, n, type, rank, nranks, chunks);
    if (rc != FFTB200_SUCCESS) return rc;
    *plan = register_plan(P);
    return FFTB200_SUCCESS;
}

int fftb200_slab_plan_2d(fftb200_handle *plan, const int *n, fftb200_type type, int rank, int nranks) {
    if (!plan) return FFTB200_INVALID_VALUE;
    *plan = 0;
    if (!n) return FFTB200_INVALID_VALUE;
    if (type != FFTB200_C2C && type != FFTB200_Z2Z) return FFTB200_INVALID_TYPE;
    Plan *P = nullptr;
    const int rc = slab_create_2d(&P, n, type, rank, nranks);
    if (rc != FFTB200_SUCCESS) return rc;
    *plan = register_plan(P);
    return FFTB200_SUCCESS;
}

int fftb200_slab_get_ipc_handle(fftb200_handle plan, void *handle64) {
    const std::shared_ptr<Plan> hold = lookup_slab(plan);
    Plan *P = hold.get();
    if (!P || !handle64) return P ? FFTB200_INVALID_VALUE : FFTB200_INVALID_PLAN;
    return slab_get_ipc_handle(P, handle64);
}

int fftb200_slab_connect_ipc(fftb200_handle plan, const void *handles) {
    const std::shared_ptr<Plan> hold = lookup_slab(plan);
    Plan *P = hold.get();
    if (!P || !handles) return P ? FFTB200_INVALID_VALUE : FFTB200_INVALID_PLAN;
    return slab_connect_ipc(P, handles);
}

int fftb200_slab_get_area(fftb200_handle plan, void **area, unsigned long long *bytes) {
    const std::shared_ptr<Plan> hold = lookup_slab(plan);
    Plan *P = hold.get();
    if (!P || !area) return P ? FFTB200_INVALID_VALUE : FFTB200_INVALID_PLAN;
    return slab_get_area(P, area, bytes);
}

int fftb200_slab_connect_ptrs(fftb200_handle plan, void *const *areas) {
    const std::shared_ptr<Plan> hold = lookup_slab(plan);
    Plan *P = hold.get();
    if (!P || !areas) return P ? FFTB200_INVALID_VALUE : FFTB200_INVALID_PLAN;
    return slab_connect_ptrs(P, areas);
}

int fftb200_slab_exec(fftb200_handle plan, const void *in, void *out, int direction) {
    const std::shared_ptr<Plan> hold = lookup_slab(plan);
    Plan *P = hold.get();
    if (!P) return FFTB200_INVALID_PLAN;
    if (!in || !out) return FFTB200_INVALID_VALUE;
    int inverse = 0;
    const int rc = slab_direction(P, direction, &inverse);
    return rc ? rc : slab_exec_p2p(P, in, out, inverse);
}

int fftb200_slab_exec_pre(fftb200_handle plan, const void *in, void *send, int direction) {
    const std::shared_ptr<Plan> hold = lookup_slab(plan);
    Plan *P = hold.get();
    if (!P) return FFTB200_INVALID_PLAN;
    if (!in || !send) return FFTB200_INVALID_VALUE;
    int inverse = 0;
    const int rc = slab_direction(P, direction, &inverse);
    if (P->rank == 2) return FFTB200_UNSUPPORTED;  // 2-D slabs run the fused exchange only
    return rc ? rc : slab_exec_pre(P, in, send, inverse);
}

int fftb200_slab_exec_post(fftb200_handle plan, const void *recv, void *out, int direction) {
    const std::shared_ptr<Plan> hold = lookup_slab(plan);
    Plan *P = hold.get();
    if (!P) return FFTB200_INVALID_PLAN;
    if (!recv || !out) return FFTB200_INVALID_VALUE;
    int inverse = 0;
    const int rc = slab_direction(P, direction, &inverse);
    if (P->rank == 2) return FFTB200_UNSUPPORTED;
    return rc ? rc : slab_exec_post(P, recv, out, inverse);
}

int fftb200_slab_set_timing(fftb200_handle plan, int on) {
    const std::shared_ptr<Plan> hold = lookup_slab(plan);
    Plan *P = hold.get();
    if (!P) return FFTB200_INVALID_PLAN;
    return slab_set_timing(P, on);
}

int fftb200_slab_get_phase_ms(fftb200_handle plan, float *ms) {
    const std::shared_ptr<Plan> hold = lookup_slab(plan);
    Plan *P = hold.get();
    if (!P || !ms) return P ? FFTB200_INVALID_VALUE : FFTB200_INVALID_PLAN;
    return slab_get_phase_ms(P, ms);
}

const char *fftb200_strerror(int code) {
    switch (code) {
        case FFTB200_SUCCESS: return "success";
        case FFTB200_INVALID_PLAN: return "invalid plan handle";
        case FFTB200_ALLOC_FAILED: return "device allocation failed";
        case FFTB200_INVALID_TYPE: return "transform type does not match the plan";
        case FFTB200_INVALID_VALUE: return "invalid argument";
        case FFTB200_INTERNAL_ERROR: return "internal error";
        case FFTB200_EXEC_FAILED: return "kernel launch failed";
        case FFTB200_SETUP_FAILED: return "CUDA setup failed (no device / context?)";
        case FFTB200_INVALID_SIZE: return "invalid transform size";
        case FFTB200_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown error";
    }
}

int fftb200_version(void) { return 100; }

}  // extern "C"

