// plan_internal.h — data structures and internal entry points shared by the plan builder (plan_builder.cu), the
// executor (plan_exec.cu), the slab plans (slab_plan.cu) and the C ABI (abi.cu).  Not part of the public interface.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fft_b200.h"
#include "generic_kernels.cuh"
#include "mixed_kernel.cuh"
#include "tile_registry.h"

namespace fftb200 {

enum BufSel { BUF_IN = 0, BUF_OUT = 1, BUF_WORK0 = 2, BUF_WORK1 = 3, BUF_BLU = 4 };

struct Launch {
    enum Kind { TILE, GEN_GATHER, GEN_STAGE, GEN_TRUNC, GEN_SCATTER, BLU_PRE, BLU_MUL, BLU_POST, GEN_GATHER_HERM, GEN_SCATTER_REAL, MIXED, R2C_POST, C2R_PRE } kind = TILE;
    // TILE
    const TileKernelInfo *ki = nullptr;
    TileParams tp{};
    int variant = 0;
    // MIXED (mixed_kernel.cuh): ki / tp as for TILE (ki owned by the plan), plus the radix list
    MixedStages mixed{};
    bool mixed_row = false;
    // generic
    GenLayout lay{};
    GenLayout lay2{};  // C2R_PRE: the input lines (lay: the output lines)
    long long total = 0, outer = 0, inner = 0;
    int L = 0, p = 0, Ns = 0, Lc = 0;
    const double2 *gtw = nullptr;
    bool real_in = false;
    // Bluestein (generic path): padded length, chirp table, transformed kernel; tile passes inside a generic plan run
    // in a fixed direction whatever the plan's (dir_override: 0 = the plan's direction, 1 = forward, 2 = backward)
    int M = 0;
    const double2 *chirp = nullptr;
    const void *bhat = nullptr;
    int dir_override = 0;
    // common
    long long in_off = 0, out_off = 0;  // element offsets added to the source / destination base (chunked passes)
    int src = BUF_IN, dst = BUF_OUT;
    unsigned grid = 0;
    unsigned long long algo_bytes = 0;
    std::string desc;
};

struct SlabState;  // slab_plan.cu

struct Plan {
    int device = 0;
    SlabState *slab = nullptr;  // multi-GPU slab plans only
    cudaStream_t stream = nullptr;
    fftb200_type type = FFTB200_Z2Z;
    int prec = 1;       // 0 fp32, 1 fp64
    bool real = false;  // R2C / D2Z
    bool c2r = false;   // C2R / Z2D (complex half spectrum in, reals out)
    bool generic = false;
    bool inplace_ok = false;  // in == out allowed
    std::vector<Launch> launches;
    std::vector<void *> dev_allocs;
    std::vector<std::unique_ptr<TileKernelInfo>> mixed_infos;  // tile shapes of MIXED launches (chosen per plan)
    void *work[2] = {nullptr, nullptr};
    void *blu = nullptr;  // Bluestein convolution buffer (generic plans with a large prime factor)
    size_t blu_bytes = 0;
    size_t work_bytes = 0;
    bool profiling = false;
    // profiling: one row of (launches + 1) events per recorded exec, up to PROF_MAX_EXECS rows
    std::vector<std::vector<cudaEvent_t>> prof_rows;
    size_t prof_used = 0;
    // host-memory staging (zero-copy / pinned / pageable regions, cf. test/test_mapper.cc:45-58):
    // the transform always runs on HBM; host buffers are copied in and out on the plan's stream
    void *stage_in = nullptr, *stage_out = nullptr;
    size_t span_in = 0, span_out = 0;  // bytes from the base pointer to the last touched element + 1
    bool out_dense = true;             // every byte of the output span is written by the transform
    // saved creation arguments (lazy generic fallback for misaligned pointers)
    int rank = 0, batch = 1;
    long long n[3] = {1, 1, 1};
    long long in_stride[4] = {0, 0, 0, 0}, out_stride[4] = {0, 0, 0, 0};  // [batch, d0, d1, d2] elements
    std::unique_ptr<Plan> fallback;
    std::mutex mu;
    Plan() = default;
    Plan(const Plan &) = delete;
    Plan &operator=(const Plan &) = delete;
    ~Plan();  // releases every device resource (plan_exec.cu)
    size_t elt_in() const { return real ? (prec ? 8 : 4) : (prec ? 16 : 8); }
    size_t elt_out() const { return c2r ? (prec ? 8 : 4) : (prec ? 16 : 8); }
};

struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};


// ---- plan_exec.cu ---------------------------------------------------------------------------------------
// one launch of a tile pass; cluster kernels go through cudaLaunchKernelEx with the cluster dimension
cudaError_t launch_tile(const TileKernelInfo *ki, unsigned grid, cudaStream_t st, const TileParams &tp);
void free_plan_resources(Plan *p);
int exec_plan(Plan *P, const void *in, void *out, int direction);
int scale_plan(Plan *P, void *data, double factor);
// handle table: handle = (generation << 32) | (slot + 1); never a raw pointer
// the table owns the plan; ABI calls hold a shared_ptr for their duration (destroy racing exec is safe)
fftb200_handle register_plan(Plan *p);
std::shared_ptr<Plan> lookup_plan(fftb200_handle h);
std::shared_ptr<Plan> unregister_plan(fftb200_handle h);

// ---- plan_builder.cu ------------------------------------------------------------------------------------
int env_int_or(const char *name, int dflt);  // tuning knobs, read at plan creation
bool is_pow2(long long v);
int ilog2ll(long long v);

struct Level {
    long long n, is, os;
    bool keep = false;  // never merged with its neighbours (its index has a meaning of its own: the n2 of a two-pass line)
};

// device tables of one plan under construction
struct Builder {
    Plan *P;
    int err = FFTB200_SUCCESS;
    std::map<std::pair<std::pair<int, long long>, long long>, void *> cache;  // ((kind, n), count) -> device table
    void *upload(const void *host, size_t bytes);
    // w_n^(k*step) for k in [0, count), in the plan's precision or always fp64 (force_double)
    void *table(long long n, long long count, bool force_double, long long step = 1);
    void *stage_tables(long long L, int R);  // per-stage transposed twiddle tables of one tile pipeline
    void *alloc(size_t bytes);
};

// Add one tile pass.  levels: non-axis index levels, fastest first (levels[0] = lines of a tile).
bool add_tile_pass(Builder &B, int variant, int L, long long in_ls, long long out_ls, std::vector<Level> lv, int src, int dst,
                   long long twN /* four-step N */, const char *what);
bool add_tile_pass_with(Builder &B, const TileKernelInfo *ki, int variant, int L, long long in_ls, long long out_ls,
                        std::vector<Level> lv, int src, int dst, long long twN, const char *what);
int create_plan(Plan **out, int rank, const long long *n, int batch, const long long *in_stride, const long long *out_stride,
                fftb200_type type, bool force_generic);

// ---- slab_plan.cu ---------------------------------------------------------------------------------------
void slab_free(Plan *p);
int slab_create(Plan **out, const int *n, fftb200_type type, int rank, int G, int chunks);
int slab_create_2d(Plan **out, const int *n, fftb200_type type, int rank, int G);
int slab_exec_p2p(Plan *P, const void *in, void *out, int inverse);
int slab_exec_pre(Plan *P, const void *in, void *send, int inverse);
int slab_exec_post(Plan *P, const void *recv, void *out, int inverse);
int slab_get_ipc_handle(Plan *P, void *handle64);
int slab_connect_ipc(Plan *P, const void *handles);
int slab_get_area(Plan *P, void **area, unsigned long long *bytes);
int slab_connect_ptrs(Plan *P, void *const *areas);
int slab_set_timing(Plan *P, int on);
int slab_get_phase_ms(Plan *P, float *ms);
int slab_launches_per_exec(const Plan *P);

}  // namespace fftb200
