// slab_plan.cu — multi-GPU slab decomposition of 3-D (and 2-D complex) transforms.
//
// No reference counterpart in src/fft.rg (its "distrib" path runs independent shard FFTs,
// src/fft.rg:513-537; README.md:117-119 lists a distributed transform as future work).  The scheme is
// the one the vendored FFTW-MPI uses (fftw-3.3.8/mpi/dft-rank-geq2.c:40-59: local FFTs over the
// non-distributed dims, global transpose, local FFTs over the formerly distributed dim;
// mpi/transpose-alltoall.c:49-100: pack + all-to-all + unpack; equal blocks, mpi/block.c:39-50), with
// the output left distributed over dim 1 exactly like FFTW_MPI_TRANSPOSED_OUT (doc/mpi.texi:443-466).
//
//   rank r holds   in  [n0/G][n1][n2]      (slab r of dim 0, row-major)
//   pass 1         x-axis FFT (or fused r2c)              in  -> tmp [n0/G][n1][n2c]     HBM
//                  (tmp and recv rows are padded to whole 128-byte lines internally)
//   pass 2         y-axis FFT whose STORE is the exchange: output line index k1 belongs to rank
//                  k1 / (n1/G); it is written straight into that rank's receive buffer
//                  recv_d [n1/G][n0][n2c] at [k1 % (n1/G)][r*n0/G + p][i]
//                    - p2p mode: recv_d is peer memory mapped over NVLink (cudaIpc / peer access); no
//                      pack, no NCCL, no unpack: the all-to-all IS the FFT pass's store
//                    - staged mode: recv_d are the G blocks of a local send buffer; the host runs any
//                      all-to-all (NCCL) on it
//   pass 3         z-axis FFT                             recv -> out [n1/G][n0][n2c]    HBM
// Complex transforms run the fused path in the order y (+exchange), x, z instead (see slab_create): the x
// axis is local on both sides of the exchange, and after it it can overlap the transfer chunk by chunk.
//
// p2p mode pipelines passes 2 and 3 over J chunks of the contiguous index i: chunk j of pass 3 starts
// as soon as every rank has signalled chunk j of pass 2 (flags written into each peer's exchange area),
// so the z-axis pass hides under the NVLink transfer of the following chunks.

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "plan_internal.h"
#include "slab_fused_kernel.cuh"

namespace fftb200 {

struct SlabState {
    int rank = 0, G = 1, J = 1;
    long long n0 = 0, n1 = 0, n2 = 0, n2c = 0, n2p = 0, n0l = 0, n1l = 0;
    // internal slabs use a row pitch n2p = n2c rounded up to a whole number of 128-byte lines, so that the
    // peer stores of an R2C exchange (n2c = n2/2+1 columns) stay line-aligned on NVLink
    void *tmp = nullptr;   // [n0l][n1][n2p]
    void *area = nullptr;  // exchange area: recv [n1l][n0][n2p], then flags
    size_t recv_bytes = 0, area_bytes = 0, flags_off = 0;
    void *peer_area[MAX_PEERS] = {};
    bool peer_mapped[MAX_PEERS] = {};
    bool connected = false;
    unsigned long long epoch = 0;
    int *err_host = nullptr;  // mapped host word: a wait kernel that gave up sets it; every later exec reports it
    cudaStream_t aux = nullptr;
    std::vector<cudaEvent_t> ev_chunk;
    cudaEvent_t ev_done = nullptr, ev_t[5] = {};
    bool timing = false;
    int l_pass1 = -1, l_pre2 = -1, l_post3 = -1;
    std::vector<int> l_pass2, l_pass3;
    long long chunk_w = 0;
    // complex transforms: y-first pipeline (see slab_exec_p2p): plane chunks
    int Jp = 0;
    std::vector<int> l_y, l_x;
    int l_z = -1;
    // 2-D slabs: rank r holds in [n0/G][n1], gets out [n1/G][n0] (transposed-out): row FFT whose store is the
    // global transpose into the peers' receive slabs, hand-shake, row FFT over the received (now contiguous) n0
    bool two_d = false;
    int l_2a = -1, l_2b = -1;
    // complex cubes: the whole transform as ONE persistent kernel per rank (slab_fused_kernel.cuh)
    const SlabFusedKernelInfo *fused = nullptr;
    TileKernelInfo fused_ki{};          // the fused kernel's tile shape, for the pass builder
    int l_fy = -1, l_fx = -1, l_fz = -1;
    unsigned *fused_counters = nullptr;
    unsigned fused_grid = 0;
};

// flags: [kind 0 = "receive buffer free", 1.. = "chunk j written"][source rank] epochs
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct SlabPeers {
    unsigned long long *flags[MAX_PEERS];
};

// one thread per destination rank: tell it that this rank's slot (kind, me) reached `epoch`
__global__ void slab_signal_kernel(SlabPeers peers, int G, int me, int kind, unsigned long long epoch) {
    __threadfence_system();
    const int d = threadIdx.x;
    if (d < G) st_release_sys(peers.flags[d] + (size_t)kind * MAX_PEERS + me, epoch);
}

// A rank whose exec failed half-way (or that never called exec) must not leave its peers' GPUs spinning for ever:
// the wait gives up after SLAB_WAIT_TIMEOUT_NS and records the failure in the plan's mapped error word, which every
// later exec on that plan reports as FFTB200_EXEC_FAILED.  (Single-process use: a host thread must not issue calls
// that synchronise all devices - cudaMalloc/cudaFree with peer mappings - between its own exec and its peers'.)
constexpr unsigned long long SLAB_WAIT_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// one thread per source rank: wait until its slot (kind, src) in MY flags reached `epoch`
__global__ void slab_wait_kernel(const unsigned long long *flags, int G, int kind, unsigned long long epoch, int *err) {
    const int s = threadIdx.x;
    if (s < G) {
        const unsigned long long *f = flags + (size_t)kind * MAX_PEERS + s;
        const unsigned long long t0 = global_ns();
        unsigned spins = 0;
        while (ld_acquire_sys(f) < epoch) {
            __nanosleep(200);
            if ((++spins & 0x3ff) == 0 && global_ns() - t0 > SLAB_WAIT_TIMEOUT_NS) {
                *reinterpret_cast<volatile int *>(err) = 1 + kind;
                break;
            }
        }
    }
    __syncthreads();
    __threadfence_system();
}

// all kinds at once: lets every peer's pending and future waits of this epoch terminate (failure path)
__global__ void slab_release_all_kernel(SlabPeers peers, int G, int me, unsigned long long epoch) {
    const int d = threadIdx.x & (MAX_PEERS - 1), kind = threadIdx.x / MAX_PEERS;
    if (d < G && kind < 64) st_release_sys(peers.flags[d] + (size_t)kind * MAX_PEERS + me, epoch);
}

// see add_tile_pass_with: every kernel a slab exec launches is loaded at plan creation
static void preload_slab_kernels() {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, (const void *)slab_signal_kernel) != cudaSuccess) cudaGetLastError();
    if (cudaFuncGetAttributes(&fa, (const void *)slab_wait_kernel) != cudaSuccess) cudaGetLastError();
    if (cudaFuncGetAttributes(&fa, (const void *)slab_release_all_kernel) != cudaSuccess) cudaGetLastError();
}

void slab_free(Plan *P) {
    SlabState *S = P->slab;
    if (!S) return;
    for (int d = 0; d < S->G; ++d)
        if (S->peer_mapped[d] && S->peer_area[d]) cudaIpcCloseMemHandle(S->peer_area[d]);
    if (S->tmp) cudaFree(S->tmp);
    if (S->area) cudaFree(S->area);
    if (S->err_host) cudaFreeHost(S->err_host);
    if (S->aux) cudaStreamDestroy(S->aux);
    for (cudaEvent_t e : S->ev_chunk) cudaEventDestroy(e);
    if (S->ev_done) cudaEventDestroy(S->ev_done);
    for (cudaEvent_t e : S->ev_t)
        if (e) cudaEventDestroy(e);
    cudaGetLastError();
    delete S;
    P->slab = nullptr;
}

static int slab_launch(Plan *P, int idx, const void *src, void *dst, void *const *peers, int inverse, cudaStream_t st) {
    const int npeers = P->slab->G;
    const Launch &ln = P->launches[idx];
    const size_t ce = P->prec ? 16 : 8;
    TileParams tp = ln.tp;
    // the r2c pass addresses its input as packed complex pairs, so in_off is in those units too
    tp.in = (const char *)src + (size_t)ln.in_off * ce;
    tp.out = dst ? (char *)dst + (size_t)ln.out_off * ce : nullptr;
    tp.inverse = inverse;
    if (peers)
        for (int d = 0; d < npeers; ++d) tp.peer[d] = (char *)peers[d] + (size_t)ln.out_off * ce;
    return launch_tile(ln.ki, ln.grid, st, tp) == cudaSuccess ? FFTB200_SUCCESS : FFTB200_EXEC_FAILED;
}


int slab_create(Plan **out, const int *n, fftb200_type type, int rank, int G, int chunks) {
    if (G < 1 || G > MAX_PEERS || rank < 0 || rank >= G) return FFTB200_INVALID_VALUE;
    for (int d = 0; d < 3; ++d)
        if (n[d] < 2 || !is_pow2(n[d])) return FFTB200_INVALID_SIZE;
    if (!is_pow2(G) || n[0] % G || n[1] % G) return FFTB200_INVALID_SIZE;
    std::unique_ptr<Plan> P(new Plan);
    if (cudaGetDevice(&P->device) != cudaSuccess) { cudaGetLastError(); return FFTB200_SETUP_FAILED; }
    P->type = type;
    P->prec = (type == FFTB200_Z2Z || type == FFTB200_D2Z) ? 1 : 0;
    P->real = (type == FFTB200_R2C || type == FFTB200_D2Z);
    P->rank = 3;
    P->batch = 1;
    const int maxL = max_tile_length(P->prec);
    if (n[0] > maxL || n[1] > maxL || (P->real ? n[2] / 2 : n[2]) > maxL || (P->real && n[2] < 4)) return FFTB200_INVALID_SIZE;
    SlabState *S = new SlabState;
    P->slab = S;
    S->rank = rank;
    S->G = G;
    S->n0 = n[0]; S->n1 = n[1]; S->n2 = n[2];
    S->n2c = P->real ? n[2] / 2 + 1 : n[2];
    {
        const long long per_line = P->prec ? 8 : 16;  // complex elements per 128 bytes
        S->n2p = (S->n2c + per_line - 1) / per_line * per_line;
    }
    S->n0l = n[0] / G;
    S->n1l = n[1] / G;
    P->n[0] = S->n0l; P->n[1] = n[1]; P->n[2] = n[2];
    const size_t ce = P->prec ? 16 : 8;
    const long long vol = S->n0l * S->n1 * S->n2p;  // == n1l * n0 * n2p
    Builder B;
    B.P = P.get();
    auto fail = [&](int code) {
        free_plan_resources(P.get());
        return code;
    };
    // chunking of the contiguous index (multiples of 16 columns keep every segment >= 128 B).
    // chunks <= 0: automatic.  Single-CTA tiles of at most 64 KiB: 4 (measured best on 2 x B200 at 512^3), 2 from 8 ranks
    // on (8 x B200, 512^3 D2Z: 0.358 ms with 2, 0.375 with 4, 0.385 unchunked).  128 KiB tiles, which own their SM (the
    // exchange pass then runs on two thirds of the SMs): 1, and 2 from 8 ranks on (1024^3 D2Z: 2.23 ms with 2, 2.28
    // unchunked, 2.27 with 4).  Cluster kernels: 1.
    if (chunks <= 0) {
        const TileKernelInfo *k2 = find_tile_kernel(P->prec, V_CC_PEER, n[1]);
        const bool single = G > 1 && k2 && k2->cluster == 1;
        if (!single) chunks = 1;
        else if (k2->smem_bytes <= 64 * 1024) chunks = G >= 8 ? 2 : 4;
        else chunks = G >= 8 ? 2 : 1;
        const int forced = env_int_or("FFTB200_SLAB_COL_CHUNKS", 0);   // A/B runs
        if (forced > 0) chunks = forced;
    }
    long long cw = (S->n2c + chunks - 1) / chunks;
    cw = (cw + 15) / 16 * 16;
    S->J = (int)((S->n2c + cw - 1) / cw);
    S->chunk_w = cw;
    if (1 + S->J > 64) return fail(FFTB200_INVALID_VALUE);

    if (cudaMalloc(&S->tmp, (size_t)vol * ce) != cudaSuccess) { cudaGetLastError(); return fail(FFTB200_ALLOC_FAILED); }
    S->recv_bytes = (size_t)vol * ce;
    S->flags_off = (S->recv_bytes + 255) / 256 * 256;
    S->area_bytes = S->flags_off + sizeof(unsigned long long) * MAX_PEERS * 64;
    if (cudaMalloc(&S->area, S->area_bytes) != cudaSuccess) { cudaGetLastError(); return fail(FFTB200_ALLOC_FAILED); }
    if (cudaMemset((char *)S->area + S->flags_off, 0, S->area_bytes - S->flags_off) != cudaSuccess) return fail(FFTB200_SETUP_FAILED);
    if (cudaHostAlloc((void **)&S->err_host, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return fail(FFTB200_ALLOC_FAILED); }
    *S->err_host = 0;
    P->work_bytes = (size_t)vol * ce;

    // ---- pass 1: x axis, in [n0l][n1][n2] -> tmp [n0l][n1][n2c]
    {
        std::vector<Level> lv;
        if (P->real) {
            lv.push_back({S->n0l * S->n1, S->n2 / 2, S->n2p});  // rows; input pitch in complex pairs
            if (S->n2 & 1) return fail(FFTB200_INVALID_SIZE);
            if (!add_tile_pass(B, V_RR_R2C, (int)(S->n2 / 2), 1, 1, lv, BUF_IN, BUF_WORK0, 0, "slab pass 1: x axis r2c"))
                return fail(B.err ? B.err : FFTB200_UNSUPPORTED);
        } else {
            lv.push_back({S->n0l * S->n1, S->n2, S->n2p});
            if (!add_tile_pass(B, V_RR, (int)S->n2, 1, 1, lv, BUF_IN, BUF_WORK0, 0, "slab pass 1: x axis"))
                return fail(B.err ? B.err : FFTB200_UNSUPPORTED);
        }
        S->l_pass1 = (int)P->launches.size() - 1;
    }
    const int peer_shift = ilog2ll(S->n1l);
    auto set_peer = [&](Launch &ln) {
        ln.tp.peer_shift = peer_shift;
        ln.tp.peer_mask = (int)S->n1l - 1;
    };
    // ---- pass 2, p2p: y axis on chunk j, tmp -> peers' recv [n1l][n0][n2c] at plane r*n0l + p
    for (int j = 0; j < S->J; ++j) {
        const long long c0 = j * cw, w = std::min(cw, S->n2c - c0);
        std::vector<Level> lv = {{w, 1, 1}, {S->n0l, S->n1 * S->n2p, S->n2p}};
        if (!add_tile_pass(B, V_CC_PEER, (int)S->n1, S->n2p, S->n0 * S->n2p, lv, BUF_WORK0, BUF_OUT, 0,
                           "slab pass 2: y axis, store = exchange (peer memory)"))
            return fail(B.err ? B.err : FFTB200_UNSUPPORTED);
        Launch &ln = P->launches.back();
        ln.in_off = c0;
        ln.out_off = (long long)rank * S->n0l * S->n2p + c0;
        set_peer(ln);
        // The exchange pass is NVLink-bound, not SM-bound: when it is pipelined against the z-axis pass
        // keep it persistent on a bounded number of CTAs so that the HBM-bound pass finds free SMs.
        if (G > 1 && S->J > 1) {
            // (a 128 KiB tile owns its SM outright - all the shared memory it leaves is too small for the z pass's
            // tiles and it holds every register - so such a pass leaves a third of the SMs to the overlapped pass)
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, P->device);
            const unsigned cap = ln.ki->smem_bytes > 64 * 1024 ? (unsigned)env_int_or("FFTB200_SLAB_EX_CTAS", sms * 2 / 3) : (unsigned)sms;
            // static tile assignment on purpose: with dynamic tickets (TileParams::ticket) every capped CTA stays
            // resident until the chunk ends and the overlapped pass starves (2 x B200, 512^3: 1.61 vs 1.40 ms)
            if (cap > 0 && ln.grid > cap) ln.grid = std::max(1u, cap / ln.ki->cluster) * ln.ki->cluster;
        }
        S->l_pass2.push_back((int)P->launches.size() - 1);
    }
    // ---- pass 3, p2p: z axis on chunk j, recv [n1l][n0][n2c] -> out (same layout)
    for (int j = 0; j < S->J; ++j) {
        const long long c0 = j * cw, w = std::min(cw, S->n2c - c0);
        std::vector<Level> lv = {{w, 1, 1}, {S->n1l, S->n0 * S->n2p, S->n0 * S->n2c}};
        if (!add_tile_pass(B, V_CC, (int)S->n0, S->n2p, S->n2c, lv, BUF_WORK1, BUF_OUT, 0, "slab pass 3: z axis"))
            return fail(B.err ? B.err : FFTB200_UNSUPPORTED);
        Launch &ln = P->launches.back();
        ln.in_off = c0;
        ln.out_off = c0;
        S->l_pass3.push_back((int)P->launches.size() - 1);
    }
    // ---- staged mode: pass 2 into the G blocks [d][n0l][n1l][n2c] of a send buffer ...
    {
        std::vector<Level> lv = {{S->n2c, 1, 1}, {S->n0l, S->n1 * S->n2p, S->n1l * S->n2c}};
        if (!add_tile_pass(B, V_CC_PEER, (int)S->n1, S->n2p, S->n2c, lv, BUF_WORK0, BUF_OUT, 0,
                           "slab pass 2 (staged): y axis, store = pack into all-to-all blocks"))
            return fail(B.err ? B.err : FFTB200_UNSUPPORTED);
        set_peer(P->launches.back());
        S->l_pre2 = (int)P->launches.size() - 1;
    }
    // ... and pass 3 from the received blocks [s][n0l][n1l][n2c] == [n0][n1l][n2c] -> out [n1l][n0][n2c]
    {
        std::vector<Level> lv = {{S->n2c, 1, 1}, {S->n1l, S->n2c, S->n0 * S->n2c}};
        if (!add_tile_pass(B, V_CC, (int)S->n0, S->n1l * S->n2c, S->n2c, lv, BUF_IN, BUF_OUT, 0,
                           "slab pass 3 (staged): z axis, load = unpack"))
            return fail(B.err ? B.err : FFTB200_UNSUPPORTED);
        S->l_post3 = (int)P->launches.size() - 1;
    }
    // ---- complex transforms, fused exchange: y axis first.  Before the exchange both x and y are local, after it
    // both x and z are, so the x-axis pass can run on either side.  Doing y + exchange first, chunked over
    // PLANES (each plane is independent), lets the x-axis pass of chunk c (rows of the receive slab whose plane
    // lies in chunk c of any rank) start as soon as chunk c has arrived from every rank: it hides under the
    // NVLink transfer of the following chunks, and only the z-axis pass is left after the exchange.
    if (!P->real) {
        // (decided here because the chunk count depends on it; see the fused block below)
        const SlabFusedKernelInfo *fk = (S->n0 == S->n1 && S->n1 == S->n2) ? find_slab_fused_kernel(P->prec, (int)S->n0) : nullptr;
        const int fused_mode = env_int_or("FFTB200_SLAB_FUSED", -1);
        const bool fused_on = fk && S->n2p == S->n2 && (fused_mode == 1 || (fused_mode != 0 && fk->smem_bytes <= 64 * 1024));
        int want = env_int_or("FFTB200_SLAB_PLANE_CHUNKS", 0);
        // automatic: 4 plane chunks; the single-kernel path keeps chunks of at least 32 planes (8 x B200, 512^3:
        // 2 chunks of 32 planes 0.583 ms, 4 of 16 0.615 ms; 4 x B200: 4 chunks of 32 planes 0.925 ms, 2 of 64 0.954 ms)
        if (want <= 0) want = (G > 1) ? (fused_on ? (int)std::max<long long>(1, std::min<long long>(4, S->n0l / 32)) : 4) : 1;
        long long Jp = 1;
        while (Jp * 2 <= want && S->n0l % (Jp * 2) == 0) Jp *= 2;
        S->Jp = (int)Jp;
        const long long pc = S->n0l / Jp;
        if (1 + S->Jp > 64) return fail(FFTB200_INVALID_VALUE);
        for (int c = 0; c < S->Jp; ++c) {
            std::vector<Level> lv = {{S->n2, 1, 1}, {pc, S->n1 * S->n2, S->n2p}};
            if (!add_tile_pass(B, V_CC_PEER, (int)S->n1, S->n2, S->n0 * S->n2p, lv, BUF_IN, BUF_OUT, 0,
                               "slab y axis (first), store = exchange (peer memory)"))
                return fail(B.err ? B.err : FFTB200_UNSUPPORTED);
            Launch &ln = P->launches.back();
            ln.in_off = (long long)c * pc * S->n1 * S->n2;
            ln.out_off = ((long long)rank * S->n0l + (long long)c * pc) * S->n2p;
            set_peer(ln);
            if (G > 1 && S->Jp > 1 && ln.ki->cluster == 1) {  // (a capped cluster pass pays a cluster barrier per tile)
                const unsigned cap = 148u;
                if (cap > 0 && ln.grid > cap) ln.grid = cap;  // static assignment, see the R2C pipeline above
            }
            S->l_y.push_back((int)P->launches.size() - 1);
        }
        for (int c = 0; c < S->Jp; ++c) {
            std::vector<Level> lv = {{pc, S->n2p, S->n2p}, {(long long)G, S->n0l * S->n2p, S->n0l * S->n2p},
                                     {S->n1l, S->n0 * S->n2p, S->n0 * S->n2p}};
            if (!add_tile_pass(B, V_RR, (int)S->n2, 1, 1, lv, BUF_WORK1, BUF_WORK1, 0,
                               "slab x axis on received planes (overlaps the exchange)"))
                return fail(B.err ? B.err : FFTB200_UNSUPPORTED);
            Launch &ln = P->launches.back();
            ln.in_off = ln.out_off = (long long)c * pc * S->n2p;
            S->l_x.push_back((int)P->launches.size() - 1);
        }
        {
            std::vector<Level> lv = {{S->n2, 1, 1}, {S->n1l, S->n0 * S->n2p, S->n0 * S->n2c}};
            if (!add_tile_pass(B, V_CC, (int)S->n0, S->n2p, S->n2c, lv, BUF_WORK1, BUF_OUT, 0, "slab z axis (last)"))
                return fail(B.err ? B.err : FFTB200_UNSUPPORTED);
            S->l_z = (int)P->launches.size() - 1;
        }
        // ---- the same three passes for the fused single-kernel path (cubes whose side has a fused kernel):
        // y over ALL local planes, x over one chunk's rows (the kernel adds the chunk shift), z over everything
        // Default: only where two of its CTAs fit an SM (tiles of at most 64 KiB), so that an exchange-queue CTA and a
        // local-queue CTA share every SM.  With 128 KiB tiles (1024^3) each role gets half the SMs to itself and the
        // multi-launch path, whose 64 KiB x-axis tiles co-reside with the exchange pass, is faster (2 x B200, 1024^3:
        // 12.3 ms against 17.2 ms); FFTB200_SLAB_FUSED=1 / 0 force the choice.
        if (fused_on && pc >= fk->W) {
            S->fused_ki = TileKernelInfo{reinterpret_cast<void (*)(const TileParams)>(fk->fn), fk->L, fk->R, fk->W, fk->threads,
                                         fk->smem_bytes, 1};
            bool ok = true;
            {
                std::vector<Level> lv = {{S->n2, 1, 1}, {S->n0l, S->n1 * S->n2, S->n2p}};
                ok = ok && add_tile_pass_with(B, &S->fused_ki, V_CC_PEER, (int)S->n1, S->n2, S->n0 * S->n2p, lv, BUF_IN, BUF_OUT, 0,
                                              "fused slab kernel: y axis, store = exchange (peer memory)");
                if (ok) {
                    Launch &ln = P->launches.back();
                    ln.out_off = (long long)rank * S->n0l * S->n2p;
                    set_peer(ln);
                    S->l_fy = (int)P->launches.size() - 1;
                }
            }
            if (ok) {
                std::vector<Level> lv = {{pc, S->n2p, S->n2p}, {(long long)G, S->n0l * S->n2p, S->n0l * S->n2p},
                                         {S->n1l, S->n0 * S->n2p, S->n0 * S->n2p}};
                ok = add_tile_pass_with(B, &S->fused_ki, V_RR, (int)S->n2, 1, 1, lv, BUF_WORK1, BUF_WORK1, 0,
                                        "fused slab kernel: x axis on the received planes of one chunk");
                if (ok) S->l_fx = (int)P->launches.size() - 1;
            }
            if (ok) {
                std::vector<Level> lv = {{S->n2, 1, 1}, {S->n1l, S->n0 * S->n2p, S->n0 * S->n2c}};
                ok = add_tile_pass_with(B, &S->fused_ki, V_CC, (int)S->n0, S->n2p, S->n2c, lv, BUF_WORK1, BUF_OUT, 0,
                                        "fused slab kernel: z axis");
                if (ok) S->l_fz = (int)P->launches.size() - 1;
            }
            if (ok) {
                S->fused_counters = (unsigned *)B.alloc(sizeof(unsigned) * (size_t)(3 + S->Jp));
                int sms = 148, per_sm = 1;
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, P->device);
                if (fk->smem_bytes > 48 * 1024)
                    cudaFuncSetAttribute((const void *)fk->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, fk->smem_bytes);
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)fk->fn, fk->threads, fk->smem_bytes) != cudaSuccess ||
                    per_sm < 1)
                    per_sm = 1;
                cudaGetLastError();
                const long long total = (long long)S->Jp * (P->launches[S->l_fy].tp.n_tiles / S->Jp + P->launches[S->l_fx].tp.n_tiles) +
                                        P->launches[S->l_fz].tp.n_tiles;
                S->fused_grid = (unsigned)std::min<long long>(total, (long long)sms * per_sm);
                for (int l : {S->l_fy, S->l_fx, S->l_fz}) P->launches[l].tp.prefetch_tiles = 0;
                if (S->fused_counters && B.err == FFTB200_SUCCESS) S->fused = fk;
            }
            if (!S->fused && B.err != FFTB200_SUCCESS) return fail(B.err);
        }
    }
    if (cudaStreamCreateWithFlags(&S->aux, cudaStreamNonBlocking) != cudaSuccess) return fail(FFTB200_SETUP_FAILED);
    S->ev_chunk.assign(std::max(S->J, S->Jp), nullptr);
    for (int j = 0; j < std::max(S->J, S->Jp); ++j)
        if (cudaEventCreateWithFlags(&S->ev_chunk[j], cudaEventDisableTiming) != cudaSuccess) return fail(FFTB200_SETUP_FAILED);
    if (cudaEventCreateWithFlags(&S->ev_done, cudaEventDisableTiming) != cudaSuccess) return fail(FFTB200_SETUP_FAILED);
    for (int i = 0; i < 5; ++i)
        if (cudaEventCreate(&S->ev_t[i]) != cudaSuccess) return fail(FFTB200_SETUP_FAILED);
    S->peer_area[rank] = S->area;
    if (G == 1) S->connected = true;
    preload_slab_kernels();
    *out = P.release();
    return FFTB200_SUCCESS;
}

int slab_create_2d(Plan **out, const int *n, fftb200_type type, int rank, int G) {
    if (G < 1 || G > MAX_PEERS || rank < 0 || rank >= G) return FFTB200_INVALID_VALUE;
    for (int d = 0; d < 2; ++d)
        if (n[d] < 2 || !is_pow2(n[d])) return FFTB200_INVALID_SIZE;
    if (!is_pow2(G) || n[0] % G || n[1] % G) return FFTB200_INVALID_SIZE;
    std::unique_ptr<Plan> P(new Plan);
    if (cudaGetDevice(&P->device) != cudaSuccess) { cudaGetLastError(); return FFTB200_SETUP_FAILED; }
    P->type = type;
    P->prec = (type == FFTB200_Z2Z) ? 1 : 0;
    P->rank = 2;
    P->batch = 1;
    const int maxL = max_tile_length(P->prec);
    if (n[0] > maxL || n[1] > maxL) return FFTB200_INVALID_SIZE;
    SlabState *S = new SlabState;
    P->slab = S;
    S->two_d = true;
    S->rank = rank;
    S->G = G;
    S->J = S->Jp = 1;
    S->n0 = n[0]; S->n1 = n[1]; S->n2 = S->n2c = 1;
    S->n0l = n[0] / G;
    S->n1l = n[1] / G;
    const long long per_line = P->prec ? 8 : 16;
    const long long n0p = (S->n0 + per_line - 1) / per_line * per_line;  // receive-slab row pitch, whole 128-byte lines
    S->n2p = n0p;
    P->n[0] = S->n0l; P->n[1] = n[1];
    const size_t ce = P->prec ? 16 : 8;
    Builder B;
    B.P = P.get();
    auto fail = [&](int code) {
        free_plan_resources(P.get());
        return code;
    };
    S->recv_bytes = (size_t)(S->n1l * n0p) * ce;
    S->flags_off = (S->recv_bytes + 255) / 256 * 256;
    S->area_bytes = S->flags_off + sizeof(unsigned long long) * MAX_PEERS * 64;
    if (cudaMalloc(&S->area, S->area_bytes) != cudaSuccess) { cudaGetLastError(); return fail(FFTB200_ALLOC_FAILED); }
    if (cudaMemset((char *)S->area + S->flags_off, 0, S->area_bytes - S->flags_off) != cudaSuccess) return fail(FFTB200_SETUP_FAILED);
    if (cudaHostAlloc((void **)&S->err_host, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return fail(FFTB200_ALLOC_FAILED); }
    *S->err_host = 0;
    P->work_bytes = S->recv_bytes;
    {   // pass A: rows of n1 (contiguous) -> column k of the transpose lives on rank k / n1l
        std::vector<Level> lv = {{S->n0l, S->n1, 1}};
        if (!add_tile_pass(B, V_RC_PEER, (int)S->n1, 1, n0p, lv, BUF_IN, BUF_OUT, 0, "2-D slab: row FFT, store = global transpose (peer memory)"))
            return fail(B.err ? B.err : FFTB200_UNSUPPORTED);
        Launch &ln = P->launches.back();
        ln.out_off = (long long)rank * S->n0l;
        ln.tp.peer_shift = ilog2ll(S->n1l);
        ln.tp.peer_mask = (int)S->n1l - 1;
        S->l_2a = (int)P->launches.size() - 1;
    }
    {   // pass B: the received slab [n1l][n0p] holds whole columns: row FFT over n0 -> out [n1l][n0]
        std::vector<Level> lv = {{S->n1l, n0p, S->n0}};
        if (!add_tile_pass(B, V_RR, (int)S->n0, 1, 1, lv, BUF_WORK1, BUF_OUT, 0, "2-D slab: column FFT on the received slab"))
            return fail(B.err ? B.err : FFTB200_UNSUPPORTED);
        S->l_2b = (int)P->launches.size() - 1;
    }
    for (int i = 0; i < 5; ++i)
        if (cudaEventCreate(&S->ev_t[i]) != cudaSuccess) return fail(FFTB200_SETUP_FAILED);
    S->peer_area[rank] = S->area;
    if (G == 1) S->connected = true;
    preload_slab_kernels();
    *out = P.release();
    return FFTB200_SUCCESS;
}

static unsigned long long *slab_flags(SlabState *S, int d) {
    return (unsigned long long *)((char *)S->peer_area[d] + S->flags_off);
}

// fused exchange: every rank calls this once per transform (collective)
static int slab_exec_p2p_body(Plan *P, const void *in, void *out, int inverse, const SlabPeers &peers, void *const *recv,
                              unsigned long long epoch) {
    SlabState *S = P->slab;
    cudaStream_t st = P->stream;
    const unsigned long long *myflags = slab_flags(S, S->rank);
    if (S->timing) cudaEventRecord(S->ev_t[0], st);
    // my receive buffer is free again (stream order: after my previous transform's pass 3); the fused kernel says so itself
    if (S->G > 1 && !(S->fused && !P->real)) slab_signal_kernel<<<1, 32, 0, st>>>(peers, S->G, S->rank, 0, epoch);
    if (S->two_d) {
        if (S->G > 1) slab_wait_kernel<<<1, 32, 0, st>>>(myflags, S->G, 0, epoch, S->err_host);
        if (S->timing) cudaEventRecord(S->ev_t[1], st);
        int rc2 = slab_launch(P, S->l_2a, in, nullptr, recv, inverse, st);
        if (rc2) return rc2;
        if (S->G > 1) {
            slab_signal_kernel<<<1, 32, 0, st>>>(peers, S->G, S->rank, 1, epoch);
            slab_wait_kernel<<<1, 32, 0, st>>>(myflags, S->G, 1, epoch, S->err_host);
        }
        if (S->timing) cudaEventRecord(S->ev_t[2], st);
        rc2 = slab_launch(P, S->l_2b, S->area, out, nullptr, inverse, st);
        if (rc2) return rc2;
        if (S->timing) cudaEventRecord(S->ev_t[3], st);
        return cudaGetLastError() == cudaSuccess ? FFTB200_SUCCESS : FFTB200_EXEC_FAILED;
    }
    if (!P->real && S->fused) {
        // one persistent kernel: y axis + exchange, hand-shakes, x axis on arrived chunks, z axis
        const size_t ce = P->prec ? 16 : 8;
        SlabFusedParams fp;
        const Launch &ly = P->launches[S->l_fy], &lx = P->launches[S->l_fx], &lz = P->launches[S->l_fz];
        fp.y = ly.tp;
        fp.y.in = in;
        fp.y.out = nullptr;
        fp.y.inverse = inverse;
        for (int d = 0; d < S->G; ++d) fp.y.peer[d] = (char *)recv[d] + (size_t)ly.out_off * ce;
        fp.x = lx.tp;
        fp.x.in = S->area;
        fp.x.out = S->area;
        fp.x.inverse = inverse;
        fp.z = lz.tp;
        fp.z.in = S->area;
        fp.z.out = out;
        fp.z.inverse = inverse;
        fp.counters = S->fused_counters;
        for (int d = 0; d < MAX_PEERS; ++d) fp.flags[d] = peers.flags[d];
        fp.err = S->err_host;
        fp.epoch = epoch;
        fp.x_chunk_shift = (S->n0l / S->Jp) * S->n2p;
        fp.G = S->G;
        fp.me = S->rank;
        fp.n_chunks = S->Jp;
        fp.tiles_y_chunk = ly.tp.n_tiles / S->Jp;
        fp.tiles_x_chunk = lx.tp.n_tiles;
        fp.tiles_z = lz.tp.n_tiles;
        if (cudaMemsetAsync(S->fused_counters, 0, sizeof(unsigned) * (size_t)(3 + S->Jp), st) != cudaSuccess) return FFTB200_EXEC_FAILED;
        if (S->timing) cudaEventRecord(S->ev_t[1], st);
        S->fused->fn<<<S->fused_grid, S->fused->threads, S->fused->smem_bytes, st>>>(fp);
        if (S->timing) { cudaEventRecord(S->ev_t[2], st); cudaEventRecord(S->ev_t[3], st); }
        return cudaGetLastError() == cudaSuccess ? FFTB200_SUCCESS : FFTB200_EXEC_FAILED;
    }
    if (!P->real) {
        // y axis + exchange (plane chunks) -> x axis on arrived chunks (second stream) -> z axis
        if (S->G > 1) slab_wait_kernel<<<1, 32, 0, st>>>(myflags, S->G, 0, epoch, S->err_host);
        if (S->timing) cudaEventRecord(S->ev_t[1], st);
        for (int c = 0; c < S->Jp; ++c) {
            int rc2 = slab_launch(P, S->l_y[c], in, nullptr, recv, inverse, st);
            if (rc2) return rc2;
            if (S->G > 1) slab_signal_kernel<<<1, 32, 0, st>>>(peers, S->G, S->rank, 1 + c, epoch);
            if (c == S->Jp - 1 && S->timing) cudaEventRecord(S->ev_t[2], st);
            cudaStream_t sx = (S->Jp > 1) ? S->aux : st;
            if (S->Jp > 1) {
                cudaEventRecord(S->ev_chunk[c], st);
                cudaStreamWaitEvent(sx, S->ev_chunk[c], 0);
            }
            if (S->G > 1) slab_wait_kernel<<<1, 32, 0, sx>>>(myflags, S->G, 1 + c, epoch, S->err_host);
            rc2 = slab_launch(P, S->l_x[c], S->area, S->area, nullptr, inverse, sx);
            if (rc2) return rc2;
        }
        if (S->Jp > 1) {
            cudaEventRecord(S->ev_done, S->aux);
            cudaStreamWaitEvent(st, S->ev_done, 0);
        }
        const int rc3 = slab_launch(P, S->l_z, S->area, out, nullptr, inverse, st);
        if (rc3) return rc3;
        if (S->timing) cudaEventRecord(S->ev_t[3], st);
        return cudaGetLastError() == cudaSuccess ? FFTB200_SUCCESS : FFTB200_EXEC_FAILED;
    }
    int rc = slab_launch(P, S->l_pass1, in, S->tmp, nullptr, inverse, st);
    if (rc) return rc;
    if (S->timing) cudaEventRecord(S->ev_t[1], st);
    if (S->G > 1) slab_wait_kernel<<<1, 32, 0, st>>>(myflags, S->G, 0, epoch, S->err_host);
    for (int j = 0; j < S->J; ++j) {
        rc = slab_launch(P, S->l_pass2[j], S->tmp, nullptr, recv, inverse, st);
        if (rc) return rc;
        if (S->G > 1) slab_signal_kernel<<<1, 32, 0, st>>>(peers, S->G, S->rank, 1 + j, epoch);
        if (j == S->J - 1 && S->timing) cudaEventRecord(S->ev_t[2], st);
        cudaStream_t s3 = (S->J > 1) ? S->aux : st;
        if (S->J > 1) {
            cudaEventRecord(S->ev_chunk[j], st);
            cudaStreamWaitEvent(s3, S->ev_chunk[j], 0);
        }
        if (S->G > 1) slab_wait_kernel<<<1, 32, 0, s3>>>(myflags, S->G, 1 + j, epoch, S->err_host);
        rc = slab_launch(P, S->l_pass3[j], S->area, out, nullptr, inverse, s3);
        if (rc) return rc;
    }
    if (S->J > 1) {
        cudaEventRecord(S->ev_done, S->aux);
        cudaStreamWaitEvent(st, S->ev_done, 0);
    }
    if (S->timing) cudaEventRecord(S->ev_t[3], st);
    return cudaGetLastError() == cudaSuccess ? FFTB200_SUCCESS : FFTB200_EXEC_FAILED;
}

int slab_exec_p2p(Plan *P, const void *in, void *out, int inverse) {
    SlabState *S = P->slab;
    if (!S->connected) return FFTB200_INVALID_PLAN;
    if (*reinterpret_cast<volatile int *>(S->err_host) != 0) {  // an earlier hand-shake timed out (1 + flag kind) or failed
        if (getenv("FFTB200_DEBUG"))
            fprintf(stderr, "libfft_b200: slab plan of rank %d is poisoned: error word %d (epoch %llu)\n", S->rank, *S->err_host, S->epoch);
        return FFTB200_EXEC_FAILED;
    }
    DeviceGuard g(P->device);
    std::lock_guard<std::mutex> lk(P->mu);
    const unsigned long long epoch = ++S->epoch;
    SlabPeers peers;
    void *recv[MAX_PEERS];
    for (int d = 0; d < MAX_PEERS; ++d) {
        peers.flags[d] = d < S->G ? slab_flags(S, d) : nullptr;
        recv[d] = d < S->G ? S->peer_area[d] : nullptr;
    }
    const int rc = slab_exec_p2p_body(P, in, out, inverse, peers, recv, epoch);
    if (rc != FFTB200_SUCCESS && getenv("FFTB200_DEBUG"))
        fprintf(stderr, "libfft_b200: slab exec failed on rank %d (code %d), last CUDA error: %s\n", S->rank, rc,
                cudaGetErrorString(cudaPeekAtLastError()));
    if (rc != FFTB200_SUCCESS && S->G > 1) {
        // a launch failed in the middle of the collective: publish every flag of this epoch anyway so that the peers'
        // wait kernels terminate (their result is garbage, they learn of it through the caller's error handling),
        // and poison this plan
        cudaGetLastError();
        slab_release_all_kernel<<<1, 1024, 0, P->stream>>>(peers, S->G, S->rank, epoch);
        cudaGetLastError();
        *S->err_host = 100;
    }
    return rc;
}

int slab_exec_pre(Plan *P, const void *in, void *send, int inverse) {
    SlabState *S = P->slab;
    DeviceGuard g(P->device);
    std::lock_guard<std::mutex> lk(P->mu);
    const size_t ce = P->prec ? 16 : 8;
    int rc = slab_launch(P, S->l_pass1, in, S->tmp, nullptr, inverse, P->stream);
    if (rc) return rc;
    void *blocks[MAX_PEERS] = {};
    for (int d = 0; d < S->G; ++d) blocks[d] = (char *)send + (size_t)d * S->n0l * S->n1l * S->n2c * ce;
    return slab_launch(P, S->l_pre2, S->tmp, nullptr, blocks, inverse, P->stream);
}

int slab_exec_post(Plan *P, const void *recv, void *out, int inverse) {
    SlabState *S = P->slab;
    DeviceGuard g(P->device);
    std::lock_guard<std::mutex> lk(P->mu);
    return slab_launch(P, S->l_post3, recv, out, nullptr, inverse, P->stream);
}

// ---- accessors used by the C ABI (abi.cu) ---------------------------------------------------------------
int slab_get_ipc_handle(Plan *P, void *handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard g(P->device);
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, P->slab->area) != cudaSuccess) { cudaGetLastError(); return FFTB200_SETUP_FAILED; }
    memcpy(handle64, &h, 64);
    return FFTB200_SUCCESS;
}

int slab_connect_ipc(Plan *P, const void *handles) {
    SlabState *S = P->slab;
    DeviceGuard g(P->device);
    std::lock_guard<std::mutex> lk(P->mu);
    for (int d = 0; d < S->G; ++d) {
        if (d == S->rank || S->peer_mapped[d]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * (size_t)d, 64);
        void *ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            return FFTB200_SETUP_FAILED;
        }
        S->peer_area[d] = ptr;
        S->peer_mapped[d] = true;
    }
    S->connected = true;
    return FFTB200_SUCCESS;
}

int slab_get_area(Plan *P, void **area, unsigned long long *bytes) {
    *area = P->slab->area;
    if (bytes) *bytes = P->slab->area_bytes;
    return FFTB200_SUCCESS;
}

int slab_connect_ptrs(Plan *P, void *const *areas) {
    SlabState *S = P->slab;
    DeviceGuard g(P->device);
    std::lock_guard<std::mutex> lk(P->mu);
    for (int d = 0; d < S->G; ++d) {
        if (d == S->rank) continue;
        if (!areas[d]) return FFTB200_INVALID_VALUE;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, areas[d]) != cudaSuccess) { cudaGetLastError(); return FFTB200_INVALID_VALUE; }
        if (a.device != P->device) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(a.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return FFTB200_SETUP_FAILED; }
            cudaGetLastError();
        }
        S->peer_area[d] = areas[d];
    }
    S->connected = true;
    return FFTB200_SUCCESS;
}

int slab_set_timing(Plan *P, int on) {
    std::lock_guard<std::mutex> lk(P->mu);
    P->slab->timing = on != 0;
    return FFTB200_SUCCESS;
}

// ms[0] = first phase, ms[1] = pass(es) with the exchange, ms[2] = remaining local work
int slab_get_phase_ms(Plan *P, float *ms) {
    SlabState *S = P->slab;
    std::lock_guard<std::mutex> lk(P->mu);
    DeviceGuard g(P->device);
    if (cudaEventSynchronize(S->ev_t[3]) != cudaSuccess) { cudaGetLastError(); return FFTB200_INVALID_VALUE; }
    for (int i = 0; i < 3; ++i)
        if (cudaEventElapsedTime(&ms[i], S->ev_t[i], S->ev_t[i + 1]) != cudaSuccess) { cudaGetLastError(); return FFTB200_INVALID_VALUE; }
    return FFTB200_SUCCESS;
}

// kernels one fused slab exec issues: pass 1, J x (pass 2 + pass 3), hand-shake kernels
int slab_launches_per_exec(const Plan *P) {
    const SlabState *S = P->slab;
    if (S->two_d) return 2 + (S->G > 1 ? 4 : 0);
    if (S->fused && !P->real) return 1;
    const int J = P->real ? S->J : S->Jp;
    return 1 + 2 * J + (S->G > 1 ? 2 + 2 * J : 0);
}

}  // namespace fftb200
