// slab_fused_kernel.cuh — one persistent kernel per rank for a whole slab-decomposed complex 3-D transform:
// the y-axis pass whose store is the exchange (peer memory over NVLink), the hand-shake with the peers, the x-axis
// pass on the planes as they arrive and the z-axis pass, without returning to the host-side streams in between.
//
// Round 1 ran this as 2 + 2 * chunks tile launches plus 2 + 2 * chunks one-warp signal / wait kernels on two streams
// chained by events; at 8 GPUs x 512^3 those hand-shake launches and stream hops were 0.10-0.17 ms of a 0.59 ms step.
// Here the CTAs of ONE launch draw tickets from two queues,
//     exchange queue:  Y(0) Y(1) ... Y(J-1)          (Y(c) = y-axis tiles of plane chunk c; NVLink-bound)
//     local queue:     X(0) X(1) ... X(J-1) Z        (X(c) = x-axis tiles of the rows of the receive slab whose plane lies
//                                                      in chunk c of any rank; Z = z-axis tiles; HBM-bound)
// half of them serving the exchange queue first, the other half the local queue, and
//   * the last tile of Y(c) to finish publishes "chunk c written" to every peer (system-scope release store),
//   * an X(c) tile first waits (one thread, system-scope acquire loads) until every peer has published chunk c,
//   * a Z tile first waits until every X tile of this rank has been stored.
// A waiting CTA keeps its SM slot but cannot block what it waits for: Y tiles wait for nothing but the peers' "receive
// slab free" flags of this epoch, which every rank publishes when its kernel starts, and each queue is handed out in
// order.  No cooperative launch is needed.
//
// No reference counterpart (src/fft.rg has no distributed transform); the decomposition is FFTW-MPI's
// (fftw-3.3.8/mpi/dft-rank-geq2.c:40-59, mpi/transpose-alltoall.c:49-100, doc/mpi.texi:443-466), see slab_plan.cu.
#pragma once
#include "tile_kernel.cuh"

namespace fftb200 {

struct SlabFusedParams {
    TileParams y, x, z;       // whole passes: y over all local planes, x over one chunk's rows (+ chunk shift), z over all
    unsigned *counters;        // [0] exchange-queue ticket, [1] local-queue ticket, [2] finished x tiles,
                               // [3 + c] finished y tiles of chunk c; zeroed before every launch
    unsigned long long *flags[MAX_PEERS];  // every rank's flag block (mine included): [kind][source rank] epochs
    int *err;                  // mapped host word: a wait that gave up records it here
    unsigned long long epoch;
    long long x_chunk_shift;   // elements between consecutive chunks' rows in the receive slab
    int G, me;
    int n_chunks;
    int tiles_y_chunk, tiles_x_chunk, tiles_z;
};

__device__ __forceinline__ void fused_st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long fused_ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned fused_ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long fused_global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long FUSED_WAIT_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;

// thread 0: wait until all G ranks' slot (kind, src) in MY flag block reached epoch
__device__ __forceinline__ void fused_wait_peers(const SlabFusedParams &p, int kind) {
    const unsigned long long *mine = p.flags[p.me] + (size_t)kind * MAX_PEERS;
    const unsigned long long t0 = fused_global_ns();
    for (int s = 0; s < p.G; ++s) {
        unsigned spins = 0;
        while (fused_ld_acquire_sys(mine + s) < p.epoch) {
            __nanosleep(100);
            if ((++spins & 0x3ff) == 0 && fused_global_ns() - t0 > FUSED_WAIT_TIMEOUT_NS) {
                *reinterpret_cast<volatile int *>(p.err) = 1 + kind;
                return;
            }
        }
    }
}

template <typename T, int L, int R, int W>
__global__ void __launch_bounds__(TileTraits<T, L, R, W, V_CC>::THREADS, TileTraits<T, L, R, W, V_CC>::MIN_CTAS)
fft_slab_fused_kernel(const SlabFusedParams p) {
    using TY = TileTraits<T, L, R, W, V_CC_PEER>;
    using TX = TileTraits<T, L, R, W, V_RR>;
    using TZ = TileTraits<T, L, R, W, V_CC>;
    static_assert(TY::THREADS == TX::THREADS && TX::THREADS == TZ::THREADS, "the three passes share one CTA shape");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_ticket;
    const int J = p.n_chunks;
    const int per_y = p.tiles_y_chunk, per_x = p.tiles_x_chunk;
    const int total_ex = J * per_y;                 // exchange queue: Y(0) .. Y(J-1)
    const int total_loc = J * per_x + p.tiles_z;    // local queue:    X(0) .. X(J-1), Z
    unsigned *ex_ticket = p.counters, *loc_ticket = p.counters + 1, *x_done = p.counters + 2, *y_done = p.counters + 3;

    // this rank's receive slab is free again (stream order: its previous transform's z pass has finished)
    if (blockIdx.x == 0 && (int)threadIdx.x < p.G)
        fused_st_release_sys(p.flags[threadIdx.x] + (size_t)0 * MAX_PEERS + p.me, p.epoch);

    // Two queues, two roles.  The exchange pass is NVLink-bound and its CTAs sit in store back-pressure; the x / z passes
    // are HBM-bound.  Half the CTAs (the first half of the grid: one per SM when two fit) serve the exchange queue first,
    // the others the local queue, so both links stay busy; a CTA whose own queue is empty helps with the other one.
    //
    // Completion is published per CTA and per chunk, not per tile: a system-scope fence after peer stores costs a full
    // NVLink round trip, and one per tile serialised every CTA (measured: 2 x B200, 512^3: 1.88 ms against 1.36 ms for
    // the multi-launch path).  A CTA counts the tiles it finished in the current (kind, chunk) and adds them to the
    // chunk's counter - one fence, one atomic - when its next ticket belongs to something else or its queue runs dry;
    // always before it waits for anything.
    const bool ex_role = blockIdx.x < (gridDim.x + 1) / 2;
    bool ex_open = true, loc_open = true, peers_free = false, x_all_done = false;
    unsigned long long ready_mask = 0;             // chunks (< 64) whose arrival this CTA has already observed
    int pend_kind = 0, pend_chunk = 0;             // 1 = Y tiles, 2 = X tiles finished but not yet published
    unsigned pend_count = 0;
    auto flush = [&]() {
        if (pend_kind != 0 && threadIdx.x == 0) {
            if (pend_kind == 1) {
                __threadfence_system();
                const unsigned done = atomicAdd(y_done + pend_chunk, pend_count);
                if (done + pend_count == (unsigned)per_y) {
                    __threadfence_system();
                    for (int d = 0; d < p.G; ++d)
                        fused_st_release_sys(p.flags[d] + (size_t)(1 + pend_chunk) * MAX_PEERS + p.me, p.epoch);
                }
            } else {
                __threadfence();
                atomicAdd(x_done, pend_count);
            }
        }
        pend_kind = 0;
        pend_count = 0;
    };
    while (ex_open || loc_open) {
        const bool take_ex = ex_open && (ex_role || !loc_open);
        if (threadIdx.x == 0) s_ticket = (int)atomicAdd(take_ex ? ex_ticket : loc_ticket, 1u);
        __syncthreads();
        const int ticket = s_ticket;
        __syncthreads();  // s_ticket may be rewritten
        if (take_ex) {
            if (ticket >= total_ex) { ex_open = false; flush(); continue; }
            const int chunk = ticket / per_y;
            if (pend_kind != 0 && (pend_kind != 1 || pend_chunk != chunk)) flush();
            if (!peers_free) {
                // the first store into a peer's receive slab waits until that peer has released it for this epoch
                if (threadIdx.x == 0 && p.G > 1) fused_wait_peers(p, 0);
                __syncthreads();
                peers_free = true;
            }
            fft_tile_body<T, L, R, W, V_CC_PEER>(p.y, ticket, smem_raw);
            pend_kind = 1;
            pend_chunk = chunk;
            ++pend_count;
        } else {
            if (ticket >= total_loc) { loc_open = false; flush(); continue; }
            if (ticket < J * per_x) {
                const int chunk = ticket / per_x, t = ticket - chunk * per_x;
                if (pend_kind != 0 && pend_kind != 2) flush();
                if (!((ready_mask >> chunk) & 1ull)) {
                    flush();  // never wait with unpublished work
                    if (threadIdx.x == 0) fused_wait_peers(p, 1 + chunk);  // (G == 1: my own flag, set by my last Y tile)
                    __syncthreads();
                    ready_mask |= 1ull << chunk;
                }
                const long long sh = (long long)chunk * p.x_chunk_shift;
                fft_tile_body<T, L, R, W, V_RR, true>(p.x, t, smem_raw, sh, sh);
                pend_kind = 2;
                pend_chunk = 0;
                ++pend_count;
            } else {
                flush();
                if (!x_all_done) {
                    if (threadIdx.x == 0) {
                        const unsigned want = (unsigned)(J * per_x);
                        const unsigned long long t0 = fused_global_ns();
                        unsigned spins = 0;
                        while (fused_ld_acquire_gpu(x_done) < want) {
                            __nanosleep(100);
                            if ((++spins & 0x3ff) == 0 && fused_global_ns() - t0 > FUSED_WAIT_TIMEOUT_NS) {
                                *reinterpret_cast<volatile int *>(p.err) = 99;
                                break;
                            }
                        }
                    }
                    __syncthreads();
                    x_all_done = true;
                }
                fft_tile_body<T, L, R, W, V_CC, true>(p.z, ticket - J * per_x, smem_raw);
            }
        }
        __syncthreads();  // this tile's stores are issued by every thread; its shared memory is free
    }
    flush();
}

struct SlabFusedKernelInfo {
    void (*fn)(const SlabFusedParams);
    int prec, L, R, W, threads, smem_bytes;
};
// fused slab kernel for cubes of side L (all three passes with the tile shape the column tables use for L), or nullptr
const SlabFusedKernelInfo *find_slab_fused_kernel(int prec, int L);

}  // namespace fftb200
