// slab_fused_kernel.cuh — one persistent kernel per rank for a whole slab-decomposed complex 3-D transform:
// the y-axis pass whose store is the exchange (peer memory over NVLink), the hand-shake with the peers, the x-axis
// pass on the planes as they arrive and the z-axis pass, without returning to the host-side streams in between.
//
// Round 1 ran this as 2 + 2 * chunks tile launches plus 2 + 2 * chunks one-warp signal / wait kernels on two streams
// chained by events; at 8 GPUs x 512^3 those hand-shake launches and stream hops were 0.10-0.17 ms of a 0.59 ms step.
// Here the CTAs of ONE launch draw tickets from a counter; the ticket order is
//     Y(0) Y(1) | X(0) Y(2) | X(1) Y(3) | ... | X(J-2) | X(J-1) | Z
// (Y(c) = y-axis tiles of plane chunk c, X(c) = x-axis tiles of the rows of the receive slab whose plane lies in chunk c
// of any rank, Z = z-axis tiles), so that
//   * the last tile of Y(c) to finish publishes "chunk c written" to every peer (system-scope release store),
//   * an X(c) tile first waits (one thread, system-scope acquire loads) until every peer has published chunk c,
//   * a Z tile first waits until every X tile of this rank has been stored.
// A waiting CTA keeps its SM slot but never blocks the tickets before it: Y tiles wait for nothing but the peers'
// "receive slab free" flags of this epoch, which every rank publishes when its kernel starts.  No cooperative launch is
// needed: tickets are handed out in order, so whatever a ticket waits for is already running or finished somewhere.
//
// No reference counterpart (src/fft.rg has no distributed transform); the decomposition is FFTW-MPI's
// (fftw-3.3.8/mpi/dft-rank-geq2.c:40-59, mpi/transpose-alltoall.c:49-100, doc/mpi.texi:443-466), see slab_plan.cu.
#pragma once
#include "tile_kernel.cuh"

namespace fftb200 {

struct SlabFusedParams {
    TileParams y, x, z;       // whole passes: y over all local planes, x over one chunk's rows (+ chunk shift), z over all
    unsigned *counters;        // [0] ticket, [1] finished x tiles, [2 + c] finished y tiles of chunk c; zeroed before launch
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
    __shared__ int s_ticket[2];
    const int J = p.n_chunks;
    // ticket layout: group g in [0, J+1]: X(g-2)... see the header.  Slot s = 0..J+1 holds Y(s) (if s < J) after X(s-2) (if s >= 2)
    const int per_y = p.tiles_y_chunk, per_x = p.tiles_x_chunk;
    const int total = J * (per_y + per_x) + p.tiles_z;

    // this rank's receive slab is free again (stream order: its previous transform's z pass has finished)
    if (blockIdx.x == 0 && (int)threadIdx.x < p.G)
        fused_st_release_sys(p.flags[threadIdx.x] + (size_t)0 * MAX_PEERS + p.me, p.epoch);

    bool peers_free = false;
    int cur = 0;
    if (threadIdx.x == 0) s_ticket[0] = (int)atomicAdd(p.counters, 1u);
    __syncthreads();
    for (;;) {
        const int ticket = s_ticket[cur];
        if (ticket >= total) break;
        unsigned next_ticket = 0;
        if (threadIdx.x == 0) next_ticket = atomicAdd(p.counters, 1u);  // its latency hides behind this tile
        // decode: slots s = 0 .. J+1; slot s holds X(s-2) first (s >= 2), then Y(s) (s < J); then all of Z
        int phase, chunk, t;  // phase 0 = Y, 1 = X, 2 = Z
        {
            int r = ticket;
            const int head = 2 * per_y;                       // slots 0 and 1: Y(0), Y(1) only (fewer if J < 2)
            const int n_head = J < 2 ? J : 2;
            if (r < n_head * per_y) {
                phase = 0; chunk = r / per_y; t = r - chunk * per_y;
            } else {
                r -= n_head * per_y;
                const int mid_slots = J > 2 ? J - 2 : 0;      // slots 2 .. J-1: X(s-2) then Y(s)
                const int per_slot = per_x + per_y;
                if (r < mid_slots * per_slot) {
                    const int s = r / per_slot, q = r - s * per_slot;
                    if (q < per_x) { phase = 1; chunk = s; t = q; }
                    else { phase = 0; chunk = s + 2; t = q - per_x; }
                } else {
                    r -= mid_slots * per_slot;
                    const int tail_x = (J < 2 ? J : 2) * per_x;  // X(J-2), X(J-1)  (X(0) only if J == 1)
                    if (r < tail_x) {
                        const int k = r / per_x;
                        phase = 1; chunk = (J < 2 ? 0 : J - 2) + k; t = r - k * per_x;
                    } else {
                        phase = 2; chunk = 0; t = r - tail_x;
                    }
                }
            }
            (void)head;
        }
        if (phase == 0) {
            if (!peers_free) {
                // the first store into a peer's receive slab waits until that peer has released it for this epoch
                if (threadIdx.x == 0 && p.G > 1) fused_wait_peers(p, 0);
                __syncthreads();
                peers_free = true;
            }
            fft_tile_body<T, L, R, W, V_CC_PEER>(p.y, chunk * per_y + t, smem_raw);
            __syncthreads();  // every thread's peer stores are issued
            if (threadIdx.x == 0) {
                __threadfence_system();
                const unsigned done = atomicAdd(p.counters + 2 + chunk, 1u);
                if (done == (unsigned)per_y - 1) {
                    __threadfence_system();
                    for (int d = 0; d < p.G; ++d) fused_st_release_sys(p.flags[d] + (size_t)(1 + chunk) * MAX_PEERS + p.me, p.epoch);
                }
            }
        } else if (phase == 1) {
            if (threadIdx.x == 0) fused_wait_peers(p, 1 + chunk);  // (G == 1: my own flag, published by my last Y tile)
            __syncthreads();
            const long long sh = (long long)chunk * p.x_chunk_shift;
            fft_tile_body<T, L, R, W, V_RR, true>(p.x, t, smem_raw, sh, sh);
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                atomicAdd(p.counters + 1, 1u);
            }
        } else {
            if (threadIdx.x == 0) {
                const unsigned want = (unsigned)(J * per_x);
                const unsigned long long t0 = fused_global_ns();
                unsigned spins = 0;
                while (fused_ld_acquire_gpu(p.counters + 1) < want) {
                    __nanosleep(100);
                    if ((++spins & 0x3ff) == 0 && fused_global_ns() - t0 > FUSED_WAIT_TIMEOUT_NS) {
                        *reinterpret_cast<volatile int *>(p.err) = 99;
                        break;
                    }
                }
            }
            __syncthreads();
            fft_tile_body<T, L, R, W, V_CC, true>(p.z, t, smem_raw);
        }
        if (threadIdx.x == 0) s_ticket[cur ^ 1] = (int)next_ticket;
        __syncthreads();  // next ticket visible; this tile's shared memory free
        cur ^= 1;
    }
}

struct SlabFusedKernelInfo {
    void (*fn)(const SlabFusedParams);
    int prec, L, R, W, threads, smem_bytes;
};
// fused slab kernel for cubes of side L (all three passes with the tile shape the column tables use for L), or nullptr
const SlabFusedKernelInfo *find_slab_fused_kernel(int prec, int L);

}  // namespace fftb200
