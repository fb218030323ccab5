// peer_probe.cu — raw all-to-all bandwidth over NVLink from kernels, push (remote STG.128) against pull (remote LDG.128),
// with the access shapes of the slab exchange (tools only, not product code).  One process, G GPUs, peer access.
//
// Every GPU moves (G-1)/G of a `mb`-MiB buffer: block d of its source goes to GPU d.
//   push-tile   each warp stores 4 rows of 128 B (16 B per lane), rows `stride` apart in the destination: the COL->peers
//               store of fft_tile_kernel
//   push-flat   each warp stores 512 contiguous bytes
//   pull-row    each warp loads 512 contiguous bytes from the peer and stores them locally (what a row pass reading a
//               peer's slab would do)
//   pull-tile   each warp loads 4 rows of 128 B from the peer
// Prints GB/s per direction per GPU (bytes that crossed NVLink out of one GPU / time, max over GPUs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o peer_probe peer_probe.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

struct Ptrs { double2 *p[16]; };

// elements: a block is `blk` double2; tile mode views a block as [rows][8] with a row pitch of `pitch` elements
template <bool PUSH, bool TILE, int R>
__global__ void __launch_bounds__(512) xfer(Ptrs remote, double2 *local, int G, int me, long long blk, int ctas_per_peer) {
    const int d = (blockIdx.x / ctas_per_peer);           // peer index handled by this CTA (skips me below)
    const int peer = d >= me ? d + 1 : d;
    if (peer >= G) return;
    const int c = blockIdx.x % ctas_per_peer;
    const long long per_cta = blk / ctas_per_peer;          // elements
    // push: my block `peer` -> remote[peer] block `me`;  pull: remote[peer] block `me` -> my block `peer`
    double2 *rem = remote.p[peer] + (long long)me * blk + c * per_cta;
    double2 *loc = local + (long long)peer * blk + c * per_cta;
    const int t = threadIdx.x;
    if (TILE) {
        // the block as [nrows][512] elements (8 KiB rows); a tile = 512 rows x 8 elements, warp = 4 rows x 128 B,
        // consecutive k are 64 rows apart (as in the FFT tile: u + d * T_LINE); tiles are dealt round-robin to the CTAs
        double2 *remb = remote.p[peer] + (long long)me * blk;
        double2 *locb = local + (long long)peer * blk;
        const long long nrows = blk / 512, n_tiles = (nrows / 512) * 64;
        const int w = t & 7, u = t >> 3;
        for (long long tile = c; tile < n_tiles; tile += ctas_per_peer) {
            const long long rb = (tile / 64) * 512, cg = tile % 64;
            double2 v[R];
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const long long idx = (rb + u + k * 64) * 512 + cg * 8 + w;
                v[k] = PUSH ? locb[idx] : remb[idx];
            }
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const long long idx = (rb + u + k * 64) * 512 + cg * 8 + w;
                if (PUSH) remb[idx] = v[k]; else locb[idx] = v[k];
            }
        }
    } else {
        for (long long base = 0; base + 512ll * R <= per_cta; base += 512ll * R) {
            double2 v[R];
#pragma unroll
            for (int k = 0; k < R; ++k) v[k] = PUSH ? loc[base + k * 512 + t] : rem[base + k * 512 + t];
#pragma unroll
            for (int k = 0; k < R; ++k) {
                if (PUSH) rem[base + k * 512 + t] = v[k]; else loc[base + k * 512 + t] = v[k];
            }
        }
    }
}

int main(int argc, char **argv) {
    int G = 0;
    CK(cudaGetDeviceCount(&G));
    if (argc > 1) G = atoi(argv[1]) < G ? atoi(argv[1]) : G;
    const long long mb = argc > 2 ? atoll(argv[2]) : 2048;
    if (G < 2) { printf("{\"error\": \"needs >= 2 GPUs\"}\n"); return 0; }
    const long long total = mb * 1024 * 1024 / 16;            // double2 elements per GPU buffer
    const long long blk = total / G / (512 * 512) * (512 * 512);  // per-peer block: whole [512 rows][512 elements] groups
    std::vector<double2 *> src(G), dst(G);
    std::vector<cudaStream_t> st(G);
    std::vector<cudaEvent_t> e0(G), e1(G);
    for (int g = 0; g < G; ++g) {
        CK(cudaSetDevice(g));
        for (int p = 0; p < G; ++p)
            if (p != g) {
                cudaError_t e = cudaDeviceEnablePeerAccess(p, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { printf("{\"error\": \"no peer access %d->%d\"}\n", g, p); return 0; }
                cudaGetLastError();
            }
        CK(cudaMalloc(&src[g], total * 16));
        CK(cudaMalloc(&dst[g], total * 16));
        CK(cudaMemset(src[g], g + 1, total * 16));
        CK(cudaMemset(dst[g], 0, total * 16));
        CK(cudaStreamCreate(&st[g]));
        CK(cudaEventCreate(&e0[g]));
        CK(cudaEventCreate(&e1[g]));
    }
    const double bytes_out = (double)blk * 16 * (G - 1);
    auto run = [&](const char *name, auto kern, bool push, int ctas_per_peer) {
        Ptrs rp{};
        const int reps = 10;
        for (int it = 0; it < 2; ++it) {
            for (int g = 0; g < G; ++g) {
                CK(cudaSetDevice(g));
                for (int p = 0; p < G; ++p) rp.p[p] = push ? dst[p] : src[p];
                if (it == 1) CK(cudaEventRecord(e0[g], st[g]));
                for (int r = 0; r < (it == 0 ? 2 : reps); ++r)
                    kern<<<(G - 1) * ctas_per_peer, 512, 0, st[g]>>>(rp, push ? src[g] : dst[g], G, g, blk, ctas_per_peer);
                if (it == 1) CK(cudaEventRecord(e1[g], st[g]));
            }
            for (int g = 0; g < G; ++g) { CK(cudaSetDevice(g)); CK(cudaStreamSynchronize(st[g])); }
        }
        float worst = 0;
        for (int g = 0; g < G; ++g) {
            float ms = 0;
            CK(cudaSetDevice(g));
            CK(cudaEventElapsedTime(&ms, e0[g], e1[g]));
            worst = ms > worst ? ms : worst;
        }
        printf("{\"gpus\": %d, \"MiB_per_gpu\": %lld, \"path\": \"%s\", \"ctas\": %d, \"ms\": %.4f, \"GB/s_out_per_gpu\": %.1f}\n", G, mb, name,
               (G - 1) * ctas_per_peer, worst / reps, bytes_out / (worst / reps) / 1e6);
        fflush(stdout);
    };
    for (int cpp : {5, 10, 15, 21, 42, 84}) {         // x (G-1) peers: 35 .. 588 CTAs at G = 8 (148 SMs)
        const int c = cpp * 7 / (G - 1);
        run("push-tile (STG.128, 4 rows x 128 B per warp)", xfer<true, true, 8>, true, c);
        run("push-flat (STG.128, 512 B per warp)", xfer<true, false, 8>, true, c);
        run("pull-tile (LDG.128 from peer, 4 rows x 128 B per warp)", xfer<false, true, 8>, false, c);
        run("pull-flat (LDG.128 from peer, 512 B per warp)", xfer<false, false, 8>, false, c);
    }
    run("pull-flat R=16", xfer<false, false, 16>, false, 42 * 7 / (G - 1));
    run("push-flat R=16", xfer<true, false, 16>, true, 42 * 7 / (G - 1));
    return 0;
}
