// tma_probe.cu — data-movement ceilings for the strided-axis tiles of a 3-D transform (tools only, not product code).
//
// Moves a complex64 array [n0][n1][n2] tile by tile exactly the way a strided-axis FFT pass does (tile = L rows of
// W = 8 adjacent columns = 128 bytes per row, rows `stride` apart), without any arithmetic, through four data paths:
//   flat      grid-stride LDG.128 / STG.128 over the flat array (the ceiling at this working-set size)
//   ldg       one tile per CTA: LDG.128 -> registers -> STG.128 (what fft_tile_kernel does)
//   tma       persistent CTAs, one thread: cp.async.bulk.tensor box loads into an NBUF ring, TMA stores out of it
//   tma_stg   persistent CTAs: a producer thread issues the box loads, 512 consumer threads read the landing
//             buffer (LDS.128) and store with STG.128 (the data path of a TMA-fed FFT pass)
// Prints GB/s (read + written bytes / time) per configuration as JSON lines.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiled get_encode() {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn) { fprintf(stderr, "no cuTensorMapEncodeTiled\n"); exit(1); }
    return (EncodeTiled)fn;
}

constexpr int W = 8;          // complex64 per row of a tile: 128 bytes
constexpr int BOX_ROWS = 256;  // rows per TMA box (boxDim <= 256)

struct Geo {
    long long n0, n1, n2;
    int axis;  // 0: lines along n0 (stride n1*n2), 1: lines along n1 (stride n2)
    int L;
    long long ls;       // line stride in elements
    int tiles_x;        // n2 / W
    long long n_tiles;  // other * tiles_x
};

__device__ __forceinline__ void tile_origin(const Geo &g, long long tile, long long &base, int &o, int &xb) {
    o = (int)(tile / g.tiles_x);
    xb = (int)(tile - (long long)o * g.tiles_x);
    base = (g.axis == 0 ? (long long)o * g.n2 : (long long)o * g.n1 * g.n2) + (long long)xb * W;
}

__global__ void __launch_bounds__(512) flat_copy(const double2 *__restrict__ in, double2 *__restrict__ out, long long n) {
    for (long long i = blockIdx.x * 512ll + threadIdx.x; i < n; i += (long long)gridDim.x * 512) out[i] = in[i];
}

template <int R, int MINB>
__global__ void __launch_bounds__(512, MINB) ldg_copy(const double2 *__restrict__ in, double2 *__restrict__ out, Geo g) {
    const int t = threadIdx.x, w = t & (W - 1), u = t >> 3;  // 64 row slots
    long long base;
    int o, xb;
    tile_origin(g, blockIdx.x, base, o, xb);
    double2 v[R];
#pragma unroll
    for (int d = 0; d < R; ++d) v[d] = __ldg(in + base + (long long)(u + d * 64) * g.ls + w);
#pragma unroll
    for (int d = 0; d < R; ++d) out[base + (long long)(u + d * 64) * g.ls + w] = v[d];
}

// narrower tiles: WW complex64 per row (64- or 32-byte row segments), 64 row slots, WW * 64 threads
template <int R, int WW, int MINB>
__global__ void __launch_bounds__(WW * 64, MINB) ldg_copy_narrow(const double2 *__restrict__ in, double2 *__restrict__ out, Geo g) {
    const int t = threadIdx.x, w = t % WW, u = t / WW;
    const long long tiles_x = g.n2 / WW;
    const long long o = blockIdx.x / tiles_x, xb = blockIdx.x - o * tiles_x;
    const long long base = (g.axis == 0 ? o * g.n2 : o * g.n1 * g.n2) + xb * WW;
    double2 v[R];
#pragma unroll
    for (int d = 0; d < R; ++d) v[d] = __ldg(in + base + (long long)(u + d * 64) * g.ls + w);
#pragma unroll
    for (int d = 0; d < R; ++d) out[base + (long long)(u + d * 64) * g.ls + w] = v[d];
}

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n .reg .pred p;\n W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_%=;\n bra W_%=;\n D_%=:\n}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap *tm, int c0, int c1, int c2, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tm, int c0, int c1, int c2, unsigned src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm), "r"(c0), "r"(c1), "r"(c2),
                 "r"(src)
                 : "memory");
}

// tma: one thread per CTA runs the whole pipeline (loads into an NBUF ring, stores out of it)
template <int NBUF>
__global__ void __launch_bounds__(32) tma_copy(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, Geo g) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[NBUF];
    if (threadIdx.x != 0) return;
    const unsigned tile_bytes = (unsigned)g.L * W * 16u;
    const unsigned sm0 = (unsigned)__cvta_generic_to_shared(smem);
    unsigned bar[NBUF];
    for (int b = 0; b < NBUF; ++b) {
        bar[b] = (unsigned)__cvta_generic_to_shared(&bars[b]);
        mbar_init(bar[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int boxes = g.L / BOX_ROWS;
    auto issue_load = [&](long long tile, int b) {
        const int o = (int)(tile / g.tiles_x), xb = (int)(tile - (long long)o * g.tiles_x);
        mbar_expect_tx(bar[b], tile_bytes);
        for (int j = 0; j < boxes; ++j) {
            const unsigned dst = sm0 + (unsigned)b * tile_bytes + (unsigned)j * BOX_ROWS * W * 16u;
            if (g.axis == 0) tma_load_3d(dst, &tin, xb * W * 2, o, j * BOX_ROWS, bar[b]);
            else tma_load_3d(dst, &tin, xb * W * 2, j * BOX_ROWS, o, bar[b]);
        }
    };
    long long k_issue = 0, n_mine = 0;
    for (long long t = blockIdx.x; t < g.n_tiles; t += gridDim.x) ++n_mine;
    for (; k_issue < NBUF - 1 && k_issue < n_mine; ++k_issue) issue_load(blockIdx.x + k_issue * gridDim.x, (int)(k_issue % NBUF));
    for (long long k = 0; k < n_mine; ++k) {
        const int b = (int)(k % NBUF);
        mbar_wait(bar[b], (unsigned)((k / NBUF) & 1));
        const long long tile = blockIdx.x + k * gridDim.x;
        const int o = (int)(tile / g.tiles_x), xb = (int)(tile - (long long)o * g.tiles_x);
        for (int j = 0; j < boxes; ++j) {
            const unsigned src = sm0 + (unsigned)b * tile_bytes + (unsigned)j * BOX_ROWS * W * 16u;
            if (g.axis == 0) tma_store_3d(&tout, xb * W * 2, o, j * BOX_ROWS, src);
            else tma_store_3d(&tout, xb * W * 2, j * BOX_ROWS, o, src);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (k_issue < n_mine) {
            // the buffer tile k_issue lands in was stored by group k - 1: wait until it has been read
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            issue_load(blockIdx.x + k_issue * gridDim.x, (int)(k_issue % NBUF));
            ++k_issue;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// tma_stg: producer thread + 512 consumers (LDS.128 -> STG.128)
template <int NBUF, int R>
__global__ void __launch_bounds__(544, 1) tma_stg_copy(const __grid_constant__ CUtensorMap tin, double2 *__restrict__ out, Geo g) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[2 * NBUF];
    const unsigned tile_bytes = (unsigned)g.L * W * 16u;
    const unsigned sm0 = (unsigned)__cvta_generic_to_shared(smem);
    unsigned full[NBUF], empty[NBUF];
    for (int b = 0; b < NBUF; ++b) {
        full[b] = (unsigned)__cvta_generic_to_shared(&bars[b]);
        empty[b] = (unsigned)__cvta_generic_to_shared(&bars[NBUF + b]);
    }
    if (threadIdx.x == 0) {
        for (int b = 0; b < NBUF; ++b) { mbar_init(full[b], 1); mbar_init(empty[b], 512); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long n_mine = 0;
    for (long long t = blockIdx.x; t < g.n_tiles; t += gridDim.x) ++n_mine;
    const int boxes = g.L / BOX_ROWS;
    if (threadIdx.x >= 512) {
        if (threadIdx.x != 512) return;
        for (long long k = 0; k < n_mine; ++k) {
            const int b = (int)(k % NBUF);
            if (k >= NBUF) mbar_wait(empty[b], (unsigned)(((k / NBUF) - 1) & 1));
            const long long tile = blockIdx.x + k * gridDim.x;
            const int o = (int)(tile / g.tiles_x), xb = (int)(tile - (long long)o * g.tiles_x);
            mbar_expect_tx(full[b], tile_bytes);
            for (int j = 0; j < boxes; ++j) {
                const unsigned dst = sm0 + (unsigned)b * tile_bytes + (unsigned)j * BOX_ROWS * W * 16u;
                if (g.axis == 0) tma_load_3d(dst, &tin, xb * W * 2, o, j * BOX_ROWS, full[b]);
                else tma_load_3d(dst, &tin, xb * W * 2, j * BOX_ROWS, o, full[b]);
            }
        }
        return;
    }
    const int t = threadIdx.x, w = t & (W - 1), u = t >> 3;
    for (long long k = 0; k < n_mine; ++k) {
        const int b = (int)(k % NBUF);
        mbar_wait(full[b], (unsigned)((k / NBUF) & 1));
        const double2 *sm = reinterpret_cast<const double2 *>(smem + (size_t)b * tile_bytes);
        double2 v[R];
#pragma unroll
        for (int d = 0; d < R; ++d) v[d] = sm[(u + d * 64) * W + w];
        mbar_arrive(empty[b]);
        long long base;
        int o, xb;
        tile_origin(g, blockIdx.x + k * gridDim.x, base, o, xb);
#pragma unroll
        for (int d = 0; d < R; ++d) out[base + (long long)(u + d * 64) * g.ls + w] = v[d];
    }
}

static CUtensorMap make_map(EncodeTiled enc, void *ptr, const Geo &g) {
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)g.n2 * 2, (cuuint64_t)g.n1, (cuuint64_t)g.n0};
    cuuint64_t strides[2] = {(cuuint64_t)g.n2 * 16, (cuuint64_t)g.n1 * g.n2 * 16};
    cuuint32_t box[3] = {W * 2, g.axis == 1 ? (cuuint32_t)BOX_ROWS : 1u, g.axis == 0 ? (cuuint32_t)BOX_ROWS : 1u};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

template <typename F> static float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

static void check(const double2 *in, const double2 *out, long long n, const char *what) {
    // spot check: 1M samples
    std::vector<double2> a(4096), b(4096);
    for (int s = 0; s < 16; ++s) {
        const long long off = (n / 16) * s + 12345 % (n / 16 - 4096);
        CK(cudaMemcpy(a.data(), in + off, 4096 * 16, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), out + off, 4096 * 16, cudaMemcpyDeviceToHost));
        if (memcmp(a.data(), b.data(), 4096 * 16)) { printf("{\"error\": \"%s copy mismatch at %lld\"}\n", what, off); return; }
    }
}

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: tma_probe n0 n1 n2 axis [reps]\n"); return 2; }
    Geo g;
    g.n0 = atoll(argv[1]); g.n1 = atoll(argv[2]); g.n2 = atoll(argv[3]); g.axis = atoi(argv[4]);
    const int reps = argc > 5 ? atoi(argv[5]) : 5;
    g.L = (int)(g.axis == 0 ? g.n0 : g.n1);
    g.ls = g.axis == 0 ? g.n1 * g.n2 : g.n2;
    g.tiles_x = (int)(g.n2 / W);
    g.n_tiles = (g.axis == 0 ? g.n1 : g.n0) * g.tiles_x;
    const long long n = g.n0 * g.n1 * g.n2;
    const double gb = 2.0 * n * 16 / 1e9;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    double2 *in, *out;
    CK(cudaMalloc(&in, n * 16));
    CK(cudaMalloc(&out, n * 16));
    CK(cudaMemset(in, 0x3c, n * 16));
    {   // recognisable content
        std::vector<double> h(1 << 20);
        for (size_t i = 0; i < h.size(); ++i) h[i] = (double)i * 1.25;
        for (long long off = 0; off + (long long)h.size() * 8 <= n * 16; off += (n * 16 / 64) / 8 * 8)
            CK(cudaMemcpy((char *)in + off, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
    }
    EncodeTiled enc = get_encode();
    const CUtensorMap tin = make_map(enc, in, g), tout = make_map(enc, out, g);
    const char *hdr = "{\"shape\": [%lld, %lld, %lld], \"axis\": %d, \"L\": %d, \"stride_bytes\": %lld, \"path\": \"%s\", \"ms\": %.4f, \"GB/s\": %.1f}\n";
    auto report = [&](const char *path, float ms) {
        printf(hdr, g.n0, g.n1, g.n2, g.axis, g.L, g.ls * 16, path, ms, gb / ms * 1e3 / 1e3 * 1e0);
        fflush(stdout);
    };
    report("flat LDG/STG grid-stride", time_ms([&] { flat_copy<<<sms * 4, 512>>>(in, out, n); }, reps));
    if (g.L == 512) {
        CK(cudaMemset(out, 0, n * 16));
        report("ldg tile R=8 2CTA/SM", time_ms([&] { ldg_copy<8, 2><<<(unsigned)g.n_tiles, 512>>>(in, out, g); }, reps));
        check(in, out, n, "ldg");
    } else if (g.L == 1024) {
        CK(cudaMemset(out, 0, n * 16));
        report("ldg tile R=16 1CTA/SM", time_ms([&] { ldg_copy<16, 1><<<(unsigned)g.n_tiles, 512>>>(in, out, g); }, reps));
        check(in, out, n, "ldg");
    }
    if (g.L == 1024) {
        CK(cudaMemset(out, 0, n * 16));
        report("ldg tile W=4 (64 B rows) R=16 2CTA/SM", time_ms([&] { ldg_copy_narrow<16, 4, 2><<<(unsigned)(g.n_tiles * 2), 256>>>(in, out, g); }, reps));
        check(in, out, n, "ldg w4");
        CK(cudaMemset(out, 0, n * 16));
        report("ldg tile W=2 (32 B rows) R=16 4CTA/SM", time_ms([&] { ldg_copy_narrow<16, 2, 4><<<(unsigned)(g.n_tiles * 4), 128>>>(in, out, g); }, reps));
        check(in, out, n, "ldg w2");
    }
    const unsigned tile_bytes = (unsigned)g.L * W * 16u;
    auto run_tma = [&](auto kern, int nbuf, int ctas_per_sm, const char *name) {
        const size_t smem = (size_t)nbuf * tile_bytes;
        if (smem * ctas_per_sm > 220 * 1024) return;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaMemset(out, 0, n * 16));
        char buf[128];
        snprintf(buf, sizeof buf, "%s NBUF=%d %dCTA/SM", name, nbuf, ctas_per_sm);
        report(buf, time_ms([&] { kern<<<sms * ctas_per_sm, 32, smem>>>(tin, tout, g); }, reps));
        check(in, out, n, buf);
    };
    run_tma(tma_copy<2>, 2, 1, "tma ld+st");
    run_tma(tma_copy<3>, 3, 1, "tma ld+st");
    run_tma(tma_copy<2>, 2, 2, "tma ld+st");
    run_tma(tma_copy<3>, 3, 2, "tma ld+st");
    run_tma(tma_copy<4>, 4, 1, "tma ld+st");
    auto run_stg = [&](auto kern, int nbuf, const char *name) {
        const size_t smem = (size_t)nbuf * tile_bytes;
        if (smem > 220 * 1024) return;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaMemset(out, 0, n * 16));
        char buf[128];
        snprintf(buf, sizeof buf, "%s NBUF=%d 1CTA/SM", name, nbuf);
        report(buf, time_ms([&] { kern<<<sms, 544, smem>>>(tin, out, g); }, reps));
        check(in, out, n, buf);
    };
    if (g.L == 512) {
        run_stg(tma_stg_copy<2, 8>, 2, "tma ld, LDS->STG");
        run_stg(tma_stg_copy<3, 8>, 3, "tma ld, LDS->STG");
    } else if (g.L == 1024) {
        run_stg(tma_stg_copy<1, 16>, 1, "tma ld, LDS->STG");
    }
    return 0;
}
