// mixed_kernel.cuh — shared-memory FFT passes for the lengths the power-of-two tile kernels do not cover:
// L = 2^a 3^b 5^c 7^d 11^e 13^f (96, 100, 120, 360, 384, 480, 1000, 1001, 1536, 6000, ...) in ONE pass per axis, and axes
// too long for one tile (10^4 ... 4 * 10^7 points, powers of two included) in two.
//
// The power-of-two tile kernels (tile_kernel.cuh) keep R points per thread in registers across all stages and are tuned to
// the HBM roofline; this kernel is the general form of the same idea.  A CTA owns a tile of W lines of length L and runs
// one Stockham autosort stage per radix - the first reading global memory, the last writing it, the exchanges in between
// through shared memory - so an axis costs one HBM round trip instead of the generic path's gather + one global-memory
// pass per prime factor + scatter.  The radices are the in-register DFTs of radix_dft.cuh (2 ... 16, products of 2, 3, 5,
// 7, 11, 13), chosen at plan time so that a line needs as few stages as possible (1000 = 10 x 10 x 10, 384 = 6 x 8 x 8,
// 96 = 12 x 8); every stage has its own twiddle table, laid out so that a warp reads it as contiguous runs.  L, W and the
// radices are run-time values (no per-length instantiation).  The reference's CPU path covers these sizes with FFTW's
// n1_3 / n1_5 / n1_7 / ... codelets and its Cooley-Tukey solver (fftw-3.3.8/dft/ct.c, dft/scalar/codelets/); its own test
// shapes 3, 5, {3,2,2}, {3,3,2} (test/fft_test.rg:143,247,328,349) are of this kind.
//
// Stage s (radix P, Ns = product of the earlier radices, Lp = L / P), butterfly j in [0, Lp):
//     k = j mod Ns;   a_t = x[j + t Lp] * w_L^(t k L / (Ns P)),  t in [0, P);   y[(j - k) P + k + q Ns] = DFT_P(a)[q]
//
// Addressing and tiling are TileParams' (tile_kernel.cuh): in[o1*in_os1 + o2*in_os2 + i*in_is + l*in_ls].
//   ROWMAP = true   contiguous axis (in_ls == out_ls == 1): nfast threads run along a line, the other thread index
//                   picks the line; shared layout [w][pitch] (pitch odd: lanes of one warp may span several lines)
//   ROWMAP = false  strided axis (in_is == out_is == 1): the W adjacent lines are the fast thread index; shared layout
//                   [l][W]
#pragma once
#include "radix_dft.cuh"
#include "tile_kernel.cuh"

namespace fftb200 {

constexpr int MIXED_MAX_STAGES = 8;
constexpr int MIXED_MAX_THREADS = 256;

struct MixedStages {
    int n;          // number of stages
    int L, W;       // line length, lines per tile
    int nfast;      // ROWMAP: threads along a line; else W (the adjacent lines are the fast thread index)
    int nslow;      // blockDim.x / nfast
    int pitch;      // ROWMAP: shared-memory elements between consecutive lines of the tile
    int tw_o2;      // MIXED_TW: the column index n2 of the twiddle w_L^(k1 n2) is the tile's o2 index (else its line index)
    unsigned char r[MIXED_MAX_STAGES];   // radix of each stage, product = L
    int tw_off[MIXED_MAX_STAGES];        // stage s: table at tw + tw_off[s], entry [(t-1)*Ns + k] = w_L^(t k L / (Ns P))
    unsigned ns_m[MIXED_MAX_STAGES], ns_s[MIXED_MAX_STAGES];  // fast_div constants of Ns
};

// One Stockham stage of radix P over the tile.  The first stage reads its points straight from global memory and the last
// one stores straight to it (both coalesced: for a fixed t consecutive butterflies touch consecutive elements), so only
// the exchanges between stages go through shared memory.
//   SRC_G / DST_G: source / destination is global memory (strides in TileParams) instead of shared memory
//   IO (contiguous axis only): MIXED_C2C complex lines in and out;
//       MIXED_R2C  lines of L reals in (taken as complex with zero imaginary part), the first L/2+1 outputs stored;
//       MIXED_C2R  lines of L/2+1 complex in, completed to the full Hermitian line while loading, the backward transform's
//                  real part stored as L reals (strides of the real side are in real elements)
//       MIXED_R2C_HALF / MIXED_C2R_HALF  even L, rows starting on a pair of reals: the line of 2 Lh reals is taken as Lh
//                  complex z[m] = x[2m] + i x[2m+1]; one Lh-point complex transform (ms.L = Lh) and an even/odd pass over
//                  the pairs (k, Lh - k) in shared memory - after the last stage for R2C
//                      X[k] = E + w B,  X[Lh-k] = conj(E - w B),  E = (Z[k] + conj Z[Lh-k]) / 2,  B = -i (Z[k] - conj Z[Lh-k]) / 2
//                  and before the first stage for C2R (fftw-3.3.8/rdft/ct-hc2c.c:59-70 is the CPU path's form of it);
//                  half the arithmetic and shared-memory traffic of MIXED_R2C / MIXED_C2R.  TileParams::tw_aux = w_2Lh^k,
//                  k in [0, Lh/2]; the real side is addressed as complex pairs (strides halved by the plan)
//   Two-pass lines (L = N1 N2 too long for one tile; n = n1 N2 + n2, k = k1 + N1 k2; fftw-3.3.8/dft/ct.c is the CPU form):
//       MIXED_TW   strided pass over n1 whose store multiplies row k1, column n2 by w_L^(k1 n2) (two-level table
//                  TileParams::tw4_hi / tw4_lo, fp64), into the work buffer
//       MIXED_RC   contiguous pass over n2 of the work buffer's rows; the last stage stays in shared memory and a final
//                  pass with the adjacent lines as the fast thread index stores out[k2 N1 + k1] in 128-byte runs
enum { MIXED_C2C = 0, MIXED_R2C = 1, MIXED_C2R = 2, MIXED_R2C_HALF = 3, MIXED_C2R_HALF = 4, MIXED_TW = 5, MIXED_RC = 6 };

template <typename T, int P, bool ROWMAP, bool SRC_G, bool DST_G, int IO>
__device__ __forceinline__ void mixed_stage(const TileParams &p, const MixedStages &ms, const void *__restrict__ gin,
                                            void *__restrict__ gout, const cplx<T> *__restrict__ ssrc,
                                            cplx<T> *__restrict__ sdst, const cplx<T> *__restrict__ tws, const int Ns,
                                            const unsigned nm, const unsigned nsh, const int i0, const int f, const int sl,
                                            const unsigned cmask, const int o2) {
    using C = cplx<T>;
    const int Lp = ms.L / P;
    const int W = ms.W;
    // ROWMAP: f runs over the butterflies of a line, sl over the lines; otherwise f is the line, sl the butterfly
    const int j0 = ROWMAP ? f : sl, jstep = ROWMAP ? ms.nfast : ms.nslow;
    const int w0 = ROWMAP ? sl : f, wstep = ROWMAP ? ms.nslow : W;
    const int s_line = ROWMAP ? ms.pitch : 1, s_elem = ROWMAP ? 1 : W;  // shared index = w * s_line + l * s_elem
    for (int j = j0; j < Lp; j += jstep) {
        const int q = fast_div(j, nm, nsh), k = j - q * Ns;
        const int ob = j + q * Ns * (P - 1);
        for (int w = w0; w < W; w += wstep) {
            C a[P];
            if (SRC_G) {
                const int wi = min(i0 + w, p.n_inner - 1);  // (a ragged last tile re-reads its last valid line; never stored)
                if (IO == MIXED_R2C) {
                    const T *g = reinterpret_cast<const T *>(gin) + (long long)wi * p.in_is + j;
#pragma unroll
                    for (int t = 0; t < P; ++t) a[t] = mk<T>(__ldg(g + (unsigned)(t * Lp)), (T)0);
                } else if (IO == MIXED_C2R) {
                    // conj(X_full[l]): X_full[l] = X[l] for l <= L/2, conj(X[L - l]) above (conj in, conj out = backward)
                    const C *g = reinterpret_cast<const C *>(gin) + (long long)wi * p.in_is;
                    const int half = ms.L >> 1;
#pragma unroll
                    for (int t = 0; t < P; ++t) {
                        const int l = j + t * Lp;
                        const bool low = l <= half;
                        const C v = __ldg(g + (unsigned)(low ? l : ms.L - l));
                        a[t] = low ? mk<T>(v.x, -v.y) : v;
                    }
                } else if (ROWMAP) {  // unit element stride: 32-bit offsets from the line's base
                    const C *g = reinterpret_cast<const C *>(gin) + (long long)wi * p.in_is + j;
#pragma unroll
                    for (int t = 0; t < P; ++t) a[t] = conj_if(__ldg(g + (unsigned)(t * Lp)), cmask);
                } else {
                    // (the plan guarantees (L - 1) * in_ls < 2^32: 32-bit offsets from the line's base)
                    const C *g = reinterpret_cast<const C *>(gin) + wi;
                    const unsigned ls = (unsigned)p.in_ls;
#pragma unroll
                    for (int t = 0; t < P; ++t) a[t] = conj_if(__ldg(g + (unsigned)(j + t * Lp) * ls), cmask);
                }
            } else {
                const C *s = ssrc + w * s_line + j * s_elem;
#pragma unroll
                for (int t = 0; t < P; ++t) a[t] = s[t * Lp * s_elem];
            }
            if (Ns > 1) {
                // (byte offsets: one 32 x 32 + 64-bit multiply-add per address)
                const char *twk = reinterpret_cast<const char *>(tws + k);
                const unsigned stepb = (unsigned)Ns * (unsigned)sizeof(C);
#pragma unroll
                for (int t = 1; t < P; ++t)
                    a[t] = cmul(a[t], __ldg(reinterpret_cast<const C *>(twk + (unsigned long long)((unsigned)(t - 1) * stepb))));
            }
            Dft<T, P>::run(a);
            if (DST_G) {
                if (i0 + w < p.n_inner) {
                    if (IO == MIXED_R2C) {
                        C *g = reinterpret_cast<C *>(gout) + (long long)(i0 + w) * p.out_is;
                        const int half = ms.L >> 1;
#pragma unroll
                        for (int t = 0; t < P; ++t)
                            if (ob + t * Ns <= half) g[(unsigned)(ob + t * Ns)] = a[t];
                    } else if (IO == MIXED_C2R) {
                        T *g = reinterpret_cast<T *>(gout) + (long long)(i0 + w) * p.out_is + ob;
#pragma unroll
                        for (int t = 0; t < P; ++t) g[(unsigned)(t * Ns)] = a[t].x;
                    } else if (ROWMAP) {
                        C *g = reinterpret_cast<C *>(gout) + (long long)(i0 + w) * p.out_is + ob;
#pragma unroll
                        for (int t = 0; t < P; ++t) g[(unsigned)(t * Ns)] = conj_if(a[t], cmask);
                    } else {
                        C *g = reinterpret_cast<C *>(gout) + (i0 + w);
                        const unsigned ls = (unsigned)p.out_ls;
                        if (IO == MIXED_TW) {
                            const unsigned col = ms.tw_o2 ? (unsigned)o2 : (unsigned)(i0 + w);
#pragma unroll
                            for (int t = 0; t < P; ++t) {
                                const unsigned m = (unsigned)(ob + t * Ns) * col;  // < L
                                const double2 wh = __ldg(p.tw4_hi + (m >> p.tw4_shift)), wl = __ldg(p.tw4_lo + (m & p.tw4_mask));
                                const C wv = mk<T>((T)(wh.x * wl.x - wh.y * wl.y), (T)(wh.x * wl.y + wh.y * wl.x));
                                a[t] = cmul(a[t], wv);
                            }
                        }
#pragma unroll
                        for (int t = 0; t < P; ++t) g[(unsigned)(ob + t * Ns) * ls] = conj_if(a[t], cmask);
                    }
                }
            } else {
                C *d = sdst + w * s_line + ob * s_elem;
#pragma unroll
                for (int t = 0; t < P; ++t) d[t * Ns * s_elem] = a[t];
            }
        }
    }
}

template <typename T, bool ROWMAP, int MAXR, int IO>
__global__ void __launch_bounds__(MIXED_MAX_THREADS) fft_mixed_kernel(const TileParams p, const MixedStages ms) {
    using C = cplx<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = ms.L, W = ms.W;
    const size_t buf_elems = ROWMAP ? (size_t)W * ms.pitch : (size_t)W * L;
    C *buf0 = reinterpret_cast<C *>(smem_raw);
    C *buf1 = buf0 + buf_elems;  // (present when the line needs three stages or more)
    const int tile = (int)blockIdx.x;
    const int o = fast_div(tile, p.div_tpo_m, p.div_tpo_s);
    const int i0 = (tile - o * p.tiles_per_outer) * W;
    const int o1 = fast_div(o, p.div_o2_m, p.div_o2_s), o2 = o - o1 * p.n_o2;
    static_assert(IO == MIXED_C2C || (IO == MIXED_TW ? !ROWMAP : ROWMAP), "real transforms and the transposing pass: contiguous axis; twiddled store: strided axis");
    // (the real side of a real transform is addressed in real elements)
    const void *__restrict__ gin = reinterpret_cast<const char *>(p.in) +
                                   (o1 * p.in_os1 + o2 * p.in_os2) * (long long)(IO == MIXED_R2C ? sizeof(T) : sizeof(C));
    void *__restrict__ gout = reinterpret_cast<char *>(p.out) +
                              (o1 * p.out_os1 + o2 * p.out_os2) * (long long)(IO == MIXED_C2R ? sizeof(T) : sizeof(C));
    const C *__restrict__ tw = reinterpret_cast<const C *>(p.tw);
    const unsigned cmask = (((IO == MIXED_C2C || IO == MIXED_TW || IO == MIXED_RC) && p.inverse) || IO == MIXED_C2R_HALF) ? 0x80000000u : 0u;
    int sl = (int)threadIdx.x / ms.nfast;
    const int f = (int)threadIdx.x - sl * ms.nfast;
    if (sl >= ms.nslow) sl = 1 << 30;  // threads past nfast * nslow (block rounded up to whole warps) only take part in the barriers

    const C *__restrict__ twx = reinterpret_cast<const C *>(p.tw_aux);  // half-length real modes: w_2L^k
    const int nfast_line = ms.nfast;
    C *src = buf0, *dst = buf0;
    if (IO == MIXED_C2R_HALF) {
        // even/odd pre-pass: half spectrum X[0 .. L] (L = ms.L = half the real length) -> conj(Z'), Z'[k] = Ee + i Oo with
        // Ee = X[k] + conj X[L-k], Oo = (X[k] - conj X[L-k]) conj(w^k); Z'[L-k] = conj(Ee) + i conj(Oo)
        for (int w = sl; w < W; w += ms.nslow) {
            const int wi = min(i0 + w, p.n_inner - 1);
            const C *g = reinterpret_cast<const C *>(gin) + (long long)wi * p.in_is;
            C *d = buf0 + w * ms.pitch;
            for (int k = f; k <= L / 2; k += nfast_line) {
                const C xa = __ldg(g + k), xb = __ldg(g + (L - k));
                const C wk = __ldg(twx + k);                       // w^k = c - i s  (wk.x = c, wk.y = -s)
                const C ee = mk<T>(xa.x + xb.x, xa.y - xb.y);      // X[k] + conj X[L-k]
                const C dd = mk<T>(xa.x - xb.x, xa.y + xb.y);      // X[k] - conj X[L-k]
                const C oo = mk<T>(dd.x * wk.x + dd.y * wk.y, dd.y * wk.x - dd.x * wk.y);  // dd * conj(w^k)
                // Z'[k] = ee + i oo = (ee.x - oo.y, ee.y + oo.x); stored conjugated (backward = conj F conj)
                d[k] = mk<T>(ee.x - oo.y, -(ee.y + oo.x));
                // Z'[L-k] = conj(ee) + i conj(oo) = (ee.x + oo.y, -ee.y + oo.x); conjugated
                if (k > 0) d[L - k] = mk<T>(ee.x + oo.y, ee.y - oo.x);
            }
        }
        __syncthreads();
        dst = buf1;
    }
    int Ns = 1;
    for (int s = 0; s < ms.n; ++s) {
        const int P = ms.r[s];
        const C *tws = tw + ms.tw_off[s];
        const unsigned nm = ms.ns_m[s], nsh = ms.ns_s[s];
        const bool first = s == 0, last = s == ms.n - 1;
        const bool src_g = first && IO != MIXED_C2R_HALF, dst_g = last && IO != MIXED_R2C_HALF && IO != MIXED_RC;
#define FFTB200_MIXED_CASE(R)                                                                                        \
    case R:                                                                                                          \
        if constexpr (R <= MAXR) {                                                                                   \
            if (src_g && dst_g) mixed_stage<T, R, ROWMAP, true, true, IO>(p, ms, gin, gout, src, dst, tws, Ns, nm, nsh, i0, f, sl, cmask, o2);   \
            else if (src_g) mixed_stage<T, R, ROWMAP, true, false, IO>(p, ms, gin, gout, src, dst, tws, Ns, nm, nsh, i0, f, sl, cmask, o2);  \
            else if (dst_g) mixed_stage<T, R, ROWMAP, false, true, IO>(p, ms, gin, gout, src, dst, tws, Ns, nm, nsh, i0, f, sl, cmask, o2);  \
            else mixed_stage<T, R, ROWMAP, false, false, IO>(p, ms, gin, gout, src, dst, tws, Ns, nm, nsh, i0, f, sl, cmask, o2);            \
        }                                                                                                            \
        break;
        switch (P) {
            FFTB200_MIXED_CASE(2) FFTB200_MIXED_CASE(3) FFTB200_MIXED_CASE(4) FFTB200_MIXED_CASE(5) FFTB200_MIXED_CASE(6)
            FFTB200_MIXED_CASE(7) FFTB200_MIXED_CASE(8) FFTB200_MIXED_CASE(9) FFTB200_MIXED_CASE(10) FFTB200_MIXED_CASE(11)
            FFTB200_MIXED_CASE(12) FFTB200_MIXED_CASE(13) FFTB200_MIXED_CASE(14) FFTB200_MIXED_CASE(15) FFTB200_MIXED_CASE(16)
            default: break;
        }
#undef FFTB200_MIXED_CASE
        if (first && p.prefetch_tiles > 0) {
            // L2 prefetch of the tile that will run next in this CTA slot (this tile's own loads have been consumed): the
            // next CTA's first stage then waits for L2, not for HBM
            const int ft = tile + p.prefetch_tiles;
            if (ft < p.n_tiles) {
                const int fo = fast_div(ft, p.div_tpo_m, p.div_tpo_s);
                const int fi0 = (ft - fo * p.tiles_per_outer) * W;
                const int fo1 = fast_div(fo, p.div_o2_m, p.div_o2_s), fo2 = fo - fo1 * p.n_o2;
                if (fi0 + W <= p.n_inner) {  // whole tiles only: never touch addresses past the array
                    constexpr int IN_ELT = IO == MIXED_R2C ? (int)sizeof(T) : (int)sizeof(C);
                    const char *fin = reinterpret_cast<const char *>(p.in) +
                                      (fo1 * p.in_os1 + fo2 * p.in_os2 + (long long)fi0 * p.in_is) * IN_ELT;
                    const int n_run = ROWMAP ? W : L;                                   // contiguous runs of the tile
                    const int run_elems = ROWMAP ? (IO == MIXED_C2R ? L / 2 + 1 : (IO == MIXED_C2R_HALF ? L + 1 : L)) : W;
                    const int ch_run = (run_elems * IN_ELT + 127) / 128;                 // 128-byte chunks per run
                    const long long run_stride = ROWMAP ? p.in_is : p.in_ls;
                    for (int ch = (int)threadIdx.x; ch < n_run * ch_run; ch += (int)blockDim.x) {
                        const int r = ch / ch_run, cc = ch - r * ch_run;
                        const char *a = fin + (long long)r * run_stride * IN_ELT + cc * 128;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
                    }
                }
            }
        }
        Ns *= P;
        if (!dst_g) {
            __syncthreads();
            // stage s wrote dst; the next one reads it and writes the other buffer
            src = dst;
            dst = (dst == buf0) ? buf1 : buf0;
        }
    }
    if (IO == MIXED_RC) {
        // transposing store: the W adjacent lines are the fast thread index (out_is == 1), element l of a line goes
        // out_ls further; the odd line pitch keeps the shared-memory reads free of bank conflicts
        const int fw = (int)threadIdx.x % W, fl = (int)threadIdx.x / W, nl = (int)blockDim.x / W;
        if (fl < nl && i0 + fw < p.n_inner) {
            C *g = reinterpret_cast<C *>(gout) + (i0 + fw);
            const C *z = src + fw * ms.pitch;
            const unsigned ls = (unsigned)p.out_ls;
            for (int l = fl; l < L; l += nl) g[(unsigned)l * ls] = conj_if(z[l], cmask);
        }
    }
    if (IO == MIXED_R2C_HALF) {
        // even/odd post-pass over the pairs (k, L-k) of Z (in src): X[k] = E + w^k B, X[L-k] = conj(E - w^k B)
        for (int w = sl; w < W; w += ms.nslow) {
            if (i0 + w >= p.n_inner) continue;
            const C *z = src + w * ms.pitch;
            C *g = reinterpret_cast<C *>(gout) + (long long)(i0 + w) * p.out_is;
            for (int k = f; k <= L / 2; k += nfast_line) {
                const C za = z[k], zb = z[k == 0 ? 0 : L - k];
                const C wk = __ldg(twx + k);                                      // c - i s
                const C e = mk<T>((T)0.5 * (za.x + zb.x), (T)0.5 * (za.y - zb.y));  // (Z[k] + conj Z[L-k]) / 2
                const C b = mk<T>((T)0.5 * (za.y + zb.y), (T)-0.5 * (za.x - zb.x));  // -i (Z[k] - conj Z[L-k]) / 2
                const C wb = cmul(wk, b);
                g[k] = mk<T>(e.x + wb.x, e.y + wb.y);
                g[L - k] = mk<T>(e.x - wb.x, -(e.y - wb.y));
            }
        }
    }
}

// host side: the instantiations (precision x mapping x largest radix compiled in: 8, 10 or 16 - the register count follows
// the largest in-register DFT); radices the kernel has code for
typedef void (*MixedKernelFn)(const TileParams, const MixedStages);
// (holding the register allocation to 80 or 64 with a minimum-blocks launch bound was measured and is not a gain:
// 384^3 fp64 1.62 -> 1.72 ms, 1536^2 0.26 -> 0.34 ms; 1000-point batches unchanged)
MixedKernelFn mixed_kernel(int prec, bool rowmap, int maxr, int io);
template <typename T, int MAXR> MixedKernelFn mixed_kernel_inst(bool rowmap, int io);
constexpr int MIXED_RADICES[] = {16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2};

}  // namespace fftb200
