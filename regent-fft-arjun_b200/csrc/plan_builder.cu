// plan_builder.cu — turns cufftPlanMany-style arguments into a list of kernel launches (include/fft_b200.h).
//
// This is the layer that replaces cuFFT behind Regent-FFT's GPU branch
// (reference src/fft.rg:233-242, 389-398, 571-580, 638).  A plan is a short list of kernel
// launches ("passes"); every pass is one HBM round trip over one axis:
//
//   3-D C2C  [n0][n1][n2] : ROW pass over n2 (in -> out), COL pass over n1, COL pass over n0 (in place)
//   R2C                   : the first pass is the fused half-length FFT + even/odd post-pass,
//                           later passes run over n_last/2+1 columns
//   1-D N > one tile      : four-step, N = N1*N2(*N3): COL+twiddle pass(es), then a transposing
//                           ROW->COL pass, through the plan's work buffer
//   anything else         : generic global-memory path (generic_kernels.cuh)
//
// No CPU fallback exists: if a kernel cannot be launched the call returns an error code.
#include <algorithm>
#include <functional>
#include <tuple>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "plan_internal.h"

namespace fftb200 {

// ------------------------------------------------------------------------------------------
// twiddles: w_n^m = exp(-2*pi*i*m/n), correctly rounded to within an ulp or so for every m (the accuracy
// contract of SURVEY.md §8 a10: FFTW's tables come from an argument reduced to [0, pi/4] before libm sees it,
// fftw-3.3.8/kernel/trig.c:57-80).  Here: the angle is m/n of a turn, an exact rational; count whole eighths of
// a turn in integers (q = floor(8m/n), r = 8m - q*n), fold odd eighths back (r -> n - r) so the remaining angle
// phi = (pi/4) * r/n lies in [0, pi/4], evaluate cos/sin of phi in long double, and place them by the table below.
// ------------------------------------------------------------------------------------------
static void twiddle(long long m, long long n, double *re, double *im) {
    static const long double QUARTER_PI = 0.78539816339744830961566084581987572104929234984378L;
    m %= n;
    if (m < 0) m += n;
    const long long q = (8 * m) / n;        // which eighth of the turn
    long long r = 8 * m - q * n;            // position inside it, in units of 1/(8n) turn
    if (q & 1) r = n - r;                   // odd eighths are measured back from their upper end
    const long double phi = QUARTER_PI * (long double)r / (long double)n;
    const long double c = cosl(phi), s = sinl(phi);
    // (cos, sin) of the full angle for eighth q, from (c, s) of the folded angle
    static const int cos_from_s[8] = {0, 1, 1, 0, 0, 1, 1, 0};
    static const int cos_sign[8] = {+1, +1, -1, -1, -1, -1, +1, +1};
    static const int sin_sign[8] = {+1, +1, +1, +1, -1, -1, -1, -1};
    const long double co = cos_sign[q] * (cos_from_s[q] ? s : c);
    const long double si = sin_sign[q] * (cos_from_s[q] ? c : s);
    *re = (double)co;
    *im = (double)(-si);  // forward transform: exp(-i theta)
}

int env_int_or(const char *name, int dflt) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// ------------------------------------------------------------------------------------------
// device tables
// ------------------------------------------------------------------------------------------
void *Builder::upload(const void *host, size_t bytes) {
    void *d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) { err = FFTB200_ALLOC_FAILED; cudaGetLastError(); return nullptr; }
    P->dev_allocs.push_back(d);
    if (cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { err = FFTB200_SETUP_FAILED; cudaGetLastError(); return nullptr; }
    return d;
}

void *Builder::table(long long n, long long count, bool force_double, long long step) {
    const int kind = (force_double ? 1 : 0) + (step != 1 ? 2 : 0);
    auto key = std::make_pair(std::make_pair(kind, n), count);
    if (step == 1) {
        auto it = cache.find(key);
        if (it != cache.end()) return it->second;
    }
    void *d = nullptr;
    if (P->prec == 1 || force_double) {
        std::vector<double> h(2 * (size_t)count);
        for (long long k = 0; k < count; ++k) twiddle((k * step) % n, n, &h[2 * k], &h[2 * k + 1]);
        d = upload(h.data(), h.size() * sizeof(double));
    } else {
        std::vector<float> h(2 * (size_t)count);
        for (long long k = 0; k < count; ++k) {
            double re, im;
            twiddle((k * step) % n, n, &re, &im);
            h[2 * k] = (float)re;
            h[2 * k + 1] = (float)im;
        }
        d = upload(h.data(), h.size() * sizeof(float));
    }
    if (step == 1) cache[key] = d;
    return d;
}

// per-stage transposed twiddle tables of a tile pipeline of length L and radix R (layout: tile_kernel.cuh,
// TileParams::tw / stage_tw_offset)
void *Builder::stage_tables(long long L, int R) {
    auto key = std::make_pair(std::make_pair(8, L), (long long)R);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    std::vector<double> h;
    long long scale = 1;  // R^(s-1)
    for (long long m = L / R; m >= 1 && scale * R < L; m /= R, scale *= R) {
        for (long long d = 1; d < R; ++d)
            for (long long lo = 0; lo < m; ++lo) {
                double re, im;
                twiddle((d * lo * scale) % L, L, &re, &im);
                h.push_back(re);
                h.push_back(im);
            }
    }
    if (h.empty()) { h.push_back(1.0); h.push_back(0.0); }
    void *d = nullptr;
    if (P->prec == 1) {
        d = upload(h.data(), h.size() * sizeof(double));
    } else {
        std::vector<float> hf(h.begin(), h.end());
        d = upload(hf.data(), hf.size() * sizeof(float));
    }
    cache[key] = d;
    return d;
}

void *Builder::alloc(size_t bytes) {
    void *d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) { err = FFTB200_ALLOC_FAILED; cudaGetLastError(); return nullptr; }
    P->dev_allocs.push_back(d);
    return d;
}

bool is_pow2(long long v) { return v > 0 && (v & (v - 1)) == 0; }
int ilog2ll(long long v) { int l = 0; while ((1ll << (l + 1)) <= v) ++l; return l; }


// merge adjacent index levels that are dense in both buffers
static void merge_levels(std::vector<Level> &lv) {
    std::vector<Level> out;
    for (size_t i = 0; i < lv.size(); ++i) {
        const Level &l = lv[i];
        // a unit-stride level of extent 1 stays: it is the tile's inner index (ragged last chunk of a slab pass)
        if (l.n == 1 && !l.keep && !(i == 0 && l.is == 1 && l.os == 1 && lv.size() > 1)) continue;
        if (!out.empty() && !l.keep && !out.back().keep && l.is == out.back().n * out.back().is &&
            l.os == out.back().n * out.back().os)
            out.back().n *= l.n;
        else
            out.push_back(l);
    }
    lv.swap(out);
}

static const char *variant_name(int v) {
    static const char *names[] = {"row", "col", "col+twiddle", "row->col", "r2c-row", "col->peers", "c2r-row", "row->peers"};
    return names[v];
}

// Add one tile pass.  levels: non-axis index levels, fastest first (levels[0] = lines of a tile).
bool add_tile_pass(Builder &B, int variant, int L, long long in_ls, long long out_ls, std::vector<Level> lv,
                          int src, int dst, long long twN /* four-step N */, const char *what) {
    return add_tile_pass_with(B, find_tile_kernel(B.P->prec, variant, L), variant, L, in_ls, out_ls, lv, src, dst, twN, what);
}

// same, with the kernel (tile shape) chosen by the caller
bool add_tile_pass_with(Builder &B, const TileKernelInfo *ki, int variant, int L, long long in_ls, long long out_ls,
                        std::vector<Level> lv, int src, int dst, long long twN, const char *what) {
    Plan *P = B.P;
    if (!ki) return false;
    {
        // Load the kernel NOW (CUDA loads kernels lazily, at their first launch, and that load synchronises with the
        // device): a first launch issued while a hand-shake kernel of a slab plan spins on this GPU would block the host
        // thread - in a single-process multi-GPU program (Legion) the very thread that still has to launch the peer's
        // work, i.e. a deadlock until the wait times out.  Seen on 2 x B200 with both plans driven by one thread.
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, (const void *)ki->fn) != cudaSuccess) cudaGetLastError();
    }
    merge_levels(lv);
    if (lv.size() > 3) return false;
    while (lv.size() < 3) lv.push_back({1, 0, 0});
    const bool load_row =
        (variant == V_RR || variant == V_RC || variant == V_RR_R2C || variant == V_RR_C2R || variant == V_RC_PEER);
    const bool store_row = (variant == V_RR || variant == V_RR_R2C || variant == V_RR_C2R);
    if (load_row ? (in_ls != 1) : (lv[0].n > 1 && lv[0].is != 1)) return false;
    if (store_row ? (out_ls != 1) : (lv[0].n > 1 && lv[0].os != 1)) return false;
    if (lv[0].n > 0x7fffffffll || lv[1].n > 0x7fffffffll) return false;

    Launch ln;
    ln.kind = Launch::TILE;
    ln.ki = ki;
    ln.variant = variant;
    ln.src = src;
    ln.dst = dst;
    TileParams &tp = ln.tp;
    const int parts = ki->cluster;  // CTAs that share one line
    tp.tw = (L / parts > ki->R) ? B.stage_tables(L / parts, ki->R) : nullptr;  // stage twiddles of the CTA-local length
    tp.tw_aux = nullptr;
    if (parts > 1) tp.tw_aux = B.table(L, L, false);  // cross-CTA stage twiddles w_L
    if (variant == V_RR_R2C) tp.tw_aux = B.table(2ll * L, L / 2 + 1, false);
    if (variant == V_RR_C2R) tp.tw_aux = B.table(2ll * L, L, false);
    tp.tw4_hi = tp.tw4_lo = nullptr;
    tp.tw4_shift = 0;
    tp.tw4_mask = 0;
    if (variant == V_CC_TW) {
        const int bits = ilog2ll(twN);
        const int sh = (bits + 1) / 2;
        tp.tw4_shift = sh;
        tp.tw4_mask = (1 << sh) - 1;
        tp.tw4_lo = (const double2 *)B.table(twN, 1ll << sh, true);
        tp.tw4_hi = (const double2 *)B.table(twN, (twN >> sh) + 1, true, 1ll << sh);
    }
    tp.in_ls = in_ls;
    tp.out_ls = out_ls;
    tp.n_inner = (int)lv[0].n;
    tp.n_inner_last_o2 = 0;
    tp.in_is = lv[0].is;
    tp.out_is = lv[0].os;
    tp.n_o2 = (int)lv[1].n;
    tp.in_os2 = lv[1].is;
    tp.out_os2 = lv[1].os;
    tp.in_os1 = lv[2].is;
    tp.out_os1 = lv[2].os;
    tp.tiles_per_outer = (int)((lv[0].n + ki->W - 1) / ki->W);
    fast_div_make(tp.tiles_per_outer, &tp.div_tpo_m, &tp.div_tpo_s);
    fast_div_make(tp.n_o2, &tp.div_o2_m, &tp.div_o2_s);
    tp.inverse = 0;
    const long long tiles = (long long)tp.tiles_per_outer * lv[1].n * lv[2].n;
    if (tiles <= 0 || tiles > 0x7fffffffll) return false;
    if (tiles * parts > 0x7fffffffll) return false;
    ln.grid = (unsigned)(tiles * parts);
    tp.n_tiles = (int)tiles;
    tp.prefetch_tiles = 0;
    {
        // L2 prefetch of the tile that will run next in this CTA slot (distance = CTAs resident on the GPU).
        // Measured on B200 (512^3): strided-axis passes whose lines stay inside a few 2 MiB pages gain ~9 %
        // (fp64 y axis 0.760 -> 0.692 ms, 6.2 TB/s); passes with a multi-MiB line stride lose (z axis
        // 0.86 -> 1.12 ms) and contiguous-axis passes do not change, so only the first kind prefetches.
        // (1024^3 fp64, 128 KiB tiles: y axis 8.11 -> 7.23 ms with the prefetch.)
        const bool col_load = !load_row;
        const bool page_local = in_ls * (long long)(P->prec ? 16 : 8) <= 65536;
        const int k = (col_load && page_local) ? 1 : 0;
        if (k > 0 && ki->cluster == 1) {
            int sms = 148, per_sm = 1;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, P->device);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)ki->fn, ki->threads, ki->smem_bytes) != cudaSuccess)
                cudaGetLastError();
            tp.prefetch_tiles = k * sms * (per_sm > 0 ? per_sm : 1);
        }
    }
    const long long lines = lv[0].n * lv[1].n * lv[2].n;
    const size_t ce = P->prec ? 16 : 8;
    if (variant == V_RR_R2C || variant == V_RR_C2R)
        ln.algo_bytes = (unsigned long long)lines * ((unsigned long long)L * ce + (unsigned long long)(L + 1) * ce);
    else
        ln.algo_bytes = (unsigned long long)lines * L * ce * 2ull;
    char buf[256];
    snprintf(buf, sizeof buf, "tile %-11s %s L=%d R=%d W=%d threads=%d smem=%d cluster=%d lines=%lld tiles=%lld (%s)",
             variant_name(variant), P->prec ? "fp64" : "fp32", L, ki->R, ki->W, ki->threads, ki->smem_bytes, ki->cluster,
             lines, tiles, what);
    ln.desc = buf;
    if (ki->smem_bytes > 48 * 1024) {
        if (cudaFuncSetAttribute((const void *)ki->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, ki->smem_bytes) !=
            cudaSuccess) {
            cudaGetLastError();
            B.err = FFTB200_SETUP_FAILED;
            return false;
        }
    }
    P->launches.push_back(ln);
    return B.err == FFTB200_SUCCESS;
}

// split N = 2^k (too long for one tile) into 2 or 3 factors, each tile-friendly
static std::vector<int> split_1d(long long N, int prec) {
    const int k = ilog2ll(N);
    // COL passes want L*W*elt <= 64 KiB with 128-byte segments: L <= 512
    std::vector<int> f;
    // FFTB200_1D_FACTORS=2 (A/B runs): two factors whatever the length, e.g. 2^27 = 2^13 x 2^14 through the cluster kernel
    // and a 128 KiB row kernel whose transposing store is 8 bytes wide - measured, see DESIGN.md §7
    if (env_int_or("FFTB200_1D_FACTORS", 0) == 2 && k <= 28) {
        const int a = k / 2;
        f = {1 << a, 1 << (k - a)};
        return f;
    }
    if (k <= 18) {
        const int a = k / 2;
        f = {1 << a, 1 << (k - a)};
    } else if (k <= 27) {
        const int a = k / 3, b = (k - a) / 2;
        f = {1 << a, 1 << b, 1 << (k - a - b)};
    } else {
        return f;
    }
    (void)prec;
    return f;
}

// in == out is allowed when the output layout occupies the input's addresses line by line: identical strides for complex
// transforms; for real input FFTW's padded in-place format (rows of 2*(n_last/2+1) reals, i.e. every input stride,
// counted in reals, is twice the output stride counted in complex).  Each pass loads its whole tile before it stores.
static bool layouts_coincide(const Plan *P) {
    const int rank = P->rank;
    if (P->in_stride[rank] != 1 || P->out_stride[rank] != 1) return false;
    for (int d = 0; d < rank; ++d) {
        const long long want = P->real ? 2 * P->out_stride[d] : P->out_stride[d];
        if ((d > 0 || P->batch > 1) && P->in_stride[d] != want) return false;
    }
    return true;
}

// ------------------------------------------------------------------------------------------
// fast plan (power-of-two, unit element stride)
// ------------------------------------------------------------------------------------------
static bool build_fast(Builder &B) {
    Plan *P = B.P;
    const int rank = P->rank;
    const long long *n = P->n;
    const int maxL = max_tile_length(P->prec);
    for (int d = 0; d < rank; ++d)
        if (!is_pow2(n[d])) return false;
    if (P->in_stride[rank] != 1 || P->out_stride[rank] != 1) return false;
    long long total = 1;
    for (int d = 0; d < rank; ++d) total *= n[d];
    if (total == 1) return false;  // nothing to transform: generic path copies

    // output extents (last cut to n/2+1 for real input)
    long long nout[3];
    for (int d = 0; d < rank; ++d) nout[d] = n[d];
    const int last = rank - 1;
    if (P->real) {
        if (n[last] < 4 || n[last] / 2 > maxL) return false;
        for (int d = 0; d <= rank; ++d)
            if (d != rank && (P->in_stride[d] & 1)) return false;  // rows must start on a complex boundary
        nout[last] = n[last] / 2 + 1;
    }

    auto levels_for = [&](int axis, bool first_pass) {
        // fastest first: dims after the axis (reverse), dims before the axis (reverse), batch
        std::vector<Level> lv;
        for (int d = rank - 1; d >= 0; --d) {
            if (d == axis) continue;
            Level l;
            l.n = first_pass ? n[d] : nout[d];
            l.is = first_pass ? P->in_stride[d + 1] : P->out_stride[d + 1];
            l.os = P->out_stride[d + 1];
            lv.push_back(l);
        }
        Level b;
        b.n = P->batch;
        b.is = first_pass ? P->in_stride[0] : P->out_stride[0];
        b.os = P->out_stride[0];
        lv.push_back(b);
        return lv;
    };

    // ---- last axis (contiguous) ----
    bool first = true;
    if (P->real) {
        std::vector<Level> lv = levels_for(last, true);
        for (Level &l : lv) l.is /= 2;  // input addressed as packed complex pairs
        if (!add_tile_pass(B, V_RR_R2C, (int)(n[last] / 2), 1, 1, lv, BUF_IN, BUF_OUT, 0, "axis r2c")) return false;
        first = false;
    } else if (n[last] > 1) {
        if (n[last] <= maxL) {
            if (!add_tile_pass(B, V_RR, (int)n[last], 1, 1, levels_for(last, true), BUF_IN, BUF_OUT, 0, "last axis"))
                return false;
            first = false;
        } else {
            // four-step: only as the sole transformed axis of the plan
            for (int d = 0; d < last; ++d)
                if (n[d] != 1) return false;
            std::vector<int> f = split_1d(n[last], P->prec);
            if (f.empty()) return false;
            const long long N = n[last];
            const size_t ce = P->prec ? 16 : 8;
            P->work_bytes = (size_t)N * P->batch * ce;
            P->work[0] = B.alloc(P->work_bytes);
            if (!P->work[0]) return false;
            const long long bi = P->in_stride[0], bo = P->out_stride[0];
            if (f.size() == 2) {
                const long long N1 = f[0], M = f[1];
                if (!add_tile_pass(B, V_CC_TW, (int)N1, M, M, {{M, 1, 1}, {P->batch, bi, N}}, BUF_IN, BUF_WORK0, N,
                                   "four-step 1/2"))
                    return false;
                if (!add_tile_pass(B, V_RC, (int)M, 1, N1, {{N1, M, 1}, {P->batch, N, bo}}, BUF_WORK0, BUF_OUT, 0,
                                   "four-step 2/2"))
                    return false;
            } else {
                const long long N1 = f[0], N2 = f[1], N3 = f[2], M = N2 * N3;
                if (!add_tile_pass(B, V_CC_TW, (int)N1, M, M, {{M, 1, 1}, {P->batch, bi, N}}, BUF_IN, BUF_WORK0, N,
                                   "six-step 1/3"))
                    return false;
                if (!add_tile_pass(B, V_CC_TW, (int)N2, N3, N3, {{N3, 1, 1}, {N1, M, M}, {P->batch, N, N}}, BUF_WORK0,
                                   BUF_WORK0, M, "six-step 2/3"))
                    return false;
                if (!add_tile_pass(B, V_RC, (int)N3, 1, N1 * N2, {{N1, M, 1}, {N2, N3, N1}, {P->batch, N, bo}},
                                   BUF_WORK0, BUF_OUT, 0, "six-step 3/3"))
                    return false;
            }
            P->inplace_ok = true;  // every pass goes through the work buffer
            return true;
        }
    }

    // ---- 3-D complex transforms whose slowest axis has a far line stride: blocked intermediate layout.
    // A strided pass whose line stride spans many 2 MiB pages is bound by address translation / DRAM-page locality
    // on its LOADS (512^3 fp64: 0.83 ms against 0.69 ms for the page-local y axis), while far STORES cost almost
    // nothing (they are fire-and-forget).  So the middle-axis pass writes a blocked intermediate
    //     W[n1][n2/Wt][n0][Wt]      (Wt = tile width, Wt * sizeof(complex) = 128 B)
    // into a work buffer, and the last pass reads every tile as ONE contiguous L*128 B run and writes the natural
    // layout.  Measured on B200: 512^3 fp64 last pass 0.833 -> 0.701 ms, middle pass 0.687 -> 0.711 ms, transform
    // 2.21 -> 2.10 ms.  Costs one work buffer of the transform's size; FFTB200_ZBLOCK=0 (or a failed allocation)
    // keeps the in-place three-pass plan.
    // Real transforms (n2c = n2/2+1 columns, not a multiple of the tile width) get the same layout with a ragged last
    // column block (1024^3 D2Z last pass 4.70 -> 3.8 ms).
    if (env_int_or("FFTB200_ZBLOCK", 1) != 0 && rank == 3 && P->batch == 1 && !first && n[0] > 1 && n[1] > 1) {
        const TileKernelInfo *k1 = find_tile_kernel(P->prec, V_CC, (int)n[1]);
        const TileKernelInfo *k0 = find_tile_kernel(P->prec, V_CC, (int)n[0]);
        const size_t ce = P->prec ? 16 : 8;
        const long long nx = nout[2];
        const bool dense = P->out_stride[3] == 1 && P->out_stride[2] == nx && P->out_stride[1] == n[1] * nx;
        const bool far = (size_t)(n[1] * nx) * ce >= (256u << 10);
        // (single-CTA tiles only.  1024^3 fp64, one 128 KiB tile per CTA: last pass 9.27 -> 7.09 ms, middle pass unchanged)
        const bool single = k1 && k0 && k1->cluster == 1 && k0->cluster == 1;
        if (single && k1->W == k0->W && dense && far && nx >= k1->W) {
            const long long Wt = k1->W, nxb_full = nx / Wt, rag = nx - nxb_full * Wt, nxb = nxb_full + (rag ? 1 : 0);
            const size_t wbytes = (size_t)(n[0] * n[1] * nxb * Wt) * ce;
            void *w = nullptr;
            if (cudaMalloc(&w, wbytes) == cudaSuccess) {
                P->dev_allocs.push_back(w);
                P->work[0] = w;
                P->work_bytes = wbytes;
                // W element (z, y, x) at ((y * nxb + x / Wt) * n0 + z) * Wt + x % Wt
                const long long w_y = nxb * n[0] * Wt, w_xb = n[0] * Wt, w_z = Wt;
                // a ragged last column block (real transforms: nx = n2/2+1) is part of the same launch: its tiles load
                // only their valid lines (clamped) and store only those (TileParams::n_inner_last_o2)
                auto mark_ragged = [&](void) {
                    Launch &ln = P->launches.back();
                    if (rag) {
                        ln.tp.n_inner_last_o2 = (int)rag;
                        ln.algo_bytes = ln.algo_bytes / (unsigned long long)(nxb * Wt) * (unsigned long long)nx;
                    }
                };
                bool ok = true;
                {   // middle axis: natural layout -> blocked work buffer
                    std::vector<Level> lv1 = {{Wt, 1, 1}, {nxb, Wt, w_xb}, {n[0], P->out_stride[1], w_z}};
                    ok = ok && add_tile_pass(B, V_CC, (int)n[1], P->out_stride[2], w_y, lv1, BUF_OUT, BUF_WORK0, 0,
                                             "strided axis -> blocked work buffer");
                    if (ok) mark_ragged();
                }
                if (ok) {   // slowest axis: blocked work buffer (every tile one contiguous run) -> natural layout
                    std::vector<Level> lv0 = {{Wt, 1, 1}, {nxb, w_xb, Wt}, {n[1], w_y, P->out_stride[2]}};
                    ok = add_tile_pass(B, V_CC, (int)n[0], w_z, P->out_stride[1], lv0, BUF_WORK0, BUF_OUT, 0,
                                       "blocked work buffer -> strided axis");
                    if (ok) mark_ragged();
                }
                if (ok) {
                    P->inplace_ok = layouts_coincide(P);  // pass 1 is tile-wise in place, the others go through the work buffer
                    return true;
                }
                return false;  // (the caller discards the partial plan and its allocations)
            }
            cudaGetLastError();  // no memory for the work buffer: in-place plan below
        }
    }
    // ---- remaining axes (strided) ----
    for (int axis = last - 1; axis >= 0; --axis) {
        if (n[axis] == 1) continue;
        if (n[axis] > maxL) return false;
        std::vector<Level> lv = levels_for(axis, first);
        const long long in_ls = first ? P->in_stride[axis + 1] : P->out_stride[axis + 1];
        const long long out_ls = P->out_stride[axis + 1];
        if (!add_tile_pass(B, V_CC, (int)n[axis], in_ls, out_ls, lv, first ? BUF_IN : BUF_OUT, BUF_OUT, 0,
                           "strided axis"))
            return false;
        first = false;
    }
    if (first) return false;
    // in place is safe when every pass reads and writes the same addresses tile by tile
    P->inplace_ok = layouts_coincide(P);
    return true;
}

// ------------------------------------------------------------------------------------------
// mixed-radix plan: transforms with unit element stride whose axes are products of 2, 3, 5, 7, 11 and 13 that fit one
// shared-memory tile (mixed_kernel.cuh).  One kernel per axis, in place over the output array like the fast plan;
// power-of-two axes of such a shape still use the tuned tile kernels.  Anything else goes to the generic path.
// ------------------------------------------------------------------------------------------
// L as a product of the radices the kernel has in-register DFTs for (all <= maxr): fewest stages, then the smallest
// largest radix, then the smallest sum.  Empty when L has another prime factor.
static std::vector<int> mixed_radices(long long L, int maxr) {
    struct Best { int stages = 1 << 20, maxr = 0, sum = 0; std::vector<int> r; };
    std::map<long long, Best> memo;
    std::function<const Best &(long long)> go = [&](long long v) -> const Best & {
        auto it = memo.find(v);
        if (it != memo.end()) return it->second;
        Best best;
        if (v == 1) {
            best.stages = 0;
        } else {
            for (int r : MIXED_RADICES) {
                if (r > maxr || v % r) continue;
                const Best &sub = go(v / r);
                if (sub.stages >= (1 << 20)) continue;
                Best c;
                c.stages = sub.stages + 1;
                c.maxr = std::max(sub.maxr, r);
                c.sum = sub.sum + r;
                if (std::make_tuple(c.stages, c.maxr, c.sum) < std::make_tuple(best.stages, best.maxr, best.sum)) {
                    c.r = sub.r;
                    c.r.push_back(r);
                    best = c;
                }
            }
        }
        return memo[v] = best;
    };
    std::vector<int> r = go(L).r;
    // the first stage stores with a stride of its radix: odd radices first (conflict-free), powers of two last
    auto cls = [](int v) { return (v & 1) ? 0 : ((v & (v - 1)) ? 1 : 2); };
    std::stable_sort(r.begin(), r.end(), [&](int x, int y) { return cls(x) < cls(y); });
    return r;
}

static const size_t MIXED_SMEM_MAX = 200u << 10;

static int mixed_max_radix(int prec) {
    // largest in-register DFT: the register count of the kernel follows it (fp64: radix 16 needs 138 registers)
    const int dflt = 16;
    (void)prec;
    const int v = env_int_or("FFTB200_MIXED_MAXR", dflt);
    return v <= 8 ? 8 : (v <= 10 ? 10 : 16);
}

static bool mixed_axis_ok(long long L, int prec) {
    if (L < 2 || L > 0x7fffff) return false;
    const std::vector<int> r = mixed_radices(L, mixed_max_radix(prec));
    const size_t nbuf = r.size() >= 3 ? 2 : 1;
    return !r.empty() && (int)r.size() <= MIXED_MAX_STAGES && nbuf * (size_t)(L + 1) * (prec ? 16 : 8) <= MIXED_SMEM_MAX;
}

// io: MIXED_C2C, or (contiguous axis only) MIXED_R2C / MIXED_C2R with the real side's strides in real elements
//     MIXED_TW (strided axis; twN = length of the whole two-pass line) / MIXED_RC (contiguous load, transposing store)
static bool add_mixed_pass(Builder &B, bool row, int L, long long in_ls, long long out_ls, std::vector<Level> lv, int src,
                           int dst, const char *what, int io = MIXED_C2C, long long twN = 0, bool tw_o2 = false) {
    Plan *P = B.P;
    if (tw_o2) {
        // the twiddle's column index must survive as the tile's o2 index
        merge_levels(lv);
        if (lv.size() < 2 || lv.size() > 3 || !lv[1].keep) return false;
    }
    if (row ? io == MIXED_TW : (io != MIXED_C2C && io != MIXED_TW)) return false;
    const size_t ce = P->prec ? 16 : 8;
    const int maxr = mixed_max_radix(P->prec);
    const std::vector<int> rad = mixed_radices(L, maxr);
    if (rad.empty() || (int)rad.size() > MIXED_MAX_STAGES) return false;
    // the kernel addresses a line with 32-bit element offsets from its base
    if (in_ls < 0 || out_ls < 0 || (unsigned long long)(L - 1) * (unsigned long long)in_ls >= (1ull << 32) ||
        (unsigned long long)(L - 1) * (unsigned long long)out_ls >= (1ull << 32))
        return false;
    MixedStages ms{};
    ms.n = (int)rad.size();
    ms.L = L;
    ms.tw_o2 = tw_o2 ? 1 : 0;
    const long long lines0 = lv.empty() ? 1 : std::max<long long>(1, lv[0].n);
    const size_t budget = (size_t)env_int_or("FFTB200_MIXED_TILE_KB", 16) << 10;  // one shared-memory buffer
    const bool half = io == MIXED_R2C_HALF || io == MIXED_C2R_HALF;  // (L is half the real length; one more shared-memory pass)
    // exchanges between stages: ping-pong from three stages on; the even/odd pass of the half-length real modes is one more
    const bool stay = half || io == MIXED_RC;  // one more pass over shared memory after (or before) the stages
    const int nbuf = stay ? std::min(2, ms.n) : (ms.n >= 3 ? 2 : (ms.n == 2 ? 1 : 0));
    int W, threads;
    long long lp_sum = 0;
    int lp_max = 0;
    for (int r : rad) { lp_sum += L / r; lp_max = std::max(lp_max, L / r); }
    // Thread layout: `c` threads share the butterflies of one line position set (stage s has L / r_s butterflies per
    // line).  Scored by the butterfly slots that do useful work among the threads resident on an SM, estimated with
    // 512 threads per SM (the kernels need about 128 registers) and 220 KiB of shared memory.
    auto score = [&](int c, int useful_threads, int block_threads, size_t smem_bytes) {
        long long cost = 0;
        for (int r : rad) cost += (long long)((L / r + c - 1) / c) * c;
        const double eff = (double)lp_sum / (double)cost;
        const long long by_smem = smem_bytes ? (long long)((220u << 10) / smem_bytes) : 32;
        const long long ctas = std::max<long long>(1, std::min<long long>({32, by_smem, 512 / block_threads}));
        return eff * (double)(ctas * useful_threads);
    };
    if (row) {
        ms.pitch = L | 1;  // odd: threads of one warp that sit on different lines hit different banks
        const size_t line_bytes = (size_t)ms.pitch * ce * std::max(1, nbuf);
        if (line_bytes > MIXED_SMEM_MAX) return false;
        const int hi = std::min(lp_max, MIXED_MAX_THREADS), lo = std::min(hi, 8);
        double best = -1;
        int tl = hi, nslow = 1;
        // (transposing store: the tile is as wide as a strided-axis tile - 128-byte runs - whatever the line slots are)
        int w_rc = (int)(128 / ce);
        while (w_rc > 1 && (size_t)w_rc * line_bytes > MIXED_SMEM_MAX) w_rc /= 2;
        for (int c = lo; c <= hi; ++c) {
            long long ns = std::min<long long>(std::max(1, MIXED_MAX_THREADS / c), lines0);
            ns = std::min<long long>(ns, std::max<size_t>(1, (64u << 10) / line_bytes));
            if (io == MIXED_RC) {
                long long p2 = 1;
                while (p2 * 2 <= std::min<long long>(std::max(1, MIXED_MAX_THREADS / c), w_rc)) p2 *= 2;
                ns = p2;
            }
            const int bt = (int)((c * ns + 31) / 32 * 32);
            const double sc = score(c, (int)(c * ns), bt, nbuf ? line_bytes * (io == MIXED_RC ? w_rc : ns) : 0);
            if (sc >= best) { best = sc; tl = c; nslow = (int)ns; }
        }
        // small tiles: several rounds of nslow lines per CTA, up to the tile budget
        int rounds = (int)std::max<size_t>(1, budget / ((size_t)ms.pitch * ce * nslow));
        rounds = (int)std::min<long long>(rounds, std::max<long long>(1, lines0 / nslow));
        while (rounds > 1 && line_bytes * nslow * rounds > MIXED_SMEM_MAX) --rounds;
        if (io == MIXED_RC) rounds = std::max(1, w_rc / nslow);
        W = nslow * rounds;
        ms.nfast = tl;
        ms.nslow = nslow;
        threads = (tl * nslow + 31) / 32 * 32;
    } else {
        W = (int)(128 / ce);  // one 128-byte segment per line position
        while (W > 1 && (size_t)std::max(1, nbuf) * L * W * ce > MIXED_SMEM_MAX) W /= 2;
        ms.nfast = W;
        ms.pitch = 0;
        const size_t tile_bytes = (size_t)nbuf * L * W * ce;
        const int unit = std::max(1, 32 / W);  // whole warps
        double best = -1;
        int nslow = unit;
        for (int c = unit; c * W <= MIXED_MAX_THREADS && c < lp_max + unit; c += unit) {
            const double sc = score(c, c * W, c * W, tile_bytes);
            if (sc >= best) { best = sc; nslow = c; }
        }
        ms.nslow = nslow;
        threads = nslow * W;
    }
    ms.W = W;
    const size_t smem = (size_t)nbuf * ce * (row ? (size_t)W * ms.pitch : (size_t)W * L);
    if (smem > MIXED_SMEM_MAX) return false;
    // per-stage twiddle tables: stage s (radix p, Ns = product of the earlier radices) reads entry [(t-1)*Ns + k] =
    // w_L^(t k L / (Ns p)), t in [1, p), k in [0, Ns): consecutive butterflies read consecutive entries
    std::vector<double> tw;
    {
        long long Ns = 1;
        for (int s = 0; s < ms.n; ++s) {
            const int p = rad[s];
            ms.r[s] = (unsigned char)p;
            ms.tw_off[s] = (int)(tw.size() / 2);
            fast_div_make((int)Ns, &ms.ns_m[s], &ms.ns_s[s]);
            if (Ns > 1) {
                const long long step = L / (Ns * p);
                for (int t = 1; t < p; ++t)
                    for (long long k = 0; k < Ns; ++k) {
                        double re, im;
                        twiddle((t * k * step) % L, L, &re, &im);
                        tw.push_back(re);
                        tw.push_back(im);
                    }
            }
            Ns *= p;
        }
        if (tw.empty()) { tw.push_back(1.0); tw.push_back(0.0); }
    }
    void *dtw;
    if (P->prec) {
        dtw = B.upload(tw.data(), tw.size() * sizeof(double));
    } else {
        std::vector<float> twf(tw.begin(), tw.end());
        dtw = B.upload(twf.data(), twf.size() * sizeof(float));
    }
    if (!dtw) return false;
    int kmax = 0;
    for (int r : rad) kmax = std::max(kmax, r);
    std::unique_ptr<TileKernelInfo> ki(new TileKernelInfo);
    ki->fn = reinterpret_cast<void (*)(const TileParams)>(mixed_kernel(P->prec, row, kmax, io));  // (launched with its own signature)
    ki->L = L;
    ki->R = L;  // no per-stage tile tables: add_tile_pass_with leaves tp.tw alone
    ki->W = W;
    ki->threads = threads;
    // (the opt-in shared-memory limit is a property of the kernel, shared by every plan: always raise it to the maximum)
    ki->smem_bytes = (int)MIXED_SMEM_MAX;
    ki->cluster = 1;
    if (!add_tile_pass_with(B, ki.get(), io == MIXED_RC ? V_RC : (row ? V_RR : V_CC), L, in_ls, out_ls, lv, src, dst, 0, what)) return false;
    ki->smem_bytes = (int)smem;
    Launch &ln = P->launches.back();
    ln.kind = Launch::MIXED;
    ln.mixed = ms;
    ln.mixed_row = row;
    if (io != MIXED_C2C) {
        const unsigned long long lines = ln.algo_bytes / ((unsigned long long)L * ce * 2ull);
        const unsigned long long lr = half ? 2ull * L : (unsigned long long)L;  // real length of a line
        ln.algo_bytes = lines * (lr * (ce / 2) + (lr / 2 + 1) * ce);
    }
    if (half) {
        ln.tp.tw_aux = B.table(2ll * L, L / 2 + 1, false);
        if (!ln.tp.tw_aux) return false;
    }
    if (io == MIXED_TW) {
        // w_twN^m = hi[m >> sh] * lo[m & mask], both tables fp64
        const int bits = ilog2ll(twN) + 1, sh = (bits + 1) / 2;
        ln.tp.tw4_shift = sh;
        ln.tp.tw4_mask = (1 << sh) - 1;
        ln.tp.tw4_lo = (const double2 *)B.table(twN, 1ll << sh, true);
        ln.tp.tw4_hi = (const double2 *)B.table(twN, (twN >> sh) + 1, true, 1ll << sh);
        if (!ln.tp.tw4_lo || !ln.tp.tw4_hi) return false;
    }
    ln.tp.tw = dtw;
    ln.tp.prefetch_tiles = 0;
    if (env_int_or("FFTB200_MIXED_PREFETCH", 1) != 0) {
        // distance = CTAs resident on the GPU: the tile this CTA's slot runs next
        int sms = 148, per_sm = 1;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, P->device);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)ki->fn, threads, smem) != cudaSuccess) cudaGetLastError();
        ln.tp.prefetch_tiles = sms * std::max(1, per_sm);
    }
    std::string radices;
    for (int i = 0; i < ms.n; ++i) radices += (i ? "x" : "") + std::to_string((int)ms.r[i]);
    char buf[256];
    snprintf(buf, sizeof buf, "mixed-radix %s %s L=%d (%s) W=%d threads=%d (%d along %s) smem=%d tiles=%u (%s)",
             io == MIXED_R2C ? "r2c-row" : (io == MIXED_C2R ? "c2r-row" : (io == MIXED_R2C_HALF ? "r2c-row (half-length)" : (io == MIXED_C2R_HALF ? "c2r-row (half-length)" : (io == MIXED_TW ? "col+twiddle" : (io == MIXED_RC ? "row->col" : (row ? "row" : "col")))))), P->prec ? "fp64" : "fp32", L,
             radices.c_str(), W, threads, ms.nfast,
             row ? "a line" : "the lines", (int)smem, ln.grid, what);
    ln.desc = buf;
    P->mixed_infos.push_back(std::move(ki));
    return B.err == FFTB200_SUCCESS;
}

static unsigned grid_for(long long total);

static bool build_mixed(Builder &B) {
    Plan *P = B.P;
    const int rank = P->rank, last = rank - 1;
    const long long *n = P->n;
    if (env_int_or("FFTB200_MIXED", 1) == 0) return false;
    if (P->in_stride[rank] != 1 || P->out_stride[rank] != 1) return false;
    const bool real = P->real, c2r = P->c2r;
    const int maxL = max_tile_length(P->prec);
    auto short_axis = [&](long long v) { return (is_pow2(v) && v <= maxL) || mixed_axis_ok(v, P->prec); };
    // (a real last axis of even length only needs its half to fit one tile)
    const bool last_long = n[last] > 1 && !short_axis(n[last]) &&
                           !((real || c2r) && !(n[last] & 1) && mixed_axis_ok(n[last] / 2, P->prec));
    long long total = 1;
    int long_axes = 0;
    for (int d = 0; d < rank; ++d) {
        total *= n[d];
        if (n[d] == 1) continue;
        if (d == last ? last_long : !short_axis(n[d])) ++long_axes;
    }
    if (total == 1) return false;
    if (c2r && long_axes > (last_long ? 1 : 0)) return false;  // (C2R: only the last axis may be a two-pass line)
    if ((real || c2r) && last_long && (n[last] & 1)) return false;  // long real lines: even lengths (half-length scheme)
    // L = N1 N2 with both factors on the one-tile path, as balanced as the radices allow (N1 <= N2); 0 if there is none
    auto split_long = [&](long long Lw) -> long long {
        if (Lw > (1ll << 31) - 1) return 0;
        for (long long a = (long long)std::sqrt((double)Lw) + 1; a >= 2; --a) {
            if (Lw % a) continue;
            if (mixed_axis_ok(a, P->prec) && mixed_axis_ok(Lw / a, P->prec)) return a;
        }
        return 0;
    };
    // A line too long for one tile: L = N1 N2 (n = n1 N2 + n2, k = k1 + N1 k2).  Pass 1: strided transform over n1, store
    // times w_L^(k1 n2), into work buffer `wbuf` (laid out like the destination).  Pass 2: transform over n2, stored at
    // k2 N1 + k1 of BUF_OUT.  `other`: the remaining index levels, fastest first, the first n_inner_levels of them faster
    // than the axis; is = source strides, os = destination strides.
    auto two_pass = [&](bool row, long long Lw, long long in_ls, long long out_ls, const std::vector<Level> &other,
                        size_t n_inner_levels, int src, int wbuf) -> bool {
        const long long N1 = split_long(Lw);
        if (!N1) return false;
        const long long N2 = Lw / N1;  // the strided pass (whole 128-byte columns in shared memory) gets the shorter factor
        std::vector<Level> other_w = other;
        for (Level &l : other_w) l.is = l.os;
        Level l_n2{N2, in_ls, out_ls};
        l_n2.keep = true;
        std::vector<Level> lv1(other.begin(), other.begin() + n_inner_levels);
        lv1.push_back(l_n2);
        lv1.insert(lv1.end(), other.begin() + n_inner_levels, other.end());
        if (!add_mixed_pass(B, false, (int)N1, N2 * in_ls, N2 * out_ls, lv1, src, wbuf, "two-pass axis 1/2", MIXED_TW, Lw, !row))
            return false;
        std::vector<Level> lv2(other_w.begin(), other_w.begin() + n_inner_levels);
        lv2.push_back({N1, N2 * out_ls, out_ls});
        lv2.insert(lv2.end(), other_w.begin() + n_inner_levels, other_w.end());
        if (row) return add_mixed_pass(B, true, (int)N2, 1, N1, lv2, wbuf, BUF_OUT, "two-pass axis 2/2", MIXED_RC);
        return add_mixed_pass(B, false, (int)N2, out_ls, N1 * out_ls, lv2, wbuf, BUF_OUT, "two-pass axis 2/2");
    };
    // line layouts (every index but the last axis) for the even/odd kernels of long real lines, strides in complex elements
    auto line_layout = [&](const long long *strides /* [batch, d0 ..] */, long long div) {
        GenLayout g{};
        g.nd = rank;  // batch + (rank - 1) leading dims
        g.n[0] = P->batch;
        g.stride[0] = strides[0] / div;
        for (int d = 0; d < last; ++d) { g.n[d + 1] = n[d]; g.stride[d + 1] = strides[d + 1] / div; }
        return g;
    };
    long long n_lines = P->batch;
    for (int d = 0; d < last; ++d) n_lines *= n[d];
    if ((real || c2r) && n[last] < 2) return false;
    const long long nc = n[last] / 2 + 1;
    // one pass along `axis`: the tuned power-of-two tile kernel when there is one, else the mixed-radix kernel
    auto axis_pass = [&](bool row, int L, long long in_ls, long long out_ls, const std::vector<Level> &lv, int src, int dst,
                         const char *what) {
        if (is_pow2(L) && find_tile_kernel(P->prec, row ? V_RR : V_CC, L))
            return add_tile_pass(B, row ? V_RR : V_CC, L, in_ls, out_ls, lv, src, dst, 0, what);
        return add_mixed_pass(B, row, L, in_ls, out_ls, lv, src, dst, what);
    };

    if (c2r) {
        // backward passes over the n_last/2+1 columns through a work buffer (the input survives), then the last axis:
        // half-spectrum lines -> real lines
        long long outer_axes = 1;
        for (int d = 0; d < last; ++d) outer_axes *= n[d];
        const size_t ce = P->prec ? 16 : 8;
        int cur = BUF_IN;
        long long cs[4];  // strides of the current complex array, [batch, d0.., d_last]
        for (int d = 0; d <= rank; ++d) cs[d] = P->in_stride[d];
        if (outer_axes > 1) {
            long long ws[4];  // dense work layout [batch][n0]..[nc]
            ws[rank] = 1;
            for (int d = rank - 1; d >= 0; --d) ws[d] = ws[d + 1] * (d == last ? nc : n[d]);
            P->work_bytes = (size_t)(ws[0] * P->batch) * ce;
            P->work[0] = B.alloc(P->work_bytes);
            if (!P->work[0]) return false;
            for (int axis = last - 1; axis >= 0; --axis) {
                if (n[axis] == 1) continue;
                std::vector<Level> lv;
                for (int d = rank - 1; d >= 0; --d) {
                    if (d == axis) continue;
                    lv.push_back({d == last ? nc : n[d], cs[d + 1], ws[d + 1]});
                }
                lv.push_back({(long long)P->batch, cs[0], ws[0]});
                if (!axis_pass(false, (int)n[axis], cs[axis + 1], ws[axis + 1], lv, cur, BUF_WORK0, "strided axis (backward)"))
                    return false;
                cur = BUF_WORK0;
                for (int d = 0; d <= rank; ++d) cs[d] = ws[d];
            }
        }
        bool even_out = true;
        for (int d = 0; d < rank; ++d) even_out = even_out && !(P->out_stride[d] & 1);
        if (last_long) {
            // long real line: even/odd pre-pass (half spectrum -> Z' in the output lines), then the two-pass backward
            // complex transform of Z' through a second work buffer, stored as pairs of reals
            if (!even_out) return false;
            const long long Lh = n[last] / 2;
            if (!split_long(Lh)) return false;
            P->work[1] = B.alloc(P->span_out);
            if (!P->work[1]) return false;
            Launch pre;
            pre.kind = Launch::C2R_PRE;
            pre.lay2 = line_layout(cs, 1);
            pre.lay = line_layout(P->out_stride, 2);
            pre.L = (int)Lh;
            pre.total = n_lines * (Lh / 2 + 1);
            pre.bhat = B.table(2 * Lh, Lh / 2 + 1, false);
            if (!pre.bhat) return false;
            pre.src = cur;
            pre.dst = BUF_OUT;
            pre.grid = grid_for(pre.total);
            pre.algo_bytes = (unsigned long long)n_lines * (unsigned long long)(2 * Lh + 1) * ce;
            pre.desc = "c2r even/odd pre-pass (half spectrum -> spectrum of the packed pairs)";
            P->launches.push_back(pre);
            std::vector<Level> other;
            for (int d = last - 1; d >= 0; --d) other.push_back({n[d], P->out_stride[d + 1] / 2, P->out_stride[d + 1] / 2});
            other.push_back({(long long)P->batch, P->out_stride[0] / 2, P->out_stride[0] / 2});
            if (!two_pass(true, Lh, 1, 1, other, 0, BUF_OUT, BUF_WORK1)) return false;
            P->inplace_ok = outer_axes > 1 || P->batch == 1 || P->out_stride[0] == 2 * P->in_stride[0];
            return true;
        }
        const bool tile_last = is_pow2(n[last]) && n[last] >= 4 && even_out && find_tile_kernel(P->prec, V_RR_C2R, (int)(n[last] / 2));
        std::vector<Level> lv;
        const long long div = tile_last ? 2 : 1;  // (the tile kernel addresses the reals as complex pairs)
        for (int d = last - 1; d >= 0; --d) lv.push_back({n[d], cs[d + 1], P->out_stride[d + 1] / div});
        lv.push_back({(long long)P->batch, cs[0], P->out_stride[0] / div});
        if (tile_last) {
            if (!add_tile_pass(B, V_RR_C2R, (int)(n[last] / 2), 1, 1, lv, cur, BUF_OUT, 0, "axis c2r")) return false;
            P->inplace_ok = false;
        } else {
            // even length, rows on a pair of reals: half-length complex transform + even/odd pass; else the full-length form
            bool done = false;
            if (!(n[last] & 1) && even_out && mixed_axis_ok(n[last] / 2, P->prec) && env_int_or("FFTB200_MIXED_HALF", 1)) {
                std::vector<Level> lh = lv;
                for (Level &l : lh) l.os /= 2;
                done = add_mixed_pass(B, true, (int)(n[last] / 2), 1, 1, lh, cur, BUF_OUT, "axis c2r", MIXED_C2R_HALF);
            }
            if (!done && !add_mixed_pass(B, true, (int)n[last], 1, 1, lv, cur, BUF_OUT, "axis c2r", MIXED_C2R)) return false;
            // in place: through the work buffer the input is consumed before the output is written; a direct pass is
            // safe when every CTA's lines occupy the same bytes on both sides (it loads its tile before it stores)
            P->inplace_ok = outer_axes > 1 || P->batch == 1 || P->out_stride[0] == 2 * P->in_stride[0];
        }
        return true;
    }

    long long nout[3];
    for (int d = 0; d < rank; ++d) nout[d] = n[d];
    if (real) nout[last] = nc;
    auto levels_for = [&](int axis, bool first_pass) {
        std::vector<Level> lv;
        for (int d = rank - 1; d >= 0; --d) {
            if (d == axis) continue;
            lv.push_back({first_pass ? n[d] : nout[d], first_pass ? P->in_stride[d + 1] : P->out_stride[d + 1], P->out_stride[d + 1]});
        }
        lv.push_back({P->batch, first_pass ? P->in_stride[0] : P->out_stride[0], P->out_stride[0]});
        return lv;
    };
    bool first = true;
    if (real) {
        bool even_in = true;
        for (int d = 0; d < rank; ++d) even_in = even_in && !(P->in_stride[d] & 1);
        std::vector<Level> lv = levels_for(last, true);
        if (last_long) {
            // long real line: two-pass complex transform of the packed pairs into the output lines, then the even/odd
            // pass in place on them
            if (!even_in) return false;
            const long long Lh = n[last] / 2;
            P->work_bytes = P->span_out;
            P->work[0] = B.alloc(P->work_bytes);
            if (!P->work[0]) return false;
            std::vector<Level> other = lv;
            for (Level &l : other) l.is /= 2;
            if (!two_pass(true, Lh, 1, 1, other, 0, BUF_IN, BUF_WORK0)) return false;
            Launch post;
            post.kind = Launch::R2C_POST;
            post.lay = line_layout(P->out_stride, 1);
            post.L = (int)Lh;
            post.total = n_lines * (Lh / 2 + 1);
            post.bhat = B.table(2 * Lh, Lh / 2 + 1, false);
            if (!post.bhat) return false;
            post.src = BUF_OUT;
            post.dst = BUF_OUT;
            post.grid = grid_for(post.total);
            post.algo_bytes = (unsigned long long)n_lines * (unsigned long long)(2 * Lh + 1) * (P->prec ? 16 : 8);
            post.desc = "r2c even/odd post-pass (in place on the half-spectrum lines)";
            P->launches.push_back(post);
        } else if (is_pow2(n[last]) && n[last] >= 4 && even_in && find_tile_kernel(P->prec, V_RR_R2C, (int)(n[last] / 2))) {
            for (Level &l : lv) l.is /= 2;  // input addressed as packed complex pairs
            if (!add_tile_pass(B, V_RR_R2C, (int)(n[last] / 2), 1, 1, lv, BUF_IN, BUF_OUT, 0, "axis r2c")) return false;
        } else {
            bool done = false;
            if (!(n[last] & 1) && even_in && mixed_axis_ok(n[last] / 2, P->prec) && env_int_or("FFTB200_MIXED_HALF", 1)) {
                std::vector<Level> lh = lv;
                for (Level &l : lh) l.is /= 2;  // input addressed as packed complex pairs
                done = add_mixed_pass(B, true, (int)(n[last] / 2), 1, 1, lh, BUF_IN, BUF_OUT, "axis r2c", MIXED_R2C_HALF);
            }
            if (!done && !add_mixed_pass(B, true, (int)n[last], 1, 1, lv, BUF_IN, BUF_OUT, "axis r2c", MIXED_R2C)) return false;
        }
        first = false;
    }
    if (long_axes && !P->work[0]) {
        // two-pass axes go through a work buffer laid out like the output array
        P->work_bytes = P->span_out;
        P->work[0] = B.alloc(P->work_bytes);
        if (!P->work[0]) return false;
    }
    bool first_two_pass = real && last_long;  // the first pass reads all of the input before anything is written to the output
    for (int axis = real ? last - 1 : last; axis >= 0; --axis) {
        if (n[axis] == 1) continue;
        const bool row = axis == last;
        const long long in_ls = first ? P->in_stride[axis + 1] : P->out_stride[axis + 1];
        const long long out_ls = P->out_stride[axis + 1];
        const int src = first ? BUF_IN : BUF_OUT;
        if (short_axis(n[axis])) {
            if (!axis_pass(row, (int)n[axis], in_ls, out_ls, levels_for(axis, first), src, BUF_OUT, row ? "last axis" : "strided axis"))
                return false;
            first = false;
            continue;
        }
        if (first) first_two_pass = true;
        if (!two_pass(row, n[axis], in_ls, out_ls, levels_for(axis, first), (size_t)(last - axis), src, BUF_WORK0)) return false;
        first = false;
    }
    // in place: tile-wise in-place passes need coinciding layouts; a two-pass axis reads everything before it writes the output
    P->inplace_ok = layouts_coincide(P) || first_two_pass;
    return true;
}

// ------------------------------------------------------------------------------------------
// inverse real plan (C2R / Z2D): backward COL passes over the n_last/2+1 columns (through a work buffer, so the
// input survives), then the fused even/odd pre-pass + half-length backward FFT on the last axis
// ------------------------------------------------------------------------------------------
static bool build_c2r(Builder &B) {
    Plan *P = B.P;
    const int rank = P->rank, last = rank - 1;
    const long long *n = P->n;
    const int maxL = max_tile_length(P->prec);
    for (int d = 0; d < rank; ++d)
        if (!is_pow2(n[d])) return false;
    if (P->in_stride[rank] != 1 || P->out_stride[rank] != 1) return false;
    if (n[last] < 4 || n[last] / 2 > maxL) return false;
    for (int d = 0; d < rank; ++d)
        if (P->out_stride[d] & 1) return false;  // output rows / batches must start on a pair of reals
    const long long nc = n[last] / 2 + 1;
    long long outer_axes = 1;
    for (int d = 0; d < last; ++d) {
        if (n[d] > maxL) return false;
        outer_axes *= n[d];
    }
    const size_t ce = P->prec ? 16 : 8;
    int cur = BUF_IN;
    long long cs[4];  // strides of the current complex array, [batch, d0.., d_last]
    for (int d = 0; d <= rank; ++d) cs[d] = P->in_stride[d];
    if (outer_axes > 1) {
        long long ws[4];  // dense work layout [batch][n0]..[nc]
        ws[rank] = 1;
        for (int d = rank - 1; d >= 0; --d) ws[d] = ws[d + 1] * (d == last ? nc : n[d]);
        const long long total = ws[0] * P->batch;
        P->work_bytes = (size_t)total * ce;
        P->work[0] = B.alloc(P->work_bytes);
        if (!P->work[0]) return false;
        for (int axis = last - 1; axis >= 0; --axis) {
            if (n[axis] == 1) continue;
            std::vector<Level> lv;
            for (int d = rank - 1; d >= 0; --d) {
                if (d == axis) continue;
                lv.push_back({d == last ? nc : n[d], cs[d + 1], ws[d + 1]});
            }
            lv.push_back({(long long)P->batch, cs[0], ws[0]});
            if (!add_tile_pass(B, V_CC, (int)n[axis], cs[axis + 1], ws[axis + 1], lv, cur, BUF_WORK0, 0, "strided axis (backward)"))
                return false;
            cur = BUF_WORK0;
            for (int d = 0; d <= rank; ++d) cs[d] = ws[d];
        }
    }
    // last axis: lines of nc complex -> n_last reals (addressed as n_last/2 complex pairs)
    std::vector<Level> lv;
    for (int d = last - 1; d >= 0; --d) lv.push_back({n[d], cs[d + 1], P->out_stride[d + 1] / 2});
    lv.push_back({(long long)P->batch, cs[0], P->out_stride[0] / 2});
    if (!add_tile_pass(B, V_RR_C2R, (int)(n[last] / 2), 1, 1, lv, cur, BUF_OUT, 0, "axis c2r")) return false;
    P->inplace_ok = false;
    return true;
}

// ------------------------------------------------------------------------------------------
// generic plan
// ------------------------------------------------------------------------------------------
static unsigned grid_for(long long total) {
    long long g = (total + 255) / 256;
    if (g > 148ll * 64) g = 148ll * 64;
    if (g < 1) g = 1;
    return (unsigned)g;
}

static int largest_prime_factor(long long v) {
    long long best = 1;
    for (long long p = 2; p * p <= v; ++p)
        while (v % p == 0) { best = p; v /= p; }
    if (v > 1) best = v;
    return (int)best;
}

// Prime factors above this run as Bluestein convolutions instead of O(L * p) radix-p stages (fftw-3.3.8/dft/bluestein.c,
// dft/rader.c are the CPU path's answers to the same problem)
static const int BLUESTEIN_MIN_PRIME = 31;

// One axis of a generic plan by Bluestein's algorithm: packed [outer][L][inner] in buffer *cur -> the other work buffer.
// The convolution runs through BUF_BLU ([outer][M][inner], M = 2^m >= 2L-1) with the library's own power-of-two tile
// passes (forward, pointwise product with the transformed chirp, backward).  Returns false (nothing added) when M does
// not fit a single tile pass; the caller then falls back to the radix-p stages.
static bool add_bluestein_axis_impl(Builder &B, int *cur, int axis, long long outer, int L, long long inner);

static bool add_bluestein_axis(Builder &B, int *cur, int axis, long long outer, int L, long long inner) {
    // all or nothing: a failure half-way must not leave some of the axis's launches behind (the caller falls back to
    // the radix-p stages)
    const size_t n_launches = B.P->launches.size();
    const int cur0 = *cur;
    if (add_bluestein_axis_impl(B, cur, axis, outer, L, inner)) return true;
    B.P->launches.resize(n_launches);
    *cur = cur0;
    return false;
}

static bool add_bluestein_axis_impl(Builder &B, int *cur, int axis, long long outer, int L, long long inner) {
    Plan *P = B.P;
    int M = 1;
    while (M < 2 * L - 1) M *= 2;
    const int variant = inner == 1 ? V_RR : V_CC;
    if (!find_tile_kernel(P->prec, variant, M)) return false;
    const size_t ce = P->prec ? 16 : 8;
    const size_t need = (size_t)outer * M * inner * ce;
    if (P->blu_bytes < need) {
        void *nb = B.alloc(need);  // (earlier, smaller buffers stay owned by the plan; at most one per axis)
        if (!nb) return false;
        P->blu = nb;
        P->blu_bytes = need;
    }
    // chirp c[j] = exp(-i pi j^2 / L) = w_{2L}^(j^2 mod 2L), exact integer reduction
    std::vector<double> hc(2 * (size_t)L);
    for (long long j = 0; j < L; ++j) twiddle((j * j) % (2ll * L), 2ll * L, &hc[2 * j], &hc[2 * j + 1]);
    const double2 *chirp = (const double2 *)B.upload(hc.data(), hc.size() * sizeof(double));
    if (!chirp) return false;
    // b[j] = conj(c[j]) for |j| < L, wrapped into M points; Bhat = FFT_M(b), computed once on the device
    void *bhat = nullptr;
    {
        std::vector<double> hb(2 * (size_t)M, 0.0);
        for (long long j = 0; j < L; ++j) {
            hb[2 * j] = hc[2 * j];
            hb[2 * j + 1] = -hc[2 * j + 1];
            if (j > 0) { hb[2 * (M - j)] = hc[2 * j]; hb[2 * (M - j) + 1] = -hc[2 * j + 1]; }
        }
        if (P->prec) {
            bhat = B.upload(hb.data(), hb.size() * sizeof(double));
        } else {
            std::vector<float> hf(hb.begin(), hb.end());
            bhat = B.upload(hf.data(), hf.size() * sizeof(float));
        }
        if (!bhat) return false;
        Plan *sub = nullptr;
        const long long nn[1] = {M}, st[2] = {M, 1};
        if (create_plan(&sub, 1, nn, 1, st, st, P->prec ? FFTB200_Z2Z : FFTB200_C2C, false) != FFTB200_SUCCESS) return false;
        const int rc = exec_plan(sub, bhat, bhat, FFTB200_FORWARD);
        cudaStreamSynchronize(sub->stream);
        delete sub;
        if (rc != FFTB200_SUCCESS) return false;
    }
    const long long total_m = outer * M * inner, total_l = outer * L * inner;
    auto grid = [](long long t) { long long g = (t + 255) / 256; if (g > 148ll * 64) g = 148ll * 64; return (unsigned)(g < 1 ? 1 : g); };
    char buf[200];
    Launch pre;
    pre.kind = Launch::BLU_PRE;
    pre.outer = outer; pre.inner = inner; pre.L = L; pre.M = M; pre.chirp = chirp;
    pre.src = *cur; pre.dst = BUF_BLU;
    pre.total = total_m; pre.grid = grid(total_m);
    pre.algo_bytes = (unsigned long long)(total_l + total_m) * ce;
    snprintf(buf, sizeof buf, "bluestein axis=%d L=%d M=%d: chirp multiply + zero pad, lines=%lld", axis, L, M, outer * inner);
    pre.desc = buf;
    P->launches.push_back(pre);
    for (int dir = 1; dir <= 2; ++dir) {
        std::vector<Level> lv;
        if (inner == 1) lv = {{outer, (long long)M, (long long)M}};
        else lv = {{inner, 1, 1}, {outer, M * inner, M * inner}};
        if (!add_tile_pass(B, variant, M, inner, inner, lv, BUF_BLU, BUF_BLU, 0,
                           dir == 1 ? "bluestein convolution: forward FFT" : "bluestein convolution: backward FFT"))
            return false;
        P->launches.back().dir_override = dir;
        if (dir == 1) {
            Launch mul;
            mul.kind = Launch::BLU_MUL;
            mul.outer = outer; mul.inner = inner; mul.M = M; mul.bhat = bhat;
            mul.src = BUF_BLU; mul.dst = BUF_BLU;
            mul.total = total_m; mul.grid = grid(total_m);
            mul.algo_bytes = (unsigned long long)total_m * ce * 2ull;
            snprintf(buf, sizeof buf, "bluestein axis=%d M=%d: product with the transformed chirp", axis, M);
            mul.desc = buf;
            P->launches.push_back(mul);
        }
    }
    Launch post;
    post.kind = Launch::BLU_POST;
    post.outer = outer; post.inner = inner; post.L = L; post.M = M; post.chirp = chirp;
    post.src = BUF_BLU;
    post.dst = (*cur == BUF_WORK0) ? BUF_WORK1 : BUF_WORK0;
    post.total = total_l; post.grid = grid(total_l);
    post.algo_bytes = (unsigned long long)(total_l + total_m) * ce;
    snprintf(buf, sizeof buf, "bluestein axis=%d L=%d M=%d: chirp multiply, scale 1/M, truncate", axis, L, M);
    post.desc = buf;
    P->launches.push_back(post);
    *cur = post.dst;
    return B.err == FFTB200_SUCCESS;
}

static bool build_generic(Builder &B) {
    Plan *P = B.P;
    const int rank = P->rank;
    const long long *n = P->n;
    const size_t ce = P->prec ? 16 : 8;
    long long total_in = P->batch;
    for (int d = 0; d < rank; ++d) total_in *= n[d];
    P->work_bytes = (size_t)total_in * ce;
    P->work[0] = B.alloc(P->work_bytes);
    P->work[1] = B.alloc(P->work_bytes);
    if (!P->work[0] || !P->work[1]) return false;
    P->generic = true;
    P->inplace_ok = true;

    Launch g;
    g.kind = P->c2r ? Launch::GEN_GATHER_HERM : Launch::GEN_GATHER;
    g.real_in = P->real;
    g.lay.nd = rank + 1;
    g.lay.n[0] = P->batch;
    g.lay.stride[0] = P->in_stride[0];
    for (int d = 0; d < rank; ++d) { g.lay.n[d + 1] = n[d]; g.lay.stride[d + 1] = P->in_stride[d + 1]; }
    g.total = total_in;
    g.src = BUF_IN;
    g.dst = BUF_WORK0;
    g.grid = grid_for(total_in);
    g.algo_bytes = (unsigned long long)total_in * (P->elt_in() + ce);
    g.desc = P->c2r ? "generic gather (half spectrum -> full Hermitian array, packed)" : "generic gather (user layout -> packed complex)";
    P->launches.push_back(g);

    int cur = BUF_WORK0;
    long long dims[3];
    for (int d = 0; d < rank; ++d) dims[d] = n[d];
    for (int axis = rank - 1; axis >= 0; --axis) {
        const int L = (int)dims[axis];
        long long outer = P->batch, inner = 1;
        for (int d = 0; d < axis; ++d) outer *= dims[d];
        for (int d = axis + 1; d < rank; ++d) inner *= dims[d];
        if (L > 1 && largest_prime_factor(L) > BLUESTEIN_MIN_PRIME && add_bluestein_axis(B, &cur, axis, outer, L, inner)) {
            // (length with a large prime factor: chirp-z through a power-of-two convolution, O(L log L) per line)
        } else if (L > 1) {
            const double2 *tw = (const double2 *)B.table(L, L, true);
            int rem = L, Ns = 1;
            for (int p = 2; rem > 1; ++p) {
                if ((long long)p * p > rem) p = rem;
                while (rem % p == 0) {
                    Launch s;
                    s.kind = Launch::GEN_STAGE;
                    s.outer = outer;
                    s.inner = inner;
                    s.L = L;
                    s.p = p;
                    s.Ns = Ns;
                    s.gtw = tw;
                    s.src = cur;
                    s.dst = (cur == BUF_WORK0) ? BUF_WORK1 : BUF_WORK0;
                    s.total = outer * L * inner;
                    s.grid = grid_for(s.total);
                    s.algo_bytes = (unsigned long long)s.total * ce * 2ull;
                    char buf[160];
                    snprintf(buf, sizeof buf, "generic stage axis=%d L=%d radix=%d Ns=%d lines=%lld", axis, L, p, Ns,
                             outer * inner);
                    s.desc = buf;
                    P->launches.push_back(s);
                    cur = s.dst;
                    Ns *= p;
                    rem /= p;
                }
            }
        }
        if (P->real && axis == rank - 1) {
            Launch t;
            t.kind = Launch::GEN_TRUNC;
            t.outer = outer;  // lines (inner == 1 on the last axis)
            t.L = L;
            t.Lc = L / 2 + 1;
            t.src = cur;
            t.dst = (cur == BUF_WORK0) ? BUF_WORK1 : BUF_WORK0;
            t.total = outer * t.Lc;
            t.grid = grid_for(t.total);
            t.algo_bytes = (unsigned long long)t.total * ce * 2ull;
            t.desc = "generic truncate to n/2+1";
            P->launches.push_back(t);
            cur = t.dst;
            dims[axis] = t.Lc;
        }
    }
    Launch s;
    s.kind = P->c2r ? Launch::GEN_SCATTER_REAL : Launch::GEN_SCATTER;
    s.lay.nd = rank + 1;
    s.lay.n[0] = P->batch;
    s.lay.stride[0] = P->out_stride[0];
    long long total_out = P->batch;
    for (int d = 0; d < rank; ++d) { s.lay.n[d + 1] = dims[d]; s.lay.stride[d + 1] = P->out_stride[d + 1]; total_out *= dims[d]; }
    s.total = total_out;
    s.src = cur;
    s.dst = BUF_OUT;
    s.grid = grid_for(total_out);
    s.algo_bytes = (unsigned long long)total_out * ce * 2ull;
    s.desc = P->c2r ? "generic scatter (real part -> user layout)" : "generic scatter (packed complex -> user layout)";
    P->launches.push_back(s);
    return B.err == FFTB200_SUCCESS;
}

int create_plan(Plan **out, int rank, const long long *n, int batch, const long long *in_stride,
                       const long long *out_stride, fftb200_type type, bool force_generic) {
    std::unique_ptr<Plan> P(new Plan);
    if (cudaGetDevice(&P->device) != cudaSuccess) { cudaGetLastError(); return FFTB200_SETUP_FAILED; }
    P->type = type;
    P->prec = (type == FFTB200_Z2Z || type == FFTB200_D2Z || type == FFTB200_Z2D) ? 1 : 0;
    P->real = (type == FFTB200_R2C || type == FFTB200_D2Z);
    P->c2r = (type == FFTB200_C2R || type == FFTB200_Z2D);
    P->rank = rank;
    P->batch = batch;
    for (int d = 0; d < rank; ++d) P->n[d] = n[d];
    for (int d = 0; d <= rank; ++d) { P->in_stride[d] = in_stride[d]; P->out_stride[d] = out_stride[d]; }
    {
        long long li = 0, lo = 0, dense = 1;
        for (int d = 0; d < rank; ++d) {
            const long long no = (P->real && d == rank - 1) ? n[d] / 2 + 1 : n[d];
            const long long ni = (P->c2r && d == rank - 1) ? n[d] / 2 + 1 : n[d];
            li += (ni - 1) * in_stride[d + 1];
            lo += (no - 1) * out_stride[d + 1];
            dense *= no;
        }
        li += (long long)(batch - 1) * in_stride[0];
        lo += (long long)(batch - 1) * out_stride[0];
        dense *= batch;
        P->span_in = (size_t)(li + 1) * P->elt_in();
        P->span_out = (size_t)(lo + 1) * P->elt_out();
        P->out_dense = (lo + 1 == dense);
    }
    Builder B;
    B.P = P.get();
    bool ok = false;
    if (!force_generic) {
        auto discard = [&]() {
            P->launches.clear();
            for (void *d : P->dev_allocs) cudaFree(d);
            P->dev_allocs.clear();
            P->mixed_infos.clear();
            B.cache.clear();
            P->work[0] = P->work[1] = nullptr;
            P->work_bytes = 0;
        };
        ok = P->c2r ? build_c2r(B) : build_fast(B);  // (power-of-two, unit-stride layouts; anything else below)
        if (!ok) {
            discard();
            if (B.err != FFTB200_SUCCESS) return B.err;
            ok = build_mixed(B);  // axes of the form 2^a 3^b 5^c 7^d 11^e 13^f that fit one shared-memory tile
        }
        if (!ok) {
            // discard partial fast plan
            P->launches.clear();
            for (void *d : P->dev_allocs) cudaFree(d);
            P->dev_allocs.clear();
            B.cache.clear();
            P->work[0] = P->work[1] = nullptr;
            P->work_bytes = 0;
            if (B.err != FFTB200_SUCCESS) return B.err;
        }
    }
    if (!ok) ok = build_generic(B);
    if (!ok) {
        free_plan_resources(P.get());
        return B.err != FFTB200_SUCCESS ? B.err : FFTB200_UNSUPPORTED;
    }
    *out = P.release();
    return FFTB200_SUCCESS;
}

}  // namespace fftb200
