// tile_kernel.cuh — the shared-memory Stockham pass: one kernel = one HBM round trip over one axis.
//
// A pass transforms a batch of length-L lines (L = 2^k).  A CTA owns a TILE of W lines that are
// adjacent in the "inner" index; the tile (L x W complex) lives in shared memory between radix
// stages, every thread keeps R points in registers and runs radix-R butterflies
// (butterfly.cuh), and each point crosses HBM exactly once in and once out.
//
// Addressing (elements): in[o1*in_os1 + o2*in_os2 + i*in_is + l*in_ls], same for out.
//   ROW access  (x_ls == 1): lanes run along l  -> a warp touches one contiguous run of a line
//   COL access  (x_is == 1): lanes run along i  -> W*sizeof(complex) = 128 B segments per l
// Load and store sides choose independently, so a pass can also transpose (four-step 1-D).
//
// Stage algebra (decimation in frequency, index digits):  L = r_1 r_2 ... r_S, r_1..r_{S-1} = R,
// m_s = L / (r_1..r_s).  Stage s runs size-r_s DFTs over digit s, then multiplies by
// w_{m_{s-1}}^{k_s * lo} (lo = the not-yet-transformed low part), in place at smem position
// hi*m_{s-1} + d*m_s + lo.  The last stage un-reverses the digits when it computes the output
// index, so HBM stores are in natural order (autosort) and coalesced.
//
// Shared-memory bank conflicts: the linear tile index idx = pos*W + w is XOR-swizzled with the
// fold of its upper bits (3-bit groups for 16-byte elements, 4-bit groups for 8-byte elements),
// which makes every access whose lanes differ in 3 (4) consecutive index bits conflict-free,
// whichever digit the lanes run along.
//
// Reference counterpart: the FFTW plan trees of SURVEY.md §8(a9) (dft/ct.c:34-58 Cooley-Tukey,
// dft/dftw-direct.c:46-56 twiddle codelets, dft/rank-geq2.c:42-52 axis split,
// rdft/ct-hc2c.c:59-70 r2c post-pass) — here one fused kernel per axis.
#pragma once
#include <cooperative_groups.h>

#include "butterfly.cuh"

namespace fftb200 {

enum TileVariant : int {
    V_RR = 0,      // ROW load, ROW store                        (contiguous axis)
    V_CC = 1,      // COL load, COL store                        (strided axis)
    V_CC_TW = 2,   // COL/COL + multiply by w_N^(i*k) on store   (four-step, first factor)
    V_RC = 3,      // ROW load, COL store                        (four-step, last factor: transposing)
    V_RR_R2C = 4,  // ROW load of packed reals, half-length FFT, even/odd post-pass, ROW store
    V_CC_PEER = 5, // COL/COL, output index k scattered over up to 16 destination buffers: the slab
                   // exchange fused into the FFT pass (peer GPUs' memory over NVLink, or the blocks of
                   // a local all-to-all send buffer)
    V_RR_C2R = 6,  // ROW load of L+1 half-spectrum bins, even/odd pre-pass, half-length inverse FFT, ROW store of
                   // 2L reals (the inverse of V_RR_R2C; no reference call site: src/fft.rg is forward-only)
    V_RC_PEER = 7, // ROW load, COL store scattered over destination buffers by output index k: the exchange of a
                   // 2-D slab transform (row FFT + global transpose in one pass)
    V_COUNT = 8
};

constexpr int MAX_PEERS = 16;

struct TileParams {
    const void *in;
    void *out;
    const void *tw;      // per-stage twiddle tables of the CTA-local length, transposed so that a warp reads them as
                         // contiguous runs: stage s < S at offset stage_tw_offset(s), entry [(d-1)*m_s + lo] =
                         // w_L^(d * lo * R^(s-1)), d in [1,R), lo in [0, m_s), m_s = L / R^s; forward sign, complex<T>
    const void *tw_aux;  // V_RR_R2C: w_{2L}^k, k in [0, L/2]; V_RR_C2R: k in [0, L); cluster: w_L^k; complex<T>
    const double2 *tw4_hi;  // V_CC_TW: w_N^(m) = hi[m >> tw4_shift] * lo[m & tw4_mask]
    const double2 *tw4_lo;
    long long in_ls, in_is, in_os1, in_os2;
    long long out_ls, out_is, out_os1, out_os2;
    int n_tiles;   // tiles of the whole pass (a CTA loops over tile = blockIdx.x + k * gridDim.x)
    int prefetch_tiles;  // > 0: every CTA asks L2 to prefetch the input of tile + prefetch_tiles (the tile the
                         // CTA slot it occupies will run next), decoupling HBM latency from SM occupancy
    int n_inner;   // lines along the inner index
    int n_inner_last_o2;  // > 0: the last o2 block holds only this many lines (ragged last column block of the blocked
                          // intermediate layout of real transforms); 0: every block holds n_inner
    int n_o2;      // outer index o = o1*n_o2 + o2
    int tiles_per_outer;
    // divisions by tiles_per_outer / n_o2 as multiply-high + shift (fast_div below): every thread of every tile splits
    // its tile index (and the prefetched tile's) into (o1, o2, i0)
    unsigned div_tpo_m, div_tpo_s, div_o2_m, div_o2_s;
    int tw4_shift, tw4_mask;
    int inverse;   // backward transform: conjugate on load and on store
    // V_CC_PEER: output line index k goes to buffer peer[k >> peer_shift] at line (k & peer_mask);
    // the o1/o2/i offsets and out_ls apply inside that buffer
    void *peer[MAX_PEERS];
    int peer_shift, peer_mask;
};

constexpr int ilog2c(int v) { return v <= 1 ? 0 : 1 + ilog2c(v >> 1); }
// x / d for 0 <= x < 2^31 and a divisor known at plan time: d == 1 -> m = 0; else l = ceil(log2 d), m = ceil(2^(31+l) / d)
// (fits 32 bits), s = l - 1 and x / d == umulhi(x, m) >> s exactly (Granlund & Montgomery, N = 31)
__host__ __device__ inline void fast_div_make(int d, unsigned *m, unsigned *s) {
    if (d <= 1) { *m = 0; *s = 0; return; }
    int l = 0;
    while ((1ll << l) < d) ++l;
    *m = (unsigned)(((1ull << (31 + l)) + (unsigned long long)d - 1) / (unsigned long long)d);
    *s = (unsigned)(l - 1);
}
__device__ __forceinline__ int fast_div(int x, unsigned m, unsigned s) {
    return m ? (int)(__umulhi((unsigned)x, m) >> s) : x;
}
// offset (in complex elements) of stage s's table inside TileParams::tw; stage_tw_offset(L, R, S) = total size
__host__ __device__ constexpr int stage_tw_offset(int L, int R, int s) {
    int off = 0, m = L / R;
    for (int q = 1; q < s; ++q) { off += (R - 1) * m; m /= R; }
    return off;
}

template <typename T, int L_, int R_, int W_, int VAR_> struct TileTraits {
    static constexpr int L = L_, R = R_, W = W_, VAR = VAR_;
    static constexpr int LOG_L = ilog2c(L), LOG_R = ilog2c(R), LOG_W = ilog2c(W);
    static constexpr int S = (LOG_L + LOG_R - 1) / LOG_R;          // number of stages
    static constexpr int R_LAST = L >> (LOG_R * (S - 1));          // radix of the last stage
    static constexpr int T_LINE = L / R;                            // threads per line
    static constexpr int LOG_TL = ilog2c(T_LINE);
    static constexpr int THREADS = T_LINE * W;
    static constexpr bool LOAD_ROW =
        (VAR == V_RR || VAR == V_RC || VAR == V_RR_R2C || VAR == V_RR_C2R || VAR == V_RC_PEER);
    static constexpr bool STORE_ROW = (VAR == V_RR || VAR == V_RR_R2C || VAR == V_RR_C2R);
    static constexpr bool NEED_SMEM = (S > 1) || (VAR == V_RR_R2C) || (VAR == V_RR_C2R);
    static constexpr int SMEM_BYTES = NEED_SMEM ? L * W * (int)sizeof(cplx<T>) : 0;
    // swizzle group width: 8 x 16 B or 16 x 8 B = 128 B = all 32 banks.  Pure column passes need no
    // swizzle at all: their lanes run along w first, so every quarter-warp (16-byte elements) or
    // half-warp (8-byte elements) touches one contiguous, aligned 128-byte run whatever the row; the
    // index arithmetic then folds into immediate offsets.
    static constexpr bool SWIZZLED = LOAD_ROW || STORE_ROW || (W * (int)sizeof(cplx<T>) < 128);
    static constexpr int SWZ_BITS = !SWIZZLED ? 0 : (sizeof(cplx<T>) == 16 ? 3 : 4);
    // two CTAs per SM (so that one tile's HBM phase overlaps the other's butterflies) whenever the tile's
    // data fits 32 registers per thread: caps the kernel at 64 registers
    static constexpr int MIN_CTAS_BY_THREADS = 1024 / THREADS < 1 ? 1 : (1024 / THREADS > 8 ? 8 : 1024 / THREADS);
    static constexpr int MIN_CTAS_BY_SMEM = SMEM_BYTES == 0 ? 8 : (200 * 1024 / SMEM_BYTES < 1 ? 1 : 200 * 1024 / SMEM_BYTES);
    // ... and at 128 registers (as many CTAs as then fit the 64 K register file) when the payload is 64 registers
    static constexpr int MIN_CTAS_BY_REGS128 = 512 / THREADS < 1 ? 1 : 512 / THREADS;
    static constexpr int MIN_CTAS_SMALL = MIN_CTAS_BY_THREADS < MIN_CTAS_BY_SMEM ? MIN_CTAS_BY_THREADS : MIN_CTAS_BY_SMEM;
    static constexpr int MIN_CTAS_BIG = MIN_CTAS_BY_REGS128 < MIN_CTAS_BY_SMEM ? MIN_CTAS_BY_REGS128 : MIN_CTAS_BY_SMEM;
    static constexpr int MIN_CTAS = (int)sizeof(cplx<T>) * R > 128 ? MIN_CTAS_BIG : MIN_CTAS_SMALL;
    static_assert(R * T_LINE == L, "R must divide L");
    static_assert(S == 1 || R_LAST <= R, "bad stage split");
};

// XOR-fold of x's bits above the low group, in groups of BITS
template <int BITS, int TOTAL_BITS> __device__ __forceinline__ int swz_fold(int idx) {
    if constexpr (BITS == 0) {
        return 0;
    } else {
        int f = 0;
#pragma unroll
        for (int s = BITS; s < TOTAL_BITS; s += BITS) f ^= (idx >> s);
        return f & ((1 << BITS) - 1);
    }
}

template <typename T> __device__ __forceinline__ cplx<T> ld_cplx(const cplx<T> *p) { return __ldg(p); }
// Backward transforms run through the forward butterflies as conj(F(conj(x))): the imaginary part's sign bit is
// flipped on load and on store (one integer XOR each; `mask` is 0 for forward transforms, the sign bit otherwise).
__device__ __forceinline__ double2 conj_if(double2 x, unsigned mask) {
    x.y = __hiloint2double(__double2hiint(x.y) ^ (int)mask, __double2loint(x.y));
    return x;
}
__device__ __forceinline__ float2 conj_if(float2 x, unsigned mask) {
    x.y = __int_as_float(__float_as_int(x.y) ^ (int)mask);
    return x;
}
// Tile data is touched exactly once per pass.  Streaming cache hints (ld.global.cs / st.global.cs) and
// ld.global.nc.L1::no_allocate were measured on B200 and made the 512^3 passes 5-10 % slower / no different,
// so tile data uses the default policies (DESIGN.md "Experiments").
template <typename T> __device__ __forceinline__ cplx<T> ld_data(const cplx<T> *p) { return __ldg(p); }
// L2-coherent load (ld.global.cg): for tile data that other SMs or peer GPUs wrote earlier in the SAME kernel (the fused
// slab kernel); the read-only / L1 path could return lines cached before that write
template <typename T> __device__ __forceinline__ cplx<T> ld_data_cg(const cplx<T> *p) { return __ldcg(p); }
template <typename T> __device__ __forceinline__ void st_data(cplx<T> *p, cplx<T> v) { *p = v; }

// ---------------------------------------------------------------------------------------------
// All radix stages of one tile.  In: thread (w1, u1) holds x[u1 + d*T_LINE], d < R, of line w1 in v[].
// Out: thread (wl, ul) holds X[k], k = (ul + b*T_LINE) + q*(L/RL), in v[b*RL + q]  (B*RL == R).
// (w1,u1) follows the load style, (wl,ul) the store style.
// ---------------------------------------------------------------------------------------------
template <typename T, int L, int R, int W, int VAR>
__device__ __forceinline__ void tile_stages(cplx<T> *v, cplx<T> *sm, const cplx<T> *__restrict__ tw, const int w1,
                                            const int u1, const int w_col, const int u_col, const int wl, const int ul) {
    using TR = TileTraits<T, L, R, W, VAR>;
    constexpr int S = TR::S;
    constexpr int LOG_R = TR::LOG_R, LOG_W = TR::LOG_W, LOG_L = TR::LOG_L;
    constexpr int T_LINE = TR::T_LINE;
    constexpr int R_LAST = TR::R_LAST;
    constexpr int IDX_BITS = LOG_L + LOG_W;
    constexpr int SB = TR::SWZ_BITS;
    if constexpr (S > 1) {
        fft_reg<T, R>(v);
        // twiddle w_L^(d*u1), then park at position d*m_1 + u1.  The tables are stored per stage and transposed
        // ([d][lo]): with a plain w_L^k table the lanes of a warp (which run along u1 on ROW loads) would gather with a
        // stride of d elements - up to 32 different 128-byte lines per warp load, 15 such loads per thread, several
        // times the tile's own data loads in L1 cycles; transposed, every load is one contiguous run.
        {
#pragma unroll
            for (int d = 1; d < R; ++d) v[d] = cmul(v[d], ld_cplx<T>(tw + (d - 1) * T_LINE + u1));
            const int base = (u1 << LOG_W) | w1;
            const int fb = swz_fold<SB, IDX_BITS>(base);
#pragma unroll
            for (int d = 0; d < R; ++d) {
                const int dbits = (d * T_LINE) << LOG_W;
                sm[(base | dbits) ^ fb ^ swz_fold<SB, IDX_BITS>(dbits)] = v[d];
            }
        }
        __syncthreads();

        // -------------------------------------------------------------- middle stages (radix R)
#pragma unroll
        for (int s = 2; s < S; ++s) {
            // m_s = L / R^s
            const int log_ms = LOG_L - LOG_R * s;
            const int ms = 1 << log_ms;
            const int u = u_col, w = w_col;
            const int lo = u & (ms - 1);
            const int hi = u >> log_ms;
            const int pos = (hi << (log_ms + LOG_R)) | lo;
            const int base = (pos << LOG_W) | w;
            const int fb = swz_fold<SB, IDX_BITS>(base);
#pragma unroll
            for (int d = 0; d < R; ++d) {
                const int dbits = (d << log_ms) << LOG_W;
                v[d] = sm[(base | dbits) ^ fb ^ swz_fold<SB, IDX_BITS>(dbits)];
            }
            fft_reg<T, R>(v);
            // twiddle w_{m_{s-1}}^(d*lo) = w_L^(d*lo*R^(s-1)): stage table [d][lo]
            const cplx<T> *tws = tw + stage_tw_offset(L, R, s) + lo;
#pragma unroll
            for (int d = 1; d < R; ++d) v[d] = cmul(v[d], ld_cplx<T>(tws + ((d - 1) << log_ms)));
#pragma unroll
            for (int d = 0; d < R; ++d) {
                const int dbits = (d << log_ms) << LOG_W;
                sm[(base | dbits) ^ fb ^ swz_fold<SB, IDX_BITS>(dbits)] = v[d];
            }
            __syncthreads();
        }

        // -------------------------------------------------------------- last stage: smem -> registers
        {
            constexpr int B = R / R_LAST;  // butterflies per thread
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const int jp = ul + b * T_LINE;  // output-order index of this butterfly, in [0, L/R_LAST)
                // digits of jp are k_1 (lowest) .. k_{S-1}; position = sum k_i * m_i
                int pos = 0;
#pragma unroll
                for (int i = 1; i < S; ++i) {
                    const int ki = (jp >> (LOG_R * (i - 1))) & (R - 1);
                    pos |= ki << (LOG_L - LOG_R * i);
                }
                const int base = (pos << LOG_W) | wl;
                const int fb = swz_fold<SB, IDX_BITS>(base);
#pragma unroll
                for (int n = 0; n < R_LAST; ++n) {
                    const int dbits = n << LOG_W;
                    v[b * R_LAST + n] = sm[(base | dbits) ^ fb ^ swz_fold<SB, IDX_BITS>(dbits)];
                }
                fft_reg<T, R_LAST>(v + b * R_LAST);
            }
        }
    } else {
        fft_reg<T, R>(v);
    }
}

// ---------------------------------------------------------------------------------------------
// Store of one tile's results (all variants but r2c).  The line's global output index of local index k
// is KADD + KMUL*k (KMUL = KADD = trivial for single-CTA tiles; a cluster CTA owns the residue class
// k1 = KADD of a length KMUL*L line).
// ---------------------------------------------------------------------------------------------
template <typename T, int L, int R, int W, int VAR>
__device__ __forceinline__ void tile_store(const cplx<T> *v, const TileParams &p, const int o1, const int o2, const int i0,
                                           const int wl, const int ul, const int kmul, const int kadd,
                                           const long long out_shift = 0) {
    using TR = TileTraits<T, L, R, W, VAR>;
    using C = cplx<T>;
    constexpr int S = TR::S;
    constexpr int T_LINE = TR::T_LINE;
    constexpr int B = (S > 1) ? R / TR::R_LAST : 1;
    constexpr int RL = (S > 1) ? TR::R_LAST : R;
    const unsigned cmask = p.inverse ? 0x80000000u : 0u;
    const int n_inner = (p.n_inner_last_o2 > 0 && o2 == p.n_o2 - 1) ? p.n_inner_last_o2 : p.n_inner;
    const bool ok = (i0 + wl) < n_inner;
    const long long off = o1 * p.out_os1 + o2 * p.out_os2 + (long long)(i0 + wl) * p.out_is + out_shift;
    C *dst = reinterpret_cast<C *>(p.out) + off;
    if constexpr (VAR == V_CC_TW) {
        // four-step twiddles by recurrence in fp64: the exponent of w_N is affine in (b, q),
        //   m = i*(kadd + kmul*ul) + b*(i*kmul*T_LINE) + q*(i*kmul*L/RL),
        // so a few two-level table lookups per thread replace two per point (the table traffic through
        // L1/L2 was 4x the pass's HBM bytes in fp32).  fp32 data: recurrence over b and q (<= B+RL fp64
        // multiplications deep, error ~1e-15); fp64 data: one lookup per b, recurrence over q only
        // (<= RL-1 <= 7 multiplications: a few eps on the twiddle, far inside the 10*log2(N)*eps budget).
        constexpr bool RECUR_B = sizeof(T) == 4;
        auto lookup = [&](long long m) {
            const double2 wh = __ldg(p.tw4_hi + (m >> p.tw4_shift));
            const double2 wlw = __ldg(p.tw4_lo + (m & p.tw4_mask));
            double2 r;
            r.x = wh.x * wlw.x - wh.y * wlw.y;
            r.y = wh.x * wlw.y + wh.y * wlw.x;
            return r;
        };
        auto mul = [](double2 a, double2 b2) {
            double2 r;
            r.x = a.x * b2.x - a.y * b2.y;
            r.y = a.x * b2.y + a.y * b2.x;
            return r;
        };
        const long long i = i0 + wl;
        double2 wb = lookup(i * (long long)(kadd + kmul * ul));
        double2 sb, sq;
        sb.x = sq.x = 1.0;
        sb.y = sq.y = 0.0;
        if (B > 1 && RECUR_B) sb = lookup(i * (long long)(kmul * T_LINE));
        if (RL > 1) sq = lookup(i * (long long)(kmul * (L / RL)));
        // output addresses advance by fixed strides over b and q: two 64-bit adds per store, no multiplies
        const long long step_q = (long long)(kmul * (L / RL)) * p.out_ls, step_b = (long long)(kmul * T_LINE) * p.out_ls;
        C *pb = dst + (long long)(kadd + kmul * ul) * p.out_ls;
        // fp32 data, short q chains (RL <= 4; 512 = 16*16*2 has RL = 2): the recurrence over b stays in fp64, the step
        // over q is one fp32 multiplication per point by the twiddle step rounded once - half the fp64 work and half the
        // fp64 -> fp32 conversions (a slow pipe) of carrying every twiddle in fp64, at the cost of one more fp32 rounding
        constexpr bool Q_IN_FP32 = sizeof(T) == 4 && RL <= 4;
        const C sqf = mk<T>((T)sq.x, (T)sq.y);
#pragma unroll
        for (int b = 0; b < B; ++b) {
            double2 wq = wb;
            C wqf = mk<T>((T)wb.x, (T)wb.y);
            C *pq = pb;
#pragma unroll
            for (int q = 0; q < RL; ++q) {
                C x = v[b * RL + q];
                if constexpr (Q_IN_FP32) {
                    x = cmul(x, wqf);
                    if (q + 1 < RL) wqf = cmul(wqf, sqf);
                } else if constexpr (sizeof(T) == 4) {
                    // fp32 data: the twiddle is carried in fp64 (recurrence) and rounded once to fp32 for the multiply;
                    // converting the data to fp64 and back instead costs four conversions per point on a slow pipe
                    x = cmul(x, mk<T>((T)wq.x, (T)wq.y));
                } else {
                    const double xr = (double)x.x * wq.x - (double)x.y * wq.y;
                    const double xi = (double)x.x * wq.y + (double)x.y * wq.x;
                    x.x = (T)xr; x.y = (T)xi;
                }
                x = conj_if(x, cmask);
                if (ok) st_data<T>(pq, x);
                pq += step_q;
                if (!Q_IN_FP32 && q + 1 < RL) wq = mul(wq, sq);
            }
            pb += step_b;
            if (b + 1 < B) {
                if constexpr (RECUR_B) wb = mul(wb, sb);
                else wb = lookup(i * (long long)(kadd + kmul * (ul + (b + 1) * T_LINE)));
            }
        }
    } else if constexpr (VAR == V_CC_PEER || VAR == V_RC_PEER) {
#pragma unroll
        for (int b = 0; b < B; ++b)
#pragma unroll
            for (int q = 0; q < RL; ++q) {
                const int k = kadd + kmul * ((ul + b * T_LINE) + q * (L / RL));
                const C x = conj_if(v[b * RL + q], cmask);
                C *pd = reinterpret_cast<C *>(p.peer[k >> p.peer_shift]) + off;
                if (ok) pd[(long long)(k & p.peer_mask) * p.out_ls] = x;
            }
    } else {
        // output index k = kadd + kmul * ((ul + b*T_LINE) + q*(L/RL)): the address advances by fixed strides over b and
        // q, two 64-bit adds per store instead of a 64-bit multiply each (the store phase was ~19 instructions per STG)
        const long long step_q = (long long)(kmul * (L / RL)) * p.out_ls, step_b = (long long)(kmul * T_LINE) * p.out_ls;
        C *pb = dst + (long long)(kadd + kmul * ul) * p.out_ls;
#pragma unroll
        for (int b = 0; b < B; ++b) {
            C *pq = pb;
#pragma unroll
            for (int q = 0; q < RL; ++q) {
                if (ok) st_data<T>(pq, conj_if(v[b * RL + q], cmask));
                pq += step_q;
            }
            pb += step_b;
        }
    }
}

// DATA_CG: tile data is read with ld.global.cg (see ld_data_cg).  in_shift / out_shift: element offsets added to the
// pass's input / output addresses (chunked passes of the fused slab kernel).
template <typename T, int L, int R, int W, int VAR, bool DATA_CG = false>
__device__ __forceinline__ void fft_tile_body(const TileParams &p, const int tile, unsigned char *smem_raw,
                                              const long long in_shift = 0, const long long out_shift = 0) {
    using TR = TileTraits<T, L, R, W, VAR>;
    using C = cplx<T>;
    constexpr int S = TR::S;
    constexpr int LOG_W = TR::LOG_W, LOG_L = TR::LOG_L;
    constexpr int T_LINE = TR::T_LINE, LOG_TL = TR::LOG_TL;
    constexpr int IDX_BITS = LOG_L + LOG_W;
    constexpr int SB = TR::SWZ_BITS;

    C *sm = reinterpret_cast<C *>(smem_raw);

    const int t = threadIdx.x;
    // thread -> (line w within tile, slot u within line), for the two access styles
    const int w_col = t & (W - 1), u_col = t >> LOG_W;
    const int u_row = t & (T_LINE - 1), w_row = t >> LOG_TL;

    const int o = fast_div(tile, p.div_tpo_m, p.div_tpo_s);
    const int i0 = (tile - o * p.tiles_per_outer) * W;
    const int o1 = fast_div(o, p.div_o2_m, p.div_o2_s), o2 = o - o1 * p.n_o2;
    const C *__restrict__ gin = reinterpret_cast<const C *>(p.in) + o1 * p.in_os1 + o2 * p.in_os2 + in_shift;
    C *__restrict__ gout = reinterpret_cast<C *>(p.out) + o1 * p.out_os1 + o2 * p.out_os2 + out_shift;
    const C *__restrict__ tw = reinterpret_cast<const C *>(p.tw);
    const unsigned cmask = (p.inverse && VAR != V_RR_C2R) ? 0x80000000u : 0u;

    C v[R];

    // ------------------------------------------------------------------ stage 1: HBM -> registers
    const int w1 = TR::LOAD_ROW ? w_row : w_col;
    const int u1 = TR::LOAD_ROW ? u_row : u_col;
    {
        // lines past the end of a ragged last tile re-read the last valid line (their results are never stored):
        // no predicates or zero fills on the load path
        const int n_inner = (p.n_inner_last_o2 > 0 && o2 == p.n_o2 - 1) ? p.n_inner_last_o2 : p.n_inner;
        const int wi = min(i0 + w1, n_inner - 1);
        const C *src = gin + (long long)wi * p.in_is + (long long)u1 * p.in_ls;
        const long long step = (long long)T_LINE * p.in_ls;
#pragma unroll
        for (int d = 0; d < R; ++d) {
            if constexpr (DATA_CG) v[d] = conj_if(ld_data_cg<T>(src), cmask);
            else v[d] = conj_if(ld_data<T>(src), cmask);
            src += step;
        }
    }

    // ------------------------------------------------------------------ L2 prefetch of a future tile (after this tile's own loads are in flight)
    if (p.prefetch_tiles > 0) {
        const int ft = tile + p.prefetch_tiles;
        if (ft < p.n_tiles) {
            const int fo = fast_div(ft, p.div_tpo_m, p.div_tpo_s);
            const int fi0 = (ft - fo * p.tiles_per_outer) * W;
            const int fo1 = fast_div(fo, p.div_o2_m, p.div_o2_s), fo2 = fo - fo1 * p.n_o2;
            // whole tiles only: never touch addresses past the array (a ragged last block is not prefetched)
            if (fi0 + W <= p.n_inner && !(p.n_inner_last_o2 > 0 && fo2 == p.n_o2 - 1)) {
                const C *fin = reinterpret_cast<const C *>(p.in) + fo1 * p.in_os1 + fo2 * p.in_os2 + (long long)fi0 * p.in_is;
                constexpr int ELT = (int)sizeof(C);
                if constexpr (TR::LOAD_ROW) {
                    // W lines of L contiguous elements each
                    constexpr int CH_LINE = (L * ELT + 127) / 128;  // 128-byte chunks per line
                    for (int ch = t; ch < W * CH_LINE; ch += TR::THREADS) {
                        const int wl2 = ch / CH_LINE, cc = ch - wl2 * CH_LINE;
                        const char *a = reinterpret_cast<const char *>(fin + (long long)wl2 * p.in_is) + cc * 128;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
                    }
                } else {
                    // L rows of W contiguous elements each
                    constexpr int CH_ROW = (W * ELT + 127) / 128;
                    for (int ch = t; ch < L * CH_ROW; ch += TR::THREADS) {
                        const int r = ch / CH_ROW, cc = ch - r * CH_ROW;
                        const char *a = reinterpret_cast<const char *>(fin + (long long)r * p.in_ls) + cc * 128;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
                    }
                }
            }
        }
    }


    if constexpr (VAR == V_RR_C2R) {
        // ---- even/odd pre-pass: from the half spectrum X[0..L] of 2L reals build
        //   Z'[k] = (X[k] + conj X[L-k]) + i * conj(w_{2L}^k) * (X[k] - conj X[L-k])      ( = 2 Z[k] )
        // whose unnormalised inverse transform z has x[2j] = Re z[j], x[2j+1] = Im z[j]  (the launch always
        // runs as a backward transform: conj here, forward butterflies, conj on store).
        const bool ok = (i0 + w1) < p.n_inner;
        T xl_re = (T)0;  // Re X[L], needed for k = 0 only
        if (u1 == 0 && ok) xl_re = ld_cplx<T>(gin + (long long)(i0 + w1) * p.in_is + (long long)L * p.in_ls).x;
#pragma unroll
        for (int d = 0; d < R; ++d) {
            const int idx = ((u1 + d * T_LINE) << LOG_W) | w1;
            sm[idx ^ swz_fold<SB, IDX_BITS>(idx)] = v[d];
        }
        __syncthreads();
        const C *tw2 = reinterpret_cast<const C *>(p.tw_aux);
#pragma unroll
        for (int d = 0; d < R; ++d) {
            const int k = u1 + d * T_LINE;
            C zp;
            if (k == 0) {
                zp = mk<T>(v[d].x + xl_re, v[d].x - xl_re);
            } else {
                const int ib = ((L - k) << LOG_W) | w1;
                const C a = v[d];
                const C b = cconj(sm[ib ^ swz_fold<SB, IDX_BITS>(ib)]);
                const C sum = cadd(a, b);
                const C tt = cmul(csub(a, b), cconj(ld_cplx<T>(tw2 + k)));
                zp = mk<T>(sum.x - tt.y, sum.y + tt.x);
            }
            v[d] = mk<T>(zp.x, -zp.y);  // conj: backward transform through the forward butterflies (tile_store conjugates back)
        }
        __syncthreads();  // every partner has been read before the stages overwrite shared memory
    }

    const int wl = TR::STORE_ROW ? w_row : w_col;
    const int ul = TR::STORE_ROW ? u_row : u_col;
    tile_stages<T, L, R, W, VAR>(v, sm, tw, w1, u1, w_col, u_col, wl, ul);

    // After the last stage thread (wl, ul) holds, for b in [0,B) and q in [0,R_LAST):
    //   X[k],  k = (ul + b*T_LINE) + q*(L/R_LAST),  in v[b*R_LAST + q]
    constexpr int B = (S > 1) ? R / TR::R_LAST : 1;
    constexpr int RL = (S > 1) ? TR::R_LAST : R;

    if constexpr (VAR == V_RR_R2C) {
        // ---- even/odd post-pass (cf. rdft/ct-hc2c.c:59-70): the line held L complex = 2L reals.
        // Z -> smem in natural order, then X[k] = E + w^k O, X[L-k] = conj(E - w^k O),
        // E = (Z[k] + conj Z[L-k])/2, O = -i (Z[k] - conj Z[L-k])/2, w = w_{2L}.
        if constexpr (S > 1) __syncthreads();
#pragma unroll
        for (int b = 0; b < B; ++b)
#pragma unroll
            for (int q = 0; q < RL; ++q) {
                const int k = (ul + b * T_LINE) + q * (L / RL);
                const int idx = (k << LOG_W) | wl;
                sm[idx ^ swz_fold<SB, IDX_BITS>(idx)] = v[b * RL + q];
            }
        __syncthreads();
        const bool ok = (i0 + wl) < p.n_inner;
        C *dst = gout + (long long)(i0 + wl) * p.out_is;
        const C *tw2 = reinterpret_cast<const C *>(p.tw_aux);
        // pairs k in [0, L/2]: thread takes k = ul + j*T_LINE (j < R/2) and, for ul == 0, also k = L/2
        static_assert(L >= 2 && R >= 2, "r2c fast path needs at least 4 reals per line");
#pragma unroll
        for (int j = 0; j <= R / 2; ++j) {
            const int k = (j < R / 2) ? ul + j * T_LINE : L / 2;
            if (j == R / 2 && ul != 0) break;
            const int k2 = (L - k) & (L - 1);  // Z[L] == Z[0]
            const int ia = (k << LOG_W) | wl, ib = (k2 << LOG_W) | wl;
            const C a = sm[ia ^ swz_fold<SB, IDX_BITS>(ia)];
            const C bq = cconj(sm[ib ^ swz_fold<SB, IDX_BITS>(ib)]);
            const C e = mk<T>((T)0.5 * (a.x + bq.x), (T)0.5 * (a.y + bq.y));
            const C d = csub(a, bq);
            const C od = mk<T>((T)0.5 * d.y, (T)-0.5 * d.x);  // -i/2 * (a - b)
            const C wo = cmul(od, ld_cplx<T>(tw2 + k));
            if (ok) {
                dst[(long long)k * p.out_ls] = cadd(e, wo);
                dst[(long long)(L - k) * p.out_ls] = cconj(csub(e, wo));
            }
        }
    } else {
        tile_store<T, L, R, W, VAR>(v, p, o1, o2, i0, wl, ul, 1, 0, out_shift);
    }
}

// The kernel: CTA b transforms tiles b, b + gridDim.x, ...  With gridDim.x == n_tiles (the default) every
// CTA owns one tile; a smaller grid makes the pass persistent on a bounded number of SMs, which is how
// the slab plans keep an NVLink-bound exchange pass and an HBM-bound local pass running side by side.
template <typename T, int L, int R, int W, int VAR>
__global__ void __launch_bounds__(TileTraits<T, L, R, W, VAR>::THREADS, TileTraits<T, L, R, W, VAR>::MIN_CTAS)
fft_tile_kernel(const TileParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // Only the exchange passes of slab plans are ever launched with fewer CTAs than tiles.  Everything else runs
    // exactly one tile per CTA and is compiled without the loop: keeping loop state alive across the body costs
    // registers at the 64-register cap (measured: ~4 % on the 512^3 strided passes).
    if constexpr (VAR != V_CC_PEER) {
        fft_tile_body<T, L, R, W, VAR>(p, (int)blockIdx.x, smem_raw);
    } else {
        // static tile = blockIdx.x + k * gridDim.x assignment on purpose: a capped exchange pass should drain early on
        // the fast SMs so that the pass it overlaps with finds them free (dynamic tickets kept every capped CTA resident
        // to the end: 2 x B200, 512^3: 1.61 vs 1.40 ms)
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            fft_tile_body<T, L, R, W, VAR>(p, tile, smem_raw);
            if (TileTraits<T, L, R, W, VAR>::NEED_SMEM && tile + (int)gridDim.x < p.n_tiles) __syncthreads();
        }
    }
}

// distributed shared memory through 32-bit shared::cluster addresses (cheaper in registers than
// generic pointers from cluster.map_shared_rank)
__device__ __forceinline__ unsigned dsmem_map(unsigned local_addr, unsigned cta_rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ double2 dsmem_ld(unsigned addr, double2 *) {
    double2 v;
    asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float2 dsmem_ld(unsigned addr, float2 *) {
    float2 v;
    asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void dsmem_st(unsigned addr, double2 v) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ void dsmem_st(unsigned addr, float2 v) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Cluster pass for long strided axes: a thread-block cluster of CL CTAs transforms W lines of length
// L = CL*LL, so that a tile keeps 128-byte HBM segments (W*sizeof(complex) = 128 B) although L*W
// complex do not fit one SM.  CTA c of the cluster
//   A. loads, for its share of the positions j (LL/CL of them), the CL rows d*LL + j straight from HBM
//      (each still a 128-byte segment), does the CL-point DFT over d in registers, multiplies by
//      w_L^(j*k1) and scatters y_k1[j] into CTA k1's shared memory (distributed shared memory stores),
//   B. after one cluster barrier transforms the length-LL sub-sequence it now owns (k1 = c) with the
//      ordinary stage pipeline and stores X[c + CL*k'], k' < LL.
// (decimation in frequency: X[k1 + CL*k'] = DFT_LL( DFT_CL_d(x[d*LL+j])(k1) * w_L^(j*k1) )[k'])
// ---------------------------------------------------------------------------------------------
template <typename T, int LL, int CL, int R, int W, int VAR>
__global__ void __launch_bounds__(TileTraits<T, LL, R, W, VAR>::THREADS, TileTraits<T, LL, R, W, VAR>::MIN_CTAS)
fft_cluster_kernel(const TileParams p) {
    namespace cg = cooperative_groups;
    using TR = TileTraits<T, LL, R, W, VAR>;
    using C = cplx<T>;
    static_assert(!TR::LOAD_ROW && !TR::STORE_ROW, "cluster passes are column passes");
    static_assert(R % CL == 0 && TR::S > 1, "cross stage needs R/CL items per thread");
    constexpr int LOG_W = TR::LOG_W, LOG_L = TR::LOG_L;
    constexpr int T_LINE = TR::T_LINE;
    constexpr int IDX_BITS = LOG_L + LOG_W;
    constexpr int SB = TR::SWZ_BITS;
    constexpr int NI = R / CL;  // cross-stage items per thread

    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *sm = reinterpret_cast<C *>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    const int n_clusters = (int)gridDim.x / CL;
    const unsigned sm_base = (unsigned)__cvta_generic_to_shared(sm);

    const int t = threadIdx.x;
    const int w = t & (W - 1), u = t >> LOG_W;
    const C *__restrict__ tw = reinterpret_cast<const C *>(p.tw);       // w_LL^k
    const C *__restrict__ twL = reinterpret_cast<const C *>(p.tw_aux);  // w_L^k
    const unsigned cmask = p.inverse ? 0x80000000u : 0u;

    // every CTA of the cluster must be running before its shared memory is written remotely: arrive now,
    // wait just before the first scatter (the HBM loads in between hide the barrier)
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    bool first = true;
    for (int tile = (int)blockIdx.x / CL; tile < p.n_tiles; tile += n_clusters) {
        const int o = fast_div(tile, p.div_tpo_m, p.div_tpo_s);
        const int i0 = (tile - o * p.tiles_per_outer) * W;
        const int o1 = fast_div(o, p.div_o2_m, p.div_o2_s), o2 = o - o1 * p.n_o2;
        const C *__restrict__ gin = reinterpret_cast<const C *>(p.in) + o1 * p.in_os1 + o2 * p.in_os2;
        C v[R];
        // ---- A: HBM -> registers -> cross-CTA radix-CL stage -> owners' shared memory
        {
            const int wi = min(i0 + w, p.n_inner - 1);  // ragged last tile: re-read the last valid line (never stored)
            const int j0 = c * (LL / CL) + u;
            const C *src = gin + (long long)wi * p.in_is + (long long)j0 * p.in_ls;
#pragma unroll
            for (int it = 0; it < NI; ++it)
#pragma unroll
                for (int d = 0; d < CL; ++d)
                    v[it * CL + d] = conj_if(ld_data<T>(src + (long long)(d * LL + it * T_LINE) * p.in_ls), cmask);
            if (first) { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); first = false; }
#pragma unroll
            for (int it = 0; it < NI; ++it) {
                const int j = j0 + it * T_LINE;
                const int idx = (j << LOG_W) | w;
                const unsigned sa = sm_base + (unsigned)(idx ^ swz_fold<SB, IDX_BITS>(idx)) * (unsigned)sizeof(C);
                fft_reg<T, CL>(v + it * CL);
#pragma unroll
                for (int k1 = 0; k1 < CL; ++k1) {
                    C y = v[it * CL + k1];
                    if (k1 > 0) y = cmul(y, ld_cplx<T>(twL + j * k1));
                    dsmem_st(dsmem_map(sa, k1), y);
                }
            }
        }
        cluster.sync();
        // ---- B: local length-LL transform of sub-sequence k1 = c, then store rows c + CL*k'
#pragma unroll
        for (int d = 0; d < R; ++d) {
            const int idx = ((u + d * T_LINE) << LOG_W) | w;
            v[d] = sm[idx ^ swz_fold<SB, IDX_BITS>(idx)];
        }
        tile_stages<T, LL, R, W, VAR>(v, sm, tw, w, u, w, u, w, u);
        tile_store<T, LL, R, W, VAR>(v, p, o1, o2, i0, w, u, CL, c);
        // persistent launch: nobody may scatter the next tile into a CTA that still works on this one
        if (tile + n_clusters < p.n_tiles) cluster.sync();
    }
    if (first) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace fftb200
