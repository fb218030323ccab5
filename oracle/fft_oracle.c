/*
 * oracle/fft_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into libfft_b200,
 * never called by the product path; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may use it).
 *
 * A plain-C99, double-precision restatement of what the reference's CPU branch
 * computes when fft.rg calls FFTW 3.3.8:
 *
 *   src/fft.rg:319   fftw_plan_dft(dim, n, in, out, FFTW_FORWARD, FFTW_ESTIMATE)
 *   src/fft.rg:313   fftw_plan_dft_r2c(dim, n, in, out, FFTW_ESTIMATE)
 *   src/fft.rg:500   fftw_plan_many_dft(dim-1, n_batch, n[dim-1], in, n_batch, 1, i_dist, ...)
 *   src/fft.rg:483   fftw_plan_many_dft_r2c(...)
 *   src/fft.rg:605,608  fftw_execute_dft_r2c / fftw_execute_dft
 *
 * Definitions followed (all under /root/reference/fftw-3.3.8):
 *   doc/reference.texi:1863-1894   1-D DFT: Y[k] = sum_j X[j] exp(sign*2*pi*i*j*k/n), unnormalised
 *   doc/reference.texi:2342-2372   multi-dimensional DFT = separable product, row-major
 *   doc/reference.texi:2423-2438   r2c output: last dimension cut to n/2+1
 *   api/plan-many-dft.c:26-51, api/mktensor-rowmajor.c:33-40   advanced (embed/stride/dist) layout
 *   api/plan-many-dft-r2c.c:24-49, api/rdft2-pad.c:24-38       r2c embeds (NULL => n, last n/2+1 on the complex side)
 *   dft/ct.c:34-58                 Cooley-Tukey decimation in time: n = r*m, r sub-DFTs of size m then m twiddled butterflies of size r
 *   dft/generic.c                  O(n^2) DFT for prime sizes
 *   rdft/rank-geq2-rdft2.c:40-53   multi-dim r2c = r2c along the last dim, then c2c over the others on n/2+1 columns
 *   kernel/trig.c:57-80            twiddles by octant reduction to [0, pi/4] before cos/sin
 *
 * FFTW picks among many algebraically equivalent factorizations (its planner);
 * this file fixes one (smallest prime factor first).  Results agree with FFTW to
 * rounding (a few 1e-16 relative L2), which tests/test_oracle.py pins against
 * the reference's real FFTW build (oracle/_ref) and its own known answers.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cplx;

/* kernel/trig.c:57-80 — exp(2*pi*i*m/n), octant-reduced */
static void oracle_cexp(long m, long n, double *c_out, double *s_out)
{
    static const double K2PI = 6.2831853071795864769252867665590057683943388;
    unsigned octant = 0;
    long quarter_n = n;
    double theta, c, s, t;

    n += n; n += n;
    m += m; m += m;
    if (m < 0) m += n;
    if (m > n - m) { m = n - m; octant |= 4; }
    if (m - quarter_n > 0) { m = m - quarter_n; octant |= 2; }
    if (m > quarter_n - m) { m = quarter_n - m; octant |= 1; }

    theta = (K2PI * (double)m) / (double)n;
    c = cos(theta); s = sin(theta);
    if (octant & 1) { t = c; c = s; s = t; }
    if (octant & 2) { t = c; c = -s; s = t; }
    if (octant & 4) { s = -s; }
    *c_out = c; *s_out = s;
}

void oracle_twiddle(long m, long n, int sign, double *out2)
{
    double c, s;
    oracle_cexp(((m % n) + n) % n, n, &c, &s);
    out2[0] = c;
    out2[1] = sign < 0 ? -s : s;
}

static long smallest_factor(long n)
{
    long f;
    if (n % 2 == 0) return 2;
    for (f = 3; f * f <= n; f += 2)
        if (n % f == 0) return f;
    return n;
}

/* W[k] = exp(sign*2*pi*i*k/N) for the top-level size; w_n^j = W[j*(N/n)]. */
typedef struct { long N; const cplx *W; } twtab;

/* dft/ct.c:34-58 (DIT): out is contiguous, in has element stride `is`. */
static void dft_rec(long n, const cplx *in, ptrdiff_t is, cplx *out, const twtab *tw, cplx *scratch)
{
    long r, m, p, q, k, step;
    if (n == 1) { out[0] = in[0]; return; }
    r = smallest_factor(n);
    m = n / r;
    step = tw->N / n;
    if (m > 1)
        for (p = 0; p < r; ++p)
            dft_rec(m, in + p * is, is * r, out + p * m, tw, scratch);
    else
        for (p = 0; p < r; ++p) out[p] = in[p * is];

    /* m butterflies of size r with twiddles w_n^(p*k), then w_r^(p*q) (dft/generic.c for the r-point part) */
    for (k = 0; k < m; ++k) {
        for (p = 0; p < r; ++p) {
            const cplx w = tw->W[((p * k) % n) * step];
            const cplx y = out[p * m + k];
            scratch[p].re = y.re * w.re - y.im * w.im;
            scratch[p].im = y.re * w.im + y.im * w.re;
        }
        for (q = 0; q < r; ++q) {
            double sr = 0.0, si = 0.0;
            for (p = 0; p < r; ++p) {
                const cplx w = tw->W[((p * q) % r) * m * step];
                sr += scratch[p].re * w.re - scratch[p].im * w.im;
                si += scratch[p].re * w.im + scratch[p].im * w.re;
            }
            out[k + q * m].re = sr;
            out[k + q * m].im = si;
        }
    }
}

typedef struct { long n; cplx *W; cplx *line_in; cplx *line_out; cplx *scratch; } plan1d;

static int plan1d_init(plan1d *p, long n, int sign)
{
    long k;
    p->n = n;
    p->W = (cplx *)malloc(sizeof(cplx) * (size_t)n);
    p->line_in = (cplx *)malloc(sizeof(cplx) * (size_t)n);
    p->line_out = (cplx *)malloc(sizeof(cplx) * (size_t)n);
    p->scratch = (cplx *)malloc(sizeof(cplx) * (size_t)n);
    if (!p->W || !p->line_in || !p->line_out || !p->scratch) return -1;
    for (k = 0; k < n; ++k) {
        double c, s;
        oracle_cexp(k, n, &c, &s);
        p->W[k].re = c;
        p->W[k].im = sign < 0 ? -s : s;
    }
    return 0;
}

static void plan1d_free(plan1d *p)
{
    free(p->W); free(p->line_in); free(p->line_out); free(p->scratch);
}

/* transform the contiguous line p->line_in into p->line_out */
static void plan1d_run(plan1d *p)
{
    twtab tw; tw.N = p->n; tw.W = p->W;
    dft_rec(p->n, p->line_in, 1, p->line_out, &tw, p->scratch);
}

/* in-place 1-D DFTs along dimension d of a row-major block `dims[rank]` embedded in
 * `embed[rank]` with element stride `stride` (api/mktensor-rowmajor.c:33-40). */
static int transform_axis(cplx *data, int rank, const long *dims, const long *embed, long stride, int d, int sign)
{
    long pitch[8], idx[8], lines = 1, l, j;
    int i;
    plan1d p;
    if (dims[d] == 1) return 0;
    if (plan1d_init(&p, dims[d], sign)) return -1;
    pitch[rank - 1] = stride;
    for (i = rank - 2; i >= 0; --i) pitch[i] = pitch[i + 1] * embed[i + 1];
    for (i = 0; i < rank; ++i) if (i != d) lines *= dims[i];
    for (l = 0; l < lines; ++l) {
        long rem = l, base = 0;
        for (i = rank - 1; i >= 0; --i) {
            if (i == d) { idx[i] = 0; continue; }
            idx[i] = rem % dims[i]; rem /= dims[i];
            base += idx[i] * pitch[i];
        }
        for (j = 0; j < dims[d]; ++j) p.line_in[j] = data[base + j * pitch[d]];
        plan1d_run(&p);
        for (j = 0; j < dims[d]; ++j) data[base + j * pitch[d]] = p.line_out[j];
    }
    plan1d_free(&p);
    return 0;
}

static long block_elems(int rank, const long *embed) { long t = 1; int i; for (i = 0; i < rank; ++i) t *= embed[i]; return t; }

/* api/plan-many-dft.c:26-51 semantics.  embeds NULL => n.  Returns 0 on success. */
int oracle_dft_many(int rank, const int *n, int howmany,
                    const double *in, const int *inembed, int istride, int idist,
                    double *out, const int *onembed, int ostride, int odist, int sign)
{
    long dims[8], ie[8], oe[8], idx[8];
    long total = 1, e;
    int i, b, d;
    const cplx *cin = (const cplx *)in;
    cplx *cout = (cplx *)out;
    if (rank < 1 || rank > 8 || howmany < 0) return 4;
    for (i = 0; i < rank; ++i) {
        if (n[i] < 1) return 4;
        dims[i] = n[i];
        ie[i] = inembed ? inembed[i] : n[i];
        oe[i] = onembed ? onembed[i] : n[i];
        total *= dims[i];
    }
    (void)block_elems;
    for (b = 0; b < howmany; ++b) {
        const cplx *src = cin + (long)b * idist;
        cplx *dst = cout + (long)b * odist;
        /* copy into the output layout, then transform each axis in place */
        for (e = 0; e < total; ++e) {
            long rem = e, io = 0, oo = 0;
            for (i = rank - 1; i >= 0; --i) { idx[i] = rem % dims[i]; rem /= dims[i]; }
            for (i = 0; i < rank; ++i) { io = io * ie[i] + idx[i]; oo = oo * oe[i] + idx[i]; }
            dst[oo * ostride] = src[io * istride];
        }
        for (d = rank - 1; d >= 0; --d)
            if (transform_axis(dst, rank, dims, oe, ostride, d, sign)) return 2;
    }
    return 0;
}

/* api/plan-many-dft-r2c.c:24-49: real input (embed in reals), complex output with last
 * dimension n_last/2+1 (embed in complexes); rdft/rank-geq2-rdft2.c:40-53 ordering. */
int oracle_dft_r2c_many(int rank, const int *n, int howmany,
                        const double *in, const int *inembed, int istride, int idist,
                        double *out, const int *onembed, int ostride, int odist)
{
    long dims[8], cdims[8], ie[8], oe[8], idx[8];
    long lines = 1, l, j;
    int i, b, d;
    cplx *cout = (cplx *)out;
    plan1d p;
    if (rank < 1 || rank > 8 || howmany < 0) return 4;
    for (i = 0; i < rank; ++i) {
        if (n[i] < 1) return 4;
        dims[i] = cdims[i] = n[i];
        ie[i] = inembed ? inembed[i] : n[i];
        oe[i] = onembed ? onembed[i] : (i == rank - 1 ? n[i] / 2 + 1 : n[i]);
    }
    cdims[rank - 1] = n[rank - 1] / 2 + 1;
    for (i = 0; i < rank - 1; ++i) lines *= dims[i];
    if (plan1d_init(&p, dims[rank - 1], -1)) return 2;
    for (b = 0; b < howmany; ++b) {
        const double *src = in + (long)b * idist;
        cplx *dst = cout + (long)b * odist;
        for (l = 0; l < lines; ++l) {
            long rem = l, io = 0, oo = 0;
            for (i = rank - 2; i >= 0; --i) { idx[i] = rem % dims[i]; rem /= dims[i]; }
            for (i = 0; i < rank - 1; ++i) { io = io * ie[i] + idx[i]; oo = oo * oe[i] + idx[i]; }
            io *= ie[rank - 1]; oo *= oe[rank - 1];
            for (j = 0; j < dims[rank - 1]; ++j) { p.line_in[j].re = src[(io + j) * istride]; p.line_in[j].im = 0.0; }
            plan1d_run(&p);
            for (j = 0; j < cdims[rank - 1]; ++j) dst[(oo + j) * ostride] = p.line_out[j];
        }
        for (d = rank - 2; d >= 0; --d)
            if (transform_axis(dst, rank, cdims, oe, ostride, d, -1)) { plan1d_free(&p); return 2; }
    }
    plan1d_free(&p);
    return 0;
}

/* basic interface: api/plan-dft.c:23, api/plan-dft-r2c.c:23 (packed row-major, batch 1) */
int oracle_dft(int rank, const int *n, const double *in, double *out, int sign)
{
    return oracle_dft_many(rank, n, 1, in, 0, 1, 0, out, 0, 1, 0, sign);
}

int oracle_dft_r2c(int rank, const int *n, const double *in, double *out)
{
    return oracle_dft_r2c_many(rank, n, 1, in, 0, 1, 0, out, 0, 1, 0);
}
