// generic_kernels.cuh — any-length, any-layout path (global-memory Stockham, naive radix-p).
//
// Covers what the tiled power-of-two kernels do not: prime factors other than 2 (the
// reference's own test shapes are 3, 5, {3,2,2}, {3,3,2}: test/fft_test.rg:143,247,328,349),
// element strides != 1, and misaligned base pointers.  It is a CUDA path like the others
// (no CPU fallback), just not a roofline one: O(L * sum of prime factors) work per line.
//
// Data is first gathered into a packed [outer][L][inner] complex work buffer, every prime
// factor p of L is one ping-pong stage
//     y[(j/Ns)*Ns*p + j%Ns + q*Ns] = sum_t x[j + t*L/p] * w_L^(t * ((j%Ns)*L/(Ns*p) + q*L/p))
// (twiddle and p-point DFT merged into one table lookup per term; fp64 table and fp64
// accumulation for both precisions), and the result is scattered to the caller's layout.
// Same definitions as the fast path: fftw-3.3.8/doc/reference.texi:1863-1894, dft/generic.c.
#pragma once
#include "butterfly.cuh"

namespace fftb200 {

struct GenLayout {
    int nd;               // number of index levels (batch first), <= 4
    long long n[4];       // extents, slowest first
    long long stride[4];  // element strides in the user's buffer
};

// user buffer (real or complex, any strides) -> packed complex
template <typename T, bool REAL_IN>
__global__ void gen_gather_kernel(const void *__restrict__ in, cplx<T> *__restrict__ packed, GenLayout lay,
                                  long long total, int swap_reim) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        long long rem = e, off = 0;
#pragma unroll
        for (int d = 3; d >= 0; --d) {
            if (d < lay.nd) {
                const long long q = rem / lay.n[d];
                off += (rem - q * lay.n[d]) * lay.stride[d];
                rem = q;
            }
        }
        cplx<T> x;
        if constexpr (REAL_IN) {
            x.x = reinterpret_cast<const T *>(in)[off];
            x.y = (T)0;
        } else {
            // component loads: the user's base may be aligned to sizeof(T) only
            const T *q = reinterpret_cast<const T *>(in) + 2 * off;
            x.x = q[0];
            x.y = q[1];
        }
        if (swap_reim) { T s = x.x; x.x = x.y; x.y = s; }
        packed[e] = x;
    }
}

// packed complex -> user buffer (complex, any strides)
template <typename T>
__global__ void gen_scatter_kernel(const cplx<T> *__restrict__ packed, cplx<T> *__restrict__ out, GenLayout lay,
                                   long long total, int swap_reim) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        long long rem = e, off = 0;
#pragma unroll
        for (int d = 3; d >= 0; --d) {
            if (d < lay.nd) {
                const long long q = rem / lay.n[d];
                off += (rem - q * lay.n[d]) * lay.stride[d];
                rem = q;
            }
        }
        cplx<T> x = packed[e];
        if (swap_reim) { T s = x.x; x.x = x.y; x.y = s; }
        T *q = reinterpret_cast<T *>(out) + 2 * off;  // component stores, see gather
        q[0] = x.x;
        q[1] = x.y;
    }
}

// inverse real transforms on the generic path: the packed half spectrum [..][n_last/2+1] of the user's buffer is
// completed to the full Hermitian array  X[k] = conj(X[-k mod n])  (every index negated), re/im swapped for the backward
// direction like gen_gather_kernel does, and the real result is the (swapped-back) real part of the complex transform
template <typename T>
__global__ void gen_gather_herm_kernel(const cplx<T> *__restrict__ in, cplx<T> *__restrict__ packed, GenLayout lay, long long total) {
    const long long n_last = lay.n[lay.nd - 1], half = n_last / 2;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        long long idx[4] = {0, 0, 0, 0};
        long long rem = e;
#pragma unroll
        for (int d = 3; d >= 0; --d) {
            if (d < lay.nd) {
                const long long q = rem / lay.n[d];
                idx[d] = rem - q * lay.n[d];
                rem = q;
            }
        }
        const bool mirror = idx[lay.nd - 1] > half;
        long long off = 0;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            if (d < lay.nd) {
                long long j = idx[d];
                if (mirror && d > 0) j = (lay.n[d] - j) % lay.n[d];  // (d == 0 is the batch index)
                off += j * lay.stride[d];
            }
        }
        const T *q = reinterpret_cast<const T *>(in) + 2 * off;
        cplx<T> x;
        x.x = q[0];
        x.y = mirror ? -q[1] : q[1];
        cplx<T> r;
        r.x = x.y;  // swap: backward transform through the forward stages
        r.y = x.x;
        packed[e] = r;
    }
}

template <typename T>
__global__ void gen_scatter_real_kernel(const cplx<T> *__restrict__ packed, T *__restrict__ out, GenLayout lay, long long total) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        long long rem = e, off = 0;
#pragma unroll
        for (int d = 3; d >= 0; --d) {
            if (d < lay.nd) {
                const long long q = rem / lay.n[d];
                off += (rem - q * lay.n[d]) * lay.stride[d];
                rem = q;
            }
        }
        out[off] = packed[e].y;  // real part after swapping back
    }
}

// ---- long real lines (2 Lh reals, Lh too long for one shared-memory tile): the even/odd passes of the half-length
// scheme as global-memory kernels around the two-pass complex transform of z[m] = x[2m] + i x[2m+1]
// (mixed_kernel.cuh has the shared-memory form and the formulas; fftw-3.3.8/rdft/ct-hc2c.c:59-70 on the CPU path).
// `lay` enumerates the LINES (all indices but the last axis): offset of a line's first complex element.
__device__ __forceinline__ long long gen_line_offset(const GenLayout &lay, long long line) {
    long long rem = line, off = 0;
#pragma unroll
    for (int d = 3; d >= 0; --d) {
        if (d < lay.nd) {
            const long long q = rem / lay.n[d];
            off += (rem - q * lay.n[d]) * lay.stride[d];
            rem = q;
        }
    }
    return off;
}

// in place on the half-spectrum lines: Z[0 .. Lh) -> X[0 .. Lh];  tw[k] = w_{2 Lh}^k, k in [0, Lh/2]
template <typename T>
__global__ void gen_r2c_post_kernel(cplx<T> *__restrict__ data, GenLayout lay, long long total, int Lh,
                                    const cplx<T> *__restrict__ tw) {
    using C = cplx<T>;
    const int pairs = Lh / 2 + 1;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long line = e / pairs;
        const int k = (int)(e - line * pairs);
        C *z = data + gen_line_offset(lay, line);
        const C za = z[k], zb = z[k == 0 ? 0 : Lh - k];
        const C wk = __ldg(tw + k);
        const C ev = mk<T>((T)0.5 * (za.x + zb.x), (T)0.5 * (za.y - zb.y));
        const C b = mk<T>((T)0.5 * (za.y + zb.y), (T)-0.5 * (za.x - zb.x));
        const C wb = cmul(wk, b);
        z[k] = mk<T>(ev.x + wb.x, ev.y + wb.y);
        z[Lh - k] = mk<T>(ev.x - wb.x, -(ev.y - wb.y));
    }
}

// half spectrum X[0 .. Lh] (lines of lay_in) -> Z'[0 .. Lh) (lines of lay_out), the spectrum whose backward complex
// transform is x[2m] + i x[2m+1]
template <typename T>
__global__ void gen_c2r_pre_kernel(const cplx<T> *__restrict__ in, cplx<T> *__restrict__ out, GenLayout lay_in,
                                   GenLayout lay_out, long long total, int Lh, const cplx<T> *__restrict__ tw) {
    using C = cplx<T>;
    const int pairs = Lh / 2 + 1;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long line = e / pairs;
        const int k = (int)(e - line * pairs);
        const C *x = in + gen_line_offset(lay_in, line);
        C *z = out + gen_line_offset(lay_out, line);
        const C xa = x[k], xb = x[Lh - k];
        const C wk = __ldg(tw + k);
        const C ee = mk<T>(xa.x + xb.x, xa.y - xb.y);
        const C dd = mk<T>(xa.x - xb.x, xa.y + xb.y);
        const C oo = mk<T>(dd.x * wk.x + dd.y * wk.y, dd.y * wk.x - dd.x * wk.y);
        z[k] = mk<T>(ee.x - oo.y, ee.y + oo.x);
        if (k > 0) z[Lh - k] = mk<T>(ee.x + oo.y, oo.x - ee.y);
    }
}

// in-place scaling of the elements of a layout (normalisation helper); COMPONENTS = 2 for complex, 1 for real elements
template <typename T, int COMPONENTS>
__global__ void gen_scale_kernel(T *__restrict__ data, GenLayout lay, long long total, T factor) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        long long rem = e, off = 0;
#pragma unroll
        for (int d = 3; d >= 0; --d) {
            if (d < lay.nd) {
                const long long q = rem / lay.n[d];
                off += (rem - q * lay.n[d]) * lay.stride[d];
                rem = q;
            }
        }
        T *q = data + COMPONENTS * off;
#pragma unroll
        for (int c = 0; c < COMPONENTS; ++c) q[c] *= factor;
    }
}

// one radix-p stage along the middle index of packed [outer][L][inner]
template <typename T>
__global__ void gen_stage_kernel(const cplx<T> *__restrict__ x, cplx<T> *__restrict__ y,
                                 const double2 *__restrict__ tw /* w_L^k */, long long outer, int L, long long inner,
                                 int p, int Ns) {
    const long long total = outer * (long long)L * inner;
    const int Lp = L / p;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long i = e % inner;
        const long long r1 = e / inner;
        const int jq = (int)(r1 % L);  // enumerates (j, q): j = jq % Lp, q = jq / Lp
        const long long o = r1 / L;
        const int j = jq % Lp, q = jq / Lp;
        const int k = j % Ns;
        // exponent step c = k*L/(Ns*p) + q*L/p  (mod L)
        const long long c = ((long long)k * (L / (Ns * p)) + (long long)q * Lp) % L;
        const cplx<T> *src = x + (o * L + j) * inner + i;
        double sr = 0.0, si = 0.0;
        long long ex = 0;
        for (int t = 0; t < p; ++t) {
            const cplx<T> a = src[(long long)t * Lp * inner];
            const double2 w = __ldg(tw + ex);
            sr += (double)a.x * w.x - (double)a.y * w.y;
            si += (double)a.x * w.y + (double)a.y * w.x;
            ex += c;
            if (ex >= L) ex -= L;
        }
        const long long j0 = (long long)(j / Ns) * Ns * p + k;
        cplx<T> r;
        r.x = (T)sr;
        r.y = (T)si;
        y[(o * L + j0 + (long long)q * Ns) * inner + i] = r;
    }
}

// ---- Bluestein (chirp-z) for lengths with a large prime factor: a length-L DFT as a cyclic convolution of length
// M = 2^m >= 2L-1 (fftw-3.3.8/dft/bluestein.c does the same on the CPU path).  With c[j] = exp(-i*pi*j^2/L):
//   X[k] = c[k] * sum_j (x[j] c[j]) * conj(c[k-j])
// pre:  a[o][j][i] = x[o][j][i] * c[j] (j < L), 0 (L <= j < M);  then FFT_M, times Bhat = FFT_M(conj chirp, wrapped),
// inverse FFT_M;  post: y[o][k][i] = a[o][k][i] * c[k] / M (k < L).  Packed [outer][L or M][inner] layouts.
template <typename T>
__global__ void gen_blu_pre_kernel(const cplx<T> *__restrict__ x, cplx<T> *__restrict__ a, const double2 *__restrict__ chirp,
                                   long long outer, int L, int M, long long inner) {
    const long long total = outer * (long long)M * inner;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long i = e % inner, r1 = e / inner;
        const int j = (int)(r1 % M);
        const long long o = r1 / M;
        cplx<T> r;
        r.x = r.y = (T)0;
        if (j < L) {
            const cplx<T> v = x[(o * L + j) * inner + i];
            const double2 c = __ldg(chirp + j);
            r.x = (T)((double)v.x * c.x - (double)v.y * c.y);
            r.y = (T)((double)v.x * c.y + (double)v.y * c.x);
        }
        a[e] = r;
    }
}

template <typename T>
__global__ void gen_blu_mul_kernel(cplx<T> *__restrict__ a, const cplx<T> *__restrict__ bhat, long long outer, int M,
                                   long long inner) {
    const long long total = outer * (long long)M * inner;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int k = (int)((e / inner) % M);
        a[e] = cmul(a[e], __ldg(bhat + k));
    }
}

template <typename T>
__global__ void gen_blu_post_kernel(const cplx<T> *__restrict__ a, cplx<T> *__restrict__ y, const double2 *__restrict__ chirp,
                                    long long outer, int L, int M, long long inner, double scale) {
    const long long total = outer * (long long)L * inner;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long i = e % inner, r1 = e / inner;
        const int k = (int)(r1 % L);
        const long long o = r1 / L;
        const cplx<T> v = a[(o * M + k) * inner + i];
        const double2 c = __ldg(chirp + k);
        cplx<T> r;
        r.x = (T)(((double)v.x * c.x - (double)v.y * c.y) * scale);
        r.y = (T)(((double)v.x * c.y + (double)v.y * c.x) * scale);
        y[e] = r;
    }
}

// keep the first Lc of every L-long line: packed [lines][L] -> packed [lines][Lc]
template <typename T>
__global__ void gen_truncate_kernel(const cplx<T> *__restrict__ x, cplx<T> *__restrict__ y, long long lines, int L,
                                    int Lc) {
    const long long total = lines * Lc;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const long long ln = e / Lc;
        const int k = (int)(e - ln * Lc);
        y[e] = x[ln * L + k];
    }
}

}  // namespace fftb200
