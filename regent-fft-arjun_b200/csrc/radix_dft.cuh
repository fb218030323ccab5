// radix_dft.cuh — in-register forward DFTs (sign -1, natural order in and out) of size
//     2, 3, 4, 5, 7, 11, 13    written out directly (odd sizes: the conjugate-pair form, (N-1)/2 real root pairs)
//     6, 8, 9, 10, 12, 14, 15, 16   Cooley-Tukey products A x B of the above with the w_N twiddles as compile-time constants
// for the mixed-radix shared-memory kernel (mixed_kernel.cuh).  All loops unroll and every index is a compile-time
// constant, so an array argument lives in registers.  They stand where FFTW's n1_3 ... n1_16 / t1_* codelets stand on the
// reference's CPU path (fftw-3.3.8/dft/scalar/codelets/).  Host-callable too, so tests/ can check them without a GPU.
#pragma once
#include "butterfly.cuh"

#define FFTB200_HD __host__ __device__ __forceinline__

namespace fftb200 {

#include "dft_roots.inc"

template <typename T> FFTB200_HD cplx<T> r_mk(T x, T y) { cplx<T> r; r.x = x; r.y = y; return r; }
template <typename C> FFTB200_HD C r_add(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <typename C> FFTB200_HD C r_sub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }

// x * w_N^e, e a compile-time constant after unrolling
template <typename T, int N> FFTB200_HD cplx<T> mul_root(cplx<T> x, int e) {
    e %= N;
    if (e == 0) return x;
    if (4 * e == N) return r_mk<T>(x.y, -x.x);       // -i
    if (2 * e == N) return r_mk<T>(-x.x, -x.y);      // -1
    if (4 * e == 3 * N) return r_mk<T>(-x.y, x.x);   // +i
    const T c = (T)root_cos<N>(e), s = (T)root_sin<N>(e);  // w = c - i s
    return r_mk<T>(x.x * c + x.y * s, x.y * c - x.x * s);
}

template <typename T, int N> struct Dft;

template <typename T> struct Dft<T, 2> {
    static FFTB200_HD void run(cplx<T> *a) {
        const cplx<T> t = a[0];
        a[0] = r_add(t, a[1]);
        a[1] = r_sub(t, a[1]);
    }
};

template <typename T> struct Dft<T, 4> {
    static FFTB200_HD void run(cplx<T> *a) {
        const cplx<T> t0 = r_add(a[0], a[2]), t1 = r_sub(a[0], a[2]);
        const cplx<T> t2 = r_add(a[1], a[3]), d = r_sub(a[1], a[3]);
        const cplx<T> t3 = r_mk<T>(d.y, -d.x);  // -i (a1 - a3)
        a[0] = r_add(t0, t2);
        a[1] = r_add(t1, t3);
        a[2] = r_sub(t0, t2);
        a[3] = r_sub(t1, t3);
    }
};

// odd primes: X[q] = a0 + sum_m ( (a_m + a_{P-m}) cos(2 pi m q / P) - i (a_m - a_{P-m}) sin(2 pi m q / P) ), m = 1 .. (P-1)/2
template <typename T, int P> FFTB200_HD void dft_odd(cplx<T> *a) {
    constexpr int H = (P - 1) / 2;
    cplx<T> s[H], d[H];
#pragma unroll
    for (int m = 1; m <= H; ++m) {
        s[m - 1] = r_add(a[m], a[P - m]);
        d[m - 1] = r_sub(a[m], a[P - m]);
    }
    const cplx<T> x0 = a[0];
    cplx<T> sum = x0;
#pragma unroll
    for (int m = 0; m < H; ++m) sum = r_add(sum, s[m]);
    a[0] = sum;
#pragma unroll
    for (int q = 1; q <= H; ++q) {
        T re = x0.x, im = x0.y, ur = (T)0, ui = (T)0;
#pragma unroll
        for (int m = 1; m <= H; ++m) {
            const T c = (T)root_cos<P>((m * q) % P), sn = (T)root_sin<P>((m * q) % P);
            re += c * s[m - 1].x;
            im += c * s[m - 1].y;
            ur += sn * d[m - 1].x;
            ui += sn * d[m - 1].y;
        }
        a[q] = r_mk<T>(re + ui, im - ur);      // ... - i (ur + i ui)
        a[P - q] = r_mk<T>(re - ui, im + ur);
    }
}
template <typename T> struct Dft<T, 3> { static FFTB200_HD void run(cplx<T> *a) { dft_odd<T, 3>(a); } };
template <typename T> struct Dft<T, 5> { static FFTB200_HD void run(cplx<T> *a) { dft_odd<T, 5>(a); } };
template <typename T> struct Dft<T, 7> { static FFTB200_HD void run(cplx<T> *a) { dft_odd<T, 7>(a); } };
template <typename T> struct Dft<T, 11> { static FFTB200_HD void run(cplx<T> *a) { dft_odd<T, 11>(a); } };
template <typename T> struct Dft<T, 13> { static FFTB200_HD void run(cplx<T> *a) { dft_odd<T, 13>(a); } };

// N = A * B.  With n = B n1 + n2 and k = k1 + A k2:  w_N^(n k) = w_A^(n1 k1) * w_N^(n2 k1) * w_B^(n2 k2)
//   1. for every n2: A-point DFT over n1              -> y[k1][n2]
//   2. y[k1][n2] *= w_N^(n2 k1)
//   3. for every k1: B-point DFT over n2              -> X[k1 + A k2]
template <typename T, int A, int B> FFTB200_HD void dft_product(cplx<T> *a) {
    constexpr int N = A * B;
    cplx<T> y[N];
#pragma unroll
    for (int n2 = 0; n2 < B; ++n2) {
        cplx<T> t[A];
#pragma unroll
        for (int n1 = 0; n1 < A; ++n1) t[n1] = a[B * n1 + n2];
        Dft<T, A>::run(t);
#pragma unroll
        for (int k1 = 0; k1 < A; ++k1) y[k1 * B + n2] = t[k1];
    }
#pragma unroll
    for (int k1 = 0; k1 < A; ++k1) {
        cplx<T> t[B];
#pragma unroll
        for (int n2 = 0; n2 < B; ++n2) t[n2] = mul_root<T, N>(y[k1 * B + n2], k1 * n2);
        Dft<T, B>::run(t);
#pragma unroll
        for (int k2 = 0; k2 < B; ++k2) a[k1 + A * k2] = t[k2];
    }
}
template <typename T> struct Dft<T, 6> { static FFTB200_HD void run(cplx<T> *a) { dft_product<T, 2, 3>(a); } };
template <typename T> struct Dft<T, 8> { static FFTB200_HD void run(cplx<T> *a) { dft_product<T, 2, 4>(a); } };
template <typename T> struct Dft<T, 9> { static FFTB200_HD void run(cplx<T> *a) { dft_product<T, 3, 3>(a); } };
template <typename T> struct Dft<T, 10> { static FFTB200_HD void run(cplx<T> *a) { dft_product<T, 2, 5>(a); } };
template <typename T> struct Dft<T, 12> { static FFTB200_HD void run(cplx<T> *a) { dft_product<T, 3, 4>(a); } };
template <typename T> struct Dft<T, 14> { static FFTB200_HD void run(cplx<T> *a) { dft_product<T, 2, 7>(a); } };
template <typename T> struct Dft<T, 15> { static FFTB200_HD void run(cplx<T> *a) { dft_product<T, 3, 5>(a); } };
template <typename T> struct Dft<T, 16> { static FFTB200_HD void run(cplx<T> *a) { dft_product<T, 4, 4>(a); } };

}  // namespace fftb200
