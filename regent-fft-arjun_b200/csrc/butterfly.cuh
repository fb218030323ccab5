// butterfly.cuh — in-register forward DFTs of size 2/4/8/16 (sign -1, natural-order output).
//
// These are the per-thread building blocks of the shared-memory Stockham passes in
// tile_kernel.cuh.  They play the role FFTW's generated codelets play on the reference's CPU
// path (fftw-3.3.8/dft/scalar/codelets/n1_*.c, t1_*.c), but are written by hand as a radix-2
// decimation-in-time recursion with the trivial twiddles (+-1, +-i, (1-i)/sqrt2, ...) folded in.
// The backward transform is obtained by swapping re/im at load and store (tile_kernel.cuh).
#pragma once
#include <cuda_runtime.h>

namespace fftb200 {

template <typename T> struct Vec2;
template <> struct Vec2<float>  { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

template <typename T> using cplx = typename Vec2<T>::type;

template <typename T> __device__ __forceinline__ cplx<T> mk(T x, T y) { cplx<T> r; r.x = x; r.y = y; return r; }
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }
// fp32 complex add/sub as ONE packed instruction (Blackwell FADD2: add.f32x2 / sub.f32x2 on a 64-bit register
// pair).  The fp32 passes are bound by instruction issue, not by HBM, and adds are the bulk of a butterfly.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
    return *reinterpret_cast<float2 *>(&r);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
    return *reinterpret_cast<float2 *>(&r);
}
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    C r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}
// a * (-i)
template <typename C> __device__ __forceinline__ C mul_mi(C a) { C r; r.x = a.y; r.y = -a.x; return r; }
template <typename C> __device__ __forceinline__ C cconj(C a) { a.y = -a.y; return a; }

template <typename T> __device__ __forceinline__ void fft2(cplx<T> &a, cplx<T> &b) {
    cplx<T> t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

// v[0..3] at element stride S inside a register array
template <typename T, int S> __device__ __forceinline__ void fft4(cplx<T> *v) {
    cplx<T> t0 = cadd(v[0], v[2 * S]);
    cplx<T> t1 = csub(v[0], v[2 * S]);
    cplx<T> t2 = cadd(v[S], v[3 * S]);
    cplx<T> t3 = mul_mi(csub(v[S], v[3 * S]));
    v[0] = cadd(t0, t2);
    v[S] = cadd(t1, t3);
    v[2 * S] = csub(t0, t2);
    v[3 * S] = csub(t1, t3);
}

template <typename T, int S> __device__ __forceinline__ void fft8(cplx<T> *v) {
    constexpr T H = (T)0.70710678118654752440084436210484903928483593768847;
    // evens and odds (each a radix-4 at stride 2S)
    fft4<T, 2 * S>(v);
    fft4<T, 2 * S>(v + S);
    // e[k] = v[2kS], o[k] = v[(2k+1)S];  X[k] = e[k] + w8^k o[k], X[k+4] = e[k] - w8^k o[k]
    cplx<T> o0 = v[S];
    cplx<T> o1 = v[3 * S];
    cplx<T> o2 = v[5 * S];
    cplx<T> o3 = v[7 * S];
    cplx<T> e0 = v[0], e1 = v[2 * S], e2 = v[4 * S], e3 = v[6 * S];
    // w8^1 = (1 - i)/sqrt2 ; w8^2 = -i ; w8^3 = (-1 - i)/sqrt2
    cplx<T> t1 = mk<T>((o1.x + o1.y) * H, (o1.y - o1.x) * H);
    cplx<T> t2 = mul_mi(o2);
    cplx<T> t3 = mk<T>((o3.y - o3.x) * H, -(o3.x + o3.y) * H);
    v[0] = cadd(e0, o0);
    v[4 * S] = csub(e0, o0);
    v[S] = cadd(e1, t1);
    v[5 * S] = csub(e1, t1);
    v[2 * S] = cadd(e2, t2);
    v[6 * S] = csub(e2, t2);
    v[3 * S] = cadd(e3, t3);
    v[7 * S] = csub(e3, t3);
}

template <typename T, int S> __device__ __forceinline__ void fft16(cplx<T> *v) {
    constexpr T H = (T)0.70710678118654752440084436210484903928483593768847;
    constexpr T C1 = (T)0.92387953251128675612818318939678828682241662586364;  // cos(pi/8)
    constexpr T S1 = (T)0.38268343236508977172845998403039886676134456248563;  // sin(pi/8)
    fft8<T, 2 * S>(v);
    fft8<T, 2 * S>(v + S);
    cplx<T> e[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { e[k] = v[2 * k * S]; o[k] = v[(2 * k + 1) * S]; }
    // w16^k = exp(-i*pi*k/8)
    o[1] = cmul(o[1], mk<T>(C1, -S1));
    o[2] = mk<T>((o[2].x + o[2].y) * H, (o[2].y - o[2].x) * H);
    o[3] = cmul(o[3], mk<T>(S1, -C1));
    o[4] = mul_mi(o[4]);
    o[5] = cmul(o[5], mk<T>(-S1, -C1));
    o[6] = mk<T>((o[6].y - o[6].x) * H, -(o[6].x + o[6].y) * H);
    o[7] = cmul(o[7], mk<T>(-C1, -S1));
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[k * S] = cadd(e[k], o[k]);
        v[(k + 8) * S] = csub(e[k], o[k]);
    }
}

// size-R forward DFT on v[0..R-1] (contiguous registers), natural order in and out
template <typename T, int R> __device__ __forceinline__ void fft_reg(cplx<T> *v) {
    if constexpr (R == 2) fft2<T>(v[0], v[1]);
    else if constexpr (R == 4) fft4<T, 1>(v);
    else if constexpr (R == 8) fft8<T, 1>(v);
    else if constexpr (R == 16) fft16<T, 1>(v);
    else static_assert(R == 1, "unsupported radix");
}

}  // namespace fftb200
