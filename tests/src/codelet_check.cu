// codelet_check.cu — host-side check of the in-register DFTs of radix_dft.cuh against the O(N^2) definition in long
// double (test infrastructure; built and run by tests/test_codelets.py, no GPU needed).
#include <cmath>
#include <cstdio>

#include "../../regent-fft-arjun_b200/csrc/radix_dft.cuh"

using namespace fftb200;

template <typename T, int N> static double check() {
    cplx<T> a[N];
    long double xr[N], xi[N];
    unsigned s = 12345u + N;
    for (int i = 0; i < N; ++i) {
        s = s * 1664525u + 1013904223u;
        xr[i] = (T)((double)(s >> 8) / (1 << 24) - 0.5);
        s = s * 1664525u + 1013904223u;
        xi[i] = (T)((double)(s >> 8) / (1 << 24) - 0.5);
        a[i].x = (T)xr[i];
        a[i].y = (T)xi[i];
    }
    Dft<T, N>::run(a);
    const long double PI = 3.141592653589793238462643383279502884L;
    long double num = 0, den = 0;
    for (int k = 0; k < N; ++k) {
        long double sr = 0, si = 0;
        for (int n = 0; n < N; ++n) {
            const long double ang = -2 * PI * ((n * k) % N) / N;
            const long double c = cosl(ang), sn = sinl(ang);
            sr += xr[n] * c - xi[n] * sn;
            si += xr[n] * sn + xi[n] * c;
        }
        num += (a[k].x - sr) * (a[k].x - sr) + (a[k].y - si) * (a[k].y - si);
        den += sr * sr + si * si;
    }
    return (double)sqrtl(num / den);
}

template <int N> static int report() {
    const double e64 = check<double, N>(), e32 = check<float, N>();
    printf("{\"n\": %d, \"rel_l2_fp64\": %.3e, \"rel_l2_fp32\": %.3e}\n", N, e64, e32);
    return (e64 < 4e-16 && e32 < 3e-7) ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += report<2>(); bad += report<3>(); bad += report<4>(); bad += report<5>(); bad += report<6>(); bad += report<7>();
    bad += report<8>(); bad += report<9>(); bad += report<10>(); bad += report<11>(); bad += report<12>(); bad += report<13>(); bad += report<14>(); bad += report<15>();
    bad += report<16>();
    return bad;
}
