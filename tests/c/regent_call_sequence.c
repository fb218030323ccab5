/* regent_call_sequence.c — a plain-C consumer of include/fft_b200.h that replays, call for call, what the
 * patched src/fft.rg does on a GPU processor: make_plan_gpu (src/fft.rg:195-258),
 * make_plan_gpu_batch (:336-414), execute_plan (:543-611) and destroy_plan (:624-645), on the shapes and
 * constant inputs of the reference's own test program (test/fft_test.rg:138-389) and checks the known
 * answers.  C99 on purpose: this is how Terra's includec sees the header.
 * exit 0 = all good, 77 = no CUDA device (nothing run), 1 = failure. */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fft_b200.h"

typedef struct { double re, im; } c64;
typedef struct { float re, im; } c32;

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { printf("FAIL: " __VA_ARGS__); printf("\n"); ++failures; } } while (0)

/* iface.plan as patched: the handle is stored BY VALUE next to the FFTW pointers (src/fft.rg:48-65) */
typedef struct { void *p; void *float_p; fftb200_handle b200_p; unsigned address_space; } iface_plan;

static void *to_device(const void *host, size_t bytes) {
    void *d = NULL;
    if (cudaMalloc(&d, bytes) != cudaSuccess) { printf("cudaMalloc failed\n"); exit(1); }
    cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice);
    return d;
}

/* make_plan_gpu + execute_plan + destroy_plan for a dense region of `dim` dimensions */
static void run_basic(int dim, const int *n, fftb200_type type, const void *in, size_t in_bytes, void *out, size_t out_bytes) {
    iface_plan plan;
    memset(&plan, 0, sizeof plan);                                  /* plan regions start zero-filled (:523-531) */
    int ok = fftb200_plan_many(&plan.b200_p, dim, n, NULL, 0, 0, NULL, 0, 0, type, 1);   /* :233-242 */
    CHECK(ok == FFTB200_SUCCESS, "fftb200_plan_many -> %d (%s)", ok, fftb200_strerror(ok));
    iface_plan copy = plan;                                          /* Legion may copy the instance */
    void *din = to_device(in, in_bytes), *dout = to_device(out, out_bytes);
    switch (type) {                                                  /* :569-581 */
        case FFTB200_R2C: ok = fftb200_exec_r2c(copy.b200_p, din, dout); break;
        case FFTB200_C2C: ok = fftb200_exec_c2c(copy.b200_p, din, dout, FFTB200_FORWARD); break;
        case FFTB200_D2Z: ok = fftb200_exec_d2z(copy.b200_p, din, dout); break;
        default: ok = fftb200_exec_z2z(copy.b200_p, din, dout, FFTB200_FORWARD); break;
    }
    CHECK(ok == FFTB200_SUCCESS, "fftb200_exec -> %d (%s)", ok, fftb200_strerror(ok));
    cudaDeviceSynchronize();                                         /* Legion: end of the task */
    cudaMemcpy(out, dout, out_bytes, cudaMemcpyDeviceToHost);
    cudaFree(din);
    cudaFree(dout);
    CHECK(fftb200_destroy(copy.b200_p) == FFTB200_SUCCESS, "destroy");
    CHECK(fftb200_destroy(0) == FFTB200_SUCCESS, "destroy(0) must be a no-op");
    CHECK(fftb200_exec_z2z(copy.b200_p, din, dout, FFTB200_FORWARD) == FFTB200_INVALID_PLAN, "stale handle must be rejected");
}

static int near(double a, double b, double tol) { return fabs(a - b) <= tol; }

int main(void) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { printf("no CUDA device: nothing run\n"); return 77; }
    printf("libfft_b200 version %d\n", fftb200_version());
    int i;
    { /* test1d: int1d 5, 3+3i -> [15+15i, 0, 0, 0, 0] */
        int n[1] = {5};
        c64 in[5], out[5];
        for (i = 0; i < 5; ++i) { in[i].re = in[i].im = 3; out[i].re = out[i].im = 0; }
        run_basic(1, n, FFTB200_Z2Z, in, sizeof in, out, sizeof out);
        CHECK(near(out[0].re, 15, 1e-13) && near(out[0].im, 15, 1e-13), "test1d DC = %g%+gi", out[0].re, out[0].im);
        for (i = 1; i < 5; ++i) CHECK(near(out[i].re, 0, 1e-13) && near(out[i].im, 0, 1e-13), "test1d bin %d", i);
    }
    { /* test1d_real: int1d 3, 3.0 -> [9, 0], third entry untouched */
        int n[1] = {3};
        double in[3] = {3, 3, 3};
        c64 out[3];
        for (i = 0; i < 3; ++i) out[i].re = out[i].im = -7;
        run_basic(1, n, FFTB200_D2Z, in, sizeof in, out, sizeof out);
        CHECK(near(out[0].re, 9, 1e-13) && near(out[1].re, 0, 1e-13) && out[2].re == -7, "test1d_real");
    }
    { /* test1d_float: complex32 3, 3+3i -> [9+9i, 0, 0]; test1d_float_real: [9, 0] */
        int n[1] = {3};
        c32 in[3], out[3];
        float rin[3] = {3, 3, 3};
        for (i = 0; i < 3; ++i) { in[i].re = in[i].im = 3; out[i].re = out[i].im = 0; }
        run_basic(1, n, FFTB200_C2C, in, sizeof in, out, sizeof out);
        CHECK(near(out[0].re, 9, 1e-5) && near(out[0].im, 9, 1e-5) && near(out[1].re, 0, 1e-5), "test1d_float");
        for (i = 0; i < 3; ++i) out[i].re = out[i].im = -7;
        run_basic(1, n, FFTB200_R2C, rin, sizeof rin, out, sizeof out);
        CHECK(near(out[0].re, 9, 1e-5) && near(out[1].re, 0, 1e-5) && out[2].re == -7, "test1d_float_real");
    }
    { /* test2d: 2x2, 5+5i, output pre-filled with 1 -> [20+20i, 0, 0, 0] */
        int n[2] = {2, 2};
        c64 in[4], out[4];
        for (i = 0; i < 4; ++i) { in[i].re = in[i].im = 5; out[i].re = out[i].im = 1; }
        run_basic(2, n, FFTB200_Z2Z, in, sizeof in, out, sizeof out);
        CHECK(near(out[0].re, 20, 1e-13) && near(out[0].im, 20, 1e-13), "test2d DC");
        for (i = 1; i < 4; ++i) CHECK(near(out[i].re, 0, 1e-13) && near(out[i].im, 0, 1e-13), "test2d bin %d", i);
    }
    { /* test3d: {3,2,2}, 3+3i -> [36+36i, 0 x 11] */
        int n[3] = {3, 2, 2};
        c64 in[12], out[12];
        for (i = 0; i < 12; ++i) { in[i].re = in[i].im = 3; out[i].re = out[i].im = 0; }
        run_basic(3, n, FFTB200_Z2Z, in, sizeof in, out, sizeof out);
        CHECK(near(out[0].re, 36, 1e-13) && near(out[0].im, 36, 1e-13), "test3d DC");
        for (i = 1; i < 12; ++i) CHECK(near(out[i].re, 0, 1e-13) && near(out[i].im, 0, 1e-13), "test3d bin %d", i);
    }
    { /* test3d_batch / test3d_batch_real: {3,3,2}: n_batch = {3,3}, i_dist = offsets[2]/offsets[0] = 9, batch = n[2] = 2 */
        int n_batch[2] = {3, 3};
        const int i_dist = 9, batch = 2;
        c64 in[18], out[18];
        double rin[18];
        fftb200_handle h = 0;
        void *din, *dout;
        int ok;
        for (i = 0; i < 18; ++i) { in[i].re = in[i].im = 3; rin[i] = 3; out[i].re = out[i].im = 0; }
        ok = fftb200_plan_many(&h, 2, n_batch, n_batch, 1, i_dist, n_batch, 1, i_dist, FFTB200_Z2Z, batch);   /* :389-398 */
        CHECK(ok == 0, "plan_many batch -> %d", ok);
        din = to_device(in, sizeof in); dout = to_device(out, sizeof out);
        CHECK(fftb200_exec_z2z(h, din, dout, FFTB200_FORWARD) == 0, "exec batch");
        cudaMemcpy(out, dout, sizeof out, cudaMemcpyDeviceToHost);
        for (i = 0; i < 18; ++i) {
            double want = (i == 0 || i == 9) ? 27 : 0;
            CHECK(near(out[i].re, want, 1e-13) && near(out[i].im, want, 1e-13), "test3d_batch [%d] = %g%+gi", i, out[i].re, out[i].im);
        }
        fftb200_destroy(h); cudaFree(din); cudaFree(dout);
        for (i = 0; i < 18; ++i) out[i].re = out[i].im = -7;
        ok = fftb200_plan_many(&h, 2, n_batch, n_batch, 1, i_dist, n_batch, 1, i_dist, FFTB200_D2Z, batch);
        CHECK(ok == 0, "plan_many batch real -> %d", ok);
        din = to_device(rin, sizeof rin); dout = to_device(out, sizeof out);
        CHECK(fftb200_exec_d2z(h, din, dout) == 0, "exec batch real");
        cudaMemcpy(out, dout, sizeof out, cudaMemcpyDeviceToHost);
        for (i = 0; i < 18; ++i) {
            if (i % 3 == 2) { CHECK(out[i].re == -7 && out[i].im == -7, "test3d_batch_real: entry %d must stay untouched", i); }
            else { double want = (i == 0 || i == 9) ? 27 : 0; CHECK(near(out[i].re, want, 1e-13) && near(out[i].im, 0, 1e-13), "test3d_batch_real [%d]", i); }
        }
        fftb200_destroy(h); cudaFree(din); cudaFree(dout);
    }
    { /* error convention: 0 / 1 / 4 keep cuFFT's meaning (src/fft.rg:246-250, 584-591) */
        fftb200_handle h = 0;
        int n[1] = {8};
        CHECK(fftb200_plan_many(&h, 4, n, NULL, 0, 0, NULL, 0, 0, FFTB200_Z2Z, 1) == FFTB200_INVALID_VALUE && h == 0, "rank 4 must be rejected");
        CHECK(fftb200_exec_z2z(12345, n, n, FFTB200_FORWARD) == FFTB200_INVALID_PLAN, "bogus handle");
    }
    if (failures) printf("regent_call_sequence: %d FAILURES\n", failures);
    else printf("regent_call_sequence: all known answers reproduced\n");
    return failures ? 1 : 0;
}
