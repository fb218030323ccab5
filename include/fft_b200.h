/*
 * fft_b200.h — C ABI of libfft_b200, the B200 (sm_100a) FFT engine that sits behind
 * Regent-FFT's GPU branch.
 *
 * Pure C (no C++, no default arguments) so that Terra's `terralib.includec` can parse
 * it exactly as it parses cufftXt.h today (reference src/fft.rg:15-20).  Constants are
 * `enum`s, not `#define`s: includec drops macros (cf. src/fft.rg:22-25, where the
 * FFTW macros are re-declared by hand).
 *
 * Every entry point replaces one cuFFT call site of the reference:
 *
 *   fftb200_plan_many   <- cufftPlanMany   src/fft.rg:233,236,239,242 (make_plan_gpu)
 *                                          src/fft.rg:389,392,395,398 (make_plan_gpu_batch)
 *   fftb200_exec_c2c    <- cufftExecC2C    src/fft.rg:574
 *   fftb200_exec_z2z    <- cufftExecZ2Z    src/fft.rg:580
 *   fftb200_exec_d2z    <- cufftExecD2Z    src/fft.rg:577
 *   fftb200_exec_r2c    <- cufftExecR2C    src/fft.rg:571 (commented out upstream; README.md:16 promises it)
 *   fftb200_destroy     <- cufftDestroy    src/fft.rg:638
 *   fftb200_set_stream  <- (cufftSetStream; the reference relies on the default stream, src/fft.rg:563-581)
 *
 * Semantics held:
 *   - forward transforms are unnormalised, sign -1 (src/fft.rg:22,574,580); backward (+1) is also
 *     provided for C2C/Z2Z;
 *   - out-of-place, input preserved (privilege reads(input), src/fft.rg:545); in == out is accepted
 *     for C2C/Z2Z;
 *   - inembed == NULL  => cuFFT "basic" layout: packed row-major, batches back to back, R2C/D2Z
 *     output rows of n[rank-1]/2+1 complex (stride/dist arguments ignored, as cuFFT does);
 *   - inembed != NULL  => advanced layout, element (b, j0..jr-1) at
 *     b*dist + (((j0*embed[1] + j1)*embed[2] + ...) + j_last)*stride, input embed counted in
 *     input elements, output embed in output elements (fftw-3.3.8/api/plan-many-dft.c:43-46,
 *     api/plan-many-dft-r2c.c:44-49);
 *   - n[i] is taken as given (the reference passes Regent extents as row-major dims,
 *     src/fft.rg:220-223; the library does not reorder);
 *   - asynchronous with respect to the host, ordered on the plan's stream (default: stream 0);
 *   - in/out may be device (framebuffer) or host pointers.  Host memory — the zero-copy instances the
 *     reference's mapper creates (test/test_mapper.cc:45-58), pinned or pageable — is staged through
 *     plan-owned HBM buffers with one copy in and one copy out on the plan's stream, so every FFT
 *     pass still runs at HBM speed instead of re-streaming the data over PCIe three times;
 *   - return value 0 == success; 1 and 4 keep cuFFT's meaning so fft.rg's checks
 *     (src/fft.rg:246-250, 584-591) keep working; nothing throws or aborts across the ABI;
 *   - handles are plain 64-bit integers (0 = null) that survive being memcpy'd inside the plan
 *     region (src/fft.rg:48-65); destroy(0) is a no-op and destroy needs no current device
 *     (src/fft.rg:523-531, 624-645).
 *   - thread-safe across plans (Legion runs one task per GPU processor concurrently).  One plan is one
 *     execution context, as with cuFFT: it owns twiddle tables and, for some shapes, a work buffer (four/six-step
 *     1-D, blocked 3-D intermediate, host staging), so execs of the SAME plan must be ordered on one stream.
 *
 * There is no CPU fallback: every exec runs hand-written sm_100a kernels or fails.
 */
#ifndef FFT_B200_H
#define FFT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__) || defined(__clang__)
#define FFTB200_API __attribute__((visibility("default")))
#else
#define FFTB200_API
#endif

typedef unsigned long long fftb200_handle; /* POD; 0 == null; storable by value in iface.plan */

/* same numeric codes as cufftType (cufft.h) so the type selection at src/fft.rg:231-243 maps 1:1 */
typedef enum fftb200_type_e {
    FFTB200_R2C = 0x2a, /* float   -> complex32 */
    FFTB200_C2C = 0x29, /* complex32 -> complex32 */
    FFTB200_D2Z = 0x6a, /* double  -> complex64 */
    FFTB200_Z2Z = 0x69, /* complex64 -> complex64 */
    /* inverse real transforms (no reference call site: src/fft.rg is forward-only) */
    FFTB200_C2R = 0x2c, /* complex32 -> float  */
    FFTB200_Z2D = 0x6c  /* complex64 -> double */
} fftb200_type;

typedef enum fftb200_result_e {
    FFTB200_SUCCESS = 0,
    FFTB200_INVALID_PLAN = 1,
    FFTB200_ALLOC_FAILED = 2,
    FFTB200_INVALID_TYPE = 3,
    FFTB200_INVALID_VALUE = 4,
    FFTB200_INTERNAL_ERROR = 5,
    FFTB200_EXEC_FAILED = 6,
    FFTB200_SETUP_FAILED = 7,
    FFTB200_INVALID_SIZE = 8,
    FFTB200_UNSUPPORTED = 16
} fftb200_result;

enum { FFTB200_FORWARD = -1, FFTB200_INVERSE = 1 };

/* Plan creation.  Call with the target GPU current (a Legion TOC processor's host thread). */
FFTB200_API int fftb200_plan_many(fftb200_handle *plan, int rank, const int *n,
                      const int *inembed, int istride, int idist,
                      const int *onembed, int ostride, int odist,
                      fftb200_type type, int batch);

/* Bind later execs of this plan to a CUDA stream (cudaStream_t passed as void*). */
FFTB200_API int fftb200_set_stream(fftb200_handle plan, void *cuda_stream);

FFTB200_API int fftb200_exec_c2c(fftb200_handle plan, const void *in, void *out, int direction);
FFTB200_API int fftb200_exec_z2z(fftb200_handle plan, const void *in, void *out, int direction);
FFTB200_API int fftb200_exec_r2c(fftb200_handle plan, const void *in, void *out);
FFTB200_API int fftb200_exec_d2z(fftb200_handle plan, const void *in, void *out);
/* Unnormalised inverse of R2C / D2Z (cufftExecC2R / cufftExecZ2D, fftw_plan_dft_c2r semantics): in = packed half
 * spectrum [..][n_last/2+1] complex, out = [..][n_last] reals = n_total * the original data.  The input is
 * preserved (multi-dimensional plans go through a plan-owned work buffer).  Power-of-two sizes run the tiled fast path,
 * sizes 2^a 3^b 5^c 7^d 11^e 13^f the mixed-radix kernels (one launch per axis), any other size the generic path (Hermitian
 * completion, backward complex stages, real part).
 * In-place transforms (in == out): C2C / Z2Z whenever the two layouts coincide; R2C / D2Z (and C2R / Z2D off the
 * power-of-two path) with FFTW's padded in-place layout, i.e. real rows of 2*(n_last/2+1) reals (inembed[last] =
 * 2*(n_last/2+1), onembed[last] = n_last/2+1 for R2C; fftw-3.3.8/doc/reference.texi "Real-data DFT Array Format"). */
FFTB200_API int fftb200_exec_c2r(fftb200_handle plan, const void *in, void *out);
FFTB200_API int fftb200_exec_z2d(fftb200_handle plan, const void *in, void *out);

/* Normalisation helper.  Every transform here is unnormalised like FFTW's and cuFFT's (a forward transform followed
 * by the backward one multiplies the data by n_total, fftw-3.3.8/doc/reference.texi:1982-2004); the reference has no
 * backward path at all (src/fft.rg:319,574-580), so this has no call site there.  Scales, on the plan's stream and in
 * place, exactly the elements the plan's transform WRITES (its output layout: advanced-layout padding is untouched) by
 * `factor`; factor == 0 means 1 / n_total (n_total = product of n[], per batch member). */
FFTB200_API int fftb200_scale(fftb200_handle plan, void *data, double factor);

/* Free the plan's device tables and work buffers.  Any thread; no current device needed. */
FFTB200_API int fftb200_destroy(fftb200_handle plan);

/* ---- introspection (benchmarks, tests; no reference counterpart) ------------------------- */

/* Bytes of device work area the plan owns (cf. cufftGetSize). */
FFTB200_API int fftb200_get_work_size(fftb200_handle plan, unsigned long long *bytes);
/* Number of kernel launches one exec issues, and a human-readable pass list
 * ("name L=512 tiles=32768 ...", one line per launch) copied into buf (NUL-terminated). */
FFTB200_API int fftb200_get_launch_count(fftb200_handle plan, int *launches);
FFTB200_API int fftb200_describe(fftb200_handle plan, char *buf, int buflen);
/* Algorithmic HBM bytes of launch `i` (elements read * sizeof(in) + elements written * sizeof(out)). */
FFTB200_API int fftb200_get_launch_bytes(fftb200_handle plan, int i, unsigned long long *bytes);
/* Record per-launch CUDA events on the next execs (on=1 starts a new series of up to 256 execs) and
 * read back launch i's mean duration over the execs recorded since. */
FFTB200_API int fftb200_set_profiling(fftb200_handle plan, int on);
FFTB200_API int fftb200_get_launch_ms(fftb200_handle plan, int i, float *ms);

/* ---- multi-GPU slab transforms (3-D, one plan per GPU / rank) --------------------------------------
 * No call site in the reference: its distrib path (src/fft.rg:513-537) transforms independent shards
 * and README.md:117-119 lists a distributed transform as future work.  Semantics follow the vendored
 * FFTW-MPI (fftw-3.3.8/mpi/dft-rank-geq2.c:40-59, doc/mpi.texi:259-270, 443-466):
 *   rank r of G holds   in  [n0/G][n1][n2]          slab r of dimension 0, row-major
 *   and receives        out [n1/G][n0][n2c]         slab r of dimension 1 (FFTW_MPI_TRANSPOSED_OUT),
 *                                                   n2c = n2 (C2C/Z2Z) or n2/2+1 (R2C/D2Z)
 * n0, n1, n2 and G are powers of two, G <= 16.  Every rank must make the same sequence of exec calls
 * (they are collective).  `chunks` = pipeline depth of the fused exchange (1 = no overlap).
 *
 * Two ways to run the exchange:
 *  (1) fused, peer-to-peer: the y-axis FFT pass stores straight into the destination ranks' exchange
 *      areas over NVLink.  Connect the plans once: across processes exchange the 64-byte IPC handles
 *      (get_ipc_handle -> any host all-gather -> connect_ipc); inside one process (Legion: one
 *      process, one GPU processor per device) pass the areas' pointers (get_area -> connect_ptrs).
 *      Then fftb200_slab_exec.
 *  (2) staged: exec_pre leaves G packed blocks [d][n0/G][n1/G][n2c] in `send`; the caller runs any
 *      all-to-all (e.g. NCCL) into `recv`; exec_post finishes from `recv`. */
FFTB200_API int fftb200_slab_plan(fftb200_handle *plan, const int *n /* [3] */, fftb200_type type, int rank, int nranks,
                                  int chunks);
/* 2-D slabs (C2C / Z2Z): rank r holds in [n0/G][n1] and receives out [n1/G][n0] (transposed-out); the row FFT's
 * store is the global transpose.  Fused exchange only (connect as above, then fftb200_slab_exec). */
FFTB200_API int fftb200_slab_plan_2d(fftb200_handle *plan, const int *n /* [2] */, fftb200_type type, int rank, int nranks);
FFTB200_API int fftb200_slab_get_ipc_handle(fftb200_handle plan, void *handle64);
FFTB200_API int fftb200_slab_connect_ipc(fftb200_handle plan, const void *handles /* nranks x 64 bytes, rank order */);
FFTB200_API int fftb200_slab_get_area(fftb200_handle plan, void **area, unsigned long long *bytes);
FFTB200_API int fftb200_slab_connect_ptrs(fftb200_handle plan, void *const *areas /* [nranks] */);
FFTB200_API int fftb200_slab_exec(fftb200_handle plan, const void *in, void *out, int direction);
FFTB200_API int fftb200_slab_exec_pre(fftb200_handle plan, const void *in, void *send, int direction);
FFTB200_API int fftb200_slab_exec_post(fftb200_handle plan, const void *recv, void *out, int direction);
/* per-phase CUDA-event timing of the last fftb200_slab_exec: ms[0] x-axis pass, ms[1] handshake + y-axis
 * pass with the exchange, ms[2] remaining z-axis work after the last exchange chunk was issued */
FFTB200_API int fftb200_slab_set_timing(fftb200_handle plan, int on);
FFTB200_API int fftb200_slab_get_phase_ms(fftb200_handle plan, float *ms /* [3] */);

FFTB200_API const char *fftb200_strerror(int code);
/* library version: major*10000 + minor*100 + patch */
FFTB200_API int fftb200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FFT_B200_H */
