/* Hand-written build configuration for compiling the reference's vendored
 * FFTW 3.3.8 sources IN PLACE (from /root/reference/fftw-3.3.8) with the
 * Makefile in this directory.  It replaces autoconf's generated config.h:
 * x86-64 Linux, gcc, glibc.  Precision (FFTW_SINGLE) and the SIMD sets
 * (HAVE_SSE2 / HAVE_AVX / HAVE_AVX2) are passed on the compiler command line by
 * oracle/Makefile, so the same header serves the scalar, AVX2 and float
 * builds.  Test infrastructure only: nothing here is product code.
 */
#ifndef ORACLE_FFTW_CONFIG_H
#define ORACLE_FFTW_CONFIG_H

#define PACKAGE "fftw"
#define VERSION "3.3.8"
#define PACKAGE_VERSION "3.3.8"
#define FFTW_CC "gcc (oracle/Makefile)"

#define STDC_HEADERS 1
#define TIME_WITH_SYS_TIME 1
#define HAVE_ALLOCA 1
#define HAVE_ALLOCA_H 1
#define FFTW_ENABLE_ALLOCA 1
#define HAVE_ABORT 1
#define HAVE_CLOCK_GETTIME 1
#define HAVE_GETTIMEOFDAY 1
#define HAVE_DLFCN_H 1
#define HAVE_FCNTL_H 1
#define HAVE_FENV_H 1
#define HAVE_INTTYPES_H 1
#define HAVE_LIMITS_H 1
#define HAVE_MALLOC_H 1
#define HAVE_MEMORY_H 1
#define HAVE_STDDEF_H 1
#define HAVE_STDINT_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRINGS_H 1
#define HAVE_STRING_H 1
#define HAVE_SYS_STAT_H 1
#define HAVE_SYS_TIME_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_UNISTD_H 1
#define HAVE_ISNAN 1
#define HAVE_LIBM 1
#define HAVE_LONG_DOUBLE 1
#define HAVE_COSL 1
#define HAVE_SINL 1
#define HAVE_DECL_COSL 1
#define HAVE_DECL_SINL 1
#define HAVE_DECL_COSQ 0
#define HAVE_DECL_SINQ 0
#define HAVE_DRAND48 1
#define HAVE_DECL_DRAND48 1
#define HAVE_DECL_SRAND48 1
#define HAVE_MEMALIGN 1
#define HAVE_DECL_MEMALIGN 1
#define HAVE_POSIX_MEMALIGN 1
#define HAVE_DECL_POSIX_MEMALIGN 1
#define HAVE_MEMMOVE 1
#define HAVE_MEMSET 1
#define HAVE_PTRDIFF_T 1
#define HAVE_UINTPTR_T 1
#define HAVE_SNPRINTF 1
#define HAVE_SQRT 1
#define HAVE_STRCHR 1
#define HAVE_VPRINTF 1
#define HAVE_GETPAGESIZE 1

#define SIZEOF_DOUBLE 8
#define SIZEOF_FLOAT 4
#define SIZEOF_INT 4
#define SIZEOF_LONG 8
#define SIZEOF_LONG_LONG 8
#define SIZEOF_PTRDIFF_T 8
#define SIZEOF_SIZE_T 8
#define SIZEOF_UNSIGNED_INT 4
#define SIZEOF_UNSIGNED_LONG 8
#define SIZEOF_UNSIGNED_LONG_LONG 8
#define SIZEOF_FFTW_R2R_KIND 4

/* threads: POSIX threads (the image has no libgomp.spec, so no OpenMP) */
#ifdef ORACLE_FFTW_THREADS
#define HAVE_THREADS 1
#define USING_POSIX_THREADS 1
#endif

#endif
