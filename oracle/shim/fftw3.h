/*
 * fftw3.h - minimal stand-in for libfftw3's header so that the UNMODIFIED reference
 * (tests/discreteFourierTransform.c:45,189,194,361,1189,1193; utils/defs.h:31) compiles
 * and links in an image without libfftw3.  TEST INFRASTRUCTURE ONLY: it exists to build
 * oracle/_ref/sts, the binary the oracle is pinned against.  Backed by oracle/fft_f64.c.
 */
#ifndef ORACLE_SHIM_FFTW3_H
#define ORACLE_SHIM_FFTW3_H
#include <stdio.h>	/* utils/defs.h relies on fftw3.h to provide FILE */
#include <stddef.h>

#if defined(_Complex_I) && defined(complex) && defined(I)
typedef double _Complex fftw_complex;	/* same rule as the real fftw3.h: C99 complex if <complex.h> came first */
#else
typedef double fftw_complex[2];
#endif

typedef struct shim_fftw_plan_s *fftw_plan;
#define FFTW_ESTIMATE (1U << 6)

void *fftw_malloc(size_t n);
void fftw_free(void *p);
fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);

#endif
