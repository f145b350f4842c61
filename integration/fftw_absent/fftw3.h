/*
 * fftw3.h - stands in for libfftw3's header where that library is not installed (this image), so that the
 * reference's own translation units (utils/defs.h:31 includes <fftw3.h>; DiscreteFourierTransform_init plans an
 * FFT per thread, discreteFourierTransform.c:189-194) compile and link into integration/_ref/sts_gpu.  A maintainer
 * with libfftw3 links the real one instead (Makefile: FFTW=system).  In the GPU build the spectral test runs on the
 * device (sts_b200/csrc/k_dft_*.cu): the plan is only ever created and destroyed, and fftw_execute() - reachable
 * solely from the CPU iterate body that the GPU build does not link - aborts instead of computing anything.
 */
#ifndef STSGPU_FFTW_ABSENT_H
#define STSGPU_FFTW_ABSENT_H
#include <stdio.h>		/* utils/defs.h relies on fftw3.h for FILE */
#include <stddef.h>

#if defined(_Complex_I) && defined(complex) && defined(I)
typedef double _Complex fftw_complex;	/* libfftw3's own rule: C99 complex when <complex.h> was included first */
#else
typedef double fftw_complex[2];
#endif

typedef struct stsgpu_no_fftw_plan *fftw_plan;
#define FFTW_ESTIMATE (1U << 6)

void *fftw_malloc(size_t n);
void fftw_free(void *p);
fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);

#endif
