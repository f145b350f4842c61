/*
 * fft_f64.h - plain FP64 mixed-radix complex FFT (TEST INFRASTRUCTURE, not product code).
 *
 * Used by (1) the fftw3 shim that lets the unmodified reference link here
 * (oracle/shim/fftw3.h) and (2) the CPU oracle's DFT test.  The reference's DFT
 * arithmetic lives in libfftw3 (not vendored, no version pin: src/Makefile:52,
 * tests/discreteFourierTransform.c:189-194,361), so parity for the DFT test is
 * "unpinned": any correct FP64 DFT agrees on N_1 except for moduli within
 * rounding distance of the threshold T.
 */
#ifndef ORACLE_FFT_F64_H
#define ORACLE_FFT_F64_H
#include <stddef.h>

typedef struct { double re, im; } fft_cplx;

typedef struct fft_plan_s fft_plan;

/* plan for complex FFT of any length n >= 1 (forward, e^{-2 pi i jk/n}) */
fft_plan *fft_plan_create(long n);
void fft_plan_destroy(fft_plan *p);
/* out-of-place forward transform; in and out must not alias */
void fft_forward(const fft_plan *p, const fft_cplx *in, fft_cplx *out);
/* real input of length n -> first n/2+1 complex outputs (r2c convention of FFTW) */
void fft_r2c(const fft_plan *p, const double *in, fft_cplx *out_half);

#endif
