/* see fftw3.h in this directory: allocation and plan bookkeeping only; there is no CPU FFT in the GPU build */
#include <stdlib.h>
#include "fftw3.h"

struct stsgpu_no_fftw_plan {
	int n;
};

void *fftw_malloc(size_t n) { return malloc(n); }
void fftw_free(void *p) { free(p); }

fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags)
{
	(void) in; (void) out; (void) flags;
	fftw_plan p = malloc(sizeof(*p));
	if (p != NULL)
		p->n = n;
	return p;
}

void fftw_execute(const fftw_plan p)
{
	(void) p;
	fputs("FATAL: fftw_execute() called in the GPU build of sts: the spectral test runs on the device, there is no CPU fallback\n", stderr);
	abort();
}

void fftw_destroy_plan(fftw_plan p) { free(p); }
