/* fftw_shim.c - the 5 libfftw3 entry points the reference calls, on top of oracle/fft_f64.c. */
#include <stdlib.h>
#include "fftw3.h"
#include "../fft_f64.h"

struct shim_fftw_plan_s {
	fft_plan *plan;
	double *in;
	fft_cplx *out;
};

void *fftw_malloc(size_t n) { return malloc(n); }
void fftw_free(void *p) { free(p); }

fftw_plan
fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags)
{
	(void) flags;
	fftw_plan p = malloc(sizeof(*p));
	if (p == NULL) return NULL;
	p->plan = fft_plan_create(n);
	p->in = in;
	p->out = (fft_cplx *) out;	/* {re,im} pairs: layout-compatible with double[2] and double _Complex */
	if (p->plan == NULL) { free(p); return NULL; }
	return p;
}

void fftw_execute(const fftw_plan p) { fft_r2c(p->plan, p->in, p->out); }

void
fftw_destroy_plan(fftw_plan p)
{
	if (p == NULL) return;
	fft_plan_destroy(p->plan);
	free(p);
}
