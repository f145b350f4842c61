/*
 * fft_f64.c - recursive mixed-radix decimation-in-time FFT in FP64.
 * TEST INFRASTRUCTURE (see fft_f64.h).  Written for clarity, not speed.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "fft_f64.h"

struct fft_plan_s {
	long n;
	fft_cplx *w;		/* w[j] = exp(-2 pi i j / n), j in [0,n) */
	fft_cplx *work;		/* n scratch for fft_r2c */
	fft_cplx *work2;
};

static long
smallest_factor(long n)
{
	if (n % 4 == 0) return 4;
	if (n % 2 == 0) return 2;
	for (long p = 3; p * p <= n; p += 2)
		if (n % p == 0) return p;
	return n;
}

fft_plan *
fft_plan_create(long n)
{
	fft_plan *p = calloc(1, sizeof(*p));
	if (p == NULL || n < 1) return NULL;
	p->n = n;
	p->w = malloc((size_t) n * sizeof(fft_cplx));
	p->work = malloc((size_t) n * sizeof(fft_cplx));
	p->work2 = malloc((size_t) n * sizeof(fft_cplx));
	if (!p->w || !p->work || !p->work2) return NULL;
	for (long j = 0; j < n; j++) {
		/* reduce the angle to the first octant for accuracy */
		long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double) j / (long double) n;
		p->w[j].re = (double) cosl(a);
		p->w[j].im = (double) sinl(a);
	}
	return p;
}

void
fft_plan_destroy(fft_plan *p)
{
	if (!p) return;
	free(p->w); free(p->work); free(p->work2); free(p);
}

/* out[0..n) = DFT of in[0], in[stride], ... ; tw_step = N/n for the top-level table */
static void
rec(const fft_plan *pl, long n, const fft_cplx *in, long stride, fft_cplx *out, long tw_step)
{
	if (n == 1) { out[0] = in[0]; return; }
	long p = smallest_factor(n);
	long m = n / p;
	for (long r = 0; r < p; r++)
		rec(pl, m, in + r * stride, stride * p, out + r * m, tw_step * p);
	const fft_cplx *w = pl->w;
	const long N = pl->n;
	if (p == 2) {
		for (long k = 0; k < m; k++) {
			fft_cplx a = out[k], b = out[m + k], t;
			const fft_cplx wk = w[k * tw_step];
			t.re = b.re * wk.re - b.im * wk.im;
			t.im = b.re * wk.im + b.im * wk.re;
			out[k].re = a.re + t.re; out[k].im = a.im + t.im;
			out[m + k].re = a.re - t.re; out[m + k].im = a.im - t.im;
		}
		return;
	}
	if (p == 4) {
		for (long k = 0; k < m; k++) {
			fft_cplx a = out[k], b, c, d, x;
			const fft_cplx w1 = w[k * tw_step], w2 = w[2 * k * tw_step], w3 = w[3 * k * tw_step];
			x = out[m + k];     b.re = x.re * w1.re - x.im * w1.im; b.im = x.re * w1.im + x.im * w1.re;
			x = out[2 * m + k]; c.re = x.re * w2.re - x.im * w2.im; c.im = x.re * w2.im + x.im * w2.re;
			x = out[3 * m + k]; d.re = x.re * w3.re - x.im * w3.im; d.im = x.re * w3.im + x.im * w3.re;
			fft_cplx s0 = { a.re + c.re, a.im + c.im }, s1 = { a.re - c.re, a.im - c.im };
			fft_cplx s2 = { b.re + d.re, b.im + d.im }, s3 = { b.re - d.re, b.im - d.im };
			out[k].re = s0.re + s2.re;         out[k].im = s0.im + s2.im;
			out[m + k].re = s1.re + s3.im;     out[m + k].im = s1.im - s3.re;	/* -i * s3 */
			out[2 * m + k].re = s0.re - s2.re; out[2 * m + k].im = s0.im - s2.im;
			out[3 * m + k].re = s1.re - s3.im; out[3 * m + k].im = s1.im + s3.re;	/* +i * s3 */
		}
		return;
	}
	/* generic radix p butterfly, O(p^2) */
	fft_cplx *t = malloc((size_t) p * sizeof(fft_cplx));
	const long pstep = N / p;	/* w[q*r*pstep mod N] = exp(-2 pi i q r / p) */
	for (long k = 0; k < m; k++) {
		for (long r = 0; r < p; r++) {
			fft_cplx x = out[r * m + k];
			const fft_cplx wk = w[(r * k * tw_step) % N];
			t[r].re = x.re * wk.re - x.im * wk.im;
			t[r].im = x.re * wk.im + x.im * wk.re;
		}
		for (long q = 0; q < p; q++) {
			double sr = 0.0, si = 0.0;
			for (long r = 0; r < p; r++) {
				const fft_cplx wq = w[((q * r) % p) * pstep];
				sr += t[r].re * wq.re - t[r].im * wq.im;
				si += t[r].re * wq.im + t[r].im * wq.re;
			}
			out[q * m + k].re = sr; out[q * m + k].im = si;
		}
	}
	free(t);
}

void
fft_forward(const fft_plan *p, const fft_cplx *in, fft_cplx *out)
{
	rec(p, p->n, in, 1, out, 1);
}

void
fft_r2c(const fft_plan *p, const double *in, fft_cplx *out_half)
{
	for (long i = 0; i < p->n; i++) { p->work[i].re = in[i]; p->work[i].im = 0.0; }
	fft_forward(p, p->work, p->work2);
	memcpy(out_half, p->work2, (size_t) (p->n / 2 + 1) * sizeof(fft_cplx));
}
