/* dft.h - plan of the four-step FP64 FFT used by k_dft.cu (tables live in device memory) */
#pragma once
#include <cuda_runtime.h>

struct DftPlan {
	int logN1, logN2;	/* N = n/2 = 2^logN1 * 2^logN2 */
	int cw;			/* columns per CTA in the column pass */
	const double2 *wm;	/* exp(-2 pi i k / 1024), k < 1024: in-pass twiddles */
	const double2 *tl;	/* exp(-2 pi i e / N), e < 1024 */
	const double2 *th;	/* exp(-2 pi i 1024 e / N), e < max(1, N / 1024) */
	const double2 *p1;	/* exp(-2 pi i k1 / n), k1 < N1 */
	const double2 *p2;	/* exp(-2 pi i N1 k2 / n), k2 < N2 */
	const double2 *tw8;	/* [7][8]:  exp(-2 pi i k r / 64),  r = 1..7, k < 8  (radix-8 pass with Ns = 8) */
	const double2 *tw64;	/* [7][64]: exp(-2 pi i k r / 512), r = 1..7, k < 64 (radix-8 pass with Ns = 64) */
	const double2 *tw2d;	/* [512][1024]: exp(-2 pi i k1 j2 / N) for the n = 2^20 kernels, else NULL */
	double T;		/* sqrt(log(20) * n), discreteFourierTransform.c:142 */
};
