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
	double T;		/* sqrt(log(20) * n), discreteFourierTransform.c:142 */
};
