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
	/* tables of the n = 2^20 kernels (k_dft_fast.cu, N = 256 x 2048), NULL otherwise */
	const double2 *q_twA;	/* [15][16]:  exp(-2 pi i k r / 256),  r = 1..15, k < 16  (radix-16 pass, Ns = 16) */
	const double2 *q_twB;	/* [7][256]:  exp(-2 pi i j r / 2048), r = 1..7,  j < 256 (radix-8 pass, Ns = 256) */
	const double2 *q_e1;	/* [16][2048]: exp(-2 pi i j2 t / N) */
	const double2 *q_e2;	/* [2048]: exp(-2 pi i 16 j2 / N) */
	const double2 *q_p1;	/* [256]:  exp(-2 pi i k1 / n) */
	const double2 *q_p2;	/* [2048]: exp(-2 pi i 256 k2 / n) */
	int dbg;		/* experiment switches (STSGPU_DFT_DBG), 0 in production */
	double T;		/* sqrt(log(20) * n), discreteFourierTransform.c:142 */
};
