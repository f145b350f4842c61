/* dft.h - plan of the four-step FP64 FFT of the DFT test (tables live in device memory) */
#pragma once
#include <cuda_runtime.h>

#define DFT_MAX_PASSES 16

struct DftPlan {
	/* N = n/2 complex points = N1 x N2; each 1-D FFT is a Stockham pass list with radices from {8, 5, 4, 3, 2} */
	int N1, N2;
	int np1, np2;
	int rad1[DFT_MAX_PASSES], rad2[DFT_MAX_PASSES];
	int cw;			/* columns per CTA in the column pass */
	int threads;		/* CTA size of the generic kernels (256 or 512) */
	int fast;		/* 1: n = 2^20, the kernels of k_dft_fast.cu; 2: n = 10^6, k_dft_1e6.cu; 3: any n, k_dft_any.cu; 4: n = 10^7, k_dft_1e7.cu; 0: generic */
	/* per-pass twiddles, pass p at offset off[p]: entry (r - 1) Ns + k = exp(-2 pi i k r / (Ns R)), so that
	 * consecutive butterflies (consecutive k) read consecutive table entries */
	const double2 *w1;
	const double2 *w2;
	int off1[DFT_MAX_PASSES], off2[DFT_MAX_PASSES];
	const double2 *tl;	/* exp(-2 pi i e / N), e < 1024 */
	const double2 *th;	/* exp(-2 pi i 1024 e / N), e < N / 1024 + 1 */
	const double2 *p1;	/* exp(-2 pi i k1 / n), k1 < N1 */
	const double2 *p2;	/* exp(-2 pi i N1 k2 / n), k2 < N2 */
	/* tables of the n = 2^20 kernels (k_dft_fast.cu, N = 256 x 2048), NULL otherwise */
	const double2 *q_twA;	/* [15][16]:  exp(-2 pi i k r / 256),  r = 1..15, k < 16  (radix-16 pass, Ns = 16) */
	const double2 *q_twB;	/* [7][256]:  exp(-2 pi i j r / 2048), r = 1..7,  j < 256 (radix-8 pass, Ns = 256) */
	const double2 *q_e1;	/* [16][2048]: exp(-2 pi i j2 t / N) */
	const double2 *q_e2;	/* [2048]: exp(-2 pi i 16 j2 / N) */
	const double2 *q_p1;	/* [256]:  exp(-2 pi i k1 / n) */
	const double2 *q_p2;	/* [2048]: exp(-2 pi i 256 k2 / n) */
	/* tables of the n = 10^6 kernels (k_dft_1e6.cu, N = 500 x 1000), NULL otherwise */
	const double2 *m_twA;	/* [9][5]:    exp(-2 pi i k r / 50),   r = 1..9, k < 5   (column pass, Ns = 5) */
	const double2 *m_twB;	/* [9][50]:   exp(-2 pi i k r / 500),  k < 50            (column pass, Ns = 50) */
	const double2 *m_twC;	/* [9][10]:   exp(-2 pi i k r / 100),  k < 10            (row pass, Ns = 10) */
	const double2 *m_twD;	/* [9][100]:  exp(-2 pi i k r / 1000), k < 100           (row pass, Ns = 100) */
	const double2 *m_e1;	/* [50][1000]: exp(-2 pi i j2 t / N) */
	const double2 *m_e2;	/* [1000]: exp(-2 pi i 50 j2 / N) */
	const double2 *m_p1;	/* [500]:  exp(-2 pi i k1 / n) */
	const double2 *m_p2;	/* [1000]: exp(-2 pi i 500 k2 / n) */
	/* tables of the n = 10^7 kernels (k_dft_1e7.cu, N = 2000 x 2500, fast == 4), NULL otherwise */
	const double2 *v_twA;	/* [9][10]:   exp(-2 pi i k r / 100),  r = 1..9, k < 10   (Ns = 10, columns and rows) */
	const double2 *v_twB;	/* [9][100]:  exp(-2 pi i k r / 1000), k < 100            (column pass, Ns = 100) */
	const double2 *v_twC;	/* [1000]:    exp(-2 pi i k / 2000)                       (column pass, radix 2, Ns = 1000) */
	const double2 *v_rtwB;	/* [4][100]:  exp(-2 pi i k r / 500),  r = 1..4, k < 100  (row pass, radix 5, Ns = 100) */
	const double2 *v_rtwC;	/* [4][500]:  exp(-2 pi i k r / 2500), k < 500            (row pass, radix 5, Ns = 500) */
	const double2 *v_e1;	/* [200][2500]: exp(-2 pi i j2 t / N) */
	const double2 *v_e2;	/* [2500]: exp(-2 pi i 200 j2 / N) */
	const double2 *v_p1;	/* [2000]: exp(-2 pi i k1 / n) */
	const double2 *v_p2;	/* [500]:  exp(-2 pi i 2000 j / n) */
	/* Bluestein path for n/2 that is not 5-smooth (k_dft_any.cu, fast == 3), NULL / 0 otherwise */
	long long bs_M;		/* convolution length, the power of two >= 2 N - 1 */
	int bs_logM;
	const double2 *bs_chirp;	/* [N]:       exp(-i pi j^2 / N) */
	const double2 *bs_bhat;	/* [M]:       FFT_M of the wrapped conjugate chirp, divided by M */
	const double2 *bs_tw;	/* [M/2]:     exp(-2 pi i t / M) */
	const double2 *bs_wq;	/* [N/2 + 1]: exp(-2 pi i q / n) */
	double T;		/* sqrt(log(20) * n), discreteFourierTransform.c:142 */
};

/* host side (k_dft.cu): split N = n/2 into N1 x N2 and radix lists; false if n/2 is not 5-smooth (or too long) */
bool dft_make_plan(long long n, DftPlan *D);
/* host side (k_dft_any.cu): the Bluestein plan for every other even n up to 2^25; tables; launches per chunk */
bool dft_any_plan(long long n, DftPlan *D);
cudaError_t dft_any_setup(long long n, DftPlan *D, double2 *tabs[4], int num_sms);
int dft_any_launches_per_chunk(const DftPlan &D);
