/*
 * k_dft_fast.cu - the DFT test's four-step FFT specialised for n = 2^20 (N = 2^19 = 512 x 1024), the
 * stream length of the headline benchmark.  Same algorithm and same tables as k_dft.cu (which stays
 * as the generic power-of-two path); here every size is a compile-time constant so the index
 * arithmetic folds away, every table is laid out so that a warp's lookups are contiguous (the LSU
 * wavefront count, not FP64, bounded the first version), each CTA is 256 threads with <= 85 registers (3 CTAs per SM), and the
 * real-FFT post pass handles bins q and N - q together:
 *     A = (Z_q + conj Z_{N-q}) / 2,  B = -(i/2) W_n^q (Z_q - conj Z_{N-q})
 *     X_q = A + B,   X_{N-q} = conj(A - B)
 * and compares squared moduli with T^2 (falling back to the reference's sqrt form only inside a
 * 1e-10 relative band around T, where the north star allows N_1 to differ anyway).
 */
#include "common.cuh"
#include "dft.h"
#include "dft_math.cuh"

#define FN1 512
#define FN2 1024
#define FCW 4			/* columns per CTA in the column pass */
#define ROWLEN (FN2 + FN2 / 8 + 1)

__global__ void __launch_bounds__(256, 3)
dft_cols_fast(DevParams P, DftPlan D, const uint32_t *__restrict__ in, double2 *__restrict__ Y, int64_t stream0)
{
	extern __shared__ double2 buf[];			/* [padi(512)][4] */
	const int c = threadIdx.x & 3, t = threadIdx.x >> 2;	/* t < 64 */
	const int j2 = blockIdx.y * FCW + c;
	const uint32_t *__restrict__ w = in + (stream0 + blockIdx.x) * P.pitch_words + (j2 >> 4);
	double2 *__restrict__ Ys = Y + (int64_t) blockIdx.x * ((int64_t) FN1 * FN2);
	const int sh = 30 - ((2 * j2) & 31);
	double2 v[8];

	/* pass 0 (Ns = 1): z[1024 j1 + j2] straight from the packed bits, j1 = t + 64 r */
#pragma unroll
	for (int r = 0; r < 8; r++) {
		const uint32_t two = (ldw(w, 64 * (t + 64 * r)) >> sh) & 3u;
		v[r] = make_double2((two & 2u) ? 1.0 : -1.0, (two & 1u) ? 1.0 : -1.0);
	}
	bfly<8>(v);
#pragma unroll
	for (int r = 0; r < 8; r++)
		buf[padi(8 * t + r) * FCW + c] = v[r];
	__syncthreads();

	/* pass 1 (Ns = 8) */
	{
#pragma unroll
		for (int r = 0; r < 8; r++)
			v[r] = buf[padi(t + 64 * r) * FCW + c];
		const int k = t & 7;
#pragma unroll
		for (int r = 1; r < 8; r++)
			v[r] = cmul(v[r], __ldg(D.tw8 + (r - 1) * 8 + k));
		bfly<8>(v);
		__syncthreads();
		const int base = (t - k) * 8 + k;
#pragma unroll
		for (int r = 0; r < 8; r++)
			buf[padi(base + 8 * r) * FCW + c] = v[r];
		__syncthreads();
	}

	/* pass 2 (Ns = 64): output index k1 = t + 64 r, times W_N^(j2 k1), to Y[k1][j2] */
#pragma unroll
	for (int r = 0; r < 8; r++)
		v[r] = buf[padi(t + 64 * r) * FCW + c];
#pragma unroll
	for (int r = 1; r < 8; r++)
		v[r] = cmul(v[r], __ldg(D.tw64 + (r - 1) * 64 + t));
	bfly<8>(v);
#pragma unroll
	for (int r = 0; r < 8; r++) {
		Ys[(t + 64 * r) * FN2 + j2] = v[r];	/* the W_N^(j2 k1) factor is applied by the row pass */
	}
}

__global__ void __launch_bounds__(256, 3)
dft_rows_fast(DevParams P, DftPlan D, const double2 *__restrict__ Y, long long *__restrict__ st, int64_t stream0)
{
	extern __shared__ double2 buf[];			/* [2][ROWLEN] */
	__shared__ unsigned int cnt_s;
	const double2 *__restrict__ Ys = Y + (int64_t) blockIdx.x * ((int64_t) FN1 * FN2);
	const int pi = blockIdx.y;				/* row pair (pi, 512 - pi), pi = 0 .. 256 */
	const int k1a = pi, k1b = (FN1 - pi) & (FN1 - 1);
	const bool self = (k1a == k1b);
	const int rho = threadIdx.x >> 7, t = threadIdx.x & 127;
	const bool active = (rho == 0) || !self;
	const double2 *__restrict__ src = Ys + (rho ? k1b : k1a) * FN2;
	const double2 *__restrict__ twr = D.tw2d + (rho ? k1b : k1a) * FN2;
	double2 *row = buf + rho * ROWLEN;
	if (threadIdx.x == 0)
		cnt_s = 0;
	double2 v[8];

	/* pass 0 (Ns = 1) */
	if (active) {
#pragma unroll
		for (int r = 0; r < 8; r++)
			v[r] = cmul(src[t + 128 * r], __ldg(twr + t + 128 * r));	/* times W_N^(j2 k1) */
		bfly<8>(v);
#pragma unroll
		for (int r = 0; r < 8; r++)
			row[padi(8 * t + r)] = v[r];
	}
	__syncthreads();
	/* pass 1 (Ns = 8) and pass 2 (Ns = 64) */
#pragma unroll
	for (int pass = 1; pass <= 2; pass++) {
		const int Ns = (pass == 1) ? 8 : 64;
		const int k = t & (Ns - 1);
		if (active) {
#pragma unroll
			for (int r = 0; r < 8; r++)
				v[r] = row[padi(t + 128 * r)];
			const double2 *__restrict__ tw = (pass == 1) ? D.tw8 : D.tw64;
#pragma unroll
			for (int r = 1; r < 8; r++)
				v[r] = cmul(v[r], __ldg(tw + (r - 1) * Ns + k));
			bfly<8>(v);
		}
		__syncthreads();
		if (active) {
			const int base = (t - k) * 8 + k;
#pragma unroll
			for (int r = 0; r < 8; r++)
				row[padi(base + r * Ns)] = v[r];
		}
		__syncthreads();
	}
	/* pass 3 (radix 2, Ns = 512): in place, natural order */
	if (active) {
#pragma unroll
		for (int u = 0; u < 4; u++) {
			const int j = t + 128 * u;
			const double2 a = row[padi(j)];
			const double2 b = cmul(row[padi(j + 512)], __ldg(D.wm + j));
			row[padi(j)] = cadd(a, b);
			row[padi(j + 512)] = csub(a, b);
		}
	}
	__syncthreads();

	/* post pass: pairs (q, N - q), q = k1a + 512 k2 in row a, partner in row b */
	const double2 *ra = buf, *rb = self ? buf : buf + ROWLEN;
	const double T2 = D.T * D.T;
	const double2 w1 = __ldg(D.p1 + k1a);
	unsigned int cnt = 0;
#pragma unroll
	for (int u = 0; u < 4; u++) {
		const int k2 = threadIdx.x + 256 * u;
		const int k2p = (k1a == 0) ? ((FN2 - k2) & (FN2 - 1)) : (FN2 - 1 - k2);
		if (self && k2 > k2p)
			continue;				/* each unordered pair once */
		const double2 zq = ra[padi(k2)];
		const double2 zp = rb[padi(k2p)];
		const double2 wq = cmul(w1, __ldg(D.p2 + k2));	/* exp(-2 pi i q / n) */
		const double2 sum = make_double2(zq.x + zp.x, zq.y - zp.y);	/* Z_q + conj Z_{N-q} */
		const double2 dif = make_double2(zq.x - zp.x, zq.y + zp.y);	/* Z_q - conj Z_{N-q} */
		const double2 wd = cmul(wq, dif);
		const double ar = sum.x + wd.y, ai = sum.y - wd.x;		/* 2 X_q */
		const double br = sum.x - wd.y, bi = sum.y + wd.x;		/* 2 conj X_{N-q} */
		const double m2a = 0.25 * (ar * ar + ai * ai), m2b = 0.25 * (br * br + bi * bi);
		bool below_a = m2a < T2, below_b = m2b < T2;
		if (fabs(m2a - T2) < 1e-10 * T2)
			below_a = sqrt(0.25 * ar * ar + 0.25 * ai * ai) < D.T;
		if (fabs(m2b - T2) < 1e-10 * T2)
			below_b = sqrt(0.25 * br * br + 0.25 * bi * bi) < D.T;
		cnt += below_a ? 1u : 0u;
		if (!(self && k2 == k2p))			/* q = 0 and q = N/2 are their own partners */
			cnt += below_b ? 1u : 0u;
	}
	cnt = (unsigned int) warp_sum((int) cnt);
	if ((threadIdx.x & 31) == 0 && cnt)
		atomicAdd(&cnt_s, cnt);
	__syncthreads();
	if (threadIdx.x == 0 && cnt_s)
		atomicAdd((unsigned long long *) (st + (stream0 + blockIdx.x) * P.nst + P.st_off[STSGPU_TEST_DFT]),
			  (unsigned long long) cnt_s);
}

cudaError_t launch_dft_fast(const DevParams &P, const LaunchCtx &L, const DftPlan &D, double2 *d_work, int64_t work_streams)
{
	const size_t smem1 = (size_t) (FN1 + FN1 / 8 + 1) * FCW * sizeof(double2);
	const size_t smem2 = (size_t) 2 * ROWLEN * sizeof(double2);
	cudaError_t e;
	e = cudaFuncSetAttribute(dft_cols_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem1);
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(dft_rows_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem2);
	if (e != cudaSuccess) return e;
	for (int64_t s0 = 0; s0 < L.nstreams; s0 += work_streams) {
		const int64_t cs = (L.nstreams - s0 < work_streams) ? L.nstreams - s0 : work_streams;
		dim3 g1((unsigned) cs, FN2 / FCW);
		dft_cols_fast<<<g1, 256, smem1, L.stream>>>(P, D, L.d_in, d_work, s0);
		dim3 g2((unsigned) cs, FN1 / 2 + 1);
		dft_rows_fast<<<g2, 256, smem2, L.stream>>>(P, D, d_work, L.d_st, s0);
	}
	return cudaGetLastError();
}
