/*
 * k_dft_fast.cu - the DFT test's four-step FFT specialised for n = 2^20 (N = n/2 = 2^19 complex
 * points), the stream length of the headline benchmark.  k_dft.cu stays as the generic
 * power-of-two path.
 *
 * The first version (512 x 1024, radix 8) was bound by the LSU wavefront pipe: six shared-memory
 * exchanges per point plus scattered twiddle lookups.  This one is built to move less:
 *
 *   N = 256 x 2048.  Sixteen points per thread (radix 16 = 4 x 4 in registers) so the columns need
 *   ONE exchange (16 x 16) and the rows TWO (16 x 16 x 8); every table is laid out so that a warp's
 *   lookups are contiguous.
 *
 *   dft_cols16 : 16 adjacent columns j2 per CTA (one packed input word per row j1), 16 threads per
 *                column.  C[k1][j2] = sum_j1 z[2048 j1 + j2] W_256^(j1 k1), times W_N^(j2 k1)
 *                (one table value per thread, then powers of W_N^(16 j2) by repeated multiply),
 *                stored as Y[k1][j2] in 256-byte runs.
 *   dft_rows16 : rows (a, 256 - a) per CTA.  Bins q = a + 256 k2 and N - q = (256 - a) +
 *                256 (2047 - k2) are partners of the real-FFT post pass
 *                    X_q = (Z_q + conj Z_{N-q})/2 - (i/2) W_n^q (Z_q - conj Z_{N-q});
 *                in the last (radix 8) pass thread u takes butterfly u of row a and butterfly
 *                255 - u of row b, whose outputs are exactly each other's partners, so the post pass
 *                and the |X| < T count run on registers and nothing but the count leaves the CTA.
 *                Rows 0 and 128 are their own partners and finish through shared memory.
 *
 * Moduli are compared squared, |2X|^2 < 4 T^2: equal to the reference's sqrt form except within one
 * rounding of T, far inside the 1e-9 relative band where the north star allows N_1 to differ.
 *
 * Per stream: 8 MiB written + 8 MiB read of Y (HBM/L2), three shared-memory exchanges, about
 * 80 FP64 instructions per complex point.
 */
#include "common.cuh"
#include "dft.h"
#include "dft_math.cuh"
#include "dft16.cuh"

#ifndef DFT_MINCTA
#define DFT_MINCTA 2
#endif
/*
 * The row pass carries no __launch_bounds__: its register count is set by -maxrregcount=$(DFT_MAXREG) in the Makefile
 * (nvcc ignores that flag for kernels with launch bounds).  It compiles without spills down to 80 registers; fewer
 * registers make the DFT alone slower but leave room for the other kernels' CTAs beside it.  Measured per 1024
 * streams, DFT group alone / pipelined batch: 114 registers 4.26 / 9.26 ms, 104: 4.29 / 9.23, 96: 4.32 / 9.20,
 * 88: 4.33 / 9.13, 80 (three CTAs per SM): 4.74 / 9.07.  -DDFT_ROWS_BOUNDS restores the bounds.
 */
#ifdef DFT_ROWS_BOUNDS
#define DFT_BOUNDS __launch_bounds__(256, DFT_MINCTA)
#else
#define DFT_BOUNDS
#endif
#ifndef DFT_NO_TW_POWERS
#define DFT_TW_POWERS 1	/* fewer table lookups: r = 1, 2, 4, 8 loaded, the other twiddle powers by products */
#endif
#ifndef QCW
#define QCW 8				/* columns per CTA in the column pass (16 threads per column): 128-thread CTAs, four per SM.  Measured DFT group / pipelined batch per 1024 streams: QCW 16 4.45 / 9.84 ms, 8 4.25 / 9.35 ms, 4 4.73 / 9.85 ms */
#endif
#ifndef QCOLS_MINCTA
#define QCOLS_MINCTA (DFT_MINCTA * 16 / QCW)
#endif
#define QCOLS_BOUNDS __launch_bounds__(16 * QCW, QCOLS_MINCTA)
/* ------------------------------------------------------------------------------------------ */
__global__ void QCOLS_BOUNDS
dft_cols16(DevParams P, DftPlan D, const uint32_t *__restrict__ in, double2 *__restrict__ Y, int64_t stream0)
{
	extern __shared__ double2 buf[];			/* [256][QCW] */
	const int c = threadIdx.x % QCW, t = threadIdx.x / QCW;	/* t < 16 */
	const int j2 = blockIdx.y * QCW + c;
	const uint32_t *__restrict__ w = in + (stream0 + blockIdx.x) * P.pitch_words + ((2 * j2) >> 5);
	double2 *__restrict__ Ys = Y + (int64_t) blockIdx.x * ((int64_t) QN1 * QN2);
	const int sh = 30 - ((2 * j2) & 31);
	double2 v[16];

	/* pass 0 (Ns = 1): z[2048 j1 + j2] straight from the packed bits, j1 = t + 16 r */
#pragma unroll
	for (int r = 0; r < 16; r++) {
		const uint32_t two = (ldw(w, 128 * (t + 16 * r)) >> sh) & 3u;
		v[r] = make_double2((two & 2u) ? 1.0 : -1.0, (two & 1u) ? 1.0 : -1.0);
	}
	bfly16(v);
#pragma unroll
	for (int r = 0; r < 16; r++)
		buf[(16 * t + r) * QCW + c] = v[r];
	__syncthreads();

	/* pass 1 (Ns = 16): k = t, outputs k1 = t + 16 r */
	v[0] = buf[t * QCW + c];
#ifdef DFT_TW_POWERS
	{
		double2 w[16];
		w[1] = __ldg(D.q_twA + t); w[2] = __ldg(D.q_twA + 16 + t); w[4] = __ldg(D.q_twA + 3 * 16 + t);
		w[8] = __ldg(D.q_twA + 7 * 16 + t);
		w[3] = cmul(w[1], w[2]); w[5] = cmul(w[4], w[1]); w[6] = cmul(w[4], w[2]); w[7] = cmul(w[4], w[3]);
#pragma unroll
		for (int r = 9; r < 16; r++)
			w[r] = cmul(w[8], w[r - 8]);
#pragma unroll
		for (int r = 1; r < 16; r++)
			v[r] = cmul(buf[(t + 16 * r) * QCW + c], w[r]);
	}
#else
#pragma unroll
	for (int r = 1; r < 16; r++)
		v[r] = cmul(buf[(t + 16 * r) * QCW + c], __ldg(D.q_twA + (r - 1) * 16 + t));
#endif
	bfly16(v);
	/* times W_N^(j2 k1) = W_N^(j2 t) (W_N^(16 j2))^r */
	double2 tw = __ldg(D.q_e1 + t * QN2 + j2);
	const double2 step = __ldg(D.q_e2 + j2);
#pragma unroll
	for (int r = 0; r < 16; r++) {
		Ys[(t + 16 * r) * QN2 + j2] = cmul(v[r], tw);
		tw = cmul(tw, step);
	}
}

/* ------------------------------------------------------------------------------------------ */
__global__ void DFT_BOUNDS
dft_rows16(DevParams P, DftPlan D, const double2 *__restrict__ Y, long long *__restrict__ st, int64_t stream0)
{
	extern __shared__ double2 buf[];			/* [2][QROW] */
	__shared__ unsigned int cnt_s;
	const double2 *__restrict__ Ys = Y + (int64_t) blockIdx.x * ((int64_t) QN1 * QN2);
	const int ra = blockIdx.y, rb = (QN1 - ra) & (QN1 - 1);	/* row pair, ra = 0 .. 128 */
	const bool self = (ra == rb);
	const int rho = threadIdx.x >> 7, t = threadIdx.x & 127;
	const bool active = (rho == 0) || !self;
	double2 *row = buf + rho * QROW;
	if (threadIdx.x == 0)
		cnt_s = 0;
	double2 v[16];

	/* pass 0 (Ns = 1): inputs t + 128 r, outputs 16 t + r */
	if (active) {
		const double2 *__restrict__ src = Ys + (rho ? rb : ra) * QN2;
#pragma unroll
		for (int r = 0; r < 16; r++)
			v[r] = src[t + 128 * r];
		bfly16(v);
#pragma unroll
		for (int r = 0; r < 16; r++)
			row[pad16(16 * t + r)] = v[r];
	}
	__syncthreads();
	/* pass 1 (Ns = 16): k = t mod 16, outputs 16 (t - k) + k + 16 r */
	{
		const int k = t & 15;
		if (active) {
			v[0] = row[pad16(t)];
#ifdef DFT_TW_POWERS
			{
				double2 w[16];
				w[1] = __ldg(D.q_twA + k); w[2] = __ldg(D.q_twA + 16 + k); w[4] = __ldg(D.q_twA + 3 * 16 + k);
				w[8] = __ldg(D.q_twA + 7 * 16 + k);
				w[3] = cmul(w[1], w[2]); w[5] = cmul(w[4], w[1]); w[6] = cmul(w[4], w[2]); w[7] = cmul(w[4], w[3]);
#pragma unroll
				for (int r = 9; r < 16; r++)
					w[r] = cmul(w[8], w[r - 8]);
#pragma unroll
				for (int r = 1; r < 16; r++)
					v[r] = cmul(row[pad16(t + 128 * r)], w[r]);
			}
#else
#pragma unroll
			for (int r = 1; r < 16; r++)
				v[r] = cmul(row[pad16(t + 128 * r)], __ldg(D.q_twA + (r - 1) * 16 + k));
#endif
			bfly16(v);
		}
		__syncthreads();
		if (active) {
			const int base = 16 * (t - k) + k;
#pragma unroll
			for (int r = 0; r < 16; r++)
				row[pad16(base + 16 * r)] = v[r];
		}
		__syncthreads();
	}
	/* pass 2 (radix 8, Ns = 256): butterfly j reads j + 256 r and yields Z[k2 = j + 256 r] */
	const double T2x4 = 4.0 * D.T * D.T;
	const double2 w1 = __ldg(D.q_p1 + ra);			/* exp(-2 pi i ra / n) */
	unsigned int cnt = 0;
	if (!self) {
		const int ja = threadIdx.x, jb = 255 - threadIdx.x;
		double2 za[8], zb[8];
#ifdef DFT_TW_POWERS
		{	/* twiddles W_2048^(j r): r = 1, 2, 4 from the table, the rest by products */
			double2 wa[8], wb[8];
			wa[1] = __ldg(D.q_twB + ja); wa[2] = __ldg(D.q_twB + 256 + ja); wa[4] = __ldg(D.q_twB + 3 * 256 + ja);
			wb[1] = __ldg(D.q_twB + jb); wb[2] = __ldg(D.q_twB + 256 + jb); wb[4] = __ldg(D.q_twB + 3 * 256 + jb);
			wa[3] = cmul(wa[1], wa[2]); wa[5] = cmul(wa[4], wa[1]); wa[6] = cmul(wa[4], wa[2]); wa[7] = cmul(wa[4], wa[3]);
			wb[3] = cmul(wb[1], wb[2]); wb[5] = cmul(wb[4], wb[1]); wb[6] = cmul(wb[4], wb[2]); wb[7] = cmul(wb[4], wb[3]);
			za[0] = buf[pad16(ja)];
			zb[0] = buf[QROW + pad16(jb)];
#pragma unroll
			for (int r = 1; r < 8; r++) {
				za[r] = cmul(buf[pad16(ja + 256 * r)], wa[r]);
				zb[r] = cmul(buf[QROW + pad16(jb + 256 * r)], wb[r]);
			}
		}
#else
		za[0] = buf[pad16(ja)];
		zb[0] = buf[QROW + pad16(jb)];
#pragma unroll
		for (int r = 1; r < 8; r++) {
			za[r] = cmul(buf[pad16(ja + 256 * r)], __ldg(D.q_twB + (r - 1) * 256 + ja));
			zb[r] = cmul(buf[QROW + pad16(jb + 256 * r)], __ldg(D.q_twB + (r - 1) * 256 + jb));
		}
#endif
		bfly<8>(za);
		bfly<8>(zb);
		/* bin q = ra + 256 (ja + 256 r); its partner N - q = rb + 256 (jb + 256 (7 - r)) */
		/* W_n^q = W_n^(ra + 256 ja) W_16^r: one table value, then the sixteenth roots of unity */
		const double2 wq0 = cmul(w1, __ldg(D.q_p2 + ja));
#pragma unroll
		for (int r = 0; r < 8; r++)
			cnt += post_pair(za[r], zb[7 - r], cmul_w16(wq0, r), T2x4, true);
	} else {
		const int ja = threadIdx.x;
		double2 za[8];
		za[0] = buf[pad16(ja)];
#pragma unroll
		for (int r = 1; r < 8; r++)
			za[r] = cmul(buf[pad16(ja + 256 * r)], __ldg(D.q_twB + (r - 1) * 256 + ja));
		bfly<8>(za);
		__syncthreads();
#pragma unroll
		for (int r = 0; r < 8; r++)
			buf[pad16(ja + 256 * r)] = za[r];
		__syncthreads();
		/* row 0: N - q = 256 (2048 - k2); row 128: N - q = 128 + 256 (2047 - k2); each unordered pair once */
#pragma unroll
		for (int r = 0; r < 8; r++) {
			const int k2 = ja + 256 * r;
			const int k2p = (ra == 0) ? ((QN2 - k2) & (QN2 - 1)) : (QN2 - 1 - k2);
			if (k2 > k2p)
				continue;
			const double2 wq = cmul(w1, __ldg(D.q_p2 + k2));
			cnt += post_pair(buf[pad16(k2)], buf[pad16(k2p)], wq, T2x4, k2 != k2p);
		}
	}
	cnt = (unsigned int) warp_sum((int) cnt);
	if ((threadIdx.x & 31) == 0 && cnt)
		atomicAdd(&cnt_s, cnt);
	__syncthreads();
	if (threadIdx.x == 0 && cnt_s)
		atomicAdd((unsigned long long *) (st + (stream0 + blockIdx.x) * P.nst + P.st_off[STSGPU_TEST_DFT]),
			  (unsigned long long) cnt_s);
}

cudaError_t launch_dft_rows_tma(const DevParams &P, const DftPlan &D, const double2 *d_work, long long *d_st, int64_t stream0, int64_t cs,
				int num_sms, cudaStream_t stream);

cudaError_t launch_dft_fast(const DevParams &P, const LaunchCtx &L, const DftPlan &D, double2 *d_work, int64_t work_streams)
{
	/* row pass: the persistent TMA-fed kernel of k_dft_rows_tma.cu, or (STSGPU_DFT_ROWS_TMA=0) dft_rows16 below */
	static int rows_tma = -1;
	if (rows_tma < 0)
		rows_tma = getenv("STSGPU_DFT_ROWS_TMA") ? atoi(getenv("STSGPU_DFT_ROWS_TMA")) : 0;
	const size_t smem1 = (size_t) QN1 * QCW * sizeof(double2);
	const size_t smem2 = (size_t) 2 * QROW * sizeof(double2);
	cudaError_t e;
	e = cudaFuncSetAttribute(dft_cols16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem1);
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(dft_rows16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem2);
	if (e != cudaSuccess) return e;
	/*
	 * Chunks of `work_streams` streams alternate between two streams with one work buffer each, so the
	 * column pass of chunk i+1 runs beside the row pass of chunk i and the tail of one grid is filled
	 * by the head of the next.
	 */
	const bool two = L.stream2 != nullptr && L.stream2 != L.stream && L.nstreams > work_streams;
	if (two) {
		e = cudaEventRecord(L.ev2_fork, L.stream);
		if (e != cudaSuccess) return e;
		e = cudaStreamWaitEvent(L.stream2, L.ev2_fork, 0);
		if (e != cudaSuccess) return e;
	}
	int which = 0;
	for (int64_t s0 = 0; s0 < L.nstreams; s0 += work_streams, which ^= 1) {
		const int64_t cs = (L.nstreams - s0 < work_streams) ? L.nstreams - s0 : work_streams;
		cudaStream_t st = (two && which) ? L.stream2 : L.stream;
		double2 *wk = d_work + (size_t) which * work_streams * ((size_t) QN1 * QN2);
		dim3 g1((unsigned) cs, QN2 / QCW);
		dft_cols16<<<g1, 16 * QCW, smem1, st>>>(P, D, L.d_in, wk, s0);
		if (rows_tma) {
			e = launch_dft_rows_tma(P, D, wk, L.d_st, s0, cs, L.num_sms, st);
			if (e != cudaSuccess) return e;
		} else {
			dim3 g2((unsigned) cs, QN1 / 2 + 1);
			dft_rows16<<<g2, 256, smem2, st>>>(P, D, wk, L.d_st, s0);
		}
	}
	if (two) {
		e = cudaEventRecord(L.ev2_join, L.stream2);
		if (e != cudaSuccess) return e;
		e = cudaStreamWaitEvent(L.stream, L.ev2_join, 0);
		if (e != cudaSuccess) return e;
	}
	return cudaGetLastError();
}
