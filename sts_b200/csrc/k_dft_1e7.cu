/*
 * k_dft_1e7.cu - the DFT test's four-step FFT specialised for n = 10^7, the stream length of BASELINE config 4
 * (N = n/2 = 5 000 000 = 2^6 5^7 complex points).  Same design as k_dft_1e6.cu: radix 10 = 2 x 5 (prime-factor
 * butterfly, no inner twiddles) in registers, Stockham passes in ONE shared-memory buffer with the butterflies held in
 * registers across the barrier, and the real-FFT post pass finished on registers by pairing mirrored butterflies.
 * The generic kernels of k_dft.cu took this length at 128 registers with up to 1.5 KB of spills per thread, a
 * shared-memory post pass and run-time radix lists: 78 ms per 428 streams, 4.4 times the n = 2^20 kernels per point.
 *
 *   N = 2000 x 2500,  j = 2500 j1 + j2,  k = k1 + 2000 k2.
 *   dft_cols_1e7 : 2 adjacent columns j2 per CTA, 200 threads per column, ten points per thread.  Passes 10, 10, 10, 2
 *                  (the first straight from the packed bits); C[k1][j2] times W_N^(j2 k1) (one table value per thread,
 *                  then powers of W_N^(200 j2)), stored as Y[k1][j2].
 *   dft_rows_1e7 : rows (a, 2000 - a) per CTA, 250 threads per row, passes 10, 10, 5, 5.  In the last pass thread u
 *                  takes butterfly u of row a and butterfly 499 - u of row b: bin q = a + 2000 (u + 500 r) and its
 *                  partner N - q = (2000 - a) + 2000 (499 - u + 500 (4 - r)) are then both in its registers.
 *                  W_n^q = W_n^a W_n^(2000 u) W_10^r.  Rows 0 and 1000 are their own partners and finish through
 *                  shared memory.
 */
#include "common.cuh"
#include "dft.h"
#include "dft_math.cuh"

#define VN1 2000
#define VN2 2500
#define VCW 2
#define VCT 200				/* threads per column */
#define VRT 250				/* threads per row */
#define VCOLP (VN1 + VN1 / 10)		/* padded column: one spare element per 10 */
#define VROWP (VN2 + VN2 / 10)

__device__ __forceinline__ int vpad(int e) { return e + e / 10; }

/* x times W_10^r = exp(-2 pi i r / 10), r = 0..4 */
__device__ __forceinline__ double2 cmul_w10(double2 x, int r)
{
	switch (r) {
	case 0: return x;
	case 1: return cmul(x, make_double2(0.80901699437494742410, -0.58778525229247312917));
	case 2: return cmul(x, make_double2(0.30901699437494742410, -0.95105651629515357212));
	case 3: return cmul(x, make_double2(-0.30901699437494742410, -0.95105651629515357212));
	default: return cmul(x, make_double2(-0.80901699437494742410, -0.58778525229247312917));
	}
}

/* ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(VCW * VCT, 2)
dft_cols_1e7(DevParams P, DftPlan D, const uint32_t *__restrict__ in, double2 *__restrict__ Y, int64_t stream0)
{
	extern __shared__ double2 buf[];			/* [VCOLP][VCW] */
	const int c = threadIdx.x % VCW, t = threadIdx.x / VCW;	/* t < 200 */
	const int j2 = blockIdx.y * VCW + c;
	const uint32_t *__restrict__ w = in + (stream0 + blockIdx.x) * P.pitch_words;
	double2 *__restrict__ Ys = Y + (int64_t) blockIdx.x * ((int64_t) VN1 * VN2);
	double2 v[10];

	/* pass 0 (radix 10, Ns = 1): butterfly t reads z[2500 j1 + j2], j1 = t + 200 r; outputs 10 t + r */
#pragma unroll
	for (int r = 0; r < 10; r++) {
		const int64_t p = 2 * ((int64_t) VN2 * (t + 200 * r) + j2);		/* bit position of x_2j */
		const uint32_t two = (ldw(w, p >> 5) >> (30 - (int) (p & 31))) & 3u;
		v[r] = make_double2((two & 2u) ? 1.0 : -1.0, (two & 1u) ? 1.0 : -1.0);
	}
	bfly10(v);
#pragma unroll
	for (int r = 0; r < 10; r++)
		buf[vpad(10 * t + r) * VCW + c] = v[r];
	__syncthreads();

	/* pass 1 (radix 10, Ns = 10): k = t mod 10, outputs (t - k) 10 + k + 10 r */
	{
		const int k = t % 10;
		v[0] = buf[vpad(t) * VCW + c];
#pragma unroll
		for (int r = 1; r < 10; r++)
			v[r] = cmul(buf[vpad(t + 200 * r) * VCW + c], __ldg(D.v_twA + (r - 1) * 10 + k));
		bfly10(v);
		__syncthreads();
		const int base = (t - k) * 10 + k;
#pragma unroll
		for (int r = 0; r < 10; r++)
			buf[vpad(base + 10 * r) * VCW + c] = v[r];
		__syncthreads();
	}
	/* pass 2 (radix 10, Ns = 100): k = t mod 100, outputs (t - k) 10 + k + 100 r */
	{
		const int k = t % 100;
		v[0] = buf[vpad(t) * VCW + c];
#pragma unroll
		for (int r = 1; r < 10; r++)
			v[r] = cmul(buf[vpad(t + 200 * r) * VCW + c], __ldg(D.v_twB + (r - 1) * 100 + k));
		bfly10(v);
		__syncthreads();
		const int base = (t - k) * 10 + k;
#pragma unroll
		for (int r = 0; r < 10; r++)
			buf[vpad(base + 100 * r) * VCW + c] = v[r];
		__syncthreads();
	}
	/* pass 3 (radix 2, Ns = 1000): butterflies b = t + 200 i, i < 5: inputs b, b + 1000, outputs k1 = b + 1000 r */
	/* times W_N^(j2 k1), k1 = t + 200 m, m = i + 5 r: W_N^(j2 t) (W_N^(200 j2))^m */
	{
		double2 lo[5], hi[5];
#pragma unroll
		for (int i = 0; i < 5; i++) {
			const int b = t + 200 * i;
			const double2 x0 = buf[vpad(b) * VCW + c];
			const double2 x1 = cmul(buf[vpad(b + 1000) * VCW + c], __ldg(D.v_twC + b));
			lo[i] = cadd(x0, x1);
			hi[i] = csub(x0, x1);
		}
		double2 tw = __ldg(D.v_e1 + (int64_t) t * VN2 + j2);
		const double2 step = __ldg(D.v_e2 + j2);
#pragma unroll
		for (int i = 0; i < 5; i++) {
			Ys[(int64_t) (t + 200 * i) * VN2 + j2] = cmul(lo[i], tw);
			tw = cmul(tw, step);
		}
#pragma unroll
		for (int i = 0; i < 5; i++) {
			Ys[(int64_t) (t + 200 * i + 1000) * VN2 + j2] = cmul(hi[i], tw);
			tw = cmul(tw, step);
		}
	}
}

/* ------------------------------------------------------------------------------------------ */
#ifndef V_ROWS_MINCTA
#define V_ROWS_MINCTA 2		/* measured per 428 streams: 1 CTA per SM (104 registers) 40.9 ms, 2 (64 registers, 8 bytes spilled) 37.3 ms */
#endif
__global__ void __launch_bounds__(2 * VRT, V_ROWS_MINCTA)
dft_rows_1e7(DevParams P, DftPlan D, const double2 *__restrict__ Y, long long *__restrict__ st, int64_t stream0)
{
	extern __shared__ double2 buf[];			/* [2][VROWP] */
	__shared__ unsigned int cnt_s;
	const double2 *__restrict__ Ys = Y + (int64_t) blockIdx.x * ((int64_t) VN1 * VN2);
	const int ra = blockIdx.y, rb = (VN1 - ra) % VN1;	/* row pair, ra = 0 .. 1000 */
	const bool self = (ra == rb);
	const int rho = threadIdx.x / VRT, t = threadIdx.x % VRT;
	const bool active = (rho == 0) || !self;
	double2 *row = buf + rho * VROWP;
	if (threadIdx.x == 0)
		cnt_s = 0;
	double2 v[10];

	/* pass 0 (radix 10, Ns = 1): inputs t + 250 r, outputs 10 t + r */
	if (active) {
		const double2 *__restrict__ src = Ys + (int64_t) (rho ? rb : ra) * VN2;
#pragma unroll
		for (int r = 0; r < 10; r++)
			v[r] = src[t + 250 * r];
		bfly10(v);
#pragma unroll
		for (int r = 0; r < 10; r++)
			row[vpad(10 * t + r)] = v[r];
	}
	__syncthreads();
	/* pass 1 (radix 10, Ns = 10): k = t mod 10, outputs 10 (t - k) + k + 10 r */
	{
		const int k = t % 10;
		if (active) {
			v[0] = row[vpad(t)];
#pragma unroll
			for (int r = 1; r < 10; r++)
				v[r] = cmul(row[vpad(t + 250 * r)], __ldg(D.v_twA + (r - 1) * 10 + k));
			bfly10(v);
		}
		__syncthreads();
		if (active) {
			const int base = 10 * (t - k) + k;
#pragma unroll
			for (int r = 0; r < 10; r++)
				row[vpad(base + 10 * r)] = v[r];
		}
		__syncthreads();
	}
	/* pass 2 (radix 5, Ns = 100): butterflies b = t, t + 250; k = b mod 100; inputs b + 500 r; outputs 5 (b - k) + k + 100 r */
	{
		if (active) {
#pragma unroll
			for (int h = 0; h < 2; h++) {
				const int b = t + 250 * h, k = b % 100;
				v[5 * h] = row[vpad(b)];
#pragma unroll
				for (int r = 1; r < 5; r++)
					v[5 * h + r] = cmul(row[vpad(b + 500 * r)], __ldg(D.v_rtwB + (r - 1) * 100 + k));
				bfly<5>(v + 5 * h);
			}
		}
		__syncthreads();
		if (active) {
#pragma unroll
			for (int h = 0; h < 2; h++) {
				const int b = t + 250 * h, k = b % 100;
				const int base = 5 * (b - k) + k;
#pragma unroll
				for (int r = 0; r < 5; r++)
					row[vpad(base + 100 * r)] = v[5 * h + r];
			}
		}
		__syncthreads();
	}
	/* pass 3 (radix 5, Ns = 500): butterfly j reads j + 500 r and yields Z[k2 = j + 500 r] */
	const double T2x4 = 4.0 * D.T * D.T;
	const double2 w1 = __ldg(D.v_p1 + ra);			/* exp(-2 pi i ra / n) */
	unsigned int cnt = 0;
	if (!self) {
		const int ja = threadIdx.x, jb = 499 - threadIdx.x;
		double2 za[5], zb[5];
		za[0] = buf[vpad(ja)];
		zb[0] = buf[VROWP + vpad(jb)];
#pragma unroll
		for (int r = 1; r < 5; r++) {
			za[r] = cmul(buf[vpad(ja + 500 * r)], __ldg(D.v_rtwC + (r - 1) * 500 + ja));
			zb[r] = cmul(buf[VROWP + vpad(jb + 500 * r)], __ldg(D.v_rtwC + (r - 1) * 500 + jb));
		}
		bfly<5>(za);
		bfly<5>(zb);
		/* bin q = ra + 2000 (ja + 500 r); its partner N - q = rb + 2000 (jb + 500 (4 - r)) */
		const double2 wq0 = cmul(w1, __ldg(D.v_p2 + ja));
#pragma unroll
		for (int r = 0; r < 5; r++)
			cnt += post_pair(za[r], zb[4 - r], cmul_w10(wq0, r), T2x4, true);
	} else {
		const int ja = threadIdx.x;			/* all 500 threads work on the one row */
		double2 za[5];
		za[0] = buf[vpad(ja)];
#pragma unroll
		for (int r = 1; r < 5; r++)
			za[r] = cmul(buf[vpad(ja + 500 * r)], __ldg(D.v_rtwC + (r - 1) * 500 + ja));
		bfly<5>(za);
		__syncthreads();
#pragma unroll
		for (int r = 0; r < 5; r++)
			buf[vpad(ja + 500 * r)] = za[r];
		__syncthreads();
		/* row 0: N - q = 2000 (2500 - k2); row 1000: N - q = 1000 + 2000 (2499 - k2); each unordered pair once */
#pragma unroll
		for (int r = 0; r < 5; r++) {
			const int k2 = ja + 500 * r;
			const int k2p = (ra == 0) ? ((VN2 - k2) % VN2) : (VN2 - 1 - k2);
			if (k2 > k2p)
				continue;
			/* W_n^q, q = ra + 2000 k2 = ra + 2000 ja + 10^6 r */
			const double2 wq = cmul_w10(cmul(w1, __ldg(D.v_p2 + ja)), r);
			cnt += post_pair(za[r], buf[vpad(k2p)], wq, T2x4, k2 != k2p);
		}
	}
	{	/* 500 threads = fifteen warps and 20 lanes: the reduction names exactly the lanes the warp has (k_dft_1e6.cu) */
		const unsigned int left = blockDim.x - (threadIdx.x & ~31u);
		const unsigned int lanes = left >= 32u ? 0xFFFFFFFFu : ((1u << left) - 1u);
		cnt = __reduce_add_sync(lanes, cnt);
	}
	if ((threadIdx.x & 31) == 0 && cnt)
		atomicAdd(&cnt_s, cnt);
	__syncthreads();
	if (threadIdx.x == 0 && cnt_s)
		atomicAdd((unsigned long long *) (st + (stream0 + blockIdx.x) * P.nst + P.st_off[STSGPU_TEST_DFT]),
			  (unsigned long long) cnt_s);
}

cudaError_t launch_dft_1e7(const DevParams &P, const LaunchCtx &L, const DftPlan &D, double2 *d_work, int64_t work_streams)
{
	const size_t smem1 = (size_t) VCOLP * VCW * sizeof(double2);
	const size_t smem2 = (size_t) 2 * VROWP * sizeof(double2);
	cudaError_t e;
	e = cudaFuncSetAttribute(dft_cols_1e7, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem1);
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(dft_rows_1e7, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem2);
	if (e != cudaSuccess) return e;
	/* chunks alternate between two streams with one work buffer each, as in k_dft_fast.cu */
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
		double2 *wk = d_work + (size_t) which * work_streams * ((size_t) VN1 * VN2);
		dim3 g1((unsigned) cs, VN2 / VCW);
		dft_cols_1e7<<<g1, VCW * VCT, smem1, st>>>(P, D, L.d_in, wk, s0);
		dim3 g2((unsigned) cs, VN1 / 2 + 1);
		dft_rows_1e7<<<g2, 2 * VRT, smem2, st>>>(P, D, wk, L.d_st, s0);
	}
	if (two) {
		e = cudaEventRecord(L.ev2_join, L.stream2);
		if (e != cudaSuccess) return e;
		e = cudaStreamWaitEvent(L.stream, L.ev2_join, 0);
		if (e != cudaSuccess) return e;
	}
	return cudaGetLastError();
}
