/*
 * k_dft_1e6.cu - the DFT test's four-step FFT specialised for n = 10^6, the stream length NIST
 * SP800-22 recommends (N = n/2 = 500000 = 2^5 5^6 complex points).  Same design as k_dft_fast.cu
 * (n = 2^20) with radix 10 = 2 x 5 in registers instead of radix 16:
 *
 *   N = 500 x 1000.
 *   dft_cols_1e6 : 8 adjacent columns j2 per CTA, 50 threads per column.  Passes 5 (two butterflies
 *                  per thread, straight from the packed bits), 10, 10; C[k1][j2] times W_N^(j2 k1)
 *                  (one table value per thread, then powers of W_N^(50 j2)), stored as Y[k1][j2] in
 *                  128-byte runs.
 *   dft_rows_1e6 : rows (a, 500 - a) per CTA, 100 threads per row, passes 10, 10, 10.  In the last
 *                  pass thread u takes butterfly u of row a and butterfly 99 - u of row b: bin
 *                  q = a + 500 (u + 100 r) and its partner N - q = (500 - a) + 500 (99 - u + 100 (9 - r))
 *                  of the real-FFT post pass are then both in its registers, so the post pass and the
 *                  |X| < T count never touch memory.  W_n^q = W_n^a W_n^(500 u) W_20^r: one table value
 *                  and the twentieth roots of unity.  Rows 0 and 250 are their own partners and finish
 *                  through shared memory.
 *
 * Per stream: 8 MB written + 8 MB read of Y, four shared-memory exchanges.
 */
#include "common.cuh"
#include "dft.h"
#include "dft_math.cuh"

#define MN1 500
#define MN2 1000
#define MCW 8
#define MROW (MN2 + MN2 / 10)		/* padded row: one spare element per 10 */

__device__ __forceinline__ int pad10(int e) { return e + e / 10; }

/* x times W_20^r = exp(-2 pi i r / 20), r = 0..9 (compile-time r after unrolling) */
__device__ __forceinline__ double2 cmul_w20(double2 x, int r)
{
	switch (r) {
	case 0: return x;
	case 1: return cmul(x, make_double2(0.95105651629515357212, -0.30901699437494742410));
	case 2: return cmul(x, make_double2(0.80901699437494742410, -0.58778525229247312917));
	case 3: return cmul(x, make_double2(0.58778525229247312917, -0.80901699437494742410));
	case 4: return cmul(x, make_double2(0.30901699437494742410, -0.95105651629515357212));
	case 5: return cmi(x);
	case 6: return cmul(x, make_double2(-0.30901699437494742410, -0.95105651629515357212));
	case 7: return cmul(x, make_double2(-0.58778525229247312917, -0.80901699437494742410));
	case 8: return cmul(x, make_double2(-0.80901699437494742410, -0.58778525229247312917));
	default: return cmul(x, make_double2(-0.95105651629515357212, -0.30901699437494742410));
	}
}

/* ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(400, 2)
dft_cols_1e6(DevParams P, DftPlan D, const uint32_t *__restrict__ in, double2 *__restrict__ Y, int64_t stream0)
{
	extern __shared__ double2 buf[];			/* [500][8] */
	const int c = threadIdx.x & 7, t = threadIdx.x >> 3;	/* t < 50 */
	const int j2 = blockIdx.y * MCW + c;
	const uint32_t *__restrict__ w = in + (stream0 + blockIdx.x) * P.pitch_words;
	double2 *__restrict__ Ys = Y + (int64_t) blockIdx.x * ((int64_t) MN1 * MN2);
	double2 v[10];

	/* pass 0 (radix 5, Ns = 1): butterflies t and t + 50 read z[1000 j1 + j2], j1 = b + 100 r; outputs 5 b + r */
#pragma unroll
	for (int h = 0; h < 2; h++) {
		const int b = t + 50 * h;
#pragma unroll
		for (int r = 0; r < 5; r++) {
			const int p = 2 * (MN2 * (b + 100 * r) + j2);		/* bit position of x_2j */
			const uint32_t two = (ldw(w, p >> 5) >> (30 - (p & 31))) & 3u;
			v[5 * h + r] = make_double2((two & 2u) ? 1.0 : -1.0, (two & 1u) ? 1.0 : -1.0);
		}
		bfly<5>(v + 5 * h);
#pragma unroll
		for (int r = 0; r < 5; r++)
			buf[(5 * b + r) * MCW + c] = v[5 * h + r];
	}
	__syncthreads();

	/* pass 1 (radix 10, Ns = 5): butterfly t, k = t mod 5, outputs (t - k) 10 + k + 5 r */
	{
		const int k = t % 5;
		v[0] = buf[t * MCW + c];
#pragma unroll
		for (int r = 1; r < 10; r++)
			v[r] = cmul(buf[(t + 50 * r) * MCW + c], __ldg(D.m_twA + (r - 1) * 5 + k));
		bfly10(v);
		__syncthreads();
		const int base = (t - k) * 10 + k;
#pragma unroll
		for (int r = 0; r < 10; r++)
			buf[(base + 5 * r) * MCW + c] = v[r];
		__syncthreads();
	}

	/* pass 2 (radix 10, Ns = 50): k = t, outputs k1 = t + 50 r */
	v[0] = buf[t * MCW + c];
#pragma unroll
	for (int r = 1; r < 10; r++)
		v[r] = cmul(buf[(t + 50 * r) * MCW + c], __ldg(D.m_twB + (r - 1) * 50 + t));
	bfly10(v);
	/* times W_N^(j2 k1) = W_N^(j2 t) (W_N^(50 j2))^r */
	double2 tw = __ldg(D.m_e1 + t * MN2 + j2);
	const double2 step = __ldg(D.m_e2 + j2);
#pragma unroll
	for (int r = 0; r < 10; r++) {
		Ys[(t + 50 * r) * MN2 + j2] = cmul(v[r], tw);
		tw = cmul(tw, step);
	}
}

/* ------------------------------------------------------------------------------------------ */
#ifndef M_ROWS_MINCTA
#define M_ROWS_MINCTA 3		/* measured: 2 CTAs (122 regs) 6.58 ms, 3 (80) 6.31, 4 (72) 6.82 per 1024 streams */
#endif
__global__ void __launch_bounds__(200, M_ROWS_MINCTA)
dft_rows_1e6(DevParams P, DftPlan D, const double2 *__restrict__ Y, long long *__restrict__ st, int64_t stream0)
{
	extern __shared__ double2 buf[];			/* [2][MROW] */
	__shared__ unsigned int cnt_s;
	const double2 *__restrict__ Ys = Y + (int64_t) blockIdx.x * ((int64_t) MN1 * MN2);
	const int ra = blockIdx.y, rb = (MN1 - ra) % MN1;	/* row pair, ra = 0 .. 250 */
	const bool self = (ra == rb);
	const int rho = threadIdx.x / 100, t = threadIdx.x % 100;
	const bool active = (rho == 0) || !self;
	double2 *row = buf + rho * MROW;
	if (threadIdx.x == 0)
		cnt_s = 0;
	double2 v[10];

	/* pass 0 (Ns = 1): inputs t + 100 r, outputs 10 t + r */
	if (active) {
		const double2 *__restrict__ src = Ys + (rho ? rb : ra) * MN2;
#pragma unroll
		for (int r = 0; r < 10; r++)
			v[r] = src[t + 100 * r];
		bfly10(v);
#pragma unroll
		for (int r = 0; r < 10; r++)
			row[pad10(10 * t + r)] = v[r];
	}
	__syncthreads();
	/* pass 1 (Ns = 10): k = t mod 10, outputs 10 (t - k) + k + 10 r */
	{
		const int k = t % 10;
		if (active) {
			v[0] = row[pad10(t)];
#pragma unroll
			for (int r = 1; r < 10; r++)
				v[r] = cmul(row[pad10(t + 100 * r)], __ldg(D.m_twC + (r - 1) * 10 + k));
			bfly10(v);
		}
		__syncthreads();
		if (active) {
			const int base = 10 * (t - k) + k;
#pragma unroll
			for (int r = 0; r < 10; r++)
				row[pad10(base + 10 * r)] = v[r];
		}
		__syncthreads();
	}
	/* pass 2 (Ns = 100): butterfly j reads j + 100 r and yields Z[k2 = j + 100 r] */
	const double T2x4 = 4.0 * D.T * D.T;
	const double2 w1 = __ldg(D.m_p1 + ra);			/* exp(-2 pi i ra / n) */
	unsigned int cnt = 0;
	if (!self) {
		if (threadIdx.x < 100) {
			const int ja = threadIdx.x, jb = 99 - threadIdx.x;
			double2 za[10], zb[10];
			za[0] = buf[pad10(ja)];
			zb[0] = buf[MROW + pad10(jb)];
#pragma unroll
			for (int r = 1; r < 10; r++) {
				za[r] = cmul(buf[pad10(ja + 100 * r)], __ldg(D.m_twD + (r - 1) * 100 + ja));
				zb[r] = cmul(buf[MROW + pad10(jb + 100 * r)], __ldg(D.m_twD + (r - 1) * 100 + jb));
			}
			bfly10(za);
			bfly10(zb);
			/* bin q = ra + 500 (ja + 100 r); its partner N - q = rb + 500 (jb + 100 (9 - r)) */
			const double2 wq0 = cmul(w1, __ldg(D.m_p2 + ja));
#pragma unroll
			for (int r = 0; r < 10; r++)
				cnt += post_pair(za[r], zb[9 - r], cmul_w20(wq0, r), T2x4, true);
		}
	} else {
		double2 za[10];
		if (threadIdx.x < 100) {
			const int ja = threadIdx.x;
			za[0] = buf[pad10(ja)];
#pragma unroll
			for (int r = 1; r < 10; r++)
				za[r] = cmul(buf[pad10(ja + 100 * r)], __ldg(D.m_twD + (r - 1) * 100 + ja));
			bfly10(za);
		}
		__syncthreads();
		if (threadIdx.x < 100) {
#pragma unroll
			for (int r = 0; r < 10; r++)
				buf[pad10(threadIdx.x + 100 * r)] = za[r];
		}
		__syncthreads();
		/* row 0: N - q = 500 (1000 - k2); row 250: N - q = 250 + 500 (999 - k2); each unordered pair once */
		if (threadIdx.x < 100) {
#pragma unroll
			for (int r = 0; r < 10; r++) {
				const int k2 = threadIdx.x + 100 * r;
				const int k2p = (ra == 0) ? ((MN2 - k2) % MN2) : (MN2 - 1 - k2);
				if (k2 > k2p)
					continue;
				const double2 wq = cmul(w1, __ldg(D.m_p2 + k2));
				cnt += post_pair(buf[pad10(k2)], buf[pad10(k2p)], wq, T2x4, k2 != k2p);
			}
		}
	}
	/*
	 * 200 threads = six warps and a quarter: the last warp has 8 lanes, and a full-mask shuffle reduction would
	 * read lanes that do not exist (undefined values in CUDA - found by the CPU emulation of tests/emu, which
	 * poisons them; the hardware happened to deliver zeros).  The reduction names exactly the lanes the warp has.
	 */
	{
		const unsigned int left = blockDim.x - (threadIdx.x & ~31u);		/* threads from this warp's lane 0 on */
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

cudaError_t launch_dft_1e6(const DevParams &P, const LaunchCtx &L, const DftPlan &D, double2 *d_work, int64_t work_streams)
{
	const size_t smem1 = (size_t) MN1 * MCW * sizeof(double2);
	const size_t smem2 = (size_t) 2 * MROW * sizeof(double2);
	cudaError_t e;
	e = cudaFuncSetAttribute(dft_cols_1e6, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem1);
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(dft_rows_1e6, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem2);
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
		double2 *wk = d_work + (size_t) which * work_streams * ((size_t) MN1 * MN2);
		dim3 g1((unsigned) cs, MN2 / MCW);
		dft_cols_1e6<<<g1, 400, smem1, st>>>(P, D, L.d_in, wk, s0);
		dim3 g2((unsigned) cs, MN1 / 2 + 1);
		dft_rows_1e6<<<g2, 200, smem2, st>>>(P, D, wk, L.d_st, s0);
	}
	if (two) {
		e = cudaEventRecord(L.ev2_join, L.stream2);
		if (e != cudaSuccess) return e;
		e = cudaStreamWaitEvent(L.stream, L.ev2_join, 0);
		if (e != cudaSuccess) return e;
	}
	return cudaGetLastError();
}
