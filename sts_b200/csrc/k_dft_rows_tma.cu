/*
 * k_dft_rows_tma.cu - the row pass of the n = 2^20 FFT (see k_dft_fast.cu: N = 2^19 = 256 x 2048, rows a and 256 - a
 * together, real-FFT post pass and |X| < T count on registers) as a PERSISTENT kernel fed by the TMA unit.
 *
 * dft_rows16 (k_dft_fast.cu) starts every CTA with 16 global loads per thread into registers: with two 256-thread CTAs
 * per SM nothing hides their latency (24 % warps active, FP64 pipe 56 %, "long scoreboard" the first stall reason).
 * Here one CTA per SM walks over its row pairs, and the NEXT pair's 64 KiB are copied global -> shared by bulk
 * asynchronous copies (cp.async.bulk, SASS UBLKCP; completion counted in bytes on an mbarrier) while the three passes
 * of the current pair run: no load ever waits in a register, and no thread issues a load instruction at all.
 *
 * Shared memory: two pair buffers of 2 x 2304 double2 (72 KiB each).  A row is stored with one spare element per 8
 * (position e + e/8), which makes every access of the three passes conflict free for 16-byte elements; the bulk
 * copies deliver it in 128-byte pieces (8 elements) to their padded places, two pieces per thread and pair.
 *
 * The FFT runs IN PLACE (decimation in frequency, radix 16 x 16 x 8), so that no second buffer is needed per pair:
 *   e = 128 a0 + 8 a1 + a2,  k2 = c0 + 16 c1 + 256 c2:
 *   pass 0  thread (a1, a2): 16 points at stride 128, output c0 replaces input a0
 *   pass 1  thread (c0, a2): times W_256^(a1 c0), 16 points at stride 8, output c1 replaces a1
 *   pass 2  thread u = 16 c0 + c1: times W_2048^(a2 (c0 + 16 c1)), 8 contiguous points -> Z[j' + 256 c2], j' = c0 + 16 c1.
 * Butterfly u of row a and butterfly 255 - u of row b (j'_b = 255 - j'_a) hold each other's partners of the real-FFT
 * post pass, exactly as in dft_rows16, so thread u finishes both on registers.
 */
#include "common.cuh"
#include "dft.h"
#include "dft_math.cuh"
#include "dft16.cuh"
#include "tma.cuh"

#define TROWP (QN2 + QN2 / 8)			/* padded row: 2304 elements */
#define T_THREADS 256
#define T_PAIR_BYTES (2 * QN2 * (int) sizeof(double2))	/* 65536 bytes arrive per pair */

__device__ __forceinline__ int pad8(int e) { return e + (e >> 3); }

/* queue the copies of one row pair: thread tid brings piece tid (8 elements) of row a and of row b */
__device__ __forceinline__ void
issue_pair(const double2 *__restrict__ Ys, int ra, double2 *buf, uint64_t *bar)
{
	const int rb = (QN1 - ra) & (QN1 - 1);
	const int g = threadIdx.x;
	mbar_arrive_expect_tx(bar, 2 * 8 * (uint32_t) sizeof(double2));
	bulk_copy_g2s(buf + 9 * g, Ys + (int64_t) ra * QN2 + 8 * g, 8 * (uint32_t) sizeof(double2), bar);
	bulk_copy_g2s(buf + TROWP + 9 * g, Ys + (int64_t) rb * QN2 + 8 * g, 8 * (uint32_t) sizeof(double2), bar);
}

/* the three passes and the count for the pair in `buf` (rows ra and 256 - ra); returns this thread's count */
__device__ __forceinline__ unsigned int
process_pair(const DftPlan &D, int ra, double2 *buf)
{
	const int rb = (QN1 - ra) & (QN1 - 1);
	const bool self = (ra == rb);
	const int rho = threadIdx.x >> 7, t = threadIdx.x & 127;
	const bool active = (rho == 0) || !self;
	double2 *row = buf + rho * TROWP;
	double2 v[16];

	/* pass 0: points t + 128 a0, in place */
	if (active) {
#pragma unroll
		for (int r = 0; r < 16; r++)
			v[r] = row[pad8(t + 128 * r)];
		bfly16(v);
#pragma unroll
		for (int r = 0; r < 16; r++)
			row[pad8(t + 128 * r)] = v[r];
	}
	__syncthreads();
	/* pass 1: thread (c0, a2) = (t / 8, t % 8): points 128 c0 + a2 + 8 a1, times W_256^(a1 c0), in place */
	if (active) {
		const int c0 = t >> 3, base = 128 * c0 + (t & 7);
		double2 tw[16];
		tw[1] = __ldg(D.q_twA + c0); tw[2] = __ldg(D.q_twA + 16 + c0); tw[4] = __ldg(D.q_twA + 3 * 16 + c0);
		tw[8] = __ldg(D.q_twA + 7 * 16 + c0);
		tw[3] = cmul(tw[1], tw[2]); tw[5] = cmul(tw[4], tw[1]); tw[6] = cmul(tw[4], tw[2]); tw[7] = cmul(tw[4], tw[3]);
#pragma unroll
		for (int r = 9; r < 16; r++)
			tw[r] = cmul(tw[8], tw[r - 8]);
		v[0] = row[pad8(base)];
#pragma unroll
		for (int r = 1; r < 16; r++)
			v[r] = cmul(row[pad8(base + 8 * r)], tw[r]);
		bfly16(v);
#pragma unroll
		for (int r = 0; r < 16; r++)
			row[pad8(base + 8 * r)] = v[r];
	}
	__syncthreads();
	/* pass 2: butterfly u holds points 8 u .. 8 u + 7 (padded: 9 u ..) and yields Z[j' + 256 c2], j' = u / 16 + 16 (u % 16) */
	const double T2x4 = 4.0 * D.T * D.T;
	const double2 w1 = __ldg(D.q_p1 + ra);			/* exp(-2 pi i ra / n) */
	const int ua = threadIdx.x, ja = (ua >> 4) + 16 * (ua & 15);
	unsigned int cnt = 0;
	if (!self) {
		const int ub = 255 - ua, jb = 255 - ja;
		double2 za[8], zb[8];
		{	/* twiddles W_2048^(j' r): r = 1, 2, 4 from the table, the rest by products */
			double2 wa[8], wb[8];
			wa[1] = __ldg(D.q_twB + ja); wa[2] = __ldg(D.q_twB + 256 + ja); wa[4] = __ldg(D.q_twB + 3 * 256 + ja);
			wb[1] = __ldg(D.q_twB + jb); wb[2] = __ldg(D.q_twB + 256 + jb); wb[4] = __ldg(D.q_twB + 3 * 256 + jb);
			wa[3] = cmul(wa[1], wa[2]); wa[5] = cmul(wa[4], wa[1]); wa[6] = cmul(wa[4], wa[2]); wa[7] = cmul(wa[4], wa[3]);
			wb[3] = cmul(wb[1], wb[2]); wb[5] = cmul(wb[4], wb[1]); wb[6] = cmul(wb[4], wb[2]); wb[7] = cmul(wb[4], wb[3]);
			za[0] = buf[9 * ua];
			zb[0] = buf[TROWP + 9 * ub];
#pragma unroll
			for (int r = 1; r < 8; r++) {
				za[r] = cmul(buf[9 * ua + r], wa[r]);
				zb[r] = cmul(buf[TROWP + 9 * ub + r], wb[r]);
			}
		}
		bfly<8>(za);
		bfly<8>(zb);
		/* bin q = ra + 256 (ja + 256 c2); its partner N - q = rb + 256 (jb + 256 (7 - c2)); W_n^q = W_n^(ra + 256 ja) W_16^c2 */
		const double2 wq0 = cmul(w1, __ldg(D.q_p2 + ja));
#pragma unroll
		for (int r = 0; r < 8; r++)
			cnt += post_pair(za[r], zb[7 - r], cmul_w16(wq0, r), T2x4, true);
	} else {
		double2 za[8];
		za[0] = buf[9 * ua];
#pragma unroll
		for (int r = 1; r < 8; r++)
			za[r] = cmul(buf[9 * ua + r], __ldg(D.q_twB + (r - 1) * 256 + ja));
		bfly<8>(za);
		__syncthreads();
#pragma unroll
		for (int r = 0; r < 8; r++)
			buf[9 * ua + r] = za[r];		/* Z[ja + 256 r] at the place of point 8 ua + r */
		__syncthreads();
		/* row 0: N - q = 256 (2048 - k2); row 128: N - q = 128 + 256 (2047 - k2); each unordered pair once */
#pragma unroll
		for (int r = 0; r < 8; r++) {
			const int k2 = ja + 256 * r;
			const int k2p = (ra == 0) ? ((QN2 - k2) & (QN2 - 1)) : (QN2 - 1 - k2);
			if (k2 > k2p)
				continue;
			const int jp = k2p & 255, up = 16 * (jp & 15) + (jp >> 4);	/* the butterfly that produced Z[k2p] */
			const double2 wq = cmul(w1, __ldg(D.q_p2 + k2));
			cnt += post_pair(za[r], buf[9 * up + (k2p >> 8)], wq, T2x4, k2 != k2p);
		}
	}
	return cnt;
}

/*
 * items = cs * 129 row pairs of the cs streams whose intermediate lies in Y; CTA b takes items b, b + grid, ...
 * (every item costs the same, a static deal is balanced).
 */
__global__ void __launch_bounds__(T_THREADS, 2)	/* at most 128 registers: the SM's other half stays free for the column pass */
dft_rows16_tma(DevParams P, DftPlan D, const double2 *__restrict__ Y, long long *__restrict__ st, int64_t stream0, int items)
{
	extern __shared__ __align__(128) double2 tbuf[];	/* [2 buffers][2 rows][TROWP] */
	__shared__ __align__(8) uint64_t full[2];		/* "the pair in buffer i has arrived" */
	__shared__ unsigned int cnt_s[2];
	const int pairs = QN1 / 2 + 1;
	if (threadIdx.x == 0) {
		mbar_init(&full[0], T_THREADS);
		mbar_init(&full[1], T_THREADS);
		mbar_fence_init();
		cnt_s[0] = cnt_s[1] = 0;
	}
	__syncthreads();
	int it = blockIdx.x;
	if (it < items)
		issue_pair(Y + (int64_t) (it / pairs) * ((int64_t) QN1 * QN2), it % pairs, tbuf, &full[0]);
	for (int k = 0; it < items; it += gridDim.x, k++) {
		const int cur = k & 1, nxt = it + (int) gridDim.x;
		double2 *buf = tbuf + (size_t) cur * 2 * TROWP;
		if (nxt < items)	/* the other buffer was released by the barrier that ended the previous iteration */
			issue_pair(Y + (int64_t) (nxt / pairs) * ((int64_t) QN1 * QN2), nxt % pairs, tbuf + (size_t) (cur ^ 1) * 2 * TROWP, &full[cur ^ 1]);
		mbar_wait(&full[cur], (uint32_t) ((k >> 1) & 1));
		unsigned int cnt = process_pair(D, it % pairs, buf);
		cnt = (unsigned int) warp_sum((int) cnt);
		if ((threadIdx.x & 31) == 0 && cnt)
			atomicAdd(&cnt_s[cur], cnt);
		fence_proxy_async();		/* this thread's reads and writes of `buf` are ordered before the copies that refill it */
		__syncthreads();
		if (threadIdx.x == 0) {
			if (cnt_s[cur])
				atomicAdd((unsigned long long *) (st + (stream0 + it / pairs) * P.nst + P.st_off[STSGPU_TEST_DFT]),
					  (unsigned long long) cnt_s[cur]);
			cnt_s[cur] = 0;		/* next used two iterations from now, after another barrier */
		}
	}
}

cudaError_t launch_dft_rows_tma(const DevParams &P, const DftPlan &D, const double2 *d_work, long long *d_st, int64_t stream0, int64_t cs,
				int num_sms, cudaStream_t stream)
{
	const size_t smem = (size_t) 2 * 2 * TROWP * sizeof(double2);
	cudaError_t e = cudaFuncSetAttribute(dft_rows16_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	if (e != cudaSuccess) return e;
	const int64_t items = cs * (QN1 / 2 + 1);
	const int64_t grid = items < num_sms ? items : num_sms;
	dft_rows16_tma<<<(unsigned) grid, T_THREADS, smem, stream>>>(P, D, d_work, d_st, stream0, (int) items);
	return cudaGetLastError();
}
