/*
 * k_scan.cu - the prefix-sum family in one pass structure: Frequency (frequency.c:191-201),
 * CumulativeSums (cusum.c:221-236), Runs (runs.c:213-240), RandomExcursions
 * (randomExcursions.c:318-464) and RandomExcursionsVariant (randomExcursionsVariant.c:248-306).
 *
 * One CTA per bitstream.  Each of SCAN_WARPS warps owns a contiguous range of 1024-bit rows
 * (lane l holds word l of the row, so global loads are 128-byte coalesced).
 *   pass 1: popcount of the warp's range -> exclusive sum over warps gives the random walk value
 *           S at the start of every range;
 *   pass 2: per row, a warp shuffle scan gives S before every word; per word a byte LUT gives
 *           the max/min prefix (cusum), one XOR + popcount gives the run transitions, and only the
 *           words with |S| <= 41 (those that can touch the states -9..9 of the excursion tests)
 *           are walked bit by bit.
 * Excursion cycles are merged across words, rows and warps with an associative summary
 * (visits before the first zero, visits after the last zero), visit counts saturating at 5 in
 * 4-bit fields (only min(count, 5) is used, randomExcursions.c:441-451).
 *
 * Algorithmic bytes: n/8 per stream (the second pass re-reads the stream from L1/L2).
 */
#include "common.cuh"

#define SCAN_WARPS 8
#define SCAN_THREADS (SCAN_WARPS * 32)

/* saturating (at 5) add of two words of eight 4-bit counters */
__device__ __forceinline__ uint32_t nib_satadd5(uint32_t a, uint32_t b)
{
	uint32_t x = a + b;		/* nibbles <= 10, no carries */
	uint32_t e = x & 0x0F0F0F0Fu, o = (x >> 4) & 0x0F0F0F0Fu;
	uint32_t me = ((e + 0x7A7A7A7Au) & 0x80808080u) >> 7;	/* byte >= 6 */
	uint32_t mo = ((o + 0x7A7A7A7Au) & 0x80808080u) >> 7;
	me = (me << 8) - me;		/* 0xFF in the bytes that overflowed */
	mo = (mo << 8) - mo;
	e = (e & ~me) | (0x05050505u & me);
	o = (o & ~mo) | (0x05050505u & mo);
	return e | (o << 4);
}

/* a closed cycle with visit counts c: v[k][x]++ for k = min(count, 5) >= 1 (k = 0 is derived from J) */
__device__ __forceinline__ void record_cycle(uint32_t c, unsigned int *v_s)
{
#pragma unroll
	for (int x = 0; x < 8; x++) {
		uint32_t k = (c >> (4 * x)) & 15u;
		if (k)
			atomicAdd(&v_s[k * 8 + x], 1u);
	}
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(DevParams P, const uint32_t *__restrict__ in, long long *__restrict__ st, int64_t nstreams)
{
	__shared__ uint32_t lut[256];
	__shared__ int seg_ones[SCAN_WARPS];
	__shared__ long long seg_bits[SCAN_WARPS];
	__shared__ int seg_max[SCAN_WARPS], seg_min[SCAN_WARPS], seg_trans[SCAN_WARPS];
	__shared__ int seg_nz[SCAN_WARPS];
	__shared__ uint32_t seg_pre[SCAN_WARPS], seg_suf[SCAN_WARPS];
	__shared__ unsigned int v_s[6 * 8];
	__shared__ unsigned int xi_s[18][SCAN_THREADS];

	const int64_t s = blockIdx.x;
	if (s >= nstreams)
		return;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t *__restrict__ w = in + s * P.pitch_words;
	const int64_t n = P.n, nwords = P.nwords;
	const bool want_exc = (P.enabled & ((1u << STSGPU_TEST_RND_EXCURSION) | (1u << STSGPU_TEST_RND_EXCURSION_VAR))) != 0;

	/* byte LUT for the walk: (sum + 8) | (max prefix + 8) << 8 | (min prefix + 8) << 16 */
	{
		int sum = 0, mx = -9, mn = 9;
		for (int b = 7; b >= 0; b--) {
			sum += ((tid >> b) & 1) ? 1 : -1;
			mx = max(mx, sum);
			mn = min(mn, sum);
		}
		lut[tid] = (uint32_t) (sum + 8) | ((uint32_t) (mx + 8) << 8) | ((uint32_t) (mn + 8) << 16);
	}
	if (tid < 48)
		v_s[tid] = 0;
#pragma unroll
	for (int i = 0; i < 18; i++)
		xi_s[i][tid] = 0;

	const int64_t nrows = (nwords + 31) >> 5;
	const int64_t rpw = (nrows + SCAN_WARPS - 1) / SCAN_WARPS;
	const int64_t r0 = min(nrows, warp * rpw), r1 = min(nrows, r0 + rpw);

	/* ---- pass 1: ones in my range ---- */
	int ones = 0;
	for (int64_t r = r0; r < r1; r++) {
		int64_t idx = r * 32 + lane;
		if (idx < nwords)
			ones += __popc(__ldg(w + idx));
	}
	ones = warp_sum(ones);
	if (lane == 0) {
		seg_ones[warp] = ones;
		seg_bits[warp] = min(n, r1 * 1024) - min(n, r0 * 1024);
	}
	__syncthreads();
	int S = 0;		/* walk value before my range */
	int total_ones = 0;
	for (int j = 0; j < SCAN_WARPS; j++) {
		if (j < warp)
			S += 2 * seg_ones[j] - (int) seg_bits[j];
		total_ones += seg_ones[j];
	}

	/* ---- pass 2 ---- */
	int smax = 0, smin = 0, trans = 0, nzeros = 0;
	uint32_t open = 0;		/* visits of the currently open cycle (warp uniform) */
	uint32_t first_pre = 0;		/* visits before the first zero of my range */
	int range_has_zero = 0;
	uint32_t carry = 0;		/* last bit of the previous word row */
	if (r0 > 0 && r0 < nrows)
		carry = ldw(w, r0 * 32 - 1) & 1u;

	for (int64_t r = r0; r < r1; r++) {
		const int64_t idx = r * 32 + lane;
		const int64_t rem = n - idx * 32;
		const int vb = rem >= 32 ? 32 : (rem > 0 ? (int) rem : 0);
		const uint32_t wd = (idx < nwords) ? ldw(w, idx) : 0u;
		const int d = 2 * __popc(wd) - vb;
		const int incl = warp_scan_incl(d, lane);
		const int Sb = S + incl - d;

		/* cusum: max / min prefix inside the word */
		if (vb == 32) {
			int run = 0, mx = -64, mn = 64;
#pragma unroll
			for (int k = 3; k >= 0; k--) {
				uint32_t e = lut[(wd >> (8 * k)) & 255u];
				mx = max(mx, run + (int) ((e >> 8) & 255u) - 8);
				mn = min(mn, run + (int) ((e >> 16) & 255u) - 8);
				run += (int) (e & 255u) - 8;
			}
			smax = max(smax, Sb + mx);
			smin = min(smin, Sb + mn);
		} else if (vb > 0) {
			int run = Sb;
			for (int p = 0; p < vb; p++) {
				run += ((wd >> (31 - p)) & 1u) ? 1 : -1;
				smax = max(smax, run);
				smin = min(smin, run);
			}
		}

		/* runs: transitions between consecutive bits */
		uint32_t prev = __shfl_up_sync(0xffffffffu, wd & 1u, 1);
		if (lane == 0)
			prev = carry;
		uint32_t x = wd ^ ((wd >> 1) | (prev << 31));
		uint32_t m = mask_below((uint32_t) vb);
		if (idx == 0)
			m &= 0x7FFFFFFFu;
		trans += __popc(x & m);
		carry = __shfl_sync(0xffffffffu, wd & 1u, 31);

		/* excursions: only words that can reach |S| <= 9 */
		if (want_exc) {
			const bool hot = vb > 0 && abs(Sb) <= 41;
			if (__any_sync(0xffffffffu, hot)) {
				int nz = 0;
				uint32_t pre = 0, cur = 0;
				if (hot) {
					int Sx = Sb;
					for (int p = 0; p < vb; p++) {
						Sx += ((wd >> (31 - p)) & 1u) ? 1 : -1;
						if (Sx == 0) {
							if (nz == 0)
								pre = cur;
							else
								record_cycle(cur, v_s);
							nz++;
							cur = 0;
						} else {
							int a = abs(Sx);
							if (a <= 9) {
								xi_s[Sx < 0 ? Sx + 9 : Sx + 8][tid]++;
								if (a <= 4) {
									int sh = 4 * (Sx < 0 ? Sx + 4 : Sx + 3);
									if (((cur >> sh) & 15u) < 5u)
										cur += 1u << sh;
								}
							}
						}
					}
					if (nz == 0)
						pre = cur;	/* no zero: "pre" carries the whole word */
				}
				nzeros += nz;
				/* ordered merge over the lanes (warp uniform) */
				for (int l = 0; l < 32; l++) {
					int nz_l = __shfl_sync(0xffffffffu, nz, l);
					uint32_t pre_l = __shfl_sync(0xffffffffu, pre, l);
					uint32_t suf_l = __shfl_sync(0xffffffffu, cur, l);
					if (nz_l == 0) {
						open = nib_satadd5(open, pre_l);
					} else {
						uint32_t closed = nib_satadd5(open, pre_l);
						if (!range_has_zero) {
							range_has_zero = 1;
							first_pre = closed;
						} else if (lane == 0) {
							record_cycle(closed, v_s);
						}
						open = suf_l;
					}
				}
			}
		}
		S += __shfl_sync(0xffffffffu, incl, 31);
	}

	smax = warp_max(smax);
	smin = warp_min(smin);
	trans = warp_sum(trans);
	nzeros = warp_sum(nzeros);
	if (lane == 0) {
		seg_max[warp] = smax;
		seg_min[warp] = smin;
		seg_trans[warp] = trans;
		seg_nz[warp] = range_has_zero ? nzeros : 0;
		seg_pre[warp] = range_has_zero ? first_pre : open;
		seg_suf[warp] = open;
	}
	__syncthreads();

	/* ---- combine the ranges in order ---- */
	if (tid == 0) {
		int gmax = 0, gmin = 0, gtrans = 0;
		long long J = 0;
		uint32_t op = 0;
		for (int j = 0; j < SCAN_WARPS; j++) {
			gmax = max(gmax, seg_max[j]);
			gmin = min(gmin, seg_min[j]);
			gtrans += seg_trans[j];
			if (seg_nz[j] == 0) {
				op = nib_satadd5(op, seg_pre[j]);
			} else {
				record_cycle(nib_satadd5(op, seg_pre[j]), v_s);
				op = seg_suf[j];
				J += seg_nz[j];
			}
		}
		const long long Sfin = 2LL * total_ones - n;
		if (Sfin != 0) {	/* randomExcursions.c:345-347: the trailing partial cycle counts */
			record_cycle(op, v_s);
			J++;
		}
		long long *o = st + s * P.nst;
		if (P.enabled & (1u << STSGPU_TEST_FREQUENCY))
			o[P.st_off[STSGPU_TEST_FREQUENCY]] = Sfin;
		if (P.enabled & (1u << STSGPU_TEST_CUSUM)) {
			long long *c = o + P.st_off[STSGPU_TEST_CUSUM];
			c[0] = Sfin;
			c[1] = gmax;
			c[2] = gmin;
			c[3] = (gmax > -gmin) ? gmax : -gmin;				/* z_forward, cusum.c:236 */
			c[4] = (gmax - Sfin > Sfin - gmin) ? gmax - Sfin : Sfin - gmin;	/* z_backward, cusum.c:235 */
		}
		if (P.enabled & (1u << STSGPU_TEST_RUNS)) {
			long long *c = o + P.st_off[STSGPU_TEST_RUNS];
			c[0] = total_ones;
			c[1] = 1 + gtrans;
			/* c[2] (test_possible) is decided by the tail kernel in the reference's FP expression */
		}
		if (P.enabled & (1u << STSGPU_TEST_RND_EXCURSION)) {
			long long *c = o + P.st_off[STSGPU_TEST_RND_EXCURSION];
			c[0] = J;
			for (int x = 0; x < 8; x++) {
				long long rest = 0;
				for (int k = 1; k < 6; k++) {
					c[1 + 8 * k + x] = v_s[k * 8 + x];
					rest += v_s[k * 8 + x];
				}
				c[1 + x] = J - rest;
			}
		}
		if (P.enabled & (1u << STSGPU_TEST_RND_EXCURSION_VAR))
			o[P.st_off[STSGPU_TEST_RND_EXCURSION_VAR]] = J;
	}
	__syncthreads();	/* v_s atomics of thread 0 above are complete; xi_s is complete since the earlier barrier */
	if ((P.enabled & (1u << STSGPU_TEST_RND_EXCURSION_VAR)) && warp < 18 / 3 + 0) {
		/* 18 states, 3 per warp-iteration: warp j reduces states 3j..3j+2 */
		for (int q = 0; q < 3; q++) {
			int state = warp * 3 + q;
			long long acc = 0;
			for (int t = lane; t < SCAN_THREADS; t += 32)
				acc += xi_s[state][t];
			acc = warp_sum_ll(acc);
			if (lane == 0)
				st[s * P.nst + P.st_off[STSGPU_TEST_RND_EXCURSION_VAR] + 1 + state] = acc;
		}
	}
}

cudaError_t launch_scan(const DevParams &P, const LaunchCtx &L)
{
	const uint32_t need = (1u << STSGPU_TEST_FREQUENCY) | (1u << STSGPU_TEST_CUSUM) | (1u << STSGPU_TEST_RUNS) |
			      (1u << STSGPU_TEST_RND_EXCURSION) | (1u << STSGPU_TEST_RND_EXCURSION_VAR);
	if (!(P.enabled & need))
		return cudaSuccess;
	scan_kernel<<<(unsigned) L.nstreams, SCAN_THREADS, 0, L.stream>>>(P, L.d_in, L.d_st, L.nstreams);
	return cudaGetLastError();
}
