/*
 * k_scan.cu - the prefix-sum family in one pass structure: Frequency (frequency.c:191-201),
 * CumulativeSums (cusum.c:221-236), Runs (runs.c:213-240), RandomExcursions
 * (randomExcursions.c:318-464) and RandomExcursionsVariant (randomExcursionsVariant.c:248-306).
 *
 * One CTA per bitstream, the stream taken in chunks of SCAN_ROWS (256) rows of 1024 bits (lane l of a warp holds
 * word l of a row, so global loads are 128-byte coalesced):
 *   pass 1: popcount of every row -> block-wide exclusive scan = the random walk value S at the
 *           start of every row (and the last bit of every row for the run count);
 *   pass 2: rows are dealt to the warps ROUND ROBIN.  Per row a warp shuffle scan gives S before
 *           every word; per word a byte LUT gives the max/min prefix (cusum), one XOR + popcount the
 *           run transitions, and only the words with |S| <= 41 (those that can touch the states
 *           -9..9 of the excursion tests) are walked bit by bit.  The walk returns to zero in
 *           clusters, so contiguous row ranges per warp (the first version) left one warp with all
 *           the bit-by-bit rows while seven waited at the barrier; dealt round robin they spread.
 *   pass 3: excursion cycles cross word, row and chunk boundaries.  Every row leaves an associative
 *           summary (zeros inside, visits before the first zero, visits after the last zero; cycles
 *           strictly inside a row are recorded at once); the summaries are merged in row order by
 *           256 threads -> 8 warps -> 1 thread.  Visit counts saturate at 5 in 4-bit fields (only
 *           min(count, 5) is used, randomExcursions.c:441-451).
 *
 * Algorithmic bytes: n/8 per stream (the second pass re-reads a 32 KiB chunk right after the first: L1/L2 hits).
 */
#include "common.cuh"

#define SCAN_WARPS 8
#define SCAN_THREADS (SCAN_WARPS * 32)
#define SCAN_ROWS 256			/* rows per chunk: 32 KiB of stream, still in L2 (and mostly in L1) when pass 2 reads it again */
#define SCAN_RPT (SCAN_ROWS / SCAN_THREADS)	/* rows per thread in the block scans */

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

/*
 * Excursion summary of a span of the walk: nz zeros inside; pre = visit counts before the first zero
 * (of the whole span when nz == 0); suf = visit counts after the last zero (unused when nz == 0).
 * append(b) merges the span that follows; the cycle closed at the seam is recorded.
 */
struct Exc {
	uint32_t nz, pre, suf;
	__device__ __forceinline__ void append(const Exc &b, unsigned int *v_s)
	{
		if (b.nz == 0) {
			if (nz == 0)
				pre = nib_satadd5(pre, b.pre);
			else
				suf = nib_satadd5(suf, b.pre);
		} else {
			if (nz == 0)
				pre = nib_satadd5(pre, b.pre);
			else
				record_cycle(nib_satadd5(suf, b.pre), v_s);
			suf = b.suf;
			nz += b.nz;
		}
	}
};

__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(DevParams P, const uint32_t *__restrict__ in, long long *__restrict__ st, int64_t nstreams)
{
	__shared__ uint32_t lut[256];
	__shared__ int row_S[SCAN_ROWS];			/* ones per row, then walk value at the row start */
	__shared__ uint8_t row_last[SCAN_ROWS];			/* last bit of the row */
	__shared__ uint32_t row_nz[SCAN_ROWS], row_pre[SCAN_ROWS], row_suf[SCAN_ROWS];
	__shared__ int warp_tot[SCAN_WARPS];
	__shared__ Exc warp_exc[SCAN_WARPS];
	__shared__ unsigned int v_s[6 * 8];
	__shared__ unsigned int xi_s[18][SCAN_THREADS];
	__shared__ int red_max[SCAN_WARPS], red_min[SCAN_WARPS], red_trans[SCAN_WARPS];
	__shared__ int chunk_hot;				/* some row of this chunk came within the excursion states */

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
	int smax = 0, smin = 0, trans = 0;
	long long total_ones = 0;
	int S0 = 0;			/* walk value at the start of the chunk */
	uint32_t last0 = 0;		/* last bit before the chunk */
	Exc run;			/* thread 0: excursion state of everything before the chunk */
	run.nz = run.pre = run.suf = 0;

	for (int64_t c0 = 0; c0 < nrows; c0 += SCAN_ROWS) {
		const int crow = (int) min((int64_t) SCAN_ROWS, nrows - c0);
		__syncthreads();				/* previous chunk's shared arrays are consumed */
		/*
		 * ---- pass 1: ones and last bit of every row.  Sixteen bytes per thread, a row = 8 consecutive lanes; the words
		 * behind the stream's end are zero padding up to the pitch, so whole rows are read.  The last bit of a row is only
		 * ever used by the row that follows it (a ragged last row has none).
		 */
		{
			const uint4 *__restrict__ w4 = (const uint4 *) (w + c0 * 32);
			for (int q = tid; q < ((crow * 8 + 31) & ~31); q += SCAN_THREADS) {	/* whole warps: the shuffles below are full-mask */
				const uint4 v = (q < crow * 8) ? __ldg(w4 + q) : make_uint4(0u, 0u, 0u, 0u);
				int ones = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
				ones += __shfl_xor_sync(0xffffffffu, ones, 1);
				ones += __shfl_xor_sync(0xffffffffu, ones, 2);
				ones += __shfl_xor_sync(0xffffffffu, ones, 4);
				if ((lane & 7) == 0 && q < crow * 8)
					row_S[q >> 3] = ones;
				if ((lane & 7) == 7 && q < crow * 8)
					row_last[q >> 3] = (uint8_t) (bswap32(v.w) & 1u);
			}
		}
		for (int r = tid; r < SCAN_ROWS; r += SCAN_THREADS) {
			row_nz[r] = 0;
			row_pre[r] = 0;
			row_suf[r] = 0;
		}
		if (tid == 0)
			chunk_hot = 0;
		__syncthreads();
		/* ---- block scan: walk value at every row start (SCAN_RPT consecutive rows per thread) ---- */
		{
			int d[SCAN_RPT], loc = 0;
#pragma unroll
			for (int j = 0; j < SCAN_RPT; j++) {
				const int r = SCAN_RPT * tid + j;
				int v = 0;
				if (r < crow) {
					const int64_t b0 = min(n, (c0 + r) * 1024), b1 = min(n, (c0 + r + 1) * 1024);
					v = 2 * row_S[r] - (int) (b1 - b0);
					total_ones += row_S[r];
				}
				d[j] = v;
				loc += v;
			}
			const int incl = warp_scan_incl(loc, lane);
			if (lane == 31)
				warp_tot[warp] = incl;
			__syncthreads();
			int base = S0 + incl - loc;
			for (int j = 0; j < warp; j++)
				base += warp_tot[j];
#pragma unroll
			for (int j = 0; j < SCAN_RPT; j++) {
				const int r = SCAN_RPT * tid + j;
				if (r < crow)
					row_S[r] = base;
				base += d[j];
			}
			int chunk_sum = 0;
			for (int j = 0; j < SCAN_WARPS; j++)
				chunk_sum += warp_tot[j];
			__syncthreads();
			S0 += chunk_sum;			/* every thread keeps the same running value */
		}

		/* ---- pass 2: rows round robin ---- */
		for (int r = warp; r < crow; r += SCAN_WARPS) {
			const int64_t idx = (c0 + r) * 32 + lane;
			const int64_t rem = n - idx * 32;
			const int vb = rem >= 32 ? 32 : (rem > 0 ? (int) rem : 0);
			const uint32_t wd = (idx < nwords) ? ldw(w, idx) : 0u;
			const int d = 2 * __popc(wd) - vb;
			const int incl = warp_scan_incl(d, lane);
			const int Sb = row_S[r] + incl - d;

			/* cusum: max / min prefix inside the word */
			int wmx = 41, wmn = -41;			/* reach of the walk inside this word, relative to Sb */
			/* the byte LUT only where it can matter: a new extreme for this lane, or a word near the excursion states */
			const bool look = Sb + 32 > smax || Sb - 32 < smin || abs(Sb) <= 41;
			if (vb == 32 && !look) {
				/* nothing to record */
			} else if (vb == 32) {
				int run_ = 0, mx = -64, mn = 64;
#pragma unroll
				for (int k = 3; k >= 0; k--) {
					uint32_t e = lut[(wd >> (8 * k)) & 255u];
					mx = max(mx, run_ + (int) ((e >> 8) & 255u) - 8);
					mn = min(mn, run_ + (int) ((e >> 16) & 255u) - 8);
					run_ += (int) (e & 255u) - 8;
				}
				smax = max(smax, Sb + mx);
				smin = min(smin, Sb + mn);
				wmx = mx;
				wmn = mn;
			} else if (vb > 0) {
				int run_ = Sb;
				for (int p = 0; p < vb; p++) {
					run_ += ((wd >> (31 - p)) & 1u) ? 1 : -1;
					smax = max(smax, run_);
					smin = min(smin, run_);
				}
			}

			/* runs: transitions between consecutive bits */
			uint32_t prev = __shfl_up_sync(0xffffffffu, wd & 1u, 1);
			if (lane == 0)
				prev = (r > 0) ? (uint32_t) row_last[r - 1] : last0;
			uint32_t x = wd ^ ((wd >> 1) | (prev << 31));
			uint32_t m = mask_below((uint32_t) vb);
			if (idx == 0)
				m &= 0x7FFFFFFFu;
			trans += __popc(x & m);

			/* excursions: only words whose walk comes within the states -9 .. 9 (the prefix extremes are known from the cusum LUT) */
			if (want_exc) {
				const bool hot = vb > 0 && abs(Sb) <= 41 && Sb + wmn <= 9 && Sb + wmx >= -9;
				if (__any_sync(0xffffffffu, hot)) {
					Exc e;
					e.nz = e.pre = e.suf = 0;
					if (hot) {
						uint32_t cur = 0;
						int Sx = Sb;
						for (int p = 0; p < vb; p++) {
							if ((p & 7) == 0 && p + 8 <= vb) {
								/* a byte that stays outside -9 .. 9 only moves the walk */
								const uint32_t be = lut[(wd >> (24 - p)) & 255u];
								if (Sx + (int) ((be >> 16) & 255u) - 8 > 9 || Sx + (int) ((be >> 8) & 255u) - 8 < -9) {
									Sx += (int) (be & 255u) - 8;
									p += 7;
									continue;
								}
							}
							Sx += ((wd >> (31 - p)) & 1u) ? 1 : -1;
							if (Sx == 0) {
								if (e.nz == 0)
									e.pre = cur;
								else
									record_cycle(cur, v_s);
								e.nz++;
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
						if (e.nz == 0)
							e.pre = cur;	/* no zero: "pre" carries the whole word */
						else
							e.suf = cur;
					}
					/* ordered tree merge over the lanes: lane l absorbs lane l + o */
#pragma unroll
					for (int o = 1; o < 32; o <<= 1) {
						Exc b;
						b.nz = __shfl_down_sync(0xffffffffu, e.nz, o);
						b.pre = __shfl_down_sync(0xffffffffu, e.pre, o);
						b.suf = __shfl_down_sync(0xffffffffu, e.suf, o);
						if ((lane & (2 * o - 1)) == 0)
							e.append(b, v_s);
					}
					if (lane == 0) {
						row_nz[r] = e.nz;
						row_pre[r] = e.pre;
						row_suf[r] = e.suf;
						chunk_hot = 1;
					}
				}
			}
		}
		last0 = (uint32_t) 0;				/* set after the barrier below (row_last of this chunk) */
		__syncthreads();
		last0 = row_last[crow - 1];

		/* ---- pass 3: merge the row summaries in order: SCAN_RPT rows per thread, 32 threads per warp, 8 warps ---- */
		if (want_exc && chunk_hot) {			/* the summary of a chunk without hot rows is the identity */
			Exc e;
			e.nz = e.pre = e.suf = 0;
#pragma unroll
			for (int j = 0; j < SCAN_RPT; j++) {
				const int r = SCAN_RPT * tid + j;
				Exc b;
				b.nz = row_nz[r];
				b.pre = row_pre[r];
				b.suf = row_suf[r];
				e.append(b, v_s);
			}
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				Exc b;
				b.nz = __shfl_down_sync(0xffffffffu, e.nz, o);
				b.pre = __shfl_down_sync(0xffffffffu, e.pre, o);
				b.suf = __shfl_down_sync(0xffffffffu, e.suf, o);
				if ((lane & (2 * o - 1)) == 0)
					e.append(b, v_s);
			}
			if (lane == 0)
				warp_exc[warp] = e;
			__syncthreads();
			if (tid == 0)
				for (int j = 0; j < SCAN_WARPS; j++)
					run.append(warp_exc[j], v_s);
		}
	}

	smax = warp_max(smax);
	smin = warp_min(smin);
	trans = warp_sum(trans);
	total_ones = warp_sum_ll(total_ones);
	if (lane == 0) {
		red_max[warp] = smax;
		red_min[warp] = smin;
		red_trans[warp] = trans;
		row_S[warp] = (int) total_ones;			/* <= 2^31 bits per stream */
	}
	__syncthreads();

	if (tid == 0) {
		int gmax = 0, gmin = 0, gtrans = 0;
		long long ones = 0;
		for (int j = 0; j < SCAN_WARPS; j++) {
			gmax = max(gmax, red_max[j]);
			gmin = min(gmin, red_min[j]);
			gtrans += red_trans[j];
			ones += (unsigned int) row_S[j];
		}
		long long J = run.nz;
		const long long Sfin = 2LL * ones - n;
		if (run.nz)		/* the first cycle: from the start of the stream to its first zero */
			record_cycle(run.pre, v_s);
		if (Sfin != 0) {	/* randomExcursions.c:345-347: the trailing partial cycle counts */
			record_cycle(run.nz ? run.suf : run.pre, v_s);
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
			c[0] = ones;
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
	__syncthreads();	/* xi_s is complete since the barriers above */
	if ((P.enabled & (1u << STSGPU_TEST_RND_EXCURSION_VAR)) && warp < 6) {
		/* 18 states, 3 per warp: warp j reduces states 3j..3j+2 */
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

/* ------------------------------------------------------------------------------------------ */
/*
 * -s only (STSGPU_FLAG_STATS): stat.counter[] of RandomExcursions as its stats.txt prints it - the visit counts of
 * the LAST cycle alone, because the reference clears the counters at the start of every cycle
 * (randomExcursions.c:386) and prints what is left after the loop (:860-866); zero when the test was not possible
 * (:581).  With the cycle ends E = { i in [1, n-1] : S_i = 0 } plus n if S_(n-1) != 0 (:337-347), the last cycle is
 * i in [z2, z1) for the two largest elements z1 > z2 of E (z2 = 0 when E has one element, :378), and counter[x]
 * = #{ i in [z2, z1) : S_i = x } for x in -4..-1, 1..4, not saturated.
 *
 * One CTA per stream, thread t owns a contiguous run of words: local sums -> block scan = walk value at the start
 * of every run; one walk to find each thread's last two zeros, a second one to count inside [z2, z1).  Words that
 * cannot come within reach of the levels of interest are skipped by their popcount.  Runs after scan_kernel on
 * the same stream (J is read from the record).
 */
#define LASTC_THREADS 256

__global__ void __launch_bounds__(LASTC_THREADS)
last_cycle_kernel(DevParams P, const uint32_t *__restrict__ in, long long *__restrict__ st, int64_t nstreams)
{
	__shared__ int warp_tot[LASTC_THREADS / 32];
	__shared__ long long zero_a[LASTC_THREADS], zero_b[LASTC_THREADS];
	__shared__ long long range[2];
	__shared__ unsigned long long counter[8];
	const int64_t s = blockIdx.x;
	if (s >= nstreams)
		return;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t *__restrict__ w = in + s * P.pitch_words;
	const int64_t n = P.n, nwords = P.nwords;
	long long *c = st + s * P.nst + P.st_off[STSGPU_TEST_RND_EXCURSION];
	if (tid < 8)
		counter[tid] = 0;
	const int64_t per = (nwords + LASTC_THREADS - 1) / LASTC_THREADS;
	const int64_t w0 = min(nwords, (int64_t) tid * per), w1 = min(nwords, w0 + per);

	int loc = 0;
	for (int64_t i = w0; i < w1; i++) {
		const int64_t rem = n - i * 32;
		loc += 2 * __popc(ldw(w, i)) - (int) (rem >= 32 ? 32 : rem);
	}
	const int incl = warp_scan_incl(loc, lane);
	if (lane == 31)
		warp_tot[warp] = incl;
	__syncthreads();
	int S_run = incl - loc, S_total = 0;
	for (int j = 0; j < LASTC_THREADS / 32; j++) {
		if (j < warp)
			S_run += warp_tot[j];
		S_total += warp_tot[j];
	}

	/* the last two zeros of this thread's run (0 = none: position 0 is never a zero, and it is where the first cycle starts) */
	long long za = 0, zb = 0;
	{
		int S = S_run;
		for (int64_t i = w0; i < w1; i++) {
			const uint32_t wd = ldw(w, i);
			const int64_t rem = n - i * 32;
			const int vb = (int) (rem >= 32 ? 32 : rem);
			if (abs(S) <= vb) {
				int Sx = S;
				for (int p = 0; p < vb; p++) {
					Sx += ((wd >> (31 - p)) & 1u) ? 1 : -1;
					const long long pos = i * 32 + p;
					if (Sx == 0 && pos <= n - 1) {
						zb = za;
						za = pos;
					}
				}
			}
			S += 2 * __popc(wd) - vb;
		}
	}
	zero_a[tid] = za;
	zero_b[tid] = zb;
	__syncthreads();
	if (tid == 0) {
		long long top[2] = { 0, 0 };
		int found = 0;
		for (int t = LASTC_THREADS - 1; t >= 0 && found < 2; t--) {
			if (zero_a[t] > 0 && found < 2) top[found++] = zero_a[t];
			if (zero_b[t] > 0 && found < 2) top[found++] = zero_b[t];
		}
		/* S_(n-1) != 0: the trailing partial cycle [last zero, n); else the cycle that ends at the zero n - 1 */
		if (S_total != 0) {
			range[1] = n;
			range[0] = top[0];
		} else {
			range[1] = top[0];
			range[0] = top[1];
		}
	}
	__syncthreads();
	const long long z2 = range[0], z1 = range[1];
	const bool possible = c[0] >= P.min_zero_crossings;
	if (possible) {
		unsigned int cnt[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
		int S = S_run;
		for (int64_t i = w0; i < w1; i++) {
			const uint32_t wd = ldw(w, i);
			const int64_t rem = n - i * 32;
			const int vb = (int) (rem >= 32 ? 32 : rem);
			const long long b0 = i * 32;
			if (b0 + vb > z2 && b0 < z1 && abs(S) <= vb + 4) {
				int Sx = S;
				for (int p = 0; p < vb; p++) {
					Sx += ((wd >> (31 - p)) & 1u) ? 1 : -1;
					const long long pos = b0 + p;
					if (pos >= z2 && pos < z1 && Sx != 0 && abs(Sx) <= 4)
						cnt[Sx < 0 ? Sx + 4 : Sx + 3]++;
				}
			}
			S += 2 * __popc(wd) - vb;
		}
#pragma unroll
		for (int x = 0; x < 8; x++) {
			const unsigned int v = (unsigned int) warp_sum((int) cnt[x]);
			if (lane == 0 && v)
				atomicAdd(&counter[x], (unsigned long long) v);
		}
	}
	__syncthreads();
	if (tid < 8)
		c[49 + tid] = possible ? (long long) counter[tid] : 0;
}

cudaError_t launch_last_cycle(const DevParams &P, const LaunchCtx &L)
{
	last_cycle_kernel<<<(unsigned) L.nstreams, LASTC_THREADS, 0, L.stream>>>(P, L.d_in, L.d_st, L.nstreams);
	return cudaGetLastError();
}
