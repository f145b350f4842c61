/*
 * k_universal.cu - Maurer's universal statistical test (universal.c:318-375).
 *
 * The statistic is  sum += log(i - T[block_i]) / log(2)  over K = 1000 * 2^L test blocks, a
 * sequential double-precision sum of 128000 terms at L = 7 whose rounding depends on the order.
 * To reproduce the reference's value bit for bit the kernel
 *   - finds every distance i - T[.] exactly (integers): a warp takes 32 consecutive L-bit blocks
 *     (one "row"), `match.any` resolves repeats inside the row and a shared-memory table T[2^L]
 *     carries the last occurrence across rows (table starts at 0 like the reference's memset,
 *     universal.c:327);
 *   - looks the term up in a table of log(d) / log(2.0) computed ON THE HOST with the reference's
 *     expression and libm, so each addend is the reference's double - stored as the exact fixed
 *     point value x * 2^58 (every term is a double in [1, 32) or 0), which is what the integer
 *     emulation of the sum needs;
 *   - reproduces the sequential round-to-nearest-even accumulation EXACTLY but in parallel
 *     (rne_sum.cuh): while the running sum stays in one binade and no addend is an exact tie,
 *     adding the up to 1024 terms of a super row is a plain integer sum of rounded fixed point
 *     values, each lane summing a contiguous span; a super row with a tie or a binade change
 *     (about 10 of 127 at L = 7) goes through composable round-to-even maps, 128 terms at a time,
 *     and at worst term by term with DADD.
 *
 * One warp per stream (one CTA = one warp).  The packed stream is staged through shared memory in
 * "super rows" of 32 L words = 1024 blocks with cp.async, double buffered, so DRAM latency is off
 * the dependency chain; what remains is the T[] chain (LDS -> STS per row).
 *
 * Algorithmic bytes: ceil((Q + K) * L / 8) per stream (113120 at L = 7) + the L1/L2-resident table.
 */
#include "common.cuh"
#include "rne_sum.cuh"

#define UNI_GROUP 4			/* rows per summation group */

__device__ __forceinline__ void cp_async4(uint32_t *smem_dst, const uint32_t *gsrc)
{
	const unsigned d = (unsigned) __cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(32)
universal_kernel(DevParams P, const uint32_t *__restrict__ in, const unsigned long long *__restrict__ log2fix,
		 long long *__restrict__ st)
{
	extern __shared__ __align__(16) uint32_t usm[];
	const int L = P.univ_L;
	const int IB = 32 * L + 2;				/* one super row + pad word, even */
	unsigned long long *tbuf = (unsigned long long *) usm;	/* [32][33] terms of one super row as fixed point, block order */
	uint32_t *ibuf = usm + 2 * 32 * 33;			/* [2][IB] staged input */
	uint32_t *T = ibuf + 2 * IB;				/* [2^L] */
	const int64_t s = blockIdx.x;
	const int lane = threadIdx.x;
	const uint32_t *__restrict__ w = in + s * P.pitch_words;
	for (int i = lane; i < (1 << L); i += 32)
		T[i] = 0;
	if (lane < 2)
		ibuf[lane * IB + 32 * L] = 0;			/* pad words (read, never used) */
	const int64_t Q = P.univ_Q, K = P.univ_K;
	const int rows = (int) ((Q + K) >> 5);			/* 1010 * 2^(L-5), a multiple of UNI_GROUP for L >= 6 */
	const int qrows = (int) (Q >> 5);
	const int nsr = (rows + 31) >> 5;

	auto stage = [&](int sr) {
		uint32_t *dst = ibuf + (sr & 1) * IB;
		const int64_t g0 = (int64_t) sr * 32 * L + lane;
		for (int j = 0; j < L; j++) {
			const int64_t g = g0 + 32 * j;
			if (g < P.pitch_words)
				cp_async4(dst + 32 * j + lane, w + g);
		}
		cp_async_commit();
	};

	double sum = 0.0;
	stage(0);
	for (int sr = 0; sr < nsr; sr++) {
		if (sr + 1 < nsr) {
			stage(sr + 1);
			cp_async_wait<1>();
		} else {
			cp_async_wait<0>();
		}
		__syncwarp();
		const uint32_t *buf = ibuf + (sr & 1) * IB;
		const int row_end = min(32, rows - sr * 32);
		const int first_test = min(row_end, max(0, qrows - sr * 32));	/* first row of this super row that is summed */
		for (int rg = 0; rg < row_end; rg += UNI_GROUP) {
			const int row0 = sr * 32 + rg;
			const bool testing = row0 >= qrows;
			/* phase A: the four rows' patterns and their in-row repeats (independent of T[]) */
			uint32_t v[UNI_GROUP], lower[UNI_GROUP], tprev[UNI_GROUP];
			bool is_last[UNI_GROUP];
#pragma unroll
			for (int r4 = 0; r4 < UNI_GROUP; r4++) {
				const int off = ((rg + r4) * 32 + lane) * L;	/* bit offset inside the super row */
				const uint32_t hi = bswap32(buf[off >> 5]), lo = bswap32(buf[(off >> 5) + 1]);
				v[r4] = __funnelshift_l(lo, hi, off & 31) >> (32 - L);
#ifndef UNI_BALLOT
				const uint32_t same = __match_any_sync(0xffffffffu, v[r4]);
#else
				/* one ballot per pattern bit instead of MATCH.ANY: measured 1.27 against 1.07 ms for the kernel alone
				 * (twice the issue slots) and no difference beyond noise for the pipelined batch */
				uint32_t same = 0xffffffffu;
				for (int b = 0; b < L; b++) {
					const uint32_t bal = __ballot_sync(0xffffffffu, (v[r4] >> b) & 1u);
					same &= ((v[r4] >> b) & 1u) ? bal : ~bal;
				}
#endif
				lower[r4] = same & ((1u << lane) - 1u);
				is_last[r4] = (same >> lane) <= 1u;	/* no higher lane holds the same pattern */
			}
			/* phase B: the T[] chain; the loaded values are only consumed in phase C */
#pragma unroll
			for (int r4 = 0; r4 < UNI_GROUP; r4++) {
				tprev[r4] = T[v[r4]];
				__syncwarp();
				if (is_last[r4])
					T[v[r4]] = (uint32_t) (row0 + r4) * 32u + lane + 1u;	/* 1-based block number */
				__syncwarp();
			}
			/* phase C: distances -> terms, in block order in tbuf (rows of 32, pitch 33) */
			if (testing) {
				unsigned long long term[UNI_GROUP];
#pragma unroll
				for (int r4 = 0; r4 < UNI_GROUP; r4++) {
					const uint32_t i = (uint32_t) (row0 + r4) * 32u + lane + 1u;
					const uint32_t prev = lower[r4] ? (i - lane + (31 - __clz(lower[r4]))) : tprev[r4];
					term[r4] = __ldg(log2fix + (i - prev));
				}
				const int trow = rg - first_test;	/* row inside this super row's test span */
#pragma unroll
				for (int r4 = 0; r4 < UNI_GROUP; r4++)
					tbuf[(trow + r4) * 33 + lane] = term[r4];
			}
		}
		/* universal.c:369: sequential accumulation in block order, reproduced exactly, one super row at a time */
		const int nrows_t = row_end - first_test;
		if (nrows_t > 0) {
			__syncwarp();
			if (!warp_rne_sum_rows_fixed(sum, tbuf, nrows_t, lane)) {
				/* a tie or a binade change inside this super row: composable maps on 128 terms at a time, then term by term */
				for (int r = 0; r < nrows_t; r += UNI_GROUP) {
					const unsigned long long *tb = tbuf + (r + (lane >> 3)) * 33 + 4 * (lane & 7);
					const double4 t4 = make_double4(rne_fixed_to_double(tb[0]), rne_fixed_to_double(tb[1]),
									rne_fixed_to_double(tb[2]), rne_fixed_to_double(tb[3]));
					if (!warp_rne_sum128(sum, t4, lane)) {
						for (int rr = r; rr < r + UNI_GROUP; rr++) {
#pragma unroll 8
							for (int j = 0; j < 32; j++)
								sum += rne_fixed_to_double(tbuf[rr * 33 + j]);
						}
					}
				}
			}
			__syncwarp();				/* tbuf is rewritten by the next super row */
		}
	}
	if (lane == 0)
		st[s * P.nst + P.st_off[STSGPU_TEST_UNIVERSAL]] = __double_as_longlong(sum);
}

/* ------------------------------------------------------------------------------------------ */
/*
 * The same test split over UNI_SEGS warps per stream.  universal_kernel above is one dependency chain per stream - the
 * table T[] is read and written once per row of 32 blocks, every row waits for its log table lookups, and MATCH.ANY
 * costs about 70 cycles of an SM's MIO queue - with seven warps per SM when a batch of 1024 streams is resident: 11 %
 * warps active, 1.07 ms per batch whatever the batch size.  The chain splits:
 *
 *   universal_dist_kernel  one CTA of UNI_SEGS warps per stream.  Warp w takes a contiguous range of rows.  The table
 *       state at the start of its range is "the last block before the range that held pattern p" for every p: found
 *       by scanning BACKWARDS from the range until every pattern has been seen (about 2^L ln 2^L blocks, 22 rows at
 *       L = 7) or the stream's first block is reached (a pattern never seen keeps the reference's initial 0,
 *       universal.c:327).  Then the warp walks its rows like universal_kernel does - equal patterns inside a row are
 *       found with one ballot per pattern bit (ALU work, which this kernel has to spare) - and writes the DISTANCE
 *       i - T[.] of every test block, an exact integer, to a scratch array in block order.  It also leaves a rough
 *       sum of log2(distance) per "super row" of 1024 test blocks (float logarithms, fixed point, integer atomics:
 *       the same value on every run), from which the sum kernel predicts the binade of the running sum.
 *   universal_sum_kernel   one CTA per stream.  The reference's sum (universal.c:369) is sequential and its rounding
 *       depends on the order, but while the running sum stays in one binade k and no addend is an exact tie, adding
 *       the 1024 terms of a super row is S += sum of (X + half) >> sh(k) over its fixed point terms X - an integer
 *       sum, any order (rne_sum.cuh).  So: (1) prefix sums of the rough super row sums give the binade each super row
 *       most probably starts and ends in; (2) all warps compute, each for its own super rows, that integer sum and
 *       whether a tie occurred, from the distances and the host's log table; (3) one warp walks the super rows in
 *       order with the TRUE running sum: where the prediction holds (true binade = predicted, no tie, no carry out of
 *       the binade) it adds the integer, otherwise (about 8 of 125 super rows at L = 7: the binade changes) it adds
 *       that super row exactly as universal_kernel does.  A wrong prediction only costs time.
 */
#define UNI_SEGS 8
#define UNI_SUM_WARPS 4		/* 128-thread CTAs at 64 registers: eight per SM, a batch of 1024 streams is resident at once */
#define UNI_APPROX_FX 20		/* rough sums: log2 in units of 2^-20 */

/* the L-bit pattern of block row * 32 + lane (bit offsets fit 32 bits: n < 2^31) */
__device__ __forceinline__ uint32_t uni_pattern(const uint32_t *__restrict__ w, uint32_t row, int lane, int L)
{
	const uint32_t pos = (row * 32u + (uint32_t) lane) * (uint32_t) L;
	const uint32_t hi = bswap32(__ldg(w + (pos >> 5))), lo = bswap32(__ldg(w + (pos >> 5) + 1));
	return __funnelshift_l(lo, hi, pos & 31u) >> (32 - L);
}

/*
 * The lanes of the warp that hold the same pattern as this one - what MATCH.ANY returns, without MATCH.ANY (about 70
 * cycles of the SM's MIO queue each, the limit of this kernel when 64 warps per SM issue one per row) and without a
 * ballot per pattern bit (140 instructions per row): every lane ORs its bit into the word of its pattern in a per-warp
 * table M[2^L] of lane masks and reads the word back; the highest lane of every pattern clears it again.
 */
__device__ __forceinline__ uint32_t uni_same(uint32_t *M, uint32_t v, int lane)
{
	atomicOr(&M[v], 1u << lane);
	__syncwarp();
	const uint32_t same = M[v];
	__syncwarp();
	if ((same >> lane) <= 1u)
		M[v] = 0;
	return same;
}

__global__ void __launch_bounds__(32 * UNI_SEGS)
universal_dist_kernel(DevParams P, const uint32_t *__restrict__ in, uint32_t *__restrict__ dist, unsigned long long *__restrict__ rough)
{
	extern __shared__ __align__(16) uint32_t usm[];		/* [UNI_SEGS][2][2^L]: last occurrence T, lane masks M; then the rough sums */
	const int L = P.univ_L;
	const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
	uint32_t *T = usm + ((size_t) (2 * seg) << L), *M = T + ((size_t) 1 << L);
	unsigned long long *rough_s = (unsigned long long *) (usm + ((size_t) (2 * UNI_SEGS) << L));
	const int64_t s = blockIdx.x;
	const uint32_t *__restrict__ w = in + s * P.pitch_words;
	const int64_t Q = P.univ_Q, K = P.univ_K;
	const int rows = (int) ((Q + K) >> 5), qrows = (int) (Q >> 5);
	const int nsr = (int) ((K + 1023) >> 10);
	const int groups = rows / UNI_GROUP;				/* rows is a multiple of UNI_GROUP for L >= 6 */
	const int row_lo = (int) ((int64_t) groups * seg / UNI_SEGS) * UNI_GROUP;
	const int row_hi = (int) ((int64_t) groups * (seg + 1) / UNI_SEGS) * UNI_GROUP;
	uint32_t *__restrict__ out = dist + s * K;

	for (int i = lane; i < (2 << L); i += 32)
		T[i] = 0;						/* T and M */
	for (int i = threadIdx.x; i < nsr; i += 32 * UNI_SEGS)
		rough_s[i] = 0;
	__syncthreads();
	/* the table as the rows before row_lo leave it: last occurrence of every pattern, found backwards */
	{
		int found = 0;
		for (int row = row_lo - 1; row >= 0 && found < (1 << L); row--) {
			const uint32_t v = uni_pattern(w, (uint32_t) row, lane, L);
			const uint32_t same = uni_same(M, v, lane);
			const bool set = ((same >> lane) <= 1u) && T[v] == 0;	/* highest lane of its pattern, pattern not seen yet */
			__syncwarp();
			if (set)
				T[v] = (uint32_t) row * 32u + lane + 1u;
			found += __popc(__ballot_sync(0xffffffffu, set));
			__syncwarp();
		}
	}
	unsigned int part = 0;						/* rough sum of this lane's terms in the current super row */
	for (int row0 = row_lo; row0 < row_hi; row0 += UNI_GROUP) {
		const bool testing = row0 >= qrows;
		uint32_t v[UNI_GROUP], lower[UNI_GROUP], tprev[UNI_GROUP];
		bool is_last[UNI_GROUP];
#pragma unroll
		for (int r4 = 0; r4 < UNI_GROUP; r4++)
			v[r4] = uni_pattern(w, (uint32_t) (row0 + r4), lane, L);
		/* one row at a time through the two tables: the same-pattern lanes of the row, then the last occurrence */
#pragma unroll
		for (int r4 = 0; r4 < UNI_GROUP; r4++) {
			tprev[r4] = T[v[r4]];
			const uint32_t same = uni_same(M, v[r4], lane);		/* two __syncwarp inside: T was read by every lane */
			lower[r4] = same & ((1u << lane) - 1u);
			is_last[r4] = (same >> lane) <= 1u;
			if (is_last[r4])
				T[v[r4]] = (uint32_t) (row0 + r4) * 32u + lane + 1u;	/* 1-based block number */
			__syncwarp();
		}
		if (testing) {
#pragma unroll
			for (int r4 = 0; r4 < UNI_GROUP; r4++) {
				const uint32_t i = (uint32_t) (row0 + r4) * 32u + lane + 1u;
				const uint32_t prev = lower[r4] ? (i - lane + (31 - __clz(lower[r4]))) : tprev[r4];
				const uint32_t d = i - prev;
				__stcg(out + (i - 1u - (uint32_t) Q), d);
				part += (unsigned int) (__log2f((float) d) * (float) (1 << UNI_APPROX_FX));
			}
			/* UNI_GROUP divides 32: a group never straddles two super rows */
			const int trow_next = row0 + UNI_GROUP - qrows;
			if ((trow_next & 31) == 0 || row0 + UNI_GROUP >= row_hi) {
				unsigned long long tot = (unsigned long long) part;
#pragma unroll
				for (int o = 16; o > 0; o >>= 1)
					tot += __shfl_xor_sync(0xffffffffu, tot, o);
				if (lane == 0)
					atomicAdd(&rough_s[(trow_next - 1) >> 5], tot);
				part = 0;
			}
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < nsr; i += 32 * UNI_SEGS)
		rough[s * nsr + i] = rough_s[i];
}

__global__ void __launch_bounds__(32 * UNI_SUM_WARPS)
universal_sum_kernel(DevParams P, const uint32_t *__restrict__ dist, const unsigned long long *__restrict__ rough,
		     const unsigned long long *__restrict__ log2fix, long long *__restrict__ st, int ntab)
{
	extern __shared__ __align__(16) unsigned long long ssm[];
	const int64_t s = blockIdx.x, K = P.univ_K;
	const int nsr = (int) ((K + 1023) >> 10), trows = (int) (K >> 5);
	unsigned long long *tbuf = ssm;				/* [32 * 33] the terms of one super row (exact path), then 128 doubles */
	unsigned long long *acc_s = ssm + 32 * 33 + 128;	/* [nsr] integer sum under the predicted binade */
	unsigned long long *ltab = acc_s + nsr;			/* [ntab] the head of the log table: distances are about geometric with mean 2^L */
	int *pred_s = (int *) (ltab + ntab);			/* [nsr] predicted binade, -1: none; negated - 2 when a tie was seen */
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t *__restrict__ d = dist + s * K;
	/*
	 * 128000 gathers per stream from a 1 MiB table cost an L1 tag look-up per lane (32 cycles of the SM per warp
	 * instruction): the kernel was bound by exactly that.  From shared memory a 32-lane gather is a few wavefronts.
	 */
	for (int i = threadIdx.x; i < ntab; i += 32 * UNI_SUM_WARPS)
		ltab[i] = __ldg(log2fix + i);
#define UNI_TERM(dv) ((dv) < (uint32_t) ntab ? ltab[(dv)] : __ldg(log2fix + (dv)))

	/* (1) binade of the running sum before and after every super row, from the rough sums */
	if (warp == 0) {
		double prefix = 0.0;
		for (int sr0 = 0; sr0 < nsr; sr0 += 32) {
			const int sr = sr0 + lane;
			double mine = (sr < nsr) ? (double) rough[s * nsr + sr] * (1.0 / (double) (1 << UNI_APPROX_FX)) : 0.0;
			double incl = mine;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const double t = __shfl_up_sync(0xffffffffu, incl, o);
				if (lane >= o)
					incl += t;
			}
			const double before = prefix + incl - mine, after = prefix + incl;
			if (sr < nsr) {
				int k = -1;
				if (before >= 2.0 && (sr + 1) * 32 <= trows) {		/* full super rows only */
					const int kb = (int) ((unsigned long long) __double_as_longlong(before) >> 52) - 1023;
					const double lo = __longlong_as_double((long long) (kb + 1023) << 52);	/* 2^kb */
					/* both ends well inside [2^kb, 2^(kb+1)): the rough sums are good to about 1e-6 */
					if (before > lo * (1.0 + 1e-4) && after < 2.0 * lo * (1.0 - 1e-4))
						k = kb;
				}
				pred_s[sr] = k;
			}
			prefix = __shfl_sync(0xffffffffu, after, 31);
		}
	}
	__syncthreads();
	/*
	 * (2) the integer sums, every warp its own super rows.  The sum does not depend on the order of the terms: eight
	 * coalesced 16-byte loads per lane, all in flight before the first lookup.
	 */
	for (int sr = warp; sr < nsr; sr += UNI_SUM_WARPS) {
		const int k = pred_s[sr];
		const int sh = k - 52 + RNE_FX;
		if (k < 0 || sh < 1 || sh > 63)
			continue;
		const unsigned long long half = 1ull << (sh - 1), mask = (half << 1) - 1ull;
		const uint4 *__restrict__ p = (const uint4 *) (d + (int64_t) sr * 1024) + lane;
		uint4 q[8];
#pragma unroll
		for (int j = 0; j < 8; j++)
			q[j] = __ldcg(p + 32 * j);		/* 512 contiguous bytes per warp instruction */
		unsigned long long acc = 0;
		bool tie = false;
#pragma unroll
		for (int j = 0; j < 8; j++) {
			const unsigned long long X0 = UNI_TERM(q[j].x), X1 = UNI_TERM(q[j].y);
			const unsigned long long X2 = UNI_TERM(q[j].z), X3 = UNI_TERM(q[j].w);
			acc += ((X0 + half) >> sh) + ((X1 + half) >> sh) + ((X2 + half) >> sh) + ((X3 + half) >> sh);
			tie = tie || ((X0 & mask) == half) || ((X1 & mask) == half) || ((X2 & mask) == half) || ((X3 & mask) == half);
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
			acc += __shfl_xor_sync(0xffffffffu, acc, o);
		tie = __any_sync(0xffffffffu, tie);
		if (lane == 0) {
			acc_s[sr] = acc;
			if (tie)
				pred_s[sr] = -2 - k;
		}
	}
	__syncthreads();
	/* (3) universal.c:369 in block order with the true running sum */
	if (warp != 0)
		return;
	double sum = 0.0;
	for (int sr = 0; sr < nsr; sr++) {
		const int k = pred_s[sr];
		if (k >= 0) {
			const unsigned long long sb = (unsigned long long) __double_as_longlong(sum);
			if ((int) (sb >> 52) - 1023 == k) {
				const unsigned long long S = (sb & 0x000FFFFFFFFFFFFFull) | 0x0010000000000000ull;
				const unsigned long long Snew = S + acc_s[sr];
				if (Snew < 0x0020000000000000ull) {		/* stays in the binade: the prediction held */
					sum = __longlong_as_double((long long) (((unsigned long long) (k + 1023) << 52) | (Snew & 0x000FFFFFFFFFFFFFull)));
					continue;
				}
			}
		}
		/* exactly, as universal_kernel does: 1024 terms in block order through the composable maps, at worst one by one */
		const int r0 = sr * 32, nrows_t = min(32, trows - r0);
		{	/* block e = 128 j + 4 lane + c of the super row -> row e / 32 = 4 j + lane / 8, column (4 lane) % 32 + c */
			const uint4 *__restrict__ p = (const uint4 *) (d + (int64_t) r0 * 32) + lane;
			uint4 q[8];
#pragma unroll
			for (int j = 0; j < 8; j++)
				q[j] = (4 * j + (lane >> 3) < nrows_t) ? __ldcg(p + 32 * j) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
			for (int j = 0; j < 8; j++) {
				unsigned long long *tb = tbuf + (4 * j + (lane >> 3)) * 33 + 4 * (lane & 7);
				if (4 * j + (lane >> 3) < nrows_t) {
					tb[0] = UNI_TERM(q[j].x);
					tb[1] = UNI_TERM(q[j].y);
					tb[2] = UNI_TERM(q[j].z);
					tb[3] = UNI_TERM(q[j].w);
				}
			}
		}
		__syncwarp();
		if (!warp_rne_sum_rows_fixed(sum, tbuf, nrows_t, lane)) {
			/*
			 * the binade changes (or a tie) somewhere in this super row: 128 terms at a time, each group by the
			 * cheapest method that is exact for it - the integer sum, the composable round-to-even maps, and for the
			 * one group that holds the binade change the reference's own loop, one DADD per term
			 */
			for (int r = 0; r < nrows_t; r += UNI_GROUP) {
				if (warp_rne_sum_rows_fixed(sum, tbuf + r * 33, UNI_GROUP, lane))
					continue;
				const unsigned long long *tb = tbuf + (r + (lane >> 3)) * 33 + 4 * (lane & 7);
				const double4 t4 = make_double4(rne_fixed_to_double(tb[0]), rne_fixed_to_double(tb[1]),
								rne_fixed_to_double(tb[2]), rne_fixed_to_double(tb[3]));
				if (warp_rne_sum128(sum, t4, lane))
					continue;
				/* the conversions in parallel, the additions in order */
				double *seq = (double *) (tbuf + 32 * 33);
				__syncwarp();
				*(double4 *) (seq + 4 * lane) = t4;
				__syncwarp();
#pragma unroll 16
				for (int j = 0; j < 32 * UNI_GROUP; j++)
					sum += seq[j];
				__syncwarp();
			}
		}
		__syncwarp();				/* tbuf is rewritten by the next exact super row */
	}
	if (lane == 0)
		st[s * P.nst + P.st_off[STSGPU_TEST_UNIVERSAL]] = __double_as_longlong(sum);
}

/* bytes of distance scratch per stream for the split kernels, or 0 when this configuration stays on universal_kernel */
size_t universal_scratch_bytes_per_stream(const DevParams &P)
{
	if (!(P.enabled & (1u << STSGPU_TEST_UNIVERSAL)) || P.univ_L > 9)	/* 2 x UNI_SEGS tables of 2^L words must stay small */
		return 0;
	/* the distances, then one rough sum per super row of 1024 test blocks (8-byte aligned: K is a multiple of 32) */
	return (size_t) P.univ_K * sizeof(uint32_t) + (size_t) ((P.univ_K + 1023) >> 10) * sizeof(unsigned long long);
}

cudaError_t launch_universal(const DevParams &P, const LaunchCtx &L, const unsigned long long *d_log2_table, uint32_t *d_dist,
			     int64_t *launches)
{
	*launches = 0;
	if (!(P.enabled & (1u << STSGPU_TEST_UNIVERSAL)))
		return cudaSuccess;
	static int split = -1;
	if (split < 0)
		split = getenv("STSGPU_UNIVERSAL_SPLIT") ? atoi(getenv("STSGPU_UNIVERSAL_SPLIT")) : 1;
	if (split && d_dist != nullptr && universal_scratch_bytes_per_stream(P) != 0) {
		const int nsr = (int) ((P.univ_K + 1023) >> 10);
		/* scratch layout: [capacity of the batch][K] distances are not needed at full capacity: this batch's streams first */
		unsigned long long *d_rough = (unsigned long long *) (d_dist + (size_t) L.nstreams * P.univ_K + ((L.nstreams * P.univ_K) & 1));
		const size_t smem1 = ((size_t) 2 * UNI_SEGS * sizeof(uint32_t) << P.univ_L) + (size_t) nsr * sizeof(unsigned long long);
		if (smem1 > 48 * 1024) {
			cudaError_t e = cudaFuncSetAttribute(universal_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem1);
			if (e != cudaSuccess)
				return e;
		}
		/* head of the log table kept in shared memory: 16 mean distances (P(d >= 16 * 2^L) = e^-16), the rest from global */
		int64_t ntab = (int64_t) 16 << P.univ_L;
		if (ntab > P.univ_Q + P.univ_K + 1)
			ntab = P.univ_Q + P.univ_K + 1;
		const size_t smem2 = (size_t) (32 * 33 + 128 + nsr + ntab) * sizeof(unsigned long long) + (size_t) nsr * sizeof(int);
		if (smem2 > 48 * 1024) {
			cudaError_t e = cudaFuncSetAttribute(universal_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem2);
			if (e != cudaSuccess)
				return e;
		}
		universal_dist_kernel<<<(unsigned) L.nstreams, 32 * UNI_SEGS, smem1, L.stream>>>(P, L.d_in, d_dist, d_rough);
		universal_sum_kernel<<<(unsigned) L.nstreams, 32 * UNI_SUM_WARPS, smem2, L.stream>>>(P, d_dist, d_rough, d_log2_table, L.d_st, (int) ntab);
		*launches = 2;
		return cudaGetLastError();
	}
	*launches = 1;
	const size_t smem = sizeof(uint32_t) * ((size_t) 2 * 32 * 33 + 2 * (32 * P.univ_L + 2) + ((size_t) 1 << P.univ_L));
	/* L = 15 needs 143376 bytes (128 KiB of table + staging), 16 more than the 140 KiB asked for until the end of
	 * round 1: the launch would have been refused (seen on the CPU emulator, which enforces the attribute) */
	const size_t allow = smem > 140 * 1024 ? smem : (size_t) 140 * 1024;
	cudaError_t e = cudaFuncSetAttribute(universal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) allow);
	if (e != cudaSuccess)
		return e;
	universal_kernel<<<(unsigned) L.nstreams, 32, smem, L.stream>>>(P, L.d_in, d_log2_table, L.d_st);
	return cudaGetLastError();
}
