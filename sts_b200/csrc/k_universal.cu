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

cudaError_t launch_universal(const DevParams &P, const LaunchCtx &L, const unsigned long long *d_log2_table)
{
	if (!(P.enabled & (1u << STSGPU_TEST_UNIVERSAL)))
		return cudaSuccess;
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
