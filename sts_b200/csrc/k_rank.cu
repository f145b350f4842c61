/*
 * k_rank.cu - binary matrix rank test (rank.c:303-327 with utils/matrix.c:49-200).
 *
 * One 32x32 matrix per THREAD: the 32 rows live in 32 registers as packed words (row i of matrix k
 * is stream word 32k + i, def_matrix matrix.c:284-288; column j is bit 31 - j) and the whole
 * elimination is straight-line, branch-free register code - no shuffles, no ballots, no divergence.
 * The statistic is NOT the mathematical GF(2) rank: it is the number of non-zero rows left by
 * NIST's schedule - forward elimination for i = 0..30 (pivot search downwards, first hit), then
 * backward elimination for i = 31..1 (pivot search upwards, first hit).  Search, swap and
 * elimination of one column are fused into one sweep over the rows r beyond i with the pivot
 * candidate p carried in a register:
 *     row r has the bit, p has not  ->  swap (p, row r)          (matrix.c:90-133 first hit + swap)
 *     row r has the bit, p has it   ->  row r ^= p                (matrix.c:64-67 / :79-82)
 * which visits the rows in the reference's order, so the first row with the bit becomes the pivot
 * and every later one is reduced by it; a column without any pivot leaves all rows untouched.
 * Whole-row XOR equals the reference's partial-row XOR because the columns already processed are
 * zero in the rows still being combined.
 *
 * A CTA stages its 128 matrices (16 KiB of the stream, 128-bit coalesced loads) in shared memory
 * with a 33-word pitch so that the per-thread row reads are bank-conflict free.
 *
 * Algorithmic bytes: n/8 per stream (each word is read exactly once).
 */
#include <stdlib.h>
#include "common.cuh"

#define RANK_THREADS 128

__device__ __forceinline__ int nist_rank32_regs(uint32_t (&row)[32])
{
	/* forward, matrix.c:59-69 */
#pragma unroll
	for (int i = 0; i < 31; i++) {
		const uint32_t bit = 0x80000000u >> i;
		uint32_t p = row[i];
		bool ph = (p & bit) != 0u;
#pragma unroll
		for (int r = i + 1; r < 32; r++) {
			const uint32_t x = row[r];
			const bool t = (x & bit) != 0u;
			const uint32_t xp = x ^ p;
			row[r] = t ? (ph ? xp : p) : x;
			p = (t && !ph) ? x : p;
			ph = ph || t;
		}
		row[i] = p;
	}
	/* backward, matrix.c:74-84 */
#pragma unroll
	for (int i = 31; i > 0; i--) {
		const uint32_t bit = 0x80000000u >> i;
		uint32_t p = row[i];
		bool ph = (p & bit) != 0u;
#pragma unroll
		for (int r = i - 1; r >= 0; r--) {
			const uint32_t x = row[r];
			const bool t = (x & bit) != 0u;
			const uint32_t xp = x ^ p;
			row[r] = t ? (ph ? xp : p) : x;
			p = (t && !ph) ? x : p;
			ph = ph || t;
		}
		row[i] = p;
	}
	int rank = 0;							/* matrix.c:162-188 */
#pragma unroll
	for (int i = 0; i < 32; i++)
		rank += (row[i] != 0u) ? 1 : 0;
	return rank;
}

/*
 * Grid-stride over the (stream, 128-matrix group) work items: the kernel is bound by the integer ALU pipe, which a
 * few warps per scheduler already saturate, so it is launched as a resident grid of STSGPU_RANK_BG CTAs per SM
 * instead of one CTA per item - the registers a full-occupancy grid would hold are what the FFT and
 * LinearComplexity CTAs on the same SM need ("register file").  Pipelined batch per 1024 streams: one CTA
 * per item 9.17 - 9.29 ms, 2 CTAs per SM 9.08, 3: 9.09 - 9.12, 4: 9.16 - 9.18, 6: 9.15 (the kernel alone: 0.52 ms at full
 * occupancy, 1.0 ms with 3 CTAs per SM).
 */
__global__ void __launch_bounds__(RANK_THREADS)
rank_kernel(DevParams P, const uint32_t *__restrict__ in, long long *__restrict__ st, int64_t nstreams, int groups_per_stream)
{
	__shared__ uint32_t tile[RANK_THREADS * 33];
	__shared__ unsigned int f[2];
	const int tid = threadIdx.x;
	for (int64_t item = blockIdx.x; item < nstreams * groups_per_stream; item += gridDim.x) {
	const int64_t s = item / groups_per_stream;
	const int by = (int) (item - s * groups_per_stream);
	__syncthreads();					/* tile[] and f[] of the previous item are done with */
	if (tid < 2)
		f[tid] = 0;
	const int64_t m0 = (int64_t) by * RANK_THREADS;			/* first matrix of this item */
	const int nm = (int) min((int64_t) RANK_THREADS, P.rank_count - m0);
	const uint4 *__restrict__ src = (const uint4 *) (in + s * P.pitch_words + m0 * 32);
#pragma unroll
	for (int j = 0; j < 8; j++) {
		const int q = tid + RANK_THREADS * j;			/* 16-byte unit inside the tile */
		const int m = q >> 3, c = (q & 7) * 4;
		if (m < nm) {
			const uint4 v = __ldg(src + q);
			uint32_t *d = tile + m * 33 + c;
			d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
		}
	}
	__syncthreads();
	int R = -1;
	if (tid < nm) {
		uint32_t row[32];
#pragma unroll
		for (int i = 0; i < 32; i++)
			row[i] = bswap32(tile[tid * 33 + i]);
		R = nist_rank32_regs(row);
	}
	const unsigned int b32 = __popc(__ballot_sync(0xffffffffu, R == 32));
	const unsigned int b31 = __popc(__ballot_sync(0xffffffffu, R == 31));
	if ((tid & 31) == 0) {
		if (b32) atomicAdd(&f[0], b32);
		if (b31) atomicAdd(&f[1], b31);
	}
	__syncthreads();
	if (tid == 0) {
		unsigned long long *o = (unsigned long long *) (st + s * P.nst + P.st_off[STSGPU_TEST_RANK]);
		if (f[0]) atomicAdd(&o[0], (unsigned long long) f[0]);
		if (f[1]) atomicAdd(&o[1], (unsigned long long) f[1]);
		/* F_remaining = matrix_count - F_32 - F_31 is filled by the tail kernel */
	}
	}
}

cudaError_t launch_rank(const DevParams &P, const LaunchCtx &L)
{
	if (!(P.enabled & (1u << STSGPU_TEST_RANK)) || P.rank_count < 1)
		return cudaSuccess;
	const int gy = (int) ((P.rank_count + RANK_THREADS - 1) / RANK_THREADS);
	const int64_t items = L.nstreams * gy;
	static int bg = -1;
	if (bg < 0)
		bg = getenv("STSGPU_RANK_BG") ? atoi(getenv("STSGPU_RANK_BG")) : 3;
	int64_t ctas = bg > 0 ? (int64_t) bg * L.num_sms : items;
	if (ctas > items) ctas = items;
	rank_kernel<<<(unsigned) ctas, RANK_THREADS, 0, L.stream>>>(P, L.d_in, L.d_st, L.nstreams, gy);
	return cudaGetLastError();
}
