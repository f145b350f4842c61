/*
 * k_rank.cu - binary matrix rank test (rank.c:303-327 with utils/matrix.c:49-200).
 *
 * One 32x32 matrix per warp, one row per lane as a packed 32-bit word (row i of matrix k is
 * stream word 32k + i, def_matrix matrix.c:284-288; column j is bit 31 - j).  The statistic is NOT
 * the mathematical GF(2) rank: it is the number of non-zero rows left by NIST's schedule - forward
 * elimination for i = 0..30 (pivot search downwards, first hit), then backward elimination for
 * i = 31..1 (pivot search upwards, first hit) - which this kernel follows step by step with
 * ballot (pivot search) and shuffles (row swap / pivot broadcast).  Whole-row XOR equals the
 * reference's partial-row XOR because columns < i of rows >= i are already zero in the forward pass.
 *
 * Algorithmic bytes: n/8 per stream (each word is read exactly once, 128-byte coalesced).
 */
#include "common.cuh"

#define RANK_WARPS 8

__device__ __forceinline__ int nist_rank32(uint32_t row, int lane)
{
	/* forward, matrix.c:59-69 */
	for (int i = 0; i < 31; i++) {
		const int sh = 31 - i;
		uint32_t colbit = (row >> sh) & 1u;
		uint32_t col = __ballot_sync(0xffffffffu, colbit);
		if (!((col >> i) & 1u)) {
			uint32_t cand = col & (0xFFFFFFFEu << i);	/* rows > i */
			if (cand == 0)
				continue;
			int idx = __ffs(cand) - 1;			/* first row below */
			uint32_t ri = __shfl_sync(0xffffffffu, row, i);
			uint32_t rx = __shfl_sync(0xffffffffu, row, idx);
			if (lane == i)
				row = rx;
			else if (lane == idx)
				row = ri;
			colbit = (row >> sh) & 1u;
		}
		uint32_t prow = __shfl_sync(0xffffffffu, row, i);
		if (lane > i && colbit)
			row ^= prow;
	}
	/* backward, matrix.c:74-84 */
	for (int i = 31; i > 0; i--) {
		const int sh = 31 - i;
		uint32_t colbit = (row >> sh) & 1u;
		uint32_t col = __ballot_sync(0xffffffffu, colbit);
		if (!((col >> i) & 1u)) {
			uint32_t cand = col & ((1u << i) - 1u);		/* rows < i */
			if (cand == 0)
				continue;
			int idx = 31 - __clz(cand);			/* first row above */
			uint32_t ri = __shfl_sync(0xffffffffu, row, i);
			uint32_t rx = __shfl_sync(0xffffffffu, row, idx);
			if (lane == i)
				row = rx;
			else if (lane == idx)
				row = ri;
			colbit = (row >> sh) & 1u;
		}
		uint32_t prow = __shfl_sync(0xffffffffu, row, i);
		if (lane < i && colbit)
			row ^= prow;
	}
	return __popc(__ballot_sync(0xffffffffu, row != 0u));	/* matrix.c:162-188 */
}

__global__ void __launch_bounds__(RANK_WARPS * 32)
rank_kernel(DevParams P, const uint32_t *__restrict__ in, long long *__restrict__ st)
{
	__shared__ unsigned int f[2];
	const int64_t s = blockIdx.x;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x < 2)
		f[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t *__restrict__ w = in + s * P.pitch_words;
	unsigned int f32 = 0, f31 = 0;
	for (int64_t k = (int64_t) blockIdx.y * RANK_WARPS + warp; k < P.rank_count; k += (int64_t) gridDim.y * RANK_WARPS) {
		uint32_t row = ldw(w, k * 32 + lane);
		int R = nist_rank32(row, lane);
		f32 += (R == 32);
		f31 += (R == 31);
	}
	if (lane == 0) {
		if (f32) atomicAdd(&f[0], f32);
		if (f31) atomicAdd(&f[1], f31);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long *o = (unsigned long long *) (st + s * P.nst + P.st_off[STSGPU_TEST_RANK]);
		if (f[0]) atomicAdd(&o[0], (unsigned long long) f[0]);
		if (f[1]) atomicAdd(&o[1], (unsigned long long) f[1]);
		/* F_remaining = matrix_count - F_32 - F_31 is filled by the tail kernel */
	}
}

cudaError_t launch_rank(const DevParams &P, const LaunchCtx &L)
{
	if (!(P.enabled & (1u << STSGPU_TEST_RANK)))
		return cudaSuccess;
	int64_t gy = (P.rank_count + RANK_WARPS * 16 - 1) / (RANK_WARPS * 16);	/* >= 16 matrices per warp */
	if (gy < 1) gy = 1;
	if (gy > 64) gy = 64;
	dim3 grid((unsigned) L.nstreams, (unsigned) gy);
	rank_kernel<<<grid, RANK_WARPS * 32, 0, L.stream>>>(P, L.d_in, L.d_st);
	return cudaGetLastError();
}
