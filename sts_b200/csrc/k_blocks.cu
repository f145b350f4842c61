/*
 * k_blocks.cu - the three "independent M-bit block" tests, one warp per block:
 *   BlockFrequency            blockFrequency.c:217-237  ones per block (M = tp.blockFrequencyBlockLength)
 *   LongestRunOfOnes          longestRunOfOnes.c:383-409 longest 1-run per block -> 7 class counts
 *   OverlappingTemplate (m=9) overlappingTemplateMatchings.c:262-296  #windows of 9 ones per
 *                             1032-bit block -> 6 class counts
 * Blocks are not word aligned in general (10000, 1032 bits), so every lane pulls 32-bit groups at
 * arbitrary bit offsets with a funnel shift (get32) and masks the ragged end.
 *
 * Algorithmic bytes: n/8 per stream per test.
 */
#include "common.cuh"

#define BLK_WARPS 8

/* ---- BlockFrequency: grid (streams, ceil(N / BLK_WARPS)) ---- */
__global__ void __launch_bounds__(BLK_WARPS * 32)
block_frequency_kernel(DevParams P, const uint32_t *__restrict__ in, long long *__restrict__ st)
{
	const int64_t s = blockIdx.x;
	const int lane = threadIdx.x & 31;
	const int64_t b = (int64_t) blockIdx.y * BLK_WARPS + (threadIdx.x >> 5);
	if (b >= P.bf_N)
		return;
	const uint32_t *__restrict__ w = in + s * P.pitch_words;
	const int64_t start = b * P.bf_M, M = P.bf_M;
	int ones = 0;
	for (int64_t g = lane; g * 32 < M; g += 32) {
		uint32_t v = get32(w, start + g * 32);
		int64_t rem = M - g * 32;
		if (rem < 32)
			v &= mask_below((uint32_t) rem);
		ones += __popc(v);
	}
	ones = warp_sum(ones);
	if (lane == 0) {
		long long *o = st + s * P.nst + P.st_off[STSGPU_TEST_BLOCK_FREQUENCY];
		o[1 + b] = ones;
		if (b == 0)
			o[0] = P.bf_N;
	}
}

/* ---- LongestRunOfOnes ---- */
struct RunSummary {
	int pre;	/* ones at the start of the span */
	int suf;	/* ones at the end of the span */
	int best;	/* longest run inside */
	int len;	/* span length in bits */
};

__device__ __forceinline__ RunSummary run_combine(const RunSummary &a, const RunSummary &b)
{
	RunSummary r;
	r.len = a.len + b.len;
	r.pre = (a.pre == a.len) ? a.len + b.pre : a.pre;
	r.suf = (b.suf == b.len) ? b.len + a.suf : b.suf;
	r.best = max(max(a.best, b.best), a.suf + b.pre);
	return r;
}

/* summary of the first `nb` (1..32) bits of v (MSB first); bits past nb must be zero */
__device__ __forceinline__ RunSummary run_of_word(uint32_t v, int nb)
{
	RunSummary r;
	r.len = nb;
	r.pre = min(__clz(~v), nb);
	/* trailing ones of the nb-bit field */
	uint32_t f = v >> (32 - nb);			/* right-aligned field */
	r.suf = min(__ffs(~f) - 1 < 0 ? 32 : __ffs(~f) - 1, nb);
	int best = 0;
	uint32_t x = v;
	while (x) {
		x &= x << 1;
		best++;
	}
	r.best = best;
	return r;
}

__global__ void __launch_bounds__(BLK_WARPS * 32)
longest_run_kernel(DevParams P, const uint32_t *__restrict__ in, long long *__restrict__ st)
{
	__shared__ unsigned int cls[7];
	const int64_t s = blockIdx.x;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x < 7)
		cls[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t *__restrict__ w = in + s * P.pitch_words;
	const int64_t M = P.lr_M;
	const int64_t groups = (M + 31) >> 5;			/* 32-bit groups per block */
	const int64_t gpl = (groups + 31) >> 5;			/* groups per lane (contiguous) */
	for (int64_t b = (int64_t) blockIdx.y * BLK_WARPS + warp; b < P.lr_N; b += (int64_t) gridDim.y * BLK_WARPS) {
		const int64_t start = b * M;
		RunSummary acc;
		acc.pre = acc.suf = acc.best = acc.len = 0;
		for (int64_t g = lane * gpl; g < min(groups, (lane + 1) * gpl); g++) {
			int64_t rem = M - g * 32;
			int nb = rem >= 32 ? 32 : (int) rem;
			uint32_t v = get32(w, start + g * 32) & mask_below((uint32_t) nb);
			RunSummary r = run_of_word(v, nb);
			acc = (acc.len == 0) ? r : run_combine(acc, r);
		}
		/* ordered combine across lanes (empty spans have len 0 and are the identity) */
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			RunSummary nb_;
			nb_.pre = __shfl_down_sync(0xffffffffu, acc.pre, o);
			nb_.suf = __shfl_down_sync(0xffffffffu, acc.suf, o);
			nb_.best = __shfl_down_sync(0xffffffffu, acc.best, o);
			nb_.len = __shfl_down_sync(0xffffffffu, acc.len, o);
			if (lane + o < 32 && (lane & (2 * o - 1)) == 0) {
				if (acc.len == 0)
					acc = nb_;
				else if (nb_.len != 0)
					acc = run_combine(acc, nb_);
			}
		}
		if (lane == 0) {
			int v_obs = acc.best, c;
			if (v_obs <= P.lr_min_class)
				c = 0;
			else if (v_obs <= P.lr_max_class)
				c = v_obs - P.lr_min_class;
			else
				c = 6;
			atomicAdd(&cls[c], 1u);
		}
	}
	__syncthreads();
	if (threadIdx.x < 7 && cls[threadIdx.x])
		atomicAdd((unsigned long long *) (st + s * P.nst + P.st_off[STSGPU_TEST_LONGEST_RUN] + threadIdx.x),
			  (unsigned long long) cls[threadIdx.x]);
}

/* ---- Overlapping template of nine ones, blocks of 1032 bits ---- */
__global__ void __launch_bounds__(BLK_WARPS * 32)
overlapping_kernel(DevParams P, const uint32_t *__restrict__ in, long long *__restrict__ st)
{
	__shared__ unsigned int cls[6];
	const int64_t s = blockIdx.x;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x < 6)
		cls[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t *__restrict__ w = in + s * P.pitch_words;
	for (int64_t b = (int64_t) blockIdx.y * BLK_WARPS + warp; b < P.ov_N; b += (int64_t) gridDim.y * BLK_WARPS) {
		/* window starts j = 0 .. 1032 - 9 = 1023: exactly 32 groups of 32 starts, one per lane */
		const int64_t pos = b * 1032 + lane * 32;
		uint32_t cur = get32(w, pos), nxt = get32(w, pos + 32);
		uint32_t r = cur;
#pragma unroll
		for (int k = 1; k < 9; k++)
			r &= __funnelshift_l(nxt, cur, k);	/* bit p set iff bits p..p+k are ones */
		int cnt = warp_sum(__popc(r));
		if (lane == 0)
			atomicAdd(&cls[cnt < 5 ? cnt : 5], 1u);
	}
	__syncthreads();
	if (threadIdx.x < 6 && cls[threadIdx.x])
		atomicAdd((unsigned long long *) (st + s * P.nst + P.st_off[STSGPU_TEST_OVERLAPPING] + threadIdx.x),
			  (unsigned long long) cls[threadIdx.x]);
}

cudaError_t launch_blocks(const DevParams &P, const LaunchCtx &L)
{
	if ((P.enabled & (1u << STSGPU_TEST_BLOCK_FREQUENCY)) && P.bf_N > 0) {	/* N = 0: slot 0 stays 0 */
		dim3 grid((unsigned) L.nstreams, (unsigned) ((P.bf_N + BLK_WARPS - 1) / BLK_WARPS));
		block_frequency_kernel<<<grid, BLK_WARPS * 32, 0, L.stream>>>(P, L.d_in, L.d_st);
	}
	if ((P.enabled & (1u << STSGPU_TEST_LONGEST_RUN)) && P.lr_N > 0) {
		int64_t gy = (P.lr_N + BLK_WARPS - 1) / BLK_WARPS;
		if (gy > 16) gy = 16;
		if (gy < 1) gy = 1;
		dim3 grid((unsigned) L.nstreams, (unsigned) gy);
		longest_run_kernel<<<grid, BLK_WARPS * 32, 0, L.stream>>>(P, L.d_in, L.d_st);
	}
	if (P.enabled & (1u << STSGPU_TEST_OVERLAPPING)) {
		int64_t gy = (P.ov_N + BLK_WARPS - 1) / BLK_WARPS;
		if (gy > 16) gy = 16;
		dim3 grid((unsigned) L.nstreams, (unsigned) gy);
		overlapping_kernel<<<grid, BLK_WARPS * 32, 0, L.stream>>>(P, L.d_in, L.d_st);
	}
	return cudaGetLastError();
}
