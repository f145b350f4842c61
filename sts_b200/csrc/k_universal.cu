/*
 * k_universal.cu - Maurer's universal statistical test (universal.c:318-375).
 *
 * The statistic is  sum += log(i - T[block_i]) / log(2)  over K = 1000 * 2^L test blocks, a
 * sequential double-precision sum of 128000 terms at L = 7 whose rounding depends on the order.
 * To reproduce the reference's value bit for bit the kernel
 *   - finds every distance i - T[.] exactly (integers): a warp takes 32 consecutive L-bit blocks,
 *     `match.any` resolves repeats inside the warp and a shared-memory table T[2^L] carries the
 *     last occurrence across rows (table starts at 0 like the reference's memset, universal.c:327);
 *   - looks the term up in a table log2tab[d] = log(d) / log(2.0) computed ON THE HOST with the
 *     reference's expression and libm, so each addend is the reference's double;
 *   - adds the 32 terms of a row in block order, one after the other (shuffle + DADD chain).
 * One warp per stream; the DADD dependency chain is the cost (~10 cycles per block), which other
 * kernels on other streams hide.
 *
 * Algorithmic bytes: ceil((Q + K) * L / 8) per stream (113120 at L = 7) + the L2-resident table.
 */
#include "common.cuh"

__global__ void __launch_bounds__(32)
universal_kernel(DevParams P, const uint32_t *__restrict__ in, const double *__restrict__ log2tab,
		 long long *__restrict__ st)
{
	extern __shared__ uint32_t T[];
	const int64_t s = blockIdx.x;
	const int lane = threadIdx.x;
	const int L = P.univ_L;
	const uint32_t *__restrict__ w = in + s * P.pitch_words;
	for (int i = lane; i < (1 << L); i += 32)
		T[i] = 0;
	__syncwarp();
	const int64_t Q = P.univ_Q, K = P.univ_K;
	double sum = 0.0;
	for (int64_t base = 0; base < Q + K; base += 32) {
		const int64_t i = base + lane + 1;			/* 1-based block number */
		const uint32_t v = get32(w, (i - 1) * L) >> (32 - L);
		const uint32_t same = __match_any_sync(0xffffffffu, v);
		const uint32_t lower = same & ((1u << lane) - 1u);
		const bool is_last = (same >> lane) <= 1u;		/* no higher lane holds the same pattern */
		if (base >= Q) {
			uint32_t prev = lower ? (uint32_t) (base + (31 - __clz(lower)) + 1) : T[v];
			double term = __ldg(log2tab + (i - prev));
			/* universal.c:369: sequential accumulation in block order */
#pragma unroll
			for (int l = 0; l < 32; l++)
				sum += __shfl_sync(0xffffffffu, term, l);
		}
		__syncwarp();
		if (is_last)
			T[v] = (uint32_t) i;
		__syncwarp();
	}
	if (lane == 0)
		st[s * P.nst + P.st_off[STSGPU_TEST_UNIVERSAL]] = __double_as_longlong(sum);
}

cudaError_t launch_universal(const DevParams &P, const LaunchCtx &L, const double *d_log2_table)
{
	if (!(P.enabled & (1u << STSGPU_TEST_UNIVERSAL)))
		return cudaSuccess;
	size_t smem = (size_t) sizeof(uint32_t) << P.univ_L;
	cudaError_t e = cudaFuncSetAttribute(universal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
	if (e != cudaSuccess)
		return e;
	universal_kernel<<<(unsigned) L.nstreams, 32, smem, L.stream>>>(P, L.d_in, d_log2_table, L.d_st);
	return cudaGetLastError();
}
