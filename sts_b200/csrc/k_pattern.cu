/*
 * k_pattern.cu - the m-bit pattern histogram family.
 *
 * (1) NonOverlappingTemplateMatchings (nonOverlappingTemplateMatchings.c:495-575).  The accepted
 *     templates have no self-overlap (appendTemplate :374-395), so the reference's greedy scan
 *     "count a hit, skip m bits" counts every occurrence: W[t][i] is bin value(t) of the 2^m-bin
 *     histogram of all m-bit windows lying inside block i (positions 0 .. M-m).  One CTA builds
 *     that histogram for one (stream, block) in shared memory with a sliding funnel-shift window
 *     and gathers the bins of all templates: the whole template set costs one pass.
 *
 * (2) Serial (serial.c:390-419) and ApproximateEntropy (approximateEntropy.c:369-407) both count
 *     all n cyclic windows; because the windowing is cyclic, hist_{b-1}[x] = hist_b[2x] + hist_b[2x+1]
 *     exactly, so ONE histogram at H = max(serial_m, apen_m + 1) bits gives every level by folding.
 *     H can reach 21 bits (8 MiB of counters): a CTA owns the patterns whose top (H - 15) bits equal
 *     its slice index and keeps 2^min(H,15) 32-bit counters (<= 128 KiB) in shared memory.
 *     Serial leaves the kernel as exact integer sums of squares (sum v^2 <= n^2 < 2^63); ApEn
 *     leaves as its two histograms (index order matters for the FP sum done by the tail kernel).
 *
 * Algorithmic bytes: n/8 per stream per kernel (times the slice count for H > 15).
 */
#include <stdlib.h>
#include "common.cuh"

#define PAT_THREADS 256		/* nonoverlapping: 2 KiB histograms, many CTAs per SM */
#define HIST_THREADS 1024	/* cyclic histogram: 128 KiB per CTA = one CTA per SM, so make it 32 warps */

/* 32 bits starting at cyclic bit position pos (0 <= pos < n + 64), stream length n bits */
__device__ __forceinline__ uint32_t cyc32(const uint32_t *__restrict__ w, int64_t pos, int64_t n)
{
	if (pos + 32 <= n)
		return get32(w, pos);
	if (pos >= n)
		return get32(w, pos - n);	/* n >= 64 guaranteed by the callers' test preconditions */
	int k = (int) (n - pos);		/* 1..31 bits left before the wrap */
	return (get32(w, pos) & mask_below((uint32_t) k)) | (get32(w, 0) >> k);
}

/* ------------------------------------------------------------------------------------------ */
/*
 * One shared-memory atomic per window start made this kernel the slowest user of the atomic path (LSU wavefronts
 * 97 %).  Now one atomic counts S = 4 window starts at once: the histogram is over WIDE windows of m + S - 1 bits
 * taken at every S-th position - a wide window holds the m-bit windows at its S offsets - and the bin of a template
 * is folded back from it at the end: W[t] = sum over offset k and over the 2^(S-1) ways to extend the template to
 * the left by k and to the right by S - 1 - k bits.  A quarter of the atomics into eight times as many bins (fewer
 * collisions); exact, like the plain count.  The at most 31 window starts at the end of a block that do not fill a
 * word-aligned group of 32 are counted one by one in a plain 2^m-bin histogram.  S = 4 for m <= 10 (2^13 wide bins at
 * most), 2 for m <= 12, 1 above.
 */
__device__ __forceinline__ int nonov_stride(int m) { return m <= 10 ? 4 : (m <= 12 ? 2 : 1); }

__global__ void __launch_bounds__(PAT_THREADS)
nonoverlapping_kernel(DevParams P, const uint32_t *__restrict__ in, const uint32_t *__restrict__ tmpl,
		      uint32_t *__restrict__ wout)
{
	extern __shared__ uint32_t hist[];			/* [2^(m + S - 1)] wide windows, then (S > 1) [2^m] single windows */
	const int64_t s = blockIdx.x;
	const int blk = blockIdx.y;
	const int m = P.nonov_m;
	const int S = nonov_stride(m), Wd = m + S - 1;
	uint32_t *plain = S > 1 ? hist + (1 << Wd) : hist;	/* at S = 1 the two histograms are the same one */
	const int bins = S > 1 ? (1 << Wd) + (1 << m) : (1 << m);
	for (int i = threadIdx.x; i < bins; i += PAT_THREADS)
		hist[i] = 0;
	__syncthreads();
	const uint32_t *__restrict__ w = in + s * P.pitch_words;
	const int64_t M = P.nonov_M, start = (int64_t) blk * M;
	const int64_t last = M - m;				/* last window start inside the block */
	const int64_t groups = (last >> 5) + 1;
	for (int64_t g = threadIdx.x; g < groups; g += PAT_THREADS) {
		const int64_t j0 = g * 32;
		uint32_t cur = get32(w, start + j0), nxt = get32(w, start + j0 + 32);
		int cnt = (last - j0 >= 31) ? 32 : (int) (last - j0 + 1);
		if (cnt == 32) {			/* the common case: 32 / S wide windows */
			if (S == 4) {
#pragma unroll
				for (int p = 0; p < 32; p += 4)
					atomicAdd(&hist[__funnelshift_l(nxt, cur, p) >> (32 - Wd)], 1u);
			} else if (S == 2) {
#pragma unroll
				for (int p = 0; p < 32; p += 2)
					atomicAdd(&hist[__funnelshift_l(nxt, cur, p) >> (32 - Wd)], 1u);
			} else {
#pragma unroll
				for (int p = 0; p < 32; p++)
					atomicAdd(&hist[__funnelshift_l(nxt, cur, p) >> (32 - Wd)], 1u);
			}
		} else {
			for (int p = 0; p < cnt; p++)
				atomicAdd(&plain[__funnelshift_l(nxt, cur, p) >> (32 - m)], 1u);
		}
	}
	__syncthreads();
	uint32_t *__restrict__ o = wout + s * P.nw;
	for (int t = threadIdx.x; t < P.ntemplates; t += PAT_THREADS) {
		const uint32_t x = __ldg(tmpl + t);
		uint32_t c = plain[x];
		for (int k = 0; k < S && S > 1; k++) {			/* the template at offset k of a wide window */
			const int lo_bits = S - 1 - k;
			for (int e = 0; e < (1 << (S - 1)); e++) {
				const uint32_t hi = (uint32_t) e >> lo_bits, lo = (uint32_t) e & ((1u << lo_bits) - 1u);
				c += hist[(hi << (m + lo_bits)) | (x << lo_bits) | lo];
			}
		}
		o[t * 8 + blk] = c;
	}
}

cudaError_t launch_nonoverlapping(const DevParams &P, const LaunchCtx &L, const uint32_t *d_templates)
{
	if (!(P.enabled & (1u << STSGPU_TEST_NON_OVERLAPPING)))
		return cudaSuccess;
	const int m = P.nonov_m, S = m <= 10 ? 4 : (m <= 12 ? 2 : 1);
	size_t smem = ((size_t) sizeof(uint32_t) << (m + S - 1)) + (S > 1 ? (size_t) sizeof(uint32_t) << m : 0);
	cudaError_t e = cudaFuncSetAttribute(nonoverlapping_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
	if (e != cudaSuccess)
		return e;
	dim3 grid((unsigned) L.nstreams, 8);
	nonoverlapping_kernel<<<grid, PAT_THREADS, smem, L.stream>>>(P, L.d_in, d_templates, L.d_w);
	return cudaGetLastError();
}

/* ------------------------------------------------------------------------------------------ */
/* block-wide sum of unsigned 64-bit values; result valid in thread 0 */
__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long *scratch)
{
	v = (unsigned long long) warp_sum_ll((long long) v);
	__syncthreads();
	if ((threadIdx.x & 31) == 0)
		scratch[threadIdx.x >> 5] = v;
	__syncthreads();
	unsigned long long t = 0;
	if (threadIdx.x == 0)
		for (int i = 0; i < HIST_THREADS / 32; i++)
			t += scratch[i];
	return t;
}

/*
 * Statistics of every level from `Htop` bits down to the lowest one a test needs, from a histogram of
 * 32-bit counters at `Htop` bits whose top `sb` bits equal `slice`; folds in place between levels.
 */
__device__ void emit_levels(const DevParams &P, uint32_t *hist, unsigned long long *red, int Htop, int sb, uint32_t slice,
			    int64_t s, long long *__restrict__ st, uint32_t *__restrict__ apen_hist)
{
	const int H = Htop;
	const int sm = (P.enabled & (1u << STSGPU_TEST_SERIAL)) ? P.serial_m : -100;
	const int am = (P.enabled & (1u << STSGPU_TEST_APEN)) ? P.apen_m : -100;
	int lowest = H;
	if (sm > 0) lowest = min(lowest, max(sm - 2, 1));
	if (am > 0) lowest = min(lowest, am);
	long long *o = st + s * P.nst;
	for (int lvl = H; lvl >= lowest; lvl--) {
		const int b = lvl - sb;			/* local bits at this level (>= 0 by construction) */
		const int nb = 1 << b;
		const bool is_serial = (lvl == sm || lvl == sm - 1 || lvl == sm - 2);
		const bool is_apen = (lvl == am + 1 || lvl == am);
		if (is_serial || is_apen) {
			unsigned long long sq = 0, lin = 0;
			const unsigned long long base = (unsigned long long) slice << b;
			for (int i = threadIdx.x; i < nb; i += HIST_THREADS) {
				unsigned long long c = hist[i];
				sq += c * c;
				lin += (base + i) * c;
			}
			sq = block_sum_u64(sq, red);
			if (is_apen)
				lin = block_sum_u64(lin, red);
			if (threadIdx.x == 0) {
				if (is_serial)
					atomicAdd((unsigned long long *) (o + P.st_off[STSGPU_TEST_SERIAL] + (sm - lvl)), sq);
				if (is_apen) {
					int r = lvl - am;	/* 1: level m+1, 0: level m */
					atomicAdd((unsigned long long *) (o + P.st_off[STSGPU_TEST_APEN] + 2 + r), sq);
					atomicAdd((unsigned long long *) (o + P.st_off[STSGPU_TEST_APEN] + 4 + r), lin);
				}
			}
			if (is_apen) {
				uint32_t *dst = apen_hist + s * ((size_t) 3 << am) + (lvl == am + 1 ? 0 : ((size_t) 2 << am))
						+ ((size_t) slice << b);
				for (int i = threadIdx.x; i < nb; i += HIST_THREADS)
					dst[i] = hist[i];
			}
		}
		if (lvl == lowest)
			break;
		/* fold to the next level in place: new[i] = old[2i] + old[2i+1] */
		const int half = nb >> 1;
		for (int c0 = 0; c0 < max(half, 1); c0 += HIST_THREADS) {
			int i = c0 + threadIdx.x;
			uint32_t v = 0;
			if (i < half)
				v = hist[2 * i] + hist[2 * i + 1];
			__syncthreads();
			if (i < half)
				hist[i] = v;
			__syncthreads();
		}
	}
}

/*
 * apen_hist layout per stream: [2^(apen_m+1) bins of level apen_m+1][2^apen_m bins of level apen_m]
 */
__global__ void __launch_bounds__(HIST_THREADS)
cyclic_hist_kernel(DevParams P, const uint32_t *__restrict__ in, long long *__restrict__ st,
		   uint32_t *__restrict__ apen_hist, uint32_t *__restrict__ debug_hist, int debug_bits,
		   const uint32_t *__restrict__ only_flagged)
{
	extern __shared__ uint32_t hist[];
	__shared__ unsigned long long red[HIST_THREADS / 32];
	const int64_t s = blockIdx.x;
	if (only_flagged && !only_flagged[s])
		return;					/* the packed kernel already did this stream */
	const uint32_t slice = blockIdx.y;
	const int H = debug_hist ? debug_bits : P.hist_bits;
	const int sb = debug_hist ? (H > 15 ? H - 15 : 0) : P.hist_slice_bits;
	const int B = H - sb;					/* local histogram bits */
	const uint32_t lmask = (1u << B) - 1u;
	for (int i = threadIdx.x; i < (1 << B); i += HIST_THREADS)
		hist[i] = 0;
	__syncthreads();
	const uint32_t *__restrict__ w = in + s * P.pitch_words;
	const int64_t n = P.n;
	const int64_t groups = (n + 31) >> 5;
	for (int64_t g = threadIdx.x; g < groups; g += HIST_THREADS) {
		const int64_t j0 = g * 32;
		uint32_t cur, nxt;
		if (j0 + 64 <= n) {
			cur = ldw(w, g);
			nxt = ldw(w, g + 1);
		} else {
			cur = cyc32(w, j0, n);
			nxt = cyc32(w, j0 + 32, n);
		}
		int cnt = (n - j0 >= 32) ? 32 : (int) (n - j0);
		if (cnt == 32) {
#pragma unroll 8
			for (int p = 0; p < 32; p++) {
				uint32_t v = __funnelshift_l(nxt, cur, p) >> (32 - H);
				if ((v >> B) == slice)
					atomicAdd(&hist[v & lmask], 1u);
			}
		} else {
			for (int p = 0; p < cnt; p++) {
				uint32_t v = __funnelshift_l(nxt, cur, p) >> (32 - H);
				if ((v >> B) == slice)
					atomicAdd(&hist[v & lmask], 1u);
			}
		}
	}
	__syncthreads();
	if (debug_hist) {
		uint32_t *o = debug_hist + ((size_t) slice << B);
		for (int i = threadIdx.x; i < (1 << B); i += HIST_THREADS)
			o[i] = hist[i];
		return;
	}

	emit_levels(P, hist, red, H, sb, slice, s, st, apen_hist);
}

/*
 * Fast path for H <= 16: the whole 2^H-bin histogram of one stream in ONE CTA with 16-bit counters,
 * two bins per shared-memory word (bin 2x in the low half, 2x+1 in the high half), one atomic per
 * window and no slice filter.  A counter can only overflow for wildly non-random input; every
 * increment adds exactly one to the sum of all 16-bit fields unless a field wraps, which lowers it by
 * 65535 or 65536, so "field sum == n" proves the counts exact.  Otherwise the stream is flagged and
 * redone by the sliced 32-bit kernel above (never silently wrong).  Folding to H-1 bits is
 * word = low + high, after which the counters are 32-bit and the common level code takes over.
 */
__global__ void __launch_bounds__(HIST_THREADS)
cyclic_hist16_kernel(DevParams P, const uint32_t *__restrict__ in, long long *__restrict__ st,
		     uint32_t *__restrict__ apen_hist, uint32_t *__restrict__ flags)
{
	extern __shared__ uint32_t hist[];
	__shared__ unsigned long long red[HIST_THREADS / 32];
	__shared__ int ok_s;
	const int64_t s = blockIdx.x;
	const int H = P.hist_bits;				/* 2 <= H <= 16 */
	const int words = 1 << (H - 1);
	for (int i = threadIdx.x; i < words; i += HIST_THREADS)
		hist[i] = 0;
	__syncthreads();
	const uint32_t *__restrict__ w = in + s * P.pitch_words;
	const int64_t n = P.n;
	const int64_t groups = (n + 31) >> 5;
	for (int64_t g = threadIdx.x; g < groups; g += HIST_THREADS) {
		const int64_t j0 = g * 32;
		uint32_t cur, nxt;
		if (j0 + 64 <= n) {
			cur = ldw(w, g);
			nxt = ldw(w, g + 1);
		} else {
			cur = cyc32(w, j0, n);
			nxt = cyc32(w, j0 + 32, n);
		}
		const int cnt = (n - j0 >= 32) ? 32 : (int) (n - j0);
		if (cnt == 32) {			/* the common case without a test per window */
#pragma unroll
			for (int p = 0; p < 32; p++) {
				const uint32_t v = __funnelshift_l(nxt, cur, p) >> (32 - H);
				atomicAdd(&hist[v >> 1], 1u << ((v & 1u) << 4));
			}
		} else {
			for (int p = 0; p < cnt; p++) {
				const uint32_t v = __funnelshift_l(nxt, cur, p) >> (32 - H);
				atomicAdd(&hist[v >> 1], 1u << ((v & 1u) << 4));
			}
		}
	}
	__syncthreads();
	{
		unsigned long long tot = 0;
		for (int i = threadIdx.x; i < words; i += HIST_THREADS)
			tot += (hist[i] & 0xFFFFu) + (hist[i] >> 16);
		tot = block_sum_u64(tot, red);
		if (threadIdx.x == 0) {
			ok_s = (tot == (unsigned long long) n) ? 1 : 0;
			flags[s] = ok_s ? 0u : 1u;
		}
		__syncthreads();
		if (!ok_s)
			return;
	}
	const int sm = (P.enabled & (1u << STSGPU_TEST_SERIAL)) ? P.serial_m : -100;
	const int am = (P.enabled & (1u << STSGPU_TEST_APEN)) ? P.apen_m : -100;
	int lowest = H;
	if (sm > 0) lowest = min(lowest, max(sm - 2, 1));
	if (am > 0) lowest = min(lowest, am);
	long long *o = st + s * P.nst;
	{	/* level H from the packed fields */
		const bool is_serial = (H == sm || H == sm - 1 || H == sm - 2);
		const bool is_apen = (H == am + 1 || H == am);
		if (is_serial || is_apen) {
			unsigned long long sq = 0, lin = 0;
			for (int i = threadIdx.x; i < words; i += HIST_THREADS) {
				const unsigned long long c0 = hist[i] & 0xFFFFu, c1 = hist[i] >> 16;
				sq += c0 * c0 + c1 * c1;
				lin += (unsigned long long) (2 * i) * c0 + (unsigned long long) (2 * i + 1) * c1;
			}
			sq = block_sum_u64(sq, red);
			if (is_apen)
				lin = block_sum_u64(lin, red);
			if (threadIdx.x == 0) {
				if (is_serial)
					atomicAdd((unsigned long long *) (o + P.st_off[STSGPU_TEST_SERIAL] + (sm - H)), sq);
				if (is_apen) {
					const int r = H - am;
					atomicAdd((unsigned long long *) (o + P.st_off[STSGPU_TEST_APEN] + 2 + r), sq);
					atomicAdd((unsigned long long *) (o + P.st_off[STSGPU_TEST_APEN] + 4 + r), lin);
				}
			}
			if (is_apen) {
				uint32_t *dst = apen_hist + s * ((size_t) 3 << am) + (H == am + 1 ? 0 : ((size_t) 2 << am));
				for (int i = threadIdx.x; i < words; i += HIST_THREADS) {
					dst[2 * i] = hist[i] & 0xFFFFu;
					dst[2 * i + 1] = hist[i] >> 16;
				}
			}
		}
	}
	if (H == lowest)
		return;
	__syncthreads();
	for (int i = threadIdx.x; i < words; i += HIST_THREADS)
		hist[i] = (hist[i] & 0xFFFFu) + (hist[i] >> 16);
	__syncthreads();
	emit_levels(P, hist, red, H - 1, 0, 0u, s, st, apen_hist);
}

/*
 * Serial at m = 16 with ApEn below 15 (the defaults): half the atomics.  ONE atomic counts the windows at TWO
 * positions: the histogram is over the 17-bit windows at the EVEN positions, 2^17 8-bit counters in the same 128 KiB
 * (a 17-bit window at p holds the 16-bit windows at p and p + 1; n is even, the stream cyclic).  Level 16 is needed
 * only for its sum of squares, which is taken on the fly - count16[x] = W[2x] + W[2x+1] + W[x] + W[2^16 + x], four bins
 * per thread and step from four word loads - and level 15 - count15[y] = the four bytes of word y + the two bytes of
 * halfwords y and 2^15 + y - goes through registers (32 bins per thread) into 32-bit counters in the same memory,
 * where the common level code takes over.  The byte counters hold 4 on average; as in the 16-bit kernel a wrapped
 * field lowers the sum of all fields, so "sum == n / 2" proves the counts, and any other stream is flagged and redone by
 * the sliced 32-bit kernel.
 */
__device__ __forceinline__ uint32_t bytesum4(uint32_t w)
{
	const uint32_t t = (w & 0x00FF00FFu) + ((w >> 8) & 0x00FF00FFu);
	return (t & 0xFFFFu) + (t >> 16);
}

#ifndef HIST17_MAXREG
#define HIST17_MAXREG 48
#endif
__global__ void __maxnreg__(HIST17_MAXREG)
cyclic_hist17w_kernel(DevParams P, const uint32_t *__restrict__ in, long long *__restrict__ st,
		      uint32_t *__restrict__ apen_hist, uint32_t *__restrict__ flags)
{
	extern __shared__ uint32_t hist[];			/* 2^15 words: first 2^17 byte counters, then 2^15 counters of level 15 */
	__shared__ unsigned long long red[HIST_THREADS / 32];
	__shared__ int ok_s;
	const int64_t s = blockIdx.x;
	const int words = 1 << 15;
	for (int i = threadIdx.x; i < words; i += HIST_THREADS)
		hist[i] = 0;
	__syncthreads();
	const uint32_t *__restrict__ w = in + s * P.pitch_words;
	const int64_t n = P.n;
	const int64_t groups = (n + 31) >> 5;
	for (int64_t g = threadIdx.x; g < groups; g += HIST_THREADS) {
		const int64_t j0 = g * 32;
		uint32_t cur, nxt;
		if (j0 + 64 <= n) {
			cur = ldw(w, g);
			nxt = ldw(w, g + 1);
		} else {
			cur = cyc32(w, j0, n);
			nxt = cyc32(w, j0 + 32, n);
		}
		const int cnt = (n - j0 >= 32) ? 32 : (int) (n - j0);	/* even: n is a multiple of 8 */
		if (cnt == 32) {
#pragma unroll
			for (int p = 0; p < 32; p += 2) {
				const uint32_t v = __funnelshift_l(nxt, cur, p) >> 15;
				atomicAdd(&hist[v >> 2], 1u << ((v & 3u) << 3));
			}
		} else {
			for (int p = 0; p < cnt; p += 2) {
				const uint32_t v = __funnelshift_l(nxt, cur, p) >> 15;
				atomicAdd(&hist[v >> 2], 1u << ((v & 3u) << 3));
			}
		}
	}
	__syncthreads();
	{
		unsigned long long tot = 0;
		for (int i = threadIdx.x; i < words; i += HIST_THREADS)
			tot += bytesum4(hist[i]);
		tot = block_sum_u64(tot, red);
		if (threadIdx.x == 0) {
			ok_s = (tot == (unsigned long long) (n >> 1)) ? 1 : 0;
			flags[s] = ok_s ? 0u : 1u;
		}
		__syncthreads();
		if (!ok_s)
			return;
	}
	long long *o = st + s * P.nst;
	{	/* level 16: only its sum of squares (serial.c:414-421 at m) */
		unsigned long long sq = 0;
		for (int q = threadIdx.x; q < (1 << 14); q += HIST_THREADS) {
			const uint32_t h0 = hist[2 * q], h1 = hist[2 * q + 1];	/* W[2x], W[2x+1] for x = 4q .. 4q + 3 */
			const uint32_t b0 = hist[q], b1 = hist[(1 << 14) + q];	/* W[x], W[2^16 + x] */
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const uint32_t hw = ((j < 2 ? h0 : h1) >> (16 * (j & 1))) & 0xFFFFu;
				const unsigned long long c = (hw & 0xFFu) + (hw >> 8) + ((b0 >> (8 * j)) & 0xFFu) + ((b1 >> (8 * j)) & 0xFFu);
				sq += c * c;
			}
		}
		sq = block_sum_u64(sq, red);
		if (threadIdx.x == 0)
			atomicAdd((unsigned long long *) (o + P.st_off[STSGPU_TEST_SERIAL]), sq);
	}
	/* level 15 through registers: bins threadIdx.x + 1024 k, two 16-bit counts per register */
	uint32_t c15[16];
#pragma unroll
	for (int k = 0; k < 32; k++) {
		const int y = threadIdx.x + k * HIST_THREADS;
		const uint32_t ha = (hist[y >> 1] >> (16 * (y & 1))) & 0xFFFFu;
		const uint32_t hb = (hist[(1 << 14) + (y >> 1)] >> (16 * (y & 1))) & 0xFFFFu;
		const uint32_t c = bytesum4(hist[y]) + (ha & 0xFFu) + (ha >> 8) + (hb & 0xFFu) + (hb >> 8);
		if (k & 1)
			c15[k >> 1] |= c << 16;
		else
			c15[k >> 1] = c;
	}
	__syncthreads();
#pragma unroll
	for (int k = 0; k < 32; k++)
		hist[threadIdx.x + k * HIST_THREADS] = (k & 1) ? (c15[k >> 1] >> 16) : (c15[k >> 1] & 0xFFFFu);
	__syncthreads();
	emit_levels(P, hist, red, 15, 0, 0u, s, st, apen_hist);
}

cudaError_t launch_pattern(const DevParams &P, const LaunchCtx &L, uint32_t *d_apen_hist, uint32_t *d_flags)
{
	if (!(P.enabled & ((1u << STSGPU_TEST_SERIAL) | (1u << STSGPU_TEST_APEN))))
		return cudaSuccess;
	const int B = P.hist_bits - P.hist_slice_bits;
	size_t smem = (size_t) sizeof(uint32_t) << B;
	cudaError_t e = cudaFuncSetAttribute(cyclic_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
	if (e != cudaSuccess)
		return e;
	const uint32_t *only_flagged = nullptr;
	if (P.hist_bits >= 2 && P.hist_bits <= 16 && d_flags != nullptr) {
		e = cudaFuncSetAttribute(cyclic_hist16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
		if (e != cudaSuccess)
			return e;
		const bool serial = (P.enabled & (1u << STSGPU_TEST_SERIAL)) != 0, apen = (P.enabled & (1u << STSGPU_TEST_APEN)) != 0;
		static const bool wide_ok = !(getenv("STSGPU_HIST_WIDE") && atoi(getenv("STSGPU_HIST_WIDE")) == 0);
		if (wide_ok && P.hist_bits == 16 && serial && P.serial_m == 16 && (!apen || P.apen_m + 1 < 16)) {
			e = cudaFuncSetAttribute(cyclic_hist17w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
			if (e != cudaSuccess)
				return e;
			cyclic_hist17w_kernel<<<(unsigned) L.nstreams, HIST_THREADS, 128 * 1024, L.stream>>>(P, L.d_in, L.d_st, d_apen_hist, d_flags);
		} else {
			cyclic_hist16_kernel<<<(unsigned) L.nstreams, HIST_THREADS, (size_t) sizeof(uint32_t) << (P.hist_bits - 1), L.stream>>>(
				P, L.d_in, L.d_st, d_apen_hist, d_flags);
		}
		only_flagged = d_flags;
	}
	dim3 grid((unsigned) L.nstreams, 1u << P.hist_slice_bits);
	cyclic_hist_kernel<<<grid, HIST_THREADS, smem, L.stream>>>(P, L.d_in, L.d_st, d_apen_hist, nullptr, 0, only_flagged);
	return cudaGetLastError();
}

/* parity helper: full cyclic histogram of one stream at `bits` into d_out (2^bits counters) */
cudaError_t launch_debug_histogram(const DevParams &P, const LaunchCtx &L, int64_t stream_index, int bits,
				   uint32_t *d_out)
{
	const int sb = bits > 15 ? bits - 15 : 0;
	size_t smem = (size_t) sizeof(uint32_t) << (bits - sb);
	cudaError_t e = cudaFuncSetAttribute(cyclic_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
	if (e != cudaSuccess)
		return e;
	dim3 grid(1, 1u << sb);
	cyclic_hist_kernel<<<grid, HIST_THREADS, smem, L.stream>>>(P, L.d_in + stream_index * P.pitch_words, nullptr,
								  nullptr, d_out, bits, nullptr);
	return cudaGetLastError();
}
