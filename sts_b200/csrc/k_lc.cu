/*
 * k_lc.cu - linear complexity test: Berlekamp-Massey over GF(2) per M-bit block
 * (linearComplexity.c:283-377).
 *
 * One block per THREAD, bit-packed 32 coefficients per register.  The recurrence is kept in
 * "sequence-aligned reversed" form so that only one array moves per step:
 *     Cr_N[j] = c_{N-j}           the connection polynomial reversed and aligned under s_0..s_N
 *     d_N     = parity(popc(Cr_N & S))                         (linearComplexity.c:312-316)
 *     on d_N = 1:  Cr_N ^= Br   where Br = Cr_m saved at the last length change stays fixed
 *                  (x^{N-m} b(x) reversed and aligned at N is exactly Cr_m)   (:322-327)
 *                  if L <= N/2: L = N + 1 - L, m = N, Br = old Cr_N            (:331-336)
 *     N -> N + 1:  Cr shifts by one position
 * All updates are branch-free masks, so the 32 blocks of a warp stay converged.  Positions are
 * stored with an offset of one (index 0 = position -1) because the initial b(x) = 1 with m = -1
 * sits at position -1.  The statistic per block is the class of L, looked up in a table computed
 * on the host with the reference's own expression, including its `1 << M` quirk (:356-377).
 *
 * Algorithmic bytes: n/8 per stream.  Integer work: M steps x ~4.5 word-ops x ceil((M+1)/32) words.
 *
 * Tried in round 2 and measured slower (2.94 against 2.57 ms per 1024 streams): the per-step advance of C as one wide
 * multiply-add per word (c_new * 2^31 + carry << 32, the multiplier a run-time value so that it stays an IMAD.WIDE on
 * the FMA pipe) to take the shift off the ALU pipe that bounds the kernel; SASS confirmed SHF 2569 -> 0 and IMAD
 * +5292 (the wide multiply and a move per word), but two FMA-pipe instructions cost more than the one SHF they replace.
 */
#include <stdlib.h>
#include "common.cuh"

#ifndef LC_THREADS
#define LC_THREADS 128
#endif

template <int W>
__global__ void __launch_bounds__(LC_THREADS)
lc_kernel_reg(DevParams P, const uint32_t *__restrict__ in, const uint8_t *__restrict__ class_lut,
	      long long *__restrict__ st)
{
	__shared__ unsigned int cls[7];
	const int64_t s = blockIdx.x;
	if (threadIdx.x < 7)
		cls[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t *__restrict__ w = in + s * P.pitch_words;
	const int M = (int) P.lc_M;
	const int64_t blk = (int64_t) blockIdx.y * LC_THREADS + threadIdx.x;
	if (blk < P.lc_N) {
		const int64_t start = blk * M;
		uint32_t S[W], C[W], B[W];
		/* index k of the arrays <-> position k - 1; index 0 (position -1) is masked off in S */
#pragma unroll
		for (int i = 0; i < W; i++) {
			uint32_t v = (i == 0) ? (get32(w, start) >> 1) : get32(w, start + 32 * i - 1);
			int valid = M + 1 - 32 * i;			/* indices still inside the block */
			if (valid < 32)
				v &= mask_below((uint32_t) (valid > 0 ? valid : 0));
			S[i] = v;
			C[i] = 0;
			B[i] = 0;
		}
		C[0] = 0x40000000u;	/* c_0 at position 0  (index 1) */
		B[0] = 0x80000000u;	/* b_0 at position -1 (index 0) */
		int L = 0;
		for (int N = 0; N < M; N++) {
			uint32_t acc = 0;
#pragma unroll
			for (int i = 0; i < W; i++)
				acc ^= C[i] & S[i];
			const uint32_t d = (uint32_t) __popc(acc) & 1u;
			const uint32_t dm = 0u - d;
			const uint32_t um = (2 * L <= N) ? dm : 0u;
			uint32_t prev = 0;
#pragma unroll
			for (int i = 0; i < W; i++) {
				const uint32_t c_old = C[i];
				const uint32_t c_new = c_old ^ (B[i] & dm);
				B[i] = (c_old & um) | (B[i] & ~um);
				C[i] = (c_new >> 1) | (prev << 31);	/* advance to N + 1 */
				prev = c_new;
			}
			L = um ? (N + 1 - L) : L;
		}
		atomicAdd(&cls[class_lut[L]], 1u);
	}
	__syncthreads();
	if (threadIdx.x < 7 && cls[threadIdx.x])
		atomicAdd((unsigned long long *) (st + s * P.nst + P.st_off[STSGPU_TEST_LINEARCOMPLEXITY] + threadIdx.x),
			  (unsigned long long) cls[threadIdx.x]);
}

/*
 * Windowed variant (M <= 511): at step N only the words between the lowest and the highest non-zero
 * coefficient carry information.  C (= c reversed, aligned at N) is non-zero on positions [N - L, N]
 * and B (= C at the last length change, fixed in place) on positions [L - 1 ..]; for every step N'
 * of a 32-step phase starting at N0 both lie above  min(N0 - L_N0, L_N0 - 1)  because L never shrinks
 * and a length change at N sets L to N + 1 - L_old <= N' - L_N0.  The phase index fixes the top word
 * at compile time, the bottom word is the warp-wide minimum of that bound, and the pair selects one
 * of ~150 straight-line step bodies, so the registers stay statically indexed.  For random input
 * L tracks N/2 and the window is about N/64 + 2 words instead of 16.
 */
template <int W, int LO, int HI>
__device__ __forceinline__ void lc_steps(const uint32_t (&S)[W], uint32_t (&C)[W], uint32_t (&B)[W], int &L, int N0, int N1)
{
	for (int N = N0; N < N1; N++) {
		uint32_t acc = 0;
#pragma unroll
		for (int i = LO; i <= HI; i++)
			acc ^= C[i] & S[i];
		const uint32_t d = (uint32_t) __popc(acc) & 1u;
		const uint32_t dm = 0u - d;
		const uint32_t um = (2 * L <= N) ? dm : 0u;
		uint32_t prev = 0;
#pragma unroll
		for (int i = LO; i <= HI; i++) {
			const uint32_t c_old = C[i];
			const uint32_t c_new = c_old ^ (B[i] & dm);
			B[i] = (c_old & um) | (B[i] & ~um);
			C[i] = (c_new >> 1) | (prev << 31);	/* advance to N + 1 */
			prev = c_new;
		}
		L = um ? (N + 1 - L) : L;
	}
}

template <int W, int LO, int HI>
__device__ __forceinline__ void lc_dispatch(int lo, const uint32_t (&S)[W], uint32_t (&C)[W], uint32_t (&B)[W], int &L, int N0,
					    int N1)
{
	if constexpr (LO >= HI) {
		lc_steps<W, HI, HI>(S, C, B, L, N0, N1);
	} else {
		if (lo <= LO)
			lc_steps<W, LO, HI>(S, C, B, L, N0, N1);
		else
			lc_dispatch<W, LO + 1, HI>(lo, S, C, B, L, N0, N1);
	}
}

template <int W, int P>
__device__ __forceinline__ void lc_phases(const uint32_t (&S)[W], uint32_t (&C)[W], uint32_t (&B)[W], int &L, int M)
{
	if constexpr (P < W) {
		const int N0 = 32 * P;
		if (N0 < M) {						/* M is uniform: no divergence */
			constexpr int HI = (P + 1 < W) ? P + 1 : W - 1;
			/* lowest index (position + 1) that can be non-zero during this phase */
			int lo_idx = min(N0 - L + 1, L);
			lo_idx = __reduce_min_sync(0xffffffffu, lo_idx);
			lc_dispatch<W, 0, HI>(lo_idx >> 5, S, C, B, L, N0, min(M, N0 + 32));
			lc_phases<W, P + 1>(S, C, B, L, M);
		}
	}
}

/*
 * Grid-stride over the (stream, 128-block group) work items, launched as a resident "background" grid
 * of STSGPU_LC_BG (default 2) CTAs per SM: this kernel needs only the integer ALU pipe and 8K
 * registers per CTA, the DFT needs the FP64 pipe, the LSU and shared memory, and with two small CTAs
 * per SM (plus the padded shared-memory request, common.cuh) the two share every SM instead of taking
 * turns.  Alone the kernel is ~15 % slower this way (8 warps per SM); the batch is ~5 % faster.
 */
template <int W>
__global__ void __launch_bounds__(LC_THREADS, 1024 / LC_THREADS)
lc_kernel_win(DevParams P, const uint32_t *__restrict__ in, const uint8_t *__restrict__ class_lut,
	      long long *__restrict__ st, int64_t nstreams, int groups_per_stream)
{
	__shared__ unsigned int cls[7];
	const int M = (int) P.lc_M;
	for (int64_t item = blockIdx.x; item < nstreams * groups_per_stream; item += gridDim.x) {
		const int64_t s = item / groups_per_stream;
		const int by = (int) (item % groups_per_stream);
		__syncthreads();				/* cls[] of the previous item has been flushed */
		if (threadIdx.x < 7)
			cls[threadIdx.x] = 0;
		__syncthreads();
		const uint32_t *__restrict__ w = in + s * P.pitch_words;
		/* whole warps stay alive (the window bound is a warp reduction); surplus lanes redo the last block */
		const int64_t blk_raw = (int64_t) by * LC_THREADS + threadIdx.x;
		const int64_t blk = min(blk_raw, P.lc_N - 1);
		{
			const int64_t start = blk * M;
			uint32_t S[W], C[W], B[W];
#pragma unroll
			for (int i = 0; i < W; i++) {
				uint32_t v = (i == 0) ? (get32(w, start) >> 1) : get32(w, start + 32 * i - 1);
				int valid = M + 1 - 32 * i;
				if (valid < 32)
					v &= mask_below((uint32_t) (valid > 0 ? valid : 0));
				S[i] = v;
				C[i] = 0;
				B[i] = 0;
			}
			C[0] = 0x40000000u;
			B[0] = 0x80000000u;
			int L = 0;
			lc_phases<W, 0>(S, C, B, L, M);
			if (blk_raw < P.lc_N)
				atomicAdd(&cls[class_lut[L]], 1u);
		}
		__syncthreads();
		if (threadIdx.x < 7 && cls[threadIdx.x])
			atomicAdd((unsigned long long *) (st + s * P.nst + P.st_off[STSGPU_TEST_LINEARCOMPLEXITY] + threadIdx.x),
				  (unsigned long long) cls[threadIdx.x]);
	}
}

/*
 * Long blocks (M > 1022, up to the reference's limit of 5000): one block per WARP.  Word i of the three arrays
 * lives in lane i mod 32, register slot i / 32 (cyclic), so that the window of live words - the same bound as in
 * the per-thread kernel: indices [min(N0 - L + 1, L), N + 2] - covers whole slots: for random input about
 * N / 2048 .. N / 1024, i.e. 1-3 of the 5 slots of M = 5000 instead of all of them.  With one block per warp the
 * discrepancy and the length test are warp-uniform, so the updates are real (uniform) branches instead of
 * masks: a step is one AND-XOR, one rotate-shuffle and one funnel shift per live slot plus one redux.sync.
 * The slot range is re-derived every 256 steps and selects one of WPL (WPL + 1) / 2 unrolled bodies.
 * (The per-thread local-memory kernel below took 83 % of a BASELINE config 4 batch.)
 *
 * Tried in round 2 and measured slower (200 against 120 ms per 428 streams of config 4): one block per group of EIGHT
 * lanes, four blocks per warp, 20 slots per lane.  More live slots per lane amortise the per-step overhead, but the
 * four blocks of a warp take different paths, so every update becomes a mask operation that runs on every step
 * (here the discrepancy is warp-uniform and half of the steps skip the update), and the rotate-shuffle per live slot
 * - five to ten per step instead of one to three - saturates the shuffle path.
 */
template <int WPL, int SLO, int SHI>
__device__ __forceinline__ void lc_warp_steps(const uint32_t (&S)[WPL], uint32_t (&C)[WPL], uint32_t (&B)[WPL], int &L, int Na,
					      int Nb, int lane)
{
	const int src = (lane + 31) & 31;
	for (int N = Na; N < Nb; N++) {
		uint32_t acc = 0;
#pragma unroll
		for (int j = SLO; j <= SHI; j++)
			acc ^= C[j] & S[j];
		const uint32_t d = (uint32_t) __popc(__reduce_xor_sync(0xffffffffu, acc)) & 1u;
		if (d) {				/* linearComplexity.c:322-336 */
			if (2 * L <= N) {
#pragma unroll
				for (int j = SLO; j <= SHI; j++) {
					const uint32_t c_old = C[j];
					C[j] = c_old ^ B[j];
					B[j] = c_old;
				}
				L = N + 1 - L;
			} else {
#pragma unroll
				for (int j = SLO; j <= SHI; j++)
					C[j] ^= B[j];
			}
		}
		/* advance to N + 1: word i takes the bit that leaves word i - 1 (lane - 1; for lane 0 lane 31 of the slot below) */
		uint32_t below = 0;
#pragma unroll
		for (int j = SLO; j <= SHI; j++) {
			const uint32_t c = C[j];
			const uint32_t rot = __shfl_sync(0xffffffffu, c, src);
			const uint32_t carry = (lane == 0) ? below : rot;
			C[j] = (c >> 1) | (carry << 31);
			below = rot;
		}
	}
}

template <int WPL, int SLO, int SHI>
__device__ __forceinline__ void lc_warp_dispatch(int slo, int shi, const uint32_t (&S)[WPL], uint32_t (&C)[WPL], uint32_t (&B)[WPL],
						 int &L, int Na, int Nb, int lane)
{
	if (slo == SLO && shi == SHI) {
		lc_warp_steps<WPL, SLO, SHI>(S, C, B, L, Na, Nb, lane);
	} else if constexpr (SLO < SHI) {
		lc_warp_dispatch<WPL, SLO + 1, SHI>(slo, shi, S, C, B, L, Na, Nb, lane);
	} else if constexpr (SHI + 1 < WPL) {
		lc_warp_dispatch<WPL, 0, SHI + 1>(slo, shi, S, C, B, L, Na, Nb, lane);
	}
}

template <int WPL>
__global__ void __launch_bounds__(LC_THREADS)
lc_kernel_warp(DevParams P, const uint32_t *__restrict__ in, const uint8_t *__restrict__ class_lut,
	       long long *__restrict__ st)
{
	__shared__ unsigned int cls[7];
	const int64_t s = blockIdx.x;
	const int lane = threadIdx.x & 31;
	if (threadIdx.x < 7)
		cls[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t *__restrict__ w = in + s * P.pitch_words;
	const int M = (int) P.lc_M;
	const int64_t blk = (int64_t) blockIdx.y * (LC_THREADS / 32) + (threadIdx.x >> 5);
	if (blk < P.lc_N) {				/* warp-uniform */
		const int64_t start = blk * M;
		uint32_t S[WPL], C[WPL], B[WPL];
#pragma unroll
		for (int j = 0; j < WPL; j++) {
			const int i = j * 32 + lane;		/* word index; array index k <-> position k - 1 */
			uint32_t v = 0;
			const int valid = M + 1 - 32 * i;	/* indices still inside the block */
			if (valid > 0) {
				v = (i == 0) ? (get32(w, start) >> 1) : get32(w, start + 32 * (int64_t) i - 1);
				if (valid < 32)
					v &= mask_below((uint32_t) valid);
			}
			S[j] = v;
			C[j] = 0;
			B[j] = 0;
		}
		if (lane == 0) {
			C[0] = 0x40000000u;	/* c_0 at position 0  (index 1) */
			B[0] = 0x80000000u;	/* b_0 at position -1 (index 0) */
		}
		int L = 0;
		/* phases end where the top index N + 2 of the advanced array reaches a multiple of 256 */
		for (int Na = 0; Na < M;) {
			const int Nb = min(M, ((Na + 2) / 256 + 1) * 256 - 2);
			const int slo = min(Na - L + 1, L) >> 10;	/* lowest live index, as in lc_phases */
			const int shi = min((Nb + 1) >> 10, WPL - 1);
			lc_warp_dispatch<WPL, 0, 0>(slo, shi, S, C, B, L, Na, Nb, lane);
			Na = Nb;
		}
		if (lane == 0)
			atomicAdd(&cls[class_lut[L]], 1u);
	}
	__syncthreads();
	if (threadIdx.x < 7 && cls[threadIdx.x])
		atomicAdd((unsigned long long *) (st + s * P.nst + P.st_off[STSGPU_TEST_LINEARCOMPLEXITY] + threadIdx.x),
			  (unsigned long long) cls[threadIdx.x]);
}

/* any M up to 5000: the same recurrence with the arrays in (L1-resident) local memory */
#define LC_MAXW 160
__global__ void __launch_bounds__(LC_THREADS)
lc_kernel_generic(DevParams P, const uint32_t *__restrict__ in, const uint8_t *__restrict__ class_lut,
		  long long *__restrict__ st)
{
	__shared__ unsigned int cls[7];
	const int64_t s = blockIdx.x;
	if (threadIdx.x < 7)
		cls[threadIdx.x] = 0;
	__syncthreads();
	const uint32_t *__restrict__ w = in + s * P.pitch_words;
	const int M = (int) P.lc_M;
	const int W = (M + 1 + 31) >> 5;
	const int64_t blk = (int64_t) blockIdx.y * LC_THREADS + threadIdx.x;
	if (blk < P.lc_N) {
		const int64_t start = blk * M;
		uint32_t S[LC_MAXW], C[LC_MAXW], B[LC_MAXW];
		for (int i = 0; i < W; i++) {
			uint32_t v = (i == 0) ? (get32(w, start) >> 1) : get32(w, start + 32 * i - 1);
			int valid = M + 1 - 32 * i;
			if (valid < 32)
				v &= mask_below((uint32_t) (valid > 0 ? valid : 0));
			S[i] = v;
			C[i] = 0;
			B[i] = 0;
		}
		C[0] = 0x40000000u;
		B[0] = 0x80000000u;
		int L = 0;
		for (int N = 0; N < M; N++) {
			/* only words up to position N carry coefficients */
			const int hi = min(W - 1, (N + 2) >> 5);
			uint32_t acc = 0;
			for (int i = 0; i <= hi; i++)
				acc ^= C[i] & S[i];
			const uint32_t d = (uint32_t) __popc(acc) & 1u;
			const uint32_t dm = 0u - d;
			const uint32_t um = (2 * L <= N) ? dm : 0u;
			uint32_t prev = 0;
			for (int i = 0; i <= hi; i++) {
				const uint32_t c_old = C[i];
				const uint32_t c_new = c_old ^ (B[i] & dm);
				B[i] = (c_old & um) | (B[i] & ~um);
				C[i] = (c_new >> 1) | (prev << 31);
				prev = c_new;
			}
			if (hi + 1 < W)
				C[hi + 1] |= prev << 31;	/* bit shifted across the active boundary */
			L = um ? (N + 1 - L) : L;
		}
		atomicAdd(&cls[class_lut[L]], 1u);
	}
	__syncthreads();
	if (threadIdx.x < 7 && cls[threadIdx.x])
		atomicAdd((unsigned long long *) (st + s * P.nst + P.st_off[STSGPU_TEST_LINEARCOMPLEXITY] + threadIdx.x),
			  (unsigned long long) cls[threadIdx.x]);
}

cudaError_t launch_lc(const DevParams &P, const LaunchCtx &L, const uint8_t *d_class_lut)
{
	if (!(P.enabled & (1u << STSGPU_TEST_LINEARCOMPLEXITY)))
		return cudaSuccess;
	dim3 grid((unsigned) L.nstreams, (unsigned) ((P.lc_N + LC_THREADS - 1) / LC_THREADS));
	const int W = (int) ((P.lc_M + 1 + 31) >> 5);
	if (W <= 16) {
		static int bg = -1;
		if (bg < 0)
			bg = getenv("STSGPU_LC_BG") ? atoi(getenv("STSGPU_LC_BG")) : 2;
		const int64_t items = (int64_t) grid.x * grid.y;
		int64_t ctas = bg > 0 ? (int64_t) bg * L.num_sms : items;
		if (ctas > items) ctas = items;
		const size_t pad = sts_pad_smem(0, 24 * 1024);	/* a large L1/shared split: common.cuh */
		STS_PAD_KERNEL(lc_kernel_win<16>, pad);
		lc_kernel_win<16><<<(unsigned) ctas, LC_THREADS, pad, L.stream>>>(P, L.d_in, d_class_lut, L.d_st, L.nstreams, (int) grid.y);
	}
	else if (W <= 24) {
		lc_kernel_reg<24><<<grid, LC_THREADS, 0, L.stream>>>(P, L.d_in, d_class_lut, L.d_st);
	} else if (W <= 32) {
		lc_kernel_reg<32><<<grid, LC_THREADS, 0, L.stream>>>(P, L.d_in, d_class_lut, L.d_st);
	} else {
		/* one block per warp, ceil(W / 32) words per lane (W <= 157 at M = 5000) */
		dim3 gw((unsigned) L.nstreams, (unsigned) ((P.lc_N + LC_THREADS / 32 - 1) / (LC_THREADS / 32)));
		const int wpl = (W + 31) / 32;
		if (wpl <= 2)
			lc_kernel_warp<2><<<gw, LC_THREADS, 0, L.stream>>>(P, L.d_in, d_class_lut, L.d_st);
		else if (wpl == 3)
			lc_kernel_warp<3><<<gw, LC_THREADS, 0, L.stream>>>(P, L.d_in, d_class_lut, L.d_st);
		else if (wpl == 4)
			lc_kernel_warp<4><<<gw, LC_THREADS, 0, L.stream>>>(P, L.d_in, d_class_lut, L.d_st);
		else if (wpl == 5)
			lc_kernel_warp<5><<<gw, LC_THREADS, 0, L.stream>>>(P, L.d_in, d_class_lut, L.d_st);
		else
			lc_kernel_generic<<<grid, LC_THREADS, 0, L.stream>>>(P, L.d_in, d_class_lut, L.d_st);
	}
	return cudaGetLastError();
}
