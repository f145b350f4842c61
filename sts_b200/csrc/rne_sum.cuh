/*
 * rne_sum.cuh - exact reproduction of a SEQUENTIAL double-precision sum  s = fl(s + x_i)  of
 * non-negative terms, evaluated in parallel by a warp.
 *
 * The reference adds long lists of terms one after the other (universal.c:369, 128000 terms;
 * approximateEntropy.c:396-407, 3072 terms) and the rounding of every step depends on the order, so
 * a tree or an exactly rounded sum gives a different double (FP parity traps).
 * While the running sum stays inside one binade [2^k, 2^(k+1)) it is an integer S in units of
 * u = 2^(k-52), and adding x >= 0 is  S -> S + round(x / u)  where only a tie (x / u = q + 1/2)
 * depends on the state, through the parity of S (round half to even).  Every term is therefore a
 * map  S -> S + (S even ? a0 : a0 + dl), dl in {-1, 0, 1}; these maps are closed under composition
 * (after a tie both parities land on an even value), so 128 terms are composed four per lane and
 * then by a 5-step butterfly.  A group that would leave the binade, or whose terms are not at least
 * two binades below the sum, is reported back and the caller adds it term by term.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct RneMap {
	unsigned long long a0;
	int dl;
};

/* x >= 0 as a map in units of 2^(k-52); needs k >= exponent(x) + 1 */
__device__ __forceinline__ RneMap rne_of_term(double x, int k)
{
	RneMap f;
	f.a0 = 0;
	f.dl = 0;
	/* the sign bit is dropped: the terms are >= 0 by contract, and a -0.0 (the negated empty bin of the ApEn sum) must
	 * count as zero - taken as a number it has "exponent" 1025 and a negative shift below (found by tests/emu) */
	const unsigned long long xb = (unsigned long long) __double_as_longlong(x) & 0x7FFFFFFFFFFFFFFFull;
	if (xb != 0) {
		const int e = (int) (xb >> 52) - 1023;
		const unsigned long long M = (xb & 0x000FFFFFFFFFFFFFull) | 0x0010000000000000ull;
		int sh = k - e;
		sh = sh > 62 ? 62 : sh;
		unsigned long long q = M >> sh;
		const unsigned long long r = M & ((1ull << sh) - 1ull), half = 1ull << (sh - 1);
		if (r > half) {
			q++;
		} else if (r == half) {			/* tie: the result must be even */
			f.dl = (q & 1ull) ? -1 : 1;
			q += (q & 1ull);
		}
		f.a0 = q;
	}
	return f;
}

/* f first, then g */
__device__ __forceinline__ RneMap rne_compose(const RneMap &f, const RneMap &g)
{
	const unsigned long long f1 = f.a0 + (long long) f.dl;
	const unsigned long long g1 = g.a0 + (long long) g.dl;
	const unsigned long long h0 = f.a0 + ((f.a0 & 1ull) ? g1 : g.a0);
	const unsigned long long h1 = f1 + ((f1 & 1ull) ? g.a0 : g1);
	RneMap h;
	h.a0 = h0;
	h.dl = (int) (long long) (h1 - h0);
	return h;
}

/*
 * sum += t[0] + ... + t[127] in index order, lane l holding terms 4l .. 4l+3 (all >= 0, sum >= 0,
 * normal numbers or zero).  Returns true when done exactly on the fast path; false when the caller
 * has to add the 128 terms one by one (sum untouched).  All lanes hold the same `sum`.
 */
__device__ __forceinline__ bool warp_rne_sum128(double &sum, const double4 &t, int lane)
{
	const unsigned long long sb = (unsigned long long) __double_as_longlong(sum);
	const int k = (int) (sb >> 52) - 1023;
	const double tm = fmax(fmax(t.x, t.y), fmax(t.z, t.w));
	const unsigned int hi = __reduce_max_sync(0xffffffffu, (unsigned int) __double2hiint(tm) & 0x7FFFFFFFu);
	const int emax = (int) (hi >> 20) - 1023;
	if (k < emax + 2 || k < -1000)
		return false;
	RneMap f = rne_of_term(t.x, k);
	f = rne_compose(f, rne_of_term(t.y, k));
	f = rne_compose(f, rne_of_term(t.z, k));
	f = rne_compose(f, rne_of_term(t.w, k));
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		RneMap g;
		const uint32_t plo = __shfl_xor_sync(0xffffffffu, (uint32_t) f.a0, o);
		const uint32_t phi = __shfl_xor_sync(0xffffffffu, (uint32_t) (f.a0 >> 32) | ((uint32_t) (f.dl + 1) << 30), o);
		g.a0 = ((unsigned long long) (phi & 0x3FFFFFFFu) << 32) | plo;
		g.dl = (int) (phi >> 30) - 1;
		f = (lane & o) ? rne_compose(g, f) : rne_compose(f, g);
	}
	const unsigned long long S = (sb & 0x000FFFFFFFFFFFFFull) | 0x0010000000000000ull;
	const unsigned long long Snew = S + ((S & 1ull) ? f.a0 + (long long) f.dl : f.a0);
	if (Snew >= 0x0020000000000000ull)		/* would leave the binade */
		return false;
	sum = __longlong_as_double((long long) (((unsigned long long) (k + 1023) << 52) | (Snew & 0x000FFFFFFFFFFFFFull)));
	return true;
}

/*
 * The same for 32 * nrows terms kept in shared memory as rows of 32 with a pitch of 33 doubles (term i at
 * tb[(i / 32) * 33 + i % 32]): lane l composes the nrows consecutive terms [l nrows, (l + 1) nrows) one after
 * the other (decode + compose, about 40 integer instructions per term and no cross-lane traffic), then ONE
 * butterfly merges the 32 spans.  Returns false (sum untouched) when the terms are not two binades below the
 * sum or the sum would leave its binade: the caller then adds the terms one by one.
 */
__device__ __forceinline__ bool warp_rne_sum_rows(double &sum, const double *tb, int nrows, int lane)
{
	const unsigned long long sb = (unsigned long long) __double_as_longlong(sum);
	const int k = (int) (sb >> 52) - 1023;
	if (k < -1000)
		return false;
	RneMap f;
	f.a0 = 0;
	f.dl = 0;
	unsigned int himax = 0;
	for (int j = 0; j < nrows; j++) {
		const int i = lane * nrows + j;
		const double x = tb[(i >> 5) * 33 + (i & 31)];
		const unsigned int hi = (unsigned int) __double2hiint(x) & 0x7FFFFFFFu;
		himax = max(himax, hi);
		/* a term too large for this binade is caught below; keep the shift in range meanwhile */
		f = rne_compose(f, rne_of_term(x, max(k, (int) (hi >> 20) - 1023 + 2)));
	}
	himax = __reduce_max_sync(0xffffffffu, himax);
	if (k < (int) (himax >> 20) - 1023 + 2)
		return false;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		RneMap g;
		const uint32_t plo = __shfl_xor_sync(0xffffffffu, (uint32_t) f.a0, o);
		const uint32_t phi = __shfl_xor_sync(0xffffffffu, (uint32_t) (f.a0 >> 32) | ((uint32_t) (f.dl + 1) << 30), o);
		g.a0 = ((unsigned long long) (phi & 0x3FFFFFFFu) << 32) | plo;
		g.dl = (int) (phi >> 30) - 1;
		f = (lane & o) ? rne_compose(g, f) : rne_compose(f, g);
	}
	const unsigned long long S = (sb & 0x000FFFFFFFFFFFFFull) | 0x0010000000000000ull;
	const unsigned long long Snew = S + ((S & 1ull) ? f.a0 + (long long) f.dl : f.a0);
	if (Snew >= 0x0020000000000000ull)		/* would leave the binade */
		return false;
	sum = __longlong_as_double((long long) (((unsigned long long) (k + 1023) << 52) | (Snew & 0x000FFFFFFFFFFFFFull)));
	return true;
}

/*
 * Fast path for terms kept as exact fixed point X = x * 2^RNE_FX (the Universal test's table of log2 values):
 * away from ties, adding x to a sum in binade k is S += (X + half) >> sh with sh = k - 52 + RNE_FX, i.e. a plain
 * integer sum, about a dozen instructions per term and one 64-bit add-reduction per 32 * nrows terms.  A tie
 * (X mod 2^sh == half, about once per stream) or a binade change returns false with sum untouched; the caller
 * then uses the composable maps above on those terms.  Layout as in warp_rne_sum_rows.
 */
#define RNE_FX 58
__device__ __forceinline__ double rne_fixed_to_double(unsigned long long X)
{
	return __ull2double_rn(X) * 3.4694469519536141888e-18;		/* 2^-58; X has at most 53 significant bits: exact */
}
__device__ __forceinline__ bool warp_rne_sum_rows_fixed(double &sum, const unsigned long long *tb, int nrows, int lane)
{
	const unsigned long long sb = (unsigned long long) __double_as_longlong(sum);
	const int k = (int) (sb >> 52) - 1023;
	const int sh = k - 52 + RNE_FX;
	if (sh < 1 || sh > 63)
		return false;
	const unsigned long long half = 1ull << (sh - 1), mask = (half << 1) - 1ull;
	unsigned long long acc = 0;
	bool tie = false;
	for (int j = 0; j < nrows; j++) {
		const int i = lane * nrows + j;
		const unsigned long long X = tb[(i >> 5) * 33 + (i & 31)];
		acc += (X + half) >> sh;
		tie = tie || ((X & mask) == half);
	}
	if (__any_sync(0xffffffffu, tie))
		return false;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t) acc, o);
		const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t) (acc >> 32), o);
		acc += ((unsigned long long) hi << 32) | lo;
	}
	const unsigned long long S = (sb & 0x000FFFFFFFFFFFFFull) | 0x0010000000000000ull;
	const unsigned long long Snew = S + acc;
	if (Snew >= 0x0020000000000000ull)		/* would leave the binade */
		return false;
	sum = __longlong_as_double((long long) (((unsigned long long) (k + 1023) << 52) | (Snew & 0x000FFFFFFFFFFFFFull)));
	return true;
}
