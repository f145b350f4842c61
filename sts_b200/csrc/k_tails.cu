/*
 * k_tails.cu - the floating point tails: integer statistics -> p-values.
 *
 * COMPILED WITH -fmad=false: every expression below is written in the reference's operand order
 * and must round like gcc -std=c99 rounds it on x86-64 (no FMA contraction), so that the only
 * difference left against the reference is CUDA's libdevice vs glibc in erf/erfc/exp/log (<= 2 ulp).
 * lgamma(a) is only ever needed for per-configuration constants a, so the host evaluates it with
 * glibc (as the reference does through cephes.h:33-35) and passes the values in TailConsts.
 *
 * Three kernels, all latency-bound chains that only pay off when thousands of them run at once:
 *   tails_slots_kernel : one THREAD per (p-value slot, stream), a warp = one slot of 32 consecutive
 *            streams, so a warp's lanes run the same cephes loop (igamc with a = 2^15 / 2 for Serial is
 *            ~1000 dependent divisions; with one CTA per stream it serialised the whole batch);
 *   tails_cusum_kernel : CumulativeSums (cusum.c:329-370), one warp per (stream, direction): 32 terms
 *            of the two series at a time, added in the reference's k order;
 *   tails_apen_kernel  : ApproximateEntropy phi (approximateEntropy.c:396-407), one warp per stream:
 *            C*log(C/n) for 128 bins at a time, accumulated in index order exactly as the reference's
 *            sequential loop does, but in parallel (rne_sum.cuh), then igamc.
 */
#include "common.cuh"
#include "tails.h"
#include "rne_sum.cuh"

#define TAIL_THREADS 128
#define NONP STSGPU_NON_P_VALUE

/* utils/cephes.c:31-32,49-50 */
#define K_MACHEP 1.11022302462515654042363e-16
#define K_MAXLOG 7.0978271289338399673222e2
#define K_BIG 4.503599627370496e15
#define K_BIGINV 2.22044604925031308085e-16
#define K_SQRT2 1.41421356237309504880

__device__ double d_igam_series(double a, double x, double lga);

/* utils/cephes.c:60-132 */
__device__ double d_igamc(double a, double x, double lga)
{
	if ((x <= 0) || (a <= 0))
		return 1.0;
	if ((x < 1.0) || (x < a))
		return 1.e0 - d_igam_series(a, x, lga);	/* igam's own redirect (x > 1 && x > a) cannot trigger here */
	double ax = a * log(x) - x - lga;
	if (ax < -K_MAXLOG)
		return 0.0;
	ax = exp(ax);
	double y = 1.0 - a;
	double z = x + y + 1.0;
	double c = 0.0;
	double pkm2 = 1.0;
	double qkm2 = x;
	double pkm1 = x + 1.0;
	double qkm1 = z * x;
	double ans = pkm1 / qkm1;
	double t;
	do {
		c += 1.0;
		y += 1.0;
		z += 2.0;
		double yc = y * c;
		double pk = pkm1 * z - pkm2 * yc;
		double qk = qkm1 * z - qkm2 * yc;
		if (qk != 0) {
			double r = pk / qk;
			t = fabs((ans - r) / r);
			ans = r;
		} else {
			t = 1.0;
		}
		pkm2 = pkm1;
		pkm1 = pk;
		qkm2 = qkm1;
		qkm1 = qk;
		if (fabs(pk) > K_BIG) {
			pkm2 *= K_BIGINV;
			pkm1 *= K_BIGINV;
			qkm2 *= K_BIGINV;
			qkm1 *= K_BIGINV;
		}
	} while (t > K_MACHEP);
	return ans * ax;
}

/* utils/cephes.c:135-174, power series branch (reached only with x < 1 or x < a, so no recursion) */
__device__ double d_igam_series(double a, double x, double lga)
{
	if ((x <= 0) || (a <= 0))
		return 0.0;
	double ax = a * log(x) - x - lga;
	if (ax < -K_MAXLOG)
		return 0.0;
	ax = exp(ax);
	double r = a;
	double c = 1.0;
	double ans = 1.0;
	do {
		r += 1.0;
		c *= x / r;
		ans += c;
	} while (c / ans > K_MACHEP);
	return ans * ax / a;
}

/* utils/cephes.c:416-430 */
__device__ __forceinline__ double d_normal(double x)
{
	if (x > 0)
		return 0.5 * (1 + erf(x / K_SQRT2));
	return 0.5 * (1 - erf(-x / K_SQRT2));
}

__device__ __forceinline__ int test_of_slot(const DevParams &P, int q, int *sub)
{
	int t = 0;
	for (int k = 1; k <= 15; k++)
		if ((P.enabled >> k) & 1u)
			if (P.pv_off[k] <= q)
				t = k;
	*sub = q - P.pv_off[t];
	return t;
}

__global__ void __launch_bounds__(TAIL_THREADS)
tails_slots_kernel(DevParams P, TailConsts K, long long *__restrict__ st_all, const uint32_t *__restrict__ w_all,
		   double *__restrict__ pv_all, double *__restrict__ fs_all, int64_t nstreams)
{
	/* warp -> (slot q, 32 consecutive streams) */
	const int64_t groups = (nstreams + 31) >> 5;
	const int64_t gw = (int64_t) blockIdx.x * (TAIL_THREADS / 32) + (threadIdx.x >> 5);
	const int q = (int) (gw / groups);
	const int64_t s = (gw % groups) * 32 + (threadIdx.x & 31);
	if (q >= P.npv || s >= nstreams)
		return;
	long long *st = st_all + s * P.nst;
	const uint32_t *wc = w_all + s * P.nw;
	double *pv = pv_all + s * P.npv;
	const long long n = P.n;
	{
		int sub;
		const int t = test_of_slot(P, q, &sub);
		const long long *c = st + P.st_off[t];
		/* -s: the doubles of the test's private stats struct (include/stsgpu.h, fstats) */
		double *fs = fs_all ? fs_all + s * P.nfs + P.fs_off[t] : nullptr;
		double p = 0.0;
		bool write = true;
		switch (t) {
		case STSGPU_TEST_FREQUENCY: {			/* frequency.c:206-212 */
			double s_obs = fabs((double) c[0]) / K.sqrtn;
			double f = s_obs / K.sqrt2;
			p = erfc(f);
			break;
		}
		case STSGPU_TEST_BLOCK_FREQUENCY: {		/* blockFrequency.c:217-247 */
			const long long M = P.bf_M, N = P.bf_N;
			double sum = 0.0;
			for (long long i = 0; i < N; i++) {
				double pi = (double) c[1 + i] / (double) M;
				double v = pi - 0.5;
				sum += v * v;
			}
			double chi_squared = 4.0 * M * sum;
			if (fs) fs[0] = chi_squared;
			p = d_igamc(N / 2.0, chi_squared / 2.0, K.lgam_bf);
			break;
		}
		case STSGPU_TEST_RUNS: {			/* runs.c:219-246 */
			double pi = (double) c[0] / (double) n;
			bool possible = (fabs(pi - 0.5) >= K.two_over_sqrtn) ? false : true;
			st[P.st_off[t] + 2] = possible ? 1 : 0;
			if (possible) {
				long long V_n = c[1];
				double erfc_arg = fabs(V_n - 2.0 * (double) n * pi * (1.0 - pi)) /
						  (2.0 * pi * (1.0 - pi) * K.sqrt2n);
				p = erfc(erfc_arg);
				if (fs) { fs[0] = pi; fs[1] = erfc_arg; }
			} else {
				p = NONP;
				if (fs) { fs[0] = 0.0; fs[1] = 0.0; }	/* UNSET_DOUBLE, runs.c:299-301 */
			}
			break;
		}
		case STSGPU_TEST_LONGEST_RUN: {			/* longestRunOfOnes.c:411-421 */
			const long long N = P.lr_N;
			double chi2 = 0.0;
			for (int i = 0; i <= 6; i++) {
				double chi_term = c[i] - (N * K.lr_pi[i]);
				chi2 += chi_term * chi_term / (N * K.lr_pi[i]);
			}
			if (fs) fs[0] = chi2;
			p = d_igamc(6 / 2.0, chi2 / 2.0, K.lgam_3);
			break;
		}
		case STSGPU_TEST_RANK: {			/* rank.c:322-341 */
			const long long mc = P.rank_count, F_M = c[0], F_M1 = c[1];
			const long long F_rem = mc - (F_M + F_M1);
			st[P.st_off[t] + 2] = F_rem;
			double chi = (((F_M - mc * K.p_32) * (F_M - mc * K.p_32) / (mc * K.p_32)) +
				      ((F_M1 - mc * K.p_31) * (F_M1 - mc * K.p_31) / (mc * K.p_31)) +
				      ((F_rem - mc * K.p_30) * (F_rem - mc * K.p_30) / (mc * K.p_30)));
			if (fs) fs[0] = chi;
			p = exp(-chi / 2.0);
			break;
		}
		case STSGPU_TEST_DFT: {				/* discreteFourierTransform.c:401-423 */
			double N_0 = (double) 0.95 * n / 2.0;
			double d = (c[0] - N_0) / K.sqrtn4_095_005;
			if (fs) { fs[0] = N_0; fs[1] = d; }
			p = erfc(fabs(d) / K.sqrt2);
			break;
		}
		case STSGPU_TEST_NON_OVERLAPPING: {		/* nonOverlappingTemplateMatchings.c:556-568 */
			double chi2 = 0.0;
			for (int i = 0; i < 8; i++) {
				double chi2_term = ((double) wc[sub * 8 + i] - K.nonov_mu) / K.nonov_sqrt_sigma2;
				chi2 += (chi2_term * chi2_term);
			}
			if (fs) fs[sub] = chi2;
			p = d_igamc(8 / 2.0, chi2 / 2.0, K.lgam_4);
			break;
		}
		case STSGPU_TEST_OVERLAPPING: {			/* overlappingTemplateMatchings.c:290-303 */
			const long long N = P.ov_N;
			double chi2 = 0.0;
			for (int i = 0; i < 6; i++) {
				double chi2_term = (double) c[i] - (double) N * K.ov_pi[i];
				chi2 += chi2_term * chi2_term / ((double) N * K.ov_pi[i]);
			}
			if (fs) fs[0] = chi2;
			p = d_igamc(5 / 2.0, chi2 / 2.0, K.lgam_2_5);
			break;
		}
		case STSGPU_TEST_UNIVERSAL: {			/* universal.c:380-390 */
			double sum = __longlong_as_double(c[0]);
			double f_n = (sum / (double) P.univ_K);
			if (fs) { fs[0] = K.univ_sigma; fs[1] = f_n; }
			double arg = fabs(f_n - K.univ_expected) / (K.sqrt2 * K.univ_sigma);
			p = erfc(arg);
			break;
		}
		case STSGPU_TEST_RND_EXCURSION: {		/* randomExcursions.c:358, :470-503, :597-600 */
			const long long J = c[0];
			if (J < P.min_zero_crossings) {
				p = NONP;
				if (fs) fs[sub] = 0.0;
			} else {
				int x = (sub < 4) ? (sub - 4) : (sub - 3);
				int ax = x < 0 ? -x : x;
				double chi2 = 0.0;
				for (int j = 0; j < 6; j++) {
					double sum_term = (double) c[1 + 8 * j + sub] - ((double) J * K.rex_pi[ax - 1][j]);
					chi2 += sum_term * sum_term / ((double) J * K.rex_pi[ax - 1][j]);
				}
				if (fs) fs[sub] = chi2;
				p = d_igamc((double) (6 - 1) / 2.0, chi2 / 2.0, K.lgam_2_5);
			}
			break;
		}
		case STSGPU_TEST_RND_EXCURSION_VAR: {		/* randomExcursionsVariant.c:283, :311-313 */
			const long long J = c[0];
			if (J < P.min_zero_crossings) {
				p = NONP;
			} else {
				long long x = (sub < 9) ? (sub - 9) : (sub - 8);
				long long ax = x < 0 ? -x : x;
				long long diff = c[1 + sub] - J;
				if (diff < 0) diff = -diff;
				p = erfc(diff / (sqrt(2.0 * J * (4.0 * ax - 2.0))));
			}
			break;
		}
		case STSGPU_TEST_SERIAL: {			/* serial.c:205-224, :414-425 */
			const int m = P.serial_m;
			double psi[3];
			for (int r = 0; r < 3; r++) {
				int b = m - r;
				if (b <= 0) {
					psi[r] = 0.0;
				} else {
					double sum = (double) c[r];
					psi[r] = (sum * (double) ((long long) 1 << b) / (double) n) - (double) n;
				}
			}
			double del1 = psi[0] - psi[1];
			double del2 = psi[0] - 2.0 * psi[1] + psi[2];
			if (fs && sub == 0) { fs[0] = psi[0]; fs[1] = psi[1]; fs[2] = psi[2]; fs[3] = del1; fs[4] = del2; }
			if (sub == 0)
				p = d_igamc((double) ((long long) 1 << (m - 1)) / 2.0, del1 / 2.0, K.lgam_serial1);
			else
				p = d_igamc((double) ((long long) 1 << (m - 2)) / 2.0, del2 / 2.0, K.lgam_serial2);
			break;
		}
		case STSGPU_TEST_LINEARCOMPLEXITY: {		/* linearComplexity.c:382-390 */
			const long long N = P.lc_N;
			double chi2 = 0.0;
			for (int i = 0; i < 7; i++)
				chi2 += (c[i] - N * K.lc_pi[i]) * (c[i] - N * K.lc_pi[i]) / (N * K.lc_pi[i]);
			if (fs) fs[0] = chi2;
			p = d_igamc(6 / 2.0, chi2 / 2.0, K.lgam_3);
			break;
		}
		default:
			write = false;		/* CUSUM and APEN: stages 2 and 3 */
			break;
		}
		if (write)
			pv[q] = p;
	}

}

/* CumulativeSums, cusum.c:353-369: one warp per (stream, direction) */
__global__ void __launch_bounds__(TAIL_THREADS)
tails_cusum_kernel(DevParams P, TailConsts K, const long long *__restrict__ st_all, double *__restrict__ pv_all,
		   int64_t nstreams)
{
	const int64_t gw = (int64_t) blockIdx.x * (TAIL_THREADS / 32) + (threadIdx.x >> 5);
	const int64_t s = gw >> 1;
	const int warp = (int) (gw & 1);			/* 0: forward, 1: backward */
	const int lane = threadIdx.x & 31;
	if (s >= nstreams)
		return;
	const long long *st = st_all + s * P.nst;
	double *pv = pv_all + s * P.npv;
	const long long n = P.n;
	{
		const long long z = st[P.st_off[STSGPU_TEST_CUSUM] + 3 + warp];	/* warp 0: forward, 1: backward */
		double sum1 = 0.0, sum2 = 0.0;
		if (z > 0) {
			const long long k1lo = (-n / z + 1) / 4, k1hi = (n / z - 1) / 4;
			for (long long k0 = k1lo; k0 <= k1hi; k0 += 32) {
				long long k = k0 + lane;
				double a = 0.0, b = 0.0;
				if (k <= k1hi) {
					a = d_normal(((4 * k + 1) * z) / K.sqrtn);
					b = d_normal(((4 * k - 1) * z) / K.sqrtn);
				}
#pragma unroll 4
				for (int l = 0; l < 32; l++) {
					sum1 += __shfl_sync(0xffffffffu, a, l);
					sum1 -= __shfl_sync(0xffffffffu, b, l);
				}
			}
			const long long k2lo = (-n / z - 3) / 4, k2hi = (n / z - 1) / 4;
			for (long long k0 = k2lo; k0 <= k2hi; k0 += 32) {
				long long k = k0 + lane;
				double a = 0.0, b = 0.0;
				if (k <= k2hi) {
					a = d_normal(((4 * k + 3) * z) / K.sqrtn);
					b = d_normal(((4 * k + 1) * z) / K.sqrtn);
				}
#pragma unroll 4
				for (int l = 0; l < 32; l++) {
					sum2 += __shfl_sync(0xffffffffu, a, l);
					sum2 -= __shfl_sync(0xffffffffu, b, l);
				}
			}
		}
		if (lane == 0)
			pv[P.pv_off[STSGPU_TEST_CUSUM] + warp] = 1.0 - sum1 + sum2;
	}

}

/* ApproximateEntropy, approximateEntropy.c:396-407 and :230-246: one warp per stream */
__global__ void __launch_bounds__(TAIL_THREADS)
tails_apen_kernel(DevParams P, TailConsts K, long long *__restrict__ st_all, const uint32_t *__restrict__ apen_hist,
		  double *__restrict__ pv_all, double *__restrict__ fs_all, int64_t nstreams)
{
	__shared__ __align__(16) double tbuf[TAIL_THREADS / 32][128];
	const int64_t s = (int64_t) blockIdx.x * (TAIL_THREADS / 32) + (threadIdx.x >> 5);
	const int lane = threadIdx.x & 31;
	if (s >= nstreams)
		return;
	double *tb = tbuf[threadIdx.x >> 5];
	long long *st = st_all + s * P.nst;
	double *pv = pv_all + s * P.npv;
	const long long n = P.n;
	const int m = P.apen_m;
	const uint32_t *h = apen_hist + s * ((size_t) 3 << m);
	double phi[2];
	for (int r = 0; r < 2; r++) {
		const uint32_t *hr = (r == 1) ? h : h + ((size_t) 2 << m);	/* r = 0: level m, r = 1: level m+1 */
		const long long bins = 1LL << (m + r);
		/* every term C*log(C/n) is <= 0: accumulate magnitudes (rounding is symmetric), negate at the end */
		double mag = 0.0;
		for (long long base = 0; base < bins; base += 128) {
			double tm[4];
#pragma unroll
			for (int j = 0; j < 4; j++) {
				const long long i = base + 4 * lane + j;
				double mterm = 0.0;		/* -C log(C / n) >= 0; an empty bin adds +0.0, not -0.0 */
				if (i < bins) {
					const uint32_t C = hr[i];
					if (C)
						mterm = -((double) C * log(C / (double) n));
				}
				tm[j] = mterm;
			}
			const double4 t4 = make_double4(tm[0], tm[1], tm[2], tm[3]);
			if (!warp_rne_sum128(mag, t4, lane)) {
				__syncwarp();
				*(double4 *) (tb + 4 * lane) = t4;
				__syncwarp();
#pragma unroll 8
				for (int j = 0; j < 128; j++)
					mag += tb[j];
			}
		}
		phi[r] = (mag == 0.0) ? 0.0 : -mag / (double) n;	/* the reference's sum of non-positive terms is +0.0, not -0.0, when every term is zero */
	}
	if (lane == 0) {
		long long *c = st + P.st_off[STSGPU_TEST_APEN];
		c[0] = __double_as_longlong(phi[0]);
		c[1] = __double_as_longlong(phi[1]);
		double ApEn = phi[0] - phi[1];
		double chi_squared = 2.0 * n * (K.log2_ - ApEn);
		if (fs_all) {
			double *fs = fs_all + s * P.nfs + P.fs_off[STSGPU_TEST_APEN];
			fs[0] = ApEn;
			fs[1] = chi_squared;
		}
		pv[P.pv_off[STSGPU_TEST_APEN]] = d_igamc((double) ((long long) 1 << (m - 1)), chi_squared / 2.0, K.lgam_apen);
	}
}

/* ------------------------------------------------------------------------------------------ */
/*
 * Assessment tallies (include/stsgpu.h, stsgpu_assess_*): frequency.c:828-894, one thread per
 * (stream, p-value column) of the batch; per column [sampleCount, toolow, freq[bins]].
 */
/* one p-value of column q into the tallies (the loop body of frequency.c:828-894) */
__device__ __forceinline__ void tally_one(const DevParams &P, int q, double p_value, int64_t bins, double alpha,
					  unsigned long long *__restrict__ tally)
{
	if (p_value == NONP)
		return;					/* the test was not possible for this iteration */
	const int e12 = P.pv_off[STSGPU_TEST_RND_EXCURSION], e13 = P.pv_off[STSGPU_TEST_RND_EXCURSION_VAR];
	const bool is_excursion = ((P.enabled & (1u << STSGPU_TEST_RND_EXCURSION)) && q >= e12 && q < e12 + 8) ||
				  ((P.enabled & (1u << STSGPU_TEST_RND_EXCURSION_VAR)) && q >= e13 && q < e13 + 18);
	if (is_excursion && !(p_value > 0.0))
		return;					/* excursion tests only sample p-values > 0 */
	unsigned long long *t = tally + (int64_t) q * (bins + 2);
	atomicAdd(&t[0], 1ull);
	if (p_value < alpha)
		atomicAdd(&t[1], 1ull);
	int64_t b;
	if (p_value >= 1.0)
		b = bins - 1;
	else if (p_value >= 0.0)
		b = (int) floor(p_value * (double) bins);
	else
		b = 0;
	atomicAdd(&t[2 + b], 1ull);
}

__global__ void __launch_bounds__(256)
assess_tally_kernel(DevParams P, const double *__restrict__ pv_all, int64_t nstreams, int64_t bins, double alpha,
		    unsigned long long *__restrict__ tally, const unsigned long long *__restrict__ limit)
{
	const int64_t idx = (int64_t) blockIdx.x * 256 + threadIdx.x;
	if (idx >= nstreams * P.npv)
		return;
	if (limit != nullptr && (unsigned long long) (idx / P.npv) >= *limit)
		return;					/* "-F a" batch: the text ended before this stream was complete */
	tally_one(P, (int) (idx % P.npv), pv_all[idx], bins, alpha, tally);
}

/* p-values of ONE test in .pvalues order (iteration-major, utilities.c:1967-2024): value i of the test's sequence
 * belongs to column pv_off[test] + i mod pv_cnt[test] */
__global__ void __launch_bounds__(256)
assess_add_kernel(DevParams P, int pv_off, int pv_cnt, const double *__restrict__ vals, int64_t count, int64_t first_index, int64_t bins,
		  double alpha, unsigned long long *__restrict__ tally)
{
	const int64_t i = (int64_t) blockIdx.x * 256 + threadIdx.x;
	if (i >= count)
		return;
	const int q = pv_off + (int) ((first_index + i) % pv_cnt);
	tally_one(P, q, vals[i], bins, alpha, tally);
}

/* frequency.c:631-700 and :739-757, one thread per column */
__global__ void __launch_bounds__(128)
assess_finish_kernel(DevParams P, int64_t bins, double alpha, double level, double lgam,
		     const unsigned long long *__restrict__ tally, stsgpu_metric *__restrict__ out)
{
	const int q = blockIdx.x * 128 + threadIdx.x;
	if (q >= P.npv)
		return;
	const unsigned long long *t = tally + (int64_t) q * (bins + 2);
	const long long sampleCount = (long long) t[0], toolow = (long long) t[1];
	long long passCount;
	if ((sampleCount <= 0) || (sampleCount < toolow))
		passCount = 0;
	else
		passCount = sampleCount - toolow;
	const double p_hat = 1.0 - alpha;
	const double tmax = (p_hat + 3.0 * sqrt((p_hat * alpha) / sampleCount)) * sampleCount;
	const double tmin = (p_hat - 3.0 * sqrt((p_hat * alpha) / sampleCount)) * sampleCount;
	double chi2 = 0.0, uniformity;
	const double expCount = (double) sampleCount / bins;
	if (expCount <= 0.0) {
		uniformity = 0.0;
	} else {
		for (int64_t i = 0; i < bins; ++i) {
			const double f = (double) (long long) t[2 + i];
			chi2 += (f - expCount) * (f - expCount) / expCount;
		}
		uniformity = d_igamc((bins - 1.0) / 2.0, chi2 / 2.0, lgam);
	}
	stsgpu_metric m;
	m.sample_count = sampleCount;
	m.too_low = toolow;
	m.pass_count = passCount;
	m.proportion_min = tmin;
	m.proportion_max = tmax;
	m.chi2 = chi2;
	m.uniformity = uniformity;
	m.uniformity_passed = !(expCount <= 0.0 || uniformity < level);
	m.proportion_passed = !(sampleCount == 0 || (passCount < tmin) || (passCount > tmax));
	out[q] = m;
}

cudaError_t launch_assess_tally(const DevParams &P, const LaunchCtx &L, int64_t bins, double alpha, unsigned long long *d_tally,
				const unsigned long long *d_limit)
{
	const int64_t total = L.nstreams * P.npv;
	assess_tally_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, L.stream>>>(P, L.d_pv, L.nstreams, bins, alpha, d_tally, d_limit);
	return cudaGetLastError();
}

cudaError_t launch_assess_add(const DevParams &P, int pv_off, int pv_cnt, const double *d_vals, int64_t count, int64_t first_index,
			      int64_t bins, double alpha, unsigned long long *d_tally, cudaStream_t stream)
{
	assess_add_kernel<<<(unsigned) ((count + 255) / 256), 256, 0, stream>>>(P, pv_off, pv_cnt, d_vals, count, first_index, bins, alpha,
										 d_tally);
	return cudaGetLastError();
}

cudaError_t launch_assess_finish(const DevParams &P, int64_t bins, double alpha, double level, double lgam,
				 const unsigned long long *d_tally, stsgpu_metric *d_out, cudaStream_t stream)
{
	assess_finish_kernel<<<(unsigned) ((P.npv + 127) / 128), 128, 0, stream>>>(P, bins, alpha, level, lgam, d_tally, d_out);
	return cudaGetLastError();
}

cudaError_t launch_tails(const DevParams &P, const TailConsts &K, const LaunchCtx &L, const uint32_t *d_apen_hist,
			 cudaStream_t s_cusum, cudaStream_t s_apen)
{
	/* the three kernels are independent chains: the caller forks s_cusum / s_apen off L.stream and joins them */
	const int wpb = TAIL_THREADS / 32;
	const int64_t groups = (L.nstreams + 31) >> 5;
	const int64_t slot_warps = groups * P.npv;
	tails_slots_kernel<<<(unsigned) ((slot_warps + wpb - 1) / wpb), TAIL_THREADS, 0, L.stream>>>(P, K, L.d_st, L.d_w, L.d_pv,
												      L.d_fs, L.nstreams);
	if (P.enabled & (1u << STSGPU_TEST_CUSUM))
		tails_cusum_kernel<<<(unsigned) ((2 * L.nstreams + wpb - 1) / wpb), TAIL_THREADS, 0, s_cusum>>>(P, K, L.d_st, L.d_pv,
														 L.nstreams);
	if (P.enabled & (1u << STSGPU_TEST_APEN))
		tails_apen_kernel<<<(unsigned) ((L.nstreams + wpb - 1) / wpb), TAIL_THREADS, 0, s_apen>>>(P, K, L.d_st, d_apen_hist,
													   L.d_pv, L.d_fs, L.nstreams);
	return cudaGetLastError();
}
