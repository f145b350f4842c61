/*
 * sts_oracle.c - CPU restatement of the 15 X_iterate() bodies of arcetri/sts v3.2.7 and of the
 * p-value tails they call.  TEST INFRASTRUCTURE (see sts_oracle.h): the checker, never the product.
 *
 * Every function cites the reference lines it follows (paths relative to the reference's src/).
 * Like the reference it works on an "epsilon" array of one byte per bit (defs.h:212) and sums
 * floating point in the reference's order with the reference's libm, so p-values agree with the
 * reference binary to the last bit for the 14 tests that do not depend on libfftw3.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include "sts_oracle.h"
#include "fft_f64.h"

#define NON_P_VALUE (-99999999.0)	/* utils/defs.h:314 */

/* ---------------------------------------------------------------------------------------------
 * p-value tails: utils/cephes.c:60-174 (igamc continued fraction, igam power series)
 * constants utils/cephes.c:31-32,49-50; lgam is libm lgamma because -D_ISOC99_SOURCE (cephes.h:27-35)
 * ------------------------------------------------------------------------------------------- */
static const double K_MACHEP = 1.11022302462515654042363e-16;
static const double K_MAXLOG = 7.0978271289338399673222e2;
static const double K_BIG = 4.503599627370496e15;
static const double K_BIGINV = 2.22044604925031308085e-16;
static const double K_SQRT2 = 1.41421356237309504880;	/* utils/cephes.c:43-47 */

double oracle_igam(double a, double x);

double
oracle_igamc(double a, double x)
{
	if (x <= 0 || a <= 0)
		return 1.0;
	if (x < 1.0 || x < a)
		return 1.e0 - oracle_igam(a, x);
	double ax = a * log(x) - x - lgamma(a);
	if (ax < -K_MAXLOG)
		return 0.0;
	ax = exp(ax);
	double y = 1.0 - a, z = x + y + 1.0, c = 0.0;
	double pkm2 = 1.0, qkm2 = x, pkm1 = x + 1.0, qkm1 = z * x;
	double ans = pkm1 / qkm1, t;
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
		pkm2 = pkm1; pkm1 = pk; qkm2 = qkm1; qkm1 = qk;
		if (fabs(pk) > K_BIG) {
			pkm2 *= K_BIGINV; pkm1 *= K_BIGINV; qkm2 *= K_BIGINV; qkm1 *= K_BIGINV;
		}
	} while (t > K_MACHEP);
	return ans * ax;
}

double
oracle_igam(double a, double x)
{
	if (x <= 0 || a <= 0)
		return 0.0;
	if (x > 1.0 && x > a)
		return 1.e0 - oracle_igamc(a, x);
	double ax = a * log(x) - x - lgamma(a);
	if (ax < -K_MAXLOG)
		return 0.0;
	ax = exp(ax);
	double r = a, c = 1.0, ans = 1.0;
	do {
		r += 1.0;
		c *= x / r;
		ans += c;
	} while (c / ans > K_MACHEP);
	return ans * ax / a;
}

/* utils/cephes.c:416-430 */
static double
o_normal(double x)
{
	if (x > 0)
		return 0.5 * (1 + erf(x / K_SQRT2));
	return 0.5 * (1 - erf(-x / K_SQRT2));
}

/* ---------------------------------------------------------------------------------------------
 * derived constants and layout
 * ------------------------------------------------------------------------------------------- */

/* nonOverlappingTemplateMatchings.c:338-401 (appendTemplate): keep words with no self-overlap */
static int
template_ok(unsigned long value, int m)
{
	unsigned char A[32];
	for (int c = 0; c < m; c++)
		A[c] = (unsigned char) ((value >> (m - 1 - c)) & 1);
	for (int i = 1; i < m; i++) {
		int overlap = 1;
		if ((A[0] != A[m - 1]) && ((A[0] != A[m - 2]) || (A[1] != A[m - 1]))) {
			for (int c = 0; c < m - i; c++) {
				if (A[c] != A[c + i]) {
					overlap = 0;
					break;
				}
			}
		}
		if (overlap)
			return 0;
	}
	return 1;		/* (for m == 1 the reference would reject; m >= 8 here) */
}

static int
count_templates(int m, uint32_t *values)
{
	int cnt = 0;
	for (unsigned long v = 1; v < (1UL << m); v++) {
		if (template_ok(v, m)) {
			if (values)
				values[cnt] = (uint32_t) v;
			cnt++;
		}
	}
	return cnt;
}

/* universal.c:145-178: largest L in [6,16] with n >= 1010 * 2^L * L */
static int
universal_L_for(int64_t n)
{
	int L;
	for (L = 7; L <= 16; L++) {
		if (n < (1010L * (1L << L) * L))
			break;
	}
	return L - 1;
}

int
oracle_layout_for(const oracle_params *p, oracle_layout *o)
{
	memset(o, 0, sizeof(*o));
	int64_t n = p->n;
	uint32_t en = p->test_mask & 0xFFFEu;
	double logn = log((double) n), log2_ = log(2.0);

	if (n < 100) en &= ~(1u << 1);						/* frequency.c:102 */
	{									/* blockFrequency.c:109-129 */
		int64_t M = p->block_frequency_M, N = (M > 0) ? n / M : 0;
		if (n < 100 || M < 20 || M <= 0.01 * n || N >= 100) en &= ~(1u << 2);
	}
	if (n < 100) en &= ~((1u << 3) | (1u << 4));				/* cusum.c:109, runs.c:114 */
	if (n < 128) en &= ~(1u << 5);						/* longestRunOfOnes.c:261 */
	if (n / 1024 < 38) en &= ~(1u << 6);					/* rank.c:119-128 */
	if (n < 1000) en &= ~(1u << 7);						/* discreteFourierTransform.c:123 */
	{									/* nonOverlapping...c:211-241 */
		int64_t M = n / 8;
		int m = p->non_overlapping_m;
		if (M <= 0.01 * n || m < 8 || m > 15 || m > M) en &= ~(1u << 8);
	}
	{									/* overlapping...c:151-168 */
		int64_t N = n / 1032;
		if (p->overlapping_m != 9 || n < 1000000 || N * 0.070432326346398449744 <= 5) en &= ~(1u << 9);
	}
	o->universal_L = universal_L_for(n);
	if (n < 387840 || o->universal_L < 6 || o->universal_L > 16) en &= ~(1u << 10);	/* universal.c:133-192 */
	if (p->apen_m >= lround(floor(logn / log2_ - 5.0))) en &= ~(1u << 11);		/* approximateEntropy.c:109 */
	if (n < 1000000) en &= ~((1u << 12) | (1u << 13));			/* randomExcursions.c:112, Variant:110 */
	if (p->serial_m >= lround(floor(logn / log2_ - 2.0))) en &= ~(1u << 14);	/* serial.c:112 */
	{									/* linearComplexity.c:118-138 */
		int64_t M = p->linear_complexity_M, N = (M > 0) ? n / M : 0;
		if (n < 1000000 || M < 500 || M > 5000 || N < 200) en &= ~(1u << 15);
	}
	o->enabled_mask = en;
	o->ntemplates = (en & (1u << 8)) ? count_templates(p->non_overlapping_m, NULL) : 0;
	o->nw = o->ntemplates * 8;

	const int pvc[16] = { 0, 1, 1, 2, 1, 1, 1, 1, o->ntemplates, 1, 1, 1, 8, 18, 2, 1 };
	int bfN = (p->block_frequency_M > 0) ? (int) (n / p->block_frequency_M) : 0;
	const int stc[16] = { 0, 1, 1 + bfN, 5, 3, 7, 3, 1, 0, 6, 1, 6, 57, 19, 3, 7 };
	int po = 0, so = 0;
	for (int t = 1; t <= 15; t++) {
		if (!(en & (1u << t)))
			continue;
		o->pv_off[t] = po; o->pv_cnt[t] = pvc[t]; po += pvc[t];
		o->st_off[t] = so; o->st_cnt[t] = stc[t]; so += stc[t];
	}
	o->npv = po;
	o->nst = so;
	return 0;
}

/* ---------------------------------------------------------------------------------------------
 * per-run constants shared by all streams
 * ------------------------------------------------------------------------------------------- */
typedef struct {
	oracle_params p;
	oracle_layout lay;
	double sqrt2, log2_, sqrtn, logn;	/* driver.c:304-311 */
	long min_zero_crossings;		/* driver.c:316 */
	double p_32, p_31, p_30;		/* rank.c:131-170 */
	uint32_t *tmpl;				/* template values, ascending (nonOverlapping...c:292-295) */
	double rex_pi[4][6];			/* randomExcursions.c:187-207 */
	fft_plan *fft;
} oconst;

static void
oconst_init(oconst *c, const oracle_params *p)
{
	memset(c, 0, sizeof(*c));
	c->p = *p;
	oracle_layout_for(p, &c->lay);
	c->sqrt2 = sqrt(2.0);
	c->log2_ = log(2.0);
	c->sqrtn = sqrt((double) p->n);
	c->logn = log((double) p->n);
	double mz = 0.005 * c->sqrtn;
	c->min_zero_crossings = (long) (mz > 500.0 ? mz : 500.0);
	/* rank.c:131-170 with NUMBER_OF_ROWS_RANK = NUMBER_OF_COLS_RANK = 32 */
	for (int pass = 0; pass < 2; pass++) {
		int r = 32 - pass;
		double product = 1.0;
		for (int i = 0; i <= r - 1; i++)
			product *= ((1.0 - pow(2.0, i - 32)) * (1.0 - pow(2.0, i - 32))) / (1.0 - pow(2.0, i - r));
		double v = pow(2.0, r * (32 + 32 - r) - 32 * 32) * product;
		if (pass == 0) c->p_32 = v; else c->p_31 = v;
	}
	c->p_30 = 1.0 - (c->p_32 + c->p_31);
	if (c->lay.ntemplates > 0) {
		c->tmpl = malloc((size_t) c->lay.ntemplates * sizeof(uint32_t));
		count_templates(p->non_overlapping_m, c->tmpl);
	}
	for (int i = 1; i <= 4; i++)
		c->rex_pi[i - 1][0] = 1.0 - 1.0 / (2.0 * i);
	for (int j = 1; j < 5; j++)
		for (int i = 1; i <= 4; i++)
			c->rex_pi[i - 1][j] = 1.0 / (4.0 * i * i) * pow(c->rex_pi[i - 1][0], j - 1);
	for (int i = 1; i <= 4; i++)
		c->rex_pi[i - 1][5] = 1.0 / (2.0 * i) * pow(c->rex_pi[i - 1][0], 4);
}

/* per-thread scratch */
typedef struct {
	unsigned char *eps;	/* n bytes, one per bit */
	long *S;		/* n partial sums (random excursions) */
	long *hist;		/* pattern histogram scratch */
	double *X;		/* DFT input */
	fft_cplx *out;
	fft_plan *fft;
	unsigned char *bm_b, *bm_c, *bm_t;
	long *univ_T;
} oscratch;

/* utils/utilities.c:1837-1884 copyBitsToEpsilon: MSB first */
static void
unpack(const uint8_t *raw, int64_t n, unsigned char *eps)
{
	for (int64_t k = 0; k < n; k++)
		eps[k] = (unsigned char) ((raw[k >> 3] >> (7 - (k & 7))) & 1);
}

static double
bits_of(double d, int64_t *slot)
{
	memcpy(slot, &d, 8);
	return d;
}

/* ---- T1 frequency.c:191-211 ---- */
static void
t_frequency(const oconst *c, const unsigned char *e, double *pv, int64_t *st)
{
	long S_n = 0;
	for (int64_t i = 0; i < c->p.n; i++)
		S_n += e[i] ? 1 : -1;
	double s_obs = fabs((double) S_n) / c->sqrtn;
	pv[0] = erfc(s_obs / c->sqrt2);
	st[0] = S_n;
}

/* ---- T2 blockFrequency.c:210-247 ---- */
static void
t_block_frequency(const oconst *c, const unsigned char *e, double *pv, int64_t *st)
{
	long M = c->p.block_frequency_M, N = c->p.n / M;
	double sum = 0.0;
	st[0] = N;
	for (long i = 0; i < N; i++) {
		long blockSum = 0;
		for (long j = 0; j < M; j++)
			blockSum += e[j + i * M] ? 1 : 0;
		double pi = (double) blockSum / (double) M;
		double v = pi - 0.5;
		sum += v * v;
		st[1 + i] = blockSum;
	}
	double chi_squared = 4.0 * M * sum;
	pv[0] = oracle_igamc(N / 2.0, chi_squared / 2.0);
}

/* ---- T3 cusum.c:221-241 and compute_pi_value cusum.c:329-370 ---- */
static double
cusum_pi_value(const oconst *c, long z)
{
	long n = c->p.n;
	double sum1 = 0.0, sum2 = 0.0;
	for (long k = (-n / z + 1) / 4; k <= (n / z - 1) / 4; k++) {
		sum1 += o_normal(((4 * k + 1) * z) / c->sqrtn);
		sum1 -= o_normal(((4 * k - 1) * z) / c->sqrtn);
	}
	for (long k = (-n / z - 3) / 4; k <= (n / z - 1) / 4; k++) {
		sum2 += o_normal(((4 * k + 3) * z) / c->sqrtn);
		sum2 -= o_normal(((4 * k + 1) * z) / c->sqrtn);
	}
	return 1.0 - sum1 + sum2;
}

static void
t_cusum(const oconst *c, const unsigned char *e, double *pv, int64_t *st)
{
	long S = 0, S_max = 0, S_min = 0;
	for (int64_t k = 0; k < c->p.n; k++) {
		e[k] ? S++ : S--;
		if (S > S_max) S_max = S;
		if (S < S_min) S_min = S;
	}
	long z_b = (S_max - S > S - S_min) ? S_max - S : S - S_min;
	long z_f = (S_max > -S_min) ? S_max : -S_min;
	pv[0] = cusum_pi_value(c, z_f);		/* forward first: cusum.c:306-307 */
	pv[1] = cusum_pi_value(c, z_b);
	st[0] = S; st[1] = S_max; st[2] = S_min; st[3] = z_f; st[4] = z_b;
}

/* ---- T4 runs.c:213-246 (V_n is reported even when the test is not possible) ---- */
static void
t_runs(const oconst *c, const unsigned char *e, double *pv, int64_t *st)
{
	long n = c->p.n, S = 0, V_n = 1;
	for (long k = 0; k < n; k++)
		S += e[k] ? 1 : 0;
	for (long k = 1; k < n; k++)
		V_n += (e[k] != e[k - 1]);
	double pi = (double) S / (double) n;
	double two_over_sqrtn = 2.0 / c->sqrtn;			/* runs.c:139 */
	double sqrt2n = sqrt(2.0 * (double) n);				/* runs.c:138 */
	int possible = (fabs(pi - 0.5) >= two_over_sqrtn) ? 0 : 1;
	if (possible) {
		double arg = fabs(V_n - 2.0 * (double) n * pi * (1.0 - pi)) / (2.0 * pi * (1.0 - pi) * sqrt2n);
		pv[0] = erfc(arg);
	} else {
		pv[0] = NON_P_VALUE;
	}
	st[0] = S; st[1] = V_n; st[2] = possible;
}

/* ---- T5 longestRunOfOnes.c:148-210 (table), :352-421 ---- */
static const struct { long min_n, M; int min_class, max_class; double pi[7]; } lr_table[3] = {
	{ 128, 8, 1, 7, { 0.21484375, 0.3671875, 0.23046875, 0.109375, 0.046875, 0.01953125, 0.01171875 } },
	{ 6272, 128, 4, 10, { 0.11740357883779323143, 0.24295595927745485921, 0.24936348317907796964,
			      0.17517706034678234949, 0.10270117130405369391, 0.05521550943390406075,
			      0.05718333762093383558 } },
	{ 750000, 10000, 10, 16, { 0.08663231117995278587, 0.20820064838760340198, 0.24841858194169954122,
				   0.19391278674165693004, 0.12145848508900441468, 0.06801108930393995064,
				   0.07336609745614297557 } },
};

static void
t_longest_run(const oconst *c, const unsigned char *e, double *pv, int64_t *st)
{
	long n = c->p.n;
	int idx = 0;
	while (idx < 3 && n > lr_table[idx].min_n)
		++idx;
	if (idx >= 3)
		idx = 2;
	long M = lr_table[idx].M, N = n / M, count[7] = { 0 };
	int min_class = lr_table[idx].min_class, max_class = lr_table[idx].max_class;
	for (long i = 0; i < N; i++) {
		long v_obs = 0, run = 0;
		for (long j = 0; j < M; j++) {
			if (e[i * M + j] == 1) {
				run++;
				if (run > v_obs) v_obs = run;
			} else {
				run = 0;
			}
		}
		if (v_obs <= min_class) count[0]++;
		else if (v_obs <= max_class) count[v_obs - min_class]++;
		else count[6]++;
	}
	double chi2 = 0.0;
	for (int i = 0; i <= 6; i++) {
		double term = count[i] - (N * lr_table[idx].pi[i]);
		chi2 += term * term / (N * lr_table[idx].pi[i]);
		st[i] = count[i];
	}
	pv[0] = oracle_igamc(6 / 2.0, chi2 / 2.0);
}

/* ---- T6 rank.c:303-346, utils/matrix.c:49-200 (byte matrix, NIST elimination schedule) ---- */
static int
rank32(unsigned char A[32][32])
{
	/* forward, matrix.c:59-69 + :90-113 + :115-144 */
	for (int i = 0; i < 31; i++) {
		int have = A[i][i] == 1;
		if (!have) {
			int index = i + 1;
			while (index < 32 && A[index][i] == 0) index++;
			if (index < 32) {
				for (int p = 0; p < 32; p++) { unsigned char t = A[i][p]; A[i][p] = A[index][p]; A[index][p] = t; }
				have = 1;
			}
		}
		if (have)
			for (int j = i + 1; j < 32; j++)
				if (A[j][i] == 1)
					for (int k = i; k < 32; k++)
						A[j][k] = (unsigned char) ((A[j][k] + A[i][k]) % 2);
	}
	/* backward, matrix.c:74-84 */
	for (int i = 31; i > 0; i--) {
		int have = A[i][i] == 1;
		if (!have) {
			int index = i - 1;
			while (index >= 0 && A[index][i] == 0) index--;
			if (index >= 0) {
				for (int p = 0; p < 32; p++) { unsigned char t = A[i][p]; A[i][p] = A[index][p]; A[index][p] = t; }
				have = 1;
			}
		}
		if (have)
			for (int j = i - 1; j >= 0; j--)
				if (A[j][i] == 1)
					for (int k = 0; k < 32; k++)
						A[j][k] = (unsigned char) ((A[j][k] + A[i][k]) % 2);
	}
	/* matrix.c:162-188: rank = m - (all-zero rows) */
	int rank = 32;
	for (int i = 0; i < 32; i++) {
		int allz = 1;
		for (int j = 0; j < 32; j++)
			if (A[i][j] == 1) { allz = 0; break; }
		rank -= allz;
	}
	return rank;
}

static void
t_rank(const oconst *c, const unsigned char *e, double *pv, int64_t *st)
{
	long mc = c->p.n / 1024, F_M = 0, F_M1 = 0;
	unsigned char A[32][32];
	for (long k = 0; k < mc; k++) {
		for (int i = 0; i < 32; i++)		/* matrix.c:284-288 def_matrix */
			for (int j = 0; j < 32; j++)
				A[i][j] = e[k * 1024 + j + i * 32];
		int R = rank32(A);
		if (R == 32) F_M++;
		else if (R == 31) F_M1++;
	}
	long F_rem = mc - (F_M + F_M1);
	double chi = (((F_M - mc * c->p_32) * (F_M - mc * c->p_32) / (mc * c->p_32)) +
		      ((F_M1 - mc * c->p_31) * (F_M1 - mc * c->p_31) / (mc * c->p_31)) +
		      ((F_rem - mc * c->p_30) * (F_rem - mc * c->p_30) / (mc * c->p_30)));
	pv[0] = exp(-chi / 2.0);
	st[0] = F_M; st[1] = F_M1; st[2] = F_rem;
}

/* ---- T7 discreteFourierTransform.c:141-142, :326-428 (fftw branch) ---- */
static void
t_dft(const oconst *c, oscratch *s, const unsigned char *e, double *pv, int64_t *st)
{
	long n = c->p.n;
	double sqrtn4_095_005 = sqrt((double) n / 4.0 * 0.95 * 0.05);
	double sqrt_log20_n = sqrt(log(20.0) * (double) n);
	for (long i = 0; i < n; i++)
		s->X[i] = e[i] ? 1 : -1;
	fft_r2c(s->fft, s->X, s->out);
	double N_0 = (double) 0.95 * n / 2.0;
	long N_1 = 0;
	for (long i = 0; i < n / 2; i++) {
		double m = hypot(s->out[i].re, s->out[i].im);	/* cabs() */
		if (m < sqrt_log20_n)
			N_1++;
	}
	double d = (N_1 - N_0) / sqrtn4_095_005;
	pv[0] = erfc(fabs(d) / c->sqrt2);
	st[0] = N_1;
}

/* ---- T8 nonOverlappingTemplateMatchings.c:479-480, :495-575: greedy scan per template/block ---- */
static void
t_non_overlapping(const oconst *c, const unsigned char *e, double *pv, uint32_t *w)
{
	long n = c->p.n, M = n / 8;
	int m = c->p.non_overlapping_m;
	double mu = (M - m + 1) / ((double) ((long) 1 << m));
	double sigma_squared = M * (1.0 / ((double) ((long) 1 << m)) - (2.0 * m - 1.0) / ((double) ((long) 1 << m * 2)));
	unsigned char tpl[32];
	for (int jj = 0; jj < c->lay.ntemplates; jj++) {
		for (int k = 0; k < m; k++)
			tpl[k] = (unsigned char) ((c->tmpl[jj] >> (m - 1 - k)) & 1);
		double chi2 = 0.0;
		long Wj[8];
		for (int i = 0; i < 8; i++) {
			long W_obs = 0;
			for (long j = 0; j < M - m + 1; j++) {
				int match = 1;
				for (int k = 0; k < m; k++) {
					if (tpl[k] != e[i * M + j + k]) { match = 0; break; }
				}
				if (match) {
					W_obs++;
					j += m - 1;
				}
			}
			Wj[i] = W_obs;
			w[jj * 8 + i] = (uint32_t) W_obs;
		}
		for (int i = 0; i < 8; i++) {
			double term = ((double) Wj[i] - mu) / sqrt(sigma_squared);
			chi2 += term * term;
		}
		pv[jj] = oracle_igamc(8 / 2.0, chi2 / 2.0);
	}
}

/* ---- T9 overlappingTemplateMatchings.c:71-78, :255-303 ---- */
static const double ov_pi[6] = {
	0.36409105321672786245, 0.18565890010624038178, 0.13938113045903269914,
	0.10057114399877811497, 0.070432326346398449744, 0.13986544587282249192
};

static void
t_overlapping(const oconst *c, const unsigned char *e, double *pv, int64_t *st)
{
	long N = c->p.n / 1032, v[6] = { 0 };
	int m = c->p.overlapping_m;
	for (long i = 0; i < N; i++) {
		long W_obs = 0;
		for (long j = 0; j < 1032 - m + 1; j++) {
			int match = 1;
			for (int k = 0; k < m; k++)
				if (e[i * 1032 + j + k] != 1) { match = 0; break; }
			W_obs += match;
		}
		v[W_obs < 5 ? W_obs : 5]++;
	}
	double chi2 = 0.0;
	for (int i = 0; i < 6; i++) {
		double term = (double) v[i] - (double) N * ov_pi[i];
		chi2 += term * term / ((double) N * ov_pi[i]);
		st[i] = v[i];
	}
	pv[0] = oracle_igamc(5 / 2.0, chi2 / 2.0);
}

/* ---- T10 universal.c:72-80, :318-390 ---- */
static const double univ_expected[17] = { 0, 0, 0, 0, 0, 0, 5.2177052, 6.1962507, 7.1836656, 8.1764248,
	9.1723243, 10.170032, 11.168765, 12.168070, 13.167693, 14.167488, 15.167379 };
static const double univ_variance[17] = { 0, 0, 0, 0, 0, 0, 2.954, 3.125, 3.238, 3.311, 3.356, 3.384, 3.401,
	3.410, 3.416, 3.419, 3.421 };

static void
t_universal(const oconst *c, oscratch *s, const unsigned char *e, double *pv, int64_t *st)
{
	long L = c->lay.universal_L, p = 1L << L, Q = 10 * p, K = 100 * Q;
	long *T = s->univ_T;
	double sum = 0.0;
	memset(T, 0, (size_t) p * sizeof(T[0]));
	for (long i = 1; i <= Q; i++) {
		long dec = 0;
		for (long j = 0; j < L; j++)
			dec += e[(i - 1) * L + j] * (1L << (L - 1 - j));
		T[dec] = i;
	}
	for (long i = Q + 1; i <= Q + K; i++) {
		long dec = 0;
		for (long j = 0; j < L; j++)
			dec += e[(i - 1) * L + j] * (1L << (L - 1 - j));
		sum += log(i - T[dec]) / c->log2_;
		T[dec] = i;
	}
	double f_n = sum / (double) K;
	double cc = 0.7 - 0.8 / (double) L + (4 + 32 / (double) L) * pow(K, -3.0 / (double) L) / 15;
	double sigma = cc * sqrt(univ_variance[L] / (double) K);
	double arg = fabs(f_n - univ_expected[L]) / (c->sqrt2 * sigma);
	pv[0] = erfc(arg);
	bits_of(sum, &st[0]);
}

/* cyclic histogram of all n overlapping b-bit patterns: approximateEntropy.c:353-391, serial.c:374-412 */
static void
cyclic_hist(const unsigned char *e, long n, int b, long *h)
{
	long mask = (1 << b) - 1, dec = 0;
	memset(h, 0, (size_t) (1L << b) * sizeof(long));
	for (long i = 0; i < n + b; i++) {
		dec = ((dec << 1) + (int) e[i % n]) & mask;
		if (i >= b)
			h[dec]++;
	}
}

/* ---- T11 approximateEntropy.c:216-228, compute_phi :292-408 ---- */
static void
t_apen(const oconst *c, oscratch *s, const unsigned char *e, double *pv, int64_t *st)
{
	long n = c->p.n;
	int m = c->p.apen_m;
	double phi[2];
	for (int r = 0; r < 2; r++) {
		int b = m + r;
		long powLen = 1L << b;
		cyclic_hist(e, n, b, s->hist);
		double sum = 0.0;
		uint64_t sq = 0, lin = 0;
		for (long i = 0; i < powLen; i++) {
			if (s->hist[i])
				sum += (double) s->hist[i] * log(s->hist[i] / (double) n);
			sq += (uint64_t) s->hist[i] * (uint64_t) s->hist[i];
			lin += (uint64_t) i * (uint64_t) s->hist[i];
		}
		phi[r] = sum / (double) n;
		bits_of(phi[r], &st[r]);
		st[2 + r] = (int64_t) sq;
		st[4 + r] = (int64_t) lin;
	}
	double ApEn = phi[0] - phi[1];
	double chi_squared = 2.0 * n * (c->log2_ - ApEn);
	pv[0] = oracle_igamc((double) (1L << (m - 1)), chi_squared / 2.0);
}

/* ---- T12 randomExcursions.c:318-511 ---- */
static void
t_rnd_excursion(const oconst *c, oscratch *s, const unsigned char *e, double *pv, int64_t *st)
{
	long n = c->p.n, *S = s->S, J = 0;
	long v[6][8];
	memset(v, 0, sizeof(v));
	memset(st, 0, 57 * sizeof(int64_t));
	S[0] = e[0] ? 1 : -1;
	for (long i = 1; i < n; i++)
		S[i] = S[i - 1] + (e[i] ? 1 : -1);
	/* cycles end at every i >= 1 with S[i] == 0, plus n if S[n-1] != 0 (:324-347) */
	long cycleStart, cycleStop = 0;
	long counter[8];
	for (long i = 1; i <= n; i++) {
		int is_end = (i < n) ? (S[i] == 0) : (S[n - 1] != 0);
		if (!is_end)
			continue;
		J++;
		cycleStart = cycleStop;
		cycleStop = i;
		memset(counter, 0, sizeof(counter));
		for (long k = cycleStart; k < cycleStop; k++) {
			long a = labs(S[k]);
			if (a >= 1 && a <= 4)
				counter[S[k] + ((S[k] < 0) ? 4 : 3)]++;
		}
		for (int x = 0; x < 8; x++)
			v[counter[x] < 5 ? counter[x] : 5][x]++;
	}
	st[0] = J;
	if (J < c->min_zero_crossings) {
		for (int x = 0; x < 8; x++)
			pv[x] = NON_P_VALUE;
		/* v[][] is still reported (the reference never computes it in this case) */
	}
	for (int k = 0; k < 6; k++)
		for (int x = 0; x < 8; x++)
			st[1 + 8 * k + x] = v[k][x];
	if (J < c->min_zero_crossings)
		return;			/* stat.counter[] is cleared in this case (:581) */
	for (int x = 0; (c->p.flags & 1) && x < 8; x++)	/* stat.counter[] after the loop: the LAST cycle's visits (:386, :860-866) */
		st[49 + x] = counter[x];
	for (int i = 0; i < 8; i++) {
		long x = (i < 4) ? (i - 4) : (i - 3);
		long ax = labs(x);
		double chi2 = 0.0;
		for (int j = 0; j < 6; j++) {
			double term = (double) v[j][i] - ((double) J * c->rex_pi[ax - 1][j]);
			chi2 += term * term / ((double) J * c->rex_pi[ax - 1][j]);
		}
		pv[i] = oracle_igamc(5 / 2.0, chi2 / 2.0);
	}
}

/* ---- T13 randomExcursionsVariant.c:248-318 ---- */
static void
t_rnd_excursion_var(const oconst *c, oscratch *s, const unsigned char *e, double *pv, int64_t *st)
{
	long n = c->p.n, *S = s->S, J = 0, xi[18];
	memset(xi, 0, sizeof(xi));
	S[0] = e[0] ? 1 : -1;
	for (long j = 1; j < n; j++) {
		S[j] = S[j - 1] + (e[j] ? 1 : -1);
		if (S[j] == 0)
			J++;
	}
	if (S[n - 1] != 0)
		J++;
	for (long j = 0; j < n; j++) {
		long a = labs(S[j]);
		if (a >= 1 && a <= 9)
			xi[S[j] + ((S[j] < 0) ? 9 : 8)]++;
	}
	st[0] = J;
	for (int i = 0; i < 18; i++) {
		long x = (i < 9) ? (i - 9) : (i - 8);
		st[1 + i] = xi[i];
		if (J < c->min_zero_crossings)
			pv[i] = NON_P_VALUE;
		else
			pv[i] = erfc(labs(xi[i] - J) / (sqrt(2.0 * J * (4.0 * labs(x) - 2.0))));
	}
}

/* ---- T14 serial.c:205-224, compute_psi2 :313-426 ---- */
static void
t_serial(const oconst *c, oscratch *s, const unsigned char *e, double *pv, int64_t *st)
{
	long n = c->p.n;
	int m = c->p.serial_m;
	double psi[3];
	for (int r = 0; r < 3; r++) {
		int b = m - r;
		if (b == 0 || b == -1) {
			psi[r] = 0.0;
			st[r] = 0;
			continue;
		}
		long powLen = 1L << b;
		cyclic_hist(e, n, b, s->hist);
		double sum = 0.0;
		uint64_t isum = 0;
		for (long i = 0; i < powLen; i++) {
			sum += (double) s->hist[i] * (double) s->hist[i];
			isum += (uint64_t) s->hist[i] * (uint64_t) s->hist[i];
		}
		psi[r] = (sum * (double) (1L << b) / (double) n) - (double) n;
		st[r] = (int64_t) isum;
	}
	double del1 = psi[0] - psi[1];
	double del2 = psi[0] - 2.0 * psi[1] + psi[2];
	pv[0] = oracle_igamc((double) (1L << (m - 1)) / 2.0, del1 / 2.0);
	pv[1] = oracle_igamc((double) (1L << (m - 2)) / 2.0, del2 / 2.0);
}

/* ---- T15 linearComplexity.c:62, :283-390 (byte-array Berlekamp-Massey) ---- */
static const double lc_pi[7] = { 0.01047, 0.03125, 0.12500, 0.50000, 0.25000, 0.06250, 0.020833 };

static void
t_linear_complexity(const oconst *c, oscratch *s, const unsigned char *e, double *pv, int64_t *st)
{
	long M = c->p.linear_complexity_M, n = c->p.n, N = n / M, v[7] = { 0 };
	unsigned char *b_ = s->bm_b, *c_ = s->bm_c, *t_ = s->bm_t;
	/* linearComplexity.c:358: `1 << M` on an int; x86-64/gcc evaluates the shift count mod 32 */
	double two_pow = (double) (int) (1u << (M & 31));
	for (long i = 0; i < N; i++) {
		memset(b_, 0, (size_t) M);
		memset(c_, 0, (size_t) M);
		c_[0] = 1;
		b_[0] = 1;
		long L = 0, m = -1;
		for (long j = 0; j < M; j++) {
			int d = (int) e[i * M + j];
			for (long k = 1; k <= L; k++)
				d += c_[k] * e[i * M + j - k];
			d = d % 2;
			if (d == 1) {
				memcpy(t_, c_, (size_t) M);
				for (long k = j - m; k < M; k++)
					c_[k] = (unsigned char) ((c_[k] + b_[k - j + m]) % 2);
				if (L <= j / 2) {
					L = j + 1 - L;
					m = j;
					memcpy(b_, t_, (size_t) M);
				}
			}
		}
		double mean = (M / 2.0) + (((M + 1) % 2) ? 10 : 8) / 36.0 - (M / 3.0 + 2.0 / 9.0) / two_pow;
		double T = ((M % 2) ? (mean - L) : (L - mean)) + 2.0 / 9.0;
		double cls = (double) (6 - 1) / 2.0;
		if (T <= -cls) v[0]++;
		else if (T > cls) v[6]++;
		else v[(int) ceil(T + cls)]++;
	}
	double chi2 = 0.0;
	for (int i = 0; i < 7; i++) {
		chi2 += (v[i] - N * lc_pi[i]) * (v[i] - N * lc_pi[i]) / (N * lc_pi[i]);
		st[i] = v[i];
	}
	pv[0] = oracle_igamc(6 / 2.0, chi2 / 2.0);
}

/* ---------------------------------------------------------------------------------------------
 * driver: one stream at a time per thread, tests in enum order (driver.c:395-425)
 * ------------------------------------------------------------------------------------------- */
static void
run_stream(const oconst *c, oscratch *s, const uint8_t *raw, double *pv, int64_t *st, uint32_t *w)
{
	const oracle_layout *l = &c->lay;
	uint32_t en = l->enabled_mask;
	unpack(raw, c->p.n, s->eps);
	const unsigned char *e = s->eps;
#define ON(t) (en & (1u << (t)))
#define PV(t) (pv + l->pv_off[t])
#define ST(t) (st + l->st_off[t])
	if (ON(1)) t_frequency(c, e, PV(1), ST(1));
	if (ON(2)) t_block_frequency(c, e, PV(2), ST(2));
	if (ON(3)) t_cusum(c, e, PV(3), ST(3));
	if (ON(4)) t_runs(c, e, PV(4), ST(4));
	if (ON(5)) t_longest_run(c, e, PV(5), ST(5));
	if (ON(6)) t_rank(c, e, PV(6), ST(6));
	if (ON(7)) t_dft(c, s, e, PV(7), ST(7));
	if (ON(8)) t_non_overlapping(c, e, PV(8), w);
	if (ON(9)) t_overlapping(c, e, PV(9), ST(9));
	if (ON(10)) t_universal(c, s, e, PV(10), ST(10));
	if (ON(11)) t_apen(c, s, e, PV(11), ST(11));
	if (ON(12)) t_rnd_excursion(c, s, e, PV(12), ST(12));
	if (ON(13)) t_rnd_excursion_var(c, s, e, PV(13), ST(13));
	if (ON(14)) t_serial(c, s, e, PV(14), ST(14));
	if (ON(15)) t_linear_complexity(c, s, e, PV(15), ST(15));
#undef ON
#undef PV
#undef ST
}

typedef struct {
	const oconst *c;
	const uint8_t *raw;
	int64_t nstreams;
	double *pv;
	int64_t *st;
	uint32_t *w;
	int64_t *next;
	pthread_mutex_t *mu;
} ojob;

static void *
worker(void *arg)
{
	ojob *j = arg;
	const oconst *c = j->c;
	int64_t n = c->p.n;
	uint32_t en = c->lay.enabled_mask;
	oscratch s;
	memset(&s, 0, sizeof(s));
	s.eps = malloc((size_t) n);
	if (en & ((1u << 12) | (1u << 13)))
		s.S = malloc((size_t) n * sizeof(long));
	int hb = 0;
	if (en & (1u << 11)) hb = c->p.apen_m + 1;
	if ((en & (1u << 14)) && c->p.serial_m > hb) hb = c->p.serial_m;
	if (hb > 0)
		s.hist = malloc(((size_t) 1 << hb) * sizeof(long));
	if (en & (1u << 7)) {
		s.X = malloc((size_t) n * sizeof(double));
		s.out = malloc((size_t) (n / 2 + 1) * sizeof(fft_cplx));
		s.fft = fft_plan_create(n);
	}
	if (en & (1u << 15)) {
		s.bm_b = malloc((size_t) c->p.linear_complexity_M);
		s.bm_c = malloc((size_t) c->p.linear_complexity_M);
		s.bm_t = malloc((size_t) c->p.linear_complexity_M);
	}
	if (en & (1u << 10))
		s.univ_T = malloc(((size_t) 1 << c->lay.universal_L) * sizeof(long));
	for (;;) {
		pthread_mutex_lock(j->mu);
		int64_t it = (*j->next)++;
		pthread_mutex_unlock(j->mu);
		if (it >= j->nstreams)
			break;
		run_stream(c, &s, j->raw + it * (n / 8), j->pv + it * c->lay.npv, j->st + it * c->lay.nst,
			   j->w ? j->w + it * c->lay.nw : NULL);
	}
	free(s.eps); free(s.S); free(s.hist); free(s.X); free(s.out); fft_plan_destroy(s.fft);
	free(s.bm_b); free(s.bm_c); free(s.bm_t); free(s.univ_T);
	return NULL;
}

int
oracle_run(const oracle_params *p, const uint8_t *raw, int64_t nstreams, double *pv, int64_t *st, uint32_t *w,
	   int nthreads)
{
	if (p == NULL || raw == NULL || p->n <= 0 || (p->n % 8) != 0)
		return -1;
	oconst c;
	oconst_init(&c, p);
	if (nthreads < 1) nthreads = 1;
	if (nthreads > nstreams) nthreads = (int) nstreams;
	if (nthreads > 512) nthreads = 512;
	pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
	int64_t next = 0;
	ojob job = { &c, raw, nstreams, pv, st, w, &next, &mu };
	pthread_t th[512];
	for (int i = 0; i < nthreads; i++)
		pthread_create(&th[i], NULL, worker, &job);
	for (int i = 0; i < nthreads; i++)
		pthread_join(th[i], NULL);
	free(c.tmpl);
	return 0;
}

int
oracle_pattern_histogram(const uint8_t *raw, int64_t n, int bits, uint32_t *hist)
{
	unsigned char *eps = malloc((size_t) n);
	long *h = malloc(((size_t) 1 << bits) * sizeof(long));
	if (!eps || !h)
		return -1;
	unpack(raw, n, eps);
	cyclic_hist(eps, n, bits, h);
	for (long i = 0; i < (1L << bits); i++)
		hist[i] = (uint32_t) h[i];
	free(eps);
	free(h);
	return 0;
}

/*
 * tools/generators.c:797-831 (lcg_rand in doubles, exact below 2^53), :870-905 (bit = DUNIF >= 0.5),
 * write_sequence (:1612-1660) packs LSB-first: file byte j holds generator bits 8j..8j+7 at bit
 * positions 0..7.  The state carries across streams.
 */
void
oracle_lcg_bytes(int64_t first_stream, int64_t nstreams, int64_t n, uint8_t *out)
{
	uint64_t x = 23482349ULL;
	const uint64_t a = 950706376ULL, m = 2147483647ULL;
	for (int64_t k = 0; k < first_stream * n; k++)
		x = (x * a) % m;
	memset(out, 0, (size_t) (nstreams * n / 8));
	for (int64_t s = 0; s < nstreams; s++) {
		for (int64_t k = 0; k < n; k++) {
			x = (x * a) % m;
			/* DUNIF = x / (2^31-1) >= 0.5  <=>  x >= 2^30 */
			if (x >= (1ULL << 30))
				out[s * (n / 8) + (k >> 3)] |= (uint8_t) (1u << (k & 7));
		}
	}
}

/* ------------------------------------------------------------------------------------------ */
/*
 * Assessment of one run: frequency.c:828-894 (tally) and :631-700, :739-757 (thresholds, uniformity,
 * verdict); the other 14 X_metrics are the same code with their own test number.
 */
int
oracle_assess(const oracle_layout *lay, const double *pv, int64_t iterations, int64_t bins, double alpha,
	      double uniformity_level, int64_t *tally, int64_t *freq, double *fp, int32_t *flags)
{
	if (lay == NULL || pv == NULL || bins < 1 || iterations < 0)
		return -1;
	for (int t = 1; t <= 15; t++) {
		if (!((lay->enabled_mask >> t) & 1u))
			continue;
		const int is_excursion = (t == 12 || t == 13);		/* driver.c is_excursion[] */
		for (int j = 0; j < lay->pv_cnt[t]; j++) {
			const int q = lay->pv_off[t] + j;
			long sampleCount = 0, toolow = 0;
			int64_t *f = freq + (int64_t) q * bins;
			memset(f, 0, (size_t) bins * sizeof(f[0]));
			for (int64_t i = 0; i < iterations; i++) {
				const double p_value = pv[i * lay->npv + q];
				if (p_value == -99999999.0)			/* NON_P_VALUE, defs.h:314 */
					continue;
				if (is_excursion) {
					if (p_value > 0.0)
						++sampleCount;
					else
						continue;
				} else {
					++sampleCount;
				}
				if (p_value < alpha)
					++toolow;
				if (p_value >= 1.0)
					++f[bins - 1];
				else if (p_value >= 0.0)
					++f[(int) floor(p_value * (double) bins)];
				else
					++f[0];
			}
			long passCount = ((sampleCount <= 0) || (sampleCount < toolow)) ? 0 : sampleCount - toolow;
			const double p_hat = 1.0 - alpha;
			const double tmax = (p_hat + 3.0 * sqrt((p_hat * alpha) / sampleCount)) * sampleCount;
			const double tmin = (p_hat - 3.0 * sqrt((p_hat * alpha) / sampleCount)) * sampleCount;
			double chi2 = 0.0, uniformity;
			const double expCount = (double) sampleCount / bins;
			if (expCount <= 0.0) {
				uniformity = 0.0;
			} else {
				for (int64_t i = 0; i < bins; ++i)
					chi2 += (f[i] - expCount) * (f[i] - expCount) / expCount;
				uniformity = oracle_igamc((bins - 1.0) / 2.0, chi2 / 2.0);
			}
			const int uniformity_passed = !(expCount <= 0.0 || uniformity < uniformity_level);
			const int proportion_passed = !(sampleCount == 0 || (passCount < tmin) || (passCount > tmax));
			tally[3 * q] = sampleCount; tally[3 * q + 1] = toolow; tally[3 * q + 2] = passCount;
			fp[4 * q] = tmin; fp[4 * q + 1] = tmax; fp[4 * q + 2] = chi2; fp[4 * q + 3] = uniformity;
			flags[q] = uniformity_passed | (proportion_passed << 1);
		}
	}
	return 0;
}
