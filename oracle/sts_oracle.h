/*
 * sts_oracle.h - CPU restatement of the per-bitstream test kernels of arcetri/sts v3.2.7.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (sts_b200/) never links or calls it.
 * It is pinned against the unmodified reference binary built into oracle/_ref/sts
 * (tests/golden/make_golden.py generated tests/golden from that binary; tests/test_oracle_golden.py
 * checks the restatement against those vectors).
 *
 * Parity status: 14 of 15 tests pinned to the reference binary's own output.  The DFT test's
 * arithmetic lives in libfftw3, which is not vendored by the reference and not installed here:
 * "parity unpinned" for the FFT itself; N_1 is pinned against the reference linked with the
 * FP64 FFT of oracle/fft_f64.c and cross-checked against numpy.fft.rfft.
 */
#ifndef STS_ORACLE_H
#define STS_ORACLE_H
#include <stdint.h>

typedef struct oracle_params {
	int64_t n;			/* tp.n */
	uint32_t test_mask;		/* bit t = test t enabled */
	int32_t flags;			/* 1: also report RandomExcursions' stat.counter[] (istats 49..56), like STSGPU_FLAG_STATS */
	int64_t block_frequency_M;
	int32_t non_overlapping_m;
	int32_t overlapping_m;
	int32_t apen_m;
	int32_t serial_m;
	int64_t linear_complexity_M;
} oracle_params;

/* same record layout as include/stsgpu.h (restated independently in sts_oracle.c) */
typedef struct oracle_layout {
	uint32_t enabled_mask;
	int32_t npv, nst, nw, ntemplates, universal_L;
	int32_t pv_off[16], pv_cnt[16], st_off[16], st_cnt[16];
} oracle_layout;

int oracle_layout_for(const oracle_params *p, oracle_layout *out);

/*
 * Run all enabled tests on nstreams streams of raw file bytes (MSB-first bits), using up to
 * nthreads host threads (one stream per thread at a time, like the reference's -T).
 * Output arrays are stream-major: pv[nstreams][npv], st[nstreams][nst], w[nstreams][nw].
 */
int oracle_run(const oracle_params *p, const uint8_t *raw, int64_t nstreams,
	       double *pv, int64_t *st, uint32_t *w, int nthreads);

/* cyclic pattern histogram used by Serial/ApEn (serial.c:390-412), for parity tests */
int oracle_pattern_histogram(const uint8_t *raw, int64_t n, int bits, uint32_t *hist);

/* the LCG of tools/generators.c:797-905, byte-identical file contents */
void oracle_lcg_bytes(int64_t first_stream, int64_t nstreams, int64_t n, uint8_t *out);

/*
 * The uniformity / proportion assessment every X_metrics applies to one column of p-values
 * (e.g. frequency.c:828-894 tally, :631-700 thresholds and uniformity; the loop is identical in all
 * 15 tests).  pv is iteration-major [iterations][npv]; columns of the excursion tests (12, 13) sample
 * only p-values > 0.  Per column q: tally[q] = { sampleCount, toolow, passCount }, freq[q][bins],
 * fp[q] = { proportion_threshold_min, proportion_threshold_max, chi2, uniformity },
 * flags[q] = uniformity_passed | proportion_passed << 1  (the non-legacy branch of :739-757).
 */
int oracle_assess(const oracle_layout *lay, const double *pv, int64_t iterations, int64_t bins, double alpha,
		  double uniformity_level, int64_t *tally, int64_t *freq, double *fp, int32_t *flags);

/* p-value tails, exported for unit tests */
double oracle_igamc(double a, double x);
double oracle_igam(double a, double x);

#endif
