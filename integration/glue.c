/*
 * glue.c - the reference-side glue that drops libstsgpu.so (include/stsgpu.h) into arcetri/sts v3.2.7.
 *
 * integration/Makefile compiles the reference's OWN sources where they lie (REF ?= /root/reference) - sts.c,
 * parse_args.c, driver.c with its testDriver[] table and iterate(), every X_init / X_print / X_metrics / X_destroy,
 * write_p_val_to_file, read_from_p_val_file, metrics() - and renames exactly sixteen of their functions away with -D:
 * the fifteen X_iterate bodies of src/tests (*.c) and invokeTestSuite of src/utils/utilities.c:1434.  This file
 * defines those sixteen symbols over the engine, so the unmodified driver.c:63-192 table and driver.c:395-425
 * iterate() call into it:
 *
 *   invokeTestSuite()  replaces utilities.c:1434-1476 + handleFileBasedBitStreams :1480-1578 + testBits :1581-1639 +
 *                      parseBitsBinaryInput / parseBitsASCIIInput / copyBitsToEpsilon :1651-1884: batches of streams go
 *                      to stsgpu_submit / stsgpu_submit_ascii as the raw file bytes (three batches in flight), and for
 *                      every completed stream the reference's iterate() is called, in iteration order, mutex NULL.
 *   X_iterate()        the compute part is gone; what remains is the record keeping tail of each reference body
 *                      (e.g. frequency.c:218-260) over the record the GPU produced, and with -s the test's private
 *                      stats struct (layout-compatible restatements below; sizes are checked against the dyn_arrays
 *                      the reference's own X_init created).
 *
 * `struct state` cannot be edited from here, so the engine handle lives in a file-static.  There is no CPU path:
 * the renamed CPU bodies are not linked (--gc-sections, checked by the tests), and every engine failure ends in the
 * reference's err() with a code of the range of the file it replaces.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <math.h>
#include <pthread.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include "utils/externs.h"
#include "utils/utilities.h"
#include "utils/debug.h"
#include "utils/dyn_alloc.h"
#include "utils/stat_fncs.h"
#include <stsgpu.h>

#define GLUE_DEPTH 3			/* batches in flight: read + upload of one overlap the kernels of the others */
#define GLUE_BATCH 1024			/* streams per batch (STS_GPU_BATCH in the environment overrides) */
#define GLUE_MAX_READERS 64

static struct {
	stsgpu_ctx *engine;
	stsgpu_layout lay;
	/* records of the batch being iterated */
	const double *pv;
	const int64_t *st;
	const uint32_t *w;
	const double *fs;
	long int first, count;
} G;

/* ------------------------------------------------------------------------------------------------------------ */
/* private stats structs: same members in the same order as the file-private definitions they mirror */

struct Frequency_stats { bool success; long int S_n; };				/* frequency.c:43-46 */
struct BlockFrequency_stats { bool success; double chi_squared; };		/* blockFrequency.c:45-48 */
struct CumulativeSums_stats {							/* cusum.c:45-50 */
	bool success_forward; bool success_backward; long int z_forward; long int z_backward;
};
struct Runs_stats { bool success; double pi; long int V_n; double erfc_arg; bool test_possible; };	/* runs.c:45-51 */
struct LongestRunOfOnes_stats {							/* longestRunOfOnes.c:45-60 */
	bool success; long int N; long int M; double chi2; int runs_table_index;
	unsigned long count[CLASS_COUNT_LONGEST_RUN + 1];
};
struct Rank_stats { bool success; double chi_squared; long int F_M; long int F_M_minus_one; long int F_remaining; };	/* rank.c:47-53 */
struct DiscreteFourierTransform_stats { bool success; long int N_1; double N_0; double d; };	/* discreteFourierTransform.c:52-58 */
struct NonOverlappingTemplateMatchings_stats { double mu; double sigma_squared; long int M; };	/* nonOverlapping...c:45-49 */
struct OverlappingTemplateMatchings_stats {					/* overlapping...c:46-52 */
	bool success; long int N; long int v[K_OVERLAPPING + 1]; double chi2;
};
struct Universal_stats { bool success; long int Q; long int K; double sum; double sigma; double f_n; };	/* universal.c:46-53 */
struct ApproximateEntropy_stats { bool success; double phi[1 + 1]; double ApEn; double chi_squared; };	/* approximateEntropy.c:45-50 */
struct RandomExcursions_stats {							/* randomExcursions.c:45-51 */
	bool success[NUMBER_OF_STATES_RND_EXCURSION]; long int counter[NUMBER_OF_STATES_RND_EXCURSION];
	double chi2[NUMBER_OF_STATES_RND_EXCURSION]; bool test_possible; long int number_of_cycles;
};
struct RandomExcursionsVariant_stats {						/* randomExcursionsVariant.c:45-50 */
	bool success[NUMBER_OF_STATES_RND_EXCURSION_VAR]; long int counter[NUMBER_OF_STATES_RND_EXCURSION_VAR];
	bool test_possible; long int number_of_cycles;
};
struct Serial_stats { bool success1; bool success2; double psim0; double psim1; double psim2; double del1; double del2; };	/* serial.c:45-53 */
struct LinearComplexity_stats { bool success; long int v[K_LINEARCOMPLEXITY + 1]; double chi2; };	/* linearComplexity.c:45-49 */

static const size_t stats_size[NUMOFTESTS + 1] = {
	0, sizeof(struct Frequency_stats), sizeof(struct BlockFrequency_stats), sizeof(struct CumulativeSums_stats),
	sizeof(struct Runs_stats), sizeof(struct LongestRunOfOnes_stats), sizeof(struct Rank_stats),
	sizeof(struct DiscreteFourierTransform_stats), sizeof(struct NonOverlappingTemplateMatchings_stats),
	sizeof(struct OverlappingTemplateMatchings_stats), sizeof(struct Universal_stats),
	sizeof(struct ApproximateEntropy_stats), sizeof(struct RandomExcursions_stats),
	sizeof(struct RandomExcursionsVariant_stats), sizeof(struct Serial_stats), sizeof(struct LinearComplexity_stats),
};

/* ------------------------------------------------------------------------------------------------------------ */
/* what every X_iterate begins and ends with */

/* the firewall of e.g. frequency.c:161-185 minus the epsilon checks (nothing reads epsilon any more) */
static struct state *
enter(struct thread_state *thread_state, int test_num, int code, const char *name)
{
	if (thread_state == NULL) {
		err(code, name, "thread_state arg is NULL");
	}
	struct state *state = thread_state->global_state;
	if (state == NULL) {
		err(code, name, "state arg is NULL");
	}
	if (state->testVector[test_num] != true) {
		dbg(DBG_LOW, "iterate driver interface for %s[%d] called when test vector was false", state->testNames[test_num],
		    test_num);
		return NULL;
	}
	if (state->cSetup != true) {
		err(code, name, "test constants not setup prior to calling %s for %s[%d]", name, state->testNames[test_num], test_num);
	}
	if (G.engine == NULL) {
		err(code, name, "%s called outside the GPU batch loop (the GPU build has no CPU iterate)", name);
	}
	const long int k = thread_state->iteration_being_done - G.first;
	if (k < 0 || k >= G.count) {
		err(code, name, "iteration %ld is not in the batch [%ld, %ld) the GPU just completed",
		    thread_state->iteration_being_done, G.first, G.first + G.count);
	}
	if (thread_state->mutex != NULL) {
		pthread_mutex_lock(thread_state->mutex);
	}
	return state;
}

static void
leave(struct thread_state *thread_state)
{
	if (thread_state->mutex != NULL) {
		pthread_mutex_unlock(thread_state->mutex);
	}
}

#define REC(ts) ((ts)->iteration_being_done - G.first)
#define PV(ts, t) (G.pv + REC(ts) * G.lay.npv + G.lay.pv_off[t])
#define ST(ts, t) (G.st + REC(ts) * G.lay.nst + G.lay.st_off[t])
#define FS(ts, t) (G.fs + REC(ts) * G.lay.nfs + G.lay.fs_off[t])	/* only with -s */

/*
 * "Record success or failure for this iteration" (frequency.c:228-246 and its copies): returns stat.success.
 * `what` and `which` reproduce the wording of the individual warnings ("backward test", "p_value1", "template[3] of").
 */
static bool
keep(struct thread_state *thread_state, int test_num, double p_value, const char *name, const char *what, const char *which)
{
	struct state *state = thread_state->global_state;

	state->count[test_num]++;	// Count this iteration
	state->valid[test_num]++;	// Count this valid iteration
	if (isNegative(p_value)) {
		state->failure[test_num]++;	// Bogus p_value < 0.0 treated as a failure
		warn(name, "iteration %ld %s %s[%d] produced bogus %s: %f < 0.0\n", thread_state->iteration_being_done + 1,
		     what, state->testNames[test_num], test_num, which, p_value);
		return false;
	} else if (isGreaterThanOne(p_value)) {
		state->failure[test_num]++;	// Bogus p_value > 1.0 treated as a failure
		warn(name, "iteration %ld %s %s[%d] produced bogus %s: %f > 1.0\n", thread_state->iteration_being_done + 1,
		     what, state->testNames[test_num], test_num, which, p_value);
		return false;
	} else if (p_value < state->tp.alpha) {
		state->valid_p_val[test_num]++;	// Valid p_value in [0.0, 1.0] range
		state->failure[test_num]++;	// Valid p_value but too low is a failure
		return false;
	}
	state->valid_p_val[test_num]++;	// Valid p_value in [0.0, 1.0] range
	state->success[test_num]++;	// Valid p_value not too low is a success
	return true;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* the fifteen table entries (driver.c:63-192), in enum order */

void
Frequency_iterate(struct thread_state *thread_state)		/* frequency.c:150-263 */
{
	static const enum test test_num = TEST_FREQUENCY;
	struct state *state = enter(thread_state, test_num, 71, __func__);
	if (state == NULL)
		return;
	struct Frequency_stats stat;
	double p_value = PV(thread_state, test_num)[0];
	stat.S_n = ST(thread_state, test_num)[0];
	stat.success = keep(thread_state, test_num, p_value, __func__, "of test", "p_value");
	if (state->resultstxtFlag == true) {
		append_value(state->stats[test_num], &stat);
	}
	append_value(state->p_val[test_num], &p_value);
	leave(thread_state);
}

void
BlockFrequency_iterate(struct thread_state *thread_state)	/* blockFrequency.c:172-296 */
{
	static const enum test test_num = TEST_BLOCK_FREQUENCY;
	struct state *state = enter(thread_state, test_num, 21, __func__);
	if (state == NULL)
		return;
	struct BlockFrequency_stats stat;
	double p_value = PV(thread_state, test_num)[0];
	stat.chi_squared = state->resultstxtFlag ? FS(thread_state, test_num)[0] : UNSET_DOUBLE;
	stat.success = keep(thread_state, test_num, p_value, __func__, "of test", "p_value");
	if (state->resultstxtFlag == true) {
		append_value(state->stats[test_num], &stat);
	}
	append_value(state->p_val[test_num], &p_value);
	leave(thread_state);
}

void
CumulativeSums_iterate(struct thread_state *thread_state)	/* cusum.c:162-317 */
{
	static const enum test test_num = TEST_CUSUM;
	struct state *state = enter(thread_state, test_num, 31, __func__);
	if (state == NULL)
		return;
	struct CumulativeSums_stats stat;
	double p_value_forward = PV(thread_state, test_num)[0];
	double p_value_backward = PV(thread_state, test_num)[1];
	stat.z_forward = ST(thread_state, test_num)[3];
	stat.z_backward = ST(thread_state, test_num)[4];
	stat.success_forward = keep(thread_state, test_num, p_value_forward, __func__, "of test", "p_value");
	stat.success_backward = keep(thread_state, test_num, p_value_backward, __func__, "of backward test", "p_value");
	if (state->resultstxtFlag == true) {
		append_value(state->stats[test_num], &stat);
	}
	append_value(state->p_val[test_num], &p_value_forward);		/* forward first, cusum.c:306-307 */
	append_value(state->p_val[test_num], &p_value_backward);
	leave(thread_state);
}

void
Runs_iterate(struct thread_state *thread_state)			/* runs.c:172-333 */
{
	static const enum test test_num = TEST_RUNS;
	struct state *state = enter(thread_state, test_num, 181, __func__);
	if (state == NULL)
		return;
	struct Runs_stats stat;
	double p_value = PV(thread_state, test_num)[0];
	stat.test_possible = ST(thread_state, test_num)[2] != 0;
	if (stat.test_possible == true) {
		stat.pi = state->resultstxtFlag ? FS(thread_state, test_num)[0] : UNSET_DOUBLE;
		stat.V_n = ST(thread_state, test_num)[1];
		stat.erfc_arg = state->resultstxtFlag ? FS(thread_state, test_num)[1] : UNSET_DOUBLE;
		stat.success = keep(thread_state, test_num, p_value, __func__, "of test", "p_value");
	} else {
		state->count[test_num]++;	/* counted, but not valid (runs.c:292-301) */
		stat.pi = UNSET_DOUBLE;
		stat.V_n = 0;
		stat.erfc_arg = UNSET_DOUBLE;
		stat.success = false;
		p_value = NON_P_VALUE;
	}
	if (state->resultstxtFlag == true) {
		append_value(state->stats[test_num], &stat);
	}
	append_value(state->p_val[test_num], &p_value);
	leave(thread_state);
}

void
LongestRunOfOnes_iterate(struct thread_state *thread_state)	/* longestRunOfOnes.c:309-471 */
{
	static const enum test test_num = TEST_LONGEST_RUN;
	static const long int min_n[3] = { 128, 6272, 750000 }, table_M[3] = { 8, 128, 10000 };	/* runs_table[], :148-210 */
	struct state *state = enter(thread_state, test_num, 111, __func__);
	if (state == NULL)
		return;
	struct LongestRunOfOnes_stats stat;
	double p_value = PV(thread_state, test_num)[0];
	stat.runs_table_index = 0;					/* the lookup of :352-360 */
	while (stat.runs_table_index < 3 && state->tp.n > min_n[stat.runs_table_index])
		++stat.runs_table_index;
	if (stat.runs_table_index >= 3)
		stat.runs_table_index = 2;
	stat.M = table_M[stat.runs_table_index];
	stat.N = state->tp.n / stat.M;
	stat.chi2 = state->resultstxtFlag ? FS(thread_state, test_num)[0] : UNSET_DOUBLE;
	for (int i = 0; i <= CLASS_COUNT_LONGEST_RUN; i++)
		stat.count[i] = (unsigned long) ST(thread_state, test_num)[i];
	stat.success = keep(thread_state, test_num, p_value, __func__, "of test", "p_value");
	if (state->resultstxtFlag == true) {
		append_value(state->stats[test_num], &stat);
	}
	append_value(state->p_val[test_num], &p_value);
	leave(thread_state);
}

void
Rank_iterate(struct thread_state *thread_state)			/* rank.c:233-396 */
{
	static const enum test test_num = TEST_RANK;
	struct state *state = enter(thread_state, test_num, 171, __func__);
	if (state == NULL)
		return;
	struct Rank_stats stat;
	double p_value = PV(thread_state, test_num)[0];
	stat.F_M = ST(thread_state, test_num)[0];
	stat.F_M_minus_one = ST(thread_state, test_num)[1];
	stat.F_remaining = ST(thread_state, test_num)[2];
	stat.chi_squared = state->resultstxtFlag ? FS(thread_state, test_num)[0] : UNSET_DOUBLE;
	stat.success = keep(thread_state, test_num, p_value, __func__, "of test", "p_value");
	if (state->resultstxtFlag == true) {
		append_value(state->stats[test_num], &stat);
	}
	append_value(state->p_val[test_num], &p_value);
	leave(thread_state);
}

void
DiscreteFourierTransform_iterate(struct thread_state *thread_state)	/* discreteFourierTransform.c:236-471 */
{
	static const enum test test_num = TEST_DFT;
	struct state *state = enter(thread_state, test_num, 41, __func__);
	if (state == NULL)
		return;
	struct DiscreteFourierTransform_stats stat;
	double p_value = PV(thread_state, test_num)[0];
	stat.N_1 = ST(thread_state, test_num)[0];
	stat.N_0 = state->resultstxtFlag ? FS(thread_state, test_num)[0] : UNSET_DOUBLE;
	stat.d = state->resultstxtFlag ? FS(thread_state, test_num)[1] : UNSET_DOUBLE;
	stat.success = keep(thread_state, test_num, p_value, __func__, "of test", "p_value");
	if (state->resultstxtFlag == true) {
		append_value(state->stats[test_num], &stat);
	}
	append_value(state->p_val[test_num], &p_value);
	leave(thread_state);
}

void
NonOverlappingTemplateMatchings_iterate(struct thread_state *thread_state)	/* nonOverlapping...c:419-650 */
{
	static const enum test test_num = TEST_NON_OVERLAPPING;
	struct state *state = enter(thread_state, test_num, 131, __func__);
	if (state == NULL)
		return;
	struct NonOverlappingTemplateMatchings_stats stat;
	const long int m = state->tp.nonOverlappingTemplateLength;
	const double *pv = PV(thread_state, test_num);
	const uint32_t *W = G.w + REC(thread_state) * G.lay.nw;
	char which[48];

	stat.M = state->tp.n / BLOCKS_NON_OVERLAPPING;			/* :472-480 */
	stat.mu = (stat.M - m + 1) / ((double) ((long int) 1 << m));
	stat.sigma_squared = stat.M * (1.0 / ((double) ((long int) 1 << m)) - (2.0 * m - 1.0) / ((double) ((long int) 1 << m * 2)));
	for (long int jj = 0; jj < G.lay.ntemplates; jj++) {		/* template order, :589-628 */
		struct nonover_stats nonover_stat;
		memset(&nonover_stat, 0, sizeof(nonover_stat));
		nonover_stat.p_value = pv[jj];
		nonover_stat.chi2 = state->resultstxtFlag ? FS(thread_state, test_num)[jj] : UNSET_DOUBLE;
		nonover_stat.template_index = jj;
		for (int i = 0; i < BLOCKS_NON_OVERLAPPING; i++)
			nonover_stat.Wj[i] = W[jj * BLOCKS_NON_OVERLAPPING + i];
		snprintf(which, sizeof(which), "template[%ld] of test", jj);
		nonover_stat.success = keep(thread_state, test_num, nonover_stat.p_value, __func__, which, "p_value");
		append_value(state->p_val[test_num], &nonover_stat);	/* the one test whose p_val holds structs (defs.h:346-352) */
	}
	if (state->resultstxtFlag == true) {
		append_value(state->stats[test_num], &stat);
	}
	leave(thread_state);
}

void
OverlappingTemplateMatchings_iterate(struct thread_state *thread_state)	/* overlapping...c:212-360 */
{
	static const enum test test_num = TEST_OVERLAPPING;
	struct state *state = enter(thread_state, test_num, 141, __func__);
	if (state == NULL)
		return;
	struct OverlappingTemplateMatchings_stats stat;
	double p_value = PV(thread_state, test_num)[0];
	stat.N = state->tp.n / BLOCK_LENGTH_OVERLAPPING;
	for (int i = 0; i <= K_OVERLAPPING; i++)
		stat.v[i] = ST(thread_state, test_num)[i];
	stat.chi2 = state->resultstxtFlag ? FS(thread_state, test_num)[0] : UNSET_DOUBLE;
	stat.success = keep(thread_state, test_num, p_value, __func__, "of test", "p_value");
	if (state->resultstxtFlag == true) {
		append_value(state->stats[test_num], &stat);
	}
	append_value(state->p_val[test_num], &p_value);
	leave(thread_state);
}

void
Universal_iterate(struct thread_state *thread_state)		/* universal.c:258-438 */
{
	static const enum test test_num = TEST_UNIVERSAL;
	struct state *state = enter(thread_state, test_num, 201, __func__);
	if (state == NULL)
		return;
	struct Universal_stats stat;
	double p_value = PV(thread_state, test_num)[0];
	stat.Q = 10 * ((long int) 1 << state->universal_L);		/* :318-325 */
	stat.K = 100 * stat.Q;
	memcpy(&stat.sum, &ST(thread_state, test_num)[0], sizeof(double));	/* the double `sum`, bit for bit */
	stat.sigma = state->resultstxtFlag ? FS(thread_state, test_num)[0] : UNSET_DOUBLE;
	stat.f_n = state->resultstxtFlag ? FS(thread_state, test_num)[1] : UNSET_DOUBLE;
	stat.success = keep(thread_state, test_num, p_value, __func__, "of test", "p_value");
	if (state->resultstxtFlag == true) {
		append_value(state->stats[test_num], &stat);
	}
	append_value(state->p_val[test_num], &p_value);
	leave(thread_state);
}

void
ApproximateEntropy_iterate(struct thread_state *thread_state)	/* approximateEntropy.c:181-278 */
{
	static const enum test test_num = TEST_APEN;
	struct state *state = enter(thread_state, test_num, 11, __func__);
	if (state == NULL)
		return;
	struct ApproximateEntropy_stats stat;
	double p_value = PV(thread_state, test_num)[0];
	memcpy(&stat.phi[0], &ST(thread_state, test_num)[0], sizeof(double));
	memcpy(&stat.phi[1], &ST(thread_state, test_num)[1], sizeof(double));
	stat.ApEn = state->resultstxtFlag ? FS(thread_state, test_num)[0] : UNSET_DOUBLE;
	stat.chi_squared = state->resultstxtFlag ? FS(thread_state, test_num)[1] : UNSET_DOUBLE;
	stat.success = keep(thread_state, test_num, p_value, __func__, "of test", "p_value");
	if (state->resultstxtFlag == true) {
		append_value(state->stats[test_num], &stat);
	}
	append_value(state->p_val[test_num], &p_value);
	leave(thread_state);
}

void
RandomExcursions_iterate(struct thread_state *thread_state)	/* randomExcursions.c:242-611 */
{
	static const enum test test_num = TEST_RND_EXCURSION;
	struct state *state = enter(thread_state, test_num, 151, __func__);
	if (state == NULL)
		return;
	struct RandomExcursions_stats stat;
	const double *pv = PV(thread_state, test_num);
	const int64_t *c = ST(thread_state, test_num);
	double p_value;
	memset(&stat, 0, sizeof(stat));
	stat.number_of_cycles = c[0];
	stat.test_possible = (stat.number_of_cycles < state->c.min_zero_crossings) ? false : true;	/* :358 */
	if (stat.test_possible == true) {
		for (int i = 0; i < NUMBER_OF_STATES_RND_EXCURSION; i++) {
			stat.counter[i] = c[49 + i];		/* the LAST cycle's visits: what :386 leaves behind */
			stat.chi2[i] = state->resultstxtFlag ? FS(thread_state, test_num)[i] : UNSET_DOUBLE;
		}
		for (int i = 0; i < NUMBER_OF_STATES_RND_EXCURSION; i++) {
			p_value = pv[i];
			stat.success[i] = keep(thread_state, test_num, p_value, __func__, "of test", "p_value");
			append_value(state->p_val[test_num], &p_value);
		}
		if (state->resultstxtFlag == true) {
			append_value(state->stats[test_num], &stat);
		}
	} else {
		state->count[test_num]++;		/* :573-600: counted once, success false, counters cleared */
		if (state->resultstxtFlag == true) {
			append_value(state->stats[test_num], &stat);
		}
		p_value = NON_P_VALUE;
		for (int i = 0; i < NUMBER_OF_STATES_RND_EXCURSION; i++) {
			append_value(state->p_val[test_num], &p_value);
		}
	}
	leave(thread_state);
}

void
RandomExcursionsVariant_iterate(struct thread_state *thread_state)	/* randomExcursionsVariant.c:191-422 */
{
	static const enum test test_num = TEST_RND_EXCURSION_VAR;
	struct state *state = enter(thread_state, test_num, 161, __func__);
	if (state == NULL)
		return;
	struct RandomExcursionsVariant_stats stat;
	const double *pv = PV(thread_state, test_num);
	const int64_t *c = ST(thread_state, test_num);
	double p_value;
	memset(&stat, 0, sizeof(stat));
	stat.number_of_cycles = c[0];
	stat.test_possible = (stat.number_of_cycles < state->c.min_zero_crossings) ? false : true;	/* :284 */
	if (stat.test_possible == true) {
		for (int i = 0; i < NUMBER_OF_STATES_RND_EXCURSION_VAR; i++) {
			stat.counter[i] = c[1 + i];		/* xi(x), :301-306 */
			p_value = pv[i];
			stat.success[i] = keep(thread_state, test_num, p_value, __func__, "of test", "p_value");
			append_value(state->p_val[test_num], &p_value);
		}
		if (state->resultstxtFlag == true) {
			append_value(state->stats[test_num], &stat);
		}
	} else {
		state->count[test_num]++;
		if (state->resultstxtFlag == true) {
			append_value(state->stats[test_num], &stat);
		}
		p_value = NON_P_VALUE;
		for (int i = 0; i < NUMBER_OF_STATES_RND_EXCURSION_VAR; i++) {
			append_value(state->p_val[test_num], &p_value);
		}
	}
	leave(thread_state);
}

void
Serial_iterate(struct thread_state *thread_state)		/* serial.c:180-300 */
{
	static const enum test test_num = TEST_SERIAL;
	struct state *state = enter(thread_state, test_num, 191, __func__);
	if (state == NULL)
		return;
	struct Serial_stats stat;
	double p_value1 = PV(thread_state, test_num)[0];
	double p_value2 = PV(thread_state, test_num)[1];
	const double *fs = state->resultstxtFlag ? FS(thread_state, test_num) : NULL;
	stat.psim0 = fs ? fs[0] : UNSET_DOUBLE;
	stat.psim1 = fs ? fs[1] : UNSET_DOUBLE;
	stat.psim2 = fs ? fs[2] : UNSET_DOUBLE;
	stat.del1 = fs ? fs[3] : UNSET_DOUBLE;
	stat.del2 = fs ? fs[4] : UNSET_DOUBLE;
	stat.success1 = keep(thread_state, test_num, p_value1, __func__, "of test", "p_value1");
	stat.success2 = keep(thread_state, test_num, p_value2, __func__, "of test", "p_value2");
	if (state->resultstxtFlag == true) {
		append_value(state->stats[test_num], &stat);
	}
	append_value(state->p_val[test_num], &p_value1);
	append_value(state->p_val[test_num], &p_value2);
	leave(thread_state);
}

void
LinearComplexity_iterate(struct thread_state *thread_state)	/* linearComplexity.c:217-441 */
{
	static const enum test test_num = TEST_LINEARCOMPLEXITY;
	struct state *state = enter(thread_state, test_num, 101, __func__);
	if (state == NULL)
		return;
	struct LinearComplexity_stats stat;
	double p_value = PV(thread_state, test_num)[0];
	for (int i = 0; i <= K_LINEARCOMPLEXITY; i++)
		stat.v[i] = ST(thread_state, test_num)[i];
	stat.chi2 = state->resultstxtFlag ? FS(thread_state, test_num)[0] : UNSET_DOUBLE;
	stat.success = keep(thread_state, test_num, p_value, __func__, "of test", "p_value");
	if (state->resultstxtFlag == true) {
		append_value(state->stats[test_num], &stat);
	}
	append_value(state->p_val[test_num], &p_value);
	leave(thread_state);
}

/* ------------------------------------------------------------------------------------------------------------ */
/* reading: the work the reference spreads over its -T testBits threads (utilities.c:1601-1625) */

struct read_slice {
	int fd;
	uint8_t *dst;
	size_t len;
	off_t off;
	size_t got;
	pthread_t tid;
};

static void *
read_slice_main(void *arg)
{
	struct read_slice *r = arg;
	size_t done = 0;
	while (done < r->len) {
		ssize_t k = pread(r->fd, r->dst + done, r->len - done, r->off + (off_t) done);
		if (k <= 0)
			break;
		done += (size_t) k;
	}
	r->got = done;
	return NULL;
}

/* `len` bytes at `off` of the data file into `dst` with up to -T threads; returns the bytes read (short only at EOF) */
static size_t
read_parallel(struct state *state, uint8_t *dst, size_t len, off_t off)
{
	long int nt = state->numberOfThreads;
	if (nt < 1)
		nt = 1;
	if (nt > GLUE_MAX_READERS)
		nt = GLUE_MAX_READERS;
	if (len < ((size_t) 8 << 20))
		nt = 1;
	struct read_slice sl[GLUE_MAX_READERS];
	const size_t per = (((len + (size_t) nt - 1) / (size_t) nt) + 4095) & ~(size_t) 4095;
	int used = 0;
	for (size_t o = 0; o < len && used < GLUE_MAX_READERS; o += per, used++) {
		sl[used].fd = fileno(state->streamFile);
		sl[used].dst = dst + o;
		sl[used].len = (len - o < per) ? len - o : per;
		sl[used].off = off + (off_t) o;
		sl[used].got = 0;
		if (used > 0 && pthread_create(&sl[used].tid, NULL, read_slice_main, &sl[used]) != 0) {
			errp(224, __func__, "error on pthread_create()");
		}
	}
	if (used > 0)
		read_slice_main(&sl[0]);
	size_t total = 0;
	bool short_seen = false;
	for (int i = 0; i < used; i++) {
		if (i > 0 && pthread_join(sl[i].tid, NULL) != 0) {
			errp(224, __func__, "error on pthread_join()");
		}
		if (!short_seen)
			total += sl[i].got;
		if (sl[i].got < sl[i].len)
			short_seen = true;
	}
	return total;
}

/*
 * Standard input cannot seek (utilities.c:1500): the streams are consumed in sequence.  Raw bytes are the packed
 * form; ASCII digits are packed here, MSB first (parseBitsASCIIInput's "%1d" skips white space; any other character
 * ends the conversion there, which this reports as the reference's "Insufficient data").
 * Returns the number of complete streams.
 */
static long int
read_sequential(struct state *state, uint8_t *dst, long int streams, long int first_iteration)
{
	const long int n = state->tp.n, nb = n / BITS_N_BYTE;
	if (state->dataFormat == FORMAT_RAW_BINARY) {
		for (long int k = 0; k < streams; k++) {
			size_t got = fread(dst + k * nb, 1, (size_t) nb, state->streamFile);
			if (ferror(state->streamFile)) {
				errp(226, __func__, "read error while reading file: %s", state->randomDataPath);
			} else if (got != (size_t) nb) {		/* utilities.c:1789-1791: EOF inside a raw stream is fatal */
				err(226, __func__, "encounted EOF (end of file) while reading file: %s: %ld bits were read before EOF",
				    state->randomDataPath, (long int) got * BITS_N_BYTE);
			}
		}
		return streams;
	}
	for (long int k = 0; k < streams; k++) {
		uint8_t *o = dst + k * nb;
		memset(o, 0, (size_t) nb);
		for (long int bit = 0; bit < n;) {
			int c = fgetc(state->streamFile);
			if (c == '0' || c == '1') {
				if (c == '1')
					o[bit >> 3] |= (uint8_t) (0x80u >> (bit & 7));
				bit++;
			} else if (c == ' ' || c == '\n' || c == '\r' || c == '\t' || c == '\f' || c == '\v') {
				continue;
			} else {
				warn(__func__, "Insufficient data in file %s: %ld bits were read", state->randomDataPath, bit);
				(void) first_iteration;
				return k;
			}
		}
	}
	return streams;
}

/* ------------------------------------------------------------------------------------------------------------ */

static double
now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

static void
engine_fatal(int code, const char *name, const char *what, int rc)
{
	err(code, name, "%s: %s: %s", what, stsgpu_strerror(rc), stsgpu_last_error(G.engine));
}

static void
engine_start(struct state *state, long int batch)
{
	stsgpu_params p;
	int rc;

	stsgpu_default_params(&p);
	p.n = state->tp.n;
	p.max_streams = batch;
	p.test_mask = 0;
	for (int i = 1; i <= NUMOFTESTS; i++)
		if (state->testVector[i] == true)
			p.test_mask |= 1u << i;
	if (state->legacy_output == true)
		p.test_mask |= 1u << TEST_FREQUENCY;	/* freq.txt's "BITSREAD = .. 0s = .. 1s = .." line needs S_n */
	p.block_frequency_M = state->tp.blockFrequencyBlockLength;
	p.non_overlapping_m = (int32_t) state->tp.nonOverlappingTemplateLength;
	p.overlapping_m = (int32_t) state->tp.overlappingTemplateLength;
	p.apen_m = (int32_t) state->tp.approximateEntropyBlockLength;
	p.serial_m = (int32_t) state->tp.serialBlockLength;
	p.linear_complexity_M = state->tp.linearComplexitySequenceLength;
	p.pipeline_depth = GLUE_DEPTH;
	p.flags = (state->resultstxtFlag == true) ? STSGPU_FLAG_STATS : 0;
	if (getenv("STS_GPU_DEVICE") != NULL)
		p.device = atoi(getenv("STS_GPU_DEVICE"));
	rc = stsgpu_create(&p, &G.engine);
	if (rc != STSGPU_OK) {		/* no CPU fallback: fatal, in the exit code range of utilities.c */
		err(228, __func__, "cannot start the GPU engine: %s: %s", stsgpu_strerror(rc), stsgpu_last_error(NULL));
	}
	stsgpu_get_layout(G.engine, &G.lay);
	/* the X_init functions of the reference have decided which tests run; the engine applies the same rules */
	for (int i = 1; i <= NUMOFTESTS; i++) {
		const bool on = ((G.lay.enabled_mask >> i) & 1u) != 0;
		if (state->testVector[i] == true && !on) {
			err(228, __func__, "test %s[%d] is enabled in the driver but the GPU engine disabled it for these parameters",
			    state->testNames[i], i);
		}
		if (state->testVector[i] == true && state->resultstxtFlag == true &&
		    (state->stats[i] == NULL || state->stats[i]->elm_size != stats_size[i])) {
			err(228, __func__, "stats struct of test %s[%d]: the driver allocates %ld bytes per iteration, the glue fills %ld",
			    state->testNames[i], i, state->stats[i] ? (long int) state->stats[i]->elm_size : -1L, (long int) stats_size[i]);
		}
	}
	if (state->testVector[TEST_NON_OVERLAPPING] == true && state->p_val[TEST_NON_OVERLAPPING]->elm_size != sizeof(struct nonover_stats)) {
		err(228, __func__, "p_val of the non-overlapping test does not hold struct nonover_stats");
	}
}

/*
 * invokeTestSuite - utilities.c:1434-1476, with the batch loop in place of handleFileBasedBitStreams (:1480-1578)
 */
void
invokeTestSuite(struct state *state)
{
	int io_ret;		// I/O return status

	if (state == NULL) {
		err(228, __func__, "state arg is NULL");
	}

	/*
	 * Announce test if in legacy output mode (utilities.c:1448-1470)
	 */
	if (state->legacy_output == true) {
		io_ret = fprintf(state->freqFile,
				 "________________________________________________________________________________\n\n");
		if (io_ret <= 0) {
			errp(228, __func__, "error in writing to %s", state->freqFilePath);
		}
		io_ret = fprintf(state->freqFile, "\t\tFILE = %s\t\tALPHA = %6.4f\n", state->randomDataPath, state->tp.alpha);
		if (io_ret <= 0) {
			errp(228, __func__, "error in writing to %s", state->freqFilePath);
		}
		io_ret = fprintf(state->freqFile,
				 "________________________________________________________________________________\n\n");
		if (io_ret <= 0) {
			errp(228, __func__, "error in writing to %s", state->freqFilePath);
		}
		io_ret = fflush(state->freqFile);
		if (io_ret != 0) {
			errp(228, __func__, "error flushing to %s", state->freqFilePath);
		}
	}
	if (state->streamFile == NULL) {
		err(225, __func__, "streamFile arg is NULL");
	}

	const long int n = state->tp.n, nb = n / BITS_N_BYTE;
	const long int total = state->tp.numOfBitStreams;
	long int B = getenv("STS_GPU_BATCH") != NULL ? atol(getenv("STS_GPU_BATCH")) : GLUE_BATCH;
	if (B < 1)
		B = 1;
	if (B > total)
		B = total;
	engine_start(state, B);

	/*
	 * the -j seek rule of utilities.c:1500-1527
	 */
	const bool sequential = (state->stdinData == true);
	const bool text = (state->dataFormat == FORMAT_ASCII_01) && !sequential;
	long int file_size = 0;
	if (sequential) {
		state->base_seek = 0;
	} else if (state->dataFormat == FORMAT_ASCII_01) {
		state->base_seek = state->jobnum * state->tp.n * state->tp.numOfBitStreams;
	} else {
		state->base_seek = ((state->jobnum * state->tp.n * state->tp.numOfBitStreams) + BITS_N_BYTE - 1) / BITS_N_BYTE;
	}
	if (!sequential) {
		if (fseek(state->streamFile, 0, SEEK_END) != 0 || (file_size = ftell(state->streamFile)) < 0) {
			errp(226, __func__, "could not find the size of file: %s", state->randomDataPath);
		}
	}

	dbg(DBG_LOW, "Start of iterate phase");

	const size_t buf_bytes = text ? (size_t) (B + 1) * (size_t) n : (size_t) B * (size_t) nb;
	const int depth = text ? 2 : GLUE_DEPTH;	/* text is 8 bytes per tested byte: PCIe bound at any depth */
	uint8_t *buf[GLUE_DEPTH] = { NULL, NULL, NULL };
	for (int i = 0; i < depth; i++) {
		buf[i] = stsgpu_host_alloc(buf_bytes);
		if (buf[i] == NULL) {
			errp(224, __func__, "cannot allocate a pinned read buffer of %ld bytes", (long int) buf_bytes);
		}
	}
	struct thread_state thread_state;
	thread_state.thread_id = 0;
	thread_state.global_state = state;
	thread_state.iteration_being_done = 0;
	thread_state.mutex = NULL;		/* one record-keeping thread: every tail tolerates a NULL mutex (e.g. frequency.c:218) */

	long int submitted = 0, done = 0, waited = 0;
	long int want_of[GLUE_DEPTH] = { 0, 0, 0 };
	bool at_eof[GLUE_DEPTH] = { false, false, false };
	bool eof = false, short_seen = false;
	int w = 0, rc;
	char tbuf[BUFSIZ + 1];	// time string buffer
	double t_read = 0.0, t_wait = 0.0, t_keep = 0.0;
	const double t_begin = now_s();

	for (;;) {
		while (stsgpu_pending(G.engine) < depth && !eof && submitted < total) {
			const long int want = MIN(B, total - submitted);
			long int have = want;
			uint8_t *dst = buf[w % depth];
			const double t_r0 = now_s();
			if (text) {
				/* iteration k = the first n digits at or after character base_seek + k n (the fseek of :1682) */
				const long int off = state->base_seek + submitted * n;
				long int len = (want + 1) * n;
				if (off >= file_size) {
					warn(__func__, "Insufficient data in file %s: %ld bits were read", state->randomDataPath, 0L);
					eof = true;
					break;
				}
				at_eof[w % depth] = (off + len >= file_size);
				if (off + len > file_size)
					len = file_size - off;
				if ((long int) read_parallel(state, dst, (size_t) len, (off_t) off) != len) {
					errp(226, __func__, "read error while reading file: %s", state->randomDataPath);
				}
				rc = stsgpu_submit_ascii(G.engine, (const char *) dst, len, want, submitted);
			} else if (!sequential) {
				const off_t off = (off_t) state->base_seek + (off_t) submitted * nb;
				const size_t got = read_parallel(state, dst, (size_t) want * (size_t) nb, off);
				if (got != (size_t) want * (size_t) nb) {	/* utilities.c:1789-1791 */
					err(226, __func__, "encounted EOF (end of file) while reading file: %s: %ld bits were read before EOF",
					    state->randomDataPath, (long int) (got % (size_t) nb) * BITS_N_BYTE);
				}
				rc = stsgpu_submit(G.engine, dst, want, submitted);
			} else {
				have = read_sequential(state, dst, want, submitted);
				if (have < want)
					eof = true;
				if (have <= 0)
					break;
				rc = stsgpu_submit(G.engine, dst, have, submitted);
			}
			if (rc != STSGPU_OK) {
				engine_fatal(229, __func__, "cannot submit a batch to the GPU engine", rc);
			}
			want_of[w % depth] = have;
			submitted += have;
			w++;
			t_read += now_s() - t_r0;
		}
		if (stsgpu_pending(G.engine) == 0)
			break;
		const double t_w0 = now_s();
		rc = stsgpu_wait(G.engine);
		if (rc != STSGPU_OK) {
			engine_fatal(229, __func__, "the GPU engine failed on a batch", rc);
		}
		const double t_k0 = now_s();
		t_wait += t_k0 - t_w0;
		int64_t bn = 0, bf = 0;
		stsgpu_results(G.engine, &bn, &bf, &G.pv, &G.st, &G.w);
		stsgpu_results_stats(G.engine, &G.fs);
		if (text) {
			const int slot = (int) (waited % depth);
			int64_t complete = 0, bad = -1;
			stsgpu_ascii_status(G.engine, &complete, &bad);
			if (bad >= 0 && !short_seen) {
				/* fscanf("%1d") stops converting at such a character: the stream ends there (utilities.c:1692-1697) */
				warn(__func__, "Insufficient data in file %s: found the character '%c' at offset %ld", state->randomDataPath,
				     buf[slot][bad], (long int) (state->base_seek + bf * n + bad));
			}
			if (complete < want_of[slot] && !short_seen) {
				short_seen = true;
				if (!at_eof[slot] && bad < 0) {
					err(226, __func__, "more than %ld white-space characters inside one batch of %s", n, state->randomDataPath);
				}
				warn(__func__, "Insufficient data in file %s: iteration %ld is incomplete", state->randomDataPath,
				     (long int) (bf + complete));
				eof = true;	/* every later iteration starts further into the text and is incomplete too */
			}
		}
		waited++;
		G.first = (long int) bf;
		G.count = (long int) bn;
		for (long int k = 0; k < (long int) bn; k++) {
			thread_state.iteration_being_done = done + k;
			if (state->legacy_output == true) {		/* utilities.c:1714 / :1805 */
				const long int S_n = (long int) ST(&thread_state, TEST_FREQUENCY)[0];
				const long int num_1s = (n + S_n) / 2;
				io_ret = fprintf(state->freqFile, "\t\tBITSREAD = %ld 0s = %ld 1s = %ld\n", n, n - num_1s, num_1s);
				if (io_ret <= 0) {
					errp(226, __func__, "error in writing to %s", state->freqFilePath);
				}
				io_ret = fflush(state->freqFile);
				if (io_ret != 0) {
					errp(226, __func__, "error flushing to %s", state->freqFilePath);
				}
			}
			iterate(&thread_state);		/* the reference's own driver.c:395-425, over its own table */
			if (state->reportCycle > 0 && (((thread_state.iteration_being_done % state->reportCycle) == 0) ||
						       (thread_state.iteration_being_done == state->tp.numOfBitStreams))) {
				getTimestamp(tbuf, BUFSIZ);
				msg("Completed iteration %ld of %ld at %s", thread_state.iteration_being_done + 1,
				    state->tp.numOfBitStreams, tbuf);
			}
		}
		done += (long int) bn;
		t_keep += now_s() - t_k0;
	}
	dbg(DBG_LOW, "iterate phase on the GPU: %ld streams of %ld bits in %.3f s = %.2f Gbit/s (read + submit %.3f s, waiting for the "
	    "GPU %.3f s, iterate() record keeping %.3f s)", done, n, now_s() - t_begin, (double) done * (double) n / (now_s() - t_begin) / 1e9,
	    t_read, t_wait, t_keep);
	state->iterationsMissing = total - done;
	if (done < total) {
		/*
		 * ASCII input ended early.  The reference warns and still iterates on whatever its epsilon buffer holds
		 * (utilities.c:1692-1697); here the incomplete iterations are not tested and the run is as long as the data.
		 */
		state->tp.numOfBitStreams = done;
	}
	G.pv = NULL; G.st = NULL; G.w = NULL; G.fs = NULL;
	G.first = G.count = 0;
	for (int i = 0; i < depth; i++)
		stsgpu_host_free(buf[i]);
	stsgpu_destroy(G.engine);
	G.engine = NULL;

	dbg(DBG_LOW, "End of iterate phase\n");

	/*
	 * Close the input file
	 */
	errno = 0;	// paranoia
	io_ret = fclose(state->streamFile);
	if (io_ret != 0) {
		errp(224, __func__, "error closing: %s", state->randomDataPath);
	}
	state->streamFile = NULL;

	return;
}
