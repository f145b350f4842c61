/*
 * sts_host.h - host side of the drop-in, in C like the reference's own host code.
 *
 * It mirrors, for the hot path only, the reference's plugin/operator interface:
 *   struct sts_driver            <-> struct driver, src/utils/driver.c:55-61 (init/iterate/print/metrics/destroy)
 *   struct sts_thread_state      <-> struct thread_state, src/utils/defs.h:489-494
 *   struct sts_state             <-> the members of struct state (defs.h:366-487) the path touches:
 *                                    tp, testVector, jobnum, runMode, dataFormat, workDir, counters, p_val
 *   sts_init()                   <-> init(),  driver.c:205   (runs every X_init: here stsgpu_create)
 *   sts_invokeTestSuite()        <-> invokeTestSuite()/handleFileBasedBitStreams()/testBits(),
 *                                    utilities.c:1433-1639: batched "fill, submit, wait, iterate in order"
 *   sts_iterate()                <-> iterate(), driver.c:395-425: walks testDriver[1..15].iterate
 *   X_iterate (15 functions)     <-> the record-keeping tail R1 of every reference X_iterate
 *                                    (e.g. frequency.c:218-260): count/valid/success/failure and append to p_val
 *   sts_write_p_val_to_file()    <-> write_p_val_to_file(), utilities.c:1938-2042 (same bytes, same name)
 *
 * Everything above the C ABI of include/stsgpu.h; nothing here computes a statistic.
 */
#ifndef STS_HOST_H
#define STS_HOST_H

#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <pthread.h>
#include "../../include/stsgpu.h"

#define STS_NUMOFTESTS 15
#define STS_PIPELINE_DEPTH 3		/* batches in flight in sts_invokeTestSuite (stsgpu_params.pipeline_depth) */
#define STS_NON_P_VALUE (-99999999.0)

enum sts_run_mode { STS_MODE_ITERATE_AND_ASSESS = 0, STS_MODE_ITERATE_ONLY = 1, STS_MODE_ASSESS_ONLY = 2 };	/* -m b/i/a */
enum sts_format { STS_FORMAT_RAW_BINARY = 'r', STS_FORMAT_ASCII_01 = 'a' };					/* -F r/a  */

struct sts_tp {			/* struct TP, defs.h:277-289 */
	long blockFrequencyBlockLength;		/* -P 1 */
	long nonOverlappingTemplateLength;	/* -P 2 */
	long overlappingTemplateLength;		/* -P 3 */
	long approximateEntropyBlockLength;	/* -P 4 */
	long serialBlockLength;			/* -P 5 */
	long linearComplexitySequenceLength;	/* -P 6 */
	long numOfBitStreams;			/* -P 7 / -i */
	long uniformityBins;			/* -P 8 (assessment only; carried, not used) */
	long n;					/* -P 9 / -S */
	double uniformityLevel;			/* -P 10 */
	double alpha;				/* -P 11 */
};

struct sts_pvals {		/* the double payload of state->p_val[test] (dyn_alloc.h), grown by append */
	double *v;
	long count, cap;
};

struct sts_state {
	struct sts_tp tp;
	bool testVector[STS_NUMOFTESTS + 1];
	bool testVectorFlag;
	long jobnum;				/* -j */
	enum sts_run_mode runMode;		/* -m */
	enum sts_format dataFormat;		/* -F */
	long numberOfThreads;			/* -T: threads that read the data file (0 = not given: min(cores, 16)) */
	long batchStreams;			/* engine batch capacity (new: -B, default 1024) */
	int device;				/* CUDA device ordinal (new: -D, default 0) */
	long generator;				/* -g: 0 = data file (the only value sts 3 accepts); 1 = tools/generators.c generator 1 (LCG) made on the GPU */
	long reportCycle;			/* -I */
	int debuglevel;				/* -v */
	bool legacy_output;			/* -O: 10 uniformity bins, finalAnalysisReport.txt (parse_args.c:585, :792-803) */
	bool resultstxtFlag;			/* -s: results.txt and data*.txt per test (parse_args.c:606) */
	bool uniformityBinsFlag;		/* -P 8 given (parse_args.c:993) */
	char *workDir;				/* -w */
	char *randomDataPath;			/* data file */
	char *pvalues_dir;			/* -d: directory of sts.*.*.<n>.pvalues files (-m a) */
	long binsIterations;			/* the -i value sqrt(iterations) bins are derived from (parse_args.c:791-793) */
	long successful_tests;			/* p-value columns that passed both analyses (e.g. frequency.c:765) */
	FILE *streamFile;
	/* counters of defs.h:437-441 */
	long count[STS_NUMOFTESTS + 1], valid[STS_NUMOFTESTS + 1], success[STS_NUMOFTESTS + 1],
	     failure[STS_NUMOFTESTS + 1], valid_p_val[STS_NUMOFTESTS + 1];
	int partitionCount[STS_NUMOFTESTS + 1];
	struct sts_pvals p_val[STS_NUMOFTESTS + 1];
	/* engine */
	stsgpu_ctx *engine;
	stsgpu_layout layout;
	const double *batch_pvals;		/* records of the batch being iterated */
	long batch_first, batch_count;
};

struct sts_thread_state {	/* defs.h:489-494 */
	long thread_id;
	struct sts_state *global_state;
	long iteration_being_done;
	pthread_mutex_t *mutex;
};

struct sts_driver {		/* driver.c:55-61 */
	void (*init)(struct sts_state *state);
	void (*iterate)(struct sts_thread_state *thread_state);
	void (*print)(struct sts_state *state);
	void (*metrics)(struct sts_state *state);
	void (*destroy)(struct sts_state *state);
};
extern const struct sts_driver sts_testDriver[STS_NUMOFTESTS + 1];
extern const char *const sts_testNames[STS_NUMOFTESTS + 1];
/* metrics() for the uniformity / proportion table (driver.c:475, X_metrics): tallied on the GPU */
void sts_metrics(struct sts_state *s);

void sts_defaultstate(struct sts_state *state);				/* parse_args.c:55-290 defaults */
void sts_change_params(struct sts_state *state, long num, long value, double d_value);	/* parse_args.c:960 */
void sts_init(struct sts_state *state);					/* driver.c:205 */
void sts_invokeTestSuite(struct sts_state *state);			/* utilities.c:1433 */
void sts_iterate(struct sts_thread_state *thread_state);		/* driver.c:395 */
void sts_print(struct sts_state *state);				/* driver.c:434 */
void sts_write_p_val_to_file(struct sts_state *state);			/* utilities.c:1938 */
void sts_find_p_val_files(struct sts_state *state);			/* parse_args.c:837-921: count the iterations of -d dir */
void sts_read_from_p_val_file(struct sts_state *state);			/* utilities.c:2051: here streamed into the GPU tallies */
void sts_destroy(struct sts_state *state);				/* driver.c:1153 */

/* fatal error in the reference's style: message + exit(code) (debug.c:267) */
void sts_err(int exitcode, const char *name, const char *fmt, ...);

#endif
