/*
 * stsgpu.h - C ABI of the B200-native engine for the per-bitstream test kernels of sts
 *            (the improved NIST SP800-22 suite, arcetri/sts v3.2.7).
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no C++/torch types.
 * It replaces, for the hot path only,
 *
 *   - the dispatch/batching of src/utils/utilities.c:1479-1884
 *     (handleFileBasedBitStreams / testBits / parseBitsBinaryInput / copyBitsToEpsilon), and
 *   - the compute part of the 15 X_iterate() bodies in src/tests/ (*.c) reached through
 *     iterate() in src/utils/driver.c:395-425,
 *
 * while the reference's host driver, plugin table (struct driver, driver.c:55-61), record
 * keeping tails, .pvalues files and result.txt assessment stay as they are.  INTEGRATION.md shows
 * the glue a maintainer adds on the reference side.
 *
 * Data contract
 * -------------
 * Input: `nstreams` bitstreams of `n` bits each, as the RAW FILE BYTES sts reads with -F r:
 * stream s occupies bytes [s*n/8, (s+1)*n/8); bit k of a stream is
 * (byte[k/8] >> (7 - k%8)) & 1  (utilities.c:1862-1872, MSB first).  n % 8 == 0 is required,
 * exactly as parse_args.c:939 requires it.
 *
 * Output, per stream (one "record"), in three flat arrays whose per-test offsets are given by
 * stsgpu_get_layout():
 *   pvals  (double)   the p-values in exactly the order X_iterate appends them to
 *                     state->p_val[test] (and therefore the order of the .pvalues file,
 *                     utilities.c:1967-2024); NON_P_VALUE (-99999999.0, defs.h:314) where the
 *                     reference records it.
 *   istats (int64_t)  the integer statistics each test derives from the bits (see STSGPU_ST_*).
 *   wcount (uint32_t) W[template][block] of the non-overlapping template test,
 *                     ntemplates x 8, template-major.
 *
 * Threading: one context per GPU, single caller per context.  No CPU fallback exists: every entry
 * point fails with STSGPU_E_CUDA / STSGPU_E_NO_DEVICE when no usable B200 is present.
 */
#ifndef STSGPU_H
#define STSGPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STSGPU_ABI_VERSION 2

/* test numbers: identical to enum test, defs.h:217-236 */
enum {
	STSGPU_TEST_FREQUENCY = 1,
	STSGPU_TEST_BLOCK_FREQUENCY = 2,
	STSGPU_TEST_CUSUM = 3,
	STSGPU_TEST_RUNS = 4,
	STSGPU_TEST_LONGEST_RUN = 5,
	STSGPU_TEST_RANK = 6,
	STSGPU_TEST_DFT = 7,
	STSGPU_TEST_NON_OVERLAPPING = 8,
	STSGPU_TEST_OVERLAPPING = 9,
	STSGPU_TEST_UNIVERSAL = 10,
	STSGPU_TEST_APEN = 11,
	STSGPU_TEST_RND_EXCURSION = 12,
	STSGPU_TEST_RND_EXCURSION_VAR = 13,
	STSGPU_TEST_SERIAL = 14,
	STSGPU_TEST_LINEARCOMPLEXITY = 15,
	STSGPU_NUMOFTESTS = 15
};
#define STSGPU_ALL_TESTS 0xFFFEu	/* bits 1..15 */
#define STSGPU_NON_P_VALUE (-99999999.0)	/* defs.h:314 */

/* error codes (negative); 0 is success */
enum {
	STSGPU_OK = 0,
	STSGPU_E_INVAL = -1,		/* bad argument */
	STSGPU_E_NO_DEVICE = -2,	/* no CUDA device / not sm_100 */
	STSGPU_E_CUDA = -3,		/* a CUDA runtime call or kernel failed */
	STSGPU_E_NOMEM = -4,		/* host or device allocation failed */
	STSGPU_E_UNSUPPORTED = -5,	/* parameter combination outside the engine's limits */
	STSGPU_E_STATE = -6,		/* call sequence error (e.g. wait without submit) */
	STSGPU_E_NCCL = -7,		/* NCCL missing or failed */
	STSGPU_E_RANGE = -8		/* iteration / index out of range */
};

/*
 * Test parameters: the subset of struct TP (defs.h:277-289) the kernels depend on, plus the
 * batch geometry.  Field meanings and defaults are the reference's (-P 1..6,9; defs.h:110-125).
 */
typedef struct stsgpu_params {
	int32_t abi_version;		/* must be STSGPU_ABI_VERSION */
	int32_t device;			/* CUDA device ordinal */
	int64_t n;			/* tp.n: bits per stream (default 1048576) */
	int64_t max_streams;		/* capacity of one submit (streams per batch) */
	uint32_t test_mask;		/* bit t set <=> state->testVector[t] (1..15) */
	int32_t pipeline_depth;		/* batches that may be in flight at once: 0 or 1 = one (default), max 4 */
	int64_t block_frequency_M;	/* tp.blockFrequencyBlockLength      (16384) */
	int32_t non_overlapping_m;	/* tp.nonOverlappingTemplateLength   (9) */
	int32_t overlapping_m;		/* tp.overlappingTemplateLength      (9) */
	int32_t apen_m;			/* tp.approximateEntropyBlockLength  (10) */
	int32_t serial_m;		/* tp.serialBlockLength              (16) */
	int64_t linear_complexity_M;	/* tp.linearComplexitySequenceLength (500) */
	uint32_t flags;			/* STSGPU_FLAG_*; 0 by default */
	uint32_t reserved;		/* must be 0 */
} stsgpu_params;

/*
 * state->resultstxtFlag (-s, defs.h:396): also produce what the private stats structs of the 15 tests hold beyond
 * the integers and the p-values: the floating point intermediates (`fstats`, below) and the visit
 * counts of the LAST excursion cycle that RandomExcursions prints (randomExcursions.c:386, :860-866).
 */
#define STSGPU_FLAG_STATS 1u

/* per-stream record layout; offsets are in elements of the respective array */
typedef struct stsgpu_layout {
	uint32_t enabled_mask;		/* tests still enabled after the X_init style checks */
	int32_t npv;			/* p-values per stream */
	int32_t nst;			/* int64 statistics per stream */
	int32_t nw;			/* uint32 W counts per stream (= ntemplates * 8) */
	int32_t ntemplates;		/* non-overlapping templates (148 for m = 9) */
	int32_t universal_L;		/* L chosen by the rule of universal.c:145-178 */
	int32_t pv_off[STSGPU_NUMOFTESTS + 1], pv_cnt[STSGPU_NUMOFTESTS + 1];
	int32_t st_off[STSGPU_NUMOFTESTS + 1], st_cnt[STSGPU_NUMOFTESTS + 1];
	int32_t nfs;			/* double statistics per stream (produced with STSGPU_FLAG_STATS only) */
	int32_t fs_off[STSGPU_NUMOFTESTS + 1], fs_cnt[STSGPU_NUMOFTESTS + 1];
} stsgpu_layout;

/*
 * istats slots, relative to st_off[test]:
 *  FREQUENCY        [0] S_n                                          (frequency.c:191-201)
 *  BLOCK_FREQUENCY  [0] N, [1..N] ones in block i                    (blockFrequency.c:217-237)
 *  CUSUM            [0] S, [1] S_max, [2] S_min, [3] z_fwd, [4] z_bwd (cusum.c:221-236)
 *  RUNS             [0] ones, [1] V_n, [2] test_possible             (runs.c:213-240)
 *  LONGEST_RUN      [0..6] class counts                              (longestRunOfOnes.c:383-409)
 *  RANK             [0] F_32, [1] F_31, [2] F_remaining              (rank.c:303-327)
 *  DFT              [0] N_1                                          (discreteFourierTransform.c:407-411)
 *  NON_OVERLAPPING  none (see wcount)
 *  OVERLAPPING      [0..5] v[k]                                      (overlappingTemplateMatchings.c:262-296)
 *  UNIVERSAL        [0] bit pattern of the double `sum`              (universal.c:355-375)
 *  APEN             [0] bits of phi(m), [1] bits of phi(m+1), [2] sum C_m[i]^2,
 *                   [3] sum C_{m+1}[i]^2, [4] sum i*C_m[i], [5] sum i*C_{m+1}[i]
 *                                                                    (approximateEntropy.c:369-407)
 *  RND_EXCURSION    [0] J, [1 + 8*k + i] v[k][i], k = 0..5, i = 0..7 (randomExcursions.c:324-464);
 *                   [49 + i] visits to state i during the last cycle (stat.counter[], :386 and :441; zero
 *                   when the test was not possible, :581) - only with STSGPU_FLAG_STATS, otherwise 0
 *  RND_EXCURSION_VAR [0] J, [1 + i] xi(state i), i = 0..17           (randomExcursionsVariant.c:262-306)
 *  SERIAL           [0..2] sum v_i^2 for block sizes m, m-1, m-2     (serial.c:390-419)
 *  LINEARCOMPLEXITY [0..6] v[class]                                  (linearComplexity.c:367-377)
 *
 * fstats slots, relative to fs_off[test] (the doubles of each test's private stats struct, evaluated by the tail
 * kernels with the reference's expressions on the way to the p-value):
 *  BLOCK_FREQUENCY  [0] chi_squared                                  (blockFrequency.c:241)
 *  RUNS             [0] pi, [1] erfc_arg; both 0.0 (UNSET_DOUBLE) when the test is not possible (runs.c:219, :244, :299-301)
 *  LONGEST_RUN      [0] chi2                                         (longestRunOfOnes.c:412-416)
 *  RANK             [0] chi_squared                                  (rank.c:333-341)
 *  DFT              [0] N_0, [1] d                                   (discreteFourierTransform.c:401, :416)
 *  NON_OVERLAPPING  [t] chi2 of template t                           (nonOverlappingTemplateMatchings.c:561-565)
 *  OVERLAPPING      [0] chi2                                         (overlappingTemplateMatchings.c:301-305)
 *  UNIVERSAL        [0] sigma, [1] f_n                               (universal.c:380-386)
 *  APEN             [0] ApEn, [1] chi_squared                        (approximateEntropy.c:222-223)
 *  RND_EXCURSION    [i] chi2 of state i; 0.0 when the test is not possible (randomExcursions.c:492-497)
 *  SERIAL           [0] psim0, [1] psim1, [2] psim2, [3] del1, [4] del2 (serial.c:210-218)
 *  LINEARCOMPLEXITY [0] chi2                                         (linearComplexity.c:383-386)
 *  FREQUENCY, CUSUM, RND_EXCURSION_VAR: none
 */

typedef struct stsgpu_ctx stsgpu_ctx;

/* ---- life cycle --------------------------------------------------------------------------- */

/* fill *p with the reference's defaults (defs.h:110-125) for n bits, all 15 tests enabled */
void stsgpu_default_params(stsgpu_params *p);

/*
 * Create an engine.  Plays the role of the parts of the 15 X_init() that validate parameters
 * and size per-thread scratch (e.g. blockFrequency.c:102-129, universal.c:145-178,
 * nonOverlappingTemplateMatchings.c:211-241): tests whose preconditions fail are disabled, as in
 * the reference, and reported through layout.enabled_mask.
 */
int stsgpu_create(const stsgpu_params *p, stsgpu_ctx **out);
void stsgpu_destroy(stsgpu_ctx *ctx);
int stsgpu_get_layout(const stsgpu_ctx *ctx, stsgpu_layout *out);
/*
 * The same layout for a parameter set WITHOUT creating an engine and without touching a device: what the 15
 * X_init() would leave enabled (enabled_mask), and the per-test p-value / statistic counts.  The host uses it to
 * size state->p_val[] before the first batch; `device` and `pipeline_depth` are ignored.  STSGPU_E_INVAL for
 * arguments stsgpu_create would refuse, or when no test is left enabled.
 */
int stsgpu_query_layout(const stsgpu_params *p, stsgpu_layout *out);
/*
 * 1 if the engine's FFT can take streams of n bits.  Like the reference's libfftw3 plan
 * (discreteFourierTransform.c:194) it takes any even n: the four-step kernels when n/2 splits into two 5-smooth
 * factors of at most 4096 (every power of two up to 2^25, 10^6, 10^7 ...), a Bluestein FFT otherwise - up to
 * n = 2^25.  Only longer streams with a non-5-smooth half are refused: stsgpu_create then fails with
 * STSGPU_E_UNSUPPORTED unless bit 7 of test_mask is clear - never a CPU fallback.
 */
int stsgpu_dft_supported(int64_t n);

/* ---- the iterate() replacement ------------------------------------------------------------ */

/*
 * Submit `nstreams` (<= max_streams) streams of raw file bytes (nstreams * n/8 bytes at `raw`,
 * borrowed until stsgpu_wait returns).  `first_iteration` is the reference's iteration index of
 * the first stream (thread_state->iteration_being_done, defs.h:489-494).  Asynchronous: H2D copy,
 * all enabled test kernels and the D2H copy of the records are queued on the engine's streams.
 * Replaces parseBitsBinaryInput + copyBitsToEpsilon + iterate() (utilities.c:1601-1625).
 */
int stsgpu_submit(stsgpu_ctx *ctx, const uint8_t *raw, int64_t nstreams, int64_t first_iteration);

/*
 * Same, for input that is already resident in the device input buffer of the batch slot this
 * submit uses (filled by stsgpu_generate_lcg or stsgpu_upload since the previous submit, or left
 * there by the submit `pipeline_depth` batches ago): no H2D inside.
 */
int stsgpu_submit_resident(stsgpu_ctx *ctx, int64_t nstreams, int64_t first_iteration);

/*
 * "-F a" input (replaces parseBitsASCIIInput, utilities.c:1651-1727): `text` holds `text_len` characters of the
 * data file starting at the character offset of the batch's first stream (base_seek + first_iteration * n,
 * utilities.c:1511 and :1682-1683), borrowed until stsgpu_wait returns.  Stream k of the batch is the first n
 * characters '0' / '1' at or after text[k * n], white space skipped - the reference seeks to that offset for
 * every iteration, so in a file with line breaks consecutive streams overlap exactly as they do there.  Pass up
 * to (nstreams + 1) * n characters: the extra stream's worth is the room for white space.  The digits are packed
 * on the device; otherwise as stsgpu_submit.  After stsgpu_wait, stsgpu_results reports only the leading
 * streams that were complete, and stsgpu_ascii_status tells why the rest are missing.
 */
int stsgpu_submit_ascii(stsgpu_ctx *ctx, const char *text, int64_t text_len, int64_t nstreams, int64_t first_iteration);

/*
 * For the batch stsgpu_wait last completed, if it was an ASCII batch: *complete_streams = number of leading
 * streams that found their n digits inside `text` (the reference warns "Insufficient data" for the first
 * incomplete one, utilities.c:1693-1697); *bad_offset = offset in `text` of the first character that is neither
 * 0, 1 nor white space among those the reference would have scanned, or -1.
 */
int stsgpu_ascii_status(stsgpu_ctx *ctx, int64_t *complete_streams, int64_t *bad_offset);

/* copy raw bytes into the device input buffer without running the tests */
int stsgpu_upload(stsgpu_ctx *ctx, const uint8_t *raw, int64_t nstreams);

/*
 * Block until the OLDEST submitted batch is done and its records are in host memory.  Up to
 * `pipeline_depth` batches may be submitted before the first wait; they complete in submit order,
 * and the copies of one batch overlap the kernels of the others.  (The reference overlaps the same
 * way with -T threads: one reads the next stream while the others iterate, utilities.c:1601-1625.)
 */
int stsgpu_wait(stsgpu_ctx *ctx);
/* batches submitted and not yet waited for */
int stsgpu_pending(const stsgpu_ctx *ctx);

/*
 * Records of the last completed batch: pointers to engine-owned pinned host arrays of
 * nstreams*npv doubles, nstreams*nst int64 and nstreams*nw uint32, valid until the next stsgpu_wait.
 */
int stsgpu_results(stsgpu_ctx *ctx, int64_t *nstreams, int64_t *first_iteration,
		   const double **pvals, const int64_t **istats, const uint32_t **wcount);

/*
 * The double statistics (`fstats`, nstreams*nfs, layout.fs_off) of the same batch; *fstats is NULL when the engine
 * was created without STSGPU_FLAG_STATS.
 */
int stsgpu_results_stats(stsgpu_ctx *ctx, const double **fstats);

/* ---- synthetic input: tools/generators.c LCG, generator 1 ----------------------- */

/*
 * Fill the device input buffer with streams [first_stream, first_stream + nstreams) of the
 * reference's LCG generator output (generators.c:797-905: x <- 950706376 x mod (2^31-1) from
 * 23482349, bit = x >= 2^30, packed LSB-first by write_sequence), i.e. byte-identical to
 * `generators 1 <count>` for n = 2^20.  Uses skip-ahead, so any rank can start anywhere.
 */
int stsgpu_generate_lcg(stsgpu_ctx *ctx, int64_t first_stream, int64_t nstreams);

/*
 * The same for any generator of tools/generators.c that can jump ahead - 1 (LCG, as above), 2 (quadratic
 * congruential I, generators.c:915-975), 4 (cubic congruential, :1049-1108) and 5 (XOR, :1111-1186) - with every
 * behaviour of the tool that shapes its output (csrc/k_gen.cu lists them): byte-identical to `generators <g> <count>`
 * for n = 2^20.  n must be at least 1024.  Generators 3, 6, 7, 8 and 9 feed their whole state through a non-linear
 * map once per block - one chain for the whole run, which no number of threads can shorten - and are made on a host
 * core in front of stsgpu_submit (host/sts_generators.h); for them this call returns STSGPU_E_UNSUPPORTED.
 */
int stsgpu_generate(stsgpu_ctx *ctx, int generator, int64_t first_stream, int64_t nstreams);

/* copy the device input buffer back (tests / fixtures) */
int stsgpu_download_input(stsgpu_ctx *ctx, uint8_t *raw, int64_t nstreams);

/* ---- pinned host memory for the caller's read buffers -------------------------------------- */
/*
 * Page-locked memory for `raw` / `text` buffers, placed on the NUMA node of the calling thread's current CUDA
 * device (the device of the context created last) so that uploads do not cross sockets; STSGPU_NUMA=0 in the
 * environment leaves placement to the OS.
 */
void *stsgpu_host_alloc(size_t bytes);
void stsgpu_host_free(void *p);

/* ---- multi-GPU: gather records to a root rank over NCCL ------------------------- */

#define STSGPU_UNIQUE_ID_BYTES 128
/* rank 0 obtains an id and distributes it out of band (file, env, torchrun store ...) */
int stsgpu_comm_unique_id(uint8_t id[STSGPU_UNIQUE_ID_BYTES]);
int stsgpu_comm_init(stsgpu_ctx *ctx, const uint8_t id[STSGPU_UNIQUE_ID_BYTES], int rank, int nranks);
/*
 * Gather mode: from now on every batch this context submits ends with the exchange of its p-value records to
 * `root` (grouped ncclSend / ncclRecv queued behind the batch's tail kernels, then the copy to the root's pinned
 * memory), so the exchange overlaps the kernels of the next batches and stsgpu_wait returns with the gathered array
 * already on the root's host - no separate blocking call per batch.  All ranks must submit the same sequence of
 * batches; "-F a" batches are refused in this mode.  Needs an idle engine and stsgpu_comm_init.  root = -1 leaves
 * gather mode again (batches submitted afterwards are local).
 */
int stsgpu_gather_begin(stsgpu_ctx *ctx, int root);
/*
 * The p-values of the last completed batch of every rank on `root`, rank-major (rank r's streams follow rank
 * r-1's: the reference's -j job order, utilities.c:1526).  In gather mode this only hands out the array that
 * arrived with the batch; otherwise it performs one blocking exchange (collective: every rank calls it, before
 * submitting again into the slot of that batch).  On root, *pvals_all points to nranks*nstreams*npv doubles in
 * engine-owned pinned memory, valid until the batch's slot is waited for again; elsewhere it is NULL.  The ranks'
 * batch sizes travel with the data: STSGPU_E_STATE (never a hang) when a rank's batch had another nstreams.
 */
int stsgpu_gather_pvals(stsgpu_ctx *ctx, int root, const double **pvals_all);
int stsgpu_comm_barrier(stsgpu_ctx *ctx);

/* ---- assessment tallies on the device -------------------------------------------- */

/*
 * The uniformity / proportion assessment every X_metrics applies to a column of p-values
 * (e.g. frequency.c:828-894 tally, :631-700 thresholds and uniformity; identical in all 15 tests) only
 * needs, per p-value column, sampleCount, the count below alpha and the uniformity bin counts - all
 * additive.  After stsgpu_assess_begin every batch adds its records to those tallies on the device as
 * part of the batch, so a run of any length is assessed without keeping (or gathering) its p-values:
 * the reference appends every p-value to state->p_val and stores a 64-byte struct per template p-value
 * (defs.h:346-352), 79 GB at 2^23 streams.  Across GPUs the tallies are summed with one small
 * collective (stsgpu_assess_reduce) instead of gathering 1.5 KB per stream.
 */
typedef struct stsgpu_metric {
	int64_t sample_count;		/* p-values counted (excursion tests: only p > 0; NON_P_VALUE never) */
	int64_t too_low;		/* of those, p < alpha */
	int64_t pass_count;		/* sample_count - too_low (frequency.c:655-659) */
	double proportion_min, proportion_max;	/* thresholds, frequency.c:664-666 */
	double chi2, uniformity;	/* frequency.c:671-681: igamc((bins - 1) / 2, chi2 / 2) */
	int32_t uniformity_passed, proportion_passed;	/* the verdicts of frequency.c:739-757 */
} stsgpu_metric;

/* start (or restart) tallying: uniformity_bins = tp.uniformity_bins (sqrt(iterations), or 10 with -O),
 * alpha = tp.alpha (0.01), uniformity_level = tp.uniformity_level (0.0001); needs an idle engine */
int stsgpu_assess_begin(stsgpu_ctx *ctx, int64_t uniformity_bins, double alpha, double uniformity_level);
/*
 * Assess-only mode (-m a, read_from_p_val_file utilities.c:2051-2165): add `count` p-values of test `test` read from a
 * .pvalues file (host memory, file order = iteration-major) to the tallies; `first_index` is the position of pvals[0]
 * in that test's p-value sequence, so a file of any size can be streamed through a small buffer in pieces.
 * Value i lands in column pv_off[test] + (first_index + i) mod pv_cnt[test].  Synchronous.
 */
int stsgpu_assess_add(stsgpu_ctx *ctx, int test, const double *pvals, int64_t count, int64_t first_index);
/* sum the tallies of all ranks of the communicator into every rank (NCCL all-reduce, in place); no-op for one rank
 * and when already reduced; afterwards stsgpu_submit* / stsgpu_assess_add fail with STSGPU_E_STATE until the next
 * stsgpu_assess_begin (the tallies would otherwise count the other ranks' share again at the next reduce) */
int stsgpu_assess_reduce(stsgpu_ctx *ctx);
/*
 * Verdicts of everything tallied since stsgpu_assess_begin: metrics[npv] in record order, and, if freq
 * is not NULL, the bin counts freq[npv][uniformity_bins].  Needs an idle engine (every batch waited for).
 */
int stsgpu_assess_finish(stsgpu_ctx *ctx, stsgpu_metric *metrics, int64_t *freq);

/* ---- measurement --------------------------------------------------------------------------- */

/*
 * Per-kernel-group CUDA events are always recorded on the stream each group is launched on.
 * on = 1 additionally serialises all groups on one stream, so that each duration is that of the
 * group running alone (otherwise groups overlap on two streams per batch slot and durations include
 * sharing).  Kernel and batch times refer to the batch returned by the last stsgpu_wait.
 */
int stsgpu_set_profiling(stsgpu_ctx *ctx, int on);
/* kernels launched by the last submit; name/ms for entry i (ms valid after stsgpu_wait) */
int stsgpu_kernel_count(const stsgpu_ctx *ctx);
int stsgpu_kernel_info(stsgpu_ctx *ctx, int i, const char **name, float *ms);
/* device time of the last batch (first kernel start -> records ready to copy), in ms */
int stsgpu_last_batch_ms(stsgpu_ctx *ctx, float *ms);
/* total kernel launches issued by this context so far */
int64_t stsgpu_launch_count(const stsgpu_ctx *ctx);

/* ---- debugging / parity helpers ------------------------------------------------------------ */

/* cyclic pattern histogram at `bits` (<= 21) of stream `s` of the last batch, 2^bits uint32 */
int stsgpu_debug_pattern_histogram(stsgpu_ctx *ctx, int64_t s, int bits, uint32_t *hist);

const char *stsgpu_strerror(int code);
/* detail of the last failure on this context (CUDA error string, failing call) */
const char *stsgpu_last_error(const stsgpu_ctx *ctx);
int stsgpu_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif				/* STSGPU_H */
