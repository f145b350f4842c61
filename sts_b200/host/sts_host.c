/*
 * sts_host.c - batching loop and record keeping above the C ABI (see sts_host.h for the mapping to
 * the reference's functions).  Plain C99, pthread-free on the result path: records are appended in
 * iteration order by one thread, which removes the races the reference tolerates (runs.c:297,
 * serial.c:229-249) without changing any p_val content.
 */
#define _GNU_SOURCE
#define _DEFAULT_SOURCE
#include <dirent.h>
#include <errno.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include "sts_host.h"
#include "sts_generators.h"

const char *const sts_testNames[STS_NUMOFTESTS + 1] = {	/* parse_args.c:170-187 */
	"((all_tests))", "Frequency", "BlockFrequency", "CumulativeSums", "Runs", "LongestRun", "Rank", "DFT",
	"NonOverlappingTemplate", "OverlappingTemplate", "Universal", "ApproximateEntropy", "RandomExcursions",
	"RandomExcursionsVariant", "Serial", "LinearComplexity",
};

void
sts_err(int exitcode, const char *name, const char *fmt, ...)
{
	va_list ap;
	fflush(stdout);
	fprintf(stderr, "FATAL error in %s(): ", name);
	va_start(ap, fmt);
	vfprintf(stderr, fmt, ap);
	va_end(ap);
	fputc('\n', stderr);
	exit(exitcode);
}

void
sts_defaultstate(struct sts_state *s)
{
	memset(s, 0, sizeof(*s));
	s->tp.blockFrequencyBlockLength = 16384;	/* defs.h:110-125 */
	s->tp.nonOverlappingTemplateLength = 9;
	s->tp.overlappingTemplateLength = 9;
	s->tp.approximateEntropyBlockLength = 10;
	s->tp.serialBlockLength = 16;
	s->tp.linearComplexitySequenceLength = 500;
	s->tp.numOfBitStreams = 1;
	s->tp.uniformityBins = 10;
	s->tp.n = 1048576;
	s->tp.uniformityLevel = 0.0001;
	s->tp.alpha = 0.01;
	s->runMode = STS_MODE_ITERATE_AND_ASSESS;
	s->dataFormat = STS_FORMAT_RAW_BINARY;
	s->numberOfThreads = 0;			/* 0: not given, min(cores, 16) reader threads (parse_args.c:808-810 takes the cores) */
	s->batchStreams = 1024;
	s->workDir = strdup(".");
}

/* parse_args.c:960-1013 */
void
sts_change_params(struct sts_state *s, long num, long value, double d_value)
{
	switch (num) {
	case 1: s->tp.blockFrequencyBlockLength = value; break;
	case 2: s->tp.nonOverlappingTemplateLength = value; break;
	case 3: s->tp.overlappingTemplateLength = value; break;
	case 4: s->tp.approximateEntropyBlockLength = value; break;
	case 5: s->tp.serialBlockLength = value; break;
	case 6: s->tp.linearComplexitySequenceLength = value; break;
	case 7: s->tp.numOfBitStreams = value; break;
	case 8: s->tp.uniformityBins = value; s->uniformityBinsFlag = true; break;
	case 9: s->tp.n = value; break;
	case 10: s->tp.uniformityLevel = d_value; break;
	case 11: s->tp.alpha = d_value; break;
	default: sts_err(3, __func__, "invalid parameter option: %ld", num);
	}
}

static void
pv_append(struct sts_pvals *a, double v)
{
	if (a->count == a->cap) {
		long ncap = a->cap ? a->cap * 2 : 1024;
		double *nv = realloc(a->v, (size_t) ncap * sizeof(double));
		if (nv == NULL)
			sts_err(61, __func__, "cannot grow p-value array to %ld doubles", ncap);
		a->v = nv;
		a->cap = ncap;
	}
	a->v[a->count++] = v;
}

/*
 * init(): the reference runs X_init for tests 1..15 (driver.c:326-330), each of which may disable
 * itself; here stsgpu_create applies the same rules and reports them through the layout.
 */
void
sts_init(struct sts_state *s)
{
	stsgpu_params p;
	stsgpu_default_params(&p);
	p.device = s->device;
	p.n = s->tp.n;
	p.max_streams = s->batchStreams < s->tp.numOfBitStreams ? s->batchStreams : s->tp.numOfBitStreams;
	if (p.max_streams < 1 || s->runMode == STS_MODE_ASSESS_ONLY)
		p.max_streams = 1;
	p.test_mask = 0;
	for (int t = 1; t <= STS_NUMOFTESTS; t++)
		if (s->testVector[t])
			p.test_mask |= 1u << t;
	p.block_frequency_M = s->tp.blockFrequencyBlockLength;
	p.non_overlapping_m = (int32_t) s->tp.nonOverlappingTemplateLength;
	p.overlapping_m = (int32_t) s->tp.overlappingTemplateLength;
	p.apen_m = (int32_t) s->tp.approximateEntropyBlockLength;
	p.serial_m = (int32_t) s->tp.serialBlockLength;
	p.linear_complexity_M = s->tp.linearComplexitySequenceLength;
	p.pipeline_depth = STS_PIPELINE_DEPTH;			/* sts_invokeTestSuite keeps this many batches in flight */
	if ((s->tp.n % 8) != 0)					/* parse_args.c:939 */
		sts_err(1, __func__, "bitcount(n): %ld must be a multiple of 8", s->tp.n);
	int rc = stsgpu_create(&p, &s->engine);
	if (rc != STSGPU_OK)		/* no CPU fallback: fatal, like every error in the reference */
		sts_err(50, __func__, "%s: %s", stsgpu_strerror(rc), stsgpu_last_error(NULL));
	stsgpu_get_layout(s->engine, &s->layout);
	if (s->runMode != STS_MODE_ITERATE_ONLY) {
		/* parse_args.c:791-803: bins = sqrt(iterations) unless -P 8 was given; -O forces 10.  With -m a the
		 * iterations are those of -i (default 1), not the count found in the file names (parse_args.c:848) */
		if (!s->uniformityBinsFlag)
			s->tp.uniformityBins = (long) sqrt((double) (s->binsIterations > 0 ? s->binsIterations : s->tp.numOfBitStreams));
		if (s->legacy_output)
			s->tp.uniformityBins = 10;
		if (s->tp.uniformityBins < 1)
			s->tp.uniformityBins = 1;
		rc = stsgpu_assess_begin(s->engine, s->tp.uniformityBins, s->tp.alpha, s->tp.uniformityLevel);
		if (rc != STSGPU_OK)
			sts_err(51, __func__, "%s: %s", stsgpu_strerror(rc), stsgpu_last_error(s->engine));
	}
	for (int t = 1; t <= STS_NUMOFTESTS; t++) {
		bool on = (s->layout.enabled_mask >> t) & 1u;
		if (s->testVector[t] && !on)
			fprintf(stderr, "Warning: disabling test %s[%d]: preconditions on n / parameters not met\n",
				sts_testNames[t], t);
		s->testVector[t] = on;
		s->partitionCount[t] = on ? s->layout.pv_cnt[t] : 0;
	}
	struct stat st;
	if (stat(s->workDir, &st) != 0 && mkdir(s->workDir, 0777) != 0)
		sts_err(50, __func__, "cannot create workDir %s: %s", s->workDir, strerror(errno));
}

/*
 * R1, the record-keeping tail every reference X_iterate ends with (frequency.c:218-260 and its 16
 * copies): classify each p-value against alpha, bump the counters, append to p_val[test].
 * One generic body serves the 15 table entries because the arithmetic already happened on the GPU.
 */
static void
record_test(struct sts_thread_state *ts, int t)
{
	struct sts_state *s = ts->global_state;
	if (s->testVector[t] != true)
		return;
	if (ts->mutex != NULL)
		pthread_mutex_lock(ts->mutex);
	const long it = ts->iteration_being_done - s->batch_first;
	if (it < 0 || it >= s->batch_count)
		sts_err(51, __func__, "iteration %ld is not in the current batch", ts->iteration_being_done);
	const double *pv = s->batch_pvals + it * s->layout.npv + s->layout.pv_off[t];
	for (int k = 0; k < s->layout.pv_cnt[t]; k++) {
		const double p_value = pv[k];
		s->count[t]++;
		if (p_value == STS_NON_P_VALUE) {
			/* test not possible for this stream (runs.c:292-323, randomExcursions.c:597-600) */
		} else {
			s->valid[t]++;
			if (p_value < 0.0) {
				s->failure[t]++;
				fprintf(stderr, "Warning: iteration %ld of test %s[%d] produced bogus p_value: %f < 0.0\n",
					ts->iteration_being_done + 1, sts_testNames[t], t, p_value);
			} else if (p_value > 1.0) {
				s->failure[t]++;
				fprintf(stderr, "Warning: iteration %ld of test %s[%d] produced bogus p_value: %f > 1.0\n",
					ts->iteration_being_done + 1, sts_testNames[t], t, p_value);
			} else if (p_value < s->tp.alpha) {
				s->valid_p_val[t]++;
				s->failure[t]++;
			} else {
				s->valid_p_val[t]++;
				s->success[t]++;
			}
		}
		if (s->runMode == STS_MODE_ITERATE_ONLY || s->resultstxtFlag)
			pv_append(&s->p_val[t], p_value);	/* -m b without -s: the GPU tallies are the assessment, nothing else reads them */
	}
	if (ts->mutex != NULL)
		pthread_mutex_unlock(ts->mutex);
}

#define DEF_ITERATE(NAME, T) static void NAME##_iterate(struct sts_thread_state *ts) { record_test(ts, T); }
DEF_ITERATE(Frequency, 1) DEF_ITERATE(BlockFrequency, 2) DEF_ITERATE(CumulativeSums, 3) DEF_ITERATE(Runs, 4)
DEF_ITERATE(LongestRunOfOnes, 5) DEF_ITERATE(Rank, 6) DEF_ITERATE(DiscreteFourierTransform, 7)
DEF_ITERATE(NonOverlappingTemplateMatchings, 8) DEF_ITERATE(OverlappingTemplateMatchings, 9)
DEF_ITERATE(Universal, 10) DEF_ITERATE(ApproximateEntropy, 11) DEF_ITERATE(RandomExcursions, 12)
DEF_ITERATE(RandomExcursionsVariant, 13) DEF_ITERATE(Serial, 14) DEF_ITERATE(LinearComplexity, 15)

static void noop_state(struct sts_state *s) { (void) s; }

/* data_filename_format, utilities.c:479-538: "data.txt" unless the partition count exceeds a
 * power of ten: the digits are the largest d with 10^d < partitionCount (0 -> data.txt, else data%0<d>d.txt), so 2 and
 * 8 partitions all write data.txt, 18 write data1 .. data18 and 148 write data01 .. data148 */
static void
data_filename(char *out, size_t len, int partitionCount, int j)
{
	int digits = 0;
	if (partitionCount >= 2)
		for (digits = 9; digits > 0; --digits)			/* MAX_DATA_DIGITS */
			if (pow(10.0, digits) < (double) partitionCount)
				break;
	if (digits < 1)
		snprintf(out, len, "data.txt");
	else
		snprintf(out, len, "data%0*d.txt", digits, j);
}

static void
print_p_value(FILE *f, double p_value)			/* e.g. Frequency_print_p_value, frequency.c:378-408 */
{
	if (p_value == STS_NON_P_VALUE)
		fprintf(f, "__INVALID__\n");
	else
		fprintf(f, "%f\n", p_value);
}

/*
 * The -s files every X_print writes that come from p-values alone (e.g. frequency.c:424-618):
 * <workDir>/<TestName>/results.txt with every p-value in order, and, for tests with several p-values per
 * iteration, data*.txt with the values of one partition each.  stats.txt (the per-iteration intermediate values
 * of each test's private struct) is not produced.
 */
static void
print_test(struct sts_state *s, int t)
{
	if (!s->testVector[t] || !s->resultstxtFlag)
		return;
	const long count = s->p_val[t].count;
	const int parts = s->partitionCount[t];
	if (count != s->tp.numOfBitStreams * parts)
		sts_err(74, __func__, "print driver interface for %s[%d] called with p_val count: %ld != %ld*%d",
			sts_testNames[t], t, count, s->tp.numOfBitStreams, parts);
	char dir[4096], path[4400], name[64];
	snprintf(dir, sizeof(dir), "%s/%s", s->workDir, sts_testNames[t]);
	struct stat sb;
	if (stat(dir, &sb) != 0 && mkdir(dir, 0777) != 0)
		sts_err(74, __func__, "cannot create %s: %s", dir, strerror(errno));
	snprintf(path, sizeof(path), "%s/results.txt", dir);
	FILE *f = fopen(path, "w");
	if (f == NULL)
		sts_err(74, __func__, "cannot open %s: %s", path, strerror(errno));
	for (long i = 0; i < count; i++)
		print_p_value(f, s->p_val[t].v[i]);
	fclose(f);
	if (parts > 1) {
		for (int j = 0; j < parts; j++) {
			data_filename(name, sizeof(name), parts, j + 1);
			snprintf(path, sizeof(path), "%s/%s", dir, name);
			f = fopen(path, "w");		/* openTruncate: with data.txt the last partition wins, as in the reference */
			if (f == NULL)
				sts_err(74, __func__, "cannot open %s: %s", path, strerror(errno));
			for (long i = j; i < count; i += parts)
				print_p_value(f, s->p_val[t].v[i]);
			fclose(f);
		}
	}
}

#define DEF_PRINT(NAME, T) static void NAME##_print(struct sts_state *s) { print_test(s, T); }
DEF_PRINT(Frequency, 1) DEF_PRINT(BlockFrequency, 2) DEF_PRINT(CumulativeSums, 3) DEF_PRINT(Runs, 4)
DEF_PRINT(LongestRunOfOnes, 5) DEF_PRINT(Rank, 6) DEF_PRINT(DiscreteFourierTransform, 7)
DEF_PRINT(NonOverlappingTemplateMatchings, 8) DEF_PRINT(OverlappingTemplateMatchings, 9)
DEF_PRINT(Universal, 10) DEF_PRINT(ApproximateEntropy, 11) DEF_PRINT(RandomExcursions, 12)
DEF_PRINT(RandomExcursionsVariant, 13) DEF_PRINT(Serial, 14) DEF_PRINT(LinearComplexity, 15)

/* driver.c:434-470 */
void
sts_print(struct sts_state *s)
{
	for (int i = 1; i <= STS_NUMOFTESTS; i++)
		if (s->testVector[i] == true && sts_testDriver[i].print != NULL)
			sts_testDriver[i].print(s);
}

/* the plugin table, driver.c:63-192: same order, same five phases */
const struct sts_driver sts_testDriver[STS_NUMOFTESTS + 1] = {
	{ NULL, NULL, NULL, NULL, NULL },
	{ noop_state, Frequency_iterate, Frequency_print, noop_state, noop_state },
	{ noop_state, BlockFrequency_iterate, BlockFrequency_print, noop_state, noop_state },
	{ noop_state, CumulativeSums_iterate, CumulativeSums_print, noop_state, noop_state },
	{ noop_state, Runs_iterate, Runs_print, noop_state, noop_state },
	{ noop_state, LongestRunOfOnes_iterate, LongestRunOfOnes_print, noop_state, noop_state },
	{ noop_state, Rank_iterate, Rank_print, noop_state, noop_state },
	{ noop_state, DiscreteFourierTransform_iterate, DiscreteFourierTransform_print, noop_state, noop_state },
	{ noop_state, NonOverlappingTemplateMatchings_iterate, NonOverlappingTemplateMatchings_print, noop_state, noop_state },
	{ noop_state, OverlappingTemplateMatchings_iterate, OverlappingTemplateMatchings_print, noop_state, noop_state },
	{ noop_state, Universal_iterate, Universal_print, noop_state, noop_state },
	{ noop_state, ApproximateEntropy_iterate, ApproximateEntropy_print, noop_state, noop_state },
	{ noop_state, RandomExcursions_iterate, RandomExcursions_print, noop_state, noop_state },
	{ noop_state, RandomExcursionsVariant_iterate, RandomExcursionsVariant_print, noop_state, noop_state },
	{ noop_state, Serial_iterate, Serial_print, noop_state, noop_state },
	{ noop_state, LinearComplexity_iterate, LinearComplexity_print, noop_state, noop_state },
};

/* driver.c:395-425 */
void
sts_iterate(struct sts_thread_state *ts)
{
	if (ts == NULL || ts->global_state == NULL)
		sts_err(52, __func__, "thread_state arg is NULL");
	for (int i = 1; i <= STS_NUMOFTESTS; i++)
		if (ts->global_state->testVector[i] == true)
			sts_testDriver[i].iterate(ts);
}

static double
now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

/*
 * One batch is 128 MiB of file (1 GiB of text with -F a): a single fread from the page cache runs at a few GB/s, the
 * GPU tests 14 GB/s, so the file is read by -T threads (default min(cores, 16)) with pread, each a contiguous slice.
 * This is the work the reference spreads over its -T iteration threads (utilities.c:1601-1625).
 */
struct read_slice {
	int fd;
	uint8_t *dst;
	size_t len;
	off_t off;
	ssize_t got;
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
	r->got = (ssize_t) done;
	return NULL;
}

/* read `len` bytes at `off` of the data file into `dst`; returns the bytes read (short only at end of file) */
static size_t
read_parallel(struct sts_state *s, uint8_t *dst, size_t len, off_t off)
{
	long nt = s->numberOfThreads;
	if (nt <= 0) {
		nt = sysconf(_SC_NPROCESSORS_ONLN);
		if (nt > 16) nt = 16;
	}
	if (nt < 1) nt = 1;
	if (nt > 64) nt = 64;
	if (len < ((size_t) 8 << 20))
		nt = 1;
	struct read_slice sl[64];
	/* ceil(len / nt) rounded up to 4 KiB: at most nt slices, whatever len is */
	const size_t per = (((len + (size_t) nt - 1) / (size_t) nt) + 4095) & ~(size_t) 4095;
	int used = 0;
	for (size_t o = 0; o < len && used < 64; o += per, used++) {
		sl[used].fd = fileno(s->streamFile);
		sl[used].dst = dst + o;
		sl[used].len = (len - o < per) ? len - o : per;
		sl[used].off = off + (off_t) o;
		sl[used].got = 0;
		if (used > 0 && pthread_create(&sl[used].tid, NULL, read_slice_main, &sl[used]) != 0)
			sts_err(213, __func__, "cannot start a reader thread");
	}
	read_slice_main(&sl[0]);
	size_t total = 0;
	bool short_seen = false;
	for (int i = 0; i < used; i++) {
		if (i > 0)
			pthread_join(sl[i].tid, NULL);
		if (!short_seen)
			total += (size_t) sl[i].got;
		if ((size_t) sl[i].got < sl[i].len)
			short_seen = true;
	}
	return total;
}

/*
 * Fill `dst` with `streams` packed streams from the data file.
 * -F r: the file bytes ARE the packed form (utilities.c:1739-1818 + copyBitsToEpsilon :1837-1884).
 * -F a: ASCII '0'/'1' characters, whitespace ignored (parseBitsASCIIInput :1651-1727, "%1d"); packed
 *       MSB first here on the host.
 * Returns the number of complete streams read (a short raw read is fatal like utilities.c:1789-1791).
 */
static long
read_streams(struct sts_state *s, uint8_t *dst, long streams)
{
	const long nb = s->tp.n / 8;
	if (s->dataFormat == STS_FORMAT_RAW_BINARY) {
		size_t got = fread(dst, (size_t) nb, (size_t) streams, s->streamFile);
		if ((long) got != streams)
			sts_err(213, __func__, "read %zu of %ld streams of %ld bytes from %s: file too short",
				got, streams, nb, s->randomDataPath);
		return streams;
	}
	for (long k = 0; k < streams; k++) {
		uint8_t *o = dst + k * nb;
		memset(o, 0, (size_t) nb);
		for (long bit = 0; bit < s->tp.n;) {
			int c = fgetc(s->streamFile);
			if (c == EOF) {
				fprintf(stderr, "Warning: insufficient data in %s, %ld bits were read\n", s->randomDataPath,
					k * s->tp.n + bit);
				return k;
			}
			if (c == '0' || c == '1') {
				if (c == '1')
					o[bit >> 3] |= (uint8_t) (0x80u >> (bit & 7));
				bit++;
			} else if (c != ' ' && c != '\n' && c != '\r' && c != '\t') {
				sts_err(214, __func__, "found a character different than 1 or 0 in the sequence: %c", c);
			}
		}
	}
	return streams;
}

/*
 * invokeTestSuite(): replaces the pthread pool of handleFileBasedBitStreams/testBits.  Seek rule of
 * the reference (utilities.c:1511, :1526): job `jobnum` starts at byte jobnum * n * iterations / 8 of a raw
 * file, at character jobnum * n * iterations of an ASCII file; standard input ("-") is never seeked (:1500).
 * Pipelined STS_PIPELINE_DEPTH (3) deep: while the GPU tests batch k, batch k+1 is already uploaded and the host
 * reads and submits batch k+2 from the third pinned buffer, then runs the record keeping of batch k.  (With
 * two in flight the upload of the newest batch - 2.4 ms of the 9.7 ms a batch takes - leaves the GPU with
 * the kernels of a single batch: measured 103.5 against 108.9 Gbit/s end to end.)
 *
 * ASCII files go to the GPU as text (stsgpu_submit_ascii): iteration k is the first n digits at or after
 * character base_seek + k n, the reference's per-iteration fseek (utilities.c:1682-1683).  ASCII on standard
 * input cannot seek, there the digits are consumed in sequence and packed here (read_streams), as the
 * reference does for stdin.
 */
void
sts_invokeTestSuite(struct sts_state *s)
{
	const long n = s->tp.n;
	const long nb = n / 8;
	const long B = s->batchStreams < s->tp.numOfBitStreams ? s->batchStreams : s->tp.numOfBitStreams;
	const bool chained = sts_gen_is_chained((int) s->generator);	/* 3, 6, 7, 8, 9: one host thread, sts_generators.c */
	const bool generated = (s->generator != 0) && !chained;		/* 1, 2, 4, 5: made on the GPU */
	const bool no_file = generated || chained;
	const bool from_stdin = !no_file && (strcmp(s->randomDataPath, "-") == 0);
	const bool gpu_ascii = !no_file && (s->dataFormat == STS_FORMAT_ASCII_01) && !from_stdin;
	s->streamFile = no_file ? NULL : (from_stdin ? stdin : fopen(s->randomDataPath, "rb"));
	if (s->streamFile == NULL && !no_file)
		sts_err(210, __func__, "cannot open data file %s: %s", s->randomDataPath, strerror(errno));
	long base_seek = 0, file_size = 0;
	if (!from_stdin && !no_file) {
		const long per_stream = (s->dataFormat == STS_FORMAT_RAW_BINARY) ? nb : n;
		base_seek = s->jobnum * per_stream * s->tp.numOfBitStreams;
		if (fseek(s->streamFile, 0, SEEK_END) != 0 || (file_size = ftell(s->streamFile)) < 0 ||
		    fseek(s->streamFile, base_seek, SEEK_SET) != 0)
			sts_err(211, __func__, "cannot seek to job %ld in %s", s->jobnum, s->randomDataPath);
	}
	const size_t buf_bytes = gpu_ascii ? (size_t) (B + 1) * (size_t) n : (size_t) B * (size_t) nb;
	/* text is 8 bytes per tested byte: PCIe bound at any depth; generated input is resident: two suffice */
	const int depth = (gpu_ascii || generated) ? 2 : STS_PIPELINE_DEPTH;
	uint8_t *buf[STS_PIPELINE_DEPTH] = { NULL };
	for (int i = 0; i < depth && !generated; i++) {
		buf[i] = stsgpu_host_alloc(buf_bytes);
		if (buf[i] == NULL)
			sts_err(212, __func__, "cannot allocate pinned read buffer of %zu bytes", buf_bytes);
	}
	struct sts_thread_state ts = { 0, s, 0, NULL };
	const long total = s->tp.numOfBitStreams;
	struct sts_gen *chain = NULL;
	if (chained) {
		chain = sts_gen_open((int) s->generator, n);
		if (chain == NULL)
			sts_err(218, __func__, "cannot start generator %ld for streams of %ld bits", s->generator, n);
		/* job `jobnum` tests streams [jobnum * iterations, ...) of the tool's output: a chain has to be run up to there */
		for (long skip = s->jobnum * total; skip > 0;) {
			const long k = skip < B ? skip : B;
			sts_gen_next(chain, buf[0], k);
			skip -= k;
		}
	}
	double t_read = 0.0, t_wait = 0.0, t_keep = 0.0;
	const double t_begin = now_s();
	long submitted = 0, done = 0;
	int w = 0, eof = 0;
	long waited = 0;				/* batches completed: batch k was read into buf[k % depth] */
	long batch_want[STS_PIPELINE_DEPTH] = { 0 };
	bool window_at_eof[STS_PIPELINE_DEPTH] = { false };	/* ASCII: did the text window of the batch in buf[i] end at EOF? */
	bool short_seen = false;
	for (;;) {
		/* keep `depth` batches in flight: the read and the upload of one overlap the kernels of the others */
		while (stsgpu_pending(s->engine) < depth && !eof && submitted < total) {
			const long want = (total - submitted) < B ? (total - submitted) : B;
			long have = want;
			int rc;
			const double t_r0 = now_s();
			if (generated) {
				/* job `jobnum` of the reference's distributed mode would read streams [jobnum * iterations, ...) of
				 * the generator's file: skip ahead to the same place */
				rc = stsgpu_generate(s->engine, (int) s->generator, s->jobnum * total + submitted, want);
				if (rc == STSGPU_OK)
					rc = stsgpu_submit_resident(s->engine, want, submitted);
			} else if (chained) {
				rc = sts_gen_next(chain, buf[w % depth], want) == 0 ? stsgpu_submit(s->engine, buf[w % depth], want, submitted)
										     : STSGPU_E_INVAL;
			} else if (gpu_ascii) {
				const long off = base_seek + submitted * n;
				long len = (want + 1) * n;
				if (off >= file_size) {
					fprintf(stderr, "Warning: insufficient data in %s, %ld bits were read\n", s->randomDataPath, 0L);
					eof = 1;
					break;
				}
				window_at_eof[w % depth] = (off + len >= file_size);
				if (off + len > file_size)
					len = file_size - off;
				if ((long) read_parallel(s, buf[w % depth], (size_t) len, (off_t) off) != len)
					sts_err(213, __func__, "cannot read %ld characters at offset %ld of %s", len, off, s->randomDataPath);
				rc = stsgpu_submit_ascii(s->engine, (const char *) buf[w % depth], len, want, submitted);
			} else if (!from_stdin && s->dataFormat == STS_FORMAT_RAW_BINARY) {
				const off_t off = (off_t) base_seek + (off_t) submitted * nb;
				const size_t got = read_parallel(s, buf[w % depth], (size_t) want * (size_t) nb, off);
				if (got != (size_t) want * (size_t) nb)		/* utilities.c:1789-1791: a short raw read is fatal */
					sts_err(213, __func__, "read %zu of %ld streams of %ld bytes from %s: file too short",
						got / (size_t) nb, want, nb, s->randomDataPath);
				rc = stsgpu_submit(s->engine, buf[w % depth], have, submitted);
			} else {
				have = read_streams(s, buf[w % depth], want);
				if (have < want)
					eof = 1;
				if (have <= 0)
					break;
				rc = stsgpu_submit(s->engine, buf[w % depth], have, submitted);
			}
			if (rc != STSGPU_OK)
				sts_err(215, __func__, "%s: %s", stsgpu_strerror(rc), stsgpu_last_error(s->engine));
			t_read += now_s() - t_r0;
			batch_want[w % depth] = have;
			submitted += have;
			w++;
		}
		if (stsgpu_pending(s->engine) == 0)
			break;
		const double t_w0 = now_s();
		int rc = stsgpu_wait(s->engine);
		if (rc != STSGPU_OK)
			sts_err(216, __func__, "%s: %s", stsgpu_strerror(rc), stsgpu_last_error(s->engine));
		const double t_k0 = now_s();
		t_wait += t_k0 - t_w0;
		int64_t bn, bf;
		stsgpu_results(s->engine, &bn, &bf, &s->batch_pvals, NULL, NULL);
		if (gpu_ascii) {
			const int slot = (int) (waited % depth);
			int64_t complete, bad;
			stsgpu_ascii_status(s->engine, &complete, &bad);
			if (bad >= 0 && !short_seen)
				sts_err(214, __func__, "found a character different than 1 or 0 in the sequence: %c (offset %ld)",
					buf[slot][bad], (long) (base_seek + bf * n + bad));
			if (complete < batch_want[slot] && !short_seen) {
				short_seen = true;
				if (!window_at_eof[slot])
					sts_err(217, __func__, "more than %ld white-space characters inside one batch of %s", n,
						s->randomDataPath);
				fprintf(stderr, "Warning: insufficient data in %s: iteration %ld is incomplete\n", s->randomDataPath,
					(long) (bf + complete));
				eof = 1;		/* every later iteration starts further into the text and is incomplete too */
			}
		}
		waited++;
		s->batch_first = (long) bf;
		s->batch_count = (long) bn;
		for (long k = 0; k < (long) bn; k++) {		/* iteration order: reproducible .pvalues */
			ts.iteration_being_done = done + k;
			sts_iterate(&ts);
			if (s->reportCycle > 0 && ((done + k + 1) % s->reportCycle) == 0)
				printf("Completed %ld iterations of %ld\n", done + k + 1, s->tp.numOfBitStreams);
		}
		done += (long) bn;
		t_keep += now_s() - t_k0;
	}
	if (s->debuglevel > 0) {
		const double dt = now_s() - t_begin;
		printf("iterate phase: %ld streams of %ld bits in %.3f s = %.2f Gbit/s (read + submit %.3f s, waiting for the GPU %.3f s, "
		       "record keeping %.3f s)\n", done, n, dt, (double) done * (double) n / dt / 1e9, t_read, t_wait, t_keep);
	}
	s->tp.numOfBitStreams = done;		/* ASCII input may end early (utilities.c:1693-1700) */
	for (int i = 0; i < depth; i++)
		stsgpu_host_free(buf[i]);
	if (!from_stdin && !no_file)
		fclose(s->streamFile);
	sts_gen_close(chain);
	s->streamFile = NULL;
}

/* utilities.c:1938-2042: int64 test, int64 count, count doubles, per enabled test; .work then rename */
void
sts_write_p_val_to_file(struct sts_state *s)
{
	char work[4096], fin[4096];
	snprintf(work, sizeof(work), "%s/sts.%04ld.%ld.%ld.work", s->workDir, s->jobnum, s->tp.numOfBitStreams, s->tp.n);
	snprintf(fin, sizeof(fin), "%s/sts.%04ld.%ld.%ld.pvalues", s->workDir, s->jobnum, s->tp.numOfBitStreams, s->tp.n);
	FILE *f = fopen(work, "wb");
	if (f == NULL)
		sts_err(232, __func__, "cannot open %s: %s", work, strerror(errno));
	for (long i = 1; i <= STS_NUMOFTESTS; i++) {
		if (s->testVector[i] != true)
			continue;
		long cnt = s->p_val[i].count;
		if (fwrite(&i, sizeof(i), 1, f) != 1 || fwrite(&cnt, sizeof(cnt), 1, f) != 1 ||
		    (cnt > 0 && fwrite(s->p_val[i].v, sizeof(double), (size_t) cnt, f) != (size_t) cnt))
			sts_err(232, __func__, "error while writing p-values of test %ld to %s", i, work);
	}
	fclose(f);
	if (rename(work, fin) < 0)
		sts_err(232, __func__, "error in renaming %s to %s", work, fin);
}

/* does `name` look like sts.<job>.<iterations>.<n>.pvalues (parse_args.c:866-898)?  returns the iterations or -1 */
static long
p_val_file_iterations(const char *name, long n)
{
	char *copy = strdup(name), *rest = copy, *tok[5];
	int k = 0;
	while (k < 5 && (tok[k] = strsep(&rest, ".")) != NULL)
		k++;
	long it = -1;
	if (k == 5 && strcmp(tok[0], "sts") == 0 && strcmp(tok[4], "pvalues") == 0 && atoi(tok[3]) == n)
		it = atoi(tok[2]);
	free(copy);
	return it;
}

/* parse_args.c:837-921: the iterations of an assess-only run are the sum of those in the matching file names */
void
sts_find_p_val_files(struct sts_state *s)
{
	DIR *dir = opendir(s->pvalues_dir);
	if (dir == NULL)
		sts_err(1, __func__, "Could not open the directory: %s", s->pvalues_dir);
	struct dirent *e;
	s->tp.numOfBitStreams = 0;
	while ((e = readdir(dir)) != NULL) {
		const long it = p_val_file_iterations(e->d_name, s->tp.n);
		if (it > 0)
			s->tp.numOfBitStreams += it;
	}
	closedir(dir);
}

/*
 * read_from_p_val_file (utilities.c:2051-2165): records of {int64 test, int64 count, count doubles}.  The
 * reference appends every value to state->p_val (64 bytes per template p-value); here each file is streamed
 * through one 8 MiB buffer into the tallies on the GPU, so memory does not grow with the run.
 */
void
sts_read_from_p_val_file(struct sts_state *s)
{
	enum { CHUNK = 1 << 20 };
	double *buf = malloc(CHUNK * sizeof(double));
	if (buf == NULL)
		sts_err(232, __func__, "cannot allocate the p-value read buffer");
	long seen[STS_NUMOFTESTS + 1] = { 0 };
	DIR *dir = opendir(s->pvalues_dir);
	if (dir == NULL)
		sts_err(1, __func__, "Could not open the directory: %s", s->pvalues_dir);
	struct dirent *e;
	while ((e = readdir(dir)) != NULL) {
		if (p_val_file_iterations(e->d_name, s->tp.n) <= 0)
			continue;
		char path[4096];
		snprintf(path, sizeof(path), "%s/%s", s->pvalues_dir, e->d_name);
		FILE *f = fopen(path, "rb");
		if (f == NULL) {
			fprintf(stderr, "Warning: skipping p-value file due to error in opening p-value file: %s\n", e->d_name);
			continue;
		}
		long test_num = 0, count;
		do {
			if (fread(&test_num, sizeof(test_num), 1, f) != 1 || fread(&count, sizeof(count), 1, f) != 1)
				break;
			if (test_num < 1 || test_num > STS_NUMOFTESTS || count < 0) {
				fprintf(stderr, "Warning: skipping p-value file, bad record header in %s\n", e->d_name);
				break;
			}
			const bool on = s->testVector[test_num];
			bool truncated = false;
			long done = 0;
			while (done < count) {
				const long c = (count - done) < CHUNK ? (count - done) : CHUNK;
				const long got = (long) fread(buf, sizeof(double), (size_t) c, f);
				if (got > 0 && on) {	/* the reference keeps the values it could read (utilities.c:2120-2133) */
					int rc = stsgpu_assess_add(s->engine, (int) test_num, buf, got, seen[test_num] + done);
					if (rc != STSGPU_OK)
						sts_err(53, __func__, "%s: %s", stsgpu_strerror(rc), stsgpu_last_error(s->engine));
				}
				done += got > 0 ? got : 0;
				if (got != c) {
					fprintf(stderr, "Warning: EOF while reading pvalue[%ld] from file: %s\n", done, e->d_name);
					truncated = true;
					break;
				}
			}
			seen[test_num] += done;
			if (truncated)
				break;
		} while (test_num < STS_NUMOFTESTS);		/* utilities.c:2153: the record of test 15 ends a file */
		fclose(f);
	}
	closedir(dir);
	free(buf);
	for (int t = 1; t <= STS_NUMOFTESTS; t++)
		if (s->testVector[t] && seen[t] != s->tp.numOfBitStreams * s->partitionCount[t])
			fprintf(stderr, "Warning: %s[%d]: %ld p-values found != bit streams: %ld\n", sts_testNames[t], t, seen[t],
				s->tp.numOfBitStreams * s->partitionCount[t]);
}

/* finishMetricTestsSentence, driver.c:1112-1146 */
enum { PASSED_BOTH = 0, FAILED_UNIFORMITY = 1, FAILED_PROPORTION = 2, FAILED_BOTH = 3 };
static const char *const verdict_sentence[4] = {
	"passed both the analyses.\n", "FAILED the uniformity analysis.\n", "FAILED the proportion analysis.\n",
	"FAILED both the analyses.\n"
};

static int
verdict_of(const stsgpu_metric *m)		/* e.g. frequency.c:757-766 */
{
	if (!m->proportion_passed && !m->uniformity_passed)
		return FAILED_BOTH;
	if (!m->proportion_passed)
		return FAILED_PROPORTION;
	if (!m->uniformity_passed)
		return FAILED_UNIFORMITY;
	return PASSED_BOTH;
}

/* result.txt, the non-legacy report of driver.c:676-1100, from the verdicts of the GPU tallies */
static void
write_result_txt(struct sts_state *s, const stsgpu_metric *m)
{
	static const char *const single[STS_NUMOFTESTS + 1] = {
		NULL, "Frequency", "Block Frequency", NULL, "Runs", "Longest Run of Ones", "Binary Matrix Rank",
		"Discrete Fourier Transform", NULL, "Overlapping Template Matching", "Maurer's Universal Statistical",
		"Approximate Entropy", NULL, NULL, NULL, "Linear Complexity"
	};
	static const char *const multi[STS_NUMOFTESTS + 1] = {
		NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, "Non-overlapping Template Matching", NULL, NULL, NULL,
		"Random Excursions", "Random Excursions Variant", NULL, NULL
	};
	static const char *const sep = "- - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - "
				       "- - - - - - - - - - - - - - -\n";
	char path[4096];
	snprintf(path, sizeof(path), "%s/result.txt", s->workDir);
	FILE *f = fopen(path, "w");
	if (f == NULL)
		sts_err(54, __func__, "cannot open %s: %s", path, strerror(errno));
	long total = 0;
	for (int t = 1; t <= STS_NUMOFTESTS; t++)
		if (s->testVector[t])
			total += s->partitionCount[t];
	const char *source = "--SOME_FILE--";
	if (s->randomDataPath != NULL)
		source = strcmp(s->randomDataPath, "-") == 0 ? "((standard input))" :
			 (strcmp(s->randomDataPath, "/dev/null") != 0 ? s->randomDataPath : "--UNKNOWN--");
	fprintf(f, "A total of %ld tests (some of the %d tests actually consist of multiple sub-tests)\n"
		   "were conducted to evaluate the randomness of %ld bitstreams of %ld bits from:\n\n\t%s\n\n",
		total, STS_NUMOFTESTS, s->tp.numOfBitStreams, s->tp.n, source);
	if (s->runMode == STS_MODE_ASSESS_ONLY)
		fprintf(f, "using the p-values from the following files:\n\n\t%s/sts.*.*.%ld.pvalues\n\n", s->pvalues_dir, s->tp.n);
	fprintf(f, "%s\n"
		   "The numerous empirical results of these tests were then interpreted with\nan "
		   "examination of the proportion of sequences that pass a statistical test\n"
		   "(proportion analysis) and the distribution of p-values to check for uniformity\n"
		   "(uniformity analysis). The results were the following:\n\n"
		   "\t%ld/%ld tests passed successfully both the analyses.\n"
		   "\t%ld/%ld tests did not pass successfully both the analyses.\n\n"
		   "%s\nHere are the results of the single tests:\n",
		sep, s->successful_tests, total, total - s->successful_tests, total, sep);
	for (int t = 1; t <= STS_NUMOFTESTS; t++) {
		if (!s->testVector[t])
			continue;
		const stsgpu_metric *mt = m + s->layout.pv_off[t];
		if (single[t] != NULL) {
			fprintf(f, "\n - The \"%s\" test %s", single[t], verdict_sentence[verdict_of(mt)]);
		} else if (multi[t] != NULL) {
			long by[4] = { 0, 0, 0, 0 };
			for (int j = 0; j < s->partitionCount[t]; j++)
				by[verdict_of(mt + j)]++;
			static const int order[4] = { PASSED_BOTH, FAILED_PROPORTION, FAILED_UNIFORMITY, FAILED_BOTH };
			bool first = true;
			for (int k = 0; k < 4; k++) {
				if (by[order[k]] <= 0)
					continue;
				fprintf(f, first ? "\n - " : "   ");
				first = false;
				fprintf(f, "%ld/%d of the \"%s\" tests %s", by[order[k]], s->partitionCount[t], multi[t],
					verdict_sentence[order[k]]);
			}
		} else {			/* CumulativeSums and Serial: two named sub-tests */
			const char *name = (t == 3) ? "Cumulative Sums" : "Serial";
			const char *a = (t == 3) ? "forward" : "first", *b = (t == 3) ? "backward" : "second";
			fprintf(f, "\n - The \"%s\" (%s) test %s", name, a, verdict_sentence[verdict_of(mt)]);
			fprintf(f, "   The \"%s\" (%s) test %s", name, b, verdict_sentence[verdict_of(mt + 1)]);
		}
	}
	fprintf(f, "\n%s\nThe missing tests (if any) were whether disabled manually by the user or disabled\n"
		   "at run time due to input size requirements not satisfied by this run.\n\n", sep);
	fclose(f);
}

/*
 * metrics() (driver.c:475): with -O the table every X_metrics prints in legacy format (e.g. frequency.c:700-735):
 * bin counts, uniformity p-value, passCount/sampleCount, test name, one line per p-value column in test
 * order, written to workDir/finalAnalysisReport.txt under the reference's header; without -O the summary of
 * workDir/result.txt (driver.c:676-1100).  The tallies come from the GPU (stsgpu_assess_*), so no p-value has
 * to be kept for this.
 */
void
sts_metrics(struct sts_state *s)
{
	if (s->runMode == STS_MODE_ITERATE_ONLY)
		return;
	const int npv = s->layout.npv;
	const long bins = s->tp.uniformityBins;
	stsgpu_metric *m = calloc((size_t) npv, sizeof(*m));
	int64_t *freq = calloc((size_t) npv * (size_t) bins, sizeof(*freq));
	if (m == NULL || freq == NULL)
		sts_err(52, __func__, "cannot allocate the assessment table");
	int rc = stsgpu_assess_finish(s->engine, m, freq);
	if (rc != STSGPU_OK)
		sts_err(53, __func__, "%s: %s", stsgpu_strerror(rc), stsgpu_last_error(s->engine));
	s->successful_tests = 0;
	for (int q = 0; q < npv; q++)
		if (verdict_of(m + q) == PASSED_BOTH)
			s->successful_tests++;
	if (!s->legacy_output) {
		write_result_txt(s, m);
		free(m);
		free(freq);
		return;
	}
	char path[4096];
	snprintf(path, sizeof(path), "%s/finalAnalysisReport.txt", s->workDir);
	FILE *f = fopen(path, "w");
	if (f == NULL)
		sts_err(54, __func__, "cannot open %s: %s", path, strerror(errno));
	fprintf(f, "------------------------------------------------------------------------------\n");
	fprintf(f, "RESULTS FOR THE UNIFORMITY OF P-VALUES AND THE PROPORTION OF PASSING SEQUENCES\n");
	fprintf(f, "------------------------------------------------------------------------------\n");
	fprintf(f, "   generator is <%s>\n", s->randomDataPath ? s->randomDataPath : "--SOME_FILE--");
	fprintf(f, "------------------------------------------------------------------------------\n");
	for (long i = 0; i < bins; i++) {
		char name[32];
		snprintf(name, sizeof(name), "C%ld", i + 1);
		fprintf(f, "%3s ", name);
	}
	fprintf(f, " P-VALUE  PROPORTION  STATISTICAL TEST\n");
	fprintf(f, "------------------------------------------------------------------------------\n");
	for (int t = 1; t <= STS_NUMOFTESTS; t++) {
		if (!((s->layout.enabled_mask >> t) & 1u))
			continue;
		for (int j = 0; j < s->layout.pv_cnt[t]; j++) {
			const int q = s->layout.pv_off[t] + j;
			for (long i = 0; i < bins; i++)
				fprintf(f, "%3ld ", (long) freq[(long) q * bins + i]);
			if ((double) m[q].sample_count / bins <= 0.0)
				fprintf(f, "    ----    ");
			else if (!m[q].uniformity_passed)
				fprintf(f, " %8.6f * ", m[q].uniformity);
			else
				fprintf(f, " %8.6f   ", m[q].uniformity);
			if (m[q].sample_count == 0)
				fprintf(f, " ------     %s\n", sts_testNames[t]);
			else if (!m[q].proportion_passed)
				fprintf(f, "%4ld/%-4ld *\t %s\n", (long) m[q].pass_count, (long) m[q].sample_count, sts_testNames[t]);
			else
				fprintf(f, "%4ld/%-4ld\t %s\n", (long) m[q].pass_count, (long) m[q].sample_count, sts_testNames[t]);
		}
	}
	fclose(f);
	free(m);
	free(freq);
}

void
sts_destroy(struct sts_state *s)
{
	if (s->engine)
		stsgpu_destroy(s->engine);
	s->engine = NULL;
	for (int t = 0; t <= STS_NUMOFTESTS; t++) {
		free(s->p_val[t].v);
		s->p_val[t].v = NULL;
	}
	free(s->workDir);
	free(s->randomDataPath);
	free(s->pvalues_dir);
	s->workDir = s->randomDataPath = s->pvalues_dir = NULL;
}
