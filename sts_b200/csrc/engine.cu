/*
 * engine.cu - the C ABI of include/stsgpu.h: context, parameter checks that mirror the X_init()
 * self-disable rules, device buffers, batch submission and result records.
 *
 * HBM layout of one context (n = 2^20, capacity B streams), per batch slot (engine.h):
 *   d_in        B x pitch bytes   raw file bytes of every stream, pitch = 128-byte multiple >= n/8 + 16
 *   d_pv/d_st/d_w                 B records (188 doubles, 176 int64, 1184 uint32 at defaults)
 *   d_apen_hist B x 3 * 2^apen_m  uint32, the two ApEn histograms in index order
 *   d_dft_work  chunk x n/2       double2, the four-step FFT intermediate (8 MiB per stream)
 * and once per context: templates, LC class LUT, log2 table (Universal), FFT twiddles.
 * `pipeline_depth` slots (default 1; the host driver and bench.py use 3) are used round robin so that consecutive
 * batches overlap.  "-F a" batches add a text staging buffer per slot, allocated on first use (k_ascii.cu).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <new>
#include <ctype.h>
#include <sched.h>
#include "common.cuh"
#include "tails.h"
#include "dft.h"
#include "engine.h"

/* ------------------------------------------------------------------------------------------ */
static void set_err(stsgpu_ctx *c, const char *what, cudaError_t e)
{
	if (c)
		snprintf(c->err, sizeof(c->err), "%s: %s", what, cudaGetErrorString(e));
}

#define CK(call)                                                                                                  \
	do {                                                                                                      \
		cudaError_t e_ = (call);                                                                          \
		if (e_ != cudaSuccess) {                                                                          \
			set_err(ctx, #call, e_);                                                                  \
			return STSGPU_E_CUDA;                                                                     \
		}                                                                                                 \
	} while (0)

/*
 * NUMA: pinned host buffers should live on the memory node the GPU's PCIe root hangs off - with eight ranks each
 * streaming ~13 GB/s of input, remote pages cross the socket interconnect.  pinned_alloc_near() binds the calling
 * thread to the CPUs of that node for the duration of the allocation (pages are placed by first touch inside
 * cudaMallocHost) and restores the affinity afterwards.  STSGPU_NUMA=0 turns it off.
 */
static int device_numa_node(int device)
{
	char bus[32];
	if (cudaDeviceGetPCIBusId(bus, (int) sizeof(bus), device) != cudaSuccess)
		return -1;
	for (char *c = bus; *c; c++)
		*c = (char) tolower((unsigned char) *c);
	char path[128];
	snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
	FILE *f = fopen(path, "r");
	if (f == NULL)
		return -1;
	int node = -1;
	if (fscanf(f, "%d", &node) != 1)
		node = -1;
	fclose(f);
	return node;
}

static bool node_cpus(int node, cpu_set_t *set)
{
	char path[128];
	snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
	FILE *f = fopen(path, "r");
	if (f == NULL)
		return false;
	CPU_ZERO(set);
	int a, b, any = 0;
	while (fscanf(f, "%d", &a) == 1) {		/* "0-31,64-95" */
		b = a;
		int c = fgetc(f);
		if (c == '-') {
			if (fscanf(f, "%d", &b) != 1)
				break;
			c = fgetc(f);
		}
		for (int k = a; k <= b && k < CPU_SETSIZE; k++) {
			CPU_SET(k, set);
			any = 1;
		}
		if (c != ',')
			break;
	}
	fclose(f);
	return any != 0;
}

static cudaError_t pinned_alloc_near(void **p, size_t bytes, int device)
{
	const char *env = getenv("STSGPU_NUMA");
	cpu_set_t old_set, want, both;
	bool bound = false;
	if (!(env && atoi(env) == 0) && sched_getaffinity(0, sizeof(old_set), &old_set) == 0) {
		const int node = device_numa_node(device);
		if (node >= 0 && node_cpus(node, &want)) {
			CPU_AND(&both, &want, &old_set);
			if (CPU_COUNT(&both) > 0 && sched_setaffinity(0, sizeof(both), &both) == 0)
				bound = true;
		}
	}
	cudaError_t e = cudaMallocHost(p, bytes);
	if (bound)
		sched_setaffinity(0, sizeof(old_set), &old_set);
	return e;
}

extern "C" const char *stsgpu_strerror(int code)
{
	switch (code) {
	case STSGPU_OK: return "success";
	case STSGPU_E_INVAL: return "invalid argument";
	case STSGPU_E_NO_DEVICE: return "no usable CUDA device (the engine has no CPU fallback)";
	case STSGPU_E_CUDA: return "CUDA call failed";
	case STSGPU_E_NOMEM: return "out of memory";
	case STSGPU_E_UNSUPPORTED: return "parameter combination not supported by the GPU engine";
	case STSGPU_E_STATE: return "call sequence error";
	case STSGPU_E_NCCL: return "NCCL unavailable or failed";
	case STSGPU_E_RANGE: return "index out of range";
	default: return "unknown error";
	}
}

static thread_local char g_create_err[256];	/* stsgpu_last_error(NULL): the failure of this thread's last stsgpu_create */
extern "C" const char *stsgpu_last_error(const stsgpu_ctx *ctx) { return ctx ? ctx->err : g_create_err; }
extern "C" int stsgpu_abi_version(void) { return STSGPU_ABI_VERSION; }

extern "C" void stsgpu_default_params(stsgpu_params *p)
{
	memset(p, 0, sizeof(*p));
	p->abi_version = STSGPU_ABI_VERSION;
	p->device = 0;
	p->n = 1048576;			/* DEFAULT_BITCOUNT, defs.h:125 */
	p->max_streams = 1024;
	p->test_mask = STSGPU_ALL_TESTS;
	p->block_frequency_M = 16384;	/* defs.h:110-115 */
	p->non_overlapping_m = 9;
	p->overlapping_m = 9;
	p->apen_m = 10;
	p->serial_m = 16;
	p->linear_complexity_M = 500;
}

/* ------------------------------------------------------------------------------------------ */
/* nonOverlappingTemplateMatchings.c:338-401: keep the m-bit words without self-overlap, ascending */
static bool template_accepted(uint32_t value, int m)
{
	uint8_t A[32];
	for (int c = 0; c < m; c++)
		A[c] = (uint8_t) ((value >> (m - 1 - c)) & 1u);
	for (int i = 1; i < m; i++) {
		bool overlap = true;
		if ((A[0] != A[m - 1]) && ((A[0] != A[m - 2]) || (A[1] != A[m - 1]))) {
			for (int c = 0; c < m - i; c++)
				if (A[c] != A[c + i]) {
					overlap = false;
					break;
				}
		}
		if (overlap)
			return false;
	}
	return true;
}

/* the X_init() self-disable rules, one line per reference check */
static uint32_t effective_mask(const stsgpu_params *p, int *universal_L)
{
	const int64_t n = p->n;
	uint32_t en = p->test_mask & STSGPU_ALL_TESTS;
	const double logn = log((double) n), log2_ = log(2.0);
	if (n < 100) en &= ~(1u << STSGPU_TEST_FREQUENCY);				/* frequency.c:102 */
	{
		const int64_t M = p->block_frequency_M, N = M > 0 ? n / M : 0;		/* blockFrequency.c:109-129 */
		if (n < 100 || M < 20 || M <= 0.01 * n || N >= 100) en &= ~(1u << STSGPU_TEST_BLOCK_FREQUENCY);
	}
	if (n < 100) en &= ~((1u << STSGPU_TEST_CUSUM) | (1u << STSGPU_TEST_RUNS));	/* cusum.c:109 runs.c:114 */
	if (n < 128) en &= ~(1u << STSGPU_TEST_LONGEST_RUN);				/* longestRunOfOnes.c:261 */
	if (n / 1024 < 38) en &= ~(1u << STSGPU_TEST_RANK);				/* rank.c:119-128 */
	if (n < 1000) en &= ~(1u << STSGPU_TEST_DFT);					/* discreteFourierTransform.c:123 */
	{
		const int64_t M = n / 8;						/* nonOverlapping...c:211-241 */
		const int m = p->non_overlapping_m;
		if (M <= 0.01 * n || m < 8 || m > 15 || m > M) en &= ~(1u << STSGPU_TEST_NON_OVERLAPPING);
	}
	{
		const int64_t N = n / 1032;						/* overlapping...c:151-168 */
		if (p->overlapping_m != 9 || n < 1000000 || N * 0.070432326346398449744 <= 5)
			en &= ~(1u << STSGPU_TEST_OVERLAPPING);
	}
	int L;										/* universal.c:145-192 */
	for (L = 7; L <= 16; L++)
		if (n < (1010LL * (1LL << L) * L))
			break;
	L--;
	*universal_L = L;
	if (n < 387840 || L < 6 || L > 16) en &= ~(1u << STSGPU_TEST_UNIVERSAL);
	if (p->apen_m >= lround(floor(logn / log2_ - 5.0))) en &= ~(1u << STSGPU_TEST_APEN);	/* approximateEntropy.c:109 */
	if (n < 1000000)								/* randomExcursions.c:112 */
		en &= ~((1u << STSGPU_TEST_RND_EXCURSION) | (1u << STSGPU_TEST_RND_EXCURSION_VAR));
	if (p->serial_m >= lround(floor(logn / log2_ - 2.0))) en &= ~(1u << STSGPU_TEST_SERIAL);	/* serial.c:112 */
	{
		const int64_t M = p->linear_complexity_M, N = M > 0 ? n / M : 0;	/* linearComplexity.c:118-138 */
		if (n < 1000000 || M < 500 || M > 5000 || N < 200) en &= ~(1u << STSGPU_TEST_LINEARCOMPLEXITY);
	}
	return en;
}

static const struct { int64_t min_n; int M, min_class, max_class; double pi[7]; } lr_rows[3] = {
	/* longestRunOfOnes.c:148-210 */
	{ 128, 8, 1, 7, { 0.21484375, 0.3671875, 0.23046875, 0.109375, 0.046875, 0.01953125, 0.01171875 } },
	{ 6272, 128, 4, 10, { 0.11740357883779323143, 0.24295595927745485921, 0.24936348317907796964,
			      0.17517706034678234949, 0.10270117130405369391, 0.05521550943390406075,
			      0.05718333762093383558 } },
	{ 750000, 10000, 10, 16, { 0.08663231117995278587, 0.20820064838760340198, 0.24841858194169954122,
				   0.19391278674165693004, 0.12145848508900441468, 0.06801108930393995064,
				   0.07336609745614297557 } },
};
static const double univ_expected[17] = { 0, 0, 0, 0, 0, 0, 5.2177052, 6.1962507, 7.1836656, 8.1764248, 9.1723243,
	10.170032, 11.168765, 12.168070, 13.167693, 14.167488, 15.167379 };	/* universal.c:72-75 */
static const double univ_variance[17] = { 0, 0, 0, 0, 0, 0, 2.954, 3.125, 3.238, 3.311, 3.356, 3.384, 3.401, 3.410,
	3.416, 3.419, 3.421 };							/* universal.c:77-80 */

/*
 * Record layout for a parameter set (must match oracle_layout_for: same contract, written independently): which
 * tests survive the X_init() checks, how many p-values and integer statistics each contributes, and the
 * non-overlapping templates.  Pure host arithmetic - no CUDA call.
 */
static void compute_layout(const stsgpu_params *p, stsgpu_layout *out, std::vector<uint32_t> *tmpl)
{
	stsgpu_layout &lay = *out;
	memset(&lay, 0, sizeof(lay));
	int L = 0;
	const uint32_t en = effective_mask(p, &L);
	lay.enabled_mask = en;
	lay.universal_L = L;
	const int64_t n = p->n;
	tmpl->clear();
	if (en & (1u << STSGPU_TEST_NON_OVERLAPPING))
		for (uint32_t v = 1; v < (1u << p->non_overlapping_m); v++)
			if (template_accepted(v, p->non_overlapping_m))
				tmpl->push_back(v);
	lay.ntemplates = (int32_t) tmpl->size();
	lay.nw = lay.ntemplates * 8;
	const int64_t bfN = p->block_frequency_M > 0 ? n / p->block_frequency_M : 0;
	const int pvc[16] = { 0, 1, 1, 2, 1, 1, 1, 1, lay.ntemplates, 1, 1, 1, 8, 18, 2, 1 };
	const int stc[16] = { 0, 1, 1 + (int) bfN, 5, 3, 7, 3, 1, 0, 6, 1, 6, 57, 19, 3, 7 };
	const int fsc[16] = { 0, 0, 1, 0, 2, 1, 1, 2, lay.ntemplates, 1, 2, 2, 8, 0, 5, 1 };
	for (int t = 1; t <= 15; t++) {
		if (!((en >> t) & 1u))
			continue;
		lay.pv_off[t] = lay.npv; lay.pv_cnt[t] = pvc[t]; lay.npv += pvc[t];
		lay.st_off[t] = lay.nst; lay.st_cnt[t] = stc[t]; lay.nst += stc[t];
		lay.fs_off[t] = lay.nfs; lay.fs_cnt[t] = fsc[t]; lay.nfs += fsc[t];
	}
}

/* the argument checks stsgpu_create makes before it touches the device */
static int check_params(const stsgpu_params *p, const char **why)
{
	if (p == NULL || p->abi_version != STSGPU_ABI_VERSION) {
		*why = "bad params pointer or ABI version";
		return STSGPU_E_INVAL;
	}
	if (p->n < 8 || (p->n % 8) != 0 || p->max_streams < 1) {	/* parse_args.c:939: n must be a multiple of 8 */
		*why = "n must be a positive multiple of 8 and max_streams >= 1";
		return STSGPU_E_INVAL;
	}
	if ((p->flags & ~STSGPU_FLAG_STATS) != 0 || p->reserved != 0) {
		*why = "unknown flag bits";
		return STSGPU_E_INVAL;
	}
	if (p->n > (1LL << 31) - 4096) {
		*why = "n above 2^31 bits per stream";
		return STSGPU_E_UNSUPPORTED;
	}
	return STSGPU_OK;
}

extern "C" int stsgpu_query_layout(const stsgpu_params *p, stsgpu_layout *out)
{
	const char *why = "";
	if (out == NULL)
		return STSGPU_E_INVAL;
	const int rc = check_params(p, &why);
	if (rc != STSGPU_OK)
		return rc;
	std::vector<uint32_t> tmpl;
	compute_layout(p, out, &tmpl);
	return out->npv == 0 ? STSGPU_E_INVAL : STSGPU_OK;
}

static int fail_create(stsgpu_ctx *ctx, int code, const char *msg)
{
	snprintf(g_create_err, sizeof(g_create_err), "%s%s%s", msg, ctx && ctx->err[0] ? ": " : "",
		 ctx && ctx->err[0] ? ctx->err : "");
	if (ctx)
		stsgpu_destroy(ctx);
	return code;
}

template <typename T> static cudaError_t upload(T **dptr, const T *host, size_t count)
{
	cudaError_t e = cudaMalloc((void **) dptr, count * sizeof(T));
	if (e != cudaSuccess)
		return e;
	return cudaMemcpy(*dptr, host, count * sizeof(T), cudaMemcpyHostToDevice);
}

static double2 unit_root(long double num, long double den)	/* exp(-2 pi i num / den) */
{
	long double a = -2.0L * 3.14159265358979323846264338327950288L * num / den;
	return make_double2((double) cosl(a), (double) sinl(a));
}

extern "C" int stsgpu_create(const stsgpu_params *p, stsgpu_ctx **out)
{
	g_create_err[0] = 0;
	{
		const char *why = "";
		const int rc = (out == NULL) ? STSGPU_E_INVAL : check_params(p, &why);
		if (rc != STSGPU_OK)
			return fail_create(NULL, rc, out == NULL ? "bad output pointer" : why);
	}
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= p->device || p->device < 0)
		return fail_create(NULL, STSGPU_E_NO_DEVICE, "no CUDA device with the requested ordinal");
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, p->device) != cudaSuccess || prop.major < 10)
		return fail_create(NULL, STSGPU_E_NO_DEVICE, "device is not sm_100 (B200) class");

	stsgpu_ctx *ctx = new (std::nothrow) stsgpu_ctx();
	if (ctx == NULL)
		return fail_create(NULL, STSGPU_E_NOMEM, "host allocation failed");
	ctx->prm = *p;
	ctx->device = p->device;
	ctx->num_sms = prop.multiProcessorCount;
	if (cudaSetDevice(p->device) != cudaSuccess)
		return fail_create(ctx, STSGPU_E_CUDA, "cudaSetDevice failed");

	/* ---- layout ---- */
	stsgpu_layout &lay = ctx->lay;
	std::vector<uint32_t> tmpl;
	compute_layout(p, &lay, &tmpl);
	const uint32_t en = lay.enabled_mask;
	const int L = lay.universal_L;
	const int64_t n = p->n;
	const int64_t bfN = p->block_frequency_M > 0 ? n / p->block_frequency_M : 0;
	if (lay.npv == 0)
		return fail_create(ctx, STSGPU_E_INVAL, "no test left enabled for these parameters");

	/* ---- device parameters ---- */
	DevParams &P = ctx->P;
	memset(&P, 0, sizeof(P));
	P.n = n;
	P.nwords = (n + 31) / 32;
	P.pitch_words = ((P.nwords + 4 + 31) / 32) * 32;
	P.npv = lay.npv; P.nst = lay.nst; P.nw = lay.nw;
	for (int t = 0; t < 16; t++) { P.pv_off[t] = lay.pv_off[t]; P.st_off[t] = lay.st_off[t]; P.fs_off[t] = lay.fs_off[t]; }
	P.nfs = lay.nfs;
	P.want_stats = (p->flags & STSGPU_FLAG_STATS) ? 1u : 0u;
	P.enabled = en;
	P.bf_M = p->block_frequency_M; P.bf_N = bfN;
	int idx = 0;								/* longestRunOfOnes.c:352-360 */
	while (idx < 3 && n > lr_rows[idx].min_n) ++idx;
	if (idx >= 3) idx = 2;
	P.lr_M = lr_rows[idx].M; P.lr_min_class = lr_rows[idx].min_class; P.lr_max_class = lr_rows[idx].max_class;
	P.lr_N = n / P.lr_M;
	P.rank_count = n / 1024;
	P.nonov_m = p->non_overlapping_m; P.ntemplates = lay.ntemplates; P.nonov_M = n / 8;
	P.ov_N = n / 1032;
	P.univ_L = L; P.univ_Q = 10LL << L; P.univ_K = 1000LL << L;		/* universal.c:324-325 */
	P.apen_m = p->apen_m; P.serial_m = p->serial_m;
	P.lc_M = p->linear_complexity_M; P.lc_N = P.lc_M > 0 ? n / P.lc_M : 0;
	{
		const double mz = 0.005 * sqrt((double) n);			/* driver.c:316 */
		P.min_zero_crossings = (int64_t) (mz > 500.0 ? mz : 500.0);
	}
	const bool want_serial = en & (1u << STSGPU_TEST_SERIAL), want_apen = en & (1u << STSGPU_TEST_APEN);
	if (want_serial || want_apen) {
		int H = 0, lowest = 99;
		if (want_serial) { H = p->serial_m; lowest = p->serial_m - 2 > 1 ? p->serial_m - 2 : 1; }
		if (want_apen) {
			if (p->apen_m + 1 > H) H = p->apen_m + 1;
			if (p->apen_m < lowest) lowest = p->apen_m;
		}
		P.hist_bits = H;
		P.hist_slice_bits = H > 15 ? H - 15 : 0;
		if (H > 21 || (want_serial && p->serial_m < 3) || (want_apen && p->apen_m < 1) || lowest < P.hist_slice_bits)
			return fail_create(ctx, STSGPU_E_UNSUPPORTED,
					   "Serial/ApEn block lengths outside the engine's range (3 <= serial m <= 21, "
					   "1 <= apen m <= 20, levels spread < 15 bits)");
	}
	if ((en & (1u << STSGPU_TEST_LINEARCOMPLEXITY)) && P.lc_M > 1022 && (P.lc_N + 3) / 4 > 65535)
		/* the warp-per-block kernel puts four blocks in a CTA and the CTAs of a stream along grid.y (k_lc.cu) */
		return fail_create(ctx, STSGPU_E_UNSUPPORTED,
				   "LinearComplexity with M > 1022 on streams of more than 262140 blocks (n > 2.6e8 bits) is not supported");
	if ((en & (1u << STSGPU_TEST_UNIVERSAL)) && L > 15)
		return fail_create(ctx, STSGPU_E_UNSUPPORTED, "Universal test with L = 16 (n >= 1.06e9 bits) is not supported");
	int logn = 0;
	while ((1LL << logn) < n) logn++;
	if (en & (1u << STSGPU_TEST_DFT)) {
		ctx->dft.fast = 0;
		ctx->dft.bs_M = 0;
		if (!dft_make_plan(n, &ctx->dft) && !dft_any_plan(n, &ctx->dft))
			return fail_create(ctx, STSGPU_E_UNSUPPORTED,
					   "DFT test: streams longer than 2^25 bits need n/2 = two 5-smooth factors of at most 4096; "
					   "clear bit 7 of test_mask for this stream length");
	}

	/* ---- tail constants (host libm, reference expressions) ---- */
	TailConsts &K = ctx->K;
	memset(&K, 0, sizeof(K));
	K.sqrt2 = sqrt(2.0); K.log2_ = log(2.0); K.sqrtn = sqrt((double) n);	/* driver.c:304-310 */
	K.two_over_sqrtn = 2.0 / K.sqrtn;					/* runs.c:139 */
	K.sqrt2n = sqrt(2.0 * (double) n);					/* runs.c:138 */
	K.lgam_bf = lgamma(bfN / 2.0);
	K.lgam_3 = lgamma(3.0); K.lgam_4 = lgamma(4.0); K.lgam_2_5 = lgamma(2.5);
	K.lgam_apen = lgamma((double) (1LL << (p->apen_m > 0 ? p->apen_m - 1 : 0)));
	K.lgam_serial1 = lgamma((double) (1LL << (p->serial_m > 1 ? p->serial_m - 1 : 0)) / 2.0);
	K.lgam_serial2 = lgamma((double) (1LL << (p->serial_m > 2 ? p->serial_m - 2 : 0)) / 2.0);
	for (int pass = 0; pass < 2; pass++) {					/* rank.c:131-170 */
		const int r = 32 - pass;
		double product = 1.0;
		for (int i = 0; i <= r - 1; i++)
			product *= ((1.0 - pow(2.0, i - 32)) * (1.0 - pow(2.0, i - 32))) / (1.0 - pow(2.0, i - r));
		const double v = pow(2.0, r * (32 + 32 - r) - 32 * 32) * product;
		if (pass == 0) K.p_32 = v; else K.p_31 = v;
	}
	K.p_30 = 1.0 - (K.p_32 + K.p_31);					/* rank.c:172 */
	for (int i = 0; i < 7; i++) K.lr_pi[i] = lr_rows[idx].pi[i];
	K.sqrtn4_095_005 = sqrt((double) n / 4.0 * 0.95 * 0.05);		/* discreteFourierTransform.c:141 */
	{
		const long M = (long) (n / 8), m = p->non_overlapping_m;	/* nonOverlapping...c:479-480 */
		if (m >= 1 && m <= 30) {
			K.nonov_mu = (M - m + 1) / ((double) ((long) 1 << m));
			const double sigma_squared =
				M * (1.0 / ((double) ((long) 1 << m)) - (2.0 * m - 1.0) / ((double) ((long) 1 << m * 2)));
			K.nonov_sqrt_sigma2 = sqrt(sigma_squared);
		}
	}
	{
		const double ov[6] = { 0.36409105321672786245, 0.18565890010624038178, 0.13938113045903269914,
				       0.10057114399877811497, 0.070432326346398449744, 0.13986544587282249192 };
		for (int i = 0; i < 6; i++) K.ov_pi[i] = ov[i];			/* overlapping...c:71-78 */
	}
	if (L >= 6 && L <= 16) {						/* universal.c:385-386 */
		const double Kd = (double) P.univ_K;
		const double c = 0.7 - 0.8 / (double) L + (4 + 32 / (double) L) * pow(Kd, -3.0 / (double) L) / 15;
		K.univ_sigma = c * sqrt(univ_variance[L] / Kd);
		K.univ_expected = univ_expected[L];
	}
	for (int i = 1; i <= 4; i++) K.rex_pi[i - 1][0] = 1.0 - 1.0 / (2.0 * i);	/* randomExcursions.c:187-207 */
	for (int j = 1; j < 5; j++)
		for (int i = 1; i <= 4; i++)
			K.rex_pi[i - 1][j] = 1.0 / (4.0 * i * i) * pow(K.rex_pi[i - 1][0], j - 1);
	for (int i = 1; i <= 4; i++) K.rex_pi[i - 1][5] = 1.0 / (2.0 * i) * pow(K.rex_pi[i - 1][0], 4);
	{
		const double lc[7] = { 0.01047, 0.03125, 0.12500, 0.50000, 0.25000, 0.06250, 0.020833 };
		for (int i = 0; i < 7; i++) K.lc_pi[i] = lc[i];			/* linearComplexity.c:62 */
	}

	/* ---- device allocations ---- */
	const int64_t B = p->max_streams;
	ctx->in_bytes = (size_t) B * P.pitch_words * 4;
#define CKC(call, what)                                                                                           \
	do {                                                                                                      \
		cudaError_t e_ = (call);                                                                          \
		if (e_ != cudaSuccess) {                                                                          \
			set_err(ctx, what, e_);                                                                   \
			return fail_create(ctx, e_ == cudaErrorMemoryAllocation ? STSGPU_E_NOMEM : STSGPU_E_CUDA, \
					   "stsgpu_create");                                                      \
		}                                                                                                 \
	} while (0)
	ctx->nslots = p->pipeline_depth > 1 ? p->pipeline_depth : 1;
	ctx->overlap_batches = !(getenv("STSGPU_SERIALIZE_BATCHES") && atoi(getenv("STSGPU_SERIALIZE_BATCHES")) > 0);
	ctx->compute_inflight = getenv("STSGPU_COMPUTE_INFLIGHT") ? atoi(getenv("STSGPU_COMPUTE_INFLIGHT")) : 2;
	if (ctx->nslots > STS_MAX_SLOTS) ctx->nslots = STS_MAX_SLOTS;
	int prio_lo = 0, prio_hi = 0;
	cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);		/* lo = least urgent, hi = most urgent */
	for (int i = 0; i < ctx->nslots; i++) {
		Slot &S = ctx->slots[i];
		CKC(cudaMalloc((void **) &S.d_in, ctx->in_bytes), "cudaMalloc(input)");
		CKC(cudaMemset(S.d_in, 0, ctx->in_bytes), "cudaMemset(input)");
		CKC(cudaMalloc((void **) &S.d_pv, (size_t) B * lay.npv * sizeof(double)), "cudaMalloc(pvals)");
		CKC(cudaMalloc((void **) &S.d_st, (size_t) B * (lay.nst + 1) * sizeof(long long)), "cudaMalloc(istats)");
		CKC(cudaMalloc((void **) &S.d_w, (size_t) B * (lay.nw + 1) * sizeof(uint32_t)), "cudaMalloc(wcount)");
		CKC(pinned_alloc_near((void **) &S.h_pv, (size_t) B * lay.npv * sizeof(double), ctx->device), "cudaMallocHost(pvals)");
		CKC(pinned_alloc_near((void **) &S.h_st, (size_t) B * (lay.nst + 1) * sizeof(long long), ctx->device),
		    "cudaMallocHost(istats)");
		CKC(pinned_alloc_near((void **) &S.h_w, (size_t) B * (lay.nw + 1) * sizeof(uint32_t), ctx->device),
		    "cudaMallocHost(wcount)");
		if (P.want_stats && lay.nfs > 0) {
			CKC(cudaMalloc((void **) &S.d_fs, (size_t) B * lay.nfs * sizeof(double)), "cudaMalloc(fstats)");
			CKC(pinned_alloc_near((void **) &S.h_fs, (size_t) B * lay.nfs * sizeof(double), ctx->device), "cudaMallocHost(fstats)");
		}
		CKC(cudaMalloc((void **) &S.d_hist_flags, (size_t) B * sizeof(uint32_t)), "cudaMalloc(hist flags)");
		{	/* Universal split over several warps per stream: 4 bytes of scratch per test block, when that stays below 2 GiB */
			const size_t per = universal_scratch_bytes_per_stream(P);
			if (per != 0 && per * (size_t) B <= ((size_t) 2 << 30))
				CKC(cudaMalloc((void **) &S.d_uni_dist, per * (size_t) B), "cudaMalloc(universal distances)");
		}
		if (want_apen)
			CKC(cudaMalloc((void **) &S.d_apen_hist, (size_t) B * ((size_t) 3 << p->apen_m) * sizeof(uint32_t)),
			    "cudaMalloc(apen histograms)");
		CKC(cudaStreamCreateWithPriority(&S.s_light, cudaStreamNonBlocking, prio_hi), "cudaStreamCreate");
		CKC(cudaStreamCreateWithPriority(&S.s_heavy, cudaStreamNonBlocking, prio_lo), "cudaStreamCreate");
		{	/* Universal was one latency chain per stream (high priority: let it start early); split over eight warps per
			 * stream it is a throughput kernel that fills every warp slot of an SM: low priority like the other big grids
			 * (measured: no difference for the pipelined batch either way; STSGPU_UNI_PRIO=1 restores the old setting) */
			const char *up = getenv("STSGPU_UNI_PRIO");
			const int uni_prio = (up && atoi(up) > 0) ? prio_hi : prio_lo;
			CKC(cudaStreamCreateWithPriority(&S.s_uni, cudaStreamNonBlocking, uni_prio), "cudaStreamCreate");
		}
		CKC(cudaStreamCreateWithPriority(&S.s_dft[0], cudaStreamNonBlocking, prio_lo), "cudaStreamCreate");
		CKC(cudaStreamCreateWithPriority(&S.s_dft[1], cudaStreamNonBlocking, prio_lo), "cudaStreamCreate");
		CKC(cudaEventCreateWithFlags(&S.ev_join_uni, cudaEventDisableTiming), "cudaEventCreate");
		CKC(cudaEventCreateWithFlags(&S.ev_join_dft, cudaEventDisableTiming), "cudaEventCreate");
		CKC(cudaEventCreateWithFlags(&S.ev_dft_fork, cudaEventDisableTiming), "cudaEventCreate");
		CKC(cudaEventCreateWithFlags(&S.ev_dft_join, cudaEventDisableTiming), "cudaEventCreate");
		CKC(cudaEventCreateWithFlags(&S.ev_fork, cudaEventDisableTiming), "cudaEventCreate");
		CKC(cudaEventCreateWithFlags(&S.ev_join, cudaEventDisableTiming), "cudaEventCreate");
		CKC(cudaEventCreateWithFlags(&S.ev_done, cudaEventDisableTiming), "cudaEventCreate");
		CKC(cudaEventCreateWithFlags(&S.ev_kdone, cudaEventDisableTiming), "cudaEventCreate");
		CKC(cudaEventCreate(&S.ev_start), "cudaEventCreate");
		CKC(cudaEventCreate(&S.ev_end), "cudaEventCreate");
		for (int k = 0; k < STS_MAX_KERNELS; k++) {
			CKC(cudaEventCreate(&S.k_ev0[k]), "cudaEventCreate");
			CKC(cudaEventCreate(&S.k_ev1[k]), "cudaEventCreate");
		}
	}
	{	/* collectives run at the highest priority: the gather's NCCL kernel must not queue behind the CTAs of the next batch */
		int pr_lo = 0, pr_hi = 0;
		cudaDeviceGetStreamPriorityRange(&pr_lo, &pr_hi);
		CKC(cudaStreamCreateWithPriority(&ctx->s_main, cudaStreamNonBlocking, pr_hi), "cudaStreamCreate");
	}
	if (!tmpl.empty())
		CKC(upload(&ctx->d_templates, tmpl.data(), tmpl.size()), "upload(templates)");
	if (en & (1u << STSGPU_TEST_LINEARCOMPLEXITY)) {
		/*
		 * class of every possible L (linearComplexity.c:356-377), evaluated on the host by the
		 * reference's expression.  `1 << M` on an int is undefined for M >= 32; x86-64 gcc takes the
		 * count mod 32, which is what every build of the reference does and what is reproduced here.
		 */
		const int M = (int) P.lc_M;
		std::vector<uint8_t> lut((size_t) M + 2);
		const double two_pow = (double) (int) (1u << (M & 31));
		const double mean = (M / 2.0) + (((M + 1) % 2) ? 10 : 8) / 36.0 - (M / 3.0 + 2.0 / 9.0) / two_pow;
		for (int Lc = 0; Lc <= M; Lc++) {
			const double T = ((M % 2) ? (mean - Lc) : (Lc - mean)) + 2.0 / 9.0;
			const double cls = (double) (6 - 1) / 2.0;
			int v;
			if (T <= -cls) v = 0;
			else if (T > cls) v = 6;
			else v = (int) ceil(T + cls);
			lut[Lc] = (uint8_t) v;
		}
		CKC(upload(&ctx->d_class_lut, lut.data(), lut.size()), "upload(LC class table)");
	}
	if (en & (1u << STSGPU_TEST_UNIVERSAL)) {
		/* universal.c:369: log(i - T[.]) / log(2), one table entry per possible distance */
		const int64_t cnt = P.univ_Q + P.univ_K + 1;
		/* stored as the exact fixed point value x * 2^58: every x is 0 or a double in [1, 32) (k_universal.cu) */
		std::vector<unsigned long long> tab((size_t) cnt);
		tab[0] = 0;
		for (int64_t d = 1; d < cnt; d++) {
			const double x = log((double) d) / K.log2_;
			const double scaled = ldexp(x, 58);
			if (!(x == 0.0 || (x >= 1.0 && x < 32.0)) || scaled != floor(scaled))
				return fail_create(ctx, STSGPU_E_UNSUPPORTED, "Universal: log2 table value outside the fixed point range");
			tab[d] = (unsigned long long) scaled;
		}
		CKC(upload(&ctx->d_log2tab, tab.data(), tab.size()), "upload(log2 table)");
	}
	if (en & (1u << STSGPU_TEST_DFT)) {
		DftPlan &D = ctx->dft;				/* N1, N2, radix lists: dft_make_plan above */
		const bool any = D.fast == 3;			/* Bluestein (k_dft_any.cu): n/2 is not 5-smooth */
		const bool ten_million = !any && n == 10000000 && !(getenv("STSGPU_DFT_GENERIC") && atoi(getenv("STSGPU_DFT_GENERIC")) > 0);
		if (ten_million) {			/* BASELINE config 4: the split k_dft_1e7.cu is written for */
			D.N1 = 2000;
			D.N2 = 2500;
		}
		const int N1 = any ? 1 : D.N1, N2 = any ? 1 : D.N2;
		const int64_t N = any ? D.bs_M : (int64_t) N1 * N2;	/* elements of one stream in a work buffer */
		if (!any)
			D.fast = (n == (1LL << 20)) ? 1 : (n == 1000000 ? 2 : 0);
		if (ten_million)
			D.fast = 4;
		D.T = sqrt(log(20.0) * (double) n);				/* discreteFourierTransform.c:142 */
		D.w1 = D.w2 = D.tl = D.th = D.p1 = D.p2 = nullptr;
		D.bs_chirp = D.bs_bhat = D.bs_tw = D.bs_wq = nullptr;
		if (any) {
			double2 *tabs[4];
			const cudaError_t e_any = dft_any_setup(n, &D, tabs, ctx->num_sms);
			for (int i = 0; i < 4; i++)
				ctx->dft_tabs[5 + i] = tabs[i];
			CKC(e_any, "DFT (Bluestein) tables");
		}
		if (!any) {
		std::vector<double2> w1(N1), w2(N2), tl(1024), th((size_t) (N / 1024 + 1)), p1(N1), p2(N2);
		{	/* per-pass twiddle tables (dft.h): entry (r - 1) Ns + k of pass p = exp(-2 pi i k r / (Ns R)) */
			for (int which = 0; which < 2; which++) {
				std::vector<double2> &w = which ? w2 : w1;
				const int np = which ? D.np2 : D.np1;
				const int *rad = which ? D.rad2 : D.rad1;
				int *off = which ? D.off2 : D.off1;
				w.clear();
				int Ns = 1;
				for (int ps = 0; ps < np; ps++) {
					const int R = rad[ps];
					off[ps] = (int) w.size();
					for (int r = 1; r < R; r++)
						for (int k = 0; k < Ns; k++)
							w.push_back(unit_root((long double) k * r, (long double) Ns * R));
					Ns *= R;
				}
				if (w.empty())
					w.push_back(make_double2(1.0, 0.0));
			}
		}
		for (int k = 0; k < 1024; k++) tl[k] = unit_root(k, (long double) N);
		for (size_t k = 0; k < th.size(); k++) th[k] = unit_root((long double) k * 1024, (long double) N);
		for (int k = 0; k < N1; k++) p1[k] = unit_root(k, (long double) n);
		for (int k = 0; k < N2; k++) p2[k] = unit_root((long double) N1 * k, (long double) n);
		double2 *d;
		CKC(upload(&d, w1.data(), w1.size()), "upload(fft w1)"); D.w1 = d; ctx->dft_tabs[0] = d;
		CKC(upload(&d, w2.data(), w2.size()), "upload(fft w2)"); D.w2 = d; ctx->dft_tabs[11] = d;
		CKC(upload(&d, tl.data(), tl.size()), "upload(fft tl)"); D.tl = d; ctx->dft_tabs[1] = d;
		CKC(upload(&d, th.data(), th.size()), "upload(fft th)"); D.th = d; ctx->dft_tabs[2] = d;
		CKC(upload(&d, p1.data(), p1.size()), "upload(fft p1)"); D.p1 = d; ctx->dft_tabs[3] = d;
		CKC(upload(&d, p2.data(), p2.size()), "upload(fft p2)"); D.p2 = d; ctx->dft_tabs[4] = d;
		D.q_twA = D.q_twB = D.q_e1 = D.q_e2 = D.q_p1 = D.q_p2 = nullptr;
		if (D.fast == 1) {						/* n = 2^20: k_dft_fast.cu (its own 256 x 2048 split) */
			std::vector<double2> tA(15 * 16), tB(7 * 256), e1(16 * 2048), e2(2048), q1(256), q2(2048);
			for (int r = 1; r < 16; r++)
				for (int k = 0; k < 16; k++) tA[(r - 1) * 16 + k] = unit_root(k * r, 256);
			for (int r = 1; r < 8; r++)
				for (int j = 0; j < 256; j++) tB[(r - 1) * 256 + j] = unit_root(j * r, 2048);
			for (int t = 0; t < 16; t++)
				for (int j2 = 0; j2 < 2048; j2++) e1[t * 2048 + j2] = unit_root((long double) j2 * t, (long double) N);
			for (int j2 = 0; j2 < 2048; j2++) e2[j2] = unit_root((long double) 16 * j2, (long double) N);
			for (int k = 0; k < 256; k++) q1[k] = unit_root(k, (long double) n);
			for (int k = 0; k < 2048; k++) q2[k] = unit_root((long double) 256 * k, (long double) n);
			CKC(upload(&d, tA.data(), tA.size()), "upload(fft twA)"); D.q_twA = d; ctx->dft_tabs[5] = d;
			CKC(upload(&d, tB.data(), tB.size()), "upload(fft twB)"); D.q_twB = d; ctx->dft_tabs[6] = d;
			CKC(upload(&d, e1.data(), e1.size()), "upload(fft e1)"); D.q_e1 = d; ctx->dft_tabs[7] = d;
			CKC(upload(&d, e2.data(), e2.size()), "upload(fft e2)"); D.q_e2 = d; ctx->dft_tabs[8] = d;
			CKC(upload(&d, q1.data(), q1.size()), "upload(fft q1)"); D.q_p1 = d; ctx->dft_tabs[9] = d;
			CKC(upload(&d, q2.data(), q2.size()), "upload(fft q2)"); D.q_p2 = d; ctx->dft_tabs[10] = d;
		}
		D.m_twA = D.m_twB = D.m_twC = D.m_twD = D.m_e1 = D.m_e2 = D.m_p1 = D.m_p2 = nullptr;
		if (D.fast == 2) {						/* n = 10^6: k_dft_1e6.cu (500 x 1000), one allocation */
			std::vector<double2> m;
			size_t o[8];
			o[0] = m.size();
			for (int r = 1; r < 10; r++) for (int k = 0; k < 5; k++) m.push_back(unit_root(k * r, 50));
			o[1] = m.size();
			for (int r = 1; r < 10; r++) for (int k = 0; k < 50; k++) m.push_back(unit_root(k * r, 500));
			o[2] = m.size();
			for (int r = 1; r < 10; r++) for (int k = 0; k < 10; k++) m.push_back(unit_root(k * r, 100));
			o[3] = m.size();
			for (int r = 1; r < 10; r++) for (int k = 0; k < 100; k++) m.push_back(unit_root(k * r, 1000));
			o[4] = m.size();
			for (int t = 0; t < 50; t++) for (int j2 = 0; j2 < 1000; j2++) m.push_back(unit_root((long double) j2 * t, (long double) N));
			o[5] = m.size();
			for (int j2 = 0; j2 < 1000; j2++) m.push_back(unit_root((long double) 50 * j2, (long double) N));
			o[6] = m.size();
			for (int k = 0; k < 500; k++) m.push_back(unit_root(k, (long double) n));
			o[7] = m.size();
			for (int k = 0; k < 1000; k++) m.push_back(unit_root((long double) 500 * k, (long double) n));
			CKC(upload(&d, m.data(), m.size()), "upload(fft 1e6 tables)"); ctx->dft_tabs[5] = d;
			D.m_twA = d + o[0]; D.m_twB = d + o[1]; D.m_twC = d + o[2]; D.m_twD = d + o[3];
			D.m_e1 = d + o[4]; D.m_e2 = d + o[5]; D.m_p1 = d + o[6]; D.m_p2 = d + o[7];
		}
		D.v_twA = D.v_twB = D.v_twC = D.v_rtwB = D.v_rtwC = D.v_e1 = D.v_e2 = D.v_p1 = D.v_p2 = nullptr;
		if (D.fast == 4) {						/* n = 10^7: k_dft_1e7.cu (2000 x 2500), one allocation */
			std::vector<double2> m;
			size_t o[9];
			o[0] = m.size();
			for (int r = 1; r < 10; r++) for (int k = 0; k < 10; k++) m.push_back(unit_root(k * r, 100));
			o[1] = m.size();
			for (int r = 1; r < 10; r++) for (int k = 0; k < 100; k++) m.push_back(unit_root(k * r, 1000));
			o[2] = m.size();
			for (int k = 0; k < 1000; k++) m.push_back(unit_root(k, 2000));
			o[3] = m.size();
			for (int r = 1; r < 5; r++) for (int k = 0; k < 100; k++) m.push_back(unit_root(k * r, 500));
			o[4] = m.size();
			for (int r = 1; r < 5; r++) for (int k = 0; k < 500; k++) m.push_back(unit_root(k * r, 2500));
			o[5] = m.size();
			for (int t = 0; t < 200; t++) for (int j2 = 0; j2 < 2500; j2++) m.push_back(unit_root((long double) j2 * t, (long double) N));
			o[6] = m.size();
			for (int j2 = 0; j2 < 2500; j2++) m.push_back(unit_root((long double) 200 * j2, (long double) N));
			o[7] = m.size();
			for (int k = 0; k < 2000; k++) m.push_back(unit_root(k, (long double) n));
			o[8] = m.size();
			for (int k = 0; k < 500; k++) m.push_back(unit_root((long double) 2000 * k, (long double) n));
			CKC(upload(&d, m.data(), m.size()), "upload(fft 1e7 tables)"); ctx->dft_tabs[5] = d;
			D.v_twA = d + o[0]; D.v_twB = d + o[1]; D.v_twC = d + o[2]; D.v_rtwB = d + o[3]; D.v_rtwC = d + o[4];
			D.v_e1 = d + o[5]; D.v_e2 = d + o[6]; D.v_p1 = d + o[7]; D.v_p2 = d + o[8];
		}
		}	/* !any */
		int64_t chunk = 128;
		const char *envc = getenv("STSGPU_DFT_CHUNK");
		if (envc && atoi(envc) > 0) chunk = atoi(envc);
		if (chunk > B) chunk = B;
		{	/* at most 1 GiB per work buffer (two per batch slot) */
			const int64_t cap = ((int64_t) 1 << 30) / (N * (int64_t) sizeof(double2));
			if (chunk > cap) chunk = cap > 0 ? cap : 1;
		}
		ctx->dft_chunk = chunk;
		for (int i = 0; i < ctx->nslots; i++)
			CKC(cudaMalloc((void **) &ctx->slots[i].d_dft_work, (size_t) 2 * chunk * N * sizeof(double2)), "cudaMalloc(fft work)");
	}
#undef CKC
	*out = ctx;
	return STSGPU_OK;
}

extern "C" void stsgpu_destroy(stsgpu_ctx *ctx)
{
	if (ctx == NULL)
		return;
	cudaSetDevice(ctx->device);
	cudaDeviceSynchronize();
	stsgpu_comm_destroy_internal(ctx);
	for (int i = 0; i < STS_MAX_SLOTS; i++) {
		Slot &S = ctx->slots[i];
		cudaFree(S.d_in); cudaFree(S.d_pv); cudaFree(S.d_st); cudaFree(S.d_w);
		cudaFreeHost(S.h_pv); cudaFreeHost(S.h_st); cudaFreeHost(S.h_w);
		cudaFree(S.d_fs); cudaFreeHost(S.h_fs);
		cudaFree(S.d_apen_hist); cudaFree(S.d_dft_work); cudaFree(S.d_hist_flags); cudaFree(S.d_uni_dist);
		cudaFree(S.d_text); cudaFree(S.d_ascii); cudaFreeHost(S.h_ascii);
		if (S.s_light) cudaStreamDestroy(S.s_light);
		if (S.s_heavy) cudaStreamDestroy(S.s_heavy);
		if (S.s_uni) cudaStreamDestroy(S.s_uni);
		if (S.s_dft[0]) cudaStreamDestroy(S.s_dft[0]);
		if (S.s_dft[1]) cudaStreamDestroy(S.s_dft[1]);
		if (S.ev_join_uni) cudaEventDestroy(S.ev_join_uni);
		if (S.ev_join_dft) cudaEventDestroy(S.ev_join_dft);
		if (S.ev_dft_fork) cudaEventDestroy(S.ev_dft_fork);
		if (S.ev_dft_join) cudaEventDestroy(S.ev_dft_join);
		if (S.ev_fork) cudaEventDestroy(S.ev_fork);
		if (S.ev_join) cudaEventDestroy(S.ev_join);
		if (S.ev_done) cudaEventDestroy(S.ev_done);
		if (S.ev_kdone) cudaEventDestroy(S.ev_kdone);
		if (S.ev_start) cudaEventDestroy(S.ev_start);
		if (S.ev_end) cudaEventDestroy(S.ev_end);
		for (int k = 0; k < STS_MAX_KERNELS; k++) {
			if (S.k_ev0[k]) cudaEventDestroy(S.k_ev0[k]);
			if (S.k_ev1[k]) cudaEventDestroy(S.k_ev1[k]);
		}
	}
	cudaFree(ctx->d_lcg_pow32);
	cudaFree(ctx->d_gen_consts);
	cudaFree(ctx->d_templates); cudaFree(ctx->d_class_lut); cudaFree(ctx->d_log2tab); cudaFree(ctx->d_debug);
	cudaFree(ctx->d_tally);
	cudaFree(ctx->d_assess_stage);
	cudaFree(ctx->d_metrics);
	for (int i = 0; i < 12; i++) cudaFree(ctx->dft_tabs[i]);
	if (ctx->s_main) cudaStreamDestroy(ctx->s_main);
	delete ctx;
}

extern "C" int stsgpu_dft_supported(int64_t n)
{
	DftPlan D;
	return (dft_make_plan(n, &D) || dft_any_plan(n, &D)) ? 1 : 0;
}

extern "C" int stsgpu_get_layout(const stsgpu_ctx *ctx, stsgpu_layout *out)
{
	if (ctx == NULL || out == NULL)
		return STSGPU_E_INVAL;
	*out = ctx->lay;
	return STSGPU_OK;
}

/* ------------------------------------------------------------------------------------------ */
struct Group {
	const char *name;
	int kind;		/* which of the slot's streams it runs on */
};

static int run_tests(stsgpu_ctx *ctx, Slot &S, int64_t nstreams, int64_t first_iteration)
{
	const DevParams &P = ctx->P;
	const stsgpu_layout &lay = ctx->lay;
	cudaStream_t sl = S.s_light;
	/*
	 * By default the kernels of consecutive batches interleave freely (measured 1 % faster on B200).
	 * STSGPU_SERIALIZE_BATCHES=1 keeps only the copy-ahead: this batch's upload (already queued on `sl`)
	 * overlaps the previous batch's kernels, but its own kernels wait for them.
	 */
	if (ctx->nslots > 1 && ctx->submit_seq > 0 && !ctx->overlap_batches) {
		Slot &prev = ctx->slots[(ctx->submit_seq - 1) % ctx->nslots];
		CK(cudaStreamWaitEvent(sl, prev.ev_kdone, 0));
	} else if (ctx->compute_inflight > 0 && ctx->nslots > ctx->compute_inflight && ctx->submit_seq >= ctx->compute_inflight) {
		/* uploads run up to pipeline_depth batches ahead, but the kernels of at most STSGPU_COMPUTE_INFLIGHT (default 2; 0 = no
		 * limit) batches interleave: a third batch of kernels only adds contention.  Measured end to end from host memory at depth 3:
		 * no limit 9.07 ms per 1024 streams, 2: 8.91, 1: 10.3 */
		Slot &prev = ctx->slots[(ctx->submit_seq - ctx->compute_inflight) % ctx->nslots];
		CK(cudaStreamWaitEvent(sl, prev.ev_kdone, 0));
	}
	CK(cudaMemsetAsync(S.d_st, 0, (size_t) nstreams * lay.nst * sizeof(long long), sl));
	CK(cudaEventRecord(S.ev_start, sl));
	S.nkernels = 0;
	const bool prof = ctx->profiling;

	LaunchCtx L;
	L.d_in = S.d_in; L.d_pv = S.d_pv; L.d_st = S.d_st; L.d_w = S.d_w; L.d_fs = S.d_fs;
	L.nstreams = nstreams; L.num_sms = ctx->num_sms;
	L.stream2 = S.s_dft[1]; L.ev2_fork = S.ev_dft_fork; L.ev2_join = S.ev_dft_join;

	if (!prof) {
		CK(cudaEventRecord(S.ev_fork, sl));
		CK(cudaStreamWaitEvent(S.s_heavy, S.ev_fork, 0));
		CK(cudaStreamWaitEvent(S.s_uni, S.ev_fork, 0));
		CK(cudaStreamWaitEvent(S.s_dft[0], S.ev_fork, 0));
	}
	/*
	 * Kernel groups.  Without profiling each group goes to the stream of its kind (engine.h): the
	 * small latency-bound kernels get SM slots as soon as they free up (high priority) instead of
	 * queueing behind the big grids, and the ALU-bound, atomics-bound and FP64-bound grids co-run.
	 */
	enum { G_LC, G_DFT, G_PATTERN, G_UNIVERSAL, G_SCAN, G_BLOCKS, G_RANK, G_NONOV, G_COUNT };
	enum { K_LIGHT, K_UNI, K_HEAVY, K_DFT };
	static const Group groups[G_COUNT] = {
		{ "linear_complexity", K_HEAVY }, { "dft", K_DFT }, { "pattern(serial,apen)", K_HEAVY },
		{ "universal", K_UNI }, { "scan(freq,cusum,runs,excursions)", K_LIGHT },
		{ "blocks(blockfreq,longestrun,overlapping)", K_LIGHT }, { "rank", K_LIGHT }, { "nonoverlapping", K_LIGHT },
	};
	cudaStream_t kind_stream[4] = { sl, S.s_uni, S.s_heavy, S.s_dft[0] };
	for (int g = 0; g < G_COUNT; g++) {
		L.stream = prof ? sl : kind_stream[groups[g].kind];
		const int slot = S.nkernels;
		CK(cudaEventRecord(S.k_ev0[slot], L.stream));
		cudaError_t e = cudaSuccess;
		int64_t launched = 0;
		switch (g) {
		case G_SCAN:
			e = launch_scan(P, L);
			launched = 1;
			if (e == cudaSuccess && P.want_stats && (P.enabled & (1u << STSGPU_TEST_RND_EXCURSION))) {
				e = launch_last_cycle(P, L);	/* -s only: stat.counter[] of the last cycle, after J is known */
				launched = 2;
			}
			break;
		case G_BLOCKS: e = launch_blocks(P, L); launched = 3; break;
		case G_RANK: e = launch_rank(P, L); launched = 1; break;
		case G_NONOV: e = launch_nonoverlapping(P, L, ctx->d_templates); launched = 1; break;
		case G_PATTERN: e = launch_pattern(P, L, S.d_apen_hist, S.d_hist_flags); launched = 2; break;
		case G_UNIVERSAL: e = launch_universal(P, L, ctx->d_log2tab, S.d_uni_dist, &launched); break;
		case G_LC: e = launch_lc(P, L, ctx->d_class_lut); launched = 1; break;
		case G_DFT:
			e = launch_dft(P, L, ctx->dft, S.d_dft_work, ctx->dft_chunk, &launched);
			break;
		}
		if (e != cudaSuccess) {
			set_err(ctx, groups[g].name, e);
			return STSGPU_E_CUDA;
		}
		ctx->launches += launched;
		CK(cudaEventRecord(S.k_ev1[slot], L.stream));
		S.k_name[slot] = groups[g].name;
		S.nkernels++;
	}
	if (!prof) {
		CK(cudaEventRecord(S.ev_join, S.s_heavy));
		CK(cudaStreamWaitEvent(sl, S.ev_join, 0));
		CK(cudaEventRecord(S.ev_join_uni, S.s_uni));
		CK(cudaStreamWaitEvent(sl, S.ev_join_uni, 0));
		CK(cudaEventRecord(S.ev_join_dft, S.s_dft[0]));
		CK(cudaStreamWaitEvent(sl, S.ev_join_dft, 0));
	}
	L.stream = sl;
	{
		const int slot = S.nkernels;
		CK(cudaEventRecord(S.k_ev0[slot], sl));
		/* three independent kernels: p-value slots here, CumulativeSums and ApEn on two idle streams */
		cudaStream_t s_cusum = prof ? sl : S.s_uni, s_apen = prof ? sl : S.s_heavy;
		if (!prof) {
			CK(cudaEventRecord(S.ev_fork, sl));
			CK(cudaStreamWaitEvent(s_cusum, S.ev_fork, 0));
			CK(cudaStreamWaitEvent(s_apen, S.ev_fork, 0));
		}
		cudaError_t e = launch_tails(P, ctx->K, L, S.d_apen_hist, s_cusum, s_apen);
		if (e != cudaSuccess) {
			set_err(ctx, "tails", e);
			return STSGPU_E_CUDA;
		}
		if (!prof) {
			CK(cudaEventRecord(S.ev_join_uni, s_cusum));
			CK(cudaStreamWaitEvent(sl, S.ev_join_uni, 0));
			CK(cudaEventRecord(S.ev_join, s_apen));
			CK(cudaStreamWaitEvent(sl, S.ev_join, 0));
		}
		ctx->launches += 3;
		CK(cudaEventRecord(S.k_ev1[slot], sl));
		S.k_name[slot] = "tails(p-values)";
		S.nkernels++;
	}
	if (ctx->assessing) {		/* add this batch's p-values to the assessment tallies */
		cudaError_t e = launch_assess_tally(P, L, ctx->assess_bins, ctx->assess_alpha, ctx->d_tally, S.ascii ? S.d_ascii : nullptr);
		if (e != cudaSuccess) {
			set_err(ctx, "assess tally", e);
			return STSGPU_E_CUDA;
		}
		ctx->launches += 1;
	}
	CK(cudaEventRecord(S.ev_end, sl));
	CK(cudaEventRecord(S.ev_kdone, sl));
	CK(cudaMemcpyAsync(S.h_pv, S.d_pv, (size_t) nstreams * lay.npv * sizeof(double), cudaMemcpyDeviceToHost, sl));
	CK(cudaMemcpyAsync(S.h_st, S.d_st, (size_t) nstreams * lay.nst * sizeof(long long), cudaMemcpyDeviceToHost, sl));
	if (lay.nw)
		CK(cudaMemcpyAsync(S.h_w, S.d_w, (size_t) nstreams * lay.nw * sizeof(uint32_t), cudaMemcpyDeviceToHost, sl));
	if (S.d_fs)
		CK(cudaMemcpyAsync(S.h_fs, S.d_fs, (size_t) nstreams * lay.nfs * sizeof(double), cudaMemcpyDeviceToHost, sl));
	if (S.ascii)
		CK(cudaMemcpyAsync(S.h_ascii, S.d_ascii, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, sl));
	S.nstreams = nstreams;
	S.first = first_iteration;
	S.gathered = false;
	if (ctx->gather_root >= 0) {
		const int rc = stsgpu_comm_enqueue_gather(ctx, S, nstreams, sl);
		if (rc != STSGPU_OK)
			return rc;
	}
	CK(cudaEventRecord(S.ev_done, sl));
	S.nstreams = nstreams;
	S.first = first_iteration;
	S.in_flight = true;
	ctx->submit_seq++;
	return STSGPU_OK;
}

/* the slot the next submit / upload / generate works on; fails while that slot is still in flight */
static int next_slot(stsgpu_ctx *ctx, int64_t nstreams, Slot **out)
{
	if (ctx == NULL)
		return STSGPU_E_INVAL;
	if (nstreams < 1 || nstreams > ctx->prm.max_streams) {
		snprintf(ctx->err, sizeof(ctx->err), "nstreams %lld outside 1..max_streams", (long long) nstreams);
		return STSGPU_E_RANGE;
	}
	Slot &S = ctx->slots[ctx->submit_seq % ctx->nslots];
	if (ctx->assessing && ctx->tally_reduced) {
		snprintf(ctx->err, sizeof(ctx->err), "the tallies hold the sum over all ranks (stsgpu_assess_reduce): call stsgpu_assess_begin before the next batch");
		return STSGPU_E_STATE;
	}
	if (S.in_flight) {
		snprintf(ctx->err, sizeof(ctx->err), "%d batches are in flight (pipeline_depth): call stsgpu_wait first",
			 ctx->nslots);
		return STSGPU_E_STATE;
	}
	if (cudaSetDevice(ctx->device) != cudaSuccess)
		return STSGPU_E_CUDA;
	*out = &S;
	return STSGPU_OK;
}

extern "C" int stsgpu_upload(stsgpu_ctx *ctx, const uint8_t *raw, int64_t nstreams)
{
	Slot *S;
	int rc = next_slot(ctx, nstreams, &S);
	if (rc) return rc;
	if (raw == NULL) return STSGPU_E_INVAL;
	const size_t row = (size_t) (ctx->P.n / 8);
	CK(cudaMemcpy2DAsync(S->d_in, (size_t) ctx->P.pitch_words * 4, raw, row, row, (size_t) nstreams,
			     cudaMemcpyHostToDevice, S->s_light));
	return STSGPU_OK;
}

extern "C" int stsgpu_submit(stsgpu_ctx *ctx, const uint8_t *raw, int64_t nstreams, int64_t first_iteration)
{
	int rc = stsgpu_upload(ctx, raw, nstreams);
	if (rc) return rc;
	Slot &S = ctx->slots[ctx->submit_seq % ctx->nslots];
	S.ascii = false;
	return run_tests(ctx, S, nstreams, first_iteration);
}

extern "C" int stsgpu_submit_resident(stsgpu_ctx *ctx, int64_t nstreams, int64_t first_iteration)
{
	Slot *S;
	int rc = next_slot(ctx, nstreams, &S);
	if (rc) return rc;
	S->ascii = false;
	return run_tests(ctx, *S, nstreams, first_iteration);
}

/*
 * "-F a": stream k of the batch is the first n digits at or after text[k n] (the reference's seek rule,
 * utilities.c:1682-1683), packed on the device by k_ascii.cu.  The staging buffer is allocated on first use.
 */
extern "C" int stsgpu_submit_ascii(stsgpu_ctx *ctx, const char *text, int64_t text_len, int64_t nstreams, int64_t first_iteration)
{
	Slot *Sp;
	int rc = next_slot(ctx, nstreams, &Sp);
	if (rc) return rc;
	Slot &S = *Sp;
	if (text == NULL || text_len < 0) return STSGPU_E_INVAL;
	const int64_t n = ctx->P.n;
	const int64_t cap = (ctx->prm.max_streams + 1) * n;		/* one stream of slack for white space */
	if (text_len > cap)
		text_len = cap;
	if (S.d_text == nullptr || S.d_ascii == nullptr || S.h_ascii == nullptr) {
		if (cudaMalloc((void **) &S.d_text, (size_t) cap + 16) != cudaSuccess ||
		    cudaMalloc((void **) &S.d_ascii, 2 * sizeof(unsigned long long)) != cudaSuccess ||
		    cudaMallocHost((void **) &S.h_ascii, 2 * sizeof(unsigned long long)) != cudaSuccess) {
			/* all three or none: a later call must not find a half-made staging area */
			cudaFree(S.d_text); cudaFree(S.d_ascii); cudaFreeHost(S.h_ascii);
			S.d_text = nullptr; S.d_ascii = nullptr; S.h_ascii = nullptr;
			cudaGetLastError();
			snprintf(ctx->err, sizeof(ctx->err), "cannot allocate the %lld-byte ASCII staging buffer", (long long) cap);
			return STSGPU_E_NOMEM;
		}
		S.text_cap = (size_t) cap;
	}
	CK(cudaMemcpyAsync(S.d_text, text, (size_t) text_len, cudaMemcpyHostToDevice, S.s_light));
	CK(cudaMemsetAsync(S.d_ascii, 0xFF, 2 * sizeof(unsigned long long), S.s_light));
	cudaError_t e = launch_ascii_pack(ctx->P, S.d_text, text_len, S.d_in, S.d_ascii, nstreams, S.s_light);
	if (e != cudaSuccess) {
		set_err(ctx, "ascii pack", e);
		return STSGPU_E_CUDA;
	}
	ctx->launches += 1;
	S.ascii = true;
	return run_tests(ctx, S, nstreams, first_iteration);
}

extern "C" int stsgpu_ascii_status(stsgpu_ctx *ctx, int64_t *complete_streams, int64_t *bad_offset)
{
	if (ctx == NULL)
		return STSGPU_E_INVAL;
	if (!ctx->have_results) {
		snprintf(ctx->err, sizeof(ctx->err), "no completed batch");
		return STSGPU_E_STATE;
	}
	const Slot &S = ctx->slots[ctx->last_done];
	if (!S.ascii) {
		snprintf(ctx->err, sizeof(ctx->err), "the last completed batch was not submitted with stsgpu_submit_ascii");
		return STSGPU_E_STATE;
	}
	if (complete_streams) *complete_streams = S.nstreams;
	if (bad_offset) *bad_offset = (S.h_ascii[1] == ~0ull) ? -1 : (int64_t) S.h_ascii[1];
	return STSGPU_OK;
}

extern "C" int stsgpu_wait(stsgpu_ctx *ctx)
{
	if (ctx == NULL)
		return STSGPU_E_INVAL;
	const int idx = (int) (ctx->wait_seq % ctx->nslots);
	Slot &S = ctx->slots[idx];
	if (!S.in_flight || ctx->wait_seq >= ctx->submit_seq) {
		snprintf(ctx->err, sizeof(ctx->err), "stsgpu_wait without a submitted batch");
		return STSGPU_E_STATE;
	}
	cudaSetDevice(ctx->device);
	CK(cudaEventSynchronize(S.ev_done));	/* a failed wait leaves the batch pending: the pipeline state moves only on success */
	CK(cudaGetLastError());
	S.in_flight = false;
	ctx->wait_seq++;
	if (S.ascii && S.h_ascii[0] < (unsigned long long) S.nstreams)
		S.nstreams = (int64_t) S.h_ascii[0];		/* the text ended inside stream h_ascii[0] */
	ctx->last_done = idx;
	ctx->d_pv = S.d_pv;
	ctx->h_pv = S.h_pv;
	ctx->cur_nstreams = S.nstreams;
	ctx->have_results = true;
	return STSGPU_OK;
}

extern "C" int stsgpu_pending(const stsgpu_ctx *ctx)
{
	return ctx ? (int) (ctx->submit_seq - ctx->wait_seq) : 0;
}

extern "C" int stsgpu_results(stsgpu_ctx *ctx, int64_t *nstreams, int64_t *first_iteration, const double **pvals,
			      const int64_t **istats, const uint32_t **wcount)
{
	if (ctx == NULL)
		return STSGPU_E_INVAL;
	if (!ctx->have_results) {
		snprintf(ctx->err, sizeof(ctx->err), "no completed batch");
		return STSGPU_E_STATE;
	}
	const Slot &S = ctx->slots[ctx->last_done];
	if (nstreams) *nstreams = S.nstreams;
	if (first_iteration) *first_iteration = S.first;
	if (pvals) *pvals = S.h_pv;
	if (istats) *istats = (const int64_t *) S.h_st;
	if (wcount) *wcount = S.h_w;
	return STSGPU_OK;
}

extern "C" int stsgpu_results_stats(stsgpu_ctx *ctx, const double **fstats)
{
	if (ctx == NULL || fstats == NULL)
		return STSGPU_E_INVAL;
	if (!ctx->have_results) {
		snprintf(ctx->err, sizeof(ctx->err), "no completed batch");
		return STSGPU_E_STATE;
	}
	*fstats = ctx->slots[ctx->last_done].h_fs;	/* NULL without STSGPU_FLAG_STATS */
	return STSGPU_OK;
}

extern "C" int stsgpu_generate_lcg(stsgpu_ctx *ctx, int64_t first_stream, int64_t nstreams)
{
	Slot *S;
	int rc = next_slot(ctx, nstreams, &S);
	if (rc) return rc;
	if (first_stream < 0) return STSGPU_E_INVAL;
	if (ctx->d_lcg_pow32 == nullptr) {
		uint32_t tab[256];
		lcg_pow32_table(tab);
		CK(cudaMalloc((void **) &ctx->d_lcg_pow32, sizeof(tab)));
		CK(cudaMemcpy(ctx->d_lcg_pow32, tab, sizeof(tab), cudaMemcpyHostToDevice));
	}
	cudaError_t e = launch_lcg(ctx->P, S->d_in, first_stream, nstreams, ctx->d_lcg_pow32, S->s_light);
	if (e != cudaSuccess) {
		set_err(ctx, "lcg", e);
		return STSGPU_E_CUDA;
	}
	ctx->launches += 1;
	return STSGPU_OK;
}

extern "C" int stsgpu_generate(stsgpu_ctx *ctx, int generator, int64_t first_stream, int64_t nstreams)
{
	if (ctx == NULL)
		return STSGPU_E_INVAL;
	if (generator == 1)
		return stsgpu_generate_lcg(ctx, first_stream, nstreams);
	if (generator == 3 || (generator >= 6 && generator <= 9)) {
		snprintf(ctx->err, sizeof(ctx->err), "generator %d is one chain for the whole run: made on the host (host/sts_generators.h)",
			 generator);
		return STSGPU_E_UNSUPPORTED;
	}
	if (generator != 2 && generator != 4 && generator != 5)
		return STSGPU_E_INVAL;
	if (ctx->P.n < 1024) {
		snprintf(ctx->err, sizeof(ctx->err), "generator %d needs n >= 1024", generator);
		return STSGPU_E_UNSUPPORTED;
	}
	Slot *S;
	int rc = next_slot(ctx, nstreams, &S);
	if (rc) return rc;
	if (first_stream < 0) return STSGPU_E_INVAL;
	if (ctx->d_gen_consts == nullptr) {
		void *h = malloc(gen_consts_bytes());
		if (h == NULL) return STSGPU_E_NOMEM;
		gen_consts_host(h);
		cudaError_t e = cudaMalloc(&ctx->d_gen_consts, gen_consts_bytes());
		if (e == cudaSuccess)
			e = cudaMemcpy(ctx->d_gen_consts, h, gen_consts_bytes(), cudaMemcpyHostToDevice);
		free(h);
		if (e != cudaSuccess) {
			cudaFree(ctx->d_gen_consts);
			ctx->d_gen_consts = nullptr;
			set_err(ctx, "generator tables", e);
			return STSGPU_E_CUDA;
		}
	}
	cudaError_t e = launch_generator(generator, ctx->P, S->d_in, first_stream, nstreams, ctx->d_gen_consts, S->s_light);
	if (e != cudaSuccess) {
		set_err(ctx, "generator", e);
		return STSGPU_E_CUDA;
	}
	ctx->launches += 1;
	return STSGPU_OK;
}

extern "C" int stsgpu_download_input(stsgpu_ctx *ctx, uint8_t *raw, int64_t nstreams)
{
	Slot *S;
	int rc = next_slot(ctx, nstreams, &S);
	if (rc) return rc;
	if (raw == NULL) return STSGPU_E_INVAL;
	const size_t row = (size_t) (ctx->P.n / 8);
	CK(cudaMemcpy2DAsync(raw, row, S->d_in, (size_t) ctx->P.pitch_words * 4, row, (size_t) nstreams,
			     cudaMemcpyDeviceToHost, S->s_light));
	CK(cudaStreamSynchronize(S->s_light));
	return STSGPU_OK;
}

extern "C" void *stsgpu_host_alloc(size_t bytes)
{
	void *p = NULL;
	int device = 0;
	if (cudaGetDevice(&device) != cudaSuccess || pinned_alloc_near(&p, bytes, device) != cudaSuccess)
		return NULL;
	return p;
}
extern "C" void stsgpu_host_free(void *p)
{
	if (p)
		cudaFreeHost(p);
}

/* ------------------------------------------------------------------------------------------ */
extern "C" int stsgpu_assess_begin(stsgpu_ctx *ctx, int64_t bins, double alpha, double level)
{
	if (ctx == NULL || bins < 1 || bins > (1 << 20))
		return STSGPU_E_INVAL;
	if (ctx->submit_seq != ctx->wait_seq) {
		snprintf(ctx->err, sizeof(ctx->err), "stsgpu_assess_begin needs an idle engine");
		return STSGPU_E_STATE;
	}
	if (cudaSetDevice(ctx->device) != cudaSuccess)
		return STSGPU_E_CUDA;
	const size_t words = (size_t) ctx->lay.npv * (size_t) (bins + 2);
	if (ctx->d_tally) cudaFree(ctx->d_tally);
	ctx->d_tally = nullptr;
	CK(cudaMalloc((void **) &ctx->d_tally, words * sizeof(unsigned long long)));
	CK(cudaMemset(ctx->d_tally, 0, words * sizeof(unsigned long long)));
	ctx->assess_bins = bins;
	ctx->assess_alpha = alpha;
	ctx->assess_level = level;
	ctx->assessing = true;
	ctx->tally_reduced = false;
	return STSGPU_OK;
}

#define ASSESS_STAGE_VALUES ((int64_t) 1 << 22)		/* 32 MiB of p-values per copy */

extern "C" int stsgpu_assess_add(stsgpu_ctx *ctx, int test, const double *pvals, int64_t count, int64_t first_index)
{
	if (ctx == NULL || pvals == NULL || count < 0 || first_index < 0 || test < 1 || test > 15)
		return STSGPU_E_INVAL;
	if (!ctx->assessing || ctx->submit_seq != ctx->wait_seq) {
		snprintf(ctx->err, sizeof(ctx->err), "stsgpu_assess_add needs stsgpu_assess_begin and an idle engine");
		return STSGPU_E_STATE;
	}
	if (ctx->tally_reduced) {
		snprintf(ctx->err, sizeof(ctx->err), "the tallies hold the sum over all ranks (stsgpu_assess_reduce): call stsgpu_assess_begin before adding more");
		return STSGPU_E_STATE;
	}
	if (!((ctx->lay.enabled_mask >> test) & 1u)) {
		snprintf(ctx->err, sizeof(ctx->err), "test %d is not enabled in this engine", test);
		return STSGPU_E_RANGE;
	}
	if (cudaSetDevice(ctx->device) != cudaSuccess)
		return STSGPU_E_CUDA;
	if (ctx->d_assess_stage == nullptr)
		CK(cudaMalloc((void **) &ctx->d_assess_stage, (size_t) ASSESS_STAGE_VALUES * sizeof(double)));
	for (int64_t done = 0; done < count;) {
		const int64_t c = (count - done < ASSESS_STAGE_VALUES) ? count - done : ASSESS_STAGE_VALUES;
		CK(cudaMemcpyAsync(ctx->d_assess_stage, pvals + done, (size_t) c * sizeof(double), cudaMemcpyHostToDevice, ctx->s_main));
		cudaError_t e = launch_assess_add(ctx->P, ctx->lay.pv_off[test], ctx->lay.pv_cnt[test], ctx->d_assess_stage, c, first_index + done, ctx->assess_bins,
						  ctx->assess_alpha, ctx->d_tally, ctx->s_main);
		if (e != cudaSuccess) {
			set_err(ctx, "assess add", e);
			return STSGPU_E_CUDA;
		}
		CK(cudaStreamSynchronize(ctx->s_main));		/* `pvals` may be pageable and is reused by the caller */
		ctx->launches += 1;
		done += c;
	}
	return STSGPU_OK;
}

extern "C" int stsgpu_assess_finish(stsgpu_ctx *ctx, stsgpu_metric *metrics, int64_t *freq)
{
	if (ctx == NULL || metrics == NULL)
		return STSGPU_E_INVAL;
	if (!ctx->assessing || ctx->submit_seq != ctx->wait_seq) {
		snprintf(ctx->err, sizeof(ctx->err), "stsgpu_assess_finish needs stsgpu_assess_begin and an idle engine");
		return STSGPU_E_STATE;
	}
	if (cudaSetDevice(ctx->device) != cudaSuccess)
		return STSGPU_E_CUDA;
	const int npv = ctx->lay.npv;
	const int64_t bins = ctx->assess_bins;
	if (ctx->d_metrics == nullptr)
		CK(cudaMalloc((void **) &ctx->d_metrics, (size_t) npv * sizeof(stsgpu_metric)));
	stsgpu_metric *d_out = ctx->d_metrics;
	/* lgamma of the per-run constant (bins - 1) / 2 with glibc, like every other tail constant */
	cudaError_t e = launch_assess_finish(ctx->P, bins, ctx->assess_alpha, ctx->assess_level, lgamma((bins - 1.0) / 2.0),
					     ctx->d_tally, d_out, ctx->s_main);
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(metrics, d_out, (size_t) npv * sizeof(stsgpu_metric), cudaMemcpyDeviceToHost, ctx->s_main);
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(ctx->s_main);
	if (e != cudaSuccess) {
		set_err(ctx, "assess finish", e);
		return STSGPU_E_CUDA;
	}
	ctx->launches += 1;
	if (freq) {
		std::vector<unsigned long long> t((size_t) npv * (size_t) (bins + 2));
		CK(cudaMemcpy(t.data(), ctx->d_tally, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
		for (int q = 0; q < npv; q++)
			for (int64_t b = 0; b < bins; b++)
				freq[(int64_t) q * bins + b] = (int64_t) t[(size_t) q * (size_t) (bins + 2) + 2 + (size_t) b];
	}
	return STSGPU_OK;
}

/* ------------------------------------------------------------------------------------------ */
extern "C" int stsgpu_set_profiling(stsgpu_ctx *ctx, int on)
{
	if (ctx == NULL) return STSGPU_E_INVAL;
	ctx->profiling = on != 0;
	return STSGPU_OK;
}
extern "C" int stsgpu_kernel_count(const stsgpu_ctx *ctx)
{
	return (ctx && ctx->last_done >= 0) ? ctx->slots[ctx->last_done].nkernels : 0;
}
extern "C" int stsgpu_kernel_info(stsgpu_ctx *ctx, int i, const char **name, float *ms)
{
	if (ctx == NULL || ctx->last_done < 0) return STSGPU_E_RANGE;
	Slot &S = ctx->slots[ctx->last_done];
	if (i < 0 || i >= S.nkernels) return STSGPU_E_RANGE;
	if (name) *name = S.k_name[i];
	if (ms) {
		*ms = 0.f;
		CK(cudaEventElapsedTime(ms, S.k_ev0[i], S.k_ev1[i]));
	}
	return STSGPU_OK;
}
extern "C" int stsgpu_last_batch_ms(stsgpu_ctx *ctx, float *ms)
{
	if (ctx == NULL || ms == NULL) return STSGPU_E_INVAL;
	if (!ctx->have_results) return STSGPU_E_STATE;
	Slot &S = ctx->slots[ctx->last_done];
	CK(cudaEventElapsedTime(ms, S.ev_start, S.ev_end));
	return STSGPU_OK;
}
extern "C" int64_t stsgpu_launch_count(const stsgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int stsgpu_debug_pattern_histogram(stsgpu_ctx *ctx, int64_t s, int bits, uint32_t *hist)
{
	if (ctx == NULL) return STSGPU_E_INVAL;
	if (ctx->submit_seq != ctx->wait_seq) {
		snprintf(ctx->err, sizeof(ctx->err), "debug histogram needs an idle engine");
		return STSGPU_E_STATE;
	}
	if (hist == NULL || bits < 1 || bits > 21 || s < 0 || s >= ctx->prm.max_streams) return STSGPU_E_INVAL;
	if (cudaSetDevice(ctx->device) != cudaSuccess) return STSGPU_E_CUDA;
	/* input of the last completed batch, or of the slot about to be submitted when nothing ran yet */
	Slot &S = ctx->slots[ctx->last_done >= 0 ? ctx->last_done : (int) (ctx->submit_seq % ctx->nslots)];
	CK(cudaStreamSynchronize(S.s_light));
	if (ctx->d_debug == NULL)
		CK(cudaMalloc((void **) &ctx->d_debug, sizeof(uint32_t) << 21));
	LaunchCtx L;
	L.stream2 = nullptr; L.ev2_fork = L.ev2_join = nullptr;
	L.stream = ctx->s_main; L.d_in = S.d_in; L.d_pv = S.d_pv; L.d_st = S.d_st; L.d_w = S.d_w; L.d_fs = nullptr;
	L.nstreams = 1; L.num_sms = ctx->num_sms;
	cudaError_t e = launch_debug_histogram(ctx->P, L, s, bits, ctx->d_debug);
	if (e != cudaSuccess) {
		set_err(ctx, "debug histogram", e);
		return STSGPU_E_CUDA;
	}
	CK(cudaMemcpyAsync(hist, ctx->d_debug, sizeof(uint32_t) << bits, cudaMemcpyDeviceToHost, ctx->s_main));
	CK(cudaStreamSynchronize(ctx->s_main));
	return STSGPU_OK;
}
