/* engine.h - the context behind the opaque stsgpu_ctx of include/stsgpu.h */
#pragma once
#include <vector>
#include "common.cuh"
#include "tails.h"
#include "dft.h"

#define STS_MAX_SLOTS 4

/*
 * One batch in flight.  A context owns `nslots` of these and uses them round robin, so that the
 * host-to-device copy, the kernels and the record copy of consecutive batches overlap.  Each slot
 * has five streams, forked and joined by events: `s_light` (high priority: copies, the small scan /
 * block / rank / template kernels, the tails), `s_uni` (high priority: Universal, one long latency
 * chain per stream), `s_heavy` (low priority: linear complexity = integer ALU, pattern histograms =
 * shared-memory atomics) and `s_dft[2]` (low priority: the DFT's FP64 + LSU kernels, chunks
 * alternating).  Work with different bottlenecks therefore shares the SMs.
 */
struct Slot {
	uint32_t *d_in = nullptr;
	double *d_pv = nullptr;
	long long *d_st = nullptr;
	uint32_t *d_w = nullptr;
	double *h_pv = nullptr;
	long long *h_st = nullptr;
	uint32_t *h_w = nullptr;
	double *d_fs = nullptr, *h_fs = nullptr;	/* fstats (STSGPU_FLAG_STATS) */
	/* gather mode (comm.cu): this batch's records of every rank, on the root; {nstreams} of this rank's batch */
	double *d_gather = nullptr, *h_gather = nullptr;
	double *d_hdr = nullptr, *h_hdr = nullptr, *d_hdr_all = nullptr, *h_hdr_all = nullptr;
	bool gathered = false;
	uint32_t *d_apen_hist = nullptr;
	uint32_t *d_uni_dist = nullptr;		/* Universal: the distance of every test block (k_universal.cu, split kernels) */
	uint32_t *d_hist_flags = nullptr;	/* streams the packed 16-bit histogram could not count exactly */
	double2 *d_dft_work = nullptr;
	/* "-F a" batches (k_ascii.cu): text staging, {first incomplete stream, first bad character} */
	unsigned char *d_text = nullptr;
	size_t text_cap = 0;
	unsigned long long *d_ascii = nullptr, *h_ascii = nullptr;
	bool ascii = false;
	cudaStream_t s_light = nullptr, s_uni = nullptr, s_heavy = nullptr, s_dft[2] = { nullptr, nullptr };
	cudaEvent_t ev_join_uni = nullptr, ev_join_dft = nullptr;
	cudaEvent_t ev_dft_fork = nullptr, ev_dft_join = nullptr;
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_start = nullptr, ev_end = nullptr, ev_done = nullptr, ev_kdone = nullptr;
	int nkernels = 0;
	const char *k_name[STS_MAX_KERNELS] = { nullptr };
	cudaEvent_t k_ev0[STS_MAX_KERNELS] = { nullptr }, k_ev1[STS_MAX_KERNELS] = { nullptr };
	int64_t nstreams = 0, first = 0;
	bool in_flight = false;
};

struct stsgpu_ctx {
	stsgpu_params prm;
	stsgpu_layout lay;
	DevParams P;
	TailConsts K;
	DftPlan dft;
	int device = 0, num_sms = 0;

	int nslots = 2;
	int compute_inflight = 2;	/* STSGPU_COMPUTE_INFLIGHT: > 0 = kernels of at most that many batches interleave (uploads still run ahead) */
	bool overlap_batches = true;	/* false (STSGPU_SERIALIZE_BATCHES=1): kernels of consecutive batches do not interleave */
	Slot slots[STS_MAX_SLOTS];
	uint64_t submit_seq = 0;	/* the next submit / upload / generate uses slot submit_seq % nslots */
	uint64_t wait_seq = 0;		/* the next wait completes slot wait_seq % nslots */
	int last_done = -1;		/* slot whose records stsgpu_results() returns */

	cudaStream_t s_main = nullptr;	/* collectives and debug helpers */
	size_t in_bytes = 0;

	/* views of the last completed slot (comm.cu gathers from these) */
	double *d_pv = nullptr, *h_pv = nullptr;
	int64_t cur_nstreams = 0;
	bool have_results = false;

	uint32_t *d_templates = nullptr;
	uint8_t *d_class_lut = nullptr;
	unsigned long long *d_log2tab = nullptr;	/* log(d) / log(2) as exact fixed point x * 2^58 */
	double2 *dft_tabs[12] = { nullptr };
	int64_t dft_chunk = 0;
	uint32_t *d_debug = nullptr;
	uint32_t *d_lcg_pow32 = nullptr;	/* k_lcg.cu: a^(32 j) mod p, made on first use */
	void *d_gen_consts = nullptr;		/* k_gen.cu: moduli, seeds and tables of generators 2, 4 and 5, made on first use */

	/* assessment tallies (stsgpu_assess_*): per column [sample, toolow, freq[bins]] */
	bool assessing = false;
	int64_t assess_bins = 0;
	double assess_alpha = 0.01, assess_level = 0.0001;
	unsigned long long *d_tally = nullptr;
	stsgpu_metric *d_metrics = nullptr;	/* stsgpu_assess_finish: verdicts, one per p-value column */
	double *d_assess_stage = nullptr;	/* stsgpu_assess_add: staging for p-values read from files */

	bool profiling = false;
	int64_t launches = 0;

	/* NCCL (comm.cu) */
	void *nccl_comm = nullptr;
	int rank = 0, nranks = 1;
	double *d_gather = nullptr;	/* stsgpu_comm_barrier's word */
	int gather_root = -1;		/* >= 0: gather mode, every batch ends with the exchange (stsgpu_gather_begin) */
	bool tally_reduced = false;	/* the tallies already hold the sum over ranks (stsgpu_assess_reduce) */

	char err[256] = { 0 };
};

void stsgpu_comm_destroy_internal(stsgpu_ctx *ctx);
/* gather mode: queue this batch's exchange on `stream` after its tails (comm.cu); no-op outside gather mode */
int stsgpu_comm_enqueue_gather(stsgpu_ctx *ctx, Slot &S, int64_t nstreams, cudaStream_t stream);

cudaError_t launch_tails(const DevParams &P, const TailConsts &K, const LaunchCtx &L, const uint32_t *d_apen_hist,
			 cudaStream_t s_cusum, cudaStream_t s_apen);
cudaError_t launch_dft(const DevParams &P, const LaunchCtx &L, const DftPlan &D, double2 *d_work, int64_t work_streams, int64_t *launches);
cudaError_t launch_assess_tally(const DevParams &P, const LaunchCtx &L, int64_t bins, double alpha, unsigned long long *d_tally,
				const unsigned long long *d_limit);
cudaError_t launch_assess_add(const DevParams &P, int pv_off, int pv_cnt, const double *d_vals, int64_t count, int64_t first_index,
			      int64_t bins, double alpha, unsigned long long *d_tally, cudaStream_t stream);
cudaError_t launch_assess_finish(const DevParams &P, int64_t bins, double alpha, double level, double lgam, const unsigned long long *d_tally,
				 stsgpu_metric *d_out, cudaStream_t stream);
cudaError_t launch_lcg(const DevParams &P, uint32_t *d_in, int64_t first_stream, int64_t nstreams, const uint32_t *d_pow32,
		       cudaStream_t stream);
void lcg_pow32_table(uint32_t tab[256]);
/* k_gen.cu: generators 2, 4, 5 */
size_t gen_consts_bytes(void);
void gen_consts_host(void *dst);
cudaError_t launch_generator(int generator, const DevParams &P, uint32_t *d_in, int64_t first_stream, int64_t nstreams,
			     const void *d_consts, cudaStream_t stream);
cudaError_t launch_debug_histogram(const DevParams &P, const LaunchCtx &L, int64_t stream_index, int bits, uint32_t *d_out);
