/*
 * comm.cu - multi-GPU gather of the per-stream p-value records to a root rank over NCCL
 * (NVLink 5 / NVSwitch inside one B200 box).
 *
 * The path shards by independent bitstreams exactly like the reference's `-j jobnum` mode
 * (utilities.c:1500-1527): rank r of G tests iterations [r*I/G, (r+1)*I/G) and no data-path
 * collective exists.  The only exchange is this gather (188 doubles = 1.5 KB per stream at the
 * defaults), after which rank 0 owns the p-values of all iterations in job order - the in-memory
 * equivalent of the reference's `-m a` reading every rank's .pvalues file (utilities.c:2053-2180).
 *
 * NCCL is loaded with dlopen so that single-GPU users need no NCCL at all.
 */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <nccl.h>
#include "engine.h"

struct NcclApi {
	void *handle;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *);
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*GroupStart)(void);
	ncclResult_t (*GroupEnd)(void);
	const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl;

static int load_nccl(char *err, size_t errlen)
{
	if (g_nccl.handle)
		return 0;
	const char *names[] = { "libnccl.so.2", "libnccl.so" };
	void *h = NULL;
	for (int i = 0; i < 2 && !h; i++)
		h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
	if (!h) {
		snprintf(err, errlen, "dlopen(libnccl.so.2) failed: %s", dlerror());
		return -1;
	}
#define SYM(field, name)                                                                                          \
	do {                                                                                                      \
		*(void **) (&g_nccl.field) = dlsym(h, name);                                                      \
		if (!g_nccl.field) {                                                                              \
			snprintf(err, errlen, "dlsym(%s) failed", name);                                          \
			return -1;                                                                                \
		}                                                                                                 \
	} while (0)
	SYM(GetUniqueId, "ncclGetUniqueId");
	SYM(CommInitRank, "ncclCommInitRank");
	SYM(CommDestroy, "ncclCommDestroy");
	SYM(Send, "ncclSend");
	SYM(Recv, "ncclRecv");
	SYM(AllReduce, "ncclAllReduce");
	SYM(GroupStart, "ncclGroupStart");
	SYM(GroupEnd, "ncclGroupEnd");
	SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
	g_nccl.handle = h;
	return 0;
}

static char g_comm_err[256];

#define NC(ctx, call)                                                                                             \
	do {                                                                                                      \
		ncclResult_t r_ = (call);                                                                         \
		if (r_ != ncclSuccess) {                                                                          \
			snprintf((ctx)->err, sizeof((ctx)->err), "%s: %s", #call, g_nccl.GetErrorString(r_));     \
			return STSGPU_E_NCCL;                                                                     \
		}                                                                                                 \
	} while (0)

extern "C" int stsgpu_comm_unique_id(uint8_t id[STSGPU_UNIQUE_ID_BYTES])
{
	if (id == NULL)
		return STSGPU_E_INVAL;
	if (load_nccl(g_comm_err, sizeof(g_comm_err)))
		return STSGPU_E_NCCL;
	ncclUniqueId u;
	if (g_nccl.GetUniqueId(&u) != ncclSuccess)
		return STSGPU_E_NCCL;
	static_assert(sizeof(ncclUniqueId) == STSGPU_UNIQUE_ID_BYTES, "ncclUniqueId size");
	memcpy(id, &u, STSGPU_UNIQUE_ID_BYTES);
	return STSGPU_OK;
}

extern "C" int stsgpu_comm_init(stsgpu_ctx *ctx, const uint8_t id[STSGPU_UNIQUE_ID_BYTES], int rank, int nranks)
{
	if (ctx == NULL || id == NULL || nranks < 1 || rank < 0 || rank >= nranks)
		return STSGPU_E_INVAL;
	if (load_nccl(ctx->err, sizeof(ctx->err)))
		return STSGPU_E_NCCL;
	if (cudaSetDevice(ctx->device) != cudaSuccess)
		return STSGPU_E_CUDA;
	ncclUniqueId u;
	memcpy(&u, id, STSGPU_UNIQUE_ID_BYTES);
	ncclComm_t comm;
	NC(ctx, g_nccl.CommInitRank(&comm, nranks, u, rank));
	ctx->nccl_comm = comm;
	ctx->rank = rank;
	ctx->nranks = nranks;
	return STSGPU_OK;
}

void stsgpu_comm_destroy_internal(stsgpu_ctx *ctx)
{
	if (ctx->nccl_comm && g_nccl.handle)
		g_nccl.CommDestroy((ncclComm_t) ctx->nccl_comm);
	ctx->nccl_comm = NULL;
	if (ctx->d_gather) cudaFree(ctx->d_gather);
	ctx->d_gather = NULL;
	for (int i = 0; i < STS_MAX_SLOTS; i++) {
		Slot &S = ctx->slots[i];
		cudaFree(S.d_gather); cudaFreeHost(S.h_gather);
		cudaFree(S.d_hdr); cudaFreeHost(S.h_hdr); cudaFree(S.d_hdr_all); cudaFreeHost(S.h_hdr_all);
		S.d_gather = S.h_gather = S.d_hdr = S.h_hdr = S.d_hdr_all = S.h_hdr_all = NULL;
	}
}

/*
 * The exchange of one batch.  Every rank sends a FIXED number of doubles - the capacity of a batch, max_streams x npv,
 * whatever the batch holds - plus a two-word header {nstreams, first_iteration}: the Send / Recv counts of all ranks
 * agree by construction, so a rank with a shorter batch (the last one of a run) can neither hang nor corrupt the
 * exchange; the root compares the headers afterwards (stsgpu_gather_pvals).  The root compacts the received rows to
 * its own nstreams on the way to the host, so that equal batches arrive as one contiguous rank-major array.
 */
static int ensure_gather_buffers(stsgpu_ctx *ctx, Slot &S, int root)
{
	const size_t cap = (size_t) ctx->prm.max_streams * ctx->lay.npv;
	if (S.d_hdr == NULL) {
		if (cudaMalloc((void **) &S.d_hdr, 2 * sizeof(double)) != cudaSuccess ||
		    cudaMallocHost((void **) &S.h_hdr, 2 * sizeof(double)) != cudaSuccess)
			return STSGPU_E_NOMEM;
	}
	if (ctx->rank == root && S.d_gather == NULL) {
		const size_t need = cap * ctx->nranks;
		if (cudaMalloc((void **) &S.d_gather, need * sizeof(double)) != cudaSuccess ||
		    cudaMallocHost((void **) &S.h_gather, need * sizeof(double)) != cudaSuccess ||
		    cudaMalloc((void **) &S.d_hdr_all, 2 * ctx->nranks * sizeof(double)) != cudaSuccess ||
		    cudaMallocHost((void **) &S.h_hdr_all, 2 * ctx->nranks * sizeof(double)) != cudaSuccess) {
			snprintf(ctx->err, sizeof(ctx->err), "cannot allocate the gather buffers (%zu doubles)", need);
			return STSGPU_E_NOMEM;
		}
	}
	return STSGPU_OK;
}

static int enqueue_gather(stsgpu_ctx *ctx, Slot &S, int root, int64_t nstreams, cudaStream_t stream)
{
	int rc = ensure_gather_buffers(ctx, S, root);
	if (rc != STSGPU_OK)
		return rc;
	ncclComm_t comm = (ncclComm_t) ctx->nccl_comm;
	const size_t cap = (size_t) ctx->prm.max_streams * ctx->lay.npv;
	const size_t mine = (size_t) nstreams * ctx->lay.npv;
	S.h_hdr[0] = (double) nstreams;
	S.h_hdr[1] = (double) S.first;
	if (cudaMemcpyAsync(S.d_hdr, S.h_hdr, 2 * sizeof(double), cudaMemcpyHostToDevice, stream) != cudaSuccess)
		return STSGPU_E_CUDA;
	ncclResult_t r = g_nccl.GroupStart();
	if (r == ncclSuccess) {
		if (ctx->rank == root) {
			for (int p = 0; p < ctx->nranks && r == ncclSuccess; p++) {
				if (p == root)
					continue;
				r = g_nccl.Recv(S.d_gather + (size_t) p * cap, cap, ncclFloat64, p, comm, stream);
				if (r == ncclSuccess)
					r = g_nccl.Recv(S.d_hdr_all + 2 * p, 2, ncclFloat64, p, comm, stream);
			}
		} else {
			r = g_nccl.Send(S.d_pv, cap, ncclFloat64, root, comm, stream);
			if (r == ncclSuccess)
				r = g_nccl.Send(S.d_hdr, 2, ncclFloat64, root, comm, stream);
		}
		const ncclResult_t r2 = g_nccl.GroupEnd();	/* always closed, also after a failed Send / Recv */
		if (r == ncclSuccess)
			r = r2;
	}
	if (r != ncclSuccess) {
		snprintf(ctx->err, sizeof(ctx->err), "gather: %s", g_nccl.GetErrorString(r));
		return STSGPU_E_NCCL;
	}
	if (ctx->rank == root) {
		if (cudaMemcpyAsync(S.d_gather + (size_t) root * cap, S.d_pv, mine * sizeof(double), cudaMemcpyDeviceToDevice, stream) != cudaSuccess ||
		    cudaMemcpyAsync(S.d_hdr_all + 2 * root, S.d_hdr, 2 * sizeof(double), cudaMemcpyDeviceToDevice, stream) != cudaSuccess ||
		    cudaMemcpy2DAsync(S.h_gather, mine * sizeof(double), S.d_gather, cap * sizeof(double), mine * sizeof(double),
				      (size_t) ctx->nranks, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
		    cudaMemcpyAsync(S.h_hdr_all, S.d_hdr_all, 2 * ctx->nranks * sizeof(double), cudaMemcpyDeviceToHost, stream) != cudaSuccess) {
			snprintf(ctx->err, sizeof(ctx->err), "gather copy failed: %s", cudaGetErrorString(cudaGetLastError()));
			return STSGPU_E_CUDA;
		}
	}
	S.gathered = true;
	return STSGPU_OK;
}

int stsgpu_comm_enqueue_gather(stsgpu_ctx *ctx, Slot &S, int64_t nstreams, cudaStream_t stream)
{
	if (ctx->gather_root < 0)
		return STSGPU_OK;
	if (ctx->nranks == 1) {
		S.gathered = true;
		return STSGPU_OK;
	}
	if (S.ascii) {
		snprintf(ctx->err, sizeof(ctx->err), "gather mode does not take -F a batches (their stream count is known only after the batch)");
		return STSGPU_E_UNSUPPORTED;
	}
	return enqueue_gather(ctx, S, ctx->gather_root, nstreams, stream);
}

extern "C" int stsgpu_gather_begin(stsgpu_ctx *ctx, int root)
{
	if (ctx == NULL)
		return STSGPU_E_INVAL;
	if (ctx->submit_seq != ctx->wait_seq) {
		snprintf(ctx->err, sizeof(ctx->err), "stsgpu_gather_begin needs an idle engine");
		return STSGPU_E_STATE;
	}
	if (root == -1) {		/* leave gather mode */
		ctx->gather_root = -1;
		return STSGPU_OK;
	}
	if (root < 0 || root >= ctx->nranks || (ctx->nranks > 1 && ctx->nccl_comm == NULL))
		return STSGPU_E_STATE;
	if (cudaSetDevice(ctx->device) != cudaSuccess)
		return STSGPU_E_CUDA;
	for (int i = 0; ctx->nranks > 1 && i < ctx->nslots; i++) {
		const int rc = ensure_gather_buffers(ctx, ctx->slots[i], root);
		if (rc != STSGPU_OK)
			return rc;
	}
	ctx->gather_root = root;
	return STSGPU_OK;
}

extern "C" int stsgpu_gather_pvals(stsgpu_ctx *ctx, int root, const double **pvals_all)
{
	if (ctx == NULL || pvals_all == NULL)
		return STSGPU_E_INVAL;
	*pvals_all = NULL;
	if (!ctx->have_results) {
		snprintf(ctx->err, sizeof(ctx->err), "gather needs a completed batch");
		return STSGPU_E_STATE;
	}
	Slot &S = ctx->slots[ctx->last_done];
	if (ctx->nranks == 1) {		/* single rank: the batch records are the answer */
		*pvals_all = S.h_pv;
		return STSGPU_OK;
	}
	if (ctx->nccl_comm == NULL || root < 0 || root >= ctx->nranks)
		return STSGPU_E_STATE;
	if (ctx->gather_root >= 0) {
		/* gather mode: the exchange was part of the batch and completed with its stsgpu_wait */
		if (root != ctx->gather_root || !S.gathered) {
			snprintf(ctx->err, sizeof(ctx->err), "the last batch was not gathered to rank %d", root);
			return STSGPU_E_STATE;
		}
	} else {
		/* one blocking exchange of the last completed batch */
		if (S.in_flight) {
			snprintf(ctx->err, sizeof(ctx->err), "the slot of the last completed batch has been submitted again: gather before the next submit, or use stsgpu_gather_begin");
			return STSGPU_E_STATE;
		}
		cudaSetDevice(ctx->device);
		const int rc = enqueue_gather(ctx, S, root, S.nstreams, ctx->s_main);
		if (rc != STSGPU_OK)
			return rc;
		if (cudaStreamSynchronize(ctx->s_main) != cudaSuccess) {
			snprintf(ctx->err, sizeof(ctx->err), "gather failed: %s", cudaGetErrorString(cudaGetLastError()));
			return STSGPU_E_CUDA;
		}
	}
	if (ctx->rank != root)
		return STSGPU_OK;
	for (int p = 0; p < ctx->nranks; p++)
		if ((int64_t) S.h_hdr_all[2 * p] != S.nstreams) {
			snprintf(ctx->err, sizeof(ctx->err), "rank %d completed a batch of %lld streams, rank %d one of %lld: the gathered array would not be rank-major",
				 p, (long long) S.h_hdr_all[2 * p], root, (long long) S.nstreams);
			return STSGPU_E_STATE;
		}
	*pvals_all = S.h_gather;
	return STSGPU_OK;
}

extern "C" int stsgpu_comm_barrier(stsgpu_ctx *ctx)
{
	if (ctx == NULL)
		return STSGPU_E_INVAL;
	if (ctx->nranks == 1)
		return STSGPU_OK;
	if (ctx->nccl_comm == NULL)
		return STSGPU_E_STATE;
	cudaSetDevice(ctx->device);
	if (ctx->d_gather == NULL && cudaMalloc((void **) &ctx->d_gather, 64) != cudaSuccess)
		return STSGPU_E_NOMEM;
	if (cudaMemsetAsync(ctx->d_gather, 0, 64, ctx->s_main) != cudaSuccess)
		return STSGPU_E_CUDA;
	NC(ctx, g_nccl.AllReduce(ctx->d_gather, ctx->d_gather, 1, ncclFloat64, ncclSum, (ncclComm_t) ctx->nccl_comm, ctx->s_main));
	if (cudaStreamSynchronize(ctx->s_main) != cudaSuccess)
		return STSGPU_E_CUDA;
	return STSGPU_OK;
}

/*
 * Sum the assessment tallies of all ranks (include/stsgpu.h): the one collective of the assessment.  In place, so
 * afterwards every rank holds the global sum: a second reduce is a no-op, and adding more batches or p-values to
 * summed tallies is refused until the next stsgpu_assess_begin (engine.cu checks tally_reduced).
 */
extern "C" int stsgpu_assess_reduce(stsgpu_ctx *ctx)
{
	if (ctx == NULL)
		return STSGPU_E_INVAL;
	if (!ctx->assessing || ctx->submit_seq != ctx->wait_seq) {
		snprintf(ctx->err, sizeof(ctx->err), "stsgpu_assess_reduce needs stsgpu_assess_begin and an idle engine");
		return STSGPU_E_STATE;
	}
	if (ctx->nranks == 1 || ctx->tally_reduced)
		return STSGPU_OK;
	if (ctx->nccl_comm == NULL)
		return STSGPU_E_STATE;
	cudaSetDevice(ctx->device);
	const size_t words = (size_t) ctx->lay.npv * (size_t) (ctx->assess_bins + 2);
	NC(ctx, g_nccl.AllReduce(ctx->d_tally, ctx->d_tally, words, ncclUint64, ncclSum, (ncclComm_t) ctx->nccl_comm, ctx->s_main));
	if (cudaStreamSynchronize(ctx->s_main) != cudaSuccess)
		return STSGPU_E_CUDA;
	ctx->tally_reduced = true;
	return STSGPU_OK;
}
