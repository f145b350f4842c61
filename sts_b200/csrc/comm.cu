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
	if (ctx->h_gather) cudaFreeHost(ctx->h_gather);
	ctx->d_gather = ctx->h_gather = NULL;
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
	const size_t per_rank = (size_t) ctx->cur_nstreams * ctx->lay.npv;
	if (ctx->nranks == 1) {		/* single rank: the batch records are the answer */
		*pvals_all = ctx->h_pv;
		return STSGPU_OK;
	}
	if (ctx->nccl_comm == NULL || root < 0 || root >= ctx->nranks)
		return STSGPU_E_STATE;
	cudaSetDevice(ctx->device);
	cudaStream_t sm = ctx->s_main;
	ncclComm_t comm = (ncclComm_t) ctx->nccl_comm;
	if (ctx->rank == root) {
		const size_t need = per_rank * ctx->nranks;
		if (need > ctx->gather_cap) {
			if (ctx->d_gather) cudaFree(ctx->d_gather);
			if (ctx->h_gather) cudaFreeHost(ctx->h_gather);
			ctx->d_gather = ctx->h_gather = NULL;
			if (cudaMalloc((void **) &ctx->d_gather, need * sizeof(double)) != cudaSuccess ||
			    cudaMallocHost((void **) &ctx->h_gather, need * sizeof(double)) != cudaSuccess)
				return STSGPU_E_NOMEM;
			ctx->gather_cap = need;
		}
		NC(ctx, g_nccl.GroupStart());
		for (int r = 0; r < ctx->nranks; r++) {
			if (r == root)
				continue;
			NC(ctx, g_nccl.Recv(ctx->d_gather + (size_t) r * per_rank, per_rank, ncclFloat64, r, comm, sm));
		}
		NC(ctx, g_nccl.GroupEnd());
		if (cudaMemcpyAsync(ctx->d_gather + (size_t) root * per_rank, ctx->d_pv, per_rank * sizeof(double),
				    cudaMemcpyDeviceToDevice, sm) != cudaSuccess ||
		    cudaMemcpyAsync(ctx->h_gather, ctx->d_gather, need * sizeof(double), cudaMemcpyDeviceToHost, sm) != cudaSuccess ||
		    cudaStreamSynchronize(sm) != cudaSuccess) {
			snprintf(ctx->err, sizeof(ctx->err), "gather copy failed: %s", cudaGetErrorString(cudaGetLastError()));
			return STSGPU_E_CUDA;
		}
		*pvals_all = ctx->h_gather;
	} else {
		NC(ctx, g_nccl.GroupStart());
		NC(ctx, g_nccl.Send(ctx->d_pv, per_rank, ncclFloat64, root, comm, sm));
		NC(ctx, g_nccl.GroupEnd());
		if (cudaStreamSynchronize(sm) != cudaSuccess)
			return STSGPU_E_CUDA;
	}
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
	if (ctx->d_gather == NULL && ctx->gather_cap == 0) {
		if (cudaMalloc((void **) &ctx->d_gather, 64) != cudaSuccess)
			return STSGPU_E_NOMEM;
		ctx->gather_cap = 8;
		if (cudaMallocHost((void **) &ctx->h_gather, 64) != cudaSuccess)
			return STSGPU_E_NOMEM;
	}
	NC(ctx, g_nccl.AllReduce(ctx->d_gather, ctx->d_gather, 1, ncclFloat64, ncclSum, (ncclComm_t) ctx->nccl_comm, ctx->s_main));
	if (cudaStreamSynchronize(ctx->s_main) != cudaSuccess)
		return STSGPU_E_CUDA;
	return STSGPU_OK;
}

/* sum the assessment tallies of all ranks (include/stsgpu.h): the one collective of the assessment */
extern "C" int stsgpu_assess_reduce(stsgpu_ctx *ctx)
{
	if (ctx == NULL)
		return STSGPU_E_INVAL;
	if (!ctx->assessing || ctx->submit_seq != ctx->wait_seq) {
		snprintf(ctx->err, sizeof(ctx->err), "stsgpu_assess_reduce needs stsgpu_assess_begin and an idle engine");
		return STSGPU_E_STATE;
	}
	if (ctx->nranks == 1)
		return STSGPU_OK;
	if (ctx->nccl_comm == NULL)
		return STSGPU_E_STATE;
	cudaSetDevice(ctx->device);
	const size_t words = (size_t) ctx->lay.npv * (size_t) (ctx->assess_bins + 2);
	NC(ctx, g_nccl.AllReduce(ctx->d_tally, ctx->d_tally, words, ncclUint64, ncclSum, (ncclComm_t) ctx->nccl_comm, ctx->s_main));
	if (cudaStreamSynchronize(ctx->s_main) != cudaSuccess)
		return STSGPU_E_CUDA;
	return STSGPU_OK;
}
