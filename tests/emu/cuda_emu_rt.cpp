/*
 * cuda_emu_rt.cpp - TEST INFRASTRUCTURE (see cuda_emu.h): the fiber scheduler that steps the threads of a CTA
 * through barriers and warp collectives, and a synchronous stand-in for the few CUDA runtime calls the engine makes.
 */
#include <sched.h>
#include "cuda_emu.h"
#include <stdio.h>
#include <time.h>
#include <sys/mman.h>
#include <vector>
#include <thread>
#include <mutex>
#include <atomic>

/* void emu_switch(void **save_sp, void *next_sp): save the callee-saved registers, swap stacks, restore */
asm(R"(
.text
.globl emu_switch
.hidden emu_switch
.type emu_switch,@function
emu_switch:
	pushq %rbp
	pushq %rbx
	pushq %r12
	pushq %r13
	pushq %r14
	pushq %r15
	movq %rsp, (%rdi)
	movq %rsi, %rsp
	popq %r15
	popq %r14
	popq %r13
	popq %r12
	popq %rbx
	popq %rbp
	ret
.size emu_switch,.-emu_switch
)");
extern "C" void emu_switch(void **save_sp, void *next_sp);

namespace emu {

enum { RUN = 0, WAIT_BAR = 1, WAIT_WARP = 2, WAIT_DRAIN = 3 };
static const size_t STACK_BYTES = 256 * 1024;
static const int MAX_THREADS = 1024;
static const size_t DYN_SMEM_BYTES = 256 * 1024;

/*
 * Warp collectives.  Lanes deposit a value together with the member mask they name; a group is released when every
 * live lane of a mask has arrived naming that same mask, so disjoint sub-groups of a warp (k_lc.cu: four blocks of
 * eight lanes) run their collectives independently of each other, as on the device.  On release every member gets a
 * private copy of the group's values: a fast lane may deposit for its next collective before a slow one has read.
 */
struct Warp {
	uint64_t val[32];
	uint32_t lmask[32];			/* the mask each arrived lane named */
	uint32_t done[32];			/* members of the released collective this lane took part in, 0: not released */
	uint64_t outv[32][32];
	uint32_t alive, arrived;
};
struct Fiber {
	void *sp;
	Ctx ctx;
	bool done;
	int wait;
	unsigned bar_seen;
};
struct Block {
	Fiber fib[MAX_THREADS];
	Warp warps[MAX_THREADS / 32];
	int n, alive, bar_arrived;
	unsigned bar_gen;
	unsigned long ticks;			/* bumped whenever any fiber makes progress past a wait */
	void *sched_sp;
	const std::function<void()> *body;
	char *stacks;
	char *smem;
};

thread_local Ctx *cur = nullptr;
static thread_local Block *blk = nullptr;
static thread_local Fiber *curf = nullptr;

static Block *new_block()
{
	Block *b = new Block();
	b->stacks = (char *) mmap(nullptr, STACK_BYTES * MAX_THREADS, PROT_READ | PROT_WRITE,
				  MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
	if (b->stacks == (char *) MAP_FAILED) {
		fprintf(stderr, "cuda_emu: cannot map fiber stacks\n");
		abort();
	}
	if (posix_memalign((void **) &b->smem, 1024, DYN_SMEM_BYTES) != 0)
		abort();
	return b;
}

void *dyn_smem() { return blk->smem; }

/* a thread that polls (an mbarrier another thread of the CTA will complete) lets the other fibers run; polling is not
 * progress (the deadlock check must still fire when every thread polls for ever), an mbarrier update is */
void fiber_yield() { emu_switch(&curf->sp, blk->sched_sp); }
void note_progress() { blk->ticks++; }

static inline void yield() { emu_switch(&curf->sp, blk->sched_sp); }

static void warp_try_release(Warp &w, uint32_t mask)
{
	const uint32_t need = mask & w.alive;
	if (need == 0 || (w.arrived & need) != need)
		return;
	for (int l = 0; l < 32; l++)
		if (((need >> l) & 1u) && w.lmask[l] != mask)
			return;			/* that lane waits in another collective */
	for (int p = 0; p < 32; p++) {
		if (!((need >> p) & 1u))
			continue;
		for (int l = 0; l < 32; l++)	/* reading a lane that does not take part is undefined in CUDA: poison it */
			w.outv[p][l] = ((need >> l) & 1u) ? w.val[l] : 0xDEADBEEFCAFEF00Dull;
		w.done[p] = need;
	}
	w.arrived &= ~need;
	blk->ticks++;
}

/* a lane left (exit): groups that only waited for it are complete now */
static void warp_recheck(Warp &w)
{
	for (int l = 0; l < 32; l++)
		if ((w.arrived >> l) & 1u)
			warp_try_release(w, w.lmask[l]);
}

static void fiber_exit()
{
	Block *b = blk;
	Fiber *f = curf;
	f->done = true;
	b->alive--;
	b->ticks++;
	Warp &w = *f->ctx.warp;
	w.alive &= ~(1u << f->ctx.lane);
	warp_recheck(w);			/* the lanes still waiting no longer wait for this one */
	if (b->bar_arrived > 0 && b->bar_arrived == b->alive) {	/* an exited thread counts as arrived */
		b->bar_arrived = 0;
		b->bar_gen++;
	}
	void *dummy;
	emu_switch(&dummy, b->sched_sp);
	abort();				/* a finished fiber is never resumed */
}

static void fiber_entry()
{
	(*blk->body)();
	fiber_exit();
}

void syncthreads()
{
	Block *b = blk;
	Fiber *f = curf;
	const unsigned gen = b->bar_gen;
	b->ticks++;
	if (++b->bar_arrived == b->alive) {
		b->bar_arrived = 0;
		b->bar_gen++;
		return;
	}
	f->bar_seen = gen;
	f->wait = WAIT_BAR;
	while (b->bar_gen == gen)
		yield();
	f->wait = RUN;
}

uint32_t warp_collect(uint32_t mask, uint64_t v, uint64_t out[32])
{
	Fiber *f = curf;
	Warp &w = *f->ctx.warp;
	const int lane = f->ctx.lane;
	w.val[lane] = v;
	w.lmask[lane] = mask;
	w.done[lane] = 0;
	w.arrived |= 1u << lane;
	blk->ticks++;
	warp_try_release(w, mask);
	while (w.done[lane] == 0) {
		f->wait = WAIT_WARP;
		yield();
	}
	f->wait = RUN;
	const uint32_t part = w.done[lane];
	w.done[lane] = 0;
	for (int l = 0; l < 32; l++)
		out[l] = w.outv[lane][l];
	return part;
}

static inline bool runnable(const Block *b, const Fiber *f)
{
	if (f->done)
		return false;
	const Warp &w = *f->ctx.warp;
	switch (f->wait) {
	case WAIT_BAR: return b->bar_gen != f->bar_seen;
	case WAIT_WARP: return w.done[f->ctx.lane] != 0;
	default: return true;
	}
}

static void run_block(Block *b, dim3 grid, dim3 block, uint3 bid, size_t smem)
{
	const int n = (int) (block.x * block.y * block.z);
	b->n = n;
	b->alive = n;
	b->bar_arrived = 0;
	b->bar_gen = 0;
	b->ticks = 0;
	if (smem)
		memset(b->smem, 0xCD, smem);		/* shared memory starts as garbage */
	const int nwarps = (n + 31) / 32;
	for (int w = 0; w < nwarps; w++) {
		Warp &W = b->warps[w];
		const int lanes = n - 32 * w >= 32 ? 32 : n - 32 * w;
		W.alive = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1u);
		W.arrived = 0;
		memset(W.done, 0, sizeof(W.done));
	}
	for (int t = 0; t < n; t++) {
		Fiber &f = b->fib[t];
		f.done = false;
		f.wait = RUN;
		f.ctx.tid.x = t % block.x;
		f.ctx.tid.y = (t / block.x) % block.y;
		f.ctx.tid.z = t / (block.x * block.y);
		f.ctx.bid = bid;
		f.ctx.bdim = block;
		f.ctx.gdim = grid;
		f.ctx.lane = t & 31;
		f.ctx.warp = &b->warps[t >> 5];
		/* initial frame: six zeroed callee-saved registers, then the entry point as the return address */
		uintptr_t top = (uintptr_t) (b->stacks + (size_t) (t + 1) * STACK_BYTES);
		top &= ~(uintptr_t) 15;
		void **sp = (void **) top;
		*--sp = nullptr;			/* the return address fiber_entry never uses */
		*--sp = (void *) fiber_entry;
		for (int r = 0; r < 6; r++)
			*--sp = nullptr;
		f.sp = (void *) sp;
	}
	/*
	 * STS_EMU_SCHED picks the order in which runnable fibers are resumed: "forward" (default: warp 0 first, lane 0
	 * first), "reverse", or "random[:seed]" (a new permutation of warps and lanes on every pass).  Results must not
	 * depend on it: a kernel that lacks a __syncthreads / __syncwarp between a write and a read of shared memory
	 * gives different answers under different orders - a cheap race check.
	 */
	static const int sched = []() {
		const char *e = getenv("STS_EMU_SCHED");
		return e == nullptr ? 0 : (e[0] == 'r' && e[1] == 'e') ? 1 : (e[0] == 'r' && e[1] == 'a') ? 2 : 0;
	}();
	static const unsigned seed0 = []() {
		const char *e = getenv("STS_EMU_SCHED");
		const char *c = e ? strchr(e, ':') : nullptr;
		return c ? (unsigned) atoi(c + 1) : 12345u;
	}();
	uint64_t rng = 0x9E3779B97F4A7C15ull * (seed0 + 1) + bid.x * 1000003ull + bid.y * 10007ull + bid.z;
	int worder[MAX_THREADS / 32], lorder[32];
	while (b->alive > 0) {
		const unsigned long before = b->ticks;
		for (int i = 0; i < nwarps; i++)
			worder[i] = sched == 1 ? nwarps - 1 - i : i;
		for (int i = 0; i < 32; i++)
			lorder[i] = sched == 1 ? 31 - i : i;
		if (sched == 2) {
			for (int i = nwarps - 1; i > 0; i--) {
				rng = rng * 6364136223846793005ull + 1442695040888963407ull;
				const int j = (int) ((rng >> 33) % (unsigned) (i + 1));
				const int tmp = worder[i]; worder[i] = worder[j]; worder[j] = tmp;
			}
			for (int i = 31; i > 0; i--) {
				rng = rng * 6364136223846793005ull + 1442695040888963407ull;
				const int j = (int) ((rng >> 33) % (unsigned) (i + 1));
				const int tmp = lorder[i]; lorder[i] = lorder[j]; lorder[j] = tmp;
			}
		}
		for (int wi = 0; wi < nwarps; wi++) {
			const int w = worder[wi];
			/* keep a warp going until its lanes are finished or wait for the other warps */
			for (;;) {
				const unsigned long wbefore = b->ticks;
				for (int li = 0; li < 32; li++) {
					const int t = 32 * w + lorder[li];
					if (t >= n)
						continue;
					Fiber *f = &b->fib[t];
					if (!runnable(b, f))
						continue;
					curf = f;
					cur = &f->ctx;
					emu_switch(&b->sched_sp, f->sp);
				}
				if (b->ticks == wbefore)
					break;
			}
		}
		if (b->ticks == before) {
			fprintf(stderr, "cuda_emu: deadlock in CTA (%u,%u,%u): %d threads alive, %d at the barrier\n", bid.x, bid.y,
				bid.z, b->alive, b->bar_arrived);
			abort();
		}
	}
	cur = nullptr;
	curf = nullptr;
}

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body)
{
	const long total = (long) grid.x * grid.y * grid.z;
	if (total <= 0 || block.x * block.y * block.z == 0 || block.x * block.y * block.z > (unsigned) MAX_THREADS ||
	    smem > DYN_SMEM_BYTES || grid.y > 65535u || grid.z > 65535u || grid.x > 2147483647u) {	/* the device's limits */
		fprintf(stderr, "cuda_emu: bad launch configuration (%u,%u,%u) x (%u,%u,%u), %zu bytes\n", grid.x, grid.y, grid.z,
			block.x, block.y, block.z, smem);
		abort();
	}
	/* CTAs are dealt to a few OS threads; every worker slot keeps its fiber stacks between launches */
	static std::vector<Block *> pool;
	static std::mutex pool_mutex;
	std::lock_guard<std::mutex> guard(pool_mutex);		/* one launch at a time */
	int workers = (int) std::thread::hardware_concurrency();
	const char *e = getenv("STS_EMU_THREADS");
	if (e && atoi(e) > 0)
		workers = atoi(e);
	if (workers > 16) workers = 16;
	if (workers < 1) workers = 1;
	if ((long) workers > total) workers = (int) total;
	while ((int) pool.size() < workers)
		pool.push_back(new_block());
	std::atomic<long> next(0);
	auto work = [&](int slot) {
		Block *b = pool[slot];
		blk = b;
		b->body = &body;
		for (long i = next.fetch_add(1); i < total; i = next.fetch_add(1)) {
			uint3 bid;
			bid.x = (unsigned) (i % grid.x);
			bid.y = (unsigned) ((i / grid.x) % grid.y);
			bid.z = (unsigned) (i / ((long) grid.x * grid.y));
			run_block(b, grid, block, bid, smem);
		}
		blk = nullptr;
	};
	if (workers == 1) {
		work(0);
	} else {
		std::vector<std::thread> th;
		for (int w = 1; w < workers; w++)
			th.emplace_back(work, w);
		work(0);
		for (auto &t : th)
			t.join();
	}
}

}	/* namespace emu */

/* marker by which sts_b200/binding.py recognises (and, outside the test suite, refuses) this library */
extern "C" int stsgpu_is_emulation(void) { return 1; }

/* ---- runtime ------------------------------------------------------------------------------------ */
#include <map>
static std::map<const void *, int> g_max_dyn_smem;
static std::mutex g_attr_mutex;
void emu_set_max_dyn_smem(const void *kernel, int bytes)
{
	std::lock_guard<std::mutex> g(g_attr_mutex);
	g_max_dyn_smem[kernel] = bytes;
}
static cudaError_t g_last_error = cudaSuccess;
bool emu_check_dyn_smem(const void *kernel, size_t bytes, const char *name)
{
	std::lock_guard<std::mutex> g(g_attr_mutex);
	const auto it = g_max_dyn_smem.find(kernel);
	const size_t limit = it == g_max_dyn_smem.end() ? 48 * 1024 : (size_t) (it->second > 48 * 1024 ? it->second : 48 * 1024);
	if (bytes > limit) {
		fprintf(stderr, "cuda_emu: launch of %s asks for %zu bytes of dynamic shared memory, the limit set for it is %zu "
			"(cudaErrorInvalidValue on the device)\n", name, bytes, limit);
		g_last_error = cudaErrorInvalidValue;
		return false;
	}
	return true;
}
struct emu_stream { int id; };
struct emu_event { double t_ms; };

static double now_ms()
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d)
{
	if (d != 0)
		return cudaErrorInvalidValue;
	memset(p, 0, sizeof(*p));
	snprintf(p->name, sizeof(p->name), "cuda_emu (CPU, tests only)");
	p->major = 10;
	p->minor = 0;
	const char *e = getenv("STS_EMU_SMS");
	p->multiProcessorCount = e && atoi(e) > 0 ? atoi(e) : 4;
	p->totalGlobalMem = (size_t) 8 << 30;
	p->sharedMemPerBlockOptin = 227 * 1024;
	return cudaSuccess;
}
cudaError_t cudaDeviceGetPCIBusId(char *, int, int) { return cudaErrorNotSupported; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -1; return cudaSuccess; }
cudaError_t cudaGetLastError(void)
{
	std::lock_guard<std::mutex> g(g_attr_mutex);
	const cudaError_t e = g_last_error;
	g_last_error = cudaSuccess;
	return e;
}
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory" : "cuda_emu error"); }
cudaError_t cudaMalloc(void **p, size_t bytes)
{
	*p = nullptr;
	if (posix_memalign(p, 256, bytes ? bytes : 1) != 0)
		return cudaErrorMemoryAllocation;
	memset(*p, 0xA5, bytes);			/* device memory starts as garbage */
	return cudaSuccess;
}
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void **p, size_t bytes)
{
	*p = nullptr;
	return posix_memalign(p, 256, bytes ? bytes : 1) == 0 ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemset(void *p, int v, size_t bytes) { memset(p, v, bytes); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t) { memset(p, v, bytes); return cudaSuccess; }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind) { memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t) { memmove(dst, src, bytes); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaMemcpyKind,
			      cudaStream_t)
{
	for (size_t r = 0; r < height; r++)
		memmove((char *) dst + r * dpitch, (const char *) src + r * spitch, width);
	return cudaSuccess;
}
cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = new emu_stream(); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { return cudaStreamCreate(s); }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { return cudaStreamCreate(s); }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event(); (*e)->t_ms = 0; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t_ms = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float) (b->t_ms - a->t_ms); return cudaSuccess; }

void emu_nanosleep(unsigned ns)
{
	(void) ns;
	sched_yield();
}
