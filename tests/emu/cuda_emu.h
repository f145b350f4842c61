/*
 * cuda_emu.h - TEST INFRASTRUCTURE, NOT PRODUCT CODE: a functional emulation of the small part of CUDA that
 * sts_b200/csrc uses, so that the UNMODIFIED kernel sources can be compiled with g++ and stepped through on a
 * CPU (tests/emu/build_emu.py) when no GPU is at hand.  It exists for one purpose: to check the kernels' logic
 * (index arithmetic, warp collectives, barriers, atomics, record layout) against the oracle on small inputs in
 * the `-m "not gpu"` suite and while developing a kernel without GPU minutes.  It is about a thousand times slower
 * than the GPU, it is never loaded by sts_b200/ (whose library refuses to run without a B200), nothing is ever
 * timed on it, and a pass here proves nothing about performance - only `-m gpu` parity runs count as parity.
 *
 * Execution model: a launch runs its CTAs one after the other (or on a few OS threads); the threads of a CTA are
 * fibers that run until they block in __syncthreads() or in a warp collective (__shfl_*_sync, __ballot_sync,
 * __match_any_sync, __reduce_*_sync, __syncwarp), where the scheduler switches to the next fiber.  Streams and
 * events are synchronous.  `__shared__` is per-OS-thread static storage (one CTA at a time per OS thread).
 */
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <functional>
#include <type_traits>

#define STS_CUDA_EMU 1

/* ---- qualifiers ---------------------------------------------------------------------------- */
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local
#define __constant__ static

/* ---- vector types -------------------------------------------------------------------------- */
struct uint3 { unsigned x, y, z; };
struct dim3 {
	unsigned x, y, z;
	dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) double2 { double x, y; };
struct alignas(16) double4 { double x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
static inline double4 make_double4(double x, double y, double z, double w) { double4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }

/* ---- per-thread context -------------------------------------------------------------------- */
namespace emu {
struct Warp;
struct Ctx {
	uint3 tid, bid;
	dim3 bdim, gdim;
	int lane;
	Warp *warp;
};
extern thread_local Ctx *cur;
void *dyn_smem();
void syncthreads();
/* all-to-all exchange among the lanes of `mask` that are still alive: out[l] = value deposited by lane l */
uint32_t warp_collect(uint32_t mask, uint64_t v, uint64_t out[32]);
void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
template <class F> static inline void launch(dim3 grid, dim3 block, size_t smem, F &&f) { run_grid(grid, block, smem, std::function<void()>(f)); }

template <class T> static inline uint64_t to_bits(T v)
{
	static_assert(sizeof(T) <= 8, "warp collectives carry at most 64 bits");
	uint64_t b = 0;
	memcpy(&b, &v, sizeof(T));
	return b;
}
template <class T> static inline T from_bits(uint64_t b)
{
	T v;
	memcpy(&v, &b, sizeof(T));
	return v;
}
}

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::cur->bid)
#define blockDim (emu::cur->bdim)
#define gridDim (emu::cur->gdim)
#define warpSize 32

/* dynamic shared memory above 48 KiB needs cudaFuncSetAttribute(MaxDynamicSharedMemorySize) first, as on the device */
void emu_set_max_dyn_smem(const void *kernel, int bytes);
bool emu_check_dyn_smem(const void *kernel, size_t bytes, const char *name);	/* false: cudaGetLastError() reports it */
#define EMU_LAUNCH(kernel, grid, block, smem, stream, ...)                                                        \
	(emu_check_dyn_smem(reinterpret_cast<const void *>(&kernel), (size_t) (smem), #kernel)                   \
		 ? emu::launch((grid), (block), (size_t) (smem), [=]() { kernel(__VA_ARGS__); })                  \
		 : (void) 0)

static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu)
{
	uint64_t o[32];
	emu::warp_collect(mask, 0, o);
}

/* ---- warp collectives ---------------------------------------------------------------------- */
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
	uint64_t o[32];
	emu::warp_collect(mask, emu::to_bits(v), o);
	const int lane = emu::cur->lane, base = lane & ~(width - 1);
	return emu::from_bits<T>(o[base + (src & (width - 1))]);
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
	uint64_t o[32];
	emu::warp_collect(mask, emu::to_bits(v), o);
	const int lane = emu::cur->lane, base = lane & ~(width - 1);
	const int s = lane - (int) delta;
	return emu::from_bits<T>(o[s < base ? lane : s]);
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
	uint64_t o[32];
	emu::warp_collect(mask, emu::to_bits(v), o);
	const int lane = emu::cur->lane, base = lane & ~(width - 1);
	const int s = lane + (int) delta;
	return emu::from_bits<T>(o[s >= base + width ? lane : s]);
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32)
{
	uint64_t o[32];
	emu::warp_collect(mask, emu::to_bits(v), o);
	const int lane = emu::cur->lane;
	const int s = lane ^ lanemask;
	return emu::from_bits<T>(o[(s & ~(width - 1)) != (lane & ~(width - 1)) ? lane : s]);
}
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
	uint64_t o[32];
	const uint32_t part = emu::warp_collect(mask, pred ? 1 : 0, o);
	unsigned r = 0;
	for (int l = 0; l < 32; l++)
		if (((part >> l) & 1u) && o[l])
			r |= 1u << l;
	return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred)
{
	uint64_t o[32];
	const uint32_t part = emu::warp_collect(mask, pred ? 1 : 0, o);
	for (int l = 0; l < 32; l++)
		if (((part >> l) & 1u) && !o[l])
			return 0;
	return 1;
}
template <class T> static inline unsigned __match_any_sync(unsigned mask, T v)
{
	uint64_t o[32];
	const uint64_t mine = emu::to_bits(v);
	const uint32_t part = emu::warp_collect(mask, mine, o);
	unsigned r = 0;
	for (int l = 0; l < 32; l++)
		if (((part >> l) & 1u) && o[l] == mine)
			r |= 1u << l;
	return r;
}
#define EMU_REDUCE(name, T, init, expr)                                                                           \
	static inline T name(unsigned mask, T v)                                                                  \
	{                                                                                                         \
		uint64_t o[32];                                                                                   \
		const uint32_t part = emu::warp_collect(mask, emu::to_bits(v), o);                                \
		T acc = init;                                                                                     \
		for (int l = 0; l < 32; l++)                                                                      \
			if ((part >> l) & 1u) {                                                                   \
				const T x = emu::from_bits<T>(o[l]);                                              \
				acc = (expr);                                                                     \
			}                                                                                         \
		return acc;                                                                                       \
	}
EMU_REDUCE(__reduce_min_sync, int, 0x7fffffff, x < acc ? x : acc)
EMU_REDUCE(__reduce_max_sync, int, (int) 0x80000000, x > acc ? x : acc)
EMU_REDUCE(__reduce_add_sync, int, 0, acc + x)
EMU_REDUCE(__reduce_min_sync, unsigned, 0xffffffffu, x < acc ? x : acc)
EMU_REDUCE(__reduce_max_sync, unsigned, 0u, x > acc ? x : acc)
EMU_REDUCE(__reduce_add_sync, unsigned, 0u, acc + x)
EMU_REDUCE(__reduce_xor_sync, unsigned, 0u, acc ^ x)
EMU_REDUCE(__reduce_or_sync, unsigned, 0u, acc | x)
EMU_REDUCE(__reduce_and_sync, unsigned, 0xffffffffu, acc & x)

/* ---- integer / conversion intrinsics --------------------------------------------------------- */
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned) x); }
static inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long) x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned __brev(unsigned x)
{
	x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
	x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
	x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
	return __builtin_bswap32(x);
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{
	const uint64_t v = ((uint64_t) b << 32) | a;
	unsigned r = 0;
	for (int i = 0; i < 4; i++) {
		const unsigned s = (sel >> (4 * i)) & 0xfu;
		unsigned byte = (unsigned) ((v >> (8 * (s & 7u))) & 0xffu);
		if (s & 8u)
			byte = (byte & 0x80u) ? 0xffu : 0u;
		r |= byte << (8 * i);
	}
	return r;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift)
{
	const uint64_t v = ((uint64_t) hi << 32) | lo;
	return (unsigned) ((v << (shift & 31u)) >> 32);
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift)
{
	const uint64_t v = ((uint64_t) hi << 32) | lo;
	return (unsigned) (v >> (shift & 31u));
}
static inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned shift)
{
	const uint64_t v = ((uint64_t) hi << 32) | lo;
	return shift >= 32 ? lo : (unsigned) ((v << shift) >> 32);
}
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned shift)
{
	const uint64_t v = ((uint64_t) hi << 32) | lo;
	return shift >= 32 ? hi : (unsigned) (v >> shift);
}
static inline long long __double_as_longlong(double d) { return emu::from_bits<long long>(emu::to_bits(d)); }
static inline double __longlong_as_double(long long v) { return emu::from_bits<double>((uint64_t) v); }
static inline int __double2hiint(double d) { return (int) (emu::to_bits(d) >> 32); }
static inline int __double2loint(double d) { return (int) (emu::to_bits(d) & 0xffffffffu); }
static inline double __hiloint2double(int hi, int lo) { return emu::from_bits<double>(((uint64_t) (unsigned) hi << 32) | (unsigned) lo); }
static inline double __ull2double_rn(unsigned long long v) { return (double) v; }
static inline double __ll2double_rn(long long v) { return (double) v; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long) (((unsigned __int128) a * b) >> 64); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned) (((uint64_t) a * b) >> 32); }
#define __log2f(x) log2f(x)	/* glibc declares a function of this name */
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *p; }
template <class T> static inline void __stcg(T *p, T v) { *p = v; }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
/*
 * mbarrier + bulk copy (sts_b200/csrc/tma.cuh).  The 64-bit barrier word holds {phase:1, arrivals pending:15,
 * arrivals per phase:16, transaction bytes pending:32 (signed)}; the threads of a CTA are cooperative fibers on one
 * OS thread, so plain read-modify-write is enough.  A bulk copy is performed at issue time.
 */
namespace emu { void fiber_yield(); void note_progress(); }
static inline void emu_mbar_settle(uint64_t *bar)
{
	uint64_t v = *bar;
	const uint32_t pending = (uint32_t) ((v >> 1) & 0x7fffu), per_phase = (uint32_t) ((v >> 16) & 0xffffu);
	const int32_t tx = (int32_t) (v >> 32);
	if (pending == 0 && tx == 0)		/* phase complete: flip, re-arm */
		*bar = ((v & 1u) ^ 1u) | ((uint64_t) per_phase << 1) | ((uint64_t) per_phase << 16);
}
static inline void mbar_init(uint64_t *bar, uint32_t arrivals) { *bar = ((uint64_t) arrivals << 1) | ((uint64_t) arrivals << 16); }
static inline void mbar_fence_init() {}
static inline void emu_mbar_update(uint64_t *bar, int arrive, int64_t tx_delta)
{
	uint64_t v = *bar;
	uint32_t pending = (uint32_t) ((v >> 1) & 0x7fffu);
	int32_t tx = (int32_t) (v >> 32);
	if (arrive) {
		if (pending == 0) { fprintf(stderr, "cuda_emu: mbarrier arrival beyond its count\n"); abort(); }
		pending--;
	}
	tx += (int32_t) tx_delta;
	*bar = (v & 0xffff0001ull) | ((uint64_t) pending << 1) | ((uint64_t) (uint32_t) tx << 32);
	emu_mbar_settle(bar);
	emu::note_progress();
}
static inline void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) { emu_mbar_update(bar, 1, (int64_t) bytes); }
static inline void mbar_arrive(uint64_t *bar) { emu_mbar_update(bar, 1, 0); }
static inline void mbar_wait(uint64_t *bar, uint32_t parity)
{
	while ((uint32_t) (*(volatile uint64_t *) bar & 1u) == parity)	/* still in the phase with this parity */
		emu::fiber_yield();
}
static inline void bulk_copy_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
	if ((bytes & 15u) || ((uintptr_t) smem_dst & 15u) || ((uintptr_t) gmem_src & 15u)) {
		fprintf(stderr, "cuda_emu: bulk copy needs 16-byte aligned addresses and size\n");
		abort();
	}
	memcpy(smem_dst, gmem_src, bytes);
	emu_mbar_update(bar, 0, -(int64_t) bytes);
}
static inline void fence_proxy_async() {}
void emu_nanosleep(unsigned ns);	/* lets the OS thread's other work (other CTAs on other OS threads) make progress */
static inline void __nanosleep(unsigned ns) { emu_nanosleep(ns); }
template <class T> static inline size_t __cvta_generic_to_shared(const T *p) { return (size_t) p; }

/* CUDA's global min / max accept mixed integer types */
template <class A, class B> static inline typename std::common_type<A, B>::type min(A a, B b)
{
	typedef typename std::common_type<A, B>::type C;
	return (C) a < (C) b ? (C) a : (C) b;
}
template <class A, class B> static inline typename std::common_type<A, B>::type max(A a, B b)
{
	typedef typename std::common_type<A, B>::type C;
	return (C) a > (C) b ? (C) a : (C) b;
}

/* ---- atomics (CTAs may run on several OS threads) ---------------------------------------------- */
template <class T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline double atomicAdd(double *p, double v)
{
	uint64_t old = __atomic_load_n((uint64_t *) p, __ATOMIC_RELAXED), want;
	do {
		want = emu::to_bits(emu::from_bits<double>(old) + v);
	} while (!__atomic_compare_exchange_n((uint64_t *) p, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
	return emu::from_bits<double>(old);
}
template <class T> static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicAnd(T *p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicXor(T *p, T v) { return __atomic_fetch_xor(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicCAS(T *p, T cmp, T v)
{
	__atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
	return cmp;
}
template <class T> static inline T atomicMin(T *p, T v)
{
	T old = __atomic_load_n(p, __ATOMIC_RELAXED);
	while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED))
		;
	return old;
}
template <class T> static inline T atomicMax(T *p, T v)
{
	T old = __atomic_load_n(p, __ATOMIC_RELAXED);
	while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED))
		;
	return old;
}

/* ---- runtime API (synchronous; cuda_emu_rt.cpp) ------------------------------------------------- */
typedef int cudaError_t;
typedef cudaError_t cudaError;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorNoDevice = 100, cudaErrorNotSupported = 801 };
struct emu_stream;
struct emu_event;
typedef emu_stream *cudaStream_t;
typedef emu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamDefault = 0, cudaStreamNonBlocking = 1 };
enum { cudaEventDefault = 0, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
struct cudaDeviceProp {
	char name[256];
	int major, minor, multiProcessorCount;
	size_t totalGlobalMem, sharedMemPerBlockOptin;
};

cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d);
cudaError_t cudaDeviceGetPCIBusId(char *buf, int len, int d);
cudaError_t cudaDeviceSynchronize(void);
cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi);
cudaError_t cudaGetLastError(void);
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaMalloc(void **p, size_t bytes);
cudaError_t cudaFree(void *p);
cudaError_t cudaMallocHost(void **p, size_t bytes);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemset(void *p, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t s = 0);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s = 0);
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
			      cudaMemcpyKind kind, cudaStream_t s = 0);
cudaError_t cudaStreamCreate(cudaStream_t *s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned flags, int prio);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = 0);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
template <class T> static inline cudaError_t cudaFuncSetAttribute(T *f, int attr, int v)
{
	if (attr == cudaFuncAttributeMaxDynamicSharedMemorySize) {
		if (v > 227 * 1024)
			return cudaErrorInvalidValue;
		emu_set_max_dyn_smem(reinterpret_cast<const void *>(f), v);
	}
	return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc((void **) p, bytes); }
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t bytes) { return cudaMallocHost((void **) p, bytes); }
