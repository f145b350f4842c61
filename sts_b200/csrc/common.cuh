/*
 * common.cuh - shared device helpers and the engine's internal structures.
 *
 * Bit order: the device input buffer holds each stream's raw file bytes (MSB-first bits,
 * utilities.c:1862-1872) at a 128-byte aligned pitch.  A 32-bit little-endian load followed by a
 * byte swap gives a word whose bit 31 is the earliest stream bit: stream bit k lives in word k/32
 * at bit position 31 - k%32.  Bits past n (inside the last word and in the pad words) are zero.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/stsgpu.h"

#define STS_MAX_KERNELS 32
#define STS_WARP 32

/* device-side constants of one engine, passed by value to the kernels */
struct DevParams {
	int64_t n;			/* bits per stream */
	int64_t nwords;			/* ceil(n / 32) */
	int64_t pitch_words;		/* words between consecutive streams (multiple of 32) */
	int32_t npv, nst, nw;		/* record sizes */
	int32_t pv_off[16], st_off[16];
	int32_t nfs, fs_off[16];	/* double statistics (STSGPU_FLAG_STATS) */
	uint32_t enabled;
	uint32_t want_stats;		/* STSGPU_FLAG_STATS: fill fstats and the last-cycle excursion counters */
	/* per-test parameters */
	int64_t bf_M, bf_N;		/* block frequency */
	int32_t lr_M, lr_min_class, lr_max_class; int64_t lr_N;	/* longest run table row */
	int64_t rank_count;		/* matrices */
	int32_t nonov_m, ntemplates; int64_t nonov_M;
	int64_t ov_N;			/* overlapping blocks of 1032 bits, m = 9 */
	int32_t univ_L; int64_t univ_Q, univ_K;
	int32_t apen_m, serial_m;
	int32_t hist_bits;		/* H = max(serial_m, apen_m + 1) */
	int32_t hist_slice_bits;	/* H - min(H, 15): slices = 2^this */
	int64_t lc_M, lc_N;
	int64_t min_zero_crossings;	/* driver.c:316 */
};

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

/* word i of a stream, MSB-first */
__device__ __forceinline__ uint32_t ldw(const uint32_t *__restrict__ s, int64_t i)
{
	return bswap32(__ldg(s + i));
}

/*
 * 32 stream bits starting at bit position `pos` (MSB-first); reads words pos/32 and pos/32 + 1,
 * both of which exist thanks to the pad words after every stream.
 */
__device__ __forceinline__ uint32_t get32(const uint32_t *__restrict__ s, int64_t pos)
{
	int64_t w = pos >> 5;
	uint32_t sh = (uint32_t) (pos & 31);
	uint32_t hi = ldw(s, w), lo = ldw(s, w + 1);
	return __funnelshift_l(lo, hi, sh);
}

/* number of one bits in stream bit range [a, b) */
__device__ __forceinline__ uint32_t mask_from(uint32_t first_bit)	/* positions >= first_bit (0 = MSB) */
{
	return first_bit >= 32 ? 0u : (0xFFFFFFFFu >> first_bit);
}
__device__ __forceinline__ uint32_t mask_below(uint32_t nbits)	/* the first nbits positions */
{
	return nbits >= 32 ? 0xFFFFFFFFu : (nbits == 0 ? 0u : ~(0xFFFFFFFFu >> nbits));
}

__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ int warp_max(int v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
__device__ __forceinline__ int warp_min(int v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
/* inclusive prefix sum across the warp */
__device__ __forceinline__ int warp_scan_incl(int v, int lane)
{
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		int t = __shfl_up_sync(0xffffffffu, v, o);
		if (lane >= o)
			v += t;
	}
	return v;
}

/*
 * An SM changes its L1 / shared-memory split only when it is idle, and the driver picks the split of a
 * kernel from the shared memory its resident CTAs can use.  A kernel without shared memory therefore
 * leaves the SM with (almost) none, and the DFT's 64-70 KiB CTAs cannot join it: measured here as the DFT
 * making no progress at all while LinearComplexity ran in another stream.  LinearComplexity, the one
 * long kernel without shared memory, therefore asks for dynamic shared memory it never touches
 * (sts_pad_smem), which leaves the SM with a large split that the DFT CTAs can join; padding the
 * other small kernels as well measured slower.  STSGPU_PAD_SMEM=0 switches it off.
 */
#include <stdlib.h>
static inline size_t sts_pad_smem(size_t used, size_t target)
{
	static int on = -1;
	if (on < 0)
		on = getenv("STSGPU_PAD_SMEM") ? atoi(getenv("STSGPU_PAD_SMEM")) : 1;
	return (on && used < target) ? target : used;
}
#define STS_PAD_KERNEL(kernel, bytes) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (bytes))

/* host-side launch bookkeeping shared by the kernel files */
struct LaunchCtx {
	cudaStream_t stream;
	cudaStream_t stream2;		/* DFT only: odd chunks run here, overlapping the even ones (may equal `stream`) */
	cudaEvent_t ev2_fork, ev2_join;
	const uint32_t *d_in;		/* packed input */
	double *d_pv;
	long long *d_st;
	uint32_t *d_w;
	double *d_fs;			/* fstats, or nullptr without STSGPU_FLAG_STATS */
	int64_t nstreams;
	int num_sms;
};

/* kernel launchers (one per .cu file); each returns the cudaError of its launches */
cudaError_t launch_scan(const DevParams &P, const LaunchCtx &L);
cudaError_t launch_last_cycle(const DevParams &P, const LaunchCtx &L);
cudaError_t launch_blocks(const DevParams &P, const LaunchCtx &L);
cudaError_t launch_rank(const DevParams &P, const LaunchCtx &L);
cudaError_t launch_nonoverlapping(const DevParams &P, const LaunchCtx &L, const uint32_t *d_templates);
cudaError_t launch_pattern(const DevParams &P, const LaunchCtx &L, uint32_t *d_apen_hist, uint32_t *d_flags);
cudaError_t launch_universal(const DevParams &P, const LaunchCtx &L, const unsigned long long *d_log2_table, uint32_t *d_dist,
			     int64_t *launches);
size_t universal_scratch_bytes_per_stream(const DevParams &P);
cudaError_t launch_ascii_pack(const DevParams &P, const unsigned char *d_text, long long text_len, uint32_t *d_in,
			      unsigned long long *d_status, int64_t nstreams, cudaStream_t stream);
cudaError_t launch_lc(const DevParams &P, const LaunchCtx &L, const uint8_t *d_class_lut);
