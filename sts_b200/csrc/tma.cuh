/*
 * tma.cuh - the asynchronous-copy primitives of sm_90+/sm_100: mbarrier objects in shared memory and the bulk
 * (1-D TMA) copy global -> shared that completes on one (cp.async.bulk, SASS UBLKCP), as thin wrappers over the PTX.
 *
 * Usage pattern (k_dft_rows_tma.cu, k_scan.cu): every thread that issues copies first announces their bytes with
 * mbar_arrive_expect_tx (the barrier is initialised with one arrival per issuing thread), then issues them; consumers
 * spin on mbar_wait with the parity of the phase they expect.  Shared memory that was read or written through normal
 * loads and stores must be fenced with fence_proxy_async before a bulk copy may overwrite it.
 *
 * For the CPU emulation build of tests/emu (STS_CUDA_EMU) the same names are functional stand-ins defined in
 * tests/emu/cuda_emu.h: a copy is a memcpy at issue time, a wait yields to the other threads of the CTA.
 */
#pragma once
#include <stdint.h>

#ifndef STS_CUDA_EMU
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
/* makes the initialised barriers visible to the async proxy (the TMA unit) */
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
/* block until the phase with this parity has completed */
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile("{\n"
		     ".reg .pred P1;\n"
		     "LAB_WAIT:\n"
		     "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
		     "@P1 bra DONE;\n"
		     "bra LAB_WAIT;\n"
		     "DONE:\n"
		     "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
/* bytes: a multiple of 16; dst and src 16-byte aligned */
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
		     "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
/* orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy (bulk copy) accesses */
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif
