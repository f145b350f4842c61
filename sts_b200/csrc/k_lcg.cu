/*
 * k_lcg.cu - the synthetic input: generator 1 (LCG) of the reference's tools/generators.c, produced
 * directly into the engine's device input buffer.
 *
 * generators.c:797-831 iterates x <- 950706376 x mod (2^31 - 1) from x0 = 23482349 in doubles (exact
 * below 2^53, so identical to integer arithmetic); output bit k (k = 0, 1, ...) is x_{k+1} >= 2^30
 * (:870-882, DUNIF >= 0.5); write_sequence packs bit 8j + c of the sequence into bit c of file byte
 * j (LSB first, :1640-1660), and the state carries over from stream to stream.  One thread makes
 * one 32-bit word of the file: it jumps to its position with a^k mod p by square-and-multiply and
 * then steps 32 times.  A little-endian 32-bit store of "bit i = sequence bit 32 t + i" is exactly
 * the 4 file bytes.
 */
#include "common.cuh"

#define LCG_A 950706376ull
#define LCG_P 2147483647ull
#define LCG_X0 23482349ull

__device__ __forceinline__ uint64_t mulmod31(uint64_t a, uint64_t b)
{
	uint64_t v = a * b;			/* < 2^62 */
	v = (v & LCG_P) + (v >> 31);
	v = (v & LCG_P) + (v >> 31);
	return v >= LCG_P ? v - LCG_P : v;
}

/*
 * pow32[j] = a^(32 j) mod p, j < 256: the distance between the words of two threads of a CTA.  When n is a multiple of 32
 * the words of consecutive streams follow each other in the generator's bit sequence, so thread 0 of a CTA jumps to the
 * CTA's first word by square-and-multiply (about 70 modular multiplications) and every thread gets its own state with
 * ONE multiplication by pow32[threadIdx.x] instead of its own jump (3.1 -> about 1 ms per 1024 streams).
 */
__global__ void __launch_bounds__(256)
lcg_kernel(DevParams P, uint32_t *__restrict__ out, int64_t first_stream, int64_t nstreams, const uint32_t *__restrict__ pow32)
{
	__shared__ uint64_t cta_x;
	const int64_t words_per_stream = P.nwords;
	const int64_t gid = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	const bool contiguous = pow32 != nullptr;		/* n % 32 == 0 */
	if (contiguous) {
		if (threadIdx.x == 0) {
			uint64_t x = LCG_X0, base = LCG_A;
			uint64_t e = (uint64_t) (first_stream * P.n) + 32ull * (uint64_t) blockIdx.x * blockDim.x;
			while (e) {
				if (e & 1ull)
					x = mulmod31(x, base);
				base = mulmod31(base, base);
				e >>= 1;
			}
			cta_x = x;
		}
		__syncthreads();
	}
	if (gid >= nstreams * words_per_stream)
		return;
	const int64_t s = gid / words_per_stream, t = gid % words_per_stream;
	uint64_t x;
	if (contiguous) {
		x = mulmod31(cta_x, (uint64_t) pow32[threadIdx.x]);
	} else {
		const uint64_t k0 = (uint64_t) ((first_stream + s) * P.n + 32 * t);	/* index of my first output bit */
		/* state before producing bit k0 is x_{k0} = a^{k0} x0 */
		uint64_t base = LCG_A, e = k0;
		x = LCG_X0;
		while (e) {
			if (e & 1ull)
				x = mulmod31(x, base);
			base = mulmod31(base, base);
			e >>= 1;
		}
	}
	const int64_t rem = P.n - 32 * t;
	const int valid = rem >= 32 ? 32 : (int) rem;
	uint32_t word = 0;
	for (int i = 0; i < valid; i++) {
		x = mulmod31(x, LCG_A);
		if (x >= (1ull << 30))
			word |= 1u << i;
	}
	out[s * P.pitch_words + t] = word;
}

/* a^(32 j) mod p for j < 256 (host) */
void lcg_pow32_table(uint32_t tab[256])
{
	uint64_t a32 = 1;
	for (int i = 0; i < 32; i++)
		a32 = (a32 * LCG_A) % LCG_P;
	uint64_t v = 1;
	for (int j = 0; j < 256; j++) {
		tab[j] = (uint32_t) v;
		v = (v * a32) % LCG_P;
	}
}

cudaError_t launch_lcg(const DevParams &P, uint32_t *d_in, int64_t first_stream, int64_t nstreams, const uint32_t *d_pow32,
		       cudaStream_t stream)
{
	const int64_t total = nstreams * P.nwords;
	const unsigned blocks = (unsigned) ((total + 255) / 256);
	lcg_kernel<<<blocks, 256, 0, stream>>>(P, d_in, first_stream, nstreams, (P.n % 32 == 0) ? d_pow32 : nullptr);
	return cudaGetLastError();
}
