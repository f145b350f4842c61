/*
 * k_gen.cu - synthetic input: the generators of the reference's tools/generators.c that can JUMP AHEAD, made directly
 * in the engine's device input buffer, any stream from any rank (generator 1, the LCG, is k_lcg.cu):
 *
 *   2  quadratic congruential I (generators.c:915-975)   x <- x^2 mod p, p a 512-bit prime; the 512 bits of x are a block.
 *      x_J = g^(2^J mod (p - 1)) mod p.  p - 1 = 2^s m with m odd, so 2^J mod (p - 1) = 2^s (2^(J - s) mod m): one
 *      64-step power of two modulo m and one 512-bit exponentiation modulo p put a thread at block J.
 *   4  cubic congruential (:1049-1108)   x <- x^3 mod 2^512; the block is the TOP 512 bits of the 1536-bit cube (:1075-1078),
 *      the state its low 512.  x_J = g^(3^J) mod 2^512 (g is odd; any exponent congruent modulo 2^510 gives the same
 *      power, so 3^J is simply kept modulo 2^512).
 *   5  XOR (:1111-1186)   b_i = b_(i-1) ^ b_(i-127), ONE sequence that the streams are cut from.  The tool seeds its ring
 *      with the characters '0'/'1' and stores new bits as numbers, so bits 128..253 come out as ones (:1141-1154);
 *      from bit 254 on the sequence is the linear recurrence with characteristic polynomial x^127 + x^126 + 1 started
 *      from bits 127..253.  With u_t = b_(127 + t) and c(x) = x^T mod P:  u_(T + j) = sum_i c_i u_(i + j).
 *
 * The other five generators (3, 6, 7, 8, 9) push their whole state through a non-linear map per block - one chain for
 * the whole run, nothing to jump with - and are made on the host (host/sts_generators.c).
 *
 * Output convention as in k_lcg.cu: the FILE the tool writes, sequence bit 8 j + c in bit c of byte j
 * (write_sequence :1640-1660), so a little-endian word holds sequence bit 32 t + i in bit i.  A stream starts a new
 * block and drops the unused end of its last one (copyBitsToEpsilon :704-760); generator 5 has no blocks.
 *
 * Work split: one thread per SEGMENT of a stream (GEN_SEGS per stream for 2 and 4, 1024 words for 5), each with its own
 * jump.  A jump costs as much as about 550 blocks of generator 2 (140 of generator 4), so more segments mean more
 * threads and more total work; the device is otherwise idle while a batch is made, so the split is chosen for time:
 * measured per 1024 streams of 2^20 bits with 4 segments 6.3 ms (2) and 3.5 ms (4), generator 5 0.29 ms, the LCG 1.4 ms.
 */
#include <string.h>
#include "common.cuh"
#include "engine.h"

#define GL 16				/* 32-bit limbs of a 512-bit number, least significant first */
#define GEN_THREADS 32
#define GEN_SEGS 16
#define XOR_SEG_WORDS 1024

struct GenConsts {
	uint32_t p[GL], p_ninv;		/* generator 2: modulus, -1/p mod 2^32 */
	uint32_t p_one[GL];		/* R mod p, R = 2^512: 1 in Montgomery form */
	uint32_t p_g[GL];		/* the seed in Montgomery form */
	uint32_t m[GL], m_ninv;		/* the odd part of p - 1 */
	uint32_t m_one[GL];
	uint32_t s;			/* p - 1 = 2^s m */
	uint32_t ccg_g[GL];		/* generator 4: the seed */
	uint32_t xor_u[9];		/* generator 5: u_0 .. u_253 (bit t of the array = u_t), one spare word */
	uint32_t xor_head[8];		/* b_0 .. b_255 */
};

/* ---- 512-bit Montgomery arithmetic, everything unrolled into registers ----------------------------------------- */

/* r = a b / R mod n (CIOS); a, b < n */
__device__ __forceinline__ void mont_mul(uint32_t (&r)[GL], const uint32_t (&a)[GL], const uint32_t (&b)[GL],
					 const uint32_t *__restrict__ n, uint32_t ninv)
{
	uint32_t t[GL + 2];
#pragma unroll
	for (int j = 0; j < GL + 2; j++)
		t[j] = 0;
#pragma unroll
	for (int i = 0; i < GL; i++) {
		uint64_t c = 0;
#pragma unroll
		for (int j = 0; j < GL; j++) {
			c += (uint64_t) a[j] * b[i] + t[j];
			t[j] = (uint32_t) c;
			c >>= 32;
		}
		c += t[GL];
		t[GL] = (uint32_t) c;
		t[GL + 1] = (uint32_t) (c >> 32);
		const uint32_t q = t[0] * ninv;
		c = (uint64_t) q * n[0] + t[0];
		c >>= 32;
#pragma unroll
		for (int j = 1; j < GL; j++) {
			c += (uint64_t) q * n[j] + t[j];
			t[j - 1] = (uint32_t) c;
			c >>= 32;
		}
		c += t[GL];
		t[GL - 1] = (uint32_t) c;
		t[GL] = t[GL + 1] + (uint32_t) (c >> 32);
	}
	uint32_t d[GL];
	uint32_t borrow = 0;
#pragma unroll
	for (int j = 0; j < GL; j++) {
		const uint64_t v = (uint64_t) t[j] - n[j] - borrow;
		d[j] = (uint32_t) v;
		borrow = (uint32_t) (v >> 32) & 1u;
	}
	const bool sub = t[GL] != 0 || borrow == 0;
#pragma unroll
	for (int j = 0; j < GL; j++)
		r[j] = sub ? d[j] : t[j];
}

/* r = a / R mod n: out of Montgomery form */
__device__ __forceinline__ void mont_redc(uint32_t (&r)[GL], const uint32_t (&a)[GL], const uint32_t *__restrict__ n, uint32_t ninv)
{
	uint32_t t[GL + 1];
#pragma unroll
	for (int j = 0; j < GL; j++)
		t[j] = a[j];
	t[GL] = 0;
#pragma unroll
	for (int i = 0; i < GL; i++) {
		const uint32_t q = t[0] * ninv;
		uint64_t c = (uint64_t) q * n[0] + t[0];
		c >>= 32;
#pragma unroll
		for (int j = 1; j < GL; j++) {
			c += (uint64_t) q * n[j] + t[j];
			t[j - 1] = (uint32_t) c;
			c >>= 32;
		}
		c += t[GL];
		t[GL - 1] = (uint32_t) c;
		t[GL] = (uint32_t) (c >> 32);
	}
	uint32_t d[GL];
	uint32_t borrow = 0;
#pragma unroll
	for (int j = 0; j < GL; j++) {
		const uint64_t v = (uint64_t) t[j] - n[j] - borrow;
		d[j] = (uint32_t) v;
		borrow = (uint32_t) (v >> 32) & 1u;
	}
	const bool sub = t[GL] != 0 || borrow == 0;
#pragma unroll
	for (int j = 0; j < GL; j++)
		r[j] = sub ? d[j] : t[j];
}

/* r = 2 r mod n, r < n < 2^511 */
__device__ __forceinline__ void mod_double(uint32_t (&r)[GL], const uint32_t *__restrict__ n)
{
	uint32_t t[GL], d[GL];
#pragma unroll
	for (int j = GL - 1; j > 0; j--)
		t[j] = (r[j] << 1) | (r[j - 1] >> 31);
	t[0] = r[0] << 1;
	uint32_t borrow = 0;
#pragma unroll
	for (int j = 0; j < GL; j++) {
		const uint64_t v = (uint64_t) t[j] - n[j] - borrow;
		d[j] = (uint32_t) v;
		borrow = (uint32_t) (v >> 32) & 1u;
	}
#pragma unroll
	for (int j = 0; j < GL; j++)
		r[j] = borrow ? t[j] : d[j];
}

/* the top bit of e, and e shifted left by one: the exponent loops read their bits with static register indices */
__device__ __forceinline__ uint32_t take_top_bit(uint32_t (&e)[GL])
{
	const uint32_t top = e[GL - 1] >> 31;
#pragma unroll
	for (int j = GL - 1; j > 0; j--)
		e[j] = (e[j] << 1) | (e[j - 1] >> 31);
	e[0] <<= 1;
	return top;
}

/* one 512-bit block (value v, least significant limb first) as words 16 b .. 16 b + 15 of a stream: the tool emits the
 * most significant byte first and each byte MSB first, the file holds sequence bits LSB first: word t = brev(limb 15 - t) */
__device__ __forceinline__ void store_block(uint32_t *__restrict__ row, int64_t b, const uint32_t (&v)[GL], const DevParams &P)
{
	const int tail = (int) (P.n & 31);
#pragma unroll
	for (int t = 0; t < GL; t++) {
		const int64_t wi = 16 * b + t;
		if (wi < P.nwords) {
			uint32_t w = __brev(v[GL - 1 - t]);
			if (tail && wi == P.nwords - 1)
				w &= (1u << tail) - 1u;
			row[wi] = w;
		}
	}
}

/* ---- 2: quadratic congruential I -------------------------------------------------------------------------------- */

__global__ void __launch_bounds__(GEN_THREADS)
qcg1_kernel(DevParams P, uint32_t *__restrict__ out, int64_t first_stream, int64_t nstreams, const GenConsts *__restrict__ K, int segs)
{
	__shared__ GenConsts C;
	for (int i = threadIdx.x; i < (int) (sizeof(GenConsts) / 4); i += GEN_THREADS)
		((uint32_t *) &C)[i] = ((const uint32_t *) K)[i];
	__syncthreads();
	const int64_t gid = (int64_t) blockIdx.x * GEN_THREADS + threadIdx.x;
	if (gid >= nstreams * segs)
		return;
	const int64_t s = gid / segs;
	const int q = (int) (gid % segs);
	const int64_t blocks = (P.n + 511) / 512;
	const int64_t per = (blocks + segs - 1) / segs;
	const int64_t b0 = q * per, b1 = min(blocks, b0 + per);
	if (b0 >= b1)
		return;
	const uint64_t J = (uint64_t) (first_stream + s) * (uint64_t) blocks + (uint64_t) b0;	/* state index before block b0 */

	/* e = 2^J mod (p - 1) */
	uint32_t e[GL];
	if (J < 480) {
#pragma unroll
		for (int j = 0; j < GL; j++)
			e[j] = ((int) (J >> 5) == j) ? (1u << (J & 31)) : 0u;
	} else {
		uint32_t r[GL];
#pragma unroll
		for (int j = 0; j < GL; j++)
			r[j] = C.m_one[j];
		const uint64_t ex = J - C.s;
		for (int bit = 63; bit >= 0; bit--) {
			mont_mul(r, r, r, C.m, C.m_ninv);
			if ((ex >> bit) & 1ull)
				mod_double(r, C.m);
		}
		mont_redc(r, r, C.m, C.m_ninv);
		const uint32_t sh = C.s;			/* 0 < s < 32 */
#pragma unroll
		for (int j = GL - 1; j > 0; j--)
			e[j] = (r[j] << sh) | (r[j - 1] >> (32 - sh));
		e[0] = r[0] << sh;
	}
	/* x = g^e mod p, in Montgomery form */
	uint32_t x[GL], g[GL];
#pragma unroll
	for (int j = 0; j < GL; j++) {
		x[j] = C.p_one[j];
		g[j] = C.p_g[j];
	}
	for (int bit = 0; bit < 512; bit++) {
		mont_mul(x, x, x, C.p, C.p_ninv);
		if (take_top_bit(e))
			mont_mul(x, x, g, C.p, C.p_ninv);
	}
	uint32_t *__restrict__ row = out + s * P.pitch_words;
	for (int64_t b = b0; b < b1; b++) {
		mont_mul(x, x, x, C.p, C.p_ninv);
		uint32_t v[GL];
		mont_redc(v, x, C.p, C.p_ninv);
		store_block(row, b, v, P);
	}
}

/* ---- 4: cubic congruential ---------------------------------------------------------------------------------------- */

/* r = a b mod 2^512 */
__device__ __forceinline__ void mul_low(uint32_t (&r)[GL], const uint32_t (&a)[GL], const uint32_t (&b)[GL])
{
	uint32_t t[GL];
#pragma unroll
	for (int j = 0; j < GL; j++)
		t[j] = 0;
#pragma unroll
	for (int i = 0; i < GL; i++) {
		uint64_t c = 0;
#pragma unroll
		for (int j = 0; j < GL - i; j++) {
			c += (uint64_t) a[j] * b[i] + t[i + j];
			t[i + j] = (uint32_t) c;
			c >>= 32;
		}
	}
#pragma unroll
	for (int j = 0; j < GL; j++)
		r[j] = t[j];
}

__global__ void __launch_bounds__(GEN_THREADS)
ccg_kernel(DevParams P, uint32_t *__restrict__ out, int64_t first_stream, int64_t nstreams, const GenConsts *__restrict__ K, int segs)
{
	const int64_t gid = (int64_t) blockIdx.x * GEN_THREADS + threadIdx.x;
	if (gid >= nstreams * segs)
		return;
	const int64_t s = gid / segs;
	const int q = (int) (gid % segs);
	const int64_t blocks = (P.n + 511) / 512;
	const int64_t per = (blocks + segs - 1) / segs;
	const int64_t b0 = q * per, b1 = min(blocks, b0 + per);
	if (b0 >= b1)
		return;
	const uint64_t J = (uint64_t) (first_stream + s) * (uint64_t) blocks + (uint64_t) b0;
	uint32_t g[GL], e[GL], x[GL];
#pragma unroll
	for (int j = 0; j < GL; j++) {
		g[j] = __ldg(&K->ccg_g[j]);
		e[j] = (j == 0) ? 1u : 0u;
		x[j] = (j == 0) ? 1u : 0u;
	}
	/* e = 3^J mod 2^512 */
	for (int bit = 63; bit >= 0; bit--) {
		mul_low(e, e, e);
		if ((J >> bit) & 1ull) {
			uint64_t c = 0;
#pragma unroll
			for (int j = 0; j < GL; j++) {
				c += 3ull * e[j];
				e[j] = (uint32_t) c;
				c >>= 32;
			}
		}
	}
	/* x = g^e mod 2^512 */
	for (int bit = 0; bit < 512; bit++) {
		mul_low(x, x, x);
		if (take_top_bit(e))
			mul_low(x, x, g);
	}
	uint32_t *__restrict__ row = out + s * P.pitch_words;
	for (int64_t b = b0; b < b1; b++) {
		/* the full cube: sq = x^2 (32 limbs), cu = sq x (48 limbs) */
		uint32_t sq[2 * GL];
#pragma unroll
		for (int j = 0; j < 2 * GL; j++)
			sq[j] = 0;
#pragma unroll
		for (int i = 0; i < GL; i++) {
			uint64_t c = 0;
#pragma unroll
			for (int j = 0; j < GL; j++) {
				c += (uint64_t) x[j] * x[i] + sq[i + j];
				sq[i + j] = (uint32_t) c;
				c >>= 32;
			}
			sq[i + GL] = (uint32_t) c;
		}
		uint32_t cu[3 * GL];
#pragma unroll
		for (int j = 0; j < 3 * GL; j++)
			cu[j] = 0;
#pragma unroll
		for (int i = 0; i < GL; i++) {
			uint64_t c = 0;
#pragma unroll
			for (int j = 0; j < 2 * GL; j++) {
				c += (uint64_t) sq[j] * x[i] + cu[i + j];
				cu[i + j] = (uint32_t) c;
				c >>= 32;
			}
			cu[i + 2 * GL] = (uint32_t) c;
		}
		uint32_t v[GL];
#pragma unroll
		for (int j = 0; j < GL; j++) {
			x[j] = cu[j];
			v[j] = cu[2 * GL + j];
		}
		store_block(row, b, v, P);
	}
}

/* ---- 5: XOR ---------------------------------------------------------------------------------------------------------- */

struct Poly { uint32_t w[4]; };			/* GF(2)[x], degree < 128: bit i of the 128 bits = coefficient of x^i */

__device__ __forceinline__ Poly poly_shr(const Poly &a, int k)	/* k in {1, 2, 4, 8, 16} */
{
	Poly r;
	r.w[0] = __funnelshift_r(a.w[0], a.w[1], k);
	r.w[1] = __funnelshift_r(a.w[1], a.w[2], k);
	r.w[2] = __funnelshift_r(a.w[2], a.w[3], k);
	r.w[3] = a.w[3] >> k;
	return r;
}

/* (lo + x^127 h) mod (x^127 + x^126 + 1) for h of degree < 127:  x^127 h = h + x^126 h_0 + x^127 (h >> 1), so the
 * sum runs over all right shifts of h - a suffix-XOR scan - plus x^126 times the parity of h */
__device__ __forceinline__ Poly poly_fold(Poly lo, Poly h)
{
	const uint32_t par = (uint32_t) __popc(h.w[0] ^ h.w[1] ^ h.w[2] ^ h.w[3]) & 1u;
	Poly y = h, t;
	t = poly_shr(y, 1);  y.w[0] ^= t.w[0]; y.w[1] ^= t.w[1]; y.w[2] ^= t.w[2]; y.w[3] ^= t.w[3];
	t = poly_shr(y, 2);  y.w[0] ^= t.w[0]; y.w[1] ^= t.w[1]; y.w[2] ^= t.w[2]; y.w[3] ^= t.w[3];
	t = poly_shr(y, 4);  y.w[0] ^= t.w[0]; y.w[1] ^= t.w[1]; y.w[2] ^= t.w[2]; y.w[3] ^= t.w[3];
	t = poly_shr(y, 8);  y.w[0] ^= t.w[0]; y.w[1] ^= t.w[1]; y.w[2] ^= t.w[2]; y.w[3] ^= t.w[3];
	t = poly_shr(y, 16); y.w[0] ^= t.w[0]; y.w[1] ^= t.w[1]; y.w[2] ^= t.w[2]; y.w[3] ^= t.w[3];
	y.w[0] ^= y.w[1];   y.w[1] ^= y.w[2];   y.w[2] ^= y.w[3];			/* >> 32 */
	y.w[0] ^= y.w[2];   y.w[1] ^= y.w[3];					/* >> 64 */
	lo.w[0] ^= y.w[0];
	lo.w[1] ^= y.w[1];
	lo.w[2] ^= y.w[2];
	lo.w[3] ^= y.w[3] ^ (par << 30);
	return lo;
}

/* the 16 low bits of v spread to the even bit positions */
__device__ __forceinline__ uint32_t spread16(uint32_t v)
{
	v &= 0xFFFFu;
	v = (v | (v << 8)) & 0x00FF00FFu;
	v = (v | (v << 4)) & 0x0F0F0F0Fu;
	v = (v | (v << 2)) & 0x33333333u;
	v = (v | (v << 1)) & 0x55555555u;
	return v;
}

__device__ __forceinline__ Poly poly_square(const Poly &a)	/* a of degree < 127 */
{
	uint32_t s[8];
#pragma unroll
	for (int i = 0; i < 4; i++) {
		s[2 * i] = spread16(a.w[i]);
		s[2 * i + 1] = spread16(a.w[i] >> 16);
	}
	/* bits 0 .. 126 stay, bits 127 .. 252 are h */
	Poly lo, h;
	lo.w[0] = s[0]; lo.w[1] = s[1]; lo.w[2] = s[2]; lo.w[3] = s[3] & 0x7FFFFFFFu;
	h.w[0] = __funnelshift_r(s[3], s[4], 31);
	h.w[1] = __funnelshift_r(s[4], s[5], 31);
	h.w[2] = __funnelshift_r(s[5], s[6], 31);
	h.w[3] = __funnelshift_r(s[6], s[7], 31);
	return poly_fold(lo, h);
}

__device__ __forceinline__ Poly poly_mulx(const Poly &a)
{
	Poly r;
	r.w[3] = __funnelshift_l(a.w[2], a.w[3], 1);
	r.w[2] = __funnelshift_l(a.w[1], a.w[2], 1);
	r.w[1] = __funnelshift_l(a.w[0], a.w[1], 1);
	r.w[0] = a.w[0] << 1;
	if (r.w[3] >> 31) {				/* x^127 = x^126 + 1 */
		r.w[3] ^= 0xC0000000u;
		r.w[0] ^= 1u;
	}
	return r;
}

__device__ __forceinline__ uint32_t prefix_xor32(uint32_t v)
{
	v ^= v << 1;
	v ^= v << 2;
	v ^= v << 4;
	v ^= v << 8;
	v ^= v << 16;
	return v;
}

__global__ void __launch_bounds__(128)
xor_kernel(DevParams P, uint32_t *__restrict__ out, int64_t first_stream, int64_t nstreams, const GenConsts *__restrict__ K)
{
	__shared__ uint32_t ub[9];
	if (threadIdx.x < 9)
		ub[threadIdx.x] = K->xor_u[threadIdx.x];
	__syncthreads();
	const int64_t segs = (P.nwords + XOR_SEG_WORDS - 1) / XOR_SEG_WORDS;
	const int64_t gid = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (gid >= nstreams * segs)
		return;
	const int64_t s = gid / segs, q = gid % segs;
	const int64_t w0 = q * XOR_SEG_WORDS, w1 = min(P.nwords, w0 + XOR_SEG_WORDS);
	uint32_t *__restrict__ row = out + s * P.pitch_words;
	const uint64_t G = (uint64_t) (first_stream + s) * (uint64_t) P.n + 32ull * (uint64_t) w0;	/* first bit of the segment */
	const int tail = (int) (P.n & 31);
	uint32_t h0, h1, h2, h3;			/* the last four words: h3 the newest */
	int64_t w = w0;
	if (G < 256) {					/* only the very first segment (n >= 1024): the tool's irregular start */
		uint32_t hd[8];
#pragma unroll
		for (int i = 0; i < 8; i++)
			hd[i] = __ldg(&K->xor_head[i]);
#pragma unroll
		for (int i = 0; i < 8; i++)
			if (w + i < w1)
				row[w + i] = hd[i];
		h0 = hd[4]; h1 = hd[5]; h2 = hd[6]; h3 = hd[7];
		w += 8;
	} else {
		/* c = x^T mod P, T = G - 127 */
		const uint64_t T = G - 127;
		Poly c;
		c.w[0] = 1u; c.w[1] = 0u; c.w[2] = 0u; c.w[3] = 0u;
		for (int bit = 63; bit >= 0; bit--) {
			c = poly_square(c);
			if ((T >> bit) & 1ull)
				c = poly_mulx(c);
		}
		/* bit j of the segment's first 128 bits: u_(T + j) = parity(c & (u_j .. u_(j + 126))) */
		uint32_t first[4];
#pragma unroll
		for (int k = 0; k < 4; k++)
			first[k] = 0;
		for (int j = 0; j < 128; j++) {
			const int wj = j >> 5, sj = j & 31;
			uint32_t acc = 0;
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const uint32_t lo = ub[wj + k], hi = ub[wj + k + 1];
				acc ^= c.w[k] & __funnelshift_r(lo, hi, sj);
			}
			const uint32_t bit = (uint32_t) __popc(acc) & 1u;
			if (wj == 0) first[0] |= bit << sj;
			else if (wj == 1) first[1] |= bit << sj;
			else if (wj == 2) first[2] |= bit << sj;
			else first[3] |= bit << sj;
		}
		h0 = first[0]; h1 = first[1]; h2 = first[2]; h3 = first[3];
#pragma unroll
		for (int i = 0; i < 4; i++) {
			if (w + i < w1) {
				uint32_t v = first[i];
				if (tail && w + i == P.nwords - 1)
					v &= (1u << tail) - 1u;
				row[w + i] = v;
			}
		}
		w += 4;
	}
	/* b_i = b_(i-1) ^ b_(i-127): a word is the prefix-XOR of the word 127 bits back, inverted when the previous bit is 1 */
	for (; w < w1; w++) {
		uint32_t v = prefix_xor32(__funnelshift_r(h0, h1, 1));
		if (h3 >> 31)
			v = ~v;
		h0 = h1; h1 = h2; h2 = h3; h3 = v;
		if (tail && w == P.nwords - 1)
			v &= (1u << tail) - 1u;
		row[w] = v;
	}
}

/* ---- host side: the constants, made once per context ------------------------------------------------------------ */

static void hex_to_limbs(uint32_t *x, const char *hex)
{
	memset(x, 0, GL * 4);
	const size_t len = strlen(hex);
	for (size_t i = 0; i < len; i++) {
		const char ch = hex[len - 1 - i];
		const uint32_t v = (ch >= '0' && ch <= '9') ? (uint32_t) (ch - '0') : (uint32_t) ((ch | 0x20) - 'a' + 10);
		x[i / 8] |= v << (4 * (i % 8));
	}
}

static int limbs_ge(const uint32_t *a, const uint32_t *b)
{
	for (int j = GL - 1; j >= 0; j--)
		if (a[j] != b[j])
			return a[j] > b[j];
	return 1;
}

/* r = 2 r mod n on the host; n may use all 512 bits */
static void host_double_mod(uint32_t *r, const uint32_t *n)
{
	const uint32_t top = r[GL - 1] >> 31;
	for (int j = GL - 1; j > 0; j--)
		r[j] = (r[j] << 1) | (r[j - 1] >> 31);
	r[0] <<= 1;
	if (top || limbs_ge(r, n)) {
		uint32_t borrow = 0;
		for (int j = 0; j < GL; j++) {
			const uint64_t v = (uint64_t) r[j] - n[j] - borrow;
			r[j] = (uint32_t) v;
			borrow = (uint32_t) (v >> 32) & 1u;
		}
	}
}

static uint32_t neg_inverse32(uint32_t n0)
{
	uint32_t x = n0;				/* Newton: 3 correct bits, doubled four times */
	for (int i = 0; i < 5; i++)
		x *= 2u - n0 * x;
	return 0u - x;
}

void gen_consts_host(void *dst)
{
	GenConsts *K = (GenConsts *) dst;
	memset(K, 0, sizeof(*K));
	hex_to_limbs(K->p, "987b6a6bf2c56a97291c445409920032499f9ee7ad128301b5d0254aa1a9633f"
			   "dbd378d40149f1e23a13849f3d45992f5c4c6b7104099bc301f6005f9d8115e1");		/* generators.c:942 */
	K->p_ninv = neg_inverse32(K->p[0]);
	uint32_t g[GL];
	hex_to_limbs(g, "3844506a9456c564b8b8538e0cc15aff46c95e69600f084f0657c2401b3c2447"
			"34b62ea9bb95be4923b9b7e84eeaf1a224894ef0328d44bc3eb3e983644da3f5");			/* :945 */
	memset(K->p_one, 0, sizeof(K->p_one));
	K->p_one[0] = 1;
	memcpy(K->p_g, g, sizeof(g));
	for (int i = 0; i < 512; i++) {			/* times R = 2^512 */
		host_double_mod(K->p_one, K->p);
		host_double_mod(K->p_g, K->p);
	}
	uint32_t pm1[GL];
	memcpy(pm1, K->p, sizeof(pm1));
	pm1[0] -= 1;					/* p is odd */
	int s = 0;
	while (!((pm1[0] >> s) & 1u))
		s++;					/* p - 1 = ...15e0: s = 5 */
	K->s = (uint32_t) s;
	for (int j = 0; j < GL; j++)
		K->m[j] = (pm1[j] >> s) | (j + 1 < GL ? pm1[j + 1] << (32 - s) : 0u);
	K->m_ninv = neg_inverse32(K->m[0]);
	K->m_one[0] = 1;
	for (int i = 0; i < 512; i++)
		host_double_mod(K->m_one, K->m);
	hex_to_limbs(K->ccg_g, "7844506a9456c564b8b8538e0cc15aff46c95e69600f084f0657c2401b3c2447"
			       "34b62ea9bb95be4923b9b7e84eeaf1a224894ef0328d44bc3eb3e983644da3f5");		/* :1070 */
	/* generator 5: the tool's first 508 bits (:1134-1154) */
	static const char seed[] = "0001011011011001000101111001001010011011101101000100000010101111"
				   "111010100100001010110110000000000100110000101110011111111100111";
	unsigned char b[512];
	for (int i = 0; i < 127; i++)
		b[i] = (unsigned char) (seed[i] - '0');
	b[127] = b[126] ^ b[0];				/* the one comparison of two characters */
	for (int i = 128; i < 254; i++)
		b[i] = 1;				/* a number against a character: always different */
	for (int i = 254; i < 512; i++)
		b[i] = b[i - 1] ^ b[i - 127];
	for (int i = 0; i < 256; i++)
		K->xor_head[i >> 5] |= (uint32_t) b[i] << (i & 31);
	for (int t = 0; t < 288; t++)
		K->xor_u[t >> 5] |= (uint32_t) b[127 + t] << (t & 31);
}

size_t gen_consts_bytes(void)
{
	return sizeof(GenConsts);
}

cudaError_t launch_generator(int generator, const DevParams &P, uint32_t *d_in, int64_t first_stream, int64_t nstreams,
			     const void *d_consts, cudaStream_t stream)
{
	const GenConsts *K = (const GenConsts *) d_consts;
	if (generator == 2 || generator == 4) {
		const int64_t blocks = (P.n + 511) / 512;
		const int segs = (int) (blocks < GEN_SEGS ? blocks : GEN_SEGS);
		const unsigned grid = (unsigned) ((nstreams * segs + GEN_THREADS - 1) / GEN_THREADS);
		if (generator == 2)
			qcg1_kernel<<<grid, GEN_THREADS, 0, stream>>>(P, d_in, first_stream, nstreams, K, segs);
		else
			ccg_kernel<<<grid, GEN_THREADS, 0, stream>>>(P, d_in, first_stream, nstreams, K, segs);
	} else {
		const int64_t segs = (P.nwords + XOR_SEG_WORDS - 1) / XOR_SEG_WORDS;
		const unsigned grid = (unsigned) ((nstreams * segs + 127) / 128);
		xor_kernel<<<grid, 128, 0, stream>>>(P, d_in, first_stream, nstreams, K);
	}
	return cudaGetLastError();
}
