/*
 * sts_generators.c - the CHAINED generators of the reference's tools/generators.c on the host: 3 (quadratic
 * congruential II), 6 (modular exponentiation), 7 (Blum-Blum-Shub as the tool computes it), 8 (Micali-Schnorr) and
 * 9 (SHA-1 based).  Each of them feeds its whole state through a non-linear map once per block, so block t cannot be
 * made before block t - 1 and nothing can jump ahead: one chain for the whole run, however many streams it is cut
 * into.  That is one thread of work; it belongs on a host core, in front of the engine's ordinary upload path, like
 * the reference's own flow (tool -> file -> sts).  The generators that CAN jump ahead - 1 (LCG), 2 (quadratic
 * congruential I), 4 (cubic congruential) and 5 (XOR) - are made on the GPU (csrc/k_lcg.cu, csrc/k_gen.cu).
 *
 * The output is the tool's FILE: write_sequence (generators.c:1603-1660) packs sequence bit 8 j + c into bit c of
 * byte j; copyBitsToEpsilon (:704-760) ends a stream in the middle of a block and drops the rest of that block.
 * The tool fixes n = 2^20 (:141); any n that is a multiple of 8 is taken here.
 *
 * The tool's big-number routines work on big-endian byte strings.  Where they are ordinary arithmetic (Mult, Mod,
 * ModExp: :1899-2200) this file uses 64-bit limbs; where the tool's result is NOT the arithmetic its names promise,
 * the byte-level behaviour is kept, because that is what the tool writes:
 *   - smult (:1968-1979) adds its carry twice and skips the multiplicand's top byte, and quadRes2 (:996-1002) reads
 *     its 64-byte result as part of a 65-byte number;
 *   - bbs (:1288) squares in place: the accumulator of Mult aliases both factors and is not cleared;
 *   - ModExp (:2012) sets only the last byte of its accumulator, which micali_schnorr never clears (:1380): every
 *     block starts from the previous result, and the first from uninitialised memory - the tool's generator 8 gives
 *     different output on every run.  Here that memory is defined as zero; sts_gen_set_ms_initial() sets it to what a
 *     given reference run had, which is how the tests reproduce such a run;
 *   - ahtopb (:2447) turns the 47-digit seed of micali_schnorr into a last byte of 0x30;
 *   - SHA1 (:1484, :1579) swaps bytes only #ifdef LITTLE_ENDIAN, which the tool's Makefile flags leave undefined:
 *     message and output words are little-endian, as built by the reference;
 *   - bbs prints progress lines into its own stdout data (:1300); those are not data and are not reproduced.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "sts_generators.h"

typedef unsigned __int128 u128;

/* ---- little-endian 64-bit limb arithmetic --------------------------------------------------------------------- */

static void
bn_from_hex(uint64_t *x, int limbs, const char *hex)
{
	memset(x, 0, (size_t) limbs * 8);
	const size_t len = strlen(hex);
	for (size_t i = 0; i < len; i++) {
		const char ch = hex[len - 1 - i];
		const uint64_t v = (ch >= '0' && ch <= '9') ? (uint64_t) (ch - '0') : (uint64_t) ((ch | 0x20) - 'a' + 10);
		x[i / 16] |= v << (4 * (i % 16));
	}
}

/* big-endian bytes <-> limbs */
static void
bn_to_bytes(const uint64_t *x, int limbs, uint8_t *out)
{
	for (int l = 0; l < limbs; l++)
		for (int b = 0; b < 8; b++)
			out[8 * (limbs - 1 - l) + (7 - b)] = (uint8_t) (x[l] >> (8 * b));
}

static void
bn_from_bytes(uint64_t *x, int limbs, const uint8_t *in)
{
	for (int l = 0; l < limbs; l++) {
		uint64_t v = 0;
		for (int b = 0; b < 8; b++)
			v = (v << 8) | in[8 * (limbs - 1 - l) + b];
		x[l] = v;
	}
}

/* r[0 .. na + nb) = a * b */
static void
bn_mul(uint64_t *r, const uint64_t *a, int na, const uint64_t *b, int nb)
{
	memset(r, 0, (size_t) (na + nb) * 8);
	for (int i = 0; i < na; i++) {
		uint64_t carry = 0;
		for (int j = 0; j < nb; j++) {
			const u128 t = (u128) a[i] * b[j] + r[i + j] + carry;
			r[i + j] = (uint64_t) t;
			carry = (uint64_t) (t >> 64);
		}
		r[i + nb] = carry;
	}
}

/* x[0 .. nx) mod m[0 .. nm), m's top limb non-zero: Knuth's algorithm D, remainder left in x[0 .. nm), upper limbs zeroed */
static void
bn_mod(uint64_t *x, int nx, const uint64_t *m, int nm)
{
	uint64_t v[32], u[66];
	const int s = __builtin_clzll(m[nm - 1]);
	for (int i = nm - 1; i > 0; i--)
		v[i] = s ? (m[i] << s) | (m[i - 1] >> (64 - s)) : m[i];
	v[0] = m[0] << s;
	u[nx] = s ? x[nx - 1] >> (64 - s) : 0;
	for (int i = nx - 1; i > 0; i--)
		u[i] = s ? (x[i] << s) | (x[i - 1] >> (64 - s)) : x[i];
	u[0] = x[0] << s;
	for (int j = nx - nm; j >= 0; j--) {
		const u128 num = ((u128) u[j + nm] << 64) | u[j + nm - 1];
		u128 qhat = num / v[nm - 1], rhat = num % v[nm - 1];
		while (qhat >> 64 || (nm > 1 && (u128) (uint64_t) qhat * v[nm - 2] > ((rhat << 64) | u[j + nm - 2]))) {
			qhat--;
			rhat += v[nm - 1];
			if (rhat >> 64)
				break;
		}
		uint64_t borrow = 0, carry = 0;
		for (int i = 0; i < nm; i++) {
			const u128 p = (u128) (uint64_t) qhat * v[i] + carry;
			carry = (uint64_t) (p >> 64);
			const u128 t = (u128) u[i + j] - (uint64_t) p - borrow;
			u[i + j] = (uint64_t) t;
			borrow = (uint64_t) (t >> 64) & 1;
		}
		const u128 t = (u128) u[j + nm] - carry - borrow;
		u[j + nm] = (uint64_t) t;
		if ((uint64_t) (t >> 64) & 1) {			/* one too many: add back */
			uint64_t c = 0;
			for (int i = 0; i < nm; i++) {
				const u128 a = (u128) u[i + j] + v[i] + c;
				u[i + j] = (uint64_t) a;
				c = (uint64_t) (a >> 64);
			}
			u[j + nm] += c;
		}
	}
	for (int i = 0; i < nm; i++)
		x[i] = s ? (u[i] >> s) | (u[i + 1] << (64 - s)) : u[i];
	for (int i = nm; i < nx; i++)
		x[i] = 0;
}

/* r = a * b mod m, all nm limbs */
static void
bn_mulmod(uint64_t *r, const uint64_t *a, const uint64_t *b, const uint64_t *m, int nm)
{
	uint64_t t[32];
	bn_mul(t, a, nm, b, nm);
	bn_mod(t, 2 * nm, m, nm);
	memcpy(r, t, (size_t) nm * 8);
}

/* ---- the stream writer ------------------------------------------------------------------------------------------ */

struct sts_gen {
	int generator;
	long n;
	/* the stream being filled */
	uint8_t *out;
	long bits_done;
	/* generator state */
	uint64_t g[8];			/* QCG2 state */
	uint64_t p512[8], g512[8], y[3];	/* MODEXP: modulus, base, 160-bit exponent */
	uint64_t n1024[16], msx[16], msy[16];	/* BBS / MS modulus; MS seed X and the accumulator left by the last block */
	uint8_t bbs[256];		/* BBS: the tool's 256-byte buffer, big-endian */
	uint8_t xkey[20];		/* SHA1 */
};

static const uint8_t REV8[16] = { 0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15 };

static inline uint8_t
rev8(uint8_t b)
{
	return (uint8_t) ((REV8[b & 15] << 4) | REV8[b >> 4]);
}

/* append the first `nbits` bits of the big-endian bytes `blk` (MSB of each byte first); true once the stream is full */
static int
emit_block(struct sts_gen *g, const uint8_t *blk, long nbits)
{
	for (long done = 0; done < nbits; done += 8) {
		long cnt = nbits - done < 8 ? nbits - done : 8;
		if (cnt > g->n - g->bits_done)
			cnt = g->n - g->bits_done;
		const unsigned v = rev8(blk[done / 8]) & ((1u << cnt) - 1u);	/* sequence order = ascending bit */
		const long pos = g->bits_done;
		g->out[pos >> 3] |= (uint8_t) (v << (pos & 7));
		if ((pos & 7) + cnt > 8)
			g->out[(pos >> 3) + 1] |= (uint8_t) (v >> (8 - (pos & 7)));
		g->bits_done += cnt;
		if (g->bits_done >= g->n)
			return 1;
	}
	return 0;
}

/* ---- 3: quadratic congruential II (generators.c:978-1046) ------------------------------------------------------ */

static void
qcg2_block(struct sts_gen *s, uint8_t blk[64])
{
	uint8_t c[64], a[72];
	bn_to_bytes(s->g, 8, c);
	memset(a, 0, sizeof(a));
	unsigned res = 0;
	for (int i = 63; i > 0; i--) {			/* smult, as written (:1975-1979) */
		res = (a[7 + i] + 2u * c[i] + (res >> 8)) & 0xFFFFu;
		a[7 + i] = (uint8_t) res;
		a[7 + i - 1] = (uint8_t) (res >> 8);
	}
	/* the 65 bytes a[7 .. 71] are t1 (a[71] = 0 is its last byte); + 3 cannot carry out of that zero byte */
	a[71] = 3;
	uint64_t t1[9], prod[17];
	bn_from_bytes(t1, 9, a);
	bn_mul(prod, t1, 9, s->g, 8);
	uint64_t carry = 1;				/* + 1 (:1001), then the low 64 bytes (:1002) */
	for (int i = 0; i < 8 && carry; i++) {
		prod[i] += carry;
		carry = (prod[i] == 0);
	}
	memcpy(s->g, prod, 64);
	bn_to_bytes(s->g, 8, blk);
}

/* ---- 6: modular exponentiation (:1192-1253) --------------------------------------------------------------------- */

static void
modexp_block(struct sts_gen *s, uint8_t blk[64])
{
	uint64_t x[8] = { 1 };
	int top = 159;
	while (top >= 0 && !((s->y[top / 64] >> (top % 64)) & 1))
		top--;
	for (int b = top; b >= 0; b--) {			/* ModExp :2014-2037 from the top set bit */
		bn_mulmod(x, x, x, s->p512, 8);
		if ((s->y[b / 64] >> (b % 64)) & 1)
			bn_mulmod(x, x, s->g512, s->p512, 8);
	}
	bn_to_bytes(x, 8, blk);
	s->y[0] = x[0];						/* y <- the last 20 bytes of x (:1226) */
	s->y[1] = x[1];
	s->y[2] = x[2] & 0xFFFFFFFFull;
}

/* ---- 7: Blum-Blum-Shub as the tool computes it (:1256-1343) ------------------------------------------------------ */

static inline uint64_t
load_be64(const uint8_t *p)
{
	uint64_t v;
	memcpy(&v, p, 8);
	return __builtin_bswap64(v);
}

static inline void
store_be64(uint8_t *p, uint64_t v)
{
	v = __builtin_bswap64(v);
	memcpy(p, &v, 8);
}

static int
bbs_bit(struct sts_gen *s)
{
	uint8_t *a = s->bbs;
	for (int i = 127; i >= 0; i--) {
		/* row i of Mult(A, A, 128, A, 128) (:1908-1917): the 129 bytes a[i .. i + 128] become a[i + 1 .. i + 128] + a[i] * C, where
		 * C is bytes 0 .. 127 AS THEY ARE NOW, rows 127 .. i + 1 having written over bytes i + 1 .. 127 of them */
		const uint64_t m = a[i];
		uint64_t carry = 0;
		uint64_t w[16];
		for (int l = 0; l < 16; l++) {
			const u128 t = (u128) m * load_be64(a + 8 * (15 - l)) + load_be64(a + i + 1 + 8 * (15 - l)) + carry;
			w[l] = (uint64_t) t;
			carry = (uint64_t) (t >> 64);
		}
		for (int l = 0; l < 16; l++)
			store_be64(a + i + 1 + 8 * (15 - l), w[l]);
		a[i] = (uint8_t) carry;
	}
	uint64_t x[32];
	bn_from_bytes(x, 32, a);
	bn_mod(x, 32, s->n1024, 16);
	bn_to_bytes(x, 16, a);				/* memcpy(x, x + 128, 128): the value ... */
	memcpy(a + 128, a, 128);			/* ... and the remainder Mod left in bytes 128 .. 255 */
	return a[127] & 1;
}

/* ---- 8: Micali-Schnorr (:1346-1425) -------------------------------------------------------------------------------- */

static void
ms_block(struct sts_gen *s, uint8_t tail[105])
{
	uint64_t a[16];
	memcpy(a, s->msy, sizeof(a));
	a[0] = (a[0] & ~0xFFull) | 1;			/* ModExp :2012: A[LM - 1] = 1 over whatever A held */
	static const int ebits[4] = { 1, 0, 1, 1 };	/* e = 11 */
	for (int b = 0; b < 4; b++) {
		bn_mulmod(a, a, a, s->n1024, 16);
		if (ebits[b])
			bn_mulmod(a, a, s->msx, s->n1024, 16);
	}
	memcpy(s->msy, a, sizeof(a));
	uint8_t y[128];
	bn_to_bytes(a, 16, y);
	memcpy(tail, y + 23, 105);			/* :1382-1385: the low 840 bits, shifted left by 3 */
	for (int i = 0; i < 105; i++)
		tail[i] = (uint8_t) ((tail[i] << 3) | (i + 1 < 105 ? tail[i + 1] >> 5 : 0));
	memset(s->msx, 0, sizeof(s->msx));		/* :1387-1392: the top 24 bytes shifted right by 5 */
	uint64_t hi[3] = { a[13], a[14], a[15] };
	s->msx[0] = (hi[0] >> 5) | (hi[1] << 59);
	s->msx[1] = (hi[1] >> 5) | (hi[2] << 59);
	s->msx[2] = hi[2] >> 5;
}

/* ---- 9: the SHA-1 based generator as built by the reference's Makefile (:1428-1600) ---------------------------- */

static inline uint32_t
rol32(uint32_t v, int s)
{
	return (v << s) | (v >> (32 - s));
}

static void
sha_block(struct sts_gen *s, uint8_t gout[20])
{
	uint32_t w[80];
	uint8_t m[64];
	memcpy(m, s->xkey, 20);
	memset(m + 20, 0, 44);
	for (int i = 0; i < 16; i++)				/* no byte reversal: LITTLE_ENDIAN is not defined in that build */
		w[i] = (uint32_t) m[4 * i] | ((uint32_t) m[4 * i + 1] << 8) | ((uint32_t) m[4 * i + 2] << 16) | ((uint32_t) m[4 * i + 3] << 24);
	for (int t = 16; t < 80; t++)
		w[t] = rol32(w[t - 3] ^ w[t - 8] ^ w[t - 14] ^ w[t - 16], 1);
	const uint32_t h[5] = { 0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u };
	uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
	for (int t = 0; t < 80; t++) {
		uint32_t f, k;
		if (t < 20) {
			f = (b & c) | (~b & d);
			k = 0x5A827999u;
		} else if (t < 40) {
			f = b ^ c ^ d;
			k = 0x6ED9EBA1u;
		} else if (t < 60) {
			f = (b & c) | (b & d) | (c & d);
			k = 0x8F1BBCDCu;
		} else {
			f = b ^ c ^ d;
			k = 0xCA62C1D6u;
		}
		const uint32_t tmp = rol32(a, 5) + f + e + k + w[t];
		e = d;
		d = c;
		c = rol32(b, 30);
		b = a;
		a = tmp;
	}
	const uint32_t r[5] = { a + h[0], b + h[1], c + h[2], d + h[3], e + h[4] };
	for (int i = 0; i < 5; i++)
		for (int j = 0; j < 4; j++)
			gout[4 * i + j] = (uint8_t) (r[i] >> (8 * j));
	/* Xkey <- Xkey + G + 1 on big-endian bytes (:1585-1586) */
	unsigned carry = 1;
	for (int i = 19; i >= 0; i--) {
		carry += (unsigned) s->xkey[i] + gout[i];
		s->xkey[i] = (uint8_t) carry;
		carry >>= 8;
	}
}

/* ---- the interface -------------------------------------------------------------------------------------------------- */

static const char P512_HEX[] = "987b6a6bf2c56a97291c445409920032499f9ee7ad128301b5d0254aa1a9633f"
	"dbd378d40149f1e23a13849f3d45992f5c4c6b7104099bc301f6005f9d8115e1";
static const char G1_HEX[] = "3844506a9456c564b8b8538e0cc15aff46c95e69600f084f0657c2401b3c2447"
	"34b62ea9bb95be4923b9b7e84eeaf1a224894ef0328d44bc3eb3e983644da3f5";
static const char G2_HEX[] = "7844506a9456c564b8b8538e0cc15aff46c95e69600f084f0657c2401b3c2447"
	"34b62ea9bb95be4923b9b7e84eeaf1a224894ef0328d44bc3eb3e983644da3f5";
static const char BBS_P_HEX[] = "E65097BAEC92E70478CAF4ED0ED94E1C94B154466BFB9EC9BE37B2B0FF8526C2"
	"22B76E0E915017535AE8B9207250257D0A0C87C0DACEF78E17D1EF9DC44FD91F";
static const char BBS_Q_HEX[] = "E029AEFCF8EA2C29D99CB53DD5FA9BC1D0176F5DF8D9110FD16EE21F32E37BA8"
	"6FF42F00531AD5B8A43073182CC2E15F5C86E8DA059E346777C9A985F7D8A867";
static const char BBS_S_HEX[] = "10d6333cfac8e30e808d2192f7c0439480da79db9bbca1667d73be9a677ed313"
	"11f3b830937763837cb7b1b1dc75f14eea417f84d9625628750de99e7ef1e976";

int
sts_gen_is_chained(int generator)
{
	return generator == 3 || (generator >= 6 && generator <= 9);
}

struct sts_gen *
sts_gen_open(int generator, long n)
{
	if (!sts_gen_is_chained(generator) || n <= 0 || n % 8 != 0)
		return NULL;
	struct sts_gen *s = calloc(1, sizeof(*s));
	if (s == NULL)
		return NULL;
	s->generator = generator;
	s->n = n;
	uint64_t p[8], q[8];
	bn_from_hex(p, 8, BBS_P_HEX);
	bn_from_hex(q, 8, BBS_Q_HEX);
	bn_mul(s->n1024, p, 8, q, 8);
	switch (generator) {
	case 3:
		bn_from_hex(s->g, 8, G2_HEX);
		break;
	case 6:
		bn_from_hex(s->p512, 8, P512_HEX);
		bn_from_hex(s->g512, 8, G1_HEX);
		bn_from_hex(s->y, 3, "7AB36982CE1ADF832019CDFEB2393CABDF0214EC");
		break;
	case 7: {
		uint64_t sd[8], x[16];
		bn_from_hex(sd, 8, BBS_S_HEX);
		bn_mul(x, sd, 8, sd, 8);			/* :1286: a true square, into a cleared buffer */
		bn_mod(x, 16, s->n1024, 16);
		bn_to_bytes(x, 16, s->bbs);
		break;
	}
	case 8:
		/* 23 bytes of digits and the byte ahtopb makes of the last digit and the string's NUL (6 << 4) + (0 - '0') = 0x30 */
		bn_from_hex(s->msx, 16, "237c5f791c2cfe47bfb16d2d54a0d60665b20904ec822a30");
		break;
	case 9: {
		uint64_t k[3];
		bn_from_hex(k, 3, "ec822a619d6ed5d9492218a7a4c5b15d57c61601");
		uint8_t b[24];
		bn_to_bytes(k, 3, b);
		memcpy(s->xkey, b + 4, 20);
		break;
	}
	}
	return s;
}

void
sts_gen_set_ms_initial(struct sts_gen *s, const uint8_t y[128])
{
	bn_from_bytes(s->msy, 16, y);
}

void
sts_gen_close(struct sts_gen *s)
{
	free(s);
}

int
sts_gen_next(struct sts_gen *s, uint8_t *raw, long nstreams)
{
	if (s == NULL || raw == NULL || nstreams < 0)
		return -1;
	const long nb = s->n / 8;
	uint8_t blk[128];
	for (long k = 0; k < nstreams; k++) {
		s->out = raw + k * nb;
		s->bits_done = 0;
		memset(s->out, 0, (size_t) nb);
		int full = 0;
		while (!full) {
			switch (s->generator) {
			case 3:
				qcg2_block(s, blk);
				full = emit_block(s, blk, 512);
				break;
			case 6:
				modexp_block(s, blk);
				full = emit_block(s, blk, 512);
				break;
			case 7:
				blk[0] = (uint8_t) (bbs_bit(s) << 7);
				full = emit_block(s, blk, 1);
				break;
			case 8:
				ms_block(s, blk);
				full = emit_block(s, blk, 837);
				break;
			default:
				sha_block(s, blk);
				full = emit_block(s, blk, 160);
				break;
			}
		}
	}
	return 0;
}
