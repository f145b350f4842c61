/*
 * k_ascii.cu - "-F a" input: ASCII '0' / '1' characters packed into the engine's stream rows on the device
 * (replaces parseBitsASCIIInput, utilities.c:1651-1727, which costs one fscanf("%1d") per bit).
 *
 * The reference's rule, kept exactly: iteration k seeks to character offset base_seek + k n of the file
 * (utilities.c:1682-1683) and takes the next n digits, skipping white space - so with line breaks in the file
 * consecutive streams overlap by the number of white-space characters.  One CTA per stream walks the text
 * from its own offset in rounds of 4096 characters: every thread compacts the digits of its 16 characters, a block
 * scan places them, the bits are OR-ed MSB first into a shared-memory ring of words and whole words are stored
 * byte-swapped (the engine's rows hold the raw file bytes of the "-F r" format).
 *
 * status[0] = lowest stream index that ran out of text before n digits, status[1] = lowest text offset of a
 * character that is neither a digit 0/1 nor white space and that the reference would have consumed.
 */
#include "common.cuh"

#define AP_THREADS 256
#define AP_CPT 16			/* characters per thread and round */
#define AP_RING 512			/* words; two rounds (2 x 129) never wrap */

__global__ void __launch_bounds__(AP_THREADS)
ascii_pack_kernel(DevParams P, const unsigned char *__restrict__ text, long long text_len, uint32_t *__restrict__ out,
		  unsigned long long *__restrict__ status)
{
	__shared__ uint32_t ring[AP_RING];
	__shared__ int warp_tot[AP_THREADS / 32];
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const long long n = P.n;
	const long long s = blockIdx.x;
	uint32_t *__restrict__ o = out + s * P.pitch_words;
	const long long words = (n + 31) >> 5;
	long long have = 0;			/* digits placed so far */
	long long wdone = 0;			/* words already stored */
	for (int i = tid; i < AP_RING; i += AP_THREADS)
		ring[i] = 0;
	__syncthreads();
	for (long long c0 = s * n; have < n && c0 < text_len; c0 += AP_THREADS * AP_CPT) {
		const long long my = c0 + (long long) tid * AP_CPT;
		unsigned char ch[AP_CPT];
		if (my + AP_CPT <= text_len && (((uintptr_t) (text + my)) & 15) == 0) {
			const uint4 q = __ldg(reinterpret_cast<const uint4 *>(text + my));
			const uint32_t w4[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
			for (int j = 0; j < AP_CPT; j++)
				ch[j] = (unsigned char) (w4[j >> 2] >> (8 * (j & 3)));
		} else {
#pragma unroll
			for (int j = 0; j < AP_CPT; j++)
				ch[j] = (my + j < text_len) ? __ldg(text + my + j) : (unsigned char) ' ';
		}
		uint32_t v = 0;
		int cnt = 0, badj = -1, badcnt = 0;
#pragma unroll
		for (int j = 0; j < AP_CPT; j++) {
			const unsigned int c = ch[j], d = c - (unsigned int) '0';
			if (d <= 1u) {
				v = (v << 1) | d;
				cnt++;
			} else if (!(c == ' ' || (c >= 9u && c <= 13u)) && badj < 0) {	/* isspace(): space \t \n \v \f \r */
				badj = j;
				badcnt = cnt;
			}
		}
		/* block exclusive scan of cnt */
		int inc = cnt;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const int t = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane >= d) inc += t;
		}
		if (lane == 31)
			warp_tot[wid] = inc;
		__syncthreads();
		int pos = inc - cnt, total = 0;
#pragma unroll
		for (int w = 0; w < AP_THREADS / 32; w++) {
			const int t = warp_tot[w];
			if (w < wid) pos += t;
			total += t;
		}
		const long long g = have + pos;			/* stream bit index of this thread's first digit */
		if (badj >= 0 && g + badcnt < n)
			atomicMin(status + 1, (unsigned long long) (my + badj));
		int take = cnt;
		if (g + take > n)
			take = (g < n) ? (int) (n - g) : 0;
		if (take > 0) {
			v >>= (cnt - take);
			const int sh = (int) (g & 31);
			const long long wi = g >> 5;
			if (sh + take <= 32) {
				atomicOr(&ring[wi & (AP_RING - 1)], v << (32 - sh - take));
			} else {
				atomicOr(&ring[wi & (AP_RING - 1)], v >> (sh + take - 32));
				atomicOr(&ring[(wi + 1) & (AP_RING - 1)], v << (64 - sh - take));
			}
		}
		__syncthreads();
		have = (have + total < n) ? have + total : n;
		const long long upto = (have == n) ? words : (have >> 5);	/* words complete after this round */
		for (long long w = wdone + tid; w < upto; w += AP_THREADS) {
			o[w] = bswap32(ring[w & (AP_RING - 1)]);
			ring[w & (AP_RING - 1)] = 0;
		}
		wdone = upto;
		/* the next round's first barrier separates these stores from its atomics, which touch later words */
	}
	if (have < n && tid == 0)
		atomicMin(status, (unsigned long long) s);
}

cudaError_t launch_ascii_pack(const DevParams &P, const unsigned char *d_text, long long text_len, uint32_t *d_in,
			      unsigned long long *d_status, int64_t nstreams, cudaStream_t stream)
{
	ascii_pack_kernel<<<(unsigned) nstreams, AP_THREADS, 0, stream>>>(P, d_text, text_len, d_in, d_status);
	return cudaGetLastError();
}
