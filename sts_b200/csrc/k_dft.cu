/*
 * k_dft.cu - discrete Fourier transform (spectral) test, discreteFourierTransform.c:326-411:
 * the general path for every stream length whose half n/2 is 5-smooth (n = 10^7 = 2^7 5^7 of BASELINE
 * config 4, n = 10^6, every power of two ...).  n = 2^20 takes the specialised kernels of
 * k_dft_fast.cu instead.
 *
 * The reference hands X_i = +-1 (n reals) to libfftw3's r2c plan and counts the moduli below
 * T = sqrt(ln(20) n) among output bins 0 .. n/2 - 1.  Here the n reals are packed as N = n/2 complex
 * points z_j = x_2j + i x_2j+1 and transformed by a hand-written FP64 four-step FFT, N = N1 x N2 with
 * j = N2 j1 + j2 and k = k1 + N1 k2:
 *
 *   dft_cols_kernel : for CW adjacent columns j2, FFT_N1 over j1 of z[N2 j1 + j2] (input decoded
 *                     straight from the packed bits), times W_N^(j2 k1), stored as Y[k1][j2].
 *   dft_rows_kernel : for the row pair (k1, N1 - k1), FFT_N2 over j2 of Y[k1][.] gives
 *                     Z[k1 + N1 k2]; bins q and N - q live in the two rows of the pair, so
 *                     X_q = (Z_q + conj Z_{N-q})/2 - (i/2) W_n^q (Z_q - conj Z_{N-q}) and the count
 *                     are finished in shared memory; only N_1 leaves the kernel.
 *
 * Each 1-D FFT is a Stockham autosort FFT, one pass per radix of the plan (8, 5, 4, 3 or 2), in ONE
 * shared-memory buffer: a thread keeps its (at most two) butterflies of the pass in registers across
 * the barrier between reading and writing.  Twiddles come from per-pass tables computed on the host in
 * long double and laid out so that consecutive butterflies read consecutive entries.
 *
 * Roofline: the intermediate Y is the traffic: 16 N bytes written + 16 N bytes read per stream
 * against n/8 input bytes; FP64 work is about 5 N log2 N flop.
 */
#include <vector>
#include "common.cuh"
#include "dft.h"
#include "dft_math.cuh"

cudaError_t launch_dft_fast(const DevParams &P, const LaunchCtx &L, const DftPlan &D, double2 *d_work, int64_t work_streams);
cudaError_t launch_dft_1e6(const DevParams &P, const LaunchCtx &L, const DftPlan &D, double2 *d_work, int64_t work_streams);
cudaError_t launch_dft_any(const DevParams &P, const LaunchCtx &L, const DftPlan &D, double2 *d_work, int64_t work_streams);
cudaError_t launch_dft_1e7(const DevParams &P, const LaunchCtx &L, const DftPlan &D, double2 *d_work, int64_t work_streams);


/*
 * A Stockham pass of radix R with the butterflies held in registers: a thread owns up to GDFT_PTS / R
 * butterflies of the pass (16 / R for the power-of-two radices; work items tid, tid + T, ...), reads their inputs, applies the pass's
 * twiddles W_(Ns R)^(k r) (table entry (r - 1) Ns + k, k = b mod Ns) and the butterfly, and after a
 * barrier writes the outputs at ((b - k) R + k + r Ns) of the SAME buffer.  The first pass reads from
 * global memory (packed bits / the intermediate Y), the last pass of the columns writes Y.
 */
#define GDFT_PTS 20			/* points a thread holds across the barrier of a pass */
/* butterflies per thread: two in the kernels for power-of-two lengths (WIDE = false), GDFT_PTS / R (16 / R for the
 * power-of-two radices) in the kernels for mixed lengths.  Measured alternatives for the mixed kernels (ms per 1024
 * streams at n = 10^7 / 400000 / 983040): this rule 188 / 5.9 / 17.3; 2 of radix 10 / 8 and 4 of the others
 * 189 / 7.4 / 21.2; at most 10 points 208 / 7.4 / 18.7; the previous two-butterfly rule 298 / 8.8 / 17.6 */
template <int R, bool WIDE> struct GdftMaxB {
	static constexpr int value = !WIDE ? 2 : ((R == 8 || R == 4 || R == 2) ? 16 / R : GDFT_PTS / R);
};

template <int R> __device__ __forceinline__ void tw_bfly(double2 *v, const double2 *__restrict__ wtab, int Ns, int k)
{
	if (Ns > 1) {
#pragma unroll
		for (int r = 1; r < R; r++)
			v[r] = cmul(v[r], __ldg(wtab + (r - 1) * Ns + k));
	}
	if constexpr (R == 10)
		bfly10(v);
	else
		bfly<R>(v);
}

/* ------------------------------------------------------------------------------------------ */
template <int R, bool WIDE>
__device__ __forceinline__ void cols_pass(const DevParams &P, const DftPlan &D, const uint32_t *__restrict__ w, double2 *buf,
					  double2 *__restrict__ Ys, int j20, int ps, int Ns)
{
	const int N1 = D.N1, N2 = D.N2, CW = D.cw, T = D.threads;
	const int q = N1 / R, total = q * CW;
	const bool first = (ps == 0), last = (ps == D.np1 - 1);
	constexpr int GDFT_MAXB = GdftMaxB<R, WIDE>::value;
	double2 v[GDFT_MAXB][R];
	int kk[GDFT_MAXB];
#pragma unroll
	for (int i = 0; i < GDFT_MAXB; i++) {
		const int it = threadIdx.x + i * T;
		if (it < total) {
			const int b = it / CW, c = it - b * CW;
			if (first) {
#pragma unroll
				for (int r = 0; r < R; r++) {	/* z[N2 j1 + j2] = x_2j + i x_2j+1 from the packed bits */
					const int64_t p = 2 * ((int64_t) N2 * (b + r * q) + j20 + c);
					const uint32_t two = (ldw(w, p >> 5) >> (30 - (int) (p & 31))) & 3u;
					v[i][r] = make_double2((two & 2u) ? 1.0 : -1.0, (two & 1u) ? 1.0 : -1.0);
				}
			} else {
#pragma unroll
				for (int r = 0; r < R; r++)
					v[i][r] = buf[(b + r * q) * CW + c];
			}
			kk[i] = b % Ns;
			tw_bfly<R>(v[i], D.w1 + D.off1[ps], Ns, kk[i]);
		}
	}
	__syncthreads();
#pragma unroll
	for (int i = 0; i < GDFT_MAXB; i++) {
		const int it = threadIdx.x + i * T;
		if (it < total) {
			const int b = it / CW, c = it - b * CW;
			const int base = (b - kk[i]) * R + kk[i];
			if (last) {		/* times W_N^(j2 k1), out as Y[k1][j2] */
#pragma unroll
				for (int r = 0; r < R; r++) {
					const int k1 = base + r * Ns;
					const int64_t x = (int64_t) (j20 + c) * k1;		/* < N */
					const double2 tw = cmul(__ldg(D.th + (x >> 10)), __ldg(D.tl + (x & 1023)));
					Ys[(int64_t) k1 * N2 + j20 + c] = cmul(v[i][r], tw);
				}
			} else {
#pragma unroll
				for (int r = 0; r < R; r++)
					buf[(base + r * Ns) * CW + c] = v[i][r];
			}
		}
	}
	__syncthreads();
}

template <bool WIDE>
__global__ void __launch_bounds__(512, 1)
dft_cols_kernel(DevParams P, DftPlan D, const uint32_t *__restrict__ in, double2 *__restrict__ Y, int64_t stream0)
{
	extern __shared__ double2 buf[];			/* [N1][CW] */
	const int64_t sl = blockIdx.x;				/* stream inside the chunk */
	const uint32_t *__restrict__ w = in + (stream0 + sl) * P.pitch_words;
	double2 *__restrict__ Ys = Y + sl * ((int64_t) D.N1 * D.N2);
	const int j20 = blockIdx.y * D.cw;
	int Ns = 1;
	for (int ps = 0; ps < D.np1; ps++) {
		const int R = D.rad1[ps];
		switch (R) {
		case 10: if constexpr (WIDE) cols_pass<10, WIDE>(P, D, w, buf, Ys, j20, ps, Ns); break;
		case 8: cols_pass<8, WIDE>(P, D, w, buf, Ys, j20, ps, Ns); break;
		case 5: cols_pass<5, WIDE>(P, D, w, buf, Ys, j20, ps, Ns); break;
		case 4: cols_pass<4, WIDE>(P, D, w, buf, Ys, j20, ps, Ns); break;
		case 3: cols_pass<3, WIDE>(P, D, w, buf, Ys, j20, ps, Ns); break;
		default: cols_pass<2, WIDE>(P, D, w, buf, Ys, j20, ps, Ns); break;
		}
		Ns *= R;
	}
}

/* ------------------------------------------------------------------------------------------ */
template <int R, bool WIDE>
__device__ __forceinline__ void rows_pass(const DftPlan &D, const double2 *__restrict__ Ys, double2 *buf, int k1a, int k1b, int rows,
					  int ps, int Ns)
{
	const int N2 = D.N2, T = D.threads;
	const int q = N2 / R, total = rows * q;
	const bool first = (ps == 0);
	constexpr int GDFT_MAXB = GdftMaxB<R, WIDE>::value;
	double2 v[GDFT_MAXB][R];
	int kk[GDFT_MAXB];
#pragma unroll
	for (int i = 0; i < GDFT_MAXB; i++) {
		const int it = threadIdx.x + i * T;
		if (it < total) {
			const int r_ = it / q, b = it - r_ * q;
			if (first) {
				const double2 *__restrict__ src = Ys + (int64_t) (r_ ? k1b : k1a) * N2;
#pragma unroll
				for (int r = 0; r < R; r++)
					v[i][r] = src[b + r * q];
			} else {
				const double2 *src = buf + (size_t) r_ * N2;
#pragma unroll
				for (int r = 0; r < R; r++)
					v[i][r] = src[b + r * q];
			}
			kk[i] = b % Ns;
			tw_bfly<R>(v[i], D.w2 + D.off2[ps], Ns, kk[i]);
		}
	}
	__syncthreads();
#pragma unroll
	for (int i = 0; i < GDFT_MAXB; i++) {
		const int it = threadIdx.x + i * T;
		if (it < total) {
			const int r_ = it / q, b = it - r_ * q;
			double2 *dst = buf + (size_t) r_ * N2 + (b - kk[i]) * R + kk[i];
#pragma unroll
			for (int r = 0; r < R; r++)
				dst[r * Ns] = v[i][r];
		}
	}
	__syncthreads();
}

template <bool WIDE>
__global__ void __launch_bounds__(512, 1)
dft_rows_kernel(DevParams P, DftPlan D, const double2 *__restrict__ Y, long long *__restrict__ st, int64_t stream0)
{
	extern __shared__ double2 buf[];			/* [2 rows][N2] */
	__shared__ unsigned int cnt_s;
	const int N1 = D.N1, N2 = D.N2;
	const int64_t sl = blockIdx.x;
	const double2 *__restrict__ Ys = Y + sl * ((int64_t) N1 * N2);
	const int k1a = blockIdx.y, k1b = (N1 - k1a) % N1;	/* row pair, k1a = 0 .. N1/2 */
	const bool self = (k1a == k1b);
	const int rows = self ? 1 : 2;
	if (threadIdx.x == 0)
		cnt_s = 0;
	int Ns = 1;
	for (int ps = 0; ps < D.np2; ps++) {
		const int R = D.rad2[ps];
		switch (R) {
		case 10: if constexpr (WIDE) rows_pass<10, WIDE>(D, Ys, buf, k1a, k1b, rows, ps, Ns); break;
		case 8: rows_pass<8, WIDE>(D, Ys, buf, k1a, k1b, rows, ps, Ns); break;
		case 5: rows_pass<5, WIDE>(D, Ys, buf, k1a, k1b, rows, ps, Ns); break;
		case 4: rows_pass<4, WIDE>(D, Ys, buf, k1a, k1b, rows, ps, Ns); break;
		case 3: rows_pass<3, WIDE>(D, Ys, buf, k1a, k1b, rows, ps, Ns); break;
		default: rows_pass<2, WIDE>(D, Ys, buf, k1a, k1b, rows, ps, Ns); break;
		}
		Ns *= R;
	}
	/* real-FFT post pass + count, discreteFourierTransform.c:392-411 (bins 0 .. n/2 - 1 = all q < N) */
	const double T2x4 = 4.0 * D.T * D.T;			/* |X| < T  <=>  |2X|^2 < 4 T^2 (see k_dft_fast.cu) */
	unsigned int cnt = 0;
	for (int e = threadIdx.x; e < rows * N2; e += D.threads) {
		const int r_ = e / N2, k2 = e - r_ * N2;
		const int kq = r_ ? k1b : k1a;			/* q = kq + N1 k2 */
		const double2 *mine = buf + (size_t) r_ * N2;
		const double2 *other = buf + (size_t) (self ? 0 : (1 - r_)) * N2;
		const int k2p = (kq == 0) ? ((N2 - k2) % N2) : (N2 - 1 - k2);
		const double2 zq = mine[k2];
		double2 zp = other[k2p];
		zp.y = -zp.y;					/* conj Z_{N-q} */
		const double2 wq = cmul(__ldg(D.p1 + kq), __ldg(D.p2 + k2));	/* exp(-2 pi i q / n) */
		const double2 sum = cadd(zq, zp), dif = csub(zq, zp);
		const double2 wd = cmul(wq, dif);
		/* 2 X = sum - i wd = (sum.x + wd.y, sum.y - wd.x) */
		const double xr = sum.x + wd.y, xi = sum.y - wd.x;
		cnt += (xr * xr + xi * xi < T2x4) ? 1u : 0u;
	}
	cnt = (unsigned int) warp_sum((int) cnt);
	if ((threadIdx.x & 31) == 0 && cnt)
		atomicAdd(&cnt_s, cnt);
	__syncthreads();
	if (threadIdx.x == 0 && cnt_s)
		atomicAdd((unsigned long long *) (st + (stream0 + sl) * P.nst + P.st_off[STSGPU_TEST_DFT]),
			  (unsigned long long) cnt_s);
}

/* ------------------------------------------------------------------------------------------ */
/* radix list of a 5-smooth length: 10s (a 5 with a 2, prime-factor butterfly), the remaining 5s, 3s, then the power of
 * two as 8s with a 4 4 / 4 / 2 remainder */
static bool radix_list(int len, int *rad, int *np)
{
	int k = 0, two = 0, five = 0, three = 0;
	while (len % 5 == 0) { five++; len /= 5; }
	while (len % 3 == 0) { three++; len /= 3; }
	while (len % 2 == 0) { two++; len /= 2; }
	if (len != 1)
		return false;
	auto push = [&](int r) { if (k >= DFT_MAX_PASSES) return false; rad[k++] = r; return true; };
	while (five > 0 && two > 0) { if (!push(10)) return false; five--; two--; }
	while (five > 0) { if (!push(5)) return false; five--; }
	while (three > 0) { if (!push(3)) return false; three--; }
	while (two >= 3 && two != 4) { if (!push(8)) return false; two -= 3; }
	while (two >= 2) { if (!push(4)) return false; two -= 2; }
	if (two == 1 && !push(2)) return false;
	*np = k;
	return true;
}

/* points per thread the planner lets a pass of radix r hold (at most GdftMaxB<r> butterflies).  Power-of-two lengths
 * keep two butterflies per thread in every pass - their plans were tuned that way (n = 2^22: 60 ms per 1024 streams
 * against 76 ms with the wider tiles) - the mixed lengths use the full GDFT_PTS / 16 points, which at n = 10^7
 * turns the 2500 x 2000 split with single-column tiles into 2000 x 2500 with 4 columns (64-byte runs of Y) */
static int pts_of_radix(int r, bool pow2)
{
	if (pow2)
		return 2 * r;
	return (r == 8 || r == 4 || r == 2) ? 16 : (GDFT_PTS / r) * r;	/* GdftMaxB<r, true> */
}

bool dft_make_plan(long long n, DftPlan *D)
{
	if (n < 16 || (n & 1))
		return false;
	const long long N = n / 2;
	/*
	 * Every pass keeps its butterflies in registers, at most GDFT_PTS (16 for radix 8/4/2) points per thread:
	 * with T threads, rows: 2 N2 <= pts T, columns: CW N1 <= pts T for every radix of the list.  Among the
	 * feasible splits take the widest column tile (coalesced Y stores), then 256 threads (two CTAs per
	 * SM), then the longest rows.
	 */
	int best_cw = 0, best_t = 0;
	long long best1 = 0, best2 = 0;
	for (int T = 256; T <= 512; T *= 2) {
		for (long long n2 = 2; n2 <= 4096 && n2 <= N; n2++) {
			if (N % n2)
				continue;
			const long long n1 = N / n2;
			if (n1 > 4096 || n1 < 2)
				continue;
			int r1[DFT_MAX_PASSES], r2[DFT_MAX_PASSES], k1, k2;
			if (!radix_list((int) n1, r1, &k1) || !radix_list((int) n2, r2, &k2))
				continue;
			int pts1 = GDFT_PTS, pts2 = GDFT_PTS;		/* the tightest pass decides */
			const bool pow2 = (N & (N - 1)) == 0;
			for (int i = 0; i < k1; i++) pts1 = pts_of_radix(r1[i], pow2) < pts1 ? pts_of_radix(r1[i], pow2) : pts1;
			for (int i = 0; i < k2; i++) pts2 = pts_of_radix(r2[i], pow2) < pts2 ? pts_of_radix(r2[i], pow2) : pts2;
			if (2 * n2 > (long long) pts2 * T)
				continue;
			if ((size_t) 2 * n2 * sizeof(double2) > 200 * 1024)
				continue;
			int cw = 16;
			while (cw >= 1 && (n2 % cw != 0 || (long long) cw * n1 > (long long) pts1 * T ||
					   (size_t) n1 * cw * sizeof(double2) > 200 * 1024))
				cw >>= 1;
			if (cw < 1)
				continue;
			const bool better = cw > best_cw || (cw == best_cw && (T < best_t || (T == best_t && n2 > best2)));
			if (best_t == 0 || better) {
				best_cw = cw; best_t = T; best1 = n1; best2 = n2;
			}
		}
	}
	if (best_t == 0)
		return false;
	D->N1 = (int) best1;
	D->N2 = (int) best2;
	D->cw = best_cw;
	D->threads = best_t;
	radix_list(D->N1, D->rad1, &D->np1);
	radix_list(D->N2, D->rad2, &D->np2);
	return true;
}

cudaError_t launch_dft(const DevParams &P, const LaunchCtx &L, const DftPlan &D, double2 *d_work, int64_t work_streams,
		       int64_t *launches)
{
	*launches = 0;
	if (!(P.enabled & (1u << STSGPU_TEST_DFT)))
		return cudaSuccess;
	const int64_t chunks = (L.nstreams + work_streams - 1) / work_streams;
	*launches = (D.fast == 3 ? dft_any_launches_per_chunk(D) : 2) * chunks;
	if (D.fast == 1)
		return launch_dft_fast(P, L, D, d_work, work_streams);
	if (D.fast == 2)
		return launch_dft_1e6(P, L, D, d_work, work_streams);
	if (D.fast == 3)
		return launch_dft_any(P, L, D, d_work, work_streams);
	if (D.fast == 4)
		return launch_dft_1e7(P, L, D, d_work, work_streams);
	const size_t smem1 = (size_t) D.N1 * D.cw * sizeof(double2);
	const size_t smem2 = (size_t) 2 * D.N2 * sizeof(double2);
	cudaError_t e;
	const bool wide = (((int64_t) D.N1 * D.N2) & ((int64_t) D.N1 * D.N2 - 1)) != 0;	/* not a power of two: dft_make_plan */
	e = cudaFuncSetAttribute(wide ? dft_cols_kernel<true> : dft_cols_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
				 (int) smem1);
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(wide ? dft_rows_kernel<true> : dft_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
				 (int) smem2);
	if (e != cudaSuccess) return e;
	for (int64_t s0 = 0; s0 < L.nstreams; s0 += work_streams) {
		const int64_t cs = (L.nstreams - s0 < work_streams) ? L.nstreams - s0 : work_streams;
		dim3 g1((unsigned) cs, (unsigned) (D.N2 / D.cw));
		if (wide)
			dft_cols_kernel<true><<<g1, D.threads, smem1, L.stream>>>(P, D, L.d_in, d_work, s0);
		else
			dft_cols_kernel<false><<<g1, D.threads, smem1, L.stream>>>(P, D, L.d_in, d_work, s0);
		dim3 g2((unsigned) cs, (unsigned) (D.N1 / 2 + 1));
		if (wide)
			dft_rows_kernel<true><<<g2, D.threads, smem2, L.stream>>>(P, D, d_work, L.d_st, s0);
		else
			dft_rows_kernel<false><<<g2, D.threads, smem2, L.stream>>>(P, D, d_work, L.d_st, s0);
	}
	return cudaGetLastError();
}
