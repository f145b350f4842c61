/*
 * k_dft.cu - discrete Fourier transform (spectral) test, discreteFourierTransform.c:326-411.
 *
 * The reference hands X_i = +-1 (n reals) to libfftw3's r2c plan and counts the moduli below
 * T = sqrt(ln(20) n) among output bins 0 .. n/2 - 1.  Here the n reals are packed as N = n/2 complex
 * points z_j = x_2j + i x_2j+1, transformed by a hand-written FP64 four-step FFT (N = N1 x N2, both
 * <= 1024, so every 1-D FFT runs out of registers + shared memory), and un-packed with the usual
 * real-FFT post pass fused into the second kernel together with the |X| < T count:
 *
 *   dft_cols_kernel : for CW consecutive columns j2, FFT_N1 over j1 of z[N2 j1 + j2] (input decoded
 *                     straight from the packed bits), times W_N^(j2 k1), stored as Y[k1][j2]
 *                     (128-byte contiguous runs per k1).
 *   dft_rows_kernel : for the row pair (k1, N1 - k1), FFT_N2 over j2 of Y[k1][.] gives
 *                     Z[k1 + N1 k2]; bins q and N - q live in the two rows of the pair, so
 *                     X_q = (Z_q + conj Z_{N-q})/2 - (i/2) W_n^q (Z_q - conj Z_{N-q}) and the count
 *                     are finished in shared memory; only N_1 leaves the kernel.
 *
 * Each 1-D FFT is a Stockham autosort FFT, radix 8 passes (plus one radix 4 or 2 pass), 8 complex
 * points per thread per pass, twiddles from L1-resident tables computed on the host.
 *
 * Roofline: the intermediate Y is the traffic: 16 N bytes written + 16 N bytes read per stream
 * (8 MiB + 8 MiB at n = 2^20) against n/8 input bytes; FP64 work is about 5 N log2 N flop.
 * Only power-of-two n with 2^10 <= n <= 2^20 is handled by this path (N1 <= 512, N2 <= 1024).
 */
#include "common.cuh"
#include "dft.h"

#include "dft_math.cuh"

cudaError_t launch_dft_fast(const DevParams &P, const LaunchCtx &L, const DftPlan &D, double2 *d_work, int64_t work_streams);

/* ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(512)
dft_cols_kernel(DevParams P, DftPlan D, const uint32_t *__restrict__ in, double2 *__restrict__ Y, int64_t stream0)
{
	extern __shared__ double2 buf[];			/* [N1][CW] */
	const int N1 = 1 << D.logN1, N2 = 1 << D.logN2, CW = D.cw;
	const int64_t sl = blockIdx.x;				/* stream inside the chunk */
	const uint32_t *__restrict__ w = in + (stream0 + sl) * P.pitch_words;
	double2 *__restrict__ Ys = Y + sl * ((int64_t) N1 * N2);
	const int c = threadIdx.x % CW, t = threadIdx.x / CW;
	const int j2 = blockIdx.y * CW + c;
	const int n8 = D.logN1 / 3, rem = D.logN1 % 3;
	const int nst = n8 + (rem ? 1 : 0);

	auto load_bits = [&](int j1) -> double2 {
		const int64_t p = 2 * ((int64_t) N2 * j1 + j2);
		const uint32_t wd = ldw(w, p >> 5);
		const uint32_t two = (wd >> (30 - (int) (p & 31))) & 3u;
		return make_double2((two & 2u) ? 1.0 : -1.0, (two & 1u) ? 1.0 : -1.0);
	};
	auto load_smem = [&](int e) -> double2 { return buf[e * CW + c]; };
	auto store_smem = [&](int e, double2 val) { buf[e * CW + c] = val; };
	auto store_out = [&](int k1, double2 val) {
		const int e = j2 * k1;				/* < N */
		const double2 tw = cmul(__ldg(D.th + (e >> 10)), __ldg(D.tl + (e & 1023)));
		Ys[(int64_t) k1 * N2 + j2] = cmul(val, tw);
	};

	int Ns = 1;
	for (int sidx = 0; sidx < nst; sidx++) {
		const bool first = (sidx == 0), last = (sidx == nst - 1);
		const int R = (sidx < n8) ? 8 : (1 << rem);
		if (!first)
			__syncthreads();			/* previous pass's stores are visible */
		double2 v[8];
		int dst[8];
		if (R == 8) {
			if (first) {
#pragma unroll
				for (int r = 0; r < 8; r++) v[r] = load_bits(t + r * (N1 >> 3));
			} else {
#pragma unroll
				for (int r = 0; r < 8; r++) v[r] = load_smem(t + r * (N1 >> 3));
			}
			const int k = t & (Ns - 1);
			if (Ns > 1) {
				const int step = 1024 / (Ns * 8);
#pragma unroll
				for (int r = 1; r < 8; r++) v[r] = cmul(v[r], __ldg(D.wm + k * r * step));
			}
			bfly<8>(v);
			const int base = (t - k) * 8 + k;
#pragma unroll
			for (int r = 0; r < 8; r++) dst[r] = base + r * Ns;
		} else if (R == 4) {
#pragma unroll
			for (int u = 0; u < 2; u++) {
				const int j = t + u * (N1 >> 3);
#pragma unroll
				for (int r = 0; r < 4; r++)
					v[u * 4 + r] = first ? load_bits(j + r * (N1 >> 2)) : load_smem(j + r * (N1 >> 2));
				const int k = j & (Ns - 1);
				if (Ns > 1) {
					const int step = 1024 / (Ns * 4);
#pragma unroll
					for (int r = 1; r < 4; r++) v[u * 4 + r] = cmul(v[u * 4 + r], __ldg(D.wm + k * r * step));
				}
				bfly<4>(&v[u * 4]);
				const int base = (j - k) * 4 + k;
#pragma unroll
				for (int r = 0; r < 4; r++) dst[u * 4 + r] = base + r * Ns;
			}
		} else {
#pragma unroll
			for (int u = 0; u < 4; u++) {
				const int j = t + u * (N1 >> 3);
#pragma unroll
				for (int r = 0; r < 2; r++)
					v[u * 2 + r] = first ? load_bits(j + r * (N1 >> 1)) : load_smem(j + r * (N1 >> 1));
				const int k = j & (Ns - 1);
				if (Ns > 1)
					v[u * 2 + 1] = cmul(v[u * 2 + 1], __ldg(D.wm + k * (1024 / (Ns * 2))));
				bfly<2>(&v[u * 2]);
				const int base = (j - k) * 2 + k;
#pragma unroll
				for (int r = 0; r < 2; r++) dst[u * 2 + r] = base + r * Ns;
			}
		}
		if (!first)
			__syncthreads();			/* everybody has read this pass's inputs */
		if (last) {
#pragma unroll
			for (int r = 0; r < 8; r++) store_out(dst[r], v[r]);
		} else {
#pragma unroll
			for (int r = 0; r < 8; r++) store_smem(dst[r], v[r]);
		}
		Ns *= R;
	}
}

/* ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(256)
dft_rows_kernel(DevParams P, DftPlan D, const double2 *__restrict__ Y, long long *__restrict__ st, int64_t stream0)
{
	extern __shared__ double2 buf[];			/* [2][padi(N2)] */
	__shared__ unsigned int cnt_s;
	const int N1 = 1 << D.logN1, N2 = 1 << D.logN2;
	const int64_t sl = blockIdx.x;
	const double2 *__restrict__ Ys = Y + sl * ((int64_t) N1 * N2);
	const int pi = blockIdx.y;				/* pair index 0 .. N1/2 */
	const int k1a = pi, k1b = (N1 - pi) & (N1 - 1);
	const bool self = (k1a == k1b);
	const int tpr = N2 >> 3;				/* threads per row */
	const int rho = threadIdx.x / tpr, t = threadIdx.x % tpr;
	const int rowlen = padi(N2) + 1;
	const bool active = (rho == 0) || !self;
	const int k1 = rho ? k1b : k1a;
	const double2 *__restrict__ src = Ys + (int64_t) k1 * N2;
	double2 *row = buf + rho * rowlen;
	if (threadIdx.x == 0)
		cnt_s = 0;

	const int n8 = D.logN2 / 3, rem = D.logN2 % 3;
	const int nst = n8 + (rem ? 1 : 0);
	int Ns = 1;
	for (int sidx = 0; sidx < nst; sidx++) {
		const bool first = (sidx == 0);
		const int R = (sidx < n8) ? 8 : (1 << rem);
		double2 v[8];
		int dst[8];
		if (!first)
			__syncthreads();
		if (active) {
			if (R == 8) {
#pragma unroll
				for (int r = 0; r < 8; r++) {
					const int e = t + r * (N2 >> 3);
					v[r] = first ? src[e] : row[padi(e)];
				}
				const int k = t & (Ns - 1);
				if (Ns > 1) {
					const int step = 1024 / (Ns * 8);
#pragma unroll
					for (int r = 1; r < 8; r++) v[r] = cmul(v[r], __ldg(D.wm + k * r * step));
				}
				bfly<8>(v);
				const int base = (t - k) * 8 + k;
#pragma unroll
				for (int r = 0; r < 8; r++) dst[r] = base + r * Ns;
			} else if (R == 4) {
#pragma unroll
				for (int u = 0; u < 2; u++) {
					const int j = t + u * (N2 >> 3);
#pragma unroll
					for (int r = 0; r < 4; r++) {
						const int e = j + r * (N2 >> 2);
						v[u * 4 + r] = first ? src[e] : row[padi(e)];
					}
					const int k = j & (Ns - 1);
					if (Ns > 1) {
						const int step = 1024 / (Ns * 4);
#pragma unroll
						for (int r = 1; r < 4; r++)
							v[u * 4 + r] = cmul(v[u * 4 + r], __ldg(D.wm + k * r * step));
					}
					bfly<4>(&v[u * 4]);
					const int base = (j - k) * 4 + k;
#pragma unroll
					for (int r = 0; r < 4; r++) dst[u * 4 + r] = base + r * Ns;
				}
			} else {
#pragma unroll
				for (int u = 0; u < 4; u++) {
					const int j = t + u * (N2 >> 3);
#pragma unroll
					for (int r = 0; r < 2; r++) {
						const int e = j + r * (N2 >> 1);
						v[u * 2 + r] = first ? src[e] : row[padi(e)];
					}
					const int k = j & (Ns - 1);
					if (Ns > 1)
						v[u * 2 + 1] = cmul(v[u * 2 + 1], __ldg(D.wm + k * (1024 / (Ns * 2))));
					bfly<2>(&v[u * 2]);
					const int base = (j - k) * 2 + k;
#pragma unroll
					for (int r = 0; r < 2; r++) dst[u * 2 + r] = base + r * Ns;
				}
			}
		}
		if (!first)
			__syncthreads();
		if (active) {
#pragma unroll
			for (int r = 0; r < 8; r++) row[padi(dst[r])] = v[r];
		}
		Ns *= R;
	}
	__syncthreads();

	/* real-FFT post pass + count, discreteFourierTransform.c:392-411 (bins 0 .. n/2 - 1 = all q < N) */
	unsigned int cnt = 0;
	const int rows = self ? 1 : 2;
	for (int e = threadIdx.x; e < rows * N2; e += blockDim.x) {
		const int r_ = e / N2, k2 = e % N2;
		const int kq = r_ ? k1b : k1a;			/* q = kq + N1 k2 */
		const double2 *mine = buf + r_ * rowlen;
		const double2 *other = buf + (self ? 0 : (1 - r_)) * rowlen;
		const int k2p = (kq == 0) ? ((N2 - k2) & (N2 - 1)) : (N2 - 1 - k2);
		const double2 zq = mine[padi(k2)];
		double2 zp = other[padi(k2p)];
		zp.y = -zp.y;					/* conj Z_{N-q} */
		const double2 wq = cmul(__ldg(D.p1 + kq), __ldg(D.p2 + k2));	/* exp(-2 pi i q / n) */
		const double2 sum = cadd(zq, zp), dif = csub(zq, zp);
		const double2 wd = cmul(wq, dif);
		/* X = sum/2 - (i/2) wd  =  (sum.x + wd.y, sum.y - wd.x) / 2 */
		const double xr = 0.5 * (sum.x + wd.y), xi = 0.5 * (sum.y - wd.x);
		const double mod = sqrt(xr * xr + xi * xi);
		cnt += (mod < D.T) ? 1u : 0u;
	}
	cnt = (unsigned int) warp_sum((int) cnt);
	if ((threadIdx.x & 31) == 0 && cnt)
		atomicAdd(&cnt_s, cnt);
	__syncthreads();
	if (threadIdx.x == 0 && cnt_s)
		atomicAdd((unsigned long long *) (st + (stream0 + sl) * P.nst + P.st_off[STSGPU_TEST_DFT]),
			  (unsigned long long) cnt_s);
}

cudaError_t launch_dft(const DevParams &P, const LaunchCtx &L, const DftPlan &D, double2 *d_work, int64_t work_streams)
{
	if (!(P.enabled & (1u << STSGPU_TEST_DFT)))
		return cudaSuccess;
	const int N1 = 1 << D.logN1, N2 = 1 << D.logN2;
	const int threads1 = D.cw * (N1 >> 3);
	const size_t smem1 = (size_t) N1 * D.cw * sizeof(double2);
	const int threads2 = 2 * (N2 >> 3);
	const size_t smem2 = (size_t) 2 * (N2 + (N2 >> 3) + 1) * sizeof(double2);
	if (D.logN1 == 9 && D.logN2 == 10)
		return launch_dft_fast(P, L, D, d_work, work_streams);
	cudaError_t e;
	e = cudaFuncSetAttribute(dft_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem1);
	if (e != cudaSuccess) return e;
	e = cudaFuncSetAttribute(dft_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem2);
	if (e != cudaSuccess) return e;
	for (int64_t s0 = 0; s0 < L.nstreams; s0 += work_streams) {
		const int64_t cs = (L.nstreams - s0 < work_streams) ? L.nstreams - s0 : work_streams;
		dim3 g1((unsigned) cs, (unsigned) (N2 / D.cw));
		dft_cols_kernel<<<g1, threads1, smem1, L.stream>>>(P, D, L.d_in, d_work, s0);
		dim3 g2((unsigned) cs, (unsigned) (N1 / 2 + 1));
		dft_rows_kernel<<<g2, threads2, smem2, L.stream>>>(P, D, d_work, L.d_st, s0);
	}
	return cudaGetLastError();
}
