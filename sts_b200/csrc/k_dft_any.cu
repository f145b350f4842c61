/*
 * k_dft_any.cu - the DFT test for stream lengths whose half is NOT 5-smooth (n = 1000008, 2 x a prime ...): the
 * reference's libfftw3 plan (discreteFourierTransform.c:194) takes any n, so the engine must too.  Bluestein's
 * chirp-z identity turns the N = n/2 point DFT of z_j = x_2j + i x_2j+1 into a cyclic convolution of power-of-two
 * length M >= 2N - 1:
 *
 *     j k = (j^2 + k^2 - (k - j)^2) / 2
 *     Z_k = c_k  sum_j (z_j c_j) conj(c_(k-j)),        c_j = exp(-i pi j^2 / N)
 *
 * with c_j from the host (long double, j^2 reduced mod 2N in integers).  Per chunk of streams:
 *
 *   bs_load_kernel   A[j] = z_j c_j (bits decoded straight from the packed input), zero up to M
 *   bs_fft_pass_*    Stockham autosort passes (radix 4, one radix 2 if log2 M is odd) through global memory
 *   bs_mul_kernel    A[k] <- conj(A[k] Bhat[k]),  Bhat = FFT_M(conj chirp, wrapped) / M, made once per engine
 *   bs_fft_pass_*    the same forward passes again: conj(FFT(conj .)) is the inverse transform
 *   bs_post_kernel   Z_k = conj(A[k]) c_k, the real-FFT post pass on the pair (q, N - q) and the count of
 *                    |X_q| < T over q in [0, n/2)  (discreteFourierTransform.c:392-411, quirk A4)
 *
 * This is the slow lane on purpose - plain global-memory passes, about 2 log4(M) x 32 M bytes of traffic per
 * stream, roughly 50 x the specialised n = 2^20 kernels - for lengths the fast paths cannot take; the 5-smooth
 * lengths (every power of two, 10^6, 10^7 ...) never come here.
 */
#include <math.h>
#include <vector>
#include "common.cuh"
#include "dft.h"
#include "dft_math.cuh"

#define BS_THREADS 256
#define BS_MAX_LOG 25			/* M <= 2^25: 512 MiB per work buffer and stream */

static double2 bs_root(long double num, long double den)	/* exp(-2 pi i num / den) */
{
	const long double a = -2.0L * 3.14159265358979323846264338327950288L * num / den;
	return make_double2((double) cosl(a), (double) sinl(a));
}

bool dft_any_plan(long long n, DftPlan *D)
{
	if (n < 16 || (n & 1))
		return false;
	const long long N = n / 2;
	int lg = 1;
	while ((1LL << lg) < 2 * N - 1)
		lg++;
	if (lg > BS_MAX_LOG)
		return false;
	D->fast = 3;
	D->bs_logM = lg;
	D->bs_M = 1LL << lg;
	D->N1 = D->N2 = 0;
	return true;
}

/* ---- kernels: blockIdx.y = stream of the chunk, grid-stride over the elements ------------------------------- */
__global__ void __launch_bounds__(BS_THREADS)
bs_load_kernel(DevParams P, DftPlan D, const uint32_t *__restrict__ in, double2 *__restrict__ A, int64_t stream0)
{
	const int64_t M = D.bs_M, N = P.n / 2;
	const uint32_t *__restrict__ w = in + (stream0 + blockIdx.y) * P.pitch_words;
	double2 *__restrict__ As = A + (int64_t) blockIdx.y * M;
	for (int64_t j = (int64_t) blockIdx.x * BS_THREADS + threadIdx.x; j < M; j += (int64_t) gridDim.x * BS_THREADS) {
		double2 v = make_double2(0.0, 0.0);
		if (j < N) {
			/* stream bits 2j (real part) and 2j + 1: word j / 16, MSB first */
			const uint32_t two = (ldw(w, j >> 4) >> (30 - (int) ((2 * j) & 31))) & 3u;
			v = cmul(make_double2((two & 2u) ? 1.0 : -1.0, (two & 1u) ? 1.0 : -1.0), __ldg(D.bs_chirp + j));
		}
		As[j] = v;
	}
}

/* one Stockham pass of radix 4: butterfly j reads j + r M/4, twiddles exp(-2 pi i k r / (4 Ns)), k = j mod Ns */
__global__ void __launch_bounds__(BS_THREADS)
bs_fft_pass4_kernel(DftPlan D, const double2 *__restrict__ src, double2 *__restrict__ dst, int64_t Ns)
{
	const int64_t M = D.bs_M, quarter = M >> 2;
	const double2 *__restrict__ a = src + (int64_t) blockIdx.y * M;
	double2 *__restrict__ o = dst + (int64_t) blockIdx.y * M;
	const int64_t tstep = quarter / Ns;			/* exp(-2 pi i k / (4 Ns)) = tw[k M / (4 Ns)] */
	for (int64_t j = (int64_t) blockIdx.x * BS_THREADS + threadIdx.x; j < quarter; j += (int64_t) gridDim.x * BS_THREADS) {
		const int64_t k = j & (Ns - 1);
		double2 v[4] = { a[j], a[j + quarter], a[j + 2 * quarter], a[j + 3 * quarter] };
		if (k) {
			const double2 w1 = __ldg(D.bs_tw + k * tstep), w2 = __ldg(D.bs_tw + 2 * k * tstep);
			v[1] = cmul(v[1], w1);
			v[2] = cmul(v[2], w2);
			v[3] = cmul(v[3], cmul(w1, w2));
		}
		bfly<4>(v);
		const int64_t j0 = ((j - k) << 2) + k;
#pragma unroll
		for (int r = 0; r < 4; r++)
			o[j0 + r * Ns] = v[r];
	}
}

/* the same with radix 2 (the last pass when log2 M is odd) */
__global__ void __launch_bounds__(BS_THREADS)
bs_fft_pass2_kernel(DftPlan D, const double2 *__restrict__ src, double2 *__restrict__ dst, int64_t Ns)
{
	const int64_t M = D.bs_M, half = M >> 1;
	const double2 *__restrict__ a = src + (int64_t) blockIdx.y * M;
	double2 *__restrict__ o = dst + (int64_t) blockIdx.y * M;
	const int64_t tstep = half / Ns;			/* exp(-2 pi i k / (2 Ns)) = tw[k M / (2 Ns)] */
	for (int64_t j = (int64_t) blockIdx.x * BS_THREADS + threadIdx.x; j < half; j += (int64_t) gridDim.x * BS_THREADS) {
		const int64_t k = j & (Ns - 1);
		double2 v[2] = { a[j], a[j + half] };
		if (k)
			v[1] = cmul(v[1], __ldg(D.bs_tw + k * tstep));
		bfly<2>(v);
		const int64_t j0 = ((j - k) << 1) + k;
		o[j0] = v[0];
		o[j0 + Ns] = v[1];
	}
}

__global__ void __launch_bounds__(BS_THREADS)
bs_mul_kernel(DftPlan D, double2 *__restrict__ A)
{
	const int64_t M = D.bs_M;
	double2 *__restrict__ As = A + (int64_t) blockIdx.y * M;
	for (int64_t k = (int64_t) blockIdx.x * BS_THREADS + threadIdx.x; k < M; k += (int64_t) gridDim.x * BS_THREADS) {
		const double2 p = cmul(As[k], __ldg(D.bs_bhat + k));
		As[k] = make_double2(p.x, -p.y);
	}
}

__global__ void __launch_bounds__(BS_THREADS)
bs_post_kernel(DevParams P, DftPlan D, const double2 *__restrict__ R, long long *__restrict__ st, int64_t stream0)
{
	__shared__ unsigned int cnt_s;
	const int64_t M = D.bs_M, N = P.n / 2;
	const double2 *__restrict__ Rs = R + (int64_t) blockIdx.y * M;
	const double T2x4 = 4.0 * D.T * D.T;
	if (threadIdx.x == 0)
		cnt_s = 0;
	__syncthreads();
	unsigned int cnt = 0;
	/* q = 0 and q = N/2 are their own partners (Z_N = Z_0; bin n/2 is not counted, quirk A4); N is even */
	for (int64_t q = (int64_t) blockIdx.x * BS_THREADS + threadIdx.x; q <= N / 2; q += (int64_t) gridDim.x * BS_THREADS) {
		const int64_t p = (q == 0) ? 0 : N - q;
		const double2 rq = Rs[q], rp = Rs[p];
		const double2 zq = cmul(make_double2(rq.x, -rq.y), __ldg(D.bs_chirp + q));
		const double2 zp = cmul(make_double2(rp.x, -rp.y), __ldg(D.bs_chirp + p));
		cnt += post_pair(zq, zp, __ldg(D.bs_wq + q), T2x4, p != q);
	}
	if (cnt)
		atomicAdd(&cnt_s, cnt);
	__syncthreads();
	if (threadIdx.x == 0 && cnt_s)
		atomicAdd((unsigned long long *) (st + (stream0 + blockIdx.y) * P.nst + P.st_off[STSGPU_TEST_DFT]),
			  (unsigned long long) cnt_s);
}

/* ---- host ---------------------------------------------------------------------------------------------------- */
static unsigned bs_grid_x(int64_t items, int64_t streams, int num_sms)
{
	int64_t want = (items + BS_THREADS - 1) / BS_THREADS;
	int64_t cap = ((int64_t) num_sms * 16 + streams - 1) / streams;	/* about 16 CTAs per SM over the whole chunk */
	if (cap < 1) cap = 1;
	if (want > cap) want = cap;
	return (unsigned) (want < 1 ? 1 : want);
}

/* forward FFT of `streams` rows of length M: passes ping-pong between *a and *b; on return *a holds the result */
static cudaError_t bs_fft(const DftPlan &D, double2 **a, double2 **b, int64_t streams, int num_sms, cudaStream_t st)
{
	int left = D.bs_logM;
	for (int64_t Ns = 1; left > 0;) {
		if (left >= 2) {
			dim3 g(bs_grid_x(D.bs_M >> 2, streams, num_sms), (unsigned) streams);
			bs_fft_pass4_kernel<<<g, BS_THREADS, 0, st>>>(D, *a, *b, Ns);
			Ns <<= 2;
			left -= 2;
		} else {
			dim3 g(bs_grid_x(D.bs_M >> 1, streams, num_sms), (unsigned) streams);
			bs_fft_pass2_kernel<<<g, BS_THREADS, 0, st>>>(D, *a, *b, Ns);
			Ns <<= 1;
			left -= 1;
		}
		double2 *t = *a; *a = *b; *b = t;
	}
	return cudaGetLastError();
}

int dft_any_launches_per_chunk(const DftPlan &D)
{
	const int passes = D.bs_logM / 2 + (D.bs_logM & 1);
	return 1 + passes + 1 + passes + 1;
}

/*
 * Tables of one engine: tabs[0] chirp c_j (N), tabs[1] Bhat (M), tabs[2] exp(-2 pi i t / M) for t < M/2,
 * tabs[3] exp(-2 pi i q / n) for q <= N/2.  Bhat is made here with the device FFT itself.
 */
cudaError_t dft_any_setup(long long n, DftPlan *D, double2 *tabs[4], int num_sms)
{
	const long long N = n / 2, M = D->bs_M;
	cudaError_t e;
	std::vector<double2> chirp((size_t) N), tw((size_t) (M / 2)), wq((size_t) (N / 2 + 1)), b((size_t) M, make_double2(0.0, 0.0));
	for (long long j = 0; j < N; j++)
		chirp[(size_t) j] = bs_root((long double) ((j * j) % (2 * N)), (long double) (2 * N));
	for (long long t = 0; t < M / 2; t++)
		tw[(size_t) t] = bs_root((long double) t, (long double) M);
	for (long long q = 0; q <= N / 2; q++)
		wq[(size_t) q] = bs_root((long double) q, (long double) n);
	const double inv = 1.0 / (double) M;			/* exact: M is a power of two */
	for (long long j = 0; j < N; j++) {
		const double2 c = make_double2(chirp[(size_t) j].x * inv, -chirp[(size_t) j].y * inv);	/* conj c_j / M */
		b[(size_t) j] = c;
		if (j)
			b[(size_t) (M - j)] = c;
	}
	for (int i = 0; i < 4; i++)
		tabs[i] = nullptr;
	double2 *tmp = nullptr;
	const struct { double2 **dst; const std::vector<double2> *src; } up[4] = { { &tabs[0], &chirp }, { &tabs[1], &b }, { &tabs[2], &tw },
										   { &tabs[3], &wq } };
	for (int i = 0; i < 4; i++) {
		e = cudaMalloc((void **) up[i].dst, up[i].src->size() * sizeof(double2));
		if (e != cudaSuccess) return e;
		e = cudaMemcpy(*up[i].dst, up[i].src->data(), up[i].src->size() * sizeof(double2), cudaMemcpyHostToDevice);
		if (e != cudaSuccess) return e;
	}
	D->bs_chirp = tabs[0];
	D->bs_tw = tabs[2];
	D->bs_wq = tabs[3];
	e = cudaMalloc((void **) &tmp, (size_t) M * sizeof(double2));
	if (e != cudaSuccess) return e;
	double2 *a = tabs[1], *c = tmp;
	e = bs_fft(*D, &a, &c, 1, num_sms, 0);
	if (e == cudaSuccess)
		e = cudaDeviceSynchronize();
	tabs[1] = a;						/* whichever buffer holds the spectrum stays, the other goes */
	cudaFree(c);
	D->bs_bhat = tabs[1];
	return e;
}

cudaError_t launch_dft_any(const DevParams &P, const LaunchCtx &L, const DftPlan &D, double2 *d_work, int64_t work_streams)
{
	const int64_t M = D.bs_M;
	for (int64_t s0 = 0; s0 < L.nstreams; s0 += work_streams) {
		const int64_t cs = (L.nstreams - s0 < work_streams) ? L.nstreams - s0 : work_streams;
		double2 *a = d_work, *b = d_work + (size_t) work_streams * M;
		dim3 g(bs_grid_x(M, cs, L.num_sms), (unsigned) cs);
		bs_load_kernel<<<g, BS_THREADS, 0, L.stream>>>(P, D, L.d_in, a, s0);
		cudaError_t e = bs_fft(D, &a, &b, cs, L.num_sms, L.stream);
		if (e != cudaSuccess) return e;
		bs_mul_kernel<<<g, BS_THREADS, 0, L.stream>>>(D, a);
		e = bs_fft(D, &a, &b, cs, L.num_sms, L.stream);
		if (e != cudaSuccess) return e;
		dim3 gp(bs_grid_x(P.n / 4 + 1, cs, L.num_sms), (unsigned) cs);
		bs_post_kernel<<<gp, BS_THREADS, 0, L.stream>>>(P, D, a, L.d_st, s0);
	}
	return cudaGetLastError();
}
