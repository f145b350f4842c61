/* dft_math.cuh - complex helpers and the radix 2/4/8 butterflies shared by the FFT kernels */
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
	return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmi(double2 a) { return make_double2(a.y, -a.x); }	/* times -i */

template <int R> __device__ __forceinline__ void bfly(double2 *v);

template <> __device__ __forceinline__ void bfly<2>(double2 *v)
{
	double2 a = v[0], b = v[1];
	v[0] = cadd(a, b);
	v[1] = csub(a, b);
}
template <> __device__ __forceinline__ void bfly<4>(double2 *v)
{
	double2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
	double2 a2 = cadd(v[1], v[3]), a3 = cmi(csub(v[1], v[3]));
	v[0] = cadd(a0, a2);
	v[1] = cadd(a1, a3);
	v[2] = csub(a0, a2);
	v[3] = csub(a1, a3);
}
template <> __device__ __forceinline__ void bfly<8>(double2 *v)
{
	const double s = 0.70710678118654752440;
	double2 a0 = cadd(v[0], v[4]), a1 = csub(v[0], v[4]);
	double2 a2 = cadd(v[2], v[6]), a3 = cmi(csub(v[2], v[6]));
	double2 a4 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
	double2 a6 = cadd(v[3], v[7]), a7 = cmi(csub(v[3], v[7]));
	double2 b0 = cadd(a0, a2), b2 = csub(a0, a2), b1 = cadd(a1, a3), b3 = csub(a1, a3);
	double2 b4 = cadd(a4, a6), b6 = cmi(csub(a4, a6)), b5 = cadd(a5, a7), b7 = csub(a5, a7);
	double2 t5 = make_double2(s * (b5.x + b5.y), s * (b5.y - b5.x));	/* b5 * (1 - i)/sqrt2 */
	double2 t7 = make_double2(s * (b7.y - b7.x), -s * (b7.x + b7.y));	/* b7 * (-1 - i)/sqrt2 */
	v[0] = cadd(b0, b4);
	v[1] = cadd(b1, t5);
	v[2] = cadd(b2, b6);
	v[3] = cadd(b3, t7);
	v[4] = csub(b0, b4);
	v[5] = csub(b1, t5);
	v[6] = csub(b2, b6);
	v[7] = csub(b3, t7);
}

/* padded shared-memory index: one spare element per 8 keeps radix-8 scatter stores conflict free */
__device__ __forceinline__ int padi(int e) { return e + (e >> 3); }

template <> __device__ __forceinline__ void bfly<3>(double2 *v)
{
	const double s = 0.86602540378443864676;
	const double2 a = cadd(v[1], v[2]), b = csub(v[1], v[2]);
	const double2 m = make_double2(v[0].x - 0.5 * a.x, v[0].y - 0.5 * a.y);
	const double2 n = make_double2(s * b.y, -s * b.x);		/* -i s b */
	v[0] = cadd(v[0], a);
	v[1] = cadd(m, n);
	v[2] = csub(m, n);
}
template <> __device__ __forceinline__ void bfly<5>(double2 *v)
{
	const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;
	const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;
	const double2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]), b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
	const double2 m1 = make_double2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
	const double2 m2 = make_double2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
	const double2 n1 = make_double2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
	const double2 n2 = make_double2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
	v[0] = cadd(v[0], cadd(a1, a2));
	v[1] = make_double2(m1.x + n1.y, m1.y - n1.x);		/* m1 - i n1 */
	v[4] = make_double2(m1.x - n1.y, m1.y + n1.x);
	v[2] = make_double2(m2.x + n2.y, m2.y - n2.x);
	v[3] = make_double2(m2.x - n2.y, m2.y + n2.x);
}

/* 10-point DFT in registers, natural order in and out.  10 = 2 x 5 with coprime factors, so the prime-factor
 * index maps j = (5 j1 + 2 j2) mod 10, k = (5 k1 + 6 k2) mod 10 leave no twiddles between the two stages */
__device__ __forceinline__ void bfly10(double2 *v)
{
	double2 a0[5] = { v[0], v[2], v[4], v[6], v[8] }, a1[5] = { v[5], v[7], v[9], v[1], v[3] };
	bfly<5>(a0);
	bfly<5>(a1);
	v[0] = cadd(a0[0], a1[0]); v[5] = csub(a0[0], a1[0]);
	v[6] = cadd(a0[1], a1[1]); v[1] = csub(a0[1], a1[1]);
	v[2] = cadd(a0[2], a1[2]); v[7] = csub(a0[2], a1[2]);
	v[8] = cadd(a0[3], a1[3]); v[3] = csub(a0[3], a1[3]);
	v[4] = cadd(a0[4], a1[4]); v[9] = csub(a0[4], a1[4]);
}

/* X_q and X_{N-q} from Z_q = zq and Z_{N-q} = zp; wq = exp(-2 pi i q / n); returns how many of the two lie below T */
__device__ __forceinline__ unsigned int post_pair(double2 zq, double2 zp, double2 wq, double T2x4, bool both)
{
	const double2 sum = make_double2(zq.x + zp.x, zq.y - zp.y);	/* Z_q + conj Z_{N-q} */
	const double2 dif = make_double2(zq.x - zp.x, zq.y + zp.y);	/* Z_q - conj Z_{N-q} */
	const double2 wd = cmul(wq, dif);
	const double ar = sum.x + wd.y, ai = sum.y - wd.x;		/* 2 X_q */
	const double br = sum.x - wd.y, bi = sum.y + wd.x;		/* 2 conj X_{N-q} */
	/* |X| < T  <=>  |2X|^2 < 4 T^2 up to one rounding, far inside the 1e-9 band the north star allows */
	const bool below_a = (ar * ar + ai * ai) < T2x4, below_b = (br * br + bi * bi) < T2x4;
	return (below_a ? 1u : 0u) + ((both && below_b) ? 1u : 0u);
}
