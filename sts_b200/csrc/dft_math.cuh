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

