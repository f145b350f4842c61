/*
 * dft16.cuh - what the two n = 2^20 FFT kernel files share (k_dft_fast.cu: column and row pass as two launches per
 * chunk; k_dft_fused.cu: one persistent kernel): the 256 x 2048 split, the padded row layout and the radix-16
 * butterfly in registers.
 */
#pragma once
#include "dft_math.cuh"

#define QN1 256
#define QN2 2048
#define QROW (QN2 + QN2 / 16)		/* padded row: one spare element per 16 */

__device__ __forceinline__ int pad16(int e) { return e + (e >> 4); }

/* 16-point DFT in registers, natural order in and out: j = j0 + 4 j1, k = k1 + 4 k0 */
__device__ __forceinline__ void bfly16(double2 *v)
{
	const double c8 = 0.92387953251128675613, s8 = 0.38268343236508977173, r2 = 0.70710678118654752440;
	double2 a[16];
#pragma unroll
	for (int j0 = 0; j0 < 4; j0++) {
		double2 t[4] = { v[j0], v[j0 + 4], v[j0 + 8], v[j0 + 12] };
		bfly<4>(t);
#pragma unroll
		for (int k1 = 0; k1 < 4; k1++)
			a[j0 * 4 + k1] = t[k1];
	}
	/* times W_16^(j0 k1) */
	a[5] = cmul(a[5], make_double2(c8, -s8));			/* W^1 */
	a[6] = make_double2(r2 * (a[6].x + a[6].y), r2 * (a[6].y - a[6].x));	/* W^2 */
	a[7] = cmul(a[7], make_double2(s8, -c8));			/* W^3 */
	a[9] = make_double2(r2 * (a[9].x + a[9].y), r2 * (a[9].y - a[9].x));	/* W^2 */
	a[10] = cmi(a[10]);						/* W^4 = -i */
	a[11] = make_double2(r2 * (a[11].y - a[11].x), -r2 * (a[11].x + a[11].y));	/* W^6 */
	a[13] = cmul(a[13], make_double2(s8, -c8));			/* W^3 */
	a[14] = make_double2(r2 * (a[14].y - a[14].x), -r2 * (a[14].x + a[14].y));	/* W^6 */
	a[15] = cmul(a[15], make_double2(-c8, s8));			/* W^9 */
#pragma unroll
	for (int k1 = 0; k1 < 4; k1++) {
		double2 t[4] = { a[k1], a[4 + k1], a[8 + k1], a[12 + k1] };
		bfly<4>(t);
#pragma unroll
		for (int k0 = 0; k0 < 4; k0++)
			v[k1 + 4 * k0] = t[k0];
	}
}

/* x times W_16^r = exp(-2 pi i r / 16), r = 0..7 (compile-time r after unrolling) */
__device__ __forceinline__ double2 cmul_w16(double2 x, int r)
{
	const double c8 = 0.92387953251128675613, s8 = 0.38268343236508977173, r2 = 0.70710678118654752440;
	switch (r) {
	case 0: return x;
	case 1: return cmul(x, make_double2(c8, -s8));
	case 2: return make_double2(r2 * (x.x + x.y), r2 * (x.y - x.x));
	case 3: return cmul(x, make_double2(s8, -c8));
	case 4: return cmi(x);
	case 5: return cmul(x, make_double2(-s8, -c8));
	case 6: return make_double2(r2 * (x.y - x.x), -r2 * (x.x + x.y));
	default: return cmul(x, make_double2(-c8, -s8));
	}
}

