/*
 * tails.h - per-configuration floating point constants of the p-value tails, evaluated once on the
 * host with the reference's expressions and glibc (driver.c:304-316, runs.c:138-139, rank.c:131-170,
 * discreteFourierTransform.c:141-142, nonOverlappingTemplateMatchings.c:479-480, universal.c:385-386,
 * randomExcursions.c:187-207) and handed to the tail kernel by value.
 */
#pragma once

struct TailConsts {
	double sqrt2, log2_, sqrtn;
	double two_over_sqrtn, sqrt2n;		/* runs.c:138-139 */
	double lgam_bf;				/* lgamma(N/2) of BlockFrequency */
	double lgam_3, lgam_4, lgam_2_5;	/* lgamma(3.0), lgamma(4.0), lgamma(2.5) */
	double lgam_apen;			/* lgamma(2^(m-1)) */
	double lgam_serial1, lgam_serial2;	/* lgamma(2^(m-1)/2), lgamma(2^(m-2)/2) */
	double p_32, p_31, p_30;
	double lr_pi[7];
	double sqrtn4_095_005;
	double nonov_mu, nonov_sqrt_sigma2;
	double ov_pi[6];
	double univ_expected, univ_sigma;
	double rex_pi[4][6];
	double lc_pi[7];
};
