/* oracle/xo_detmath.h -- TEST INFRASTRUCTURE (CPU oracle only).
 *
 * "Portable" single-precision elementary functions for the deterministic parity
 * mode.  OpenCL leaves log/sincos/cbrt/powr/exp accurate only to a few ulp and
 * implementation defined (the reference calls the bare built-ins,
 * mcbase.template.h:576-594), so *any* faithful implementation is a legal
 * reference outcome.  These versions are specified purely in terms of IEEE-754
 * binary64 +,-,*,/ (round-to-nearest-even, no contraction) followed by one
 * rounding to binary32, so a C build (-ffp-contract=off) and the CUDA
 * deterministic kernel (pyxopto_b200/csrc/kernels/xo_math.cuh, written with
 * __dadd_rn/__dmul_rn/__ddiv_rn) produce bit-identical results.  The double
 * evaluation error is < 1e-14 relative, i.e. results are correctly rounded
 * except in ~1e-7 of cases, and agree with glibc to <= 1 ulp.
 *
 * The algorithms are restated independently in the .cuh; the two files share
 * nothing but this specification:
 *   log : x = m*2^e, m in (sqrt(.5), sqrt(2)]; s=(m-1)/(m+1);
 *         log x = e*LN2 + 2*s*P(s^2), P = sum_{k=0}^{11} z^k/(2k+1) (Horner)
 *   exp : n = rint(t*INVLN2); r = (t - n*LN2_HI) - n*LN2_LO;
 *         e^r = sum_{k=0}^{13} r^k/k! (Horner), result scaled by 2^n
 *   sincos: k = rint(x*TWO_OVER_PI); r = (x - k*PIO2_HI) - k*PIO2_LO;
 *         sin r = r*S(r^2) deg 8 in r^2 (Taylor to r^17), cos r = C(r^2) deg 9
 *         (Taylor to r^18); quadrant from k&3
 *   cbrt: y0 from exponent/3 + linear mantissa seed, 6 Newton steps in binary64
 *   pow : exp(y*log(x)) with the binary64 log/exp above (x > 0)
 *   atan2: octant reduction + atan(t), t in [0,1]: argument halving
 *         t' = t/(1+sqrt(1+t^2)) twice then Taylor (odd, 16 terms)
 */
#ifndef XO_DETMATH_H
#define XO_DETMATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline double xo_d_from_bits(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
static inline uint64_t xo_d_bits(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }

#define XO_LN2      0.6931471805599453094
#define XO_LN2_HI   6.93147180369123816490e-01  /* low 32 bits of mantissa zero */
#define XO_LN2_LO   1.90821492927058770002e-10
#define XO_INVLN2   1.44269504088896338700
#define XO_PIO2_HI  1.57079632673412561417e+00  /* 33 significant bits */
#define XO_PIO2_LO  6.07710050650619224932e-11
#define XO_2_OVER_PI 0.63661977236758134308
#define XO_SQRT2    1.41421356237309504880
#define XO_PI_D     3.14159265358979323846
#define XO_PIO2_D   1.57079632679489661923

/* binary64 log of a positive finite normal double */
static inline double xo_dlog_pos(double x) {
	uint64_t b = xo_d_bits(x);
	int e = (int)((b >> 52) & 0x7ff) - 1023;
	double m = xo_d_from_bits((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);
	if (m > XO_SQRT2) { m = m*0.5; e += 1; }
	double s = (m - 1.0)/(m + 1.0);
	double z = s*s;
	double p = 1.0/23.0;
	p = p*z + 1.0/21.0;
	p = p*z + 1.0/19.0;
	p = p*z + 1.0/17.0;
	p = p*z + 1.0/15.0;
	p = p*z + 1.0/13.0;
	p = p*z + 1.0/11.0;
	p = p*z + 1.0/9.0;
	p = p*z + 1.0/7.0;
	p = p*z + 1.0/5.0;
	p = p*z + 1.0/3.0;
	p = p*z + 1.0;
	return (double)e*XO_LN2 + (2.0*s)*p;
}

/* binary64 exp for |t| < 700 */
static inline double xo_dexp(double t) {
	double n = rint(t*XO_INVLN2);
	double r = (t - n*XO_LN2_HI) - n*XO_LN2_LO;
	double p = 1.0/6227020800.0;          /* 1/13! */
	p = p*r + 1.0/479001600.0;
	p = p*r + 1.0/39916800.0;
	p = p*r + 1.0/3628800.0;
	p = p*r + 1.0/362880.0;
	p = p*r + 1.0/40320.0;
	p = p*r + 1.0/5040.0;
	p = p*r + 1.0/720.0;
	p = p*r + 1.0/120.0;
	p = p*r + 1.0/24.0;
	p = p*r + 1.0/6.0;
	p = p*r + 0.5;
	p = p*r + 1.0;
	p = p*r + 1.0;
	int64_t ni = (int64_t)n;
	double scale = xo_d_from_bits((uint64_t)(ni + 1023) << 52);
	return p*scale;
}

static inline float xo_logf(float xf) {
	if (xf != xf) return xf;
	if (xf < 0.0f) return NAN;
	if (xf == 0.0f) return -INFINITY;
	if (isinf(xf)) return xf;
	return (float)xo_dlog_pos((double)xf);
}

static inline float xo_expf(float xf) {
	if (xf != xf) return xf;
	if (xf > 89.0f) return INFINITY;
	if (xf < -104.0f) return 0.0f;
	return (float)xo_dexp((double)xf);
}

/* OpenCL powr semantics on the domain the path uses (x >= 0) */
static inline float xo_powf(float xf, float yf) {
	if (xf != xf || yf != yf) return NAN;
	if (yf == 0.0f) return 1.0f;
	if (xf == 0.0f) return (yf > 0.0f) ? 0.0f : INFINITY;
	if (xf < 0.0f) return NAN;
	if (isinf(xf)) return (yf > 0.0f) ? INFINITY : 0.0f;
	double t = (double)yf*xo_dlog_pos((double)xf);
	if (t > 89.0) return INFINITY;
	if (t < -104.0) return 0.0f;
	return (float)xo_dexp(t);
}

static inline void xo_dsincos_reduced(double r, double *s, double *c) {
	double z = r*r;
	double ps = 1.0/355687428096000.0;    /* 1/17! */
	ps = ps*z - 1.0/1307674368000.0;      /* 1/15! */
	ps = ps*z + 1.0/6227020800.0;         /* 1/13! */
	ps = ps*z - 1.0/39916800.0;           /* 1/11! */
	ps = ps*z + 1.0/362880.0;             /* 1/9!  */
	ps = ps*z - 1.0/5040.0;               /* 1/7!  */
	ps = ps*z + 1.0/120.0;                /* 1/5!  */
	ps = ps*z - 1.0/6.0;                  /* 1/3!  */
	ps = ps*z + 1.0;
	*s = r*ps;
	double pc = -1.0/6402373705728000.0;  /* 1/18! */
	pc = pc*z + 1.0/20922789888000.0;     /* 1/16! */
	pc = pc*z - 1.0/87178291200.0;        /* 1/14! */
	pc = pc*z + 1.0/479001600.0;          /* 1/12! */
	pc = pc*z - 1.0/3628800.0;            /* 1/10! */
	pc = pc*z + 1.0/40320.0;              /* 1/8!  */
	pc = pc*z - 1.0/720.0;                /* 1/6!  */
	pc = pc*z + 1.0/24.0;                 /* 1/4!  */
	pc = pc*z - 0.5;
	pc = pc*z + 1.0;
	*c = pc;
}

/* returns sin(x), stores cos(x); valid for |x| < 1e6 (the path uses [0, 2pi]) */
static inline float xo_sincosf(float xf, float *cosout) {
	if (xf != xf || isinf(xf)) { *cosout = NAN; return NAN; }
	double x = (double)xf;
	double k = rint(x*XO_2_OVER_PI);
	double r = (x - k*XO_PIO2_HI) - k*XO_PIO2_LO;
	double s, c;
	xo_dsincos_reduced(r, &s, &c);
	int q = (int)((int64_t)k & 3);
	double ss, cc;
	switch (q) {
		case 0: ss = s; cc = c; break;
		case 1: ss = c; cc = -s; break;
		case 2: ss = -s; cc = -c; break;
		default: ss = -c; cc = s; break;
	}
	*cosout = (float)cc;
	return (float)ss;
}
static inline float xo_sinf(float x) { float c; return xo_sincosf(x, &c); }
static inline float xo_cosf(float x) { float c; xo_sincosf(x, &c); return c; }

static inline float xo_cbrtf(float xf) {
	if (xf != xf || xf == 0.0f || isinf(xf)) return xf;
	double x = fabs((double)xf);
	uint64_t b = xo_d_bits(x);
	int e = (int)((b >> 52) & 0x7ff) - 1023;
	/* e = 3*q + rem, rem in {0,1,2} (floor division) */
	int q = (e >= 0) ? e/3 : -((2 - e)/3);
	int rem = e - 3*q;
	double m = xo_d_from_bits((b & 0x000fffffffffffffULL) | ((uint64_t)(1023 + rem) << 52));
	/* m in [1, 8): seed with a line through (1,1) and (8,2) */
	double y = 0.857142857142857142 + 0.142857142857142857*m;
	for (int i = 0; i < 6; ++i)
		y = y - (y*y*y - m)/(3.0*(y*y));
	y = y*xo_d_from_bits((uint64_t)(1023 + q) << 52);
	return (float)((xf < 0.0f) ? -y : y);
}

/* atan of t in [0, 1] */
static inline double xo_datan_unit(double t) {
	double t1 = t/(1.0 + sqrt(1.0 + t*t));
	double t2 = t1/(1.0 + sqrt(1.0 + t1*t1));     /* |t2| <= tan(pi/16) */
	double z = t2*t2;
	double p = 1.0/31.0;
	p = 1.0/29.0 - p*z;
	p = 1.0/27.0 - p*z;
	p = 1.0/25.0 - p*z;
	p = 1.0/23.0 - p*z;
	p = 1.0/21.0 - p*z;
	p = 1.0/19.0 - p*z;
	p = 1.0/17.0 - p*z;
	p = 1.0/15.0 - p*z;
	p = 1.0/13.0 - p*z;
	p = 1.0/11.0 - p*z;
	p = 1.0/9.0 - p*z;
	p = 1.0/7.0 - p*z;
	p = 1.0/5.0 - p*z;
	p = 1.0/3.0 - p*z;
	p = 1.0 - p*z;
	return 4.0*(t2*p);
}

static inline float xo_atan2f(float yf, float xf) {
	if (xf != xf || yf != yf) return NAN;
	double y = (double)yf, x = (double)xf;
	double ax = fabs(x), ay = fabs(y);
	double a;
	if (ax == 0.0 && ay == 0.0)
		a = 0.0;
	else if (isinf(ax) && isinf(ay))
		a = 0.25*XO_PI_D;
	else if (ay <= ax)
		a = xo_datan_unit(ay/ax);
	else
		a = XO_PIO2_D - xo_datan_unit(ax/ay);
	if (signbit(xf)) a = XO_PI_D - a;
	if (signbit(yf)) a = -a;
	return (float)a;
}
#endif
