// xo_math.cuh -- elementary functions of the photon-packet kernels (sm_100a).
//
// Two bindings of the reference's math macro layer (mc_log, mc_sincos, mc_fdiv,
// mc_sqrt, mc_pow, mc_cbrt ..., xopto/mcbase/kernel/mcbase.template.h:522-644):
//
//  * FastMath  -- throughput mode.  One MUFU per transcendental (lg2/ex2/sin/cos/
//    rcp/rsq), FMA contraction on.  Mirrors the reference's own validated mode
//    (-cl-fast-relaxed-math + MC_USE_NATIVE_MATH, mcml/test/validate.py:1838).
//  * DetMath   -- deterministic parity mode.  Every function is specified purely
//    by IEEE-754 binary64 +,-,*,/,sqrt (round-to-nearest, no contraction) and a
//    final rounding to binary32, written with __dadd_rn/__dmul_rn/__ddiv_rn so
//    the result does not depend on compiler flags.  The specification (range
//    reductions, polynomial degrees, constants) is the one in the header comment
//    of the CPU oracle's portable math; results are bit-identical to it and
//    within 1 ulp of glibc/OpenCL built-ins.
#pragma once

#ifndef XO_DOUBLE
#define XO_DOUBLE 0
#endif
// literal in the kernels' floating-point type
#if XO_DOUBLE
#define XO_FP(x) x
#include "xo_math_double.cuh"
#else
#define XO_FP(x) x##f

namespace xo {

#define XO_INF __int_as_float(0x7f800000)
#define XO_NAN __int_as_float(0x7fffffff)
#define XO_FLT_MAX 3.402823466e+38f

// ---------------------------------------------------------------------------
struct FastMath {
	static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }
	static __device__ __forceinline__ float rcp(float a) { return __frcp_rn(a); }
	// one MUFU.RCP
	static __device__ __forceinline__ float rcp_approx(float a) {
		float r;
		asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
		return r;
	}
	// one MUFU.LG2 (callers pass normal numbers or 0)
	static __device__ __forceinline__ float lg2(float a) {
		float r;
		asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
		return r;
	}
	// sqrt.approx = one MUFU.SQRT (the IEEE version costs ~10 instructions)
	static __device__ __forceinline__ float sqrt(float a) {
		float r;
		asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
		return r;
	}
	static __device__ __forceinline__ float rsqrt(float a) { return rsqrtf(a); }
	static __device__ __forceinline__ float log(float a) { return __logf(a); }
	static __device__ __forceinline__ float exp(float a) { return __expf(a); }
	static __device__ __forceinline__ float pow(float a, float b) { return __powf(a, b); }
	// |a|^(1/3) through MUFU.LG2 / MUFU.EX2 (2 MUFU + 3 ALU instead of ~25)
	static __device__ __forceinline__ float cbrt(float a) {
		float r = exp2f(__log2f(fabsf(a))*0.3333333432674408f);
		return copysignf(r, a);
	}
	static __device__ __forceinline__ float atan2(float y, float x) { return atan2f(y, x); }
	static __device__ __forceinline__ void sincos(float a, float *s, float *c) { __sincosf(a, s, c); }
	// a*b + c with contraction allowed
	static __device__ __forceinline__ float mad(float a, float b, float c) { return fmaf(a, b, c); }
};

// ---------------------------------------------------------------------------
namespace det {
#define XO_LN2      0.6931471805599453094
#define XO_LN2_HI   6.93147180369123816490e-01
#define XO_LN2_LO   1.90821492927058770002e-10
#define XO_INVLN2   1.44269504088896338700
#define XO_PIO2_HI  1.57079632673412561417e+00
#define XO_PIO2_LO  6.07710050650619224932e-11
#define XO_2_OVER_PI 0.63661977236758134308
#define XO_SQRT2    1.41421356237309504880
#define XO_PI_D     3.14159265358979323846
#define XO_PIO2_D   1.57079632679489661923

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double quo(double a, double b) { return __ddiv_rn(a, b); }
// Horner step p*z + c as two separately rounded operations
__device__ __forceinline__ double hs(double p, double z, double c) { return add(mul(p, z), c); }

__device__ inline double dlog_pos(double x) {
	unsigned long long b = (unsigned long long)__double_as_longlong(x);
	int e = (int)((b >> 52) & 0x7ff) - 1023;
	double m = __longlong_as_double((long long)((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL));
	if (m > XO_SQRT2) { m = mul(m, 0.5); e += 1; }
	double s = quo(sub(m, 1.0), add(m, 1.0));
	double z = mul(s, s);
	double p = 1.0/23.0;
	p = hs(p, z, 1.0/21.0);
	p = hs(p, z, 1.0/19.0);
	p = hs(p, z, 1.0/17.0);
	p = hs(p, z, 1.0/15.0);
	p = hs(p, z, 1.0/13.0);
	p = hs(p, z, 1.0/11.0);
	p = hs(p, z, 1.0/9.0);
	p = hs(p, z, 1.0/7.0);
	p = hs(p, z, 1.0/5.0);
	p = hs(p, z, 1.0/3.0);
	p = hs(p, z, 1.0);
	return add(mul((double)e, XO_LN2), mul(mul(2.0, s), p));
}

__device__ inline double dexp(double t) {
	double n = rint(mul(t, XO_INVLN2));
	double r = sub(sub(t, mul(n, XO_LN2_HI)), mul(n, XO_LN2_LO));
	double p = 1.0/6227020800.0;
	p = hs(p, r, 1.0/479001600.0);
	p = hs(p, r, 1.0/39916800.0);
	p = hs(p, r, 1.0/3628800.0);
	p = hs(p, r, 1.0/362880.0);
	p = hs(p, r, 1.0/40320.0);
	p = hs(p, r, 1.0/5040.0);
	p = hs(p, r, 1.0/720.0);
	p = hs(p, r, 1.0/120.0);
	p = hs(p, r, 1.0/24.0);
	p = hs(p, r, 1.0/6.0);
	p = hs(p, r, 0.5);
	p = hs(p, r, 1.0);
	p = hs(p, r, 1.0);
	long long ni = (long long)n;
	double scale = __longlong_as_double((ni + 1023) << 52);
	return mul(p, scale);
}

__device__ inline float logf_(float xf) {
	if (xf != xf) return xf;
	if (xf < 0.0f) return XO_NAN;
	if (xf == 0.0f) return -XO_INF;
	if (xf == XO_INF) return xf;
	return __double2float_rn(dlog_pos((double)xf));
}

__device__ inline float expf_(float xf) {
	if (xf != xf) return xf;
	if (xf > 89.0f) return XO_INF;
	if (xf < -104.0f) return 0.0f;
	return __double2float_rn(dexp((double)xf));
}

__device__ inline float powf_(float xf, float yf) {
	if (xf != xf || yf != yf) return XO_NAN;
	if (yf == 0.0f) return 1.0f;
	if (xf == 0.0f) return (yf > 0.0f) ? 0.0f : XO_INF;
	if (xf < 0.0f) return XO_NAN;
	if (xf == XO_INF) return (yf > 0.0f) ? XO_INF : 0.0f;
	double t = mul((double)yf, dlog_pos((double)xf));
	if (t > 89.0) return XO_INF;
	if (t < -104.0) return 0.0f;
	return __double2float_rn(dexp(t));
}

__device__ inline void dsincos_reduced(double r, double *s, double *c) {
	double z = mul(r, r);
	double ps = 1.0/355687428096000.0;
	ps = hs(ps, z, -(1.0/1307674368000.0));
	ps = hs(ps, z, 1.0/6227020800.0);
	ps = hs(ps, z, -(1.0/39916800.0));
	ps = hs(ps, z, 1.0/362880.0);
	ps = hs(ps, z, -(1.0/5040.0));
	ps = hs(ps, z, 1.0/120.0);
	ps = hs(ps, z, -(1.0/6.0));
	ps = hs(ps, z, 1.0);
	*s = mul(r, ps);
	double pc = -1.0/6402373705728000.0;
	pc = hs(pc, z, 1.0/20922789888000.0);
	pc = hs(pc, z, -(1.0/87178291200.0));
	pc = hs(pc, z, 1.0/479001600.0);
	pc = hs(pc, z, -(1.0/3628800.0));
	pc = hs(pc, z, 1.0/40320.0);
	pc = hs(pc, z, -(1.0/720.0));
	pc = hs(pc, z, 1.0/24.0);
	pc = hs(pc, z, -0.5);
	pc = hs(pc, z, 1.0);
	*c = pc;
}

__device__ inline void sincosf_(float xf, float *sn, float *cs) {
	if (xf != xf || fabsf(xf) == XO_INF) { *sn = XO_NAN; *cs = XO_NAN; return; }
	double x = (double)xf;
	double k = rint(mul(x, XO_2_OVER_PI));
	double r = sub(sub(x, mul(k, XO_PIO2_HI)), mul(k, XO_PIO2_LO));
	double s, c;
	dsincos_reduced(r, &s, &c);
	int q = (int)((long long)k & 3);
	double ss, cc;
	if (q == 0) { ss = s; cc = c; }
	else if (q == 1) { ss = c; cc = -s; }
	else if (q == 2) { ss = -s; cc = -c; }
	else { ss = -c; cc = s; }
	*sn = __double2float_rn(ss);
	*cs = __double2float_rn(cc);
}

__device__ inline float cbrtf_(float xf) {
	if (xf != xf || xf == 0.0f || fabsf(xf) == XO_INF) return xf;
	double x = fabs((double)xf);
	unsigned long long b = (unsigned long long)__double_as_longlong(x);
	int e = (int)((b >> 52) & 0x7ff) - 1023;
	int q = (e >= 0) ? e/3 : -((2 - e)/3);
	int rem = e - 3*q;
	double m = __longlong_as_double((long long)((b & 0x000fffffffffffffULL) |
		((unsigned long long)(1023 + rem) << 52)));
	double y = add(0.857142857142857142, mul(0.142857142857142857, m));
#pragma unroll 1
	for (int i = 0; i < 6; ++i)
		y = sub(y, quo(sub(mul(mul(y, y), y), m), mul(3.0, mul(y, y))));
	y = mul(y, __longlong_as_double((long long)(1023 + q) << 52));
	return __double2float_rn((xf < 0.0f) ? -y : y);
}

__device__ inline double datan_unit(double t) {
	double t1 = quo(t, add(1.0, __dsqrt_rn(add(1.0, mul(t, t)))));
	double t2 = quo(t1, add(1.0, __dsqrt_rn(add(1.0, mul(t1, t1)))));
	double z = mul(t2, t2);
	double p = 1.0/31.0;
	p = sub(1.0/29.0, mul(p, z));
	p = sub(1.0/27.0, mul(p, z));
	p = sub(1.0/25.0, mul(p, z));
	p = sub(1.0/23.0, mul(p, z));
	p = sub(1.0/21.0, mul(p, z));
	p = sub(1.0/19.0, mul(p, z));
	p = sub(1.0/17.0, mul(p, z));
	p = sub(1.0/15.0, mul(p, z));
	p = sub(1.0/13.0, mul(p, z));
	p = sub(1.0/11.0, mul(p, z));
	p = sub(1.0/9.0, mul(p, z));
	p = sub(1.0/7.0, mul(p, z));
	p = sub(1.0/5.0, mul(p, z));
	p = sub(1.0/3.0, mul(p, z));
	p = sub(1.0, mul(p, z));
	return mul(4.0, mul(t2, p));
}

__device__ inline float atan2f_(float yf, float xf) {
	if (xf != xf || yf != yf) return XO_NAN;
	double y = (double)yf, x = (double)xf;
	double ax = fabs(x), ay = fabs(y);
	double a;
	const double dinf = __longlong_as_double(0x7ff0000000000000LL);
	if (ax == 0.0 && ay == 0.0) a = 0.0;
	else if (ax == dinf && ay == dinf) a = mul(0.25, XO_PI_D);
	else if (ay <= ax) a = datan_unit(quo(ay, ax));
	else a = sub(XO_PIO2_D, datan_unit(quo(ax, ay)));
	if (__float_as_int(xf) < 0) a = sub(XO_PI_D, a);
	if (__float_as_int(yf) < 0) a = -a;
	return __double2float_rn(a);
}
}  // namespace det

struct DetMath {
	static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
	static __device__ __forceinline__ float rcp(float a) { return __fdiv_rn(1.0f, a); }
	static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
	static __device__ __forceinline__ float rsqrt(float a) { return __fdiv_rn(1.0f, __fsqrt_rn(a)); }
	static __device__ __forceinline__ float log(float a) { return det::logf_(a); }
	static __device__ __forceinline__ float exp(float a) { return det::expf_(a); }
	static __device__ __forceinline__ float pow(float a, float b) { return det::powf_(a, b); }
	static __device__ __forceinline__ float cbrt(float a) { return det::cbrtf_(a); }
	static __device__ __forceinline__ float atan2(float y, float x) { return det::atan2f_(y, x); }
	static __device__ __forceinline__ void sincos(float a, float *s, float *c) { det::sincosf_(a, s, c); }
	static __device__ __forceinline__ float mad(float a, float b, float c) { return __fadd_rn(__fmul_rn(a, b), c); }
};

}  // namespace xo

#endif  // !XO_DOUBLE
