// xo_math_double.cuh -- the kernels in binary64 (McDataTypesDouble, XO_DOUBLE).
//
// The reference switches mc_fp_t to double (mcbase.template.h:324-345: literals through
// FP_LITERAL, the fp_random_double generator, mc_* math macros bound to the double
// built-ins) and compiles the very same kernel text.  This engine does the same with its
// own text: in double mode the token `float` IS double for everything that follows this
// header - packed plugin structs, packet state, lookup tables, trace rows - and the
// single-precision intrinsic names the kernels use map to their double counterparts.  The
// inexact constants are written through XO_FP() / named macros, exact small literals
// (0.5f, 2.0f ...) promote exactly.  Only the reference-structured loops are compiled in
// this mode (the host forces XO_DETERMINISTIC: static schedule, reference expression and
// draw order, no contraction); elementary functions are CUDA's double-precision ones
// (<= 1-2 ulp; the reference's come from the OpenCL runtime / libm).
#pragma once

namespace xo {

#define XO_INF (__longlong_as_double(0x7ff0000000000000LL))
#define XO_NAN (__longlong_as_double(0x7ff8000000000000LL))
#define XO_FLT_MAX 1.7976931348623157e308

struct DoubleMath {
	static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
	static __device__ __forceinline__ double rcp(double a) { return __ddiv_rn(1.0, a); }
	static __device__ __forceinline__ double rcp_approx(double a) { return __ddiv_rn(1.0, a); }
	static __device__ __forceinline__ double lg2(double a) { return ::log2(a); }
	static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
	static __device__ __forceinline__ double rsqrt(double a) { return __ddiv_rn(1.0, __dsqrt_rn(a)); }
	static __device__ __forceinline__ double log(double a) { return ::log(a); }
	static __device__ __forceinline__ double exp(double a) { return ::exp(a); }
	static __device__ __forceinline__ double pow(double a, double b) { return ::pow(a, b); }
	static __device__ __forceinline__ double cbrt(double a) { return ::cbrt(a); }
	static __device__ __forceinline__ double atan2(double y, double x) { return ::atan2(y, x); }
	static __device__ __forceinline__ void sincos(double a, double *s, double *c) { ::sincos(a, s, c); }
	static __device__ __forceinline__ double mad(double a, double b, double c) { return __dadd_rn(__dmul_rn(a, b), c); }
};
typedef DoubleMath FastMath;
typedef DoubleMath DetMath;

}  // namespace xo

// double-typed wrappers: calls with mixed (double, 1.0f) arguments resolve here
namespace xo { namespace dbl {
__device__ __forceinline__ double fma_(double a, double b, double c) { return ::fma(a, b, c); }
__device__ __forceinline__ double fmin_(double a, double b) { return ::fmin(a, b); }
__device__ __forceinline__ double fmax_(double a, double b) { return ::fmax(a, b); }
__device__ __forceinline__ double fabs_(double a) { return ::fabs(a); }
__device__ __forceinline__ double copysign_(double a, double b) { return ::copysign(a, b); }
__device__ __forceinline__ double sqrt_(double a) { return ::sqrt(a); }
__device__ __forceinline__ double rsqrt_(double a) { return ::rsqrt(a); }
__device__ __forceinline__ double floor_(double a) { return ::floor(a); }
__device__ __forceinline__ double round_(double a) { return ::round(a); }
__device__ __forceinline__ double atan_(double a) { return ::atan(a); }
__device__ __forceinline__ double asin_(double a) { return ::asin(a); }
__device__ __forceinline__ double acos_(double a) { return ::acos(a); }
__device__ __forceinline__ double atan2_(double a, double b) { return ::atan2(a, b); }
__device__ __forceinline__ double exp2_(double a) { return ::exp2(a); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ double add_rd(double a, double b) { return __dadd_rd(a, b); }
__device__ __forceinline__ double fma_rd(double a, double b, double c) { return __fma_rd(a, b, c); }
__device__ __forceinline__ int d2i_rz(double a) { return __double2int_rz(a); }
__device__ __forceinline__ int d2i_rd(double a) { return __double2int_rd(a); }
__device__ __forceinline__ unsigned int d2u_rz(double a) { return __double2uint_rz(a); }
} }

// ---- from here on: float is double -----------------------------------------------------
#define float double
#define fmaf xo::dbl::fma_
#define fminf xo::dbl::fmin_
#define fmaxf xo::dbl::fmax_
#define fabsf xo::dbl::fabs_
#define copysignf xo::dbl::copysign_
#define sqrtf xo::dbl::sqrt_
#define rsqrtf xo::dbl::rsqrt_
#define floorf xo::dbl::floor_
#define roundf xo::dbl::round_
#define atanf xo::dbl::atan_
#define asinf xo::dbl::asin_
#define acosf xo::dbl::acos_
#define atan2f xo::dbl::atan2_
#define exp2f xo::dbl::exp2_
#define __fadd_rn xo::dbl::add_rn
#define __fsub_rn xo::dbl::sub_rn
#define __fmul_rn xo::dbl::mul_rn
#define __fdiv_rn xo::dbl::div_rn
#define __fsqrt_rn xo::dbl::sqrt_rn
#define __fadd_rd xo::dbl::add_rd
#define __fmaf_rd xo::dbl::fma_rd
#define __frcp_rn(x) xo::dbl::div_rn(1.0, (x))
#define __float2int_rz xo::dbl::d2i_rz
#define __float2int_rd xo::dbl::d2i_rd
#define __float2uint_rz xo::dbl::d2u_rz
#define __uint2float_rn __uint2double_rn
#define __ull2float_rn __ull2double_rn
