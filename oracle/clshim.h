/* oracle/clshim.h -- TEST INFRASTRUCTURE (never shipped, never on the product path).
 *
 * Lets gcc compile the OpenCL-C 1.2 translation unit that the *reference's own
 * Python* renders (mcml/mc.py:531-626 `_build_src`), unchanged, as plain C, so the
 * reference kernel itself can be executed on CPU cores.
 * Maps address-space qualifiers to nothing, OpenCL scalar/vector typedefs to C
 * structs (3-vectors are 16 bytes as in OpenCL), work-item queries to a
 * thread-local, OpenCL atomics to GCC __atomic builtins and the overloaded math
 * built-ins to their single-precision libm versions.
 */
#ifndef XO_ORACLE_CLSHIM_H
#define XO_ORACLE_CLSHIM_H
#include <math.h>
#include <float.h>
#include <stdio.h>
#include <stdbool.h>
#include <stddef.h>
#include <string.h>

#define __kernel
#define __global
#define __constant
#define __local
#define __private
#define __OPENCL_VERSION__ 120
#define __OPENCL_C_VERSION__ 120

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long ulong;

#define XO_V2(T)  typedef struct { T x, y; } T##2;
#define XO_V3(T)  typedef struct { T x, y, z, w; } T##3;
#define XO_V4(T)  typedef struct { T x, y, z, w; } T##4;
#define XO_VN(T, N) typedef struct { T s[N]; } T##N;
#define XO_VECTORS(T) XO_V2(T) XO_V3(T) XO_V4(T) XO_VN(T, 8) XO_VN(T, 16)
XO_VECTORS(char) XO_VECTORS(uchar) XO_VECTORS(short) XO_VECTORS(ushort)
XO_VECTORS(int) XO_VECTORS(uint) XO_VECTORS(long) XO_VECTORS(ulong)
XO_VECTORS(float) XO_VECTORS(double)

/* work-item id: set by the driver before each call of the kernel function */
extern _Thread_local size_t xo_ref_global_id;
static inline size_t get_global_id(int dim) { (void)dim; return xo_ref_global_id; }

static inline uint atomic_inc(volatile uint *p) {
	return __atomic_fetch_add(p, 1u, __ATOMIC_RELAXED);
}
static inline uint atomic_add(volatile uint *p, uint v) {
	return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
static inline uint atomic_cmpxchg(volatile uint *p, uint cmp, uint val) {
	__atomic_compare_exchange_n(p, &cmp, val, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
	return cmp;
}
static inline ulong atom_inc(volatile ulong *p) {
	return __atomic_fetch_add(p, 1ul, __ATOMIC_RELAXED);
}
static inline ulong atom_add(volatile ulong *p, ulong v) {
	return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}

#define convert_int(x)   ((int)(x))
#define convert_uint(x)  ((uint)(x))

/* (with -DXO_REF_DOUBLE the rendered kernel computes in binary64: the bare OpenCL math
 * names then are libm's double functions as they stand) */
#ifdef XO_REF_DOUBLE
static inline double xo_ref_sincos(double x, double *c) { *c = cos(x); return sin(x); }
#define sincos(x, pc)    xo_ref_sincos((x), (pc))
#define rsqrt(x)         (1.0/sqrt(x))
#define powr(x, y)       pow((x), (y))
#define native_sin       sin
#define native_cos       cos
#define native_log       log
#define native_exp       exp
#define native_sqrt      sqrt
#define native_rsqrt(x)  (1.0/sqrt(x))
#define native_powr      pow
#else
static inline float xo_ref_sincos(float x, float *c) { *c = cosf(x); return sinf(x); }
#define sincos(x, pc)    xo_ref_sincos((x), (pc))
#define rsqrt(x)         (1.0f/sqrtf(x))
#define powr(x, y)       powf((x), (y))
#define native_sin       sinf
#define native_cos       cosf
#define native_log       logf
#define native_exp       expf
#define native_sqrt      sqrtf
#define native_rsqrt(x)  (1.0f/sqrtf(x))
#define native_powr      powf
#endif
#define native_divide(a, b) ((a)/(b))
/* OpenCL's min / max / clamp are functions: every argument is evaluated once
 * (mcvox/mcsource/voxel.py draws a random number inside mc_min(...)) */
#define clamp(x, lo, hi) ({ __typeof__(x) x_ = (x); __typeof__(lo) lo_ = (lo); __typeof__(hi) hi_ = (hi); \
	x_ < lo_ ? lo_ : (x_ > hi_ ? hi_ : x_); })
#define min(a, b)        ({ __typeof__(a) a_ = (a); __typeof__(b) b_ = (b); a_ < b_ ? a_ : b_; })
#define max(a, b)        ({ __typeof__(a) a_ = (a); __typeof__(b) b_ = (b); a_ > b_ ? a_ : b_; })

/* OpenCL math built-ins are overloaded on float; the rendered text uses the
 * bare names with float arguments. */
#ifndef XO_REF_DOUBLE
#define sin sinf
#define cos cosf
#define tan tanf
#define sqrt sqrtf
#define log logf
#define exp expf
#define fabs fabsf
#define fmin fminf
#define fmax fmaxf
#define floor floorf
#define ceil ceilf
#define cbrt cbrtf
#define copysign copysignf
#define asin asinf
#define acos acosf
#define atan atanf
#define atan2 atan2f
#define tanh tanhf
#define pow powf
#define fmod fmodf
#define round roundf
#endif
#endif
