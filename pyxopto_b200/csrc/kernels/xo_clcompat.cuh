// xo_clcompat.cuh -- OpenCL-C compatibility layer for user-written plugin fragments.
//
// The reference assembles its kernel from OpenCL-C text that every plugin object
// contributes (`cl_declaration` / `cl_implementation`, xopto/mcbase/mcobject.py:
// 29-170, fused by mcbase/mcsrc.py:35-93).  The built-in plugins are hand-written
// CUDA here (xo_pf.cuh, xo_detectors.cuh, mcml_sources.cuh ...); a plugin that a
// *user* wrote against the reference's kernel API still carries OpenCL-C text.
// This header lets that text compile unchanged inside the CUDA translation unit:
//
//   * address-space / attribute keywords of OpenCL-C -> nothing;
//   * the scalar, vector and matrix type names of mcbase.template.h:314-451,
//     785-1000,1757-1865 (mc_fp_t, mc_point3f_t, mc_matrix3f_t ...);
//   * the FP_* constants and the mc_* math macro layer (mcbase.template.h:
//     456-644) bound to this engine's two math bindings (xo::M: MUFU math in
//     throughput mode, the portable IEEE set in deterministic mode);
//   * the vector helpers plugin code calls (mc_dot_point3f, transform_point3f,
//     mc_normalize_point3f, reflect, refract, reflectance, cos_critical ...);
//   * a `McSim` facade with the `mcsim_*` accessors of mcml.template.h:426-1108
//     over the register-resident packet state of the CUDA kernels: the adapters
//     in xo_clcompat_glue.cuh build one on the stack around each plugin call, the
//     compiler scalarises it away.
//
// Nothing here is reference code: the names are the reference's API, the bodies
// are this engine's.  Double precision: with XO_DOUBLE the token `float` is
// `double` for everything below (xo_math_double.cuh), so the same text serves both.
// Not covered (documented in DESIGN.md): OpenCL vector swizzles / operators on float3, image and
// work-group built-ins, `printf` debugging (dbg_print* expand to nothing).
#pragma once
#include "xo_core.cuh"

// ---- OpenCL-C keywords -----------------------------------------------------------
#define __kernel
#define __global
#define __constant
#define __local
#define __private
#define __read_only
#define __write_only
#define restrict __restrict__
#define MC_STRUCT_ATTRIBUTES
#define __mc_pf_mem
#define __mc_source_mem
#define __mc_detector_mem
#define __mc_surface_mem
#define __mc_fluence_mem
#define __mc_trace_mem
#define __mc_geometry_mem
#define __mc_material_mem
#define __mc_layer_mem
#define __mc_fp_lut_mem
#define __mc_int_lut_mem
#define __constant_or_global

// ---- scalar types (single precision, 32-bit counters / sizes) ------------------------
typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long long ulong;
#ifndef XO_CLCOMPAT_NO_STDINT
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#endif
typedef float mc_fp_t;
typedef int mc_int_t;
typedef unsigned int mc_uint_t;
typedef unsigned int mc_size_t;
typedef unsigned int mc_cnt_t;
typedef unsigned long long mc_accu_t;

// ---- vectors and matrices (packed structs of scalars, as in the packed host structs) --
typedef xo::P2 mc_fpv2_t;
typedef xo::P3 mc_fpv3_t;
struct mc_fpv4_t { mc_fp_t x, y, z, w; };
struct mc_intv2_t { mc_int_t x, y; };
struct mc_intv3_t { mc_int_t x, y, z; };
struct mc_intv4_t { mc_int_t x, y, z, w; };
struct mc_sizev2_t { mc_size_t x, y; };
struct mc_sizev3_t { mc_size_t x, y, z; };
struct mc_sizev4_t { mc_size_t x, y, z, w; };
typedef mc_fpv2_t mc_point2f_t;
typedef mc_fpv3_t mc_point3f_t;
typedef mc_fpv4_t mc_point4f_t;
typedef mc_intv2_t mc_point2_t;
typedef mc_intv3_t mc_point3_t;
typedef mc_intv4_t mc_point4_t;
typedef mc_sizev2_t mc_point2s_t;
typedef mc_sizev3_t mc_point3s_t;
typedef mc_sizev4_t mc_point4s_t;
struct mc_matrix2_fp_t { mc_fp_t a_11, a_12, a_21, a_22; };
struct mc_matrix3_fp_t { mc_fp_t a_11, a_12, a_13, a_21, a_22, a_23, a_31, a_32, a_33; };
typedef mc_matrix2_fp_t mc_matrix2f_t;
typedef mc_matrix3_fp_t mc_matrix3f_t;
struct mc_rectf_t { mc_point2f_t top_left; mc_fp_t width, height; };
struct mc_circf_t { mc_point2f_t center; mc_fp_t r; };

// lookup-table descriptor (mcbase.template.h:2494-2503)
typedef xo::FpLut mc_fp_lut_t;

// ---- constants ---------------------------------------------------------------------------
#define FP_LITERAL(x) x##f
#define FP_0 0.0f
#define FP_0p25 0.25f
#define FP_0p5 0.5f
#define FP_0p75 0.75f
#define FP_1 1.0f
#define FP_1p5 1.5f
#define FP_2 2.0f
#define FP_2p5 2.5f
#define FP_4 4.0f
#define FP_HALF_PI 1.5707963267948966f
#define FP_PI 3.141592653589793f
#define FP_2PI 6.283185307179586f
#define FP_INV_PI 0.3183098861837907f
#define FP_INV_2PI 0.15915494309189535f
#define FP_COS_0 1.0f
#define FP_COS_90 0.0f
#define FP_COS_CRITICAL_MAX 1.0f
#define FP_COS_CRITICAL_MIN 0.0f
#define FP_EPS 1.1920928955078125e-07f
#define FP_INV_EPS 8388608.0f
#define FP_MAX 3.402823466e+38f
#define FP_INF XO_INF
#define FP_RMIN XO_FP_RMIN
#define FP_PLMIN XO_FP_PLMIN
#define FP_C 299792458.0f
#define FP_INV_C XO_FP_INV_C
#undef MC_INT_ACCUMULATOR_K            // (an integer when it arrives as a compile-time option)
#define MC_INT_ACCUMULATOR_K XO_ACCU_K
#ifndef INFINITY
#define INFINITY XO_INF
#endif

// ---- math macro layer (mcbase.template.h:522-644) ----------------------------------------
namespace xo { namespace clc {
__device__ __forceinline__ float sin_(float x) { float s, c; M::sincos(x, &s, &c); return s; }
__device__ __forceinline__ float cos_(float x) { float s, c; M::sincos(x, &s, &c); return c; }
__device__ __forceinline__ float tan_(float x) { float s, c; M::sincos(x, &s, &c); return M::div(s, c); }
__device__ __forceinline__ int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
}}
#define mc_sqrt(x) xo::M::sqrt((mc_fp_t)(x))
#define mc_rsqrt(x) xo::M::rsqrt((mc_fp_t)(x))
#define mc_log(x) xo::M::log((mc_fp_t)(x))
#define mc_exp(x) xo::M::exp((mc_fp_t)(x))
#define mc_pow(x, y) xo::M::pow((mc_fp_t)(x), (mc_fp_t)(y))
#define mc_cbrt(x) xo::M::cbrt((mc_fp_t)(x))
#define mc_sin(x) xo::clc::sin_((mc_fp_t)(x))
#define mc_cos(x) xo::clc::cos_((mc_fp_t)(x))
#define mc_tan(x) xo::clc::tan_((mc_fp_t)(x))
#define mc_sincos(x, psin, pcos) xo::M::sincos((mc_fp_t)(x), (psin), (pcos))
#define mc_fdiv(a, b) xo::M::div((mc_fp_t)(a), (mc_fp_t)(b))
#define mc_reciprocal(x) xo::M::div(1.0f, (mc_fp_t)(x))
#define mc_atan2(y, x) xo::M::atan2((mc_fp_t)(y), (mc_fp_t)(x))
#define mc_atan(x) atanf((mc_fp_t)(x))
#define mc_asin(x) asinf((mc_fp_t)(x))
#define mc_acos(x) acosf((mc_fp_t)(x))
#define mc_fcopysign(to, from) copysignf((mc_fp_t)(to), (mc_fp_t)(from))
#define mc_fsign(x) (((x) >= FP_0) ? 1 : -1)
#define mc_fabs(x) fabsf((mc_fp_t)(x))
#define mc_fmin(x, y) fminf((mc_fp_t)(x), (mc_fp_t)(y))
#define mc_fmax(x, y) fmaxf((mc_fp_t)(x), (mc_fp_t)(y))
#define mc_min(x, y) (((mc_int_t)(x) < (mc_int_t)(y)) ? (mc_int_t)(x) : (mc_int_t)(y))
#define mc_max(x, y) (((mc_int_t)(x) > (mc_int_t)(y)) ? (mc_int_t)(x) : (mc_int_t)(y))
#define mc_clip(x, low, high) xo::clc::clampi((mc_int_t)(x), (mc_int_t)(low), (mc_int_t)(high))
#define mc_fclip(x, low, high) xo::clipf((mc_fp_t)(x), (mc_fp_t)(low), (mc_fp_t)(high))
#define mc_fsquare(x) ((mc_fp_t)(x)*(mc_fp_t)(x))
#define mc_int(x) xo::f2i(x)
#define mc_uint(x) xo::f2u(x)
#define mc_round(x) roundf(x)
#define mc_floor(x) floorf(x)
#define mc_isfinite(x) (!isinf(x))
// OpenCL built-ins plugin text may call directly
#define convert_int(x) xo::f2i(x)
#define convert_uint(x) xo::f2u(x)
#define native_divide(a, b) xo::M::div((a), (b))
#define native_sqrt(x) xo::M::sqrt(x)
#define native_log(x) xo::M::log(x)
#define native_exp(x) xo::M::exp(x)

// ---- vector helpers (mcbase.template.h:1440-1753, 1875-2111) -----------------------------
__device__ __forceinline__ mc_fpv3_t *mc_transform_fpv3(const mc_matrix3_fp_t *m, const mc_fpv3_t *v, mc_fpv3_t *r) {
	mc_fpv3_t t;
	t.x = m->a_11*v->x + m->a_12*v->y + m->a_13*v->z;
	t.y = m->a_21*v->x + m->a_22*v->y + m->a_23*v->z;
	t.z = m->a_31*v->x + m->a_32*v->y + m->a_33*v->z;
	*r = t;
	return r;
}
__device__ __forceinline__ mc_fpv2_t *mc_transform_fpv2(const mc_matrix2_fp_t *m, const mc_fpv2_t *v, mc_fpv2_t *r) {
	mc_fpv2_t t;
	t.x = m->a_11*v->x + m->a_12*v->y;
	t.y = m->a_21*v->x + m->a_22*v->y;
	*r = t;
	return r;
}
__device__ __forceinline__ mc_fp_t mc_dot_fpv2(const mc_fpv2_t *a, const mc_fpv2_t *b) { return a->x*b->x + a->y*b->y; }
__device__ __forceinline__ mc_fp_t mc_dot_fpv3(const mc_fpv3_t *a, const mc_fpv3_t *b) { return xo::dot3(*a, *b); }
__device__ __forceinline__ mc_fp_t mc_length_fpv2(const mc_fpv2_t *a) { return mc_sqrt(mc_dot_fpv2(a, a)); }
__device__ __forceinline__ mc_fp_t mc_length_fpv3(const mc_fpv3_t *a) { return mc_sqrt(mc_dot_fpv3(a, a)); }
__device__ __forceinline__ mc_fp_t mc_distance2_fpv2(const mc_fpv2_t *a, const mc_fpv2_t *b) {
	mc_fp_t dx = a->x - b->x, dy = a->y - b->y;
	return dx*dx + dy*dy;
}
__device__ __forceinline__ mc_fp_t mc_distance2_fpv3(const mc_fpv3_t *a, const mc_fpv3_t *b) {
	mc_fp_t dx = a->x - b->x, dy = a->y - b->y, dz = a->z - b->z;
	return dx*dx + dy*dy + dz*dz;
}
__device__ __forceinline__ mc_fp_t mc_distance_fpv2(const mc_fpv2_t *a, const mc_fpv2_t *b) { return mc_sqrt(mc_distance2_fpv2(a, b)); }
__device__ __forceinline__ mc_fp_t mc_distance_fpv3(const mc_fpv3_t *a, const mc_fpv3_t *b) { return mc_sqrt(mc_distance2_fpv3(a, b)); }
__device__ __forceinline__ mc_fpv3_t *mc_cross_fpv3(const mc_fpv3_t *a, const mc_fpv3_t *b, mc_fpv3_t *r) {
	mc_fpv3_t t = { a->y*b->z - a->z*b->y, a->z*b->x - a->x*b->z, a->x*b->y - a->y*b->x };
	*r = t;
	return r;
}
__device__ __forceinline__ mc_fpv2_t *mc_reverse_fpv2(const mc_fpv2_t *a, mc_fpv2_t *r) { r->x = -a->x; r->y = -a->y; return r; }
__device__ __forceinline__ mc_fpv3_t *mc_reverse_fpv3(const mc_fpv3_t *a, mc_fpv3_t *r) { r->x = -a->x; r->y = -a->y; r->z = -a->z; return r; }
__device__ __forceinline__ mc_fpv2_t *mc_normalize_fpv2(const mc_fpv2_t *a, mc_fpv2_t *r) {
	mc_fp_t k = mc_rsqrt(a->x*a->x + a->y*a->y);
	r->x = a->x*k; r->y = a->y*k;
	return r;
}
__device__ __forceinline__ mc_fpv3_t *mc_normalize_fpv3(const mc_fpv3_t *a, mc_fpv3_t *r) {
	mc_fp_t k = mc_rsqrt(a->x*a->x + a->y*a->y + a->z*a->z);
	r->x = a->x*k; r->y = a->y*k; r->z = a->z*k;
	return r;
}
__device__ __forceinline__ mc_fpv2_t *mc_mad_fpv2(const mc_fpv2_t *a, const mc_fpv2_t *b, mc_fp_t c, mc_fpv2_t *r) {
	r->x = a->x + b->x*c; r->y = a->y + b->y*c;
	return r;
}
__device__ __forceinline__ mc_fpv3_t *mc_mad_fpv3(const mc_fpv3_t *a, const mc_fpv3_t *b, mc_fp_t c, mc_fpv3_t *r) {
	r->x = a->x + b->x*c; r->y = a->y + b->y*c; r->z = a->z + b->z*c;
	return r;
}
#define transform_point3f(pT, pt, pres) mc_transform_fpv3(pT, pt, pres)
#define transform_point2f(pT, pt, pres) mc_transform_fpv2(pT, pt, pres)
#define mc_transform_point3f(pT, pt, pres) mc_transform_fpv3(pT, pt, pres)
#define mc_transform_point2f(pT, pt, pres) mc_transform_fpv2(pT, pt, pres)
#define transform_point3f_z(pT, pt) ((pT)->a_31*(pt)->x + (pT)->a_32*(pt)->y + (pT)->a_33*(pt)->z)
#define mc_length_point2f(pt) mc_length_fpv2(pt)
#define mc_length_point3f(pt) mc_length_fpv3(pt)
#define mc_dot_point2f(a, b) mc_dot_fpv2(a, b)
#define mc_dot_point3f(a, b) mc_dot_fpv3(a, b)
#define mc_cross_point3f(a, b, r) mc_cross_fpv3(a, b, r)
#define mc_reverse_point2f(pt) mc_reverse_fpv2(pt, pt)
#define mc_reverse_point3f(pt) mc_reverse_fpv3(pt, pt)
#define mc_normalize_point2f(pv) mc_normalize_fpv2(pv, pv)
#define mc_normalize_point3f(pv) mc_normalize_fpv3(pv, pv)
#define mc_mad_point2f(a, b, c, r) mc_mad_fpv2(a, b, c, r)
#define mc_mad_point3f(a, b, c, r) mc_mad_fpv3(a, b, c, r)
#define mc_r2_point2f(pt) mc_dot_fpv2(pt, pt)
#define mc_r_point2f(pt) mc_length_fpv2(pt)
#define mc_r2_point3f(pt) mc_dot_fpv3(pt, pt)
#define mc_r_point3f(pt) mc_length_fpv3(pt)
#define mc_distance2_point2f(a, b) mc_distance2_fpv2(a, b)
#define mc_distance_point2f(a, b) mc_distance_fpv2(a, b)
#define mc_distance2_point3f(a, b) mc_distance2_fpv3(a, b)
#define mc_distance_point3f(a, b) mc_distance_fpv3(a, b)
__device__ __forceinline__ int mc_rectf_contains_ex(mc_fp_t left, mc_fp_t top, mc_fp_t width, mc_fp_t height, mc_fp_t x, mc_fp_t y) {
	return x >= left && x <= left + width && y >= top && y <= top + height;
}
__device__ __forceinline__ int mc_circf_contains_ex(mc_fp_t cx, mc_fp_t cy, mc_fp_t r, mc_fp_t x, mc_fp_t y) {
	mc_fp_t dx = x - cx, dy = y - cy;
	return dx*dx + dy*dy <= r*r;
}
#define mc_rectf_contains_point2f(prect, ppt) \
	mc_rectf_contains_ex((prect)->top_left.x, (prect)->top_left.y, (prect)->width, (prect)->height, (ppt)->x, (ppt)->y)
#define mc_circf_contains_point2f(pcirc, ppt) \
	mc_circf_contains_ex((pcirc)->center.x, (pcirc)->center.y, (pcirc)->r, (ppt)->x, (ppt)->y)

// ---- interface physics (mcbase.template.h:2612-2773) ---------------------------------------
__device__ __forceinline__ mc_fp_t cos_critical(mc_fp_t n1, mc_fp_t n2) { return xo::cos_critical(n1, n2); }
__device__ __forceinline__ mc_fp_t reflectance(mc_fp_t n1, mc_fp_t n2, mc_fp_t cos1, mc_fp_t cc) {
	return xo::reflectance(n1, n2, cos1, cc);
}
__device__ __forceinline__ mc_point3f_t *reflect(const mc_point3f_t *p, const mc_point3f_t *n, mc_point3f_t *r) {
	*r = xo::reflect3(*p, *n);
	return r;
}
__device__ __forceinline__ mc_point3f_t *refract(const mc_point3f_t *p, const mc_point3f_t *n, mc_fp_t n1, mc_fp_t n2, mc_point3f_t *r) {
	*r = xo::refract3(*p, *n, n1, n2);
	return r;
}
__device__ __forceinline__ int refract_safe(const mc_point3f_t *p, const mc_point3f_t *n, mc_fp_t n1, mc_fp_t n2, mc_point3f_t *r) {
	return xo::refract3_safe(*p, *n, n1, n2, r) ? 1 : 0;
}
__device__ __forceinline__ void scatter_direction(mc_point3f_t *dir, mc_fp_t cos_theta, mc_fp_t fi) {
	xo::scatter_direction(*dir, cos_theta, fi);
}

// ---- lookup tables -------------------------------------------------------------------------------
// linear interpolation in a table of the float pool, with the reference's
// semantics (mcbase.template.h:2507-2548, see xo::lut_sample): the first index is
// *rounded*, the weight is the fractional part, the value is written only when
// the position falls inside the table
__device__ __forceinline__ void fp_linear_lut_rel_sample(const mc_fp_t *lut_array, const mc_fp_lut_t *lut, mc_fp_t rel, mc_fp_t *value) {
	xo::lut_sample_index(lut_array, lut->n, lut->offset, rel*(mc_fp_t)(lut->n - 1), false, value);
}
__device__ __forceinline__ void fp_linear_lut_sample(const mc_fp_t *lut_array, const mc_fp_lut_t *lut, mc_fp_t x, mc_fp_t *value) {
	xo::lut_sample_index(lut_array, lut->n, lut->offset,
		(x - lut->first)*lut->inv_span*(mc_fp_t)(lut->n - 1), true, value);
}

// ---- debugging hooks: compiled out ---------------------------------------------------------------
#define INDENT "  "
#define dbg_print(...) ((void)0)
#define dbg_printf(...) ((void)0)
#define dbg_print_status(...) ((void)0)
#define dbg_print_float(...) ((void)0)
#define dbg_print_int(...) ((void)0)
#define dbg_print_uint(...) ((void)0)
#define dbg_print_size_t(...) ((void)0)
#define dbg_print_cnt_t(...) ((void)0)
#define dbg_print_point2f(...) ((void)0)
#define dbg_print_point3f(...) ((void)0)
#define dbg_print_point2(...) ((void)0)
#define dbg_print_point3(...) ((void)0)
#define dbg_print_matrix2f(...) ((void)0)
#define dbg_print_matrix3f(...) ((void)0)
#define dbg_print_fp_lut(...) ((void)0)

// ---- the simulator facade -------------------------------------------------------------------------
// Plugin text only ever sees a `McSim *`.  Slots are untyped here; the accessor
// macros cast them to the struct names the *fragments* declare (McPf, McSource,
// Mc{Top,Bottom,Specular}Detector), so a macro that names a type the user did not
// declare is simply never expanded.
struct McSimState {
	mc_point3f_t position, direction;
	mc_fp_t weight;
	mc_int_t layer_index;
	mc_cnt_t photon_index;
	mc_fp_t optical_pathlength;
	mc_point3_t voxel_index;            // (voxel geometry, xo_clcompat_mcvox.cuh)
	mc_int_t voxel_material_index;
};
struct McSim {
	McSimState state;
	xo::Rng *rng;
	const void *pf, *source, *det_top, *det_bottom, *det_specular, *layers, *fluence;
	const void *det_outer;                  // (cylindrical geometry)
	const void *surf_top, *surf_bottom;
	const void *trace;
	const void *voxel_cfg, *materials;      // (voxel geometry)
	const mc_int_t *voxels;
	mc_fp_t *float_buffer;
	mc_int_t *integer_buffer;
	mc_uint_t event_flags;
	mc_int_t num_layers;
	const mc_fp_t *fp_lut_array;
	mc_accu_t *accumulator_buffer;
	// request recorded by mcsim_specular_detector_deposit inside mcsim_launch; the
	// kernel hands it to the specular detector after the launch
	mc_point3f_t spec_dir;
	mc_fp_t spec_weight;
};

#define mcsim_random_single(psim) ((psim)->rng->next())
#define mcsim_random(psim) mcsim_random_single(psim)
#define mcsim_packet_index(psim) ((psim)->state.photon_index)
#define mcsim_position(psim) (&(psim)->state.position)
#define mcsim_position_x(psim) ((psim)->state.position.x)
#define mcsim_position_y(psim) ((psim)->state.position.y)
#define mcsim_position_z(psim) ((psim)->state.position.z)
#define mcsim_position_r2(psim) \
	((psim)->state.position.x*(psim)->state.position.x + (psim)->state.position.y*(psim)->state.position.y)
#define mcsim_position_r(psim) mc_sqrt(mcsim_position_r2(psim))
#define mcsim_set_position(psim, ppoint) ((psim)->state.position = *(ppoint))
#define mcsim_set_position_z(psim, zpos) ((psim)->state.position.z = (zpos))
#define mcsim_set_position_coordinates(psim, posx, posy, posz) \
	{ (psim)->state.position.x = (posx); (psim)->state.position.y = (posy); (psim)->state.position.z = (posz); }
#define mcsim_direction(psim) (&(psim)->state.direction)
#define mcsim_direction_x(psim) ((psim)->state.direction.x)
#define mcsim_direction_y(psim) ((psim)->state.direction.y)
#define mcsim_direction_z(psim) ((psim)->state.direction.z)
#define mcsim_set_direction(psim, pdir) ((psim)->state.direction = *(pdir))
#define mcsim_set_direction_coordinates(psim, px, py, pz) \
	{ (psim)->state.direction.x = (px); (psim)->state.direction.y = (py); (psim)->state.direction.z = (pz); }
#define mcsim_reverse_direction_z(psim) ((psim)->state.direction.z = -(psim)->state.direction.z)
#define mcsim_weight(psim) ((psim)->state.weight)
#define mcsim_set_weight(psim, w) ((psim)->state.weight = (w))
#define mcsim_adjust_weight(psim, delta) ((psim)->state.weight -= (delta))
#define mcsim_current_layer_index(psim) ((psim)->state.layer_index)
#define mcsim_set_current_layer_index(psim, index) ((psim)->state.layer_index = (index))
#define mcsim_layer_count(psim) ((psim)->num_layers)
#define mcsim_optical_pathlength(psim) ((psim)->state.optical_pathlength)
#define mcsim_fp_lut_array(psim) ((psim)->fp_lut_array)
#define mcsim_pf_lut_array(psim) ((psim)->fp_lut_array)
#define mcsim_fp_lut_array_ex(psim, offset) ((psim)->fp_lut_array + (offset))
#define mcsim_accumulator_buffer(psim) ((psim)->accumulator_buffer)
#define mcsim_accumulator_buffer_ex(psim, offset) ((psim)->accumulator_buffer + (offset))
#define mcsim_current_pf(psim) (static_cast<const McPf *>((psim)->pf))
#define mcsim_current_layer_pf(psim) mcsim_current_pf(psim)
#define mcsim_source(psim) (static_cast<const McSource *>((psim)->source))
#define mcsim_trace(psim) (static_cast<const McTrace *>((psim)->trace))
#define mcsim_float_buffer(psim) ((psim)->float_buffer)
#define mcsim_integer_buffer(psim) ((psim)->integer_buffer)
#define mcsim_event_flags(psim) ((psim)->event_flags)
#define __mc_trace_mem
#define mcsim_top_surface_layout(psim) (static_cast<const McTopSurfaceLayout *>((psim)->surf_top))
#define mcsim_bottom_surface_layout(psim) (static_cast<const McBottomSurfaceLayout *>((psim)->surf_bottom))
// return values of mcsim_{top,bottom}_surface_layout_handler (mcml.template.h:262, 999-1001)
#define MC_SURFACE_LAYOUT_CONTINUE (-1)
#define MC_REFLECTED 1
#define MC_REFRACTED 2
#define mcsim_top_detector(psim) (static_cast<const McTopDetector *>((psim)->det_top))
#define mcsim_outer_detector(psim) (static_cast<const McOuterDetector *>((psim)->det_outer))
#define mcsim_bottom_detector(psim) (static_cast<const McBottomDetector *>((psim)->det_bottom))
#define mcsim_specular_detector(psim) (static_cast<const McSpecularDetector *>((psim)->det_specular))
#define mcsim_fluence(psim) (static_cast<const McFluence *>((psim)->fluence))
// low-level deposit of a fluence fragment (mcml.template.c:229-240): one 64-bit RED to
// the accumulator at `offset`
#define mcsim_fluence_weight_deposit_ll(psim, offset, weight) \
	atomicAdd(reinterpret_cast<unsigned long long *>((psim)->accumulator_buffer + (offset)), \
		(unsigned long long)(uint32_t)(weight))
// `mcsim_specular_detector_deposit(psim, ppos, pdir, w)` is two things in the
// reference: the function a specular-detector fragment defines and the call a
// source fragment makes from mcsim_launch.  The host wraps *source* fragments in
// `#define mcsim_specular_detector_deposit XO_CLC_SPECULAR_REQUEST` ... `#undef`:
// the launch only records the request (the kernels deposit it after the launch,
// which in throughput mode parks the packet in the warp's launch queue first).
#define XO_CLC_SPECULAR_REQUEST(psim, ppos, pdir, w) \
	{ (void)(ppos); (psim)->spec_dir = *(pdir); (psim)->spec_weight = (w); }

// weight -> fixed point (mcml.template.h:606) and the 64-bit deposit
// (mcbase.template.h:729-740): one RED.E.ADD.64 to the global bin
#define weight_to_int(weight) ((weight)*MC_INT_ACCUMULATOR_K + FP_0p5)
#define accumulator_deposit(paddress, weight) \
	atomicAdd(reinterpret_cast<unsigned long long *>(paddress), (unsigned long long)(uint32_t)(weight))
