// xo_surface.cuh -- sample surface layouts of the layered kernel.
//
// Struct members = packed `Mc{Top,Bottom}SurfaceLayout` of the reference plugins
// (xopto/mcml/mcsurface/{base,lambertian}.py, mcsurface/probe/sixaroundone.py
// `cl_type`).  `handle()` restates `mcsim_{top,bottom}_surface_layout_handler`:
// it is called when a packet reaches the sample surface, before the Fresnel
// logic of the interface (mcml.template.c:100-126), and either
//   * returns SURF_CONTINUE, possibly after overriding the refractive index n2 /
//     critical cosine cc of the medium behind the surface at the point of
//     incidence (fibre core, cladding, cut-out), or
//   * reflects the packet itself (direction, weight) and returns SURF_REFLECTED.
// A user-written layout (OpenCL-C fragment, xo_clcompat_slots.cuh) may also move the packet
// across the surface itself and return SURF_REFRACTED (it sets the layer index).
#pragma once
#include "xo_core.cuh"

namespace xo {

enum { SURF_CONTINUE = 0, SURF_REFLECTED = 1, SURF_REFRACTED = 2 };

struct MlLayer;

// absent layout: SurfaceLayoutDefault {int64 dummy} (mcsurface/base.py:124-152)
struct SurfNone {
	i64 dummy;
	static constexpr bool active = false;
	__device__ __forceinline__ int handle(Rng &, const P3 &, P3 &, float &, float *, float *) const {
		return SURF_CONTINUE;
	}
};

struct SurfLambertian {             // mcsurface/lambertian.py
	float reflectance, specular;
	static constexpr bool active = true;
	__device__ __forceinline__ int handle(Rng &rng, const P3 &pos, P3 &dir, float &weight,
			float *n2, float *cc) const {
		(void)pos; (void)n2; (void)cc;
		if (rng.next() > specular) {
			float sf, cf;
			float st = M::sqrt(rng.next());
#if XO_DETERMINISTIC
			float ct = M::sqrt(__fsub_rn(1.0f, __fmul_rn(st, st)));
			M::sincos(__fmul_rn(rng.next(), XO_FP_2PI), &sf, &cf);
			float z = __fmul_rn(signf(-dir.z), ct);
			dir.x = __fmul_rn(cf, st); dir.y = __fmul_rn(sf, st); dir.z = z;
#else
			float ct = M::sqrt(1.0f - st*st);
			M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
			float z = signf(-dir.z)*ct;
			dir.x = cf*st; dir.y = sf*st; dir.z = z;
#endif
		} else {
			dir.z = -dir.z;
		}
		weight = weight*reflectance;
		return SURF_REFLECTED;
	}
};

struct SurfSixAroundOne {           // mcsurface/probe/sixaroundone.py
	M3 T; P2 position; float core_spacing;
	float cladding_r_squared, cladding_n, cladding_cc;
	float core_r_squared, core_n, core_cc;
	float cutout_r_squared, cutout_n, cutout_cc;
	float probe_r_squared, probe_reflectivity;
	static constexpr bool active = true;
	__device__ __forceinline__ bool fiber(float r2, float *n2, float *cc) const {
		if (r2 <= cladding_r_squared) {
			if (r2 <= core_r_squared) { *n2 = core_n; *cc = core_cc; }
			else { *n2 = cladding_n; *cc = cladding_cc; }
			return true;
		}
		return false;
	}
	__device__ __forceinline__ int handle(Rng &rng, const P3 &pos, P3 &dir, float &weight,
			float *n2, float *cc) const {
		(void)rng;
		float rx = pos.x - position.x, ry = pos.y - position.y;
		P3 p = { rx, ry, 0.0f };
		P3 q = transform3(T, p);
		if (fiber(q.x*q.x + q.y*q.y, n2, cc)) return SURF_CONTINUE;
		p.x = fabsf(rx) - core_spacing; p.y = ry;
		q = transform3(T, p);
		if (fiber(q.x*q.x + q.y*q.y, n2, cc)) return SURF_CONTINUE;
		p.x = fabsf(rx) - core_spacing*0.5f;
		p.y = fabsf(ry) - core_spacing*XO_FP_COS_30;
		q = transform3(T, p);
		if (fiber(q.x*q.x + q.y*q.y, n2, cc)) return SURF_CONTINUE;
		float r2 = rx*rx + ry*ry;
		if (r2 <= cutout_r_squared) { *n2 = cutout_n; *cc = cutout_cc; return SURF_CONTINUE; }
		if (r2 <= probe_r_squared) {
			dir.z = -dir.z;
			weight = weight*probe_reflectivity;
			return SURF_REFLECTED;
		}
		return SURF_CONTINUE;
	}
};

// mcsurface/probe/lineararray.py:150-230: N fibers in a row, rectangular cut-out,
// reflective probe tip.  The tip test uses the position relative to the *last*
// fiber (the loop variable of the reference survives the loop), kept as is.
template <int N>
struct SurfLinearArray {
	M3 T; float c11, c12, c21, c22; P2 position, first_position, delta_position;
	float core_spacing;
	float cladding_r_squared, cladding_n, cladding_cc;
	float core_r_squared, core_n, core_cc;
	float cutout_width_half, cutout_height_half, cutout_n, cutout_cc;
	float probe_r_squared, probe_reflectivity;
	static constexpr bool active = true;
	__device__ __forceinline__ int handle(Rng &rng, const P3 &pos, P3 &dir, float &weight,
			float *n2, float *cc) const {
		(void)rng;
		float fx = first_position.x, fy = first_position.y;
		P3 p = { 0.0f, 0.0f, 0.0f };
#pragma unroll 1
		for (u32 i = 0; i < (u32)N; ++i) {
			p.x = pos.x - fx; p.y = pos.y - fy; p.z = 0.0f;
			P3 q = transform3(T, p);
			float r2 = q.x*q.x + q.y*q.y;
			if (r2 <= cladding_r_squared) {
				if (r2 <= core_r_squared) { *n2 = core_n; *cc = core_cc; }
				else { *n2 = cladding_n; *cc = cladding_cc; }
				return SURF_CONTINUE;
			}
			fx += delta_position.x;
			fy += delta_position.y;
		}
		float cx = pos.x - position.x, cy = pos.y - position.y;
		float dx = fabsf(c11*cx + c12*cy), dy = fabsf(c21*cx + c22*cy);
		if (dx <= cutout_width_half && dy < cutout_height_half) {
			*n2 = cutout_n; *cc = cutout_cc;
			return SURF_CONTINUE;
		}
		if (p.x*p.x + p.y*p.y <= probe_r_squared) {
			dir.z = -dir.z;
			weight = weight*probe_reflectivity;
			return SURF_REFLECTED;
		}
		return SURF_CONTINUE;
	}
};

// mcsurface/probe/fiberarray.py:130-185: N individually placed / tilted fibers and
// the probe tip around `probe_position`
template <int N>
struct SurfFiberArray {
	M3 T[N]; P2 fiber_position[N];
	float cladding_r_squared[N], cladding_n[N], cladding_cc[N];
	float core_r_squared[N], core_n[N], core_cc[N];
	P2 probe_position; float probe_r_squared, probe_reflectivity;
	static constexpr bool active = true;
	__device__ __forceinline__ int handle(Rng &rng, const P3 &pos, P3 &dir, float &weight,
			float *n2, float *cc) const {
		(void)rng;
#pragma unroll 1
		for (u32 i = 0; i < (u32)N; ++i) {
			P3 p = { pos.x - fiber_position[i].x, pos.y - fiber_position[i].y, 0.0f };
			P3 q = transform3(T[i], p);
			float r2 = q.x*q.x + q.y*q.y;
			if (r2 <= cladding_r_squared[i]) {
				if (r2 <= core_r_squared[i]) { *n2 = core_n[i]; *cc = core_cc[i]; }
				else { *n2 = cladding_n[i]; *cc = cladding_cc[i]; }
				return SURF_CONTINUE;
			}
		}
		float dx = pos.x - probe_position.x, dy = pos.y - probe_position.y;
		if (dx*dx + dy*dy <= probe_r_squared) {
			dir.z = -dir.z;
			weight = weight*probe_reflectivity;
			return SURF_REFLECTED;
		}
		return SURF_CONTINUE;
	}
};

// The kernels call a layout through this dispatcher: the hand-written layouts need the
// packet only, the adapters of user-written fragments (non-template overloads in
// xo_clcompat_slots.cuh) also hand the layer stack to the fragment's McSim facade.
template <class S>
__device__ __forceinline__ int surf_handle(const S &s, Rng &rng, const P3 &pos, P3 &dir,
		float &weight, float *n2, float *cc, const MlLayer *layers, i32 num_layers, i32 &layer) {
	(void)layers; (void)num_layers; (void)layer;
	return s.handle(rng, pos, dir, weight, n2, cc);
}

template <class Top, class Bottom>
struct SurfaceLayouts {             // mcsurface/base.py:258-263
	Top top;
	Bottom bottom;
};

}  // namespace xo
