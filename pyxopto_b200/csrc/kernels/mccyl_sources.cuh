// mccyl_sources.cuh -- photon packet sources of the cylindrical simulator.
//
// Struct members = packed `McSource` of xopto/mccyl/mcsource/{line,gaussianbeam,
// uniformbeam,point}.py; launch() consumes the same uniform draws in the same
// order as `mcsim_launch`.  `Ctx` (CylCtx) exposes the layer table.
#pragma once
#include "xo_core.cuh"
#include "mcml_sources.cuh"     // struct Launch

namespace xo {

// ray / cylinder x^2 + y^2 = r^2 (mccyl.template.c:107-125)
__device__ __forceinline__ bool ray_cylinder(float r, const P3 &pos, const P3 &dir,
		float *d1, float *d2) {
	float a = dir.x*dir.x + dir.y*dir.y;
	float b = 2.0f*(pos.x*dir.x + pos.y*dir.y);
	float D = b*b - 4.0f*a*(pos.x*pos.x + pos.y*pos.y - r*r);
	if (a != 0.0f && D > 0.0f) {
		D = M::sqrt(D);
		*d1 = M::div(-b - D, 2.0f*a);
		*d2 = M::div(-b + D, 2.0f*a);
		return true;
	}
	return false;
}

// unit radial normal at `pos`, outward (sign > 0) or inward (mccyl.template.c:220-245)
__device__ __forceinline__ P3 radial_normal(const P3 &pos, bool outward) {
	float k = M::sqrt(pos.x*pos.x + pos.y*pos.y);
	k = (k > 0.0f) ? M::div(1.0f, k) : 0.0f;
	P3 n;
	if (outward) { n.x = pos.x*k; n.y = pos.y*k; }
	else { n.x = -pos.x*k; n.y = -pos.y*k; }
	n.z = 0.0f;
	return n;
}

// Shared tail of the beam / point sources: propagate to the sample surface,
// split off the specular reflection, refract into layer 1
// (gaussianbeam.py:150-190, uniformbeam.py, point.py).  A ray that misses the
// sample becomes a zero-weight packet in the innermost layer.
template <class Ctx>
__device__ __forceinline__ void cyl_enter_sample(const Ctx &ctx, const P3 &p0, const P3 &sdir, Launch &L) {
	float d1, d2;
	L.dir = sdir;
	L.spec_dir = sdir;
	L.spec_weight = 0.0f;
	if (ray_cylinder(ctx.layer_r_inner(0), p0, sdir, &d1, &d2)) {
		float k = fminf(d1, d2);
		// entry point built in a fresh variable: nvrtc 12.9 drops the update of
		// .x when a by-value copy of a __grid_constant__ member is modified in place
		P3 q = { p0.x + k*sdir.x, p0.y + k*sdir.y, p0.z + k*sdir.z };
		P3 normal = radial_normal(q, false);
		float cos1 = normal.x*sdir.x + normal.y*sdir.y;
		float n0 = ctx.layer_n(0), n1 = ctx.layer_n(1), cc = ctx.layer_cc_inner(0);
		float rs = reflectance(n0, n1, cos1, cc);
		if (cos1 > cc) L.dir = refract3(sdir, normal, n0, n1);
		L.layer = 1;
		L.weight = 1.0f - rs;
		if (Ctx::has_specular && rs > 0.0f) {
			L.spec_dir = reflect3(sdir, normal);
			L.spec_weight = rs;
		}
		L.pos = q;
	} else {
		P3 zero = { 0.0f, 0.0f, 0.0f };
		L.pos = zero;
		L.layer = ctx.num_layers - 1;
		L.weight = 0.0f;
	}
}

struct CylSrcLine {                 // mccyl/mcsource/line.py:45-51
	P3 position, direction_medium, direction_sample, direction_reflected;
	float reflectance;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, Launch &L) const {
		(void)rng; (void)ctx;
		L.weight = 1.0f - reflectance;
		L.pos = position;
		L.dir = direction_sample;
		L.spec_dir = direction_reflected;
		L.spec_weight = reflectance;
		L.layer = 1;
	}
};

struct CylSrcGaussianBeam {         // mccyl/mcsource/gaussianbeam.py:50-58 (pack=1)
	M3 T; P3 position, direction; P2 sigma; float clip;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, Launch &L) const {
		float sf, cf;
		float r = M::sqrt(-2.0f*M::log(1.0f - rng.next()));
		r = fminf(r, clip);
		M::sincos(XO_FP_2PI*rng.next(), &sf, &cf);
		P3 ps = { r*cf*sigma.x, r*sf*sigma.y, 0.0f };
		P3 p = transform3(T, ps);
		p.x += position.x; p.y += position.y; p.z += position.z;
		cyl_enter_sample(ctx, p, direction, L);
	}
};

struct CylSrcUniformBeam {          // mccyl/mcsource/uniformbeam.py:47-53
	M3 T; P3 position, direction; P2 radius;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, Launch &L) const {
		float sf, cf;
		float rs = M::sqrt(rng.next());
		M::sincos(XO_FP_2PI*rng.next(), &sf, &cf);
		P3 ps = { rs*cf*radius.x, rs*sf*radius.y, 0.0f };
		P3 p = transform3(T, ps);
		p.x += position.x; p.y += position.y; p.z += position.z;
		cyl_enter_sample(ctx, p, direction, L);
	}
};

struct CylSrcIsotropicPoint {       // mccyl/mcsource/point.py:45-49
	P3 position; u32 layer_index;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, Launch &L) const {
		float sf, cf;
		M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
		float ct = 1.0f - 2.0f*rng.next();
		float st = M::sqrt(1.0f - ct*ct);
		P3 d = { cf*st, sf*st, ct };
		const P3 p0 = { position.x, position.y, position.z };
		float r2 = p0.x*p0.x + p0.y*p0.y;
		float rsam = ctx.layer_r_inner(0);
		if (r2 >= rsam*rsam) {
			cyl_enter_sample(ctx, p0, d, L);
		} else {
			L.weight = 1.0f;
			L.layer = (i32)layer_index;
			L.pos = p0;
			L.dir = d;
			L.spec_dir = d;
			L.spec_weight = 0.0f;
		}
	}
};

}  // namespace xo
