// mcml_sources.cuh -- photon packet sources of the layered simulator.
//
// Struct members = packed `McSource` of the reference plugins
// (xopto/mcml/mcsource/{line,gaussianbeam,fiber,point}.py `cl_type`);
// `launch()` draws the same uniforms in the same order as `mcsim_launch`.
// `Ctx` gives access to layer refractive indices and the specular detector.
#pragma once
#include "xo_core.cuh"

namespace xo {

// Every launch() fills pos/dir/weight/layer and returns the direction + weight
// to hand to the specular detector (weight < 0: nothing to deposit).
struct Launch {
	P3 pos, dir;
	float weight;
	i32 layer;
	P3 spec_dir;
	float spec_weight;
};

struct SrcLine {                    // mcsource/line.py:57-63
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

struct SrcGaussianBeam {            // mcsource/gaussianbeam.py:75-85 (pack=1)
	M3 T; P3 position, direction; P2 sigma; float clip, reflectance;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, Launch &L) const {
		float sf, cf;
		float r = M::sqrt(-2.0f*M::log(1.0f - rng.next()));
		r = fminf(r, clip);
		M::sincos(XO_FP_2PI*rng.next(), &sf, &cf);
		P3 ps = { r*cf*sigma.x, r*sf*sigma.y, 0.0f };
		P3 pm = transform3(T, ps);
		float k = M::div(0.0f - pm.z, direction.z);
		pm.x += k*direction.x;
		pm.y += k*direction.y;
		L.pos.x = position.x + pm.x;
		L.pos.y = position.y + pm.y;
		L.pos.z = 0.0f;
		L.dir = direction;
		L.layer = 1;
		L.weight = 1.0f - reflectance;
		if (Ctx::has_specular) {
			P3 din = { direction.x, direction.y, -direction.z };
			P3 normal = { 0.0f, 0.0f, -1.0f };
			L.spec_dir = refract3(din, normal, ctx.layer_n(1), ctx.layer_n(0));
		}
		L.spec_weight = reflectance;
	}
};

struct SrcUniformBeam {             // mcsource/uniformbeam.py:36-47
	M3 T; P3 position, direction; P2 radius; float reflectance;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, Launch &L) const {
		float sf, cf;
		float rs = M::sqrt(rng.next());
		M::sincos(XO_FP_2PI*rng.next(), &sf, &cf);
		P3 ps = { rs*cf*radius.x, rs*sf*radius.y, 0.0f };
		P3 pm = transform3(T, ps);
		float k = M::div(0.0f - pm.z, direction.z);
		pm.x += k*direction.x;
		pm.y += k*direction.y;
		L.pos.x = position.x + pm.x;
		L.pos.y = position.y + pm.y;
		L.pos.z = 0.0f;
		L.dir = direction;
		L.layer = 1;
		L.weight = 1.0f - reflectance;
		if (Ctx::has_specular) {
			P3 din = { direction.x, direction.y, -direction.z };
			P3 normal = { 0.0f, 0.0f, -1.0f };
			L.spec_dir = refract3(din, normal, ctx.layer_n(1), ctx.layer_n(0));
		}
		L.spec_weight = reflectance;
	}
};

struct SrcUniformFiber {            // mcsource/fiber.py:224-232
	M3 T; P3 position, direction; float radius, cos_min, n;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, Launch &L) const {
		float sf, cf;
		float r = M::sqrt(rng.next())*radius;
		M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
		P3 ps = { r*cf, r*sf, 0.0f };
		P3 pm = transform3(T, ps);
		float k = M::div(0.0f - pm.z, direction.z);
		pm.x += k*direction.x;
		pm.y += k*direction.y;
		L.pos.x = position.x + pm.x;
		L.pos.y = position.y + pm.y;
		L.pos.z = 0.0f;
		M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
		float ct = 1.0f - rng.next()*(1.0f - cos_min);
		float st = M::sqrt(1.0f - ct*ct);
		st = M::div(st, n);
		ct = M::sqrt(1.0f - st*st);
		P3 ds = { cf*st, sf*st, ct };
		P3 d = transform3(T, ds);
		float n1 = ctx.layer_n(1);
		float cc = cos_critical(n, n1);
		// The reference tests the launch-plane offset z (== 0) against cc, so
		// the direction enters the sample un-refracted.
		L.dir = d;
		float rs = reflectance(n, n1, d.z, cc);
		L.weight = 1.0f - rs;
		L.spec_dir = d;
		L.spec_weight = rs;
		L.layer = 1;
	}
};

struct SrcLambertianFiber {         // mcsource/fiber.py:537-545, launch :567-632
	M3 T; P3 position, direction; float radius, na, n;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, Launch &L) const {
		float sf, cf;
		float r = M::sqrt(rng.next())*radius;
		M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
		P3 ps = { r*cf, r*sf, 0.0f };
		P3 pm = transform3(T, ps);
		float k = M::div(0.0f - pm.z, direction.z);
		pm.x += k*direction.x;
		pm.y += k*direction.y;
		L.pos.x = position.x + pm.x;
		L.pos.y = position.y + pm.y;
		L.pos.z = 0.0f;
		M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
		// lambertian: sin(theta) = sqrt(u) NA, then into the fiber core
		float st = M::sqrt(rng.next())*na;
		st = M::div(st, n);
		float ct = M::sqrt(1.0f - st*st);
		P3 ds = { cf*st, sf*st, ct };
		P3 d = transform3(T, ds);
		float n1 = ctx.layer_n(1);
		float cc = cos_critical(n, n1);
		// as in UniformFiber, the reference's refraction test never fires
		// (launch-plane offset z == 0 against cc): un-refracted direction
		L.dir = d;
		float rs = reflectance(n, n1, d.z, cc);
		L.weight = 1.0f - rs;
		L.spec_dir = d;
		L.spec_weight = rs;
		L.layer = 1;
	}
};

struct SrcUniformFiberLut {         // mcsource/fiber.py:719-727, launch :766-830
	M3 T; P3 position, direction; float radius, n; FpLut lut;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, Launch &L) const {
		float sf, cf;
		float r = M::sqrt(rng.next())*radius;
		M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
		P3 ps = { r*cf, r*sf, 0.0f };
		P3 pm = transform3(T, ps);
		float k = M::div(0.0f - pm.z, direction.z);
		pm.x += k*direction.x;
		pm.y += k*direction.y;
		L.pos.x = position.x + pm.x;
		L.pos.y = position.y + pm.y;
		L.pos.z = 0.0f;
		M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
		// emission cosine from the table (sampled with a uniform number)
		float ct = 0.0f;
		lut_sample_index(ctx.lut, lut.n, lut.offset, rng.next()*(float)(lut.n - 1), false, &ct);
		float st = M::sqrt(1.0f - ct*ct);
		st = M::div(st, n);
		ct = M::sqrt(1.0f - st*st);
		P3 ds = { cf*st, sf*st, ct };
		P3 d = transform3(T, ds);
		float n1 = ctx.layer_n(1);
		float cc = cos_critical(n, n1);
		// this source tests the *direction* against the critical cosine: refracted
		P3 normal = { 0.0f, 0.0f, 1.0f };
		L.dir = (d.z > cc) ? refract3(d, normal, n, n1) : d;
		float rs = reflectance(n, n1, d.z, cc);
		L.weight = 1.0f - rs;
		L.spec_dir = d;
		L.spec_weight = rs;
		L.layer = 1;
	}
};

// mcsource/fiberni.py:180-296 / :422-535 / :581-698: fibers at normal incidence
// (no transformation).  APERTURE 0: emission cosine uniform within the NA
// (`aperture` = cos_min), 1: lambertian, sin = sqrt(u) NA (`aperture` = na).  The
// emission angle is adjusted to the refractive index of the first sample layer
// and the weight loses reflectance_cos2(n_core, n_layer, cos); the specular
// detector sees the direction refracted back into the fiber core.
template <class Ctx>
__device__ __forceinline__ void fiberni_finish(const Ctx &ctx, float n, float sf, float cf,
		float st, Launch &L) {
	const float n1 = ctx.layer_n(1);
	st = M::div(st, n1);
	const float ct = M::sqrt(1.0f - st*st);
	L.dir.x = cf*st; L.dir.y = sf*st; L.dir.z = ct;
	const float r = reflectance_cos2(n, n1, ct);
	L.weight = 1.0f - r;
	P3 dir_in = { L.dir.x, L.dir.y, -L.dir.z };
	P3 normal = { 0.0f, 0.0f, -1.0f };
	L.spec_dir = refract3(dir_in, normal, n1, n);
	L.spec_weight = r;
	L.layer = 1;
}

template <bool LAMBERTIAN>
struct SrcFiberNI {
	P3 position; float radius, aperture, n;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, Launch &L) const {
		float sf, cf, st;
		float r = M::sqrt(rng.next())*radius;
		M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
		L.pos.x = position.x + r*cf;
		L.pos.y = position.y + r*sf;
		L.pos.z = 0.0f;
		M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
		if (LAMBERTIAN) {
			st = M::sqrt(rng.next())*aperture;
		} else {
			float ct = 1.0f - rng.next()*(1.0f - aperture);
			st = M::sqrt(1.0f - ct*ct);
		}
		fiberni_finish(ctx, n, sf, cf, st, L);
	}
};
typedef SrcFiberNI<false> SrcUniformFiberNI;
typedef SrcFiberNI<true> SrcLambertianFiberNI;

struct SrcUniformFiberLutNI {       // mcsource/fiberni.py:611-616, launch :642-696
	P3 position; float radius, n; FpLut lut;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, Launch &L) const {
		float sf, cf;
		float r = M::sqrt(rng.next())*radius;
		M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
		L.pos.x = position.x + r*cf;
		L.pos.y = position.y + r*sf;
		L.pos.z = 0.0f;
		M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
		float ct = 0.0f;
		lut_sample_index(ctx.lut, lut.n, lut.offset, rng.next()*(float)(lut.n - 1), false, &ct);
		fiberni_finish(ctx, n, sf, cf, M::sqrt(1.0f - ct*ct), L);
	}
};

// mcsource/rectangular.py:32-315 / :317-528: rectangular emitter at the surface or
// inside a layer; emission cosine uniform within the NA (Uniform) or sin = sqrt(u)
// NA (Lambertian), adjusted to the refractive index of the layer.  (The
// reference's specular branch names a struct field that does not exist, so these
// sources only build without a specular detector; the host refuses otherwise.)
template <bool LAMBERTIAN>
struct SrcRectangular {
	P3 position; P2 size; float n, cos_critical, aperture; u32 layer_index;
		// aperture: cos_min (uniform) or na (lambertian)
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, Launch &L) const {
		float sf, cf, st, ct;
		L.pos.x = position.x + (rng.next() - 0.5f)*size.x;
		L.pos.y = position.y + (rng.next() - 0.5f)*size.y;
		L.pos.z = position.z;
		M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
		if (LAMBERTIAN) {
			st = M::sqrt(rng.next())*aperture;
		} else {
			ct = 1.0f - rng.next()*(1.0f - aperture);
			st = M::sqrt(1.0f - ct*ct);
		}
		const float n_layer = ctx.layer_n((int)layer_index);
		st = M::div(st, n_layer);
		ct = M::sqrt(1.0f - st*st);
		L.dir.x = cf*st; L.dir.y = sf*st; L.dir.z = ct;
		L.weight = 1.0f - reflectance_cos2(n, n_layer, ct);
		L.spec_dir = L.dir;
		L.spec_weight = 0.0f;
		L.layer = (i32)layer_index;
	}
};
typedef SrcRectangular<false> SrcUniformRectangular;
typedef SrcRectangular<true> SrcLambertianRectangular;

// mcsource/rectangular.py:530-854: rectangular emitter with a tabulated emission
// cosine (EmissionLut in the float pool, sampled with a uniform number).  Like the
// other rectangular sources its specular branch names a missing field.
struct SrcUniformRectangularLut {
	P3 position; P2 size; float n, cos_critical; FpLut lut; u32 layer_index;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, Launch &L) const {
		float sf, cf;
		L.pos.x = position.x + (rng.next() - 0.5f)*size.x;
		L.pos.y = position.y + (rng.next() - 0.5f)*size.y;
		L.pos.z = position.z;
		M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
		float ct = 0.0f;
		lut_sample_index(ctx.lut, lut.n, lut.offset, rng.next()*(float)(lut.n - 1), false, &ct);
		const float n_layer = ctx.layer_n((int)layer_index);
		const float st = M::div(M::sqrt(1.0f - ct*ct), n_layer);
		ct = M::sqrt(1.0f - st*st);
		L.dir.x = cf*st; L.dir.y = sf*st; L.dir.z = ct;
		L.weight = 1.0f - reflectance_cos2(n, n_layer, ct);
		L.spec_dir = L.dir;
		L.spec_weight = 0.0f;
		L.layer = (i32)layer_index;
	}
};

struct SrcIsotropicPoint {          // mcsource/point.py:46-49
	P3 position; u32 layer_index;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, Launch &L) const {
		float sf, cf, rs = 0.0f;
		// component-wise copy: nvrtc 12.9 mis-forwards a struct copy of a
		// __grid_constant__ member that is modified on one path only
		P3 p = { position.x, position.y, position.z };
		M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
		float ct = 1.0f - 2.0f*rng.next();
		float st = M::sqrt(1.0f - ct*ct);
		P3 d = { cf*st, sf*st, ct };
		P3 rd = d;
		if (p.z <= 0.0f) {
			float cc = ctx.layer_cc_bottom(0);
			if (d.z > cc) {
				P3 normal = { 0.0f, 0.0f, 1.0f };
				float ns = ctx.layer_n(1), nm = ctx.layer_n(0);
				rd = refract3(d, normal, ns, nm);
				rs = reflectance(nm, ns, d.z, cc);
				float t = M::div(-p.z, d.z);
				p.x += d.x*t;
				p.y += d.y*t;
			} else {
				rs = 1.0f;
			}
			p.z = 0.0f;
		}
		L.dir = rd;
		L.pos = p;
		L.spec_dir = d;
		L.spec_weight = rs;
		L.weight = 1.0f - rs;
		L.layer = (i32)layer_index;
	}
};

}  // namespace xo
