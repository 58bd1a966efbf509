// mcvox_sources.cuh -- photon packet sources of the voxelised simulator.
//
// Struct members = packed `McSource` of xopto/mcvox/mcsource/{line,gaussianbeam,
// point}.py; launch() consumes the same uniform draws as `mcsim_launch`.
// `Ctx` (VoxCtx) exposes the voxel box, the voxel -> material lookup and the
// material refractive indices.
#pragma once
#include "xo_core.cuh"
#include "mcml_sources.cuh"     // struct Launch

namespace xo {

struct VoxSrcLine {                 // mcvox/mcsource/line.py:57-63
	P3 position, direction_medium, direction_sample, direction_reflected;
	float reflectance;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, const P3 &prev_pos, Launch &L) const {
		(void)rng; (void)ctx; (void)prev_pos;
		L.weight = 1.0f - reflectance;
		L.pos = position;
		L.dir = direction_sample;
		L.spec_dir = direction_reflected;
		L.spec_weight = reflectance;
	}
};

// Specular reflectance + refracted direction of a ray entering the box at a
// point whose voxel is looked up from `lookup_pos` (see the IsotropicPoint note).
template <class Ctx>
__device__ __forceinline__ float vox_enter_n(const Ctx &ctx, const P3 &lookup_pos,
		const P3 &normal, const P3 &direction, P3 *refracted, float n_out) {
	i32 vx, vy, vz;
	ctx.position_to_voxel(lookup_pos, &vx, &vy, &vz);
	vx = clipi(vx, 0, ctx.cfg.nx - 1);
	vy = clipi(vy, 0, ctx.cfg.ny - 1);
	vz = clipi(vz, 0, ctx.cfg.nz - 1);
	float n_in = ctx.material_n(ctx.voxel_material(vx, vy, vz));
	*refracted = direction;
	refract3_safe(direction, normal, n_out, n_in, refracted);
	return reflectance(n_out, n_in, dot3(normal, direction), cos_critical(n_out, n_in));
}

template <class Ctx>
__device__ __forceinline__ float vox_enter(const Ctx &ctx, const P3 &lookup_pos,
		const P3 &normal, const P3 &direction, P3 *refracted) {
	return vox_enter_n(ctx, lookup_pos, normal, direction, refracted, ctx.material_n(0));
}

// Collimated beam whose launch point `pm` (source plane) is projected onto the
// box surface along the beam direction: shared tail of GaussianBeam / UniformBeam
// (mcvox/mcsource/gaussianbeam.py:140-188, uniformbeam.py:95-148)
template <class Ctx>
__device__ __forceinline__ void vox_beam_enter(const Ctx &ctx, const P3 &pm, const P3 &direction, Launch &L) {
	float rs = 0.0f;
	P3 d = { direction.x, direction.y, direction.z };
	if (ctx.box_contains(pm)) { d.x = -d.x; d.y = -d.y; d.z = -d.z; }
	P3 isect, normal;
	if (ctx.box_intersect(pm, d, &isect, &normal)) {
		L.pos = isect;
		d.x = -d.x; d.y = -d.y; d.z = -d.z;
		normal.x = -normal.x; normal.y = -normal.y; normal.z = -normal.z;
		P3 refracted;
		rs = vox_enter(ctx, isect, normal, d, &refracted);
		L.dir = refracted;
		L.spec_dir = d;
		L.spec_weight = rs;
	} else {
		L.pos = ctx.cfg.top_left;
		L.dir.x = 0.0f; L.dir.y = 0.0f; L.dir.z = 1.0f;
		rs = 1.0f;
		L.spec_weight = -1.0f;      // the reference deposits nothing on a miss
	}
	L.weight = 1.0f - rs;
}

struct VoxSrcUniformBeam {          // mcvox/mcsource/uniformbeam.py:36-44
	M3 T; P3 position, direction; P2 radius;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, const P3 &prev_pos, Launch &L) const {
		(void)prev_pos;
		float sf, cf;
		float rs = M::sqrt(rng.next());
		M::sincos(XO_FP_2PI*rng.next(), &sf, &cf);
		P3 ps = { rs*cf*radius.x, rs*sf*radius.y, 0.0f };
		P3 pm = transform3(T, ps);
		pm.x += position.x; pm.y += position.y; pm.z += position.z;
		vox_beam_enter(ctx, pm, direction, L);
	}
};

// mcvox/mcsource/fiber.py: the three fiber sources share everything but the sampling
// of the emission angle.  `st` is the sine of the polar angle in air; the packet
// leaves the core at st / n, enters the box where the fiber axis meets it and is
// refracted into the material of the entry voxel.
template <class Ctx>
__device__ __forceinline__ void vox_fiber_emit(const Ctx &ctx, const M3 &T, const P3 &position,
		const P3 &direction, float n, const P3 &pm, float sf, float cf, float st, Launch &L) {
	float rs = 0.0f;
	st = M::div(st, n);             // emission angle inside the fibre core
	const float ct = M::sqrt(1.0f - st*st);
	P3 pd = { cf*st, sf*st, ct };
	P3 packet_dir = transform3(T, pd);
	P3 sd = { direction.x, direction.y, direction.z };
	if (ctx.box_contains(pm)) { sd.x = -sd.x; sd.y = -sd.y; sd.z = -sd.z; }
	P3 refracted = packet_dir, p = position, isect, normal;
	if (ctx.box_intersect(pm, sd, &isect, &normal)) {
		normal.x = -normal.x; normal.y = -normal.y; normal.z = -normal.z;
		p = isect;
		rs = vox_enter_n(ctx, isect, normal, packet_dir, &refracted, n);
	} else {
		p = ctx.cfg.top_left;
		rs = 1.0f;
	}
	L.pos = p;
	L.dir = refracted;
	L.spec_dir = packet_dir;
	L.spec_weight = rs;
	L.weight = 1.0f - rs;
}

// core position (2 draws) and azimuth of the emission direction (1 draw)
__device__ __forceinline__ P3 vox_fiber_core_point(Rng &rng, const M3 &T, const P3 &position,
		float radius, float *sf, float *cf) {
	float r = M::sqrt(rng.next())*radius;
	M::sincos(rng.next()*XO_FP_2PI, sf, cf);
	P3 ps = { r*(*cf), r*(*sf), 0.0f };
	P3 pm = transform3(T, ps);
	pm.x += position.x; pm.y += position.y; pm.z += position.z;
	M::sincos(rng.next()*XO_FP_2PI, sf, cf);
	return pm;
}

struct VoxSrcUniformFiber {         // mcvox/mcsource/fiber.py:178-521 (UniformFiber)
	M3 T; P3 position, direction; float radius, cos_min, n;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, const P3 &prev_pos, Launch &L) const {
		(void)prev_pos;
		float sf, cf;
		P3 pm = vox_fiber_core_point(rng, T, position, radius, &sf, &cf);
		float ct = 1.0f - rng.next()*(1.0f - cos_min);
		vox_fiber_emit(ctx, T, position, direction, n, pm, sf, cf, M::sqrt(1.0f - ct*ct), L);
	}
};

struct VoxSrcLambertianFiber {      // mcvox/mcsource/fiber.py:523-745 (sin = sqrt(u) NA)
	M3 T; P3 position, direction; float radius, na, n;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, const P3 &prev_pos, Launch &L) const {
		(void)prev_pos;
		float sf, cf;
		P3 pm = vox_fiber_core_point(rng, T, position, radius, &sf, &cf);
		vox_fiber_emit(ctx, T, position, direction, n, pm, sf, cf, M::sqrt(rng.next())*na, L);
	}
};

struct VoxSrcUniformFiberLut {      // mcvox/mcsource/fiber.py:747- (tabulated emission cosine)
	M3 T; P3 position, direction; float radius, n; FpLut lut;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, const P3 &prev_pos, Launch &L) const {
		(void)prev_pos;
		float sf, cf;
		P3 pm = vox_fiber_core_point(rng, T, position, radius, &sf, &cf);
		float ct = 0.0f;
		lut_sample_index(ctx.lut, lut.n, lut.offset, rng.next()*(float)(lut.n - 1), false, &ct);
		vox_fiber_emit(ctx, T, position, direction, n, pm, sf, cf, M::sqrt(1.0f - ct*ct), L);
	}
};

struct VoxSrcGaussianBeam {         // mcvox/mcsource/gaussianbeam.py:71-77
	M3 T; P3 position, direction; P2 sigma; float clip;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, const P3 &prev_pos, Launch &L) const {
		(void)prev_pos;
		float sf, cf;
		float r = M::sqrt(-2.0f*M::log(1.0f - rng.next()));
		r = fminf(r, clip);
		M::sincos(XO_FP_2PI*rng.next(), &sf, &cf);
		P3 ps = { r*cf*sigma.x, r*sf*sigma.y, 0.0f };
		P3 pm = transform3(T, ps);
		pm.x += position.x; pm.y += position.y; pm.z += position.z;
		vox_beam_enter(ctx, pm, direction, L);
	}
};

struct VoxSrcIsotropicPoint {       // mcvox/mcsource/point.py:44-46
	P3 position;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, const P3 &prev_pos, Launch &L) const {
		float sf, cf, rs = 0.0f;
		P3 p = { position.x, position.y, position.z };
		M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
		float ct = 1.0f - 2.0f*rng.next();
		float st = M::sqrt(1.0f - ct*ct);
		P3 d = { cf*st, sf*st, ct };
		if (!ctx.box_contains(p)) {
			P3 isect, normal;
			if (ctx.box_intersect(p, d, &isect, &normal)) {
				p = isect;
				// Reference quirks kept for parity (point.py:105-117): the voxel
				// under the entry point is looked up from the simulator's
				// *previous* position, and the refracted direction is computed
				// into a shadowed variable, i.e. the packet keeps `d`.
				P3 unused;
				rs = vox_enter(ctx, prev_pos, normal, d, &unused);
			} else {
				p = ctx.cfg.top_left;
				rs = 1.0f;
			}
		}
		L.pos = p;
		L.dir = d;
		L.spec_dir = d;
		L.spec_weight = rs;
		L.weight = 1.0f - rs;
	}
};

// mcvox/mcsource/voxel.py:31-110: a uniformly random point of one voxel, isotropic
// direction, no surface passage (weight 1, nothing for the specular detector)
struct VoxSrcIsotropicVoxel {
	P3 position; i32 vx, vy, vz;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, const P3 &prev_pos, Launch &L) const {
		(void)prev_pos;
		const float lim = 1.0f - XO_FP_EPS;      // FP_1 - FP_EPS
		L.pos.x = ((float)vx + fminf(rng.next(), lim))*ctx.cfg.size.x + ctx.cfg.top_left.x;
		L.pos.y = ((float)vy + fminf(rng.next(), lim))*ctx.cfg.size.y + ctx.cfg.top_left.y;
		L.pos.z = ((float)vz + fminf(rng.next(), lim))*ctx.cfg.size.z + ctx.cfg.top_left.z;
		float sf, cf;
		M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
		float ct = 1.0f - 2.0f*rng.next();
		float st = M::sqrt(1.0f - ct*ct);
		L.dir.x = cf*st; L.dir.y = sf*st; L.dir.z = ct;
		L.weight = 1.0f;
		L.spec_dir = L.dir;
		L.spec_weight = 0.0f;
	}
};

// mcvox/mcsource/voxel.py:195-300: one of `n` source voxels is drawn uniformly, its
// weight and (float) indices come from the float pool (4 floats per voxel)
struct VoxSrcIsotropicVoxels {
	P3 position; u32 n, offset;
	__device__ __forceinline__ P3 origin() const { return position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, const P3 &prev_pos, Launch &L) const {
		(void)prev_pos;
		i32 pick = f2i(rng.next()*(float)n);
		if (pick > (i32)(n - 1u)) pick = (i32)(n - 1u);
		const float *entry = ctx.lut + offset + (u32)pick*4u;
		const float w = entry[0], vx = entry[1], vy = entry[2], vz = entry[3];
		const float lim = 1.0f - XO_FP_EPS;      // FP_1 - FP_EPS
		L.pos.x = (vx + fminf(rng.next(), lim))*ctx.cfg.size.x + ctx.cfg.top_left.x;
		L.pos.y = (vy + fminf(rng.next(), lim))*ctx.cfg.size.y + ctx.cfg.top_left.y;
		L.pos.z = (vz + fminf(rng.next(), lim))*ctx.cfg.size.z + ctx.cfg.top_left.z;
		float sf, cf;
		M::sincos(rng.next()*XO_FP_2PI, &sf, &cf);
		float ct = 1.0f - 2.0f*rng.next();
		float st = M::sqrt(1.0f - ct*ct);
		L.dir.x = cf*st; L.dir.y = sf*st; L.dir.z = ct;
		L.weight = w;
		L.spec_dir = L.dir;
		L.spec_weight = 0.0f;
	}
};

}  // namespace xo
