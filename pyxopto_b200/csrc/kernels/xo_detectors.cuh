// xo_detectors.cuh -- surface / specular detectors.
//
// Struct members = packed `Mc{Top,Bottom,Specular}Detector` of the reference
// plugins (xopto/mcml/mcdetector/*.py `cl_type`); `deposit()` restates the bin
// selection and acceptance test of the plugin's `mcsim_*_detector_deposit` and
// pushes the fixed-point weight through the CTA-private accumulator window.
#pragma once
#include "xo_core.cuh"

namespace xo {

// absent detector: DetectorDefault {int64 dummy} (mcdetector/base.py:145-146)
struct DetNone {
	i64 dummy;
	static constexpr bool active = false;
	static constexpr bool needs_opl = false;
	__device__ __forceinline__ void deposit(const Accu &, const P3 &, const P3 &, float, float) const {}
};

struct DetTotal {                   // mcdetector/total.py
	P3 direction; float cos_min; u32 offset;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float) const {
		(void)pos;
		u32 iw = weight_u32(w, cos_min <= fabsf(dot3(dir, direction)));
		if (iw > 0) acc.add(offset, iw);
	}
};

// mcdetector/total.py:240-330 TotalLut: the weight is scaled by an angular
// sensitivity looked up (linear table in the float pool) at |cos| of the
// incidence angle against the detector direction
struct DetTotalLut {
	FpLut lut; P3 direction; u32 offset;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float) const {
		(void)pos;
		float sensitivity = 0.0f;
		lut_sample(acc.lut, lut, fabsf(dot3(dir, direction)), &sensitivity);
		u32 iw = weight_u32(w*sensitivity, true);
		if (iw > 0) acc.add(offset, iw);
	}
};

struct DetRadial {                  // mcdetector/radial.py
	P3 direction; P2 position; float r_min, inv_dr, cos_min; u32 n, offset; i32 log_scale;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float) const {
		float dx = pos.x - position.x, dy = pos.y - position.y;
		float r = M::sqrt(dx*dx + dy*dy);
		if (log_scale) r = M::log(fmaxf(r, XO_FP_RMIN));
		i32 ri = clipi(f2i((r - r_min)*inv_dr), 0, (i32)(n - 1));
		u32 iw = weight_u32(w, cos_min <= fabsf(dot3(dir, direction)));
		if (iw > 0) acc.add(offset + (u32)ri, iw);
	}
};

struct DetCartesian {               // mcdetector/cartesian.py
	P3 direction; float x_min, inv_dx, y_min, inv_dy, cos_min; u32 n_x, n_y, offset;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float) const {
		i32 ix = clipi(f2i((pos.x - x_min)*inv_dx), 0, (i32)(n_x - 1));
		i32 iy = clipi(f2i((pos.y - y_min)*inv_dy), 0, (i32)(n_y - 1));
		u32 iw = weight_u32(w, cos_min <= fabsf(dot3(dir, direction)));
		if (iw > 0) acc.add((u32)iy*n_x + (u32)ix + offset, iw);
	}
};

struct DetSixAroundOne {            // mcdetector/probe/sixaroundone.py
	M3 T; P2 position; float core_r_squared, core_spacing, cos_min; u32 offset;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float) const {
		u32 fiber = 7;
		float rx = pos.x - position.x, ry = pos.y - position.y;
#if !XO_DETERMINISTIC
		// Early out (throughput mode): T is a rotation, so a point inside the core
		// of any of the seven fibres lies within spacing + r_core/|a33| of the
		// probe centre; (a + b)^2 <= 2 (a^2 + b^2) avoids the square root.  Most
		// packets leave the surface far from the probe.
		if (rx*rx + ry*ry > 2.0f*fmaf(core_r_squared, FastMath::rcp_approx(T.a33*T.a33),
				core_spacing*core_spacing)) return;
#endif
		P3 p = { rx, ry, 0.0f };
		P3 q = transform3(T, p);
		if (q.x*q.x + q.y*q.y <= core_r_squared) fiber = 0;
		p.x = fabsf(rx) - core_spacing; p.y = ry;
		q = transform3(T, p);
		if (q.x*q.x + q.y*q.y <= core_r_squared) fiber = (rx >= 0.0f) ? 1 : 4;
		p.x = fabsf(rx) - core_spacing*0.5f;
		p.y = fabsf(ry) - core_spacing*XO_FP_COS_30;
		q = transform3(T, p);
		if (q.x*q.x + q.y*q.y <= core_r_squared)
			fiber = (rx >= 0.0f) ? ((ry >= 0.0f) ? 2 : 6) : ((ry >= 0.0f) ? 3 : 5);
		if (fiber > 6) return;
		float pz = T.a31*dir.x + T.a32*dir.y + T.a33*dir.z;
		u32 iw = weight_u32(w, cos_min <= fabsf(pz));
		if (iw > 0) acc.add(offset + fiber, iw);
	}
};

// mcdetector/probe/lineararray.py:69-170: N equally spaced fibers; the first fiber
// whose core contains the exit point (in fiber coordinates) takes the packet
template <int N>
struct DetLinearArray {
	M3 T; P2 first_position, delta_position; float core_r_squared, cos_min; u32 offset;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float) const {
		u32 fiber = N;
		float fx = first_position.x, fy = first_position.y;
#pragma unroll 1
		for (u32 i = 0; i < (u32)N; ++i) {
			P3 p = { pos.x - fx, pos.y - fy, 0.0f };
			P3 q = transform3(T, p);
			if (q.x*q.x + q.y*q.y <= core_r_squared) { fiber = i; break; }
			fx += delta_position.x;
			fy += delta_position.y;
		}
		if (fiber >= (u32)N) return;
		float pz = T.a31*dir.x + T.a32*dir.y + T.a33*dir.z;
		u32 iw = weight_u32(w, cos_min <= fabsf(pz));
		if (iw > 0) acc.add(offset + fiber, iw);
	}
};

// mcdetector/probe/fiberarray.py:66-160: N individually placed / tilted fibers
template <int N>
struct DetFiberArray {
	M3 T[N]; P2 core_position[N]; float core_r_squared[N], cos_min[N]; u32 offset;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float) const {
		u32 fiber = N;
#pragma unroll 1
		for (u32 i = 0; i < (u32)N; ++i) {
			P3 p = { pos.x - core_position[i].x, pos.y - core_position[i].y, 0.0f };
			P3 q = transform3(T[i], p);
			if (q.x*q.x + q.y*q.y <= core_r_squared[i]) { fiber = i; break; }
		}
		if (fiber >= (u32)N) return;
		float pz = T[fiber].a31*dir.x + T[fiber].a32*dir.y + T[fiber].a33*dir.z;
		u32 iw = weight_u32(w, cos_min[fiber] <= fabsf(pz));
		if (iw > 0) acc.add(offset + fiber, iw);
	}
};

// path-length resolved fiber arrays (probe/lineararraypl.py, fiberarraypl.py):
// bins indexed [pl][fiber]
__device__ __forceinline__ u32 pl_bin(float opl, float pl_min, float inv_dpl, u32 n_pl, i32 log_scale) {
	float pl = opl;
	if (log_scale) pl = M::log(fmaxf(pl, XO_FP_PLMIN));
	return (u32)clipi(f2i((pl - pl_min)*inv_dpl), 0, (i32)(n_pl - 1));
}

// mcdetector/probe/fiberlutarray.py: N fibers, each with its own tabulated
// collection sensitivity (sampled at |cos| in the fiber's frame)
template <int N>
struct DetFiberLutArray {
	M3 T[N]; P2 core_position[N]; float core_r_squared[N]; FpLut lut[N]; u32 offset;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float) const {
		u32 fiber = N;
#pragma unroll 1
		for (u32 i = 0; i < (u32)N; ++i) {
			P3 p = { pos.x - core_position[i].x, pos.y - core_position[i].y, 0.0f };
			P3 q = transform3(T[i], p);
			if (q.x*q.x + q.y*q.y <= core_r_squared[i]) { fiber = i; break; }
		}
		if (fiber >= (u32)N) return;
		float pz = T[fiber].a31*dir.x + T[fiber].a32*dir.y + T[fiber].a33*dir.z;
		float sensitivity = 0.0f;
		lut_sample(acc.lut, lut[fiber], fabsf(pz), &sensitivity);
		u32 iw = weight_u32(w*sensitivity, true);
		if (iw > 0) acc.add(offset + fiber, iw);
	}
};

struct DetTotalLutPl {              // mcdetector/totalpl.py:315-420
	FpLut lut; P3 direction; float pl_min, inv_dpl; u32 n_pl, offset; i32 pl_log_scale;
	static constexpr bool active = true;
	static constexpr bool needs_opl = true;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float opl) const {
		(void)pos;
		u32 pi = pl_bin(opl, pl_min, inv_dpl, n_pl, pl_log_scale);
		float sensitivity = 0.0f;
		lut_sample(acc.lut, lut, fabsf(dot3(dir, direction)), &sensitivity);
		u32 iw = weight_u32(w*sensitivity, true);
		if (iw > 0) acc.add(offset + pi, iw);
	}
};

template <int N>
struct DetLinearArrayPl {
	M3 T; P2 first_position, delta_position; float core_r_squared, cos_min, pl_min, inv_dpl;
	u32 n_pl, offset; i32 pl_log_scale;
	static constexpr bool active = true;
	static constexpr bool needs_opl = true;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float opl) const {
		u32 fiber = N;
		float fx = first_position.x, fy = first_position.y;
#pragma unroll 1
		for (u32 i = 0; i < (u32)N; ++i) {
			P3 p = { pos.x - fx, pos.y - fy, 0.0f };
			P3 q = transform3(T, p);
			if (q.x*q.x + q.y*q.y <= core_r_squared) { fiber = i; break; }
			fx += delta_position.x;
			fy += delta_position.y;
		}
		if (fiber >= (u32)N) return;
		u32 pi = pl_bin(opl, pl_min, inv_dpl, n_pl, pl_log_scale);
		float pz = T.a31*dir.x + T.a32*dir.y + T.a33*dir.z;
		u32 iw = weight_u32(w, cos_min <= fabsf(pz));
		if (iw > 0) acc.add(offset + pi*(u32)N + fiber, iw);
	}
};

template <int N>
struct DetFiberArrayPl {
	M3 T[N]; P2 core_position[N]; float core_r_squared[N], cos_min[N]; float pl_min, inv_dpl;
	u32 n_pl, offset; i32 pl_log_scale;
	static constexpr bool active = true;
	static constexpr bool needs_opl = true;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float opl) const {
		u32 fiber = N;
#pragma unroll 1
		for (u32 i = 0; i < (u32)N; ++i) {
			P3 p = { pos.x - core_position[i].x, pos.y - core_position[i].y, 0.0f };
			P3 q = transform3(T[i], p);
			if (q.x*q.x + q.y*q.y <= core_r_squared[i]) { fiber = i; break; }
		}
		if (fiber >= (u32)N) return;
		u32 pi = pl_bin(opl, pl_min, inv_dpl, n_pl, pl_log_scale);
		float pz = T[fiber].a31*dir.x + T[fiber].a32*dir.y + T[fiber].a33*dir.z;
		u32 iw = weight_u32(w, cos_min[fiber] <= fabsf(pz));
		if (iw > 0) acc.add(offset + pi*(u32)N + fiber, iw);
	}
};

struct DetRadialPl {                // mcdetector/radialpl.py
	P3 direction; P2 position; float r_min, inv_dr, pl_min, inv_dpl, cos_min;
	u32 n_r, n_pl, offset; i32 r_log_scale, pl_log_scale;
	static constexpr bool active = true;
	static constexpr bool needs_opl = true;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float opl) const {
		float dx = pos.x - position.x, dy = pos.y - position.y;
		float r = M::sqrt(dx*dx + dy*dy);
		if (r_log_scale) r = M::log(fmaxf(r, XO_FP_RMIN));
		i32 ri = clipi(f2i((r - r_min)*inv_dr), 0, (i32)(n_r - 1));
		float pl = opl;
		if (pl_log_scale) pl = M::log(fmaxf(pl, XO_FP_PLMIN));
		i32 pi = clipi(f2i((pl - pl_min)*inv_dpl), 0, (i32)(n_pl - 1));
		u32 iw = weight_u32(w, cos_min <= fabsf(dot3(dir, direction)));
		if (iw > 0) acc.add(offset + (u32)pi*n_r + (u32)ri, iw);
	}
};

struct DetSymmetricX {              // mcdetector/symmetric.py
	P3 direction; float position_x, x_offset, inv_step, cos_min; u32 n_half; i32 log_scale; u32 offset;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float) const {
		float x = fabsf(pos.x - position_x);
		if (log_scale) x = M::log(fmaxf(x, XO_FP_RMIN));
		i32 ix = clipi(f2i((x - x_offset)*inv_step), 0, (i32)(n_half - 1));
		u32 index = (pos.x - position_x >= 0.0f) ? (u32)ix + n_half : n_half - (u32)ix - 1u;
		u32 iw = weight_u32(w, cos_min <= fabsf(dot3(dir, direction)));
		if (iw > 0) acc.add(offset + index, iw);
	}
};

struct DetCartesianPl {             // mcdetector/cartesianpl.py
	P3 direction; float x_min, inv_dx, y_min, inv_dy, pl_min, inv_dpl, cos_min;
	u32 n_x, n_y, n_pl, offset; i32 pl_log_scale;
	static constexpr bool active = true;
	static constexpr bool needs_opl = true;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float opl) const {
		i32 ix = clipi(f2i((pos.x - x_min)*inv_dx), 0, (i32)(n_x - 1));
		i32 iy = clipi(f2i((pos.y - y_min)*inv_dy), 0, (i32)(n_y - 1));
		float pl = opl;
		if (pl_log_scale) pl = M::log(fmaxf(pl, XO_FP_PLMIN));
		i32 pi = clipi(f2i((pl - pl_min)*inv_dpl), 0, (i32)(n_pl - 1));
		u32 iw = weight_u32(w, cos_min <= fabsf(dot3(dir, direction)));
		if (iw > 0) acc.add(offset + ((u32)pi*n_y + (u32)iy)*n_x + (u32)ix, iw);
	}
};

struct DetSixAroundOnePl {          // mcdetector/probe/sixaroundonepl.py
	M3 T; P2 position; float core_r_squared, core_spacing, pl_min, inv_dpl, cos_min;
	u32 n_pl, offset; i32 pl_log_scale;
	static constexpr bool active = true;
	static constexpr bool needs_opl = true;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float opl) const {
		u32 fiber = 7;
		float rx = pos.x - position.x, ry = pos.y - position.y;
		P3 p = { rx, ry, 0.0f };
		P3 q = transform3(T, p);
		if (q.x*q.x + q.y*q.y <= core_r_squared) fiber = 0;
		p.x = fabsf(rx) - core_spacing; p.y = ry;
		q = transform3(T, p);
		if (q.x*q.x + q.y*q.y <= core_r_squared) fiber = (rx >= 0.0f) ? 1 : 4;
		p.x = fabsf(rx) - core_spacing*0.5f;
		p.y = fabsf(ry) - core_spacing*XO_FP_COS_30;
		q = transform3(T, p);
		if (q.x*q.x + q.y*q.y <= core_r_squared)
			fiber = (rx >= 0.0f) ? ((ry >= 0.0f) ? 2 : 6) : ((ry >= 0.0f) ? 3 : 5);
		if (fiber > 6) return;
		float pl = opl;
		if (pl_log_scale) pl = M::log(fmaxf(pl, XO_FP_PLMIN));
		i32 pi = clipi(f2i((pl - pl_min)*inv_dpl), 0, (i32)(n_pl - 1));
		float pz = T.a31*dir.x + T.a32*dir.y + T.a33*dir.z;
		u32 iw = weight_u32(w, cos_min <= fabsf(pz));
		if (iw > 0) acc.add(offset + (u32)pi*7u + fiber, iw);
	}
};

struct DetTotalPl {                 // mcdetector/totalpl.py
	P3 direction; float cos_min, pl_min, inv_dpl; u32 n_pl, offset; i32 pl_log_scale;
	static constexpr bool active = true;
	static constexpr bool needs_opl = true;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float opl) const {
		(void)pos;
		float pl = opl;
		if (pl_log_scale) pl = M::log(fmaxf(pl, XO_FP_PLMIN));
		i32 pi = clipi(f2i((pl - pl_min)*inv_dpl), 0, (i32)(n_pl - 1));
		u32 iw = weight_u32(w, cos_min <= fabsf(dot3(dir, direction)));
		if (iw > 0) acc.add(offset + (u32)pi, iw);
	}
};

// ---- cylindrical geometry (mccyl): acceptance against the radial normal ----
struct DetFiZ {                     // mccyl/mcdetector/fiz.py:44-56 (pack=1)
	float fi_min, inv_dfi, z_min, inv_dz, cos_min; u32 n_fi, n_z, offset;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float) const {
		float fi = M::atan2(pos.y, pos.x);
		i32 ifi = clipi(f2i((fi - fi_min)*inv_dfi), 0, (i32)(n_fi - 1));
		i32 iz = clipi(f2i((pos.z - z_min)*inv_dz), 0, (i32)(n_z - 1));
		float k = M::sqrt(pos.x*pos.x + pos.y*pos.y);
		k = (k > 0.0f) ? M::div(1.0f, k) : 0.0f;
		P3 normal = { pos.x*k, pos.y*k, 0.0f };
		u32 iw = weight_u32(w, cos_min <= fabsf(dot3(dir, normal)));
		if (iw > 0) acc.add((u32)iz*n_fi + (u32)ifi + offset, iw);
	}
};

struct DetTotalCyl {                // mccyl/mcdetector/total.py:44-50
	float cos_min; u32 offset;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float) const {
		float r = M::sqrt(pos.x*pos.x + pos.y*pos.y);
		float k = (r > 0.0f) ? M::div(1.0f, r) : 0.0f;
		P3 normal = { pos.x*k, pos.y*k, 0.0f };
		u32 iw = weight_u32(w, cos_min <= fabsf(dot3(dir, normal)));
		if (iw > 0) acc.add(offset, iw);
	}
};

// packed McDetectors {top, bottom, specular} (mcdetector/base.py:321-327);
// natural alignment as laid out by ctypes.
template <class Top, class Bottom, class Specular>
struct Detectors {
	Top top;
	Bottom bottom;
	Specular specular;
};
// packed McDetectors {outer, specular} of mccyl (mccyl/mcdetector/base.py:302-307)
template <class Outer, class Specular>
struct CylDetectors {
	Outer outer;
	Specular specular;
};

}  // namespace xo
