// xo_core.cuh -- shared device code of the mcml / mcvox / mccyl kernels.
//
// Hand-written CUDA for sm_100a.  Functionally this is the B200 counterpart of
// xopto/mcbase/kernel/mcbase.template.{h,c} (RNG, vector helpers, boundary
// physics, scattering rotation, 64-bit fixed-point accumulators); structure and
// data layout are this engine's own:
//   * the per-thread MWC state lives in registers (one IMAD.WIDE per draw),
//   * packed plugin structs arrive BY VALUE as __grid_constant__ kernel
//     parameters (constant bank -> uniform operands, no loads),
//   * per-thread-indexed tables (layers / materials / pf LUT) are staged in
//     shared memory,
//   * small accumulators (detectors) are privatised per CTA in shared memory as
//     lo/hi 32-bit words and flushed with 64-bit REDs at CTA exit; large grids
//     (fluence) use RED.E.ADD.64 straight to L2.
#pragma once
#include "xo_math.cuh"

namespace xo {

typedef unsigned int u32;
typedef int i32;
typedef unsigned long long u64;
typedef long long i64;

#ifndef XO_DETERMINISTIC
#define XO_DETERMINISTIC 0
#endif
#if XO_DETERMINISTIC
typedef DetMath M;
#else
typedef FastMath M;
#endif

#define XO_FP_2PI XO_FP(6.283185307179586)
#define XO_FP_PI XO_FP(3.141592653589793)
#define XO_FP_COS_30 XO_FP(0.8660254037844386)
#define XO_FP_INV_C XO_FP(3.3356409519815204e-09)
#define XO_FP_RMIN XO_FP(1e-12)
#define XO_FP_PLMIN XO_FP(1e-12)
#define XO_ACCU_K 8388607.0f
// machine epsilon of the kernels' floating-point type (FP_EPS, mcbase.template.h:457-470)
#if XO_DOUBLE
#define XO_FP_EPS 2.220446049250313e-16
#else
#define XO_FP_EPS 1.1920929e-07f
#endif

// event flags (values as in mcbase.template.h:664-681 so Trace event masks mean the same)
enum : u32 {
	EV_REFLECTION = 1u, EV_REFRACTION = 2u, EV_BOUNDARY_HIT = 4u, EV_LAUNCH = 8u,
	EV_ABSORPTION = 16u, EV_SCATTERING = 32u, EV_TERMINATED = 64u, EV_ESCAPED = 128u
};
enum { METHOD_AW = 0, METHOD_AR = 1, METHOD_MBL = 2 };
enum { LOC_TOP = 0, LOC_BOTTOM = 1, LOC_SPECULAR = 2 };

// ---- packed vector types (ctypes layout: tightly packed 4-byte members) -----
struct P3 { float x, y, z; };
struct P2 { float x, y; };
struct M3 { float a11, a12, a13, a21, a22, a23, a31, a32, a33; };

__device__ __forceinline__ float dot3(const P3 &a, const P3 &b) {
#if XO_DETERMINISTIC
	return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z));
#else
	return a.x*b.x + a.y*b.y + a.z*b.z;
#endif
}
// p T p' (mcbase.template.h:2227-2230), the reference's association
__device__ __forceinline__ float tensor_project(const M3 &T, const P3 &p) {
	return p.x*(T.a11*p.x + T.a12*p.y + T.a13*p.z) +
		p.y*(T.a21*p.x + T.a22*p.y + T.a23*p.z) +
		p.z*(T.a31*p.x + T.a32*p.y + T.a33*p.z);
}

__device__ __forceinline__ P3 transform3(const M3 &m, const P3 &v) {
	P3 r;
	r.x = m.a11*v.x + m.a12*v.y + m.a13*v.z;
	r.y = m.a21*v.x + m.a22*v.y + m.a23*v.z;
	r.z = m.a31*v.x + m.a32*v.y + m.a33*v.z;
	return r;
}
__device__ __forceinline__ float clipf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
__device__ __forceinline__ i32 clipi(i32 x, i32 lo, i32 hi) { return x < lo ? lo : (x > hi ? hi : x); }
__device__ __forceinline__ float signf(float x) { return (x >= 0.0f) ? 1.0f : -1.0f; }
// C float->int conversion of the reference (convert_int / (uint32_t) casts):
// CUDA's cvt.rzi saturates where C is undefined; identical on every reachable value.
// floor(x) for |x| < 2^22 on the FMA pipe: bits(x + 1.5 * 2^23, rounded down) - bits(1.5 * 2^23)
#define XO_FLOOR_MAGIC 12582912.0f
#define XO_FLOOR_MAGIC_BITS 0x4B400000u
__device__ __forceinline__ i32 f2i(float x) { return __float2int_rz(x); }
__device__ __forceinline__ u32 f2u(float x) { return __float2uint_rz(x); }

#ifndef XO_ENHANCED_RNG
#define XO_ENHANCED_RNG 0
#endif

// ---- RNG: 64-bit multiply-with-carry, state in registers -------------------
// Same recurrence and float mapping as fp_random_single (mcbase.template.c:1576):
//   x <- lo32(x)*a + hi32(x);  u = RN(float(lo32(x))) / RN(float(0xFFFFFFFF)) = RN(float(lo32 x))*2^-32
// (the divisor rounds to 2^32, so the IEEE division is an exact scaling).
struct Rng {
#if XO_DETERMINISTIC
	u64 x;
	u32 a;
	__device__ __forceinline__ void load(u64 v) { x = v; }
	__device__ __forceinline__ u64 state() const { return x; }
	__device__ __forceinline__ u32 low() const { return (u32)x; }
	__device__ __forceinline__ void step() { x = (u64)(u32)x*(u64)a + (x >> 32); }
#else
	// the two state words as independent 32-bit registers, stepped in place: one
	// step is IMAD.WIDE.U32 + IADD3 + IADD3.X (the plain C expression on a u64
	// compiles to twice as many instructions: the compiler materialises the
	// zero-extended carry word as a register pair)
	u32 lo, hi;
	u32 a;
	__device__ __forceinline__ void load(u64 v) { lo = (u32)v; hi = (u32)(v >> 32); }
	__device__ __forceinline__ u64 state() const { return ((u64)hi << 32) | lo; }
	__device__ __forceinline__ u32 low() const { return lo; }
	__device__ __forceinline__ void step() {
		asm("{\n\t"
			".reg .u64 p;\n\t"
			".reg .u32 pl, ph;\n\t"
			"mul.wide.u32 p, %0, %2;\n\t"
			"mov.b64 {pl, ph}, p;\n\t"
			"add.cc.u32 %0, pl, %1;\n\t"
			"addc.u32 %1, ph, 0;\n\t"
			"}" : "+r"(lo), "+r"(hi) : "r"(a));
	}
#endif
#if XO_DOUBLE
	// fp_random_double (mcbase.template.c:1610-1628): the divisor 0xFFFFFFFF is exact in
	// binary64 (a true division); the enhanced one, 2^64 - 1, rounds to 2^64
	__device__ __forceinline__ double next() {
		step();
#if XO_ENHANCED_RNG
		const u32 high = low();
		step();
		return __ull2double_rn(((u64)high << 32) + low())*5.421010862427522e-20;
#else
		return __ddiv_rn(__uint2double_rn(low()), 4294967295.0);
#endif
	}
	__device__ __forceinline__ double next_raw() { return next()*4294967296.0; }
#elif XO_ENHANCED_RNG
	// MC_USE_ENHANCED_RNG (mcbase.template.c:1577-1586): two steps per draw,
	// u = RN(float(u64)) * 2^-64 (the reference's divisor rounds to 2^64)
	__device__ __forceinline__ float next() {
		step();
		const u32 high = low();
		step();
		return __ull2float_rn(((u64)high << 32) + low())*5.421010862427522e-20f;
	}
	__device__ __forceinline__ float next_raw() { return next()*4294967296.0f; }
#else
	__device__ __forceinline__ float next() {
		step();
		return __uint2float_rn(low())*2.3283064365386963e-10f;
	}
	// the same draw before the 2^-32 scaling, RN(float(lo32 x)) in [0, 2^32]:
	// throughput-mode callers fold the scale into their own constants
	__device__ __forceinline__ float next_raw() {
		step();
		return __uint2float_rn(low());
	}
#endif
};
#define XO_RNG_SCALE 2.3283064365386963e-10f

// ---- boundary physics (Fresnel, Snell) --------------------------------------
__device__ __forceinline__ float cos_critical(float n1, float n2) {
	return (n1 > n2) ? M::sqrt(1.0f - M::div(n2*n2, n1*n1)) : 0.0f;
}

// unpolarised Fresnel reflectance; cos1 = incidence cosine, cc = critical cosine
__device__ inline float reflectance(float n1, float n2, float cos1, float cc) {
	float R = 1.0f;
	cos1 = fabsf(cos1);
	if (n1 == n2) return 0.0f;
	if (cos1 > cc) {
		float n12 = M::div(n1, n2);
		float sin1 = M::sqrt(1.0f - cos1*cos1);
		if (cos1 >= 1.0f) sin1 = 0.0f;
		float sin2 = fminf(1.0f, n12*sin1);
		float cos2 = M::sqrt(1.0f - sin2*sin2);
		float nc1 = n12*cos1, nc2 = n12*cos2;
		float Rs = M::div(nc1 - cos2, nc1 + cos2); Rs *= Rs;
		float Rp = M::div(nc2 - cos1, nc2 + cos1); Rp *= Rp;
		R = 0.5f*(Rp + Rs);
		if (cos1 <= 0.0f || sin2 == 1.0f) return 1.0f;
	}
	return R;
}

// unpolarised Fresnel reflectance from the cosine on the *far* side of the
// interface (mcbase.template.c:1303-1348 reflectance_cos2)
__device__ inline float reflectance_cos2(float n1, float n2, float cos2) {
	float R = 1.0f;
	cos2 = fabsf(cos2);
	if (n1 == n2) return 0.0f;
	float sin2 = M::sqrt(1.0f - cos2*cos2);
	if (cos2 >= 1.0f) sin2 = 0.0f;
	float sin1 = M::div(n2, n1)*sin2;
	if (sin1 < 1.0f) {
		float cos1 = M::sqrt(1.0f - sin1*sin1);
		float n12 = M::div(n1, n2);
		float nc1 = n12*cos1, nc2 = n12*cos2;
		float Rs = M::div(nc1 - cos2, nc1 + cos2); Rs *= Rs;
		float Rp = M::div(nc2 - cos1, nc2 + cos1); Rp *= Rp;
		R = 0.5f*(Rp + Rs);
		if (cos1 <= 0.0f || sin2 == 1.0f) return 1.0f;
	}
	return R;
}

__device__ __forceinline__ P3 reflect3(const P3 &p, const P3 &n) {
	float k = 2.0f*dot3(p, n);
	P3 r = { p.x - n.x*k, p.y - n.y*k, p.z - n.z*k };
	return r;
}
__device__ __forceinline__ P3 refract3(const P3 &p, const P3 &n, float n1, float n2) {
	float cos1 = dot3(p, n);
	float n12 = M::div(n1, n2);
	float sin2sq = n12*n12*(1.0f - cos1*cos1);
	float k = signf(cos1)*(n12*fabsf(cos1) - M::sqrt(1.0f - sin2sq));
	P3 r = { n12*p.x - k*n.x, n12*p.y - k*n.y, n12*p.z - k*n.z };
	return r;
}
__device__ __forceinline__ bool refract3_safe(const P3 &p, const P3 &n, float n1, float n2, P3 *r) {
	float cos1 = dot3(p, n);
	float n12 = M::div(n1, n2);
	float sin2sq = n12*n12*(1.0f - cos1*cos1);
	if (sin2sq > 1.0f) return true;
	float k = signf(cos1)*(n12*fabsf(cos1) - M::sqrt(1.0f - sin2sq));
	r->x = n12*p.x - k*n.x; r->y = n12*p.y - k*n.y; r->z = n12*p.z - k*n.z;
	return false;
}

#if !XO_DETERMINISTIC
// Fresnel / Snell at an interface whose normal is a coordinate axis (throughput
// mode): dn = direction component along the normal, da / db the tangential ones,
// n12 = n1/n2, cc = critical cosine.  One square root instead of two; the special
// cases of `reflectance` (cos1 <= 0, sin2 == 1) fall out of the formula
// (cos1 > cc >= 0; cos2 == 0 gives Rs = Rp^2 = 1).  Consumes one draw only above
// the critical angle, like the reference.  Returns true when the packet passes.
__device__ __forceinline__ bool fresnel_axis_fast(float n12, float cc, float &dn, float &da,
		float &db, Rng &rng) {
	float cos1 = fabsf(dn);
	if (cos1 > cc) {
		float s2 = (n12*n12)*fmaf(-cos1, cos1, 1.0f);
		float cos2 = FastMath::sqrt(fmaxf(1.0f - s2, 0.0f));
		float a = n12*cos1, b = n12*cos2;
		float rs = (a - cos2)*FastMath::rcp_approx(a + cos2);
		float rp = (b - cos1)*FastMath::rcp_approx(b + cos1);
		float R = 0.5f*fmaf(rs, rs, rp*rp);
		if (R*4294967296.0f < rng.next_raw()) {
			da *= n12;
			db *= n12;
			dn = copysignf(cos2, dn);
			return true;
		}
	}
	dn = -dn;
	return false;
}
#endif

// ---- scattering rotation -----------------------------------------------------
// Rotates `d` by polar cosine ct and azimuth fi, then renormalises (the
// reference always renormalises in single precision, mcbase.template.c:1551).
__device__ __forceinline__ void scatter_direction(P3 &d, float ct, float fi) {
#if XO_DETERMINISTIC
	float sf, cf;
	float st = M::sqrt(1.0f - ct*ct);
	M::sincos(fi, &sf, &cf);
	float stcf = st*cf, stsf = st*sf;
	float px = d.x;
	if (fabsf(d.z) >= 1.0f) {
		d.x = stcf;
		d.y = stsf;
		d.z = copysignf(ct, d.z*ct);
	} else {
		float k = M::sqrt(1.0f - d.z*d.z);
		d.x = M::div(stcf*px*d.z - stsf*d.y, k) + px*ct;
		d.y = M::div(stcf*d.y*d.z + stsf*px, k) + d.y*ct;
		d.z = (-stcf)*k + d.z*ct;
	}
	float k = M::div(1.0f, M::sqrt(d.x*d.x + d.y*d.y + d.z*d.z));
	d.x *= k; d.y *= k; d.z *= k;
#else
	// branch-free: the general rotation with 1 - dz^2 kept away from zero, the
	// |dz| >= 1 case of the reference selected afterwards (three selects instead
	// of a divergent region)
	float sf, cf;
	float st = M::sqrt(1.0f - ct*ct);
	M::sincos(fi, &sf, &cf);
	const float stcf = st*cf, stsf = st*sf;
	const float px = d.x, py = d.y, pz = d.z;
	const bool polar = fabsf(pz) >= 1.0f;
	const float k2 = fmaxf(fmaf(-pz, pz, 1.0f), 1e-30f);
	const float ik = rsqrtf(k2);
	float nx = (stcf*px*pz - stsf*py)*ik + px*ct;
	float ny = (stcf*py*pz + stsf*px)*ik + py*ct;
	float nz = pz*ct - stcf*(k2*ik);
	nx = polar ? stcf : nx;
	ny = polar ? stsf : ny;
	nz = polar ? copysignf(ct, pz*ct) : nz;
	// renormalisation (the reference always renormalises).  A Newton step 1.5 - 0.5 s
	// in place of the MUFU.RSQ is NOT enough: within ~1e-6 of the poles 1 - pz^2 loses
	// all its digits and the rotated vector is off by O(1) (a handful of events per
	// 1e7 leave with |dir| - 1 > 1e-4; measured with the trace consistency tests)
	const float k = rsqrtf(nx*nx + ny*ny + nz*nz);
	d.x = nx*k; d.y = ny*k; d.z = nz*k;
#endif
}

// ---- linear lookup tables in the float pool --------------------------------------
// descriptor mc_fp_lut_t (mcbase.template.h:2494-2503) and the sampling macros
// fp_linear_lut_(rel_)sample (:2507-2548), restated with their quirks: the first
// index is the *rounded* position, the interpolation weight its fractional part,
// and `value` is left untouched when the position falls outside the table.
struct FpLut { float first, inv_span; u32 n, offset; };
__device__ __forceinline__ void lut_sample_index(const float *pool, u32 n, u32 offset,
		float fp_index, bool range_on_position, float *value) {
	const u32 i1 = f2u(fp_index + 0.5f);
	const bool inside = range_on_position ? (fp_index >= 0.0f && fp_index <= (float)(n - 1))
		: (i1 < n);
	if (inside) {
		const float w2 = fp_index - floorf(fp_index);
		const i32 nxt = (i32)i1 + 1;
		const u32 i2 = (u32)clipi(nxt, 0, (i32)n - 1);
		*value = pool[offset + i1]*(1.0f - w2) + pool[offset + i2]*w2;
	}
}
__device__ __forceinline__ void lut_sample(const float *pool, const FpLut &lut, float x, float *value) {
	lut_sample_index(pool, lut.n, lut.offset, (x - lut.first)*lut.inv_span*(float)(lut.n - 1), true, value);
}

// ---- accumulators ------------------------------------------------------------
// The flat accumulator buffer holds detector bins first (pack order) and the
// fluence grid after them.  Bins [0, priv_len) are privatised per CTA in shared
// memory as (lo, hi) 32-bit pairs -- the reference's accu_64_deposit_32 carry
// trick (mcbase.template.c:58-61) applied to shared memory, where 32-bit ATOMS
// are native.  Integer adds commute, so totals are exact under any schedule.
// CTA-private window of the fluence / deposition grid (kernel parameter, chosen
// by the host around the source): grid cells with index (i0,i1,i2) such that
// (ik - org_k) < ext_k accumulate in shared memory as 32-bit partial sums; a
// wrap-around of the partial sum is forwarded to the global 64-bit bin as one
// RED of 2^32, so the 64-bit totals stay exact.  Measured on B200
// (tools/atomic_probe.cu): shared-memory deposits sustain > 1.2e12 /s even when
// peaked on a few bins, RED.E.ADD.64 to L2 1.8e11 /s uniform and 1.2-2.9e10 /s
// when peaked (per-address serialisation in the L2 slice).
struct FluWindow { u32 org0, org1, org2, ext0, ext1, ext2; };

// 32-bit shared-memory atomic add on a shared-space address (ATOMS.ADD with the
// address in one register; through a generic pointer the compiler rebuilds the
// shared window base from SR_CgaCtaId before every atomic)
__device__ __forceinline__ u32 atoms_add(u32 saddr, u32 v) {
	u32 old;
	asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(saddr), "r"(v) : "memory");
	return old;
}
__device__ __forceinline__ u32 shared_address(const void *p) {
	return (u32)__cvta_generic_to_shared(p);
}

struct Accu {
	u64 *global;
	u32 *priv;        // shared memory, 2*priv_len words
	u32 priv_len;
	u32 *win;         // shared memory, ext0*ext1*ext2 words (0 extents: no window)
	u32 priv_s, win_s;  // shared-space addresses of priv / win (set by bind())
	const float *lut;   // float lookup-table pool (shared-memory copy if staged): *Lut plugins
	// (the empty asm keeps both addresses in registers: the compiler otherwise
	// rematerialises them from SR_CgaCtaId and the kernel parameters before every
	// atomic, 7 instructions per deposit)
	__device__ __forceinline__ void bind() {
		priv_s = shared_address(priv); win_s = shared_address(win);
		asm volatile("" : "+r"(priv_s), "+r"(win_s));
	}
	// returns true when the 32-bit partial sum wrapped around (the caller then
	// forwards 2^32 to the global bin with carry_global)
	__device__ __forceinline__ bool add_window(u32 local, u32 w) const {
		u32 old = atoms_add(win_s + 4u*local, w);
		return old + w < old;
	}
	// the 16 bytes in front of the window (constants of the deposits that miss it)
	__device__ __forceinline__ uint4 load_far() const {
		uint4 q;
		asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+-16];"
			: "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(win_s));
		return q;
	}
	__device__ __forceinline__ void carry_global(u32 index) const {
		atomicAdd(global + index, 1ull << 32);
	}
	__device__ __forceinline__ void add(u32 index, u32 w) const {
		if (index < priv_len) {
			u32 old = atoms_add(priv_s + 8u*index, w);
			if (old + w < old) atoms_add(priv_s + 8u*index + 4u, 1u);
		} else {
			atomicAdd(global + index, (u64)w);
		}
	}
	// large-grid deposit that bypasses the private window test
	__device__ __forceinline__ void add_global(u32 index, u32 w) const {
		atomicAdd(global + index, (u64)w);
	}
	__device__ __forceinline__ void zero_private() const {
		for (u32 i = threadIdx.x; i < 2*priv_len; i += blockDim.x) priv[i] = 0;
	}
	__device__ __forceinline__ void flush_private() const {
		for (u32 i = threadIdx.x; i < priv_len; i += blockDim.x) {
			u64 v = ((u64)priv[2*i + 1] << 32) | priv[2*i];
			if (v) atomicAdd(global + i, v);
		}
	}
};

// weight -> fixed point with the detector acceptance test folded in:
// (uint32)((w*K + 0.5)*(int)accept)   (mcml.template.h:606 + mcdetector/*.py)
__device__ __forceinline__ u32 weight_u32(float w, bool accept) {
#if XO_DETERMINISTIC
	float v = __fadd_rn(__fmul_rn(w, XO_ACCU_K), 0.5f);
#else
	float v = fmaf(w, XO_ACCU_K, 0.5f);
#endif
	return accept ? f2u(v) : 0u;
}

// ---- packet budget ------------------------------------------------------------
// Deterministic mode: work-item t owns the packets of its static quota.
// Throughput mode: the reference's global packet counter (mcml.template.c:460,
// 790), claimed `chunk` packets per atomic; `dry` is set once the counter has
// passed the budget so a finished lane never touches the counter again.
struct Budget {
	u32 next, end;
	bool dry;
	__device__ __forceinline__ bool claim(u32 N, u32 *counter, u32 chunk, u32 *packet) {
#if !XO_DETERMINISTIC
		if (next >= end && !dry) {
			next = atomicAdd(counter, chunk);
			if (next >= N) { dry = true; end = next; }
			else end = (N - next < chunk) ? N : next + chunk;
		}
#else
		(void)N; (void)counter; (void)chunk;
#endif
		if (next < end) { *packet = next++; return true; }
		return false;
	}
};

// block-schedule quota of work-item t (deterministic mode)
__device__ __forceinline__ void static_quota(u32 N, u32 T, u32 t, u32 *first, u32 *end) {
	u32 q = N/T, r = N % T;
	u32 n_t = q + (t < r ? 1u : 0u);
	u32 base = t*q + (t < r ? t : r);
	*first = base;
	*end = base + n_t;
}

}  // namespace xo
