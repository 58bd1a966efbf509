// xo_pf.cuh -- built-in scattering phase functions (Hg, MHg, Gk, Lut).
//
// Each struct is the packed `McPf` of the corresponding reference plugin
// (xopto/mcbase/mcpf/{hg,mhg,gk,lut}.py `cl_type`) plus a `sample()` that
// consumes the same uniform draws in the same order as the plugin's
// `mcsim_pf_sample_angles`, so the MWC streams stay aligned with the reference.
// `lut` is the shared-memory (or global) copy of the float lookup-table pool.
//
// Throughput mode additionally uses a *prepared* form (`Fast`): constants that
// depend only on the layer / material (1-g, 2g, (1+g^2)/2g ...) are computed
// once per CTA when the medium table is staged in shared memory and are kept in
// registers while a packet stays in the layer, so the per-event work is the
// random draw plus a handful of FMAs and one MUFU.  Same distribution, same
// draws; only the association of the floating-point operations differs.
#pragma once
#include "xo_core.cuh"

// 0: the host found no scattering layer / material with g == 0, the isotropic
// special case of Hg / MHg (one extra draw) is compiled out of the throughput loop
#ifndef XO_PF_G0
#define XO_PF_G0 1
#endif

namespace xo {

// Henyey-Greenstein polar cosine from the prepared constants:
//   k = (1-g^2)/(1 + g(2r-1)),  ct = (1+g^2-k^2)/(2g)
//   with q = 1/((1-g) + 2g r):  ct = A - B q^2,  A=(1+g^2)/(2g),  B=(1-g^2)^2/(2g)
// `f` is the raw draw RN(float(u32)) in [0, 2^32]; d1 carries the 2^-32.
struct HgFast {
	float d0, d1, A, B;             // d1 == 0  <=>  g == 0 (isotropic)
	__device__ __forceinline__ void prepare(float g) {
		d0 = 1.0f - g;
		d1 = 2.0f*g*2.3283064365386963e-10f;
		if (g != 0.0f) {
			float inv2g = 1.0f/(2.0f*g);
			A = (1.0f + g*g)*inv2g;
			B = (1.0f - g*g)*(1.0f - g*g)*inv2g;
		} else {
			A = 0.0f; B = 0.0f;
		}
	}
	// f: the raw draw; `wanted` is false when the caller discards the result
	// (then the extra isotropic draw of g == 0 must not be consumed either)
	__device__ __forceinline__ float polar(float f, Rng &rng, bool wanted = true) const {
		float q = FastMath::rcp_approx(fmaf(f, d1, d0));
		float ct = fmaf(-(B*q), q, A);
		if (XO_PF_G0 && d1 == 0.0f && wanted) ct = fmaf(rng.next_raw(), -2.0f*2.3283064365386963e-10f, 1.0f);
		return ct;
	}
};

// prepared form of the phase functions that have nothing worth precomputing
template <class Pf>
struct PfPlainFast {
	Pf pf;
	__device__ __forceinline__ float sample(Rng &rng, const float *lut, float *azimuth) const {
		return pf.sample(rng, lut, azimuth);
	}
};

struct PfHg {                       // mcpf/hg.py:49-50
	float g;
	static constexpr bool uses_lut = false;
	__device__ __forceinline__ float sample(Rng &rng, const float *lut, float *azimuth) const {
		(void)lut;
		*azimuth = XO_FP_2PI*rng.next();
		float r = rng.next();
		float k = M::div(1.0f - g*g, 1.0f + g*(2.0f*r - 1.0f));
		float ct = M::div(1.0f + g*g - k*k, 2.0f*g);
		if (g == 0.0f) ct = 1.0f - 2.0f*rng.next();
		return fmaxf(fminf(ct, 1.0f), -1.0f);
	}
	struct Fast {
		HgFast hg;
		__device__ __forceinline__ float sample(Rng &rng, const float *lut, float *azimuth) const {
			(void)lut;
			*azimuth = rng.next_raw()*(XO_FP_2PI*2.3283064365386963e-10f);
			return fmaxf(fminf(hg.polar(rng.next_raw(), rng), 1.0f), -1.0f);
		}
	};
	__device__ __forceinline__ void prepare(Fast &f) const { f.hg.prepare(g); }
};

struct PfMHg {                      // mcpf/mhg.py:52-58
	float g, beta;
	static constexpr bool uses_lut = false;
	__device__ __forceinline__ float sample(Rng &rng, const float *lut, float *azimuth) const {
		(void)lut;
		float ct;
		*azimuth = XO_FP_2PI*rng.next();
		if (rng.next() <= beta) {
			float r = rng.next();
			float k = M::div(1.0f - g*g, 1.0f + g*(2.0f*r - 1.0f));
			ct = M::div(1.0f + g*g - k*k, 2.0f*g);
			if (g == 0.0f) ct = 1.0f - 2.0f*rng.next();
		} else {
			ct = M::cbrt(2.0f*rng.next() - 1.0f);
		}
		return clipf(ct, -1.0f, 1.0f);
	}
	struct Fast {
		// One draw serves the branch choice AND the polar angle: given f <= beta the
		// draw f/beta is uniform again (and (f - beta)/(1 - beta) given f > beta), so
		// the two affine maps are folded into the constants of the two samplers -
		// the same distribution as the reference's three draws with two (one MWC
		// step, one int->float conversion and the divergent third draw less).
		// Both branches are evaluated and selected: nearly every warp would take
		// both sides of a branch.
		HgFast hg;                  // d1 carries 1/beta
		float beta_raw;             // beta * 2^32 (compared against the raw draw)
		float rl_k, rl_b;           // Rayleigh-like part: ct = cbrt(f*rl_k + rl_b)
		__device__ __forceinline__ float sample(Rng &rng, const float *lut, float *azimuth) const {
			(void)lut;
			*azimuth = rng.next_raw()*(XO_FP_2PI*2.3283064365386963e-10f);
			const float f = rng.next_raw();
			const bool use_hg = f <= beta_raw;
			float ct_hg = hg.polar(f, rng, use_hg);
			float ct_rl = FastMath::cbrt(fmaf(f, rl_k, rl_b));
			float ct = use_hg ? ct_hg : ct_rl;
			return fmaxf(fminf(ct, 1.0f), -1.0f);
		}
	};
	__device__ __forceinline__ void prepare(Fast &f) const {
		f.hg.prepare(g);
		f.beta_raw = beta*4294967296.0f;
		if (beta > 0.0f) f.hg.d1 = f.hg.d1/beta;
		// 2 (f 2^-32 - beta)/(1 - beta) - 1
		const float w = (beta < 1.0f) ? 1.0f/(1.0f - beta) : 0.0f;
		f.rl_k = 2.0f*2.3283064365386963e-10f*w;
		f.rl_b = -2.0f*beta*w - 1.0f;
	}
};

struct PfGk {                       // mcpf/gk.py:58-66
	float g, a, inv_a, a1, a2;
	static constexpr bool uses_lut = false;
	__device__ __forceinline__ float sample(Rng &rng, const float *lut, float *azimuth) const {
		(void)lut;
		float ct;
		*azimuth = XO_FP_2PI*rng.next();
		if (g == 0.0f) {
			ct = 1.0f - 2.0f*rng.next();
		} else if (a == 0.0f) {
			ct = a1 + M::pow(M::div(1.0f - g, 1.0f + g), 2.0f*rng.next())*a2;
		} else {
			float tmp = a1*rng.next() + a2;
			tmp = 1.0f + g*g - M::pow(tmp, -inv_a);
			ct = M::div(tmp, 2.0f*g);
		}
		return clipf(ct, -1.0f, 1.0f);
	}
	typedef PfPlainFast<PfGk> Fast;
	__device__ __forceinline__ void prepare(Fast &f) const { f.pf = *this; }
};

// polar cosine of the Gegenbauer kernel from its packed constants
// (mcpf/gk.py:99-132; shared by Gk, MGk and Gk2)
__device__ __forceinline__ float gk_polar(float g, float a, float inv_a, float a1, float a2, Rng &rng) {
	float ct;
	if (g == 0.0f) {
		ct = 1.0f - 2.0f*rng.next();
	} else if (a == 0.0f) {
		ct = a1 + M::pow(M::div(1.0f - g, 1.0f + g), 2.0f*rng.next())*a2;
	} else {
		float tmp = a1*rng.next() + a2;
		tmp = 1.0f + g*g - M::pow(tmp, -inv_a);
		ct = M::div(tmp, 2.0f*g);
	}
	return ct;
}

struct PfHg2 {                      // mcpf/hg2.py:58-66
	float g1, g2, b;
	static constexpr bool uses_lut = false;
	__device__ __forceinline__ float sample(Rng &rng, const float *lut, float *azimuth) const {
		(void)lut;
		*azimuth = XO_FP_2PI*rng.next();
		float g = (rng.next() >= b) ? g1 : g2;
		float k = M::div(1.0f - g*g, 1.0f + g*(2.0f*rng.next() - 1.0f));
		float ct = M::div(1.0f + g*g - k*k, 2.0f*g);
		if (g == 0.0f) ct = 1.0f - 2.0f*rng.next();
		return fmaxf(fminf(ct, 1.0f), -1.0f);
	}
	typedef PfPlainFast<PfHg2> Fast;
	__device__ __forceinline__ void prepare(Fast &f) const { f.pf = *this; }
};

struct PfMGk {                      // mcpf/mgk.py:64-72
	float g, a, beta, inv_a, a1, a2;
	static constexpr bool uses_lut = false;
	__device__ __forceinline__ float sample(Rng &rng, const float *lut, float *azimuth) const {
		(void)lut;
		float ct;
		*azimuth = XO_FP_2PI*rng.next();
		if (rng.next() <= beta) ct = gk_polar(g, a, inv_a, a1, a2, rng);
		else ct = M::cbrt(2.0f*rng.next() - 1.0f);
		return clipf(ct, -1.0f, 1.0f);
	}
	typedef PfPlainFast<PfMGk> Fast;
	__device__ __forceinline__ void prepare(Fast &f) const { f.pf = *this; }
};

struct PfGk2 {                      // mcpf/gk2.py:61-69
	PfGk gk_1, gk_2; float b;
	static constexpr bool uses_lut = false;
	__device__ __forceinline__ float sample(Rng &rng, const float *lut, float *azimuth) const {
		(void)lut;
		*azimuth = XO_FP_2PI*rng.next();
		const bool first = rng.next() >= b;
		const float g = first ? gk_1.g : gk_2.g, a = first ? gk_1.a : gk_2.a;
		const float inv_a = first ? gk_1.inv_a : gk_2.inv_a;
		const float a1 = first ? gk_1.a1 : gk_2.a1, a2 = first ? gk_1.a2 : gk_2.a2;
		return clipf(gk_polar(g, a, inv_a, a1, a2, rng), -1.0f, 1.0f);
	}
	typedef PfPlainFast<PfGk2> Fast;
	__device__ __forceinline__ void prepare(Fast &f) const { f.pf = *this; }
};

struct PfPc {                       // mcpf/pc.py:50-52
	float n;
	static constexpr bool uses_lut = false;
	__device__ __forceinline__ float sample(Rng &rng, const float *lut, float *azimuth) const {
		(void)lut;
		*azimuth = XO_FP_2PI*rng.next();
		float r = rng.next();
		float ct = 2.0f*M::pow(r, M::div(1.0f, n + 1.0f)) - 1.0f;
		return clipf(ct, -1.0f, 1.0f);
	}
	typedef PfPlainFast<PfPc> Fast;
	__device__ __forceinline__ void prepare(Fast &f) const { f.pf = *this; }
};

// Rayleigh (mcpf/rayleigh.py:54-58 packed struct; Frisvad's importance sampling by Cardano's
// formula).  The sampling text of the reference does not compile (rayleigh.py:96-100: a
// missing `;` and a stray `);`): restated here with its evident meaning.
struct PfRayleigh {
	float gamma, a, b;
	static constexpr bool uses_lut = false;
	__device__ __forceinline__ float sample(Rng &rng, const float *lut, float *azimuth) const {
		(void)lut;
		float ct;
		*azimuth = XO_FP_2PI*rng.next();
		if (gamma == 1.0f) {
			ct = 2.0f*rng.next() - 1.0f;                // isotropic
		} else {
			const float bb = b*(1.0f - 2.0f*rng.next());
			const float tmp = M::sqrt(bb*bb*0.25f + a*a*a*0.037037037037037035f);
			ct = M::cbrt(-0.5f*bb + tmp) + M::cbrt(-0.5f*bb - tmp);
		}
		return clipf(ct, -1.0f, 1.0f);
	}
	typedef PfPlainFast<PfRayleigh> Fast;
	__device__ __forceinline__ void prepare(Fast &f) const { f.pf = *this; }
};

struct PfMPc {                      // mcpf/mpc.py:54-58
	float n, beta;
	static constexpr bool uses_lut = false;
	__device__ __forceinline__ float sample(Rng &rng, const float *lut, float *azimuth) const {
		(void)lut;
		float ct;
		*azimuth = XO_FP_2PI*rng.next();
		if (rng.next() <= beta) {
			float r = rng.next();
			ct = 2.0f*M::pow(r, M::div(1.0f, n + 1.0f)) - 1.0f;
		} else {
			ct = M::cbrt(2.0f*rng.next() - 1.0f);
		}
		return clipf(ct, -1.0f, 1.0f);
	}
	typedef PfPlainFast<PfMPc> Fast;
	__device__ __forceinline__ void prepare(Fast &f) const { f.pf = *this; }
};

struct PfLut {                      // mcpf/lut.py:78-85
	float a, b, c;
	u32 offset, size;
	static constexpr bool uses_lut = true;
	__device__ __forceinline__ float sample(Rng &rng, const float *lut, float *azimuth) const {
		u32 last = size - 1;
		*azimuth = XO_FP_2PI*rng.next();
		float fidx = (M::div(a, rng.next() - c) - b + 1.0f)*(float)last*0.5f;
		float ffl = floorf(fidx);
		float d = fidx - ffl;
		u32 i0 = (u32)f2i(ffl);
		i32 i1 = (i32)(i0 + 1);
		if (i1 > (i32)last) i1 = (i32)last;
		return lut[offset + i0]*(1.0f - d) + lut[offset + (u32)i1]*d;
	}
	typedef PfPlainFast<PfLut> Fast;
	__device__ __forceinline__ void prepare(Fast &f) const { f.pf = *this; }
};

// ---- direction-sampling phase functions (MC_PF_SAMPLE_DIRECTION) -----------------
// A phase function may produce the new direction itself instead of (cos theta,
// azimuth) (mcml.template.c:277-290 mcsim_scatter).  `samples_direction` marks it;
// every scattering site of the kernels goes through pf_scatter().
template <class Pf, class = void>
struct pf_samples_direction { static constexpr bool value = false; };
template <class Pf>
struct pf_samples_direction<Pf, decltype((void)Pf::samples_direction)> {
	static constexpr bool value = Pf::samples_direction;
};

template <class Pf>
__device__ __forceinline__ void pf_scatter(const Pf &pf, Rng &rng, const float *lut, P3 &dir) {
	if constexpr (pf_samples_direction<Pf>::value) {
		pf.sample_dir(rng, lut, dir);
	} else {
		float fi, ct = pf.sample(rng, lut, &fi);
		scatter_direction(dir, ct, fi);
	}
}

// prepared form of a direction-sampling phase function (nothing to precompute)
template <class Pf>
struct PfDirFast {
	Pf pf;
	static constexpr bool samples_direction = true;
	__device__ __forceinline__ void sample_dir(Rng &rng, const float *lut, P3 &dir) const { pf.sample_dir(rng, lut, dir); }
	__device__ __forceinline__ float sample(Rng &, const float *, float *azimuth) const { *azimuth = 0.0f; return 1.0f; }
};

// mcpf/hgdir.py:60-120: Henyey-Greenstein deflection measured from a preferred
// direction with probability p, from the current direction otherwise (draw
// order: polar [, isotropic extra], the choice, azimuth)
struct PfHgDir {
	P3 direction; float g, p;
	static constexpr bool uses_lut = false;
	static constexpr bool samples_direction = true;
	__device__ __forceinline__ void sample_dir(Rng &rng, const float *lut, P3 &dir) const {
		(void)lut;
		float k = M::div(1.0f - g*g, 1.0f + g*(2.0f*rng.next() - 1.0f));
		float ct = M::div(1.0f + g*g - k*k, 2.0f*g);
		if (g == 0.0f) ct = 1.0f - 2.0f*rng.next();
		ct = fmaxf(fminf(ct, 1.0f), -1.0f);
		P3 out = (rng.next() < p) ? direction : dir;
		ct = (dot3(dir, out) < 0.0f) ? -ct : ct;
		scatter_direction(out, ct, XO_FP_2PI*rng.next());
		dir = out;
	}
	// (sample() is never used; present so that generic code compiles)
	__device__ __forceinline__ float sample(Rng &, const float *, float *azimuth) const { *azimuth = 0.0f; return 1.0f; }
	typedef PfDirFast<PfHgDir> Fast;
	__device__ __forceinline__ void prepare(Fast &f) const { f.pf = *this; }
};

}  // namespace xo
