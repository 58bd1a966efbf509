// xo_pf.cuh -- built-in scattering phase functions (Hg, MHg, Gk, Lut).
//
// Each struct is the packed `McPf` of the corresponding reference plugin
// (xopto/mcbase/mcpf/{hg,mhg,gk,lut}.py `cl_type`) plus a `sample()` that
// consumes the same uniform draws in the same order as the plugin's
// `mcsim_pf_sample_angles`, so the MWC streams stay aligned with the reference.
// `lut` is the shared-memory (or global) copy of the float lookup-table pool.
#pragma once
#include "xo_core.cuh"

namespace xo {

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
};

}  // namespace xo
