// xo_aux_kernels.cuh -- helper kernels: RNG known-answer hook and the
// elementary-function probe used by the parity tests.
//  * RngKernel  : counterpart of the reference's RngKernel
//                 (xopto/mcbase/kernel/mcbase.template.c:1640-1646)
//  * MathProbe  : evaluates one function of the active math binding
//                 (DetMath when XO_DETERMINISTIC=1, FastMath otherwise)
#pragma once
#include "xo_core.cuh"

extern "C" __global__ void RngKernel(xo::u64 x, xo::u32 a, xo::u32 n, float *buffer) {
	if (blockIdx.x*blockDim.x + threadIdx.x == 0) {
		xo::Rng rng; rng.load(x); rng.a = a;
		for (xo::u32 i = 0; i < n; ++i) buffer[i] = rng.next();
	}
}

extern "C" __global__ void MathProbe(int fn, xo::u32 n, const float *in0, const float *in1,
		float *out0, float *out1) {
	using namespace xo;
	u32 i = blockIdx.x*blockDim.x + threadIdx.x;
	if (i >= n) return;
	float a = in0[i], b = in1[i];
	switch (fn) {
		case 0: out0[i] = M::log(a); break;
		case 1: M::sincos(a, &out0[i], &out1[i]); break;
		case 2: out0[i] = M::cbrt(a); break;
		case 3: out0[i] = M::pow(a, b); break;
		case 4: out0[i] = M::exp(a); break;
		case 5: out0[i] = M::atan2(a, b); break;
		case 6: out0[i] = M::sqrt(a); break;
		case 7: out0[i] = M::div(a, b); break;
	}
}

// AccuScale: out[i] = (double)accu[i] * inv_k -- the host's `update_data`
// (mcfluence/fluence.py:368-376: raw += accumulators*(1/k), float64) evaluated on
// the device for large grids; same IEEE operations (u64 -> binary64 round to
// nearest, one binary64 multiply), so the result is bit-identical to NumPy's.
extern "C" __global__ void AccuScale(const xo::u64 *accu, double *out, xo::u64 n, double inv_k) {
	const xo::u64 stride = (xo::u64)gridDim.x*blockDim.x;
	for (xo::u64 i = (xo::u64)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += stride)
		out[i] = __dmul_rn(__ull2double_rn(accu[i]), inv_k);
}

// AccuScaleAdd: grid[i] += (double)accu[i] * inv_k -- the same update for a result that
// already holds data (fluence.py:349-376 with `out=`: raw += accumulators*(1/k)), for a
// float64 grid that stays on the device between runs (Mc.lazy_fluence).  Multiply and add
// are separate round-to-nearest operations (no FMA), as in NumPy.
extern "C" __global__ void AccuScaleAdd(const xo::u64 *accu, double *grid, xo::u64 n, double inv_k) {
	const xo::u64 stride = (xo::u64)gridDim.x*blockDim.x;
	for (xo::u64 i = (xo::u64)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += stride)
		grid[i] = __dadd_rn(grid[i], __dmul_rn(__ull2double_rn(accu[i]), inv_k));
}
