// mccyl_layer.cuh -- packed layer of the cylindrical geometry
// (xopto/mccyl/mclayer/layer.py:119-130,477-491).  Shared by the kernel and by the accessors
// user-written fragments see (xo_clcompat_mccyl.cuh).  Included after XoPf is bound.
#pragma once
#include "xo_core.cuh"
#include "xo_pf.cuh"

namespace xo {

#ifndef XO_ANISO
#define XO_ANISO 0
#endif
#if XO_ANISO
// AnisotropicLayer (mccyl/mclayer/layer.py:477-491): coefficient tensors projected on
// the propagation direction
struct CylLayer {
	float r_inner, r_outer, n, cc_inner, cc_outer;
	M3 mus_t, mua_t, mut_t;
	XoPf pf;
	__device__ __forceinline__ float mus_at(const P3 &d) const { return tensor_project(mus_t, d); }
	__device__ __forceinline__ float mua_at(const P3 &d) const { return tensor_project(mua_t, d); }
	__device__ __forceinline__ float inv_mut_at(const P3 &d) const {
		const float mut = tensor_project(mut_t, d);
		return (mut != 0.0f) ? M::div(1.0f, mut) : XO_INF;
	}
	__device__ __forceinline__ float mua_inv_mut_at(const P3 &d) const {
		const float mua = tensor_project(mua_t, d), mut = tensor_project(mut_t, d);
		return (mua != 0.0f) ? ((mut != 0.0f) ? M::div(mua, mut) : XO_INF) : 0.0f;
	}
};
#else
struct CylLayer {                   // mccyl/mclayer/layer.py:119-130
	float r_inner, r_outer, n, cc_inner, cc_outer, mus, mua, inv_mut, mua_inv_mut;
	XoPf pf;
	__device__ __forceinline__ float mus_at(const P3 &) const { return mus; }
	__device__ __forceinline__ float mua_at(const P3 &) const { return mua; }
	__device__ __forceinline__ float inv_mut_at(const P3 &) const { return inv_mut; }
	__device__ __forceinline__ float mua_inv_mut_at(const P3 &) const { return mua_inv_mut; }
};
#endif

}  // namespace xo
