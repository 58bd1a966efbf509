// mcvox_medium.cuh -- packed medium of the voxel geometry: materials and the voxel box
// (xopto/mcbase/mcmaterial.py:52-62,330-341, xopto/mcvox/mcgeometry/voxel.py:96-121).
// Shared by the kernel and by the accessors user-written fragments see
// (xo_clcompat_mcvox.cuh).  Included after the phase-function slot XoPf is bound.
#pragma once
#include "xo_core.cuh"
#include "xo_pf.cuh"

namespace xo {

#ifndef XO_ANISO
#define XO_ANISO 0
#endif
#if XO_ANISO
// AnisotropicMaterial (mcbase/mcmaterial.py:330-341): coefficient tensors projected on
// the propagation direction (:390-455)
struct VoxMaterial {
	float n;
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
struct VoxMaterial {                // mcbase/mcmaterial.py:52-62
	float n, mus, mua, inv_mut, mua_inv_mut;
	XoPf pf;
	__device__ __forceinline__ float mus_at(const P3 &) const { return mus; }
	__device__ __forceinline__ float mua_at(const P3 &) const { return mua; }
	__device__ __forceinline__ float inv_mut_at(const P3 &) const { return inv_mut; }
	__device__ __forceinline__ float mua_inv_mut_at(const P3 &) const { return mua_inv_mut; }
};
#endif
struct VoxCfg {                     // mcvox/mcgeometry/voxel.py:96-121
	P3 top_left, bottom_right, size;
	i32 nx, ny, nz;
};

}  // namespace xo
