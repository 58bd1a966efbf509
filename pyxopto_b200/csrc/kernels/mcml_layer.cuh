// mcml_layer.cuh -- packed layer record of the layered simulator and the view of
// the layer stack that sources get.  Split from mcml_kernel.cuh so that
// user-written plugin fragments (xo_clcompat*.cuh) can see the layer type before
// the kernel body.  The including translation unit has already bound XoPf and
// XoDetSpecular.
#pragma once
#include "xo_core.cuh"

namespace xo {

#ifndef XO_ANISO
#define XO_ANISO 0
#endif

#if XO_ANISO
// AnisotropicLayer (mcml/mclayer/layer.py:412-424): the coefficients seen by a packet
// are the tensors projected on its propagation direction (:497-551)
struct MlLayer {
	float thickness, top, bottom, n, cc_top, cc_bottom;
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
struct MlLayer {                    // mcml/mclayer/layer.py:57-69
	float thickness, top, bottom, n, cc_top, cc_bottom, mus, mua, inv_mut, mua_inv_mut;
	XoPf pf;
	__device__ __forceinline__ float mus_at(const P3 &) const { return mus; }
	__device__ __forceinline__ float mua_at(const P3 &) const { return mua; }
	__device__ __forceinline__ float inv_mut_at(const P3 &) const { return inv_mut; }
	__device__ __forceinline__ float mua_inv_mut_at(const P3 &) const { return mua_inv_mut; }
};
#endif

struct MlCtx {
	const MlLayer *layers;          // shared memory
	i32 num_layers;
	const float *lut;               // float lookup-table pool (*Lut sources)
	static constexpr bool has_specular = XoDetSpecular::active;
	__device__ __forceinline__ float layer_n(int i) const { return layers[i].n; }
	__device__ __forceinline__ float layer_cc_bottom(int i) const { return layers[i].cc_bottom; }
};

}  // namespace xo
