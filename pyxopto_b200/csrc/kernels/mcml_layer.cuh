// mcml_layer.cuh -- packed layer record of the layered simulator and the view of
// the layer stack that sources get.  Split from mcml_kernel.cuh so that
// user-written plugin fragments (xo_clcompat*.cuh) can see the layer type before
// the kernel body.  The including translation unit has already bound XoPf and
// XoDetSpecular.
#pragma once
#include "xo_core.cuh"

namespace xo {

struct MlLayer {                    // mcml/mclayer/layer.py:57-69
	float thickness, top, bottom, n, cc_top, cc_bottom, mus, mua, inv_mut, mua_inv_mut;
	XoPf pf;
};

struct MlCtx {
	const MlLayer *layers;          // shared memory
	i32 num_layers;
	const float *lut;               // float lookup-table pool (*Lut sources)
	static constexpr bool has_specular = XoDetSpecular::active;
	__device__ __forceinline__ float layer_n(int i) const { return layers[i].n; }
	__device__ __forceinline__ float layer_cc_bottom(int i) const { return layers[i].cc_bottom; }
};

}  // namespace xo
