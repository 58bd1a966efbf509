// xo_clcompat_glue.cuh -- bodies of the user-plugin adapters (xo_clcompat_slots.cuh).
//
// Included after the fragments' implementations: each adapter builds a McSim
// facade around the kernel's register state, calls the user's function and copies
// the results back.  Everything inlines; the facade never reaches memory.
#pragma once
#include "xo_clcompat_slots.cuh"

namespace xo {

__device__ __forceinline__ void clc_sim_init(McSim &sim, Rng *rng) {
	sim.rng = rng;
	sim.pf = nullptr; sim.source = nullptr;
	sim.det_top = nullptr; sim.det_bottom = nullptr; sim.det_specular = nullptr;
	sim.det_outer = nullptr;
	sim.layers = nullptr; sim.num_layers = 0; sim.fluence = nullptr;
	sim.surf_top = nullptr; sim.surf_bottom = nullptr;
	sim.trace = nullptr; sim.float_buffer = nullptr; sim.integer_buffer = nullptr;
	sim.event_flags = 0u;
	sim.voxel_cfg = nullptr; sim.materials = nullptr; sim.voxels = nullptr;
	sim.state.voxel_index.x = 0; sim.state.voxel_index.y = 0; sim.state.voxel_index.z = 0;
	sim.state.voxel_material_index = 0;
	sim.fp_lut_array = nullptr;
	sim.accumulator_buffer = nullptr;
	sim.state.position = P3{ 0.0f, 0.0f, 0.0f };
	sim.state.direction = P3{ 0.0f, 0.0f, 1.0f };
	sim.state.weight = 1.0f;
	sim.state.layer_index = 1;
	sim.state.photon_index = 0;
	sim.state.optical_pathlength = 0.0f;
	sim.spec_dir = P3{ 0.0f, 0.0f, -1.0f };
	sim.spec_weight = 0.0f;
}

#if XO_USER_PF
__device__ __forceinline__ float PfUser::sample(Rng &rng, const float *lut, float *azimuth) const {
	McSim sim;
	clc_sim_init(sim, &rng);
	sim.pf = &p;
	sim.fp_lut_array = lut;
	return mcsim_pf_sample_angles(&sim, azimuth);
}
#endif

#if XO_USER_SOURCE
template <class Ctx>
__device__ __forceinline__ void SrcUser::launch(Rng &rng, const Ctx &ctx, Launch &L) const {
	McSim sim;
	clc_sim_init(sim, &rng);
	sim.source = &s;
	sim.layers = ctx.layers;
	sim.num_layers = ctx.num_layers;
	mcsim_launch(&sim);
	L.pos = sim.state.position;
	L.dir = sim.state.direction;
	L.weight = sim.state.weight;
	L.layer = sim.state.layer_index;
	L.spec_dir = sim.spec_dir;
	L.spec_weight = sim.spec_weight;
}
#endif

#if XO_USER_SOURCE
template <class Ctx>
__device__ __forceinline__ void SrcUser::launch(Rng &rng, const Ctx &ctx, const P3 &prev_pos, Launch &L) const {
	McSim sim;
	clc_sim_init(sim, &rng);
	sim.source = &s;
	sim.voxel_cfg = &ctx.cfg;
	sim.materials = ctx.materials;
	sim.voxels = ctx.voxels;
	sim.fp_lut_array = ctx.lut;
	sim.state.position = prev_pos;
	sim.spec_weight = -1.0f;            // (no specular deposit unless the fragment asks for one)
	mcsim_launch(&sim);
	L.pos = sim.state.position;
	L.dir = sim.state.direction;
	L.weight = sim.state.weight;
	L.layer = 0;
	L.spec_dir = sim.spec_dir;
	L.spec_weight = sim.spec_weight;
}
#endif

#if XO_USER_FLUENCE
__device__ __forceinline__ void FluUser::deposit(const Accu &acc, const FluWindow &, const P3 &pos, float w, float mua, float opl) const {
	McSim sim;
	Rng none; none.load(0ull); none.a = 0u;
	clc_sim_init(sim, &none);
	sim.fluence = &f;
	sim.accumulator_buffer = acc.global;
	sim.state.position = pos;
	sim.state.weight = w; sim.state.optical_pathlength = opl;
	mc_point3f_t pos_ = pos;
#if XO_FLUENCE_RATE
	mcsim_fluence_deposit_at(&sim, &pos_, w, mua);
#else
	(void)mua;
	mcsim_fluence_deposit_at(&sim, &pos_, w);
#endif
}
#endif

#if XO_USER_TRACE
__device__ __forceinline__ bool trace_event(const TraceUser &t, float *fbuf, u32 packet,
		u32 count, u32 flags, const P3 &pos, const P3 &dir, float w, float opl) {
	McSim sim;
	Rng none; none.load(0ull); none.a = 0u;
	clc_sim_init(sim, &none);
	sim.trace = &t.t;
	sim.float_buffer = fbuf;
	sim.event_flags = flags;
	sim.state.position = pos; sim.state.direction = dir;
	sim.state.weight = w; sim.state.optical_pathlength = opl;
	sim.state.photon_index = packet;
	return mcsim_trace_event(&sim, count) != 0;
}
__device__ __forceinline__ void trace_complete(const TraceUser &t, i32 *ibuf, u32 packet, u32 count) {
	McSim sim;
	Rng none; none.load(0ull); none.a = 0u;
	clc_sim_init(sim, &none);
	sim.trace = &t.t;
	sim.integer_buffer = ibuf;
	sim.state.photon_index = packet;
	mcsim_trace_complete(&sim, count);
}
#endif

#define XO_CLC_SURFACE_BODY(slot, fn) \
	McSim sim; \
	clc_sim_init(sim, &rng); \
	sim.slot = &s.l; \
	sim.layers = layers; sim.num_layers = num_layers; \
	sim.state.position = pos; sim.state.direction = dir; \
	sim.state.weight = weight; sim.state.layer_index = layer; \
	const int r_ = fn(&sim, n2, cc); \
	if (r_ == MC_SURFACE_LAYOUT_CONTINUE) return SURF_CONTINUE; \
	dir = sim.state.direction; weight = sim.state.weight; layer = sim.state.layer_index; \
	return (r_ == MC_REFRACTED) ? SURF_REFRACTED : SURF_REFLECTED;

#if XO_USER_SURF_TOP
__device__ __forceinline__ int surf_handle(const SurfUserTop &s, Rng &rng, const P3 &pos, P3 &dir,
		float &weight, float *n2, float *cc, const MlLayer *layers, i32 num_layers, i32 &layer) {
	XO_CLC_SURFACE_BODY(surf_top, mcsim_top_surface_layout_handler)
}
#endif
#if XO_USER_SURF_BOTTOM
__device__ __forceinline__ int surf_handle(const SurfUserBottom &s, Rng &rng, const P3 &pos, P3 &dir,
		float &weight, float *n2, float *cc, const MlLayer *layers, i32 num_layers, i32 &layer) {
	XO_CLC_SURFACE_BODY(surf_bottom, mcsim_bottom_surface_layout_handler)
}
#endif

#define XO_CLC_DETECTOR_BODY(slot, fn) \
	McSim sim; \
	Rng none; none.load(0ull); none.a = 0u; \
	clc_sim_init(sim, &none); \
	sim.slot = &d; \
	sim.accumulator_buffer = acc.global; \
	sim.state.position = pos; sim.state.direction = dir; \
	sim.state.weight = w; sim.state.optical_pathlength = opl; \
	mc_point3f_t pos_ = pos, dir_ = dir; \
	fn(&sim, &pos_, &dir_, w);

#if XO_USER_DET_TOP
__device__ __forceinline__ void DetUserTop::deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float opl) const {
	XO_CLC_DETECTOR_BODY(det_top, mcsim_top_detector_deposit)
}
#endif
#if XO_USER_DET_BOTTOM
__device__ __forceinline__ void DetUserBottom::deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float opl) const {
	XO_CLC_DETECTOR_BODY(det_bottom, mcsim_bottom_detector_deposit)
}
#endif
#if XO_USER_DET_OUTER
__device__ __forceinline__ void DetUserOuter::deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float opl) const {
	XO_CLC_DETECTOR_BODY(det_outer, mcsim_outer_detector_deposit)
}
#endif
#if XO_USER_DET_SPECULAR
__device__ __forceinline__ void DetUserSpecular::deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float opl) const {
	XO_CLC_DETECTOR_BODY(det_specular, mcsim_specular_detector_deposit)
}
#endif

}  // namespace xo
