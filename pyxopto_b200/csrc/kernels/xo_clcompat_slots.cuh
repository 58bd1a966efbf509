// xo_clcompat_slots.cuh -- plugin slots filled by user-written OpenCL-C fragments.
//
// Included after the fragments' declarations (`struct McPf {...}` etc.).  Each
// adapter wraps the user's packed struct in the interface the CUDA kernels bind
// their plugin slots to (xo_pf.cuh, mcml_sources.cuh, xo_detectors.cuh); the
// method bodies follow in xo_clcompat_glue.cuh, after the fragments'
// implementations.  XO_USER_* (0/1) are written by the host for the slots whose
// plugin object carries `cl_implementation` text instead of a `cu_type`.
#pragma once
#include "xo_clcompat.cuh"
#include "xo_pf.cuh"
#include "xo_detectors.cuh"

#ifndef XO_USER_PF
#define XO_USER_PF 0
#endif
#ifndef XO_USER_SOURCE
#define XO_USER_SOURCE 0
#endif
#ifndef XO_USER_DET_TOP
#define XO_USER_DET_TOP 0
#endif
#ifndef XO_USER_DET_BOTTOM
#define XO_USER_DET_BOTTOM 0
#endif
#ifndef XO_USER_DET_OUTER
#define XO_USER_DET_OUTER 0
#endif
#ifndef XO_USER_DET_SPECULAR
#define XO_USER_DET_SPECULAR 0
#endif

#ifndef XO_USER_FLUENCE
#define XO_USER_FLUENCE 0
#endif
#ifndef XO_USER_TRACE
#define XO_USER_TRACE 0
#endif
#ifndef XO_USER_SURF_TOP
#define XO_USER_SURF_TOP 0
#endif
#ifndef XO_USER_SURF_BOTTOM
#define XO_USER_SURF_BOTTOM 0
#endif

namespace xo {

struct Launch;

#if XO_USER_FLUENCE
// fluence / deposition accumulator: `inline void mcsim_fluence_deposit_at(McSim *,
// mc_point3f_t const *position, mc_fp_t weight[, mc_fp_t mua])` (the mua argument with
// MC_FLUENCE_MODE_RATE, mcfluence/fluence.py:103-108); the fragment deposits with
// mcsim_fluence_weight_deposit_ll
struct FluUser {
	McFluence f;
	static constexpr bool active = true;
	static constexpr bool needs_opl = true;     // (the facade carries the optical path length)
	static constexpr bool fixed_point = false;  // the fragment converts the weight itself
	struct Prep { };
	struct Far { };
	__device__ __forceinline__ Prep prepare(const FluWindow &) const { return Prep(); }
	__device__ __forceinline__ Far prepare_far(const FluWindow &) const { return Far(); }
	__device__ __forceinline__ float fixed_scale(float) const { return 1.0f; }
	__device__ __forceinline__ void deposit(const Accu &acc, const FluWindow &, const P3 &pos, float w, float mua, float opl) const;
	__device__ __forceinline__ void deposit_prep(const Accu &, const Prep &, const FluWindow &, const P3 &, u32, float) const { }
	__device__ __forceinline__ u32 window_index(const FluWindow &, u32) const { return 0u; }
};
#endif

#if XO_USER_PF
// scattering phase function: `inline mc_fp_t mcsim_pf_sample_angles(McSim *, mc_fp_t *azimuth)`
struct PfUser {
	McPf p;
	static constexpr bool uses_lut = true;      // the pool is staged whenever the plugin appended a table
	__device__ __forceinline__ float sample(Rng &rng, const float *lut, float *azimuth) const;
	typedef PfPlainFast<PfUser> Fast;
	__device__ __forceinline__ void prepare(Fast &f) const { f.pf = *this; }
};
#endif

#if XO_USER_SOURCE
// packet source: `inline void mcsim_launch(McSim *)`; the struct must have a
// `position` member (the origin of the rmax sphere, mcml.template.c:759)
struct SrcUser {
	McSource s;
	__device__ __forceinline__ P3 origin() const { return s.position; }
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, Launch &L) const;
	// voxel geometry: the facade starts at the position of the previous packet of the
	// work-item, like the reference's simulator state
	template <class Ctx>
	__device__ __forceinline__ void launch(Rng &rng, const Ctx &ctx, const P3 &prev_pos, Launch &L) const;
};
#endif

// surface detectors: `inline void mcsim_<loc>_detector_deposit(McSim *,
// mc_point3f_t const *pos, mc_point3f_t const *dir, mc_fp_t weight)`
#if XO_USER_DET_TOP
struct DetUserTop {
	McTopDetector d;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float opl) const;
};
#endif
#if XO_USER_DET_BOTTOM
struct DetUserBottom {
	McBottomDetector d;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float opl) const;
};
#endif
#if XO_USER_DET_OUTER
struct DetUserOuter {               // cylindrical geometry: mcsim_outer_detector_deposit
	McOuterDetector d;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float opl) const;
};
#endif
#if XO_USER_DET_SPECULAR
struct DetUserSpecular {
	McSpecularDetector d;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	__device__ __forceinline__ void deposit(const Accu &acc, const P3 &pos, const P3 &dir, float w, float opl) const;
};
#endif

#if XO_USER_TRACE
// trace: `inline int mcsim_trace_event(McSim *, mc_uint_t event_count)` (1: the event was
// recorded) and `inline void mcsim_trace_complete(McSim *, mc_uint_t event_count)`
// (mctrace.py:541-585); the loops reach them through the trace_event / trace_complete
// overloads below
struct TraceUser { McTrace t; };
__device__ __forceinline__ bool trace_event(const TraceUser &t, float *fbuf, u32 packet,
	u32 count, u32 flags, const P3 &pos, const P3 &dir, float w, float opl);
__device__ __forceinline__ void trace_complete(const TraceUser &t, i32 *ibuf, u32 packet, u32 count);
#endif

// sample surface layouts (mcml): `inline int mcsim_<loc>_surface_layout_handler(McSim *,
// mc_fp_t *n2, mc_fp_t *cc)` returning MC_SURFACE_LAYOUT_CONTINUE / MC_REFLECTED /
// MC_REFRACTED (mcsurface/base.py:265-287, mcml.template.h:262-305); called through the
// surf_handle overloads below instead of the `handle` of the hand-written layouts
struct MlLayer;
#if XO_USER_SURF_TOP
struct SurfUserTop {
	McTopSurfaceLayout l;
	static constexpr bool active = true;
};
__device__ __forceinline__ int surf_handle(const SurfUserTop &s, Rng &rng, const P3 &pos, P3 &dir,
	float &weight, float *n2, float *cc, const MlLayer *layers, i32 num_layers, i32 &layer);
#endif
#if XO_USER_SURF_BOTTOM
struct SurfUserBottom {
	McBottomSurfaceLayout l;
	static constexpr bool active = true;
};
__device__ __forceinline__ int surf_handle(const SurfUserBottom &s, Rng &rng, const P3 &pos, P3 &dir,
	float &weight, float *n2, float *cc, const MlLayer *layers, i32 num_layers, i32 &layer);
#endif

}  // namespace xo
