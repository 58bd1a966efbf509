// mcml_kernel.cuh -- persistent-thread photon-packet kernel, planar layers.
//
// B200 counterpart of `McKernel` in xopto/mcml/kernel/mcml.template.c:346-824.
// Same physics, same per-work-item MWC stream usage (draw order: step ->
// [Fresnel] -> azimuth -> polar -> [lottery]; launch draws first), different
// machine mapping:
//   * one packet per thread, regenerated in place until the packet budget is
//     exhausted, packet state entirely in registers;
//   * lane state machine (throughput mode): RUN / BND (step ends on an interface)
//     / DEAD (needs a packet) / DRY.  The common trip pays one VOTE; a *service
//     round* fires when `refill` lanes of the warp wait (or every lane that still
//     has work waits) and runs, jointly for the waiting lanes: the interface
//     physics (move onto the interface, Fresnel, detector deposit, reload of the
//     layer constants), then the refill of the warp's launch queue (all 32 lanes
//     run the source code together, one atomic claims 32 packet indices, the
//     launched packets park in shared memory), then the pops.  The rare, long
//     paths therefore run with several lanes instead of the 1-2 that happen to
//     need them;
//   * layer table (+ derived per-layer constants + pf lookup tables) staged once
//     per CTA in shared memory; the constants of the *current* layer are cached
//     in registers and reloaded only when the packet changes layer;
//   * plugin parameter structs are __grid_constant__ kernel parameters;
//   * detector bins privatised per CTA in shared memory (Accu); fluence grids
//     through a CTA-private window in shared memory (32-bit partial sums, exact
//     64-bit totals) and RED.E.ADD.64 for the cells outside it;
//   * packet scheduling: deterministic mode = static block schedule (work-item t
//     owns packets [base_t, base_t+n_t)), a legal outcome of the reference's
//     racing atomic counter (mcml.template.c:460,790); throughput mode = the
//     same counter, claimed 32 packets at a time by a warp (launch queue).
//
// Two loops share prologue and epilogue: the deterministic loop evaluates the
// reference's expressions in the reference's order with DetMath, each work-item
// launching its own packets from its own MWC stream (bit-exact against the
// oracle); the throughput loop is the same physics re-associated for the SM
// (step constants folded, one MUFU per transcendental, the boundary division
// only on the divergent boundary path, Fresnel from a precomputed index ratio).
//
// The translation unit that includes this header must define the configuration:
//   typedef ... XoPf; XoSource; XoDetTop; XoDetBottom; XoDetSpecular; XoFluence;
//   XO_METHOD, XO_USE_LOTTERY, XO_WEIGHT_MIN, XO_LOTTERY_CHANCE, XO_TRACE,
//   XO_TRACK_OPL, XO_DETERMINISTIC, XO_USE_RMAX, XO_BLOCK, XO_MIN_BLOCKS
#pragma once
#include "xo_core.cuh"
#include "xo_pf.cuh"
#include "xo_detectors.cuh"
#include "xo_fluence.cuh"
#include "xo_surface.cuh"
#include "mcml_sources.cuh"
#include "mcml_layer.cuh"

#ifndef XO_USE_RMAX
#define XO_USE_RMAX 1
#endif
#ifndef XO_DEPOSIT_MAGIC
#define XO_DEPOSIT_MAGIC 0
#endif

namespace xo {

// Per-layer records of the throughput body, derived once per CTA when the medium
// is staged in shared memory.  `hot`, `pf` (+ `aux`) are exactly the register
// cache of the current layer, 16-byte aligned so a layer change reloads them
// with a few LDS.128; `iface` holds what the interface physics needs.
struct __align__(16) MlHot { float top, bottom, step_k, step_b; };
	// step_k = -ln2/mut (AW, AR) or -ln2/mus (MBL), step_b = -32 step_k:
	// step = lg2(raw draw)*step_k + step_b = -ln(u)/mut
struct __align__(16) MlAbs { float absorb, survive, dep_k, mua; };
	// absorb = mua/mut, survive = 1 - absorb, dep_k = absorb x (weight -> fixed
	// point factor of the fluence plugin): AW deposits f2u(w*dep_k + 0.5)
struct __align__(16) MlIface { float n12_top, cc_top, n12_bottom, cc_bottom; };
	// n12 = n/n(neighbour), exactly 1 when the indices are equal
struct __align__(16) MlAux { float n, pad0, pad1, pad2; };
struct __align__(16) MlPfFast { XoPf::Fast v; };
struct MlFastLayer { MlHot hot; MlAbs abs; MlPfFast pf; MlAux aux; MlIface iface; };

typedef Detectors<XoDetTop, XoDetBottom, XoDetSpecular> XoDetectors;
typedef SurfaceLayouts<XoSurfTop, XoSurfBottom> XoSurface;

#define XO_NEEDS_OPL (XO_TRACK_OPL || XoDetTop::needs_opl || XoDetBottom::needs_opl || \
	XoDetSpecular::needs_opl || XoFluence::needs_opl)

// Fresnel / Snell at a layer interface (mcml.template.c:80-203), reference
// operation order.  Returns the event flag and updates dir / layer index.  A
// uniform draw is consumed only when the indices differ and incidence is above
// the critical angle.
__device__ __forceinline__ u32 ml_boundary(const MlLayer &cur, const MlLayer &nxt,
		P3 &dir, i32 &layer, i32 next_layer, Rng &rng,
		const XoSurface &surface, const P3 &pos, float &weight, i32 num_layers) {
	float cc = (dir.z < 0.0f) ? cur.cc_top : cur.cc_bottom;
	float n1 = cur.n, n2 = nxt.n;
	// surface layouts (mcml.template.c:100-126): may override n2 / cc at the
	// point of incidence, or reflect the packet themselves
	if (XoSurfTop::active && next_layer == 0) {
		const int r = surf_handle(surface.top, rng, pos, dir, weight, &n2, &cc, &cur - layer, num_layers, layer);
		if (r == SURF_REFLECTED) return EV_REFLECTION;
		if (r == SURF_REFRACTED) return EV_REFRACTION;
	} else if (XoSurfBottom::active && next_layer == num_layers - 1) {
		const int r = surf_handle(surface.bottom, rng, pos, dir, weight, &n2, &cc, &cur - layer, num_layers, layer);
		if (r == SURF_REFLECTED) return EV_REFLECTION;
		if (r == SURF_REFRACTED) return EV_REFRACTION;
	}
	if (n1 == n2) { layer = next_layer; return EV_REFRACTION; }
	dir.z = -dir.z;
	float cos1 = fabsf(dir.z);
	if (cos1 > cc) {
		float n12 = M::div(n1, n2);
		float sin1 = M::sqrt(1.0f - cos1*cos1);
		if (cos1 >= 1.0f) sin1 = 0.0f;
		float sin2 = fminf(1.0f, n12*sin1);
		float cos2 = M::sqrt(1.0f - sin2*sin2);
		float nc1 = n12*cos1, nc2 = n12*cos2;
		float Rs = M::div(nc1 - cos2, nc1 + cos2); Rs *= Rs;
		float Rp = M::div(nc2 - cos1, nc2 + cos1); Rp *= Rp;
		float R = 0.5f*(Rs + Rp);
		if (cos1 <= 0.0f || sin2 == 1.0f) R = 1.0f;
		if (R < rng.next()) {
			layer = next_layer;
			dir.x *= n12;
			dir.y *= n12;
			dir.z = -copysignf(cos2, dir.z);
			return EV_REFRACTION;
		}
	}
	return EV_REFLECTION;
}

#if !XO_DETERMINISTIC
// The same interface physics for the throughput body: n12 = n1/n2 comes from
// the staged table, sin2^2 = n12^2 (1 - cos1^2) needs one square root instead
// of two, and the special cases of the reference (cos1 <= 0, sin2 == 1) fall
// out of the formula (cos1 > cc >= 0; cos2 == 0 gives Rs = Rp^2 = 1).
__device__ __forceinline__ bool ml_boundary_fast(float n12, float cc, P3 &dir, Rng &rng) {
	if (n12 == 1.0f) return true;
	float cos1 = fabsf(dir.z);
	if (cos1 > cc) {
		float s2 = (n12*n12)*fmaf(-cos1, cos1, 1.0f);
		float cos2 = FastMath::sqrt(fmaxf(1.0f - s2, 0.0f));
		float a = n12*cos1, b = n12*cos2;
		float rs = (a - cos2)*FastMath::rcp_approx(a + cos2);
		float rp = (b - cos1)*FastMath::rcp_approx(b + cos1);
		float R = 0.5f*fmaf(rs, rs, rp*rp);
		if (R*4294967296.0f < rng.next_raw()) {
			dir.x *= n12;
			dir.y *= n12;
			dir.z = copysignf(cos2, dir.z);
			return true;
		}
	}
	dir.z = -dir.z;
	return false;
}
#endif

}  // namespace xo

extern "C" __global__ void __launch_bounds__(XO_BLOCK, XO_MIN_BLOCKS)
McKernel(
	xo::u32 num_packets,
	xo::u32 *num_packets_done,
	xo::u32 *num_kernels,
	float rmax,
	xo::u64 *rng_state_x,
	const xo::u32 *rng_state_a,
	xo::u32 num_layers,
	const xo::MlLayer *layers,
	const __grid_constant__ XoSource source,
	const __grid_constant__ xo::XoSurface surface,
	const __grid_constant__ xo::XoTrace trace,
	const __grid_constant__ XoFluence fluence,
	const __grid_constant__ xo::XoDetectors detectors,
	const float *fp_lut,
	xo::i32 *int_buffer,
	float *float_buffer,
	xo::u64 *accumulator_buffer,
	xo::u32 lut_len,            // floats of fp_lut staged in shared memory (0: read global)
	xo::u32 priv_len,           // accumulator bins privatised per CTA
	const __grid_constant__ xo::FluWindow window,   // fluence cells privatised per CTA
	xo::u32 chunk,              // (unused by this kernel: packets are claimed 32 at a time per warp)
	xo::u32 refill)             // throughput mode: waiting lanes per warp that trigger a service round
{
	using namespace xo;
	extern __shared__ __align__(16) unsigned char xo_smem[];

	// ---- stage per-CTA tables -------------------------------------------------
	MlLayer *sh_layers = reinterpret_cast<MlLayer *>(xo_smem);
	u32 layer_words = num_layers*(u32)(sizeof(MlLayer)/4);
	{
		const u32 *src = reinterpret_cast<const u32 *>(layers);
		u32 *dst = reinterpret_cast<u32 *>(sh_layers);
		for (u32 i = threadIdx.x; i < layer_words; i += blockDim.x) dst[i] = src[i];
	}
	u32 off_words = (layer_words + 3u) & ~3u;
#if !XO_DETERMINISTIC
	MlFastLayer *sh_fast = reinterpret_cast<MlFastLayer *>(reinterpret_cast<u32 *>(xo_smem) + off_words);
	for (u32 i = threadIdx.x; i < num_layers; i += blockDim.x) {
		const MlLayer &Lg = layers[i];
		MlFastLayer F;
		F.hot.top = Lg.top; F.hot.bottom = Lg.bottom;
#if XO_ANISO
		// direction dependent: derived per flight segment (XO_DIR_CONSTS below)
		F.hot.step_k = 0.0f; F.hot.step_b = 0.0f;
		F.abs.absorb = 0.0f; F.abs.survive = 1.0f; F.abs.dep_k = 0.0f; F.abs.mua = 0.0f;
#else
#if XO_METHOD == 2
		F.hot.step_k = -0.6931471805599453f/Lg.mus;
#else
		F.hot.step_k = -0.6931471805599453f*Lg.inv_mut;
#endif
		F.hot.step_b = -32.0f*F.hot.step_k;
		F.abs.absorb = Lg.mua_inv_mut; F.abs.survive = 1.0f - Lg.mua_inv_mut;
		F.abs.dep_k = Lg.mua_inv_mut*fluence.fixed_scale(Lg.mua);
		F.abs.mua = Lg.mua;
#endif
		F.aux.n = Lg.n; F.aux.pad0 = 0.0f; F.aux.pad1 = 0.0f; F.aux.pad2 = 0.0f;
		float n_up = (i > 0) ? layers[i - 1].n : Lg.n;
		float n_dn = (i + 1 < num_layers) ? layers[i + 1].n : Lg.n;
		F.iface.n12_top = (n_up == Lg.n) ? 1.0f : Lg.n/n_up;
		F.iface.n12_bottom = (n_dn == Lg.n) ? 1.0f : Lg.n/n_dn;
		F.iface.cc_top = Lg.cc_top; F.iface.cc_bottom = Lg.cc_bottom;
		Lg.pf.prepare(F.pf.v);
		sh_fast[i] = F;
	}
	off_words += num_layers*(u32)(sizeof(MlFastLayer)/4);
#endif
	float *sh_lut = reinterpret_cast<float *>(xo_smem) + off_words;
	const float *lut = fp_lut;
	if (lut_len) {      // staged for the pf and the *Lut plugins
		for (u32 i = threadIdx.x; i < lut_len; i += blockDim.x) sh_lut[i] = fp_lut[i];
		lut = sh_lut;
		off_words += (lut_len + 3u) & ~3u;
	}
	Accu acc;
	acc.global = accumulator_buffer;
	acc.priv = reinterpret_cast<u32 *>(xo_smem) + off_words;
	acc.priv_len = priv_len;
	acc.zero_private();
	// 16 bytes in front of the window hold the constants of the deposits that
	// miss it (XoFluence::Far, one LDS.128 relative to the window address)
	const u32 far_words = (off_words + 2u*priv_len + 3u) & ~3u;
	acc.win = reinterpret_cast<u32 *>(xo_smem) + far_words + 4u;
	acc.bind();
	acc.lut = lut;
	const u32 win_len = window.ext0*window.ext1*window.ext2;
	for (u32 i = threadIdx.x; i < win_len; i += blockDim.x) acc.win[i] = 0;
#if !XO_DETERMINISTIC
	// per-warp launch queue: 32 slots of {pos, weight | dir, packet | layer, trace count}
	off_words = far_words + 4u + win_len;
	off_words = (off_words + 3u) & ~3u;
	float4 *q_a = reinterpret_cast<float4 *>(reinterpret_cast<u32 *>(xo_smem) + off_words) + (threadIdx.x & ~31u)*2u;
	float4 *q_b = q_a + 32;
	u32 *q_l = reinterpret_cast<u32 *>(xo_smem) + off_words + blockDim.x*8u + (threadIdx.x & ~31u);
	// constants of the fluence deposit, pinned in registers by a round trip through
	// shared memory (XoFluence::Prep)
	__shared__ typename XoFluence::Prep sh_flu_prep;
	if (threadIdx.x == 0) {
		sh_flu_prep = fluence.prepare(window);
		if (sizeof(typename XoFluence::Far) == 16)
			*reinterpret_cast<typename XoFluence::Far *>(acc.win - 4) = fluence.prepare_far(window);
	}
#endif
	__syncthreads();

	const u32 gid = blockIdx.x*blockDim.x + threadIdx.x;
	Rng rng;
	rng.load(rng_state_x[gid]);
	rng.a = rng_state_a[gid];
	MlCtx ctx; ctx.layers = sh_layers; ctx.num_layers = (i32)num_layers; ctx.lut = lut;
#if XO_USE_RMAX
	const P3 src_pos = source.origin();
	const float rmax2 = rmax*rmax;
#else
	(void)rmax;
#endif
	const XoTraceCfg &tcfg = *reinterpret_cast<const XoTraceCfg *>(&trace);
	(void)tcfg;

	bool started = false;
	u32 iterations = 0;
	// packet state (registers)
	P3 pos = { 0.0f, 0.0f, 0.0f }, dir = { 0.0f, 0.0f, 1.0f };
	float weight = 0.0f;
	i32 layer = 1;
	float opl = 0.0f;
	u32 packet = 0, trace_count = 0, flags = 0;
	(void)opl; (void)packet; (void)trace_count; (void)flags;

#if XO_USE_RMAX
#define XO_RMAX_TEST() do { \
		float ex_ = pos.x - src_pos.x, ey_ = pos.y - src_pos.y, ez_ = pos.z - src_pos.z; \
		if (ex_*ex_ + ey_*ey_ + ez_*ez_ > rmax2) { done = true; flags |= EV_ESCAPED; } \
	} while (0)
#else
#define XO_RMAX_TEST() do { } while (0)
#endif
#if XO_TRACE
#define XO_TRACE_TRIP() do { \
		flags |= done ? EV_TERMINATED : 0u; \
		if (XO_TRACE == XO_TRACE_ALL || ((XO_TRACE & XO_TRACE_END) && done)) { \
			if (trace_event(tcfg, float_buffer, packet, trace_count, flags, \
					pos, dir, weight, opl)) ++trace_count; \
		} \
		if (done) trace_complete(tcfg, int_buffer, packet, trace_count); \
	} while (0)
#else
#define XO_TRACE_TRIP() do { } while (0)
#endif

#if XO_DETERMINISTIC
	// ======== deterministic loop: reference expressions, reference order =========
	// (mcml.template.c:456-816; one packet after the other from this work-item's
	// static quota, launch and interface physics inline)
	u32 pk_next, pk_end;
	static_quota(num_packets, gridDim.x*blockDim.x, gid, &pk_next, &pk_end);
	(void)refill;
	if (pk_next < pk_end) {
		started = true;
		bool done = false;

#define XO_LAUNCH_PACKET() do { \
		Launch L_; \
		packet = pk_next++; \
		source.launch(rng, ctx, L_); \
		pos = L_.pos; dir = L_.dir; weight = L_.weight; layer = L_.layer; \
		if (XoDetSpecular::active) \
			detectors.specular.deposit(acc, L_.pos, L_.spec_dir, L_.spec_weight, 0.0f); \
		flags |= EV_LAUNCH; \
		if (XO_TRACE & XO_TRACE_START) { \
			if (trace_event(tcfg, float_buffer, packet, trace_count, flags, \
					pos, dir, weight, opl)) ++trace_count; \
		} \
	} while (0)

		XO_LAUNCH_PACKET();
		while (!done) {
			const MlLayer &L = sh_layers[layer];
			float step;
			++iterations;
#if XO_METHOD == 2
			step = M::div(-M::log(rng.next()), L.mus_at(dir));
#else
			step = -M::log(rng.next())*L.inv_mut_at(dir);
#endif
			step = fminf(step, XO_FLT_MAX);
			i32 next_layer = layer;
			const float top = L.top, bottom = L.bottom;
			if (pos.z + step*dir.z < top) {
				--next_layer;
				if (fabsf(dir.z) != 0.0f) step = M::div(top - pos.z, dir.z);
			}
			if (pos.z + step*dir.z >= bottom) {
				++next_layer;
				if (fabsf(dir.z) != 0.0f) step = M::div(bottom - pos.z, dir.z);
			}
			pos.x = pos.x + dir.x*step;
			pos.y = pos.y + dir.y*step;
			pos.z = pos.z + dir.z*step;
			if (XO_NEEDS_OPL) opl += L.n*step;
			if (layer < next_layer) pos.z = bottom;
			if (layer > next_layer) pos.z = top;
#if XO_METHOD == 2
			{   // microscopic Beer-Lambert (mcml.template.c:584-666)
				float mua = L.mua_at(dir);
				float frac = 1.0f - M::exp(-mua*step);
				float deposit = frac*weight;
				weight -= deposit;
				flags |= EV_ABSORPTION;
				if (XoFluence::active) {
					float back = (mua != 0.0f) ?
						step - M::div(-M::log(1.0f - rng.next()*frac), mua) : 0.0f;
					P3 dp = { pos.x - back*dir.x, pos.y - back*dir.y, pos.z - back*dir.z };
					fluence.deposit(acc, window, dp, deposit, mua, opl);
				}
			}
#endif
			if (next_layer != layer) {
				u32 bf = ml_boundary(L, sh_layers[next_layer], dir, layer, next_layer, rng,
					surface, pos, weight, (i32)num_layers);
				flags |= bf | EV_BOUNDARY_HIT;
				if (layer <= 0) {
					if (XoDetTop::active) detectors.top.deposit(acc, pos, dir, weight, opl);
					done = true;
				} else if (layer >= (i32)num_layers - 1) {
					if (XoDetBottom::active) detectors.bottom.deposit(acc, pos, dir, weight, opl);
					done = true;
				}
			} else {
#if XO_METHOD == 1
				// albedo rejection (mcml.template.c:705-721)
				if (rng.next() < L.mua_inv_mut_at(dir)) {
					float deposit = weight;
					done = true;
					weight -= deposit;
					flags |= EV_ABSORPTION;
					if (XoFluence::active) fluence.deposit(acc, window, pos, deposit, L.mua_at(dir), opl);
				} else {
					pf_scatter(L.pf, rng, lut, dir);
					flags |= EV_SCATTERING;
				}
#else
#if XO_METHOD == 0
				{   // albedo weight (mcml.template.c:722-731)
					float deposit = weight*L.mua_inv_mut_at(dir);
					weight -= deposit;
					flags |= EV_ABSORPTION;
					if (XoFluence::active) fluence.deposit(acc, window, pos, deposit, L.mua_at(dir), opl);
				}
#endif
				pf_scatter(L.pf, rng, lut, dir);
				flags |= EV_SCATTERING;
#endif
			}
#if XO_METHOD != 1
			// survival lottery (mcml.template.c:737-753); for AW only after a
			// scattering event, for MBL after every step
			if ((XO_METHOD == 2 || !(flags & EV_BOUNDARY_HIT)) && weight < XO_WEIGHT_MIN) {
#if XO_USE_LOTTERY
				if (rng.next() > XO_LOTTERY_CHANCE) done = true;
				else weight = M::div(weight, XO_LOTTERY_CHANCE);
#else
				done = true;
#endif
			}
#endif
			XO_RMAX_TEST();
			XO_TRACE_TRIP();
			flags = 0;
			if (done && pk_next < pk_end) {
				trace_count = 0;
				opl = 0.0f;
				XO_LAUNCH_PACKET();
				done = false;
			}
		}
	}
#undef XO_LAUNCH_PACKET
#else
	// ======== throughput loop ======================================================
	// Lane states.  RUN: a packet is in flight.  BND: the step of the packet ends
	// on the top / bottom interface of its layer, interface physics pending.  DEAD:
	// needs a packet from the warp's launch queue.  DRY: no packets left.
	enum : u32 { ST_RUN = 0, ST_DRY = 1, ST_BND = 2, ST_DEAD = 3 };   // waiting: >= ST_BND
	const u32 lane = threadIdx.x & 31u;
	const u32 lanemask_lt = (1u << lane) - 1u;
	u32 state = ST_DEAD;
	u32 q_count = 0;                // warp-uniform: packets parked in the warp's launch queue
	bool q_dry = false;             // warp-uniform: the packet budget is exhausted
	u32 thr_eff = refill < 1u ? 1u : (refill > 32u ? 32u : refill);   // waiting lanes that trigger a service round
	(void)chunk;
	// constants of the current layer (registers; reloaded on layer change)
	MlHot c_hot = { 0.0f, 0.0f, 0.0f, 0.0f };
	MlAbs c_abs = { 0.0f, 1.0f, 0.0f, 0.0f };
	MlAux c_aux = { 1.0f, 0.0f, 0.0f, 0.0f };
	XoPf::Fast c_pf;
	(void)c_aux;
	const typename XoFluence::Prep flu_prep = sh_flu_prep;
	(void)flu_prep;
#if XO_ANISO
	// anisotropic layers: the step / absorption constants belong to a (layer, direction)
	// pair and are rederived whenever either changes (launch, interface, scattering)
#define XO_DIR_CONSTS() do { \
		const MlLayer &L_ = sh_layers[layer]; \
		const float mut_ = tensor_project(L_.mut_t, dir), mua_ = tensor_project(L_.mua_t, dir); \
		const float inv_ = (mut_ != 0.0f) ? FastMath::rcp_approx(mut_) : XO_INF; \
		c_hot.step_k = (XO_METHOD == 2) ? \
			-0.6931471805599453f*FastMath::rcp_approx(tensor_project(L_.mus_t, dir)) : \
			-0.6931471805599453f*inv_; \
		c_hot.step_b = -32.0f*c_hot.step_k; \
		c_abs.absorb = (mua_ != 0.0f) ? mua_*inv_ : 0.0f; \
		c_abs.survive = 1.0f - c_abs.absorb; \
		c_abs.dep_k = c_abs.absorb*fluence.fixed_scale(mua_); \
		c_abs.mua = mua_; \
	} while (0)
#else
#define XO_DIR_CONSTS() do { } while (0)
#endif
#define XO_LOAD_LAYER(idx) do { \
		const MlFastLayer &F_ = sh_fast[idx]; \
		c_hot = F_.hot; c_abs = F_.abs; c_pf = F_.pf.v; \
		if (XO_NEEDS_OPL) c_aux = F_.aux; \
		XO_DIR_CONSTS(); \
	} while (0)
#define XO_END_TRIP() do { \
		XO_RMAX_TEST(); \
		XO_TRACE_TRIP(); \
		flags = 0; \
		state = done ? ST_DEAD : ST_RUN; \
	} while (0)
#if XO_USE_LOTTERY
#define XO_LOTTERY() do { \
		if (weight < XO_WEIGHT_MIN) { \
			if (rng.next_raw() > XO_LOTTERY_CHANCE*4294967296.0f) done = true; \
			else weight *= (1.0f/XO_LOTTERY_CHANCE); \
		} \
	} while (0)
#else
#define XO_LOTTERY() do { if (weight < XO_WEIGHT_MIN) done = true; } while (0)
#endif

#define XO_NEEDS_PACKET() (state == ST_DEAD)
	for (;;) {
		// ---- service round ------------------------------------------------------------
		// One vote per trip.  Lanes waiting at an interface (BND_*) or for a new packet
		// (DEAD) idle until `refill` lanes of the warp wait (or every lane that still
		// has work waits); then the interface physics, the queue refill and the pops
		// run jointly, in this order -- a packet that leaves the medium in the
		// interface handler is replaced in the same round.  The rare, long paths run
		// with several lanes instead of 1-2, and the common trip pays one VOTE.
		const u32 wait_mask = __ballot_sync(0xffffffffu, state >= ST_BND);
		if (__builtin_expect((u32)__popc(wait_mask) >= thr_eff, 0)) {
			// ---- interface physics ----------------------------------------------------
			if (state == ST_BND) {
				// (a packet reaches the top interface moving up, the bottom one moving down)
				const bool up = dir.z < 0.0f;
				bool done = false;
#if XO_METHOD != 2
				{   // move onto the interface (mcml.template.c:541-582)
					const float zb = up ? c_hot.top : c_hot.bottom;
					const float step = (dir.z != 0.0f) ? (zb - pos.z)*FastMath::rcp_approx(dir.z) : 0.0f;
					pos.x = fmaf(dir.x, step, pos.x);
					pos.y = fmaf(dir.y, step, pos.y);
					pos.z = zb;
					if (XO_NEEDS_OPL) opl = fmaf(c_aux.n, step, opl);
				}
#endif
				const MlIface I = sh_fast[layer].iface;
				float n12 = up ? I.n12_top : I.n12_bottom, cc = up ? I.cc_top : I.cc_bottom;
				bool through;
				int surf = SURF_CONTINUE;
				if ((XoSurfTop::active && up && layer == 1) ||
						(XoSurfBottom::active && !up && layer == (i32)num_layers - 2)) {
					// sample surface with a layout (mcml.template.c:100-126)
					const float n_out = sh_layers[up ? 0 : (i32)num_layers - 1].n;
					float n2 = n_out;
					surf = up ? surf_handle(surface.top, rng, pos, dir, weight, &n2, &cc, sh_layers, (i32)num_layers, layer)
						: surf_handle(surface.bottom, rng, pos, dir, weight, &n2, &cc, sh_layers, (i32)num_layers, layer);
					if (n2 != n_out) {
						const float n1 = sh_layers[layer].n;
						n12 = (n1 == n2) ? 1.0f : n1*FastMath::rcp_approx(n2);
					}
				}
				if (surf == SURF_REFLECTED) through = false;
				else if (surf == SURF_REFRACTED) through = true;     // (a user-written layout moved the packet across)
				else through = ml_boundary_fast(n12, cc, dir, rng);
				flags |= EV_BOUNDARY_HIT | (through ? EV_REFRACTION : EV_REFLECTION);
				if (through) {
					if (surf != SURF_REFRACTED) layer += up ? -1 : 1;
					if (layer <= 0) {
						if (XoDetTop::active) detectors.top.deposit(acc, pos, dir, weight, opl);
						done = true;
					} else if (layer >= (i32)num_layers - 1) {
						if (XoDetBottom::active) detectors.bottom.deposit(acc, pos, dir, weight, opl);
						done = true;
					} else {
						XO_LOAD_LAYER(layer);
					}
				} else {
					XO_DIR_CONSTS();                // reflected: new direction, same layer
				}
#if XO_METHOD == 2
				XO_LOTTERY();                   // MBL: lottery after every step
#endif
				XO_END_TRIP();
			}
			// ---- new packets for the lanes that need one -----------------------------------
			u32 dead_mask = __ballot_sync(0xffffffffu, XO_NEEDS_PACKET());
			while (dead_mask != 0u) {
				if (q_count == 0u) {
					if (q_dry) break;
					// queue empty: all 32 lanes launch one packet each into the queue
					u32 base = 0;
					if (lane == 0u) base = atomicAdd(num_packets_done, 32u);
					base = __shfl_sync(0xffffffffu, base, 0);
					const u32 n_new = base < num_packets ?
						(num_packets - base < 32u ? num_packets - base : 32u) : 0u;
					q_dry = n_new < 32u;
					if (lane < n_new) {
						Launch L_;
						source.launch(rng, ctx, L_);
						if (XoDetSpecular::active)
							detectors.specular.deposit(acc, L_.pos, L_.spec_dir, L_.spec_weight, 0.0f);
						u32 tc = 0;
						if (XO_TRACE & XO_TRACE_START) {
							if (trace_event(tcfg, float_buffer, base + lane, 0u, EV_LAUNCH,
									L_.pos, L_.dir, L_.weight, 0.0f)) tc = 1u;
						}
						q_a[lane] = make_float4(L_.pos.x, L_.pos.y, L_.pos.z, L_.weight);
						q_b[lane] = make_float4(L_.dir.x, L_.dir.y, L_.dir.z, __uint_as_float(base + lane));
						q_l[lane] = (u32)L_.layer | (tc << 16);
					}
					__syncwarp();
					q_count = n_new;
					if (n_new == 0u) break;
				}
				if (XO_NEEDS_PACKET()) {
					const u32 rank = (u32)__popc(dead_mask & lanemask_lt);
					if (rank < q_count) {
						const u32 slot = q_count - 1u - rank;
						const float4 a = q_a[slot], b = q_b[slot];
						const u32 l = q_l[slot];
						pos.x = a.x; pos.y = a.y; pos.z = a.z; weight = a.w;
						dir.x = b.x; dir.y = b.y; dir.z = b.z; packet = __float_as_uint(b.w);
						layer = (i32)(l & 0xffffu); trace_count = l >> 16;
						XO_LOAD_LAYER(layer);
						opl = 0.0f;
						flags = EV_LAUNCH;
						state = ST_RUN;
						started = true;
					}
				}
				__syncwarp();
				const u32 n_dead = (u32)__popc(dead_mask);
				if (n_dead <= q_count) { q_count -= n_dead; break; }
				// the queue ran out before every waiting lane had a packet: refill, go on
				q_count = 0u;
				dead_mask = __ballot_sync(0xffffffffu, XO_NEEDS_PACKET());
			}
			if (q_dry && q_count == 0u) {
				// packet budget exhausted: the lanes still waiting for a packet retire
				if (XO_NEEDS_PACKET()) state = ST_DRY;
				const u32 n_dry = (u32)__popc(__ballot_sync(0xffffffffu, state == ST_DRY));
				if (n_dry == 32u) break;
				thr_eff = refill < 32u - n_dry ? refill : 32u - n_dry;
			}
		}
		if (state != ST_RUN) continue;

		// ---- one step of the packet ----------------------------------------------------
		++iterations;
#if XO_METHOD == 2
		float step = fminf(fmaf(FastMath::lg2(rng.next_raw()), c_hot.step_k, c_hot.step_b), XO_FLT_MAX);
#else
		// (no clamp to FLT_MAX: an infinite step (draw == 0) always "hits" - z +- inf
		// or NaN fails the in-layer test - and the hit path recomputes the step)
		const float step = fmaf(FastMath::lg2(rng.next_raw()), c_hot.step_k, c_hot.step_b);
#endif
		const float zs = fmaf(step, dir.z, pos.z);
		const bool hit = !(zs >= c_hot.top && zs < c_hot.bottom);
#if XO_METHOD == 2
		if (hit) {
			const float zb = (zs < c_hot.top) ? c_hot.top : c_hot.bottom;
			if (dir.z != 0.0f) step = (zb - pos.z)*FastMath::rcp_approx(dir.z);
			pos.z = zb;
		} else {
			pos.z = zs;
		}
		pos.x = fmaf(dir.x, step, pos.x);
		pos.y = fmaf(dir.y, step, pos.y);
		if (XO_NEEDS_OPL) opl = fmaf(c_aux.n, step, opl);
		{
			float frac = 1.0f - FastMath::exp(-c_abs.mua*step);
			float deposit = frac*weight;
			weight -= deposit;
			flags |= EV_ABSORPTION;
			if (XoFluence::active) {
				float back = (c_abs.mua != 0.0f) ?
					step + FastMath::log(1.0f - rng.next()*frac)*FastMath::rcp_approx(c_abs.mua) : 0.0f;
				P3 dp = { pos.x - back*dir.x, pos.y - back*dir.y, pos.z - back*dir.z };
				fluence.deposit(acc, window, dp, deposit, c_abs.mua, opl);
			}
		}
		if (hit) {
			state = ST_BND;
			continue;
		}
#else
		if (__builtin_expect(hit, 0)) {
			// the move onto the interface (one division) is part of the deferred
			// interface handling: the packet stays where it is until then
			state = ST_BND;
			continue;
		}
		pos.z = zs;
		pos.x = fmaf(dir.x, step, pos.x);
		pos.y = fmaf(dir.y, step, pos.y);
		if (XO_NEEDS_OPL) opl = fmaf(c_aux.n, step, opl);
#endif
		bool done = false;
#if XO_METHOD == 1
		if (rng.next() < c_abs.absorb) {
			float deposit = weight;
			done = true;
			weight = 0.0f;
			flags |= EV_ABSORPTION;
			if (XoFluence::active) fluence.deposit(acc, window, pos, deposit, c_abs.mua, opl);
		} else {
			pf_scatter(c_pf, rng, lut, dir);
			flags |= EV_SCATTERING;
			XO_DIR_CONSTS();
		}
#else
#if XO_METHOD == 0
		{   // albedo weight: the deposit leaves as fixed point in one FFMA + LOP3
#if XO_DEPOSIT_MAGIC
			// round(w*dep_k) from the mantissa of w*dep_k + 2^23 (the host guarantees
			// w*dep_k < 2^23 - 1): no F2I on the XU pipe.  Round-to-nearest-even
			// instead of the reference's floor(x + 0.5): differs only on exact ties.
			const u32 wfix = __float_as_uint(fmaf(weight, c_abs.dep_k, 8388608.0f)) & 0x7fffffu;
#else
			const u32 wfix = f2u(fmaf(weight, c_abs.dep_k, 0.5f));
#endif
			if (XoFluence::active && !XoFluence::fixed_point)     // user-written fragment
				fluence.deposit(acc, window, pos, weight*c_abs.absorb, c_abs.mua, opl);
			weight *= c_abs.survive;
			flags |= EV_ABSORPTION;
			if (XoFluence::active && XoFluence::fixed_point)
				fluence.deposit_prep(acc, flu_prep, window, pos, wfix, opl);
		}
#endif
		pf_scatter(c_pf, rng, lut, dir);
		flags |= EV_SCATTERING;
		XO_DIR_CONSTS();
		XO_LOTTERY();
#endif
		XO_END_TRIP();
	}
#undef XO_LOAD_LAYER
#undef XO_DIR_CONSTS
#undef XO_NEEDS_PACKET
#undef XO_END_TRIP
#undef XO_LOTTERY
	// every lane drew from its stream (queue refills), whether or not it ever
	// carried a packet: all states go back
	rng_state_x[gid] = rng.state();
#endif  // XO_DETERMINISTIC
#undef XO_RMAX_TEST
#undef XO_TRACE_TRIP
	if (started) {
#if XO_DETERMINISTIC
		rng_state_x[gid] = rng.state();
#endif
		atomicAdd(num_kernels, 1u);
	}
	// loop-trip count (the roofline's unit of work): one 64-bit RED per warp
	{
		u32 warp_iters = __reduce_add_sync(0xffffffffu, iterations);
		if ((threadIdx.x & 31u) == 0u && warp_iters)
			atomicAdd(reinterpret_cast<u64 *>(num_kernels + 1), (u64)warp_iters);
	}

	__syncthreads();
	acc.flush_private();
	if (XoFluence::active) flush_window(fluence, acc, window);
#if XO_DETERMINISTIC
	if (gid == 0) *num_packets_done = num_packets;
#endif
	(void)int_buffer; (void)float_buffer;
}
