// mcml_kernel.cuh -- persistent-thread photon-packet kernel, planar layers.
//
// B200 counterpart of `McKernel` in xopto/mcml/kernel/mcml.template.c:346-824.
// Same physics, same per-work-item MWC stream usage (draw order: step ->
// [Fresnel] -> azimuth -> polar -> [lottery]; launch draws first), different
// machine mapping:
//   * one packet per thread, regenerated in place until the packet budget is
//     exhausted, packet state entirely in registers;
//   * deferred rare work: the long, rarely taken paths of the loop -- a layer
//     interface (Fresnel, detector deposit, reload of the layer constants) and
//     the launch of a new packet -- are not executed where they occur.  The lane
//     parks in a PENDING state and idles until at least `refill` lanes of its
//     warp are pending; then all of them run the deferred code together.  With
//     one packet per lane these paths otherwise execute with 1-2 active lanes in
//     most loop trips (ncu: 27 % of the issue slots of the 5-layer skin case);
//   * layer table (+ derived per-layer constants + pf lookup tables) staged once
//     per CTA in shared memory; the constants of the *current* layer are cached
//     in registers and reloaded only when the packet changes layer;
//   * plugin parameter structs are __grid_constant__ kernel parameters;
//   * detector bins privatised per CTA in shared memory (Accu), fluence through
//     RED.E.ADD.64;
//   * packet scheduling: deterministic mode = static block schedule (work-item t
//     owns packets [base_t, base_t+n_t)), a legal outcome of the reference's
//     racing atomic counter (mcml.template.c:460,790); throughput mode = the
//     same counter, but claimed in chunks of `chunk` packets per atomic.
//
// Two loop bodies share the skeleton: the deterministic body evaluates the
// reference's expressions in the reference's order with DetMath (bit-exact
// against the oracle); the throughput body is the same physics re-associated
// for the SM (step constants folded, one MUFU per transcendental, the boundary
// division only on the divergent boundary path, Fresnel from a precomputed
// index ratio).
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
#include "mcml_sources.cuh"

#ifndef XO_USE_RMAX
#define XO_USE_RMAX 1
#endif

namespace xo {

struct MlLayer {                    // mcml/mclayer/layer.py:57-69
	float thickness, top, bottom, n, cc_top, cc_bottom, mus, mua, inv_mut, mua_inv_mut;
	XoPf pf;
};

// Per-layer records of the throughput body, derived once per CTA when the medium
// is staged in shared memory.  `hot`, `pf` (+ `aux`) are exactly the register
// cache of the current layer, 16-byte aligned so a layer change reloads them
// with a few LDS.128; `iface` holds what the interface physics needs.
struct __align__(16) MlHot { float top, bottom, step_k, absorb; };
	// step_k = -ln2/mut (AW, AR) or -ln2/mus (MBL): step = lg2(u)*step_k
struct __align__(16) MlIface { float n12_top, cc_top, n12_bottom, cc_bottom; };
	// n12 = n/n(neighbour), exactly 1 when the indices are equal
struct __align__(16) MlAux { float mua, n, pad0, pad1; };
struct __align__(16) MlPfFast { XoPf::Fast v; };
struct MlFastLayer { MlHot hot; MlPfFast pf; MlAux aux; MlIface iface; };

typedef Detectors<XoDetTop, XoDetBottom, XoDetSpecular> XoDetectors;
#if XO_TRACE
typedef TraceCfg XoTrace;
#else
typedef TraceNone XoTrace;
#endif

#define XO_NEEDS_OPL (XO_TRACK_OPL || XoDetTop::needs_opl || XoDetBottom::needs_opl || \
	XoDetSpecular::needs_opl || XoFluence::needs_opl)

struct MlCtx {
	const MlLayer *layers;          // shared memory
	static constexpr bool has_specular = XoDetSpecular::active;
	__device__ __forceinline__ float layer_n(int i) const { return layers[i].n; }
	__device__ __forceinline__ float layer_cc_bottom(int i) const { return layers[i].cc_bottom; }
};

// Fresnel / Snell at a layer interface (mcml.template.c:80-203), reference
// operation order.  Returns the event flag and updates dir / layer index.  A
// uniform draw is consumed only when the indices differ and incidence is above
// the critical angle.
__device__ __forceinline__ u32 ml_boundary(const MlLayer &cur, const MlLayer &nxt,
		P3 &dir, i32 &layer, i32 next_layer, Rng &rng) {
	float cc = (dir.z < 0.0f) ? cur.cc_top : cur.cc_bottom;
	float n1 = cur.n, n2 = nxt.n;
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
	const __grid_constant__ xo::XoTrace trace,
	const __grid_constant__ XoFluence fluence,
	const __grid_constant__ xo::XoDetectors detectors,
	const float *fp_lut,
	xo::i32 *int_buffer,
	float *float_buffer,
	xo::u64 *accumulator_buffer,
	xo::u32 lut_len,            // floats of fp_lut staged in shared memory (0: read global)
	xo::u32 priv_len,           // accumulator bins privatised per CTA
	xo::u32 chunk,              // packets claimed per atomic (throughput mode)
	xo::u32 refill)             // pending lanes per warp that trigger the deferred work (1..32)
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
		F.hot.top = Lg.top; F.hot.bottom = Lg.bottom; F.hot.absorb = Lg.mua_inv_mut;
#if XO_METHOD == 2
		F.hot.step_k = -0.6931471805599453f/Lg.mus;
#else
		F.hot.step_k = -0.6931471805599453f*Lg.inv_mut;
#endif
		F.aux.mua = Lg.mua; F.aux.n = Lg.n; F.aux.pad0 = 0.0f; F.aux.pad1 = 0.0f;
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
	if (XoPf::uses_lut && lut_len) {
		for (u32 i = threadIdx.x; i < lut_len; i += blockDim.x) sh_lut[i] = fp_lut[i];
		lut = sh_lut;
		off_words += (lut_len + 3u) & ~3u;
	}
	Accu acc;
	acc.global = accumulator_buffer;
	acc.priv = reinterpret_cast<u32 *>(xo_smem) + off_words;
	acc.priv_len = priv_len;
	acc.zero_private();
	__syncthreads();

	const u32 gid = blockIdx.x*blockDim.x + threadIdx.x;
	Rng rng;
	rng.x = rng_state_x[gid];
	rng.a = rng_state_a[gid];
	MlCtx ctx; ctx.layers = sh_layers;
#if XO_USE_RMAX
	const P3 src_pos = source.origin();
	const float rmax2 = rmax*rmax;
#else
	(void)rmax;
#endif
	const TraceCfg &tcfg = *reinterpret_cast<const TraceCfg *>(&trace);
	(void)tcfg;

	// ---- packet budget -------------------------------------------------------
	Budget budget;
	budget.dry = false;
#if XO_DETERMINISTIC
	static_quota(num_packets, gridDim.x*blockDim.x, gid, &budget.next, &budget.end);
#else
	budget.next = 0; budget.end = 0;
#endif

	// ---- lane state ------------------------------------------------------------
	// RUN: a packet is in flight.  BND_*: the packet sits on the top / bottom
	// interface of its layer, interface physics pending.  DEAD: needs a new
	// packet.  DRY: no packets left for this lane.
	enum : u32 { ST_RUN = 0, ST_BND_TOP = 1, ST_BND_BOTTOM = 2, ST_DEAD = 3, ST_DRY = 4 };
	u32 state = ST_DEAD;
	u32 n_dry = 0;                  // warp-uniform count of DRY lanes
	bool started = false;
	u32 iterations = 0;

	// packet state (registers)
	P3 pos = { 0.0f, 0.0f, 0.0f }, dir = { 0.0f, 0.0f, 1.0f };
	float weight = 0.0f;
	i32 layer = 1;
	float opl = 0.0f;
	u32 packet = 0, trace_count = 0, flags = 0;
	(void)opl; (void)packet; (void)trace_count; (void)flags;
#if !XO_DETERMINISTIC
	// constants of the current layer (registers; reloaded on layer change)
	MlHot c_hot = { 0.0f, 0.0f, 0.0f, 0.0f };
	MlAux c_aux = { 0.0f, 1.0f, 0.0f, 0.0f };
	XoPf::Fast c_pf;
	(void)c_aux;
#define XO_LOAD_LAYER(idx) do { \
		const MlFastLayer &F_ = sh_fast[idx]; \
		c_hot = F_.hot; c_pf = F_.pf.v; \
		if (XO_NEEDS_OPL || XO_METHOD != 0 || XO_FLUENCE_RATE) c_aux = F_.aux; \
	} while (0)
#else
#define XO_LOAD_LAYER(idx) do { } while (0)
#endif

	// end of a loop trip for this lane: rmax test, trace event, termination
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
		if (done) int_buffer[tcfg.count_off + packet] = (i32)trace_count; \
	} while (0)
#else
#define XO_TRACE_TRIP() do { } while (0)
#endif
#define XO_END_TRIP() do { \
		XO_RMAX_TEST(); \
		XO_TRACE_TRIP(); \
		flags = 0; \
		state = done ? ST_DEAD : ST_RUN; \
	} while (0)
#if XO_USE_LOTTERY
#define XO_LOTTERY() do { \
		if (weight < XO_WEIGHT_MIN) { \
			if (rng.next() > XO_LOTTERY_CHANCE) done = true; \
			else weight = XO_DETERMINISTIC ? M::div(weight, XO_LOTTERY_CHANCE) \
				: weight*(1.0f/XO_LOTTERY_CHANCE); \
		} \
	} while (0)
#else
#define XO_LOTTERY() do { if (weight < XO_WEIGHT_MIN) done = true; } while (0)
#endif

	for (;;) {
		const u32 n_run = (u32)__popc(__ballot_sync(0xffffffffu, state == ST_RUN));
		const u32 n_pending = 32u - n_dry - n_run;
		if (n_pending >= refill || n_run == 0u) {
			if (n_pending == 0u) break;         // every lane is DRY
			// ======== deferred work, executed jointly by the pending lanes ========
			if (state == ST_BND_TOP || state == ST_BND_BOTTOM) {
				const bool up = (state == ST_BND_TOP);
				bool done = false;
#if XO_DETERMINISTIC
				const i32 next_layer = up ? layer - 1 : layer + 1;
				u32 bf = ml_boundary(sh_layers[layer], sh_layers[next_layer], dir, layer, next_layer, rng);
				flags |= bf | EV_BOUNDARY_HIT;
				const bool through = (layer == next_layer);
#else
				const MlIface I = sh_fast[layer].iface;
				const bool through = ml_boundary_fast(up ? I.n12_top : I.n12_bottom,
					up ? I.cc_top : I.cc_bottom, dir, rng);
				flags |= EV_BOUNDARY_HIT | (through ? EV_REFRACTION : EV_REFLECTION);
				if (through) layer += up ? -1 : 1;
#endif
				if (through) {
					if (layer <= 0) {
						if (XoDetTop::active) detectors.top.deposit(acc, pos, dir, weight, opl);
						done = true;
					} else if (layer >= (i32)num_layers - 1) {
						if (XoDetBottom::active) detectors.bottom.deposit(acc, pos, dir, weight, opl);
						done = true;
					} else {
						XO_LOAD_LAYER(layer);
					}
				}
#if XO_METHOD == 2
				XO_LOTTERY();                   // MBL: lottery after every step
#endif
				XO_END_TRIP();
			}
			if (state == ST_DEAD) {
				if (budget.claim(num_packets, num_packets_done, chunk, &packet)) {
					Launch L_;
					source.launch(rng, ctx, L_);
					pos = L_.pos; dir = L_.dir; weight = L_.weight; layer = L_.layer;
					if (XoDetSpecular::active)
						detectors.specular.deposit(acc, L_.pos, L_.spec_dir, L_.spec_weight, 0.0f);
					XO_LOAD_LAYER(layer);
					opl = 0.0f;
					trace_count = 0;
					flags = EV_LAUNCH;
					if (XO_TRACE & XO_TRACE_START) {
						if (trace_event(tcfg, float_buffer, packet, trace_count, flags,
								pos, dir, weight, opl)) ++trace_count;
					}
					state = ST_RUN;
					started = true;
				} else {
					state = ST_DRY;
				}
			}
			n_dry = (u32)__popc(__ballot_sync(0xffffffffu, state == ST_DRY));
			continue;
		}
		if (state != ST_RUN) continue;

		++iterations;
#if XO_DETERMINISTIC
		// ======== deterministic body: reference expressions, reference order ====
		const MlLayer &L = sh_layers[layer];
		float step;
#if XO_METHOD == 2
		step = M::div(-M::log(rng.next()), L.mus);
#else
		step = -M::log(rng.next())*L.inv_mut;
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
			float mua = L.mua;
			float frac = 1.0f - M::exp(-mua*step);
			float deposit = frac*weight;
			weight -= deposit;
			flags |= EV_ABSORPTION;
			if (XoFluence::active) {
				float back = (mua != 0.0f) ?
					step - M::div(-M::log(1.0f - rng.next()*frac), mua) : 0.0f;
				P3 dp = { pos.x - back*dir.x, pos.y - back*dir.y, pos.z - back*dir.z };
				fluence.deposit(acc, dp, deposit, mua, opl);
			}
		}
#endif
		if (next_layer != layer) {
			// interface physics deferred (mcml.template.c:669-703)
			state = (next_layer < layer) ? ST_BND_TOP : ST_BND_BOTTOM;
			continue;
		}
		bool done = false;
#if XO_METHOD == 1
		// albedo rejection (mcml.template.c:705-721)
		if (rng.next() < L.mua_inv_mut) {
			float deposit = weight;
			done = true;
			weight -= deposit;
			flags |= EV_ABSORPTION;
			if (XoFluence::active) fluence.deposit(acc, pos, deposit, L.mua, opl);
		} else {
			float fi, ct = L.pf.sample(rng, lut, &fi);
			scatter_direction(dir, ct, fi);
			flags |= EV_SCATTERING;
		}
#else
#if XO_METHOD == 0
		{   // albedo weight (mcml.template.c:722-731)
			float deposit = weight*L.mua_inv_mut;
			weight -= deposit;
			flags |= EV_ABSORPTION;
			if (XoFluence::active) fluence.deposit(acc, pos, deposit, L.mua, opl);
		}
#endif
		float fi, ct = L.pf.sample(rng, lut, &fi);
		scatter_direction(dir, ct, fi);
		flags |= EV_SCATTERING;
		// survival lottery (mcml.template.c:737-753); for AW only after a
		// scattering event, for MBL after every step
		XO_LOTTERY();
#endif
		XO_END_TRIP();
#else
		// ======== throughput body ===============================================
		float step = fminf((FastMath::lg2(rng.next_raw()) - 32.0f)*c_hot.step_k, XO_FLT_MAX);
		const float zs = fmaf(step, dir.z, pos.z);
		const bool hit_top = zs < c_hot.top;
		const bool hit = hit_top || zs >= c_hot.bottom;
		if (hit) {
			const float zb = hit_top ? c_hot.top : c_hot.bottom;
			if (dir.z != 0.0f) step = (zb - pos.z)*FastMath::rcp_approx(dir.z);
			pos.z = zb;
		} else {
			pos.z = zs;
		}
		pos.x = fmaf(dir.x, step, pos.x);
		pos.y = fmaf(dir.y, step, pos.y);
		if (XO_NEEDS_OPL) opl = fmaf(c_aux.n, step, opl);
#if XO_METHOD == 2
		{
			float frac = 1.0f - FastMath::exp(-c_aux.mua*step);
			float deposit = frac*weight;
			weight -= deposit;
			flags |= EV_ABSORPTION;
			if (XoFluence::active) {
				float back = (c_aux.mua != 0.0f) ?
					step + FastMath::log(1.0f - rng.next()*frac)*FastMath::rcp_approx(c_aux.mua) : 0.0f;
				P3 dp = { pos.x - back*dir.x, pos.y - back*dir.y, pos.z - back*dir.z };
				fluence.deposit(acc, dp, deposit, c_aux.mua, opl);
			}
		}
#endif
		if (hit) {
			state = hit_top ? ST_BND_TOP : ST_BND_BOTTOM;
			continue;
		}
		bool done = false;
#if XO_METHOD == 1
		if (rng.next() < c_hot.absorb) {
			float deposit = weight;
			done = true;
			weight = 0.0f;
			flags |= EV_ABSORPTION;
			if (XoFluence::active) fluence.deposit(acc, pos, deposit, c_aux.mua, opl);
		} else {
			float fi, ct = c_pf.sample(rng, lut, &fi);
			scatter_direction(dir, ct, fi);
			flags |= EV_SCATTERING;
		}
#else
#if XO_METHOD == 0
		{
			float deposit = weight*c_hot.absorb;
			weight -= deposit;
			flags |= EV_ABSORPTION;
			if (XoFluence::active) fluence.deposit(acc, pos, deposit, c_aux.mua, opl);
		}
#endif
		float fi, ct = c_pf.sample(rng, lut, &fi);
		scatter_direction(dir, ct, fi);
		flags |= EV_SCATTERING;
		XO_LOTTERY();
#endif
		XO_END_TRIP();
#endif  // XO_DETERMINISTIC
	}
#undef XO_LOAD_LAYER
#undef XO_RMAX_TEST
#undef XO_TRACE_TRIP
#undef XO_END_TRIP
#undef XO_LOTTERY
	if (started) {
		rng_state_x[gid] = rng.x;
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
#if XO_DETERMINISTIC
	if (gid == 0) *num_packets_done = num_packets;
#endif
	(void)int_buffer; (void)float_buffer;
}
