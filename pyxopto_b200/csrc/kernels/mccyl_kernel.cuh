// mccyl_kernel.cuh -- persistent-thread photon-packet kernel, concentric cylinders.
//
// B200 counterpart of `McKernel` in xopto/mccyl/kernel/mccyl.template.c:560-1001.
// The sample is a stack of infinitely long concentric cylinders around the z
// axis; layer 0 is the surrounding medium and the layer index grows inwards.
// Every loop iteration draws a fresh exponential step, solves the two quadratic
// ray / cylinder equations of the current layer (inner and outer radius), moves
// min(distance, step) and either handles the interface (Fresnel with the radial
// normal) or absorbs + scatters.  Machine mapping as in mcml_kernel.cuh:
//   * one packet per thread, regenerated in place, state in registers;
//   * layer table (+ pf lookup tables) staged per CTA in shared memory, plugin
//     structs as __grid_constant__ parameters;
//   * detector bins privatised per CTA in shared memory (Accu), fluence grids
//     through the CTA-private window / RED.E.ADD.64;
//   * deterministic mode: static block schedule, reference expressions in
//     reference order with DetMath; throughput mode: global packet counter
//     claimed in chunks, the quadratic solved once for both radii (shared a, b,
//     1/2a; one MUFU.RSQ-free square root per radius), direction sanity check
//     only where the direction is not renormalised (interface events).
//
// Reference quirks kept: per-work-item budget of 1e6 loop trips in deterministic
// mode (mccyl.template.c:681, `num_steps` is never reset between packets);
// termination on |dir| drifting from 1 by more than 10 eps (:928-934) and on
// weight <= 0 (:937); MBL is not available (the reference branch does not
// compile); AR follows mcml (the reference line :875 does not preprocess).
#pragma once
#include "xo_core.cuh"
#include "xo_pf.cuh"
#include "xo_detectors.cuh"
#include "xo_fluence.cuh"
#include "mccyl_sources.cuh"
#include "mccyl_layer.cuh"

#ifndef XO_USE_RMAX
#define XO_USE_RMAX 1
#endif
#if XO_METHOD == 2
#error "mccyl: the microscopic Beer-Lambert method is not available (mccyl.template.c:774)"
#endif

namespace xo {

// Per-layer records of the throughput loop, derived once per CTA when the medium
// is staged in shared memory; the record of the *current* layer is cached in
// registers (cf. MlFastLayer in mcml_kernel.cuh).
struct __align__(16) CylHot { float ri2, ro2, step_k, step_b; };
	// squared radii; step = lg2(raw draw)*step_k + step_b = -ln(u)/mut
struct __align__(16) CylAbs { float absorb, mua, n, pad; };
struct __align__(16) CylGeo { float ri, ro, pad0, pad1; };
	// radii for the radial-clearance test (ri = -inf for the innermost layer: no inner surface)
struct __align__(16) CylPfFast { XoPf::Fast v; };
struct CylFastLayer { CylHot hot; CylAbs abs; CylGeo geo; CylPfFast pf; };

typedef CylDetectors<XoDetOuter, XoDetSpecular> XoDetectors;

#define XO_NEEDS_OPL (XO_TRACK_OPL || XoDetOuter::needs_opl || \
	XoDetSpecular::needs_opl || XoFluence::needs_opl)
#define XO_CYL_MAX_STEPS 1000000

struct CylCtx {
	const CylLayer *layers;         // shared memory
	i32 num_layers;
	static constexpr bool has_specular = XoDetSpecular::active;
	__device__ __forceinline__ float layer_n(int i) const { return layers[i].n; }
	__device__ __forceinline__ float layer_r_inner(int i) const { return layers[i].r_inner; }
	__device__ __forceinline__ float layer_cc_inner(int i) const { return layers[i].cc_inner; }
};

// Fresnel / Snell at a cylinder surface (mccyl.template.c:249-400).  `inward`:
// the packet crosses the inner radius of its layer.  Returns the event flag,
// updates dir / layer; one uniform draw only when the indices differ and the
// incidence is above the critical angle.
__device__ __forceinline__ u32 cyl_boundary(const CylLayer &cur, const CylLayer &nxt,
		const P3 &pos, P3 &dir, i32 &layer, i32 next_layer, Rng &rng) {
	const bool inward = layer < next_layer;
	P3 normal = radial_normal(pos, !inward);
	float cc = inward ? cur.cc_inner : cur.cc_outer;
	float n1 = cur.n, n2 = nxt.n;
	if (n1 == n2) { layer = next_layer; return EV_REFRACTION; }
	float cos1 = fminf(fabsf(normal.x*dir.x + normal.y*dir.y), 1.0f);
	if (cos1 > cc) {
		float n12 = M::div(n1, n2);
		float sin1 = M::sqrt(1.0f - cos1*cos1);
		if (cos1 >= 1.0f) sin1 = 0.0f;
		float sin2 = fminf(1.0f, n12*sin1);
		float cos2 = M::sqrt(1.0f - sin2*sin2);
		float nc1 = n12*cos1, nc2 = n12*cos2;
		float Rs = M::div(nc1 - cos2, nc1 + cos2); Rs *= Rs;
		float Rp = M::div(nc2 - cos1, nc2 + cos1); Rp *= Rp;
		float R = 0.5f*(Rp + Rs);
		if (cos1 <= 0.0f || sin2 == 1.0f) R = 1.0f;
		if (R < rng.next()) {
			layer = next_layer;
			float k = n12*cos1 - cos2;
			dir.x = n12*dir.x - k*normal.x;
			dir.y = n12*dir.y - k*normal.y;
			dir.z = n12*dir.z;
			return EV_REFRACTION;
		}
	}
	dir.x = dir.x - 2.0f*cos1*normal.x;
	dir.y = dir.y - 2.0f*cos1*normal.y;
	return EV_REFLECTION;
}

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
	const xo::CylLayer *layers,
	const __grid_constant__ XoSource source,
	const __grid_constant__ xo::XoTrace trace,
	const __grid_constant__ XoFluence fluence,
	const __grid_constant__ xo::XoDetectors detectors,
	const float *fp_lut,
	xo::i32 *int_buffer,
	float *float_buffer,
	xo::u64 *accumulator_buffer,
	xo::u32 lut_len,
	xo::u32 priv_len,
	const __grid_constant__ xo::FluWindow window,
	xo::u32 chunk,
	xo::u32 refill)             // throughput mode: waiting lanes per warp that trigger a service round
{
	using namespace xo;
	extern __shared__ __align__(16) unsigned char xo_smem[];

	CylLayer *sh_layers = reinterpret_cast<CylLayer *>(xo_smem);
	u32 layer_words = num_layers*(u32)(sizeof(CylLayer)/4);
	{
		const u32 *src = reinterpret_cast<const u32 *>(layers);
		u32 *dst = reinterpret_cast<u32 *>(sh_layers);
		for (u32 i = threadIdx.x; i < layer_words; i += blockDim.x) dst[i] = src[i];
	}
	u32 off_words = (layer_words + 3u) & ~3u;
#if !XO_DETERMINISTIC
	CylFastLayer *sh_fast = reinterpret_cast<CylFastLayer *>(reinterpret_cast<u32 *>(xo_smem) + off_words);
	for (u32 i = threadIdx.x; i < num_layers; i += blockDim.x) {
		const CylLayer &Lg = layers[i];
		CylFastLayer F;
		F.hot.ri2 = Lg.r_inner*Lg.r_inner; F.hot.ro2 = Lg.r_outer*Lg.r_outer;
#if XO_ANISO
		F.hot.step_k = 0.0f; F.hot.step_b = 0.0f;       // per direction (XO_DIR_CONSTS)
		F.abs.absorb = 0.0f; F.abs.mua = 0.0f;
#else
		F.hot.step_k = -0.6931471805599453f*Lg.inv_mut;
		F.hot.step_b = -32.0f*F.hot.step_k;
		F.abs.absorb = Lg.mua_inv_mut; F.abs.mua = Lg.mua;
#endif
		F.abs.n = Lg.n; F.abs.pad = 0.0f;
		F.geo.ri = (Lg.r_inner > 0.0f) ? Lg.r_inner : -XO_INF; F.geo.ro = Lg.r_outer;
		F.geo.pad0 = 0.0f; F.geo.pad1 = 0.0f;
		Lg.pf.prepare(F.pf.v);
		sh_fast[i] = F;
	}
	off_words += num_layers*(u32)(sizeof(CylFastLayer)/4);
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
	acc.win = acc.priv + 2*priv_len;
	acc.bind();
	acc.lut = lut;
	for (u32 i = threadIdx.x; i < window.ext0*window.ext1*window.ext2; i += blockDim.x) acc.win[i] = 0;
	__syncthreads();

	const u32 gid = blockIdx.x*blockDim.x + threadIdx.x;
	Rng rng;
	rng.load(rng_state_x[gid]);
	rng.a = rng_state_a[gid];
	CylCtx ctx; ctx.layers = sh_layers; ctx.num_layers = (i32)num_layers;
	const P3 src_pos = source.origin();
	const float rmax2 = rmax*rmax;
	const XoTraceCfg &tcfg = *reinterpret_cast<const XoTraceCfg *>(&trace);
	(void)tcfg;

#if !XO_DETERMINISTIC
	// ======== throughput loop =====================================================
	// Lane states as in mcml_kernel.cuh: RUN / BND_IN, BND_OUT (the step ended on the
	// inner / outer cylinder of the layer, interface physics pending) / DEAD (needs a
	// packet) / DRY.  One VOTE per trip; a service round runs the interface physics
	// (Fresnel against the radial normal, detector deposit), the packet claims and
	// the launches jointly once `refill` lanes wait.
	enum : u32 { ST_RUN = 0, ST_DRY = 1, ST_BND_IN = 2, ST_BND_OUT = 3, ST_DEAD = 4 };
	bool started = false;
	u32 iterations = 0;
	{
		P3 pos = { 0.0f, 0.0f, 0.0f }, dir = { 0.0f, 0.0f, 1.0f };
		float weight = 0.0f;
		i32 layer = 1;
		float opl = 0.0f;
		u32 packet = 0, trace_count = 0, flags = 0;
		(void)opl; (void)packet; (void)trace_count; (void)flags;
		u32 state = ST_DEAD;
		u32 pk_next = 0, pk_end = 0;
		bool budget_dry = false;
		u32 thr_eff = refill < 1u ? 1u : (refill > 32u ? 32u : refill);
		CylHot c_hot = { 0.0f, 0.0f, 0.0f, 0.0f };
		CylAbs c_abs = { 0.0f, 0.0f, 1.0f, 0.0f };
		CylGeo c_geo = { 0.0f, 0.0f, 0.0f, 0.0f };
		XoPf::Fast c_pf;
#if XO_ANISO
	// anisotropic layers: step / absorption constants of a (layer, direction) pair
#define XO_DIR_CONSTS() do { \
		const CylLayer &L_ = sh_layers[layer]; \
		const float mut_ = tensor_project(L_.mut_t, dir), mua_ = tensor_project(L_.mua_t, dir); \
		const float inv_ = (mut_ != 0.0f) ? FastMath::rcp_approx(mut_) : XO_INF; \
		c_hot.step_k = -0.6931471805599453f*inv_; \
		c_hot.step_b = -32.0f*c_hot.step_k; \
		c_abs.absorb = (mua_ != 0.0f) ? mua_*inv_ : 0.0f; \
		c_abs.mua = mua_; \
	} while (0)
#else
#define XO_DIR_CONSTS() do { } while (0)
#endif
#define XO_CYL_LOAD_LAYER(idx) do { \
		const CylFastLayer &F_ = sh_fast[idx]; \
		c_hot = F_.hot; c_abs = F_.abs; c_geo = F_.geo; c_pf = F_.pf.v; \
		XO_DIR_CONSTS(); \
	} while (0)

#define XO_CYL_END_TRIP() do { \
		{ \
			float ex_ = pos.x - src_pos.x, ey_ = pos.y - src_pos.y, ez_ = pos.z - src_pos.z; \
			if (ex_*ex_ + ey_*ey_ + ez_*ez_ > rmax2 || weight <= 0.0f) { done = true; flags |= EV_ESCAPED; } \
		} \
		if (XO_TRACE) { \
			flags |= done ? EV_TERMINATED : 0u; \
			if (XO_TRACE == XO_TRACE_ALL || ((XO_TRACE & XO_TRACE_END) && done)) { \
				if (trace_event(tcfg, float_buffer, packet, trace_count, flags, \
						pos, dir, weight, opl)) ++trace_count; \
			} \
			if (done) trace_complete(tcfg, int_buffer, packet, trace_count); \
		} \
		flags = 0; \
		state = done ? ST_DEAD : ST_RUN; \
	} while (0)

		for (;;) {
			const u32 wait_mask = __ballot_sync(0xffffffffu, state >= ST_BND_IN);
			if (__builtin_expect((u32)__popc(wait_mask) >= thr_eff, 0)) {
				// ---- interface physics ---------------------------------------------
				if (state == ST_BND_IN || state == ST_BND_OUT) {
					const i32 next_layer = layer + (state == ST_BND_IN ? 1 : -1);
					bool done = false;
					u32 bf = cyl_boundary(sh_layers[layer], sh_layers[next_layer], pos, dir,
						layer, next_layer, rng);
					flags |= bf | EV_BOUNDARY_HIT;
					if (!(layer > 0 && layer < (i32)num_layers)) {
						if (layer <= 0 && XoDetOuter::active)
							detectors.outer.deposit(acc, pos, dir, weight, opl);
						done = true;
					} else {
						XO_CYL_LOAD_LAYER(layer);
					}
					{   // direction sanity check (mccyl.template.c:928-934)
						float len = M::sqrt(dir.x*dir.x + dir.y*dir.y + dir.z*dir.z);
						if (fabsf(len - 1.0f) > 10.0f*XO_FP_EPS) done = true;
					}
					XO_CYL_END_TRIP();
				}
				// ---- new packets ------------------------------------------------------
				if (state == ST_DEAD) {
					if (pk_next >= pk_end && !budget_dry) {
						pk_next = atomicAdd(num_packets_done, chunk);
						pk_end = (pk_next < num_packets && num_packets - pk_next > chunk) ? pk_next + chunk : num_packets;
						if (pk_next >= num_packets) { pk_end = pk_next; budget_dry = true; }
					}
					if (pk_next < pk_end) {
						Launch L_;
						packet = pk_next++;
						source.launch(rng, ctx, L_);
						pos = L_.pos; dir = L_.dir; weight = L_.weight; layer = L_.layer;
						if (XoDetSpecular::active)
							detectors.specular.deposit(acc, L_.pos, L_.spec_dir, L_.spec_weight, 0.0f);
						trace_count = 0;
						opl = 0.0f;
						flags = EV_LAUNCH;
						if (XO_TRACE & XO_TRACE_START) {
							if (trace_event(tcfg, float_buffer, packet, trace_count, flags,
									pos, dir, weight, opl)) ++trace_count;
						}
						if (layer > 0 && layer < (i32)num_layers) XO_CYL_LOAD_LAYER(layer);
						state = ST_RUN;
						started = true;
					} else {
						state = ST_DRY;
					}
				}
				const u32 n_dry = (u32)__popc(__ballot_sync(0xffffffffu, state == ST_DRY));
				if (n_dry == 32u) break;
				thr_eff = refill < 32u - n_dry ? refill : 32u - n_dry;
				if (thr_eff < 1u) thr_eff = 1u;
			}
			if (state != ST_RUN) continue;

			// ---- one step of the packet ------------------------------------------------
			++iterations;
			float step = fminf(fmaf(FastMath::lg2(rng.next_raw()), c_hot.step_k, c_hot.step_b), XO_FLT_MAX);
			bool hit = false, inwards = false;
			// Radial clearance: a packet at radius r cannot reach a surface of its layer
			// within a path shorter than min(ro - r, r - ri), whatever its direction - the
			// ray / cylinder quadratic (2 square roots, 1 reciprocal, ~25 FMA / compares)
			// is evaluated only for the steps that are longer than that (one square root
			// and 5 FMA / compares otherwise; in a 5 mm cylinder with a 0.1 mm free path
			// nearly every step).  A pure geometric bound: no statistical change.
			const float r_now = FastMath::sqrt(fmaf(pos.x, pos.x, pos.y*pos.y));
			const float room = fminf(c_geo.ro - r_now, r_now - c_geo.ri);
			if (!(step < 0.9999f*room) && (dir.x != 0.0f || dir.y != 0.0f)) {
				// distance to the inner / outer cylinder of the layer
				// (mccyl.template.c:147-209) from the half-b form of the quadratic:
				// a d^2 + 2 hb d + (c - r^2) = 0,  d = (-hb +- sqrt(hb^2 - a (c - r^2)))/a
				const float a = fmaf(dir.x, dir.x, dir.y*dir.y);
				const float hb = fmaf(pos.x, dir.x, pos.y*dir.y);
				const float c = fmaf(pos.x, pos.x, pos.y*pos.y);
				const float inv_a = FastMath::rcp_approx(a);
				float d_inner = XO_INF, d_outer = XO_INF;
				float D = fmaf(hb, hb, -a*(c - c_hot.ri2));
				if (c_hot.ri2 > 0.0f && D > 0.0f) {
					D = FastMath::sqrt(D);
					const float d1 = (-hb - D)*inv_a;
					const float d2 = (D - hb)*inv_a;
					d_inner = (d2 > 2.0f*XO_FP_EPS) ? fmaxf(d1, 0.0f) : XO_INF;
				}
				D = fmaf(hb, hb, -a*(c - c_hot.ro2));
				if (D >= 0.0f) {
					D = FastMath::sqrt(D);
					d_outer = fmaxf((D - hb)*inv_a, 0.0f);
				}
				const float d = fminf(d_outer, d_inner);
				hit = step > d;
				inwards = d_inner <= d_outer;
				step = fminf(d, step);
			}
			pos.x = fmaf(dir.x, step, pos.x);
			pos.y = fmaf(dir.y, step, pos.y);
			pos.z = fmaf(dir.z, step, pos.z);
			if (XO_NEEDS_OPL) opl = fmaf(c_abs.n, step, opl);
			if (__builtin_expect(hit, 0)) {
				state = inwards ? ST_BND_IN : ST_BND_OUT;
				continue;
			}
			bool done = false;
#if XO_METHOD == 1
			if (rng.next() < c_abs.absorb) {
				float deposit = weight;
				done = true;
				weight -= deposit;
				flags |= EV_ABSORPTION;
				if (XoFluence::active) fluence.deposit(acc, window, pos, deposit, c_abs.mua, opl);
			} else {
				pf_scatter(c_pf, rng, lut, dir);
				flags |= EV_SCATTERING;
				XO_DIR_CONSTS();
			}
#else
			{
				float deposit = weight*c_abs.absorb;
				weight -= deposit;
				flags |= EV_ABSORPTION;
				if (XoFluence::active) fluence.deposit(acc, window, pos, deposit, c_abs.mua, opl);
			}
			pf_scatter(c_pf, rng, lut, dir);
			flags |= EV_SCATTERING;
			XO_DIR_CONSTS();
			if (weight < XO_WEIGHT_MIN) {
#if XO_USE_LOTTERY
				if (rng.next_raw() > XO_LOTTERY_CHANCE*4294967296.0f) done = true;
				else weight *= (1.0f/XO_LOTTERY_CHANCE);
#else
				done = true;
#endif
			}
#endif
			XO_CYL_END_TRIP();
		}
#undef XO_CYL_END_TRIP
#undef XO_CYL_LOAD_LAYER
#undef XO_DIR_CONSTS
		rng_state_x[gid] = rng.state();
	}
#else
	// ======== deterministic loop: reference expressions, reference order =========
	(void)refill;
	u32 pk_next, pk_end;
#if XO_DETERMINISTIC
	static_quota(num_packets, gridDim.x*blockDim.x, gid, &pk_next, &pk_end);
	(void)chunk;
#else
	pk_next = atomicAdd(num_packets_done, chunk);
	pk_end = (pk_next < num_packets && num_packets - pk_next > chunk) ? pk_next + chunk : num_packets;
	if (pk_next >= num_packets) pk_end = pk_next;
#endif
	bool started = false;
	u32 iterations = 0;

	if (pk_next < pk_end) {
		started = true;
		P3 pos = { 0.0f, 0.0f, 0.0f }, dir = { 0.0f, 0.0f, 1.0f };
		float weight = 0.0f;
		i32 layer = 1;
		float opl = 0.0f;
		u32 packet = 0, trace_count = 0, flags = 0;
		bool done = false;
		(void)opl; (void)packet; (void)trace_count; (void)flags;

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
#if XO_DETERMINISTIC
		i32 num_steps = 0;
		while (!done && num_steps++ < XO_CYL_MAX_STEPS) {
#else
		while (!done) {
#endif
			const CylLayer &L = sh_layers[layer];
			++iterations;
			float step = -M::log(rng.next())*L.inv_mut_at(dir);
			step = fminf(step, XO_FLT_MAX);
			i32 next_layer = layer;
			if (dir.x != 0.0f || dir.y != 0.0f) {
				// distance to the inner / outer cylinder of the layer
				// (mccyl.template.c:147-209): a d^2 + b d + c - r^2 = 0
				float a = dir.x*dir.x + dir.y*dir.y;
				float b = 2.0f*(pos.x*dir.x + pos.y*dir.y);
				float c = pos.x*pos.x + pos.y*pos.y;
				float d_inner = XO_INF, d_outer = XO_INF;
				float inv_2a = M::div(1.0f, 2.0f*a);
				float D = b*b - 4.0f*a*(c - L.r_inner*L.r_inner);
				if (L.r_inner > 0.0f && D > 0.0f) {
					D = M::sqrt(D);
					float d1 = (-b - D)*inv_2a;
					float d2 = (-b + D)*inv_2a;
					d_inner = (d2 > 2.0f*XO_FP_EPS) ? fmaxf(d1, 0.0f) : XO_INF;
				}
				D = b*b - 4.0f*a*(c - L.r_outer*L.r_outer);
				if (D >= 0.0f) {
					D = M::sqrt(D);
					float d2 = (-b + D)*inv_2a;
					d_outer = fmaxf(d2, 0.0f);
				}
				float d = fminf(d_outer, d_inner);
				if (step > d) next_layer += (d_inner <= d_outer) ? 1 : -1;
				step = fminf(d, step);
			}
			pos.x = pos.x + dir.x*step;
			pos.y = pos.y + dir.y*step;
			pos.z = pos.z + dir.z*step;
			if (XO_NEEDS_OPL) opl += L.n*step;

			if (next_layer != layer) {
				u32 bf = cyl_boundary(L, sh_layers[next_layer], pos, dir, layer, next_layer, rng);
				flags |= bf | EV_BOUNDARY_HIT;
				if (!(layer > 0 && layer < (i32)num_layers)) {
					if (layer <= 0 && XoDetOuter::active)
						detectors.outer.deposit(acc, pos, dir, weight, opl);
					done = true;
				}
			} else {
#if XO_METHOD == 1
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
				{
					float deposit = weight*L.mua_inv_mut_at(dir);
					weight -= deposit;
					flags |= EV_ABSORPTION;
					if (XoFluence::active) fluence.deposit(acc, window, pos, deposit, L.mua_at(dir), opl);
				}
				pf_scatter(L.pf, rng, lut, dir);
				flags |= EV_SCATTERING;
				if (weight < XO_WEIGHT_MIN) {
#if XO_USE_LOTTERY
					if (rng.next() > XO_LOTTERY_CHANCE) done = true;
					else weight = M::div(weight, XO_LOTTERY_CHANCE);
#else
					done = true;
#endif
				}
#endif
			}
			// direction sanity check (mccyl.template.c:928-934); scattering
			// renormalises, so the throughput mode tests interface events only
#if !XO_DETERMINISTIC
			if (flags & EV_BOUNDARY_HIT)
#endif
			{
				float len = M::sqrt(dir.x*dir.x + dir.y*dir.y + dir.z*dir.z);
				if (fabsf(len - 1.0f) > 10.0f*XO_FP_EPS) done = true;
			}
			{
				float ex = pos.x - src_pos.x, ey = pos.y - src_pos.y, ez = pos.z - src_pos.z;
				if (ex*ex + ey*ey + ez*ez > rmax2 || weight <= 0.0f) { done = true; flags |= EV_ESCAPED; }
			}
#if XO_TRACE
			flags |= done ? EV_TERMINATED : 0u;
			if (XO_TRACE == XO_TRACE_ALL || ((XO_TRACE & XO_TRACE_END) && done)) {
				if (trace_event(tcfg, float_buffer, packet, trace_count, flags,
						pos, dir, weight, opl)) ++trace_count;
			}
#endif
			flags = 0;

			if (done) {
#if XO_TRACE
				trace_complete(tcfg, int_buffer, packet, trace_count);
#endif
#if !XO_DETERMINISTIC
				if (pk_next >= pk_end) {
					pk_next = atomicAdd(num_packets_done, chunk);
					pk_end = (pk_next < num_packets && num_packets - pk_next > chunk) ? pk_next + chunk : num_packets;
					if (pk_next >= num_packets) pk_end = pk_next;
				}
#endif
				if (pk_next < pk_end) {
					trace_count = 0;
					opl = 0.0f;
					XO_LAUNCH_PACKET();
					done = false;
				}
			}
		}
		rng_state_x[gid] = rng.state();
	}
#undef XO_LAUNCH_PACKET
#endif  // XO_DETERMINISTIC
	if (started) atomicAdd(num_kernels, 1u);
	{
		const u32 mask = __activemask();
		u32 warp_iters = __reduce_add_sync(mask, iterations);
		if ((threadIdx.x & 31u) == (u32)(__ffs(mask) - 1) && warp_iters)
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
