// mcml_kernel.cuh -- persistent-thread photon-packet kernel, planar layers.
//
// B200 counterpart of `McKernel` in xopto/mcml/kernel/mcml.template.c:346-824.
// Same physics, same per-work-item MWC stream usage (draw order: step ->
// [Fresnel] -> azimuth -> polar -> [lottery]; launch draws first), different
// machine mapping:
//   * one packet per thread, regenerated in place until the packet budget is
//     exhausted, packet state entirely in registers;
//   * layer table (+ pf lookup tables) staged once per CTA in shared memory;
//   * plugin parameter structs are __grid_constant__ kernel parameters;
//   * detector bins privatised per CTA in shared memory (Accu), fluence through
//     RED.E.ADD.64;
//   * packet scheduling: deterministic mode = static block schedule (work-item t
//     owns packets [base_t, base_t+n_t)), a legal outcome of the reference's
//     racing atomic counter (mcml.template.c:460,790); throughput mode = the
//     same counter, but claimed in chunks of `chunk` packets per atomic.
//
// The translation unit that includes this header must define the configuration:
//   typedef ... XoPf; XoSource; XoDetTop; XoDetBottom; XoDetSpecular; XoFluence;
//   XO_METHOD, XO_USE_LOTTERY, XO_WEIGHT_MIN, XO_LOTTERY_CHANCE, XO_TRACE,
//   XO_TRACK_OPL, XO_DETERMINISTIC, XO_BLOCK, XO_MIN_BLOCKS
#pragma once
#include "xo_core.cuh"
#include "xo_pf.cuh"
#include "xo_detectors.cuh"
#include "xo_fluence.cuh"
#include "mcml_sources.cuh"

namespace xo {

struct MlLayer {                    // mcml/mclayer/layer.py:57-69
	float thickness, top, bottom, n, cc_top, cc_bottom, mus, mua, inv_mut, mua_inv_mut;
	XoPf pf;
};

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

// Fresnel / Snell at a layer interface (mcml.template.c:80-203).  Returns the
// event flag and updates dir / layer index.  A uniform draw is consumed only
// when the indices differ and incidence is above the critical angle.
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
	xo::u32 chunk)              // packets claimed per atomic (throughput mode)
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
	const u32 nthreads = gridDim.x*blockDim.x;
	Rng rng;
	rng.x = rng_state_x[gid];
	rng.a = rng_state_a[gid];
	MlCtx ctx; ctx.layers = sh_layers;
	const P3 src_pos = source.origin();
	const float rmax2 = rmax*rmax;

	// ---- packet budget -------------------------------------------------------
	u32 pk_next, pk_end;
#if XO_DETERMINISTIC
	static_quota(num_packets, nthreads, gid, &pk_next, &pk_end);
	(void)chunk;
#else
	(void)nthreads;
	pk_next = atomicAdd(num_packets_done, chunk);
	pk_end = pk_next + chunk < num_packets ? pk_next + chunk : num_packets;
	if (pk_next >= num_packets) pk_end = pk_next;
#endif
	bool started = false;
	u32 iterations = 0;

	if (pk_next < pk_end) {
		started = true;
		// packet state (registers)
		P3 pos, dir;
		float weight;
		i32 layer;
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
			if (trace_event(*reinterpret_cast<const TraceCfg *>(&trace), float_buffer, packet, \
					trace_count, flags, pos, dir, weight, opl)) ++trace_count; \
		} \
	} while (0)

		XO_LAUNCH_PACKET();

		while (!done) {
			const MlLayer &L = sh_layers[layer];
			float step;
			++iterations;
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
				u32 bf = ml_boundary(L, sh_layers[next_layer], dir, layer, next_layer, rng);
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
			{
				float ex = pos.x - src_pos.x, ey = pos.y - src_pos.y, ez = pos.z - src_pos.z;
				if (ex*ex + ey*ey + ez*ez > rmax2) { done = true; flags |= EV_ESCAPED; }
			}
#if XO_TRACE
			flags |= done ? EV_TERMINATED : 0u;
			if (XO_TRACE == XO_TRACE_ALL || ((XO_TRACE & XO_TRACE_END) && done)) {
				if (trace_event(*reinterpret_cast<const TraceCfg *>(&trace), float_buffer, packet,
						trace_count, flags, pos, dir, weight, opl)) ++trace_count;
			}
#endif
			flags = 0;

			if (done) {
#if XO_TRACE
				int_buffer[reinterpret_cast<const TraceCfg *>(&trace)->count_off + packet] = (i32)trace_count;
#endif
#if !XO_DETERMINISTIC
				if (pk_next >= pk_end) {
					pk_next = atomicAdd(num_packets_done, chunk);
					pk_end = pk_next + chunk < num_packets ? pk_next + chunk : num_packets;
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
		rng_state_x[gid] = rng.x;
	}
#undef XO_LAUNCH_PACKET
	if (started) atomicAdd(num_kernels, 1u);
	// loop-iteration count (the roofline's unit of work): one 64-bit RED per warp
	{
		const u32 mask = __activemask();
		u32 warp_iters = __reduce_add_sync(mask, iterations);
		if ((threadIdx.x & 31u) == (u32)(__ffs(mask) - 1) && warp_iters)
			atomicAdd(reinterpret_cast<u64 *>(num_kernels + 1), (u64)warp_iters);
	}

	__syncthreads();
	acc.flush_private();
#if XO_DETERMINISTIC
	if (gid == 0) *num_packets_done = num_packets;
#endif
	(void)int_buffer; (void)float_buffer;
}
