// mcvox_kernel.cuh -- persistent-thread photon-packet kernel, voxel grid.
//
// B200 counterpart of `McKernel` in xopto/mcvox/kernel/mcvox.template.c:548-1048.
// Every loop iteration draws a fresh exponential step, measures the distance to
// the three exit faces of the current voxel, moves min(distance, step) and then
// either handles the voxel boundary (same-material fast path, else Fresnel) or
// absorbs + scatters.  Machine mapping as in mcml_kernel.cuh; additionally
//   * the material table lives in shared memory, the voxel box in the constant
//     bank (kernel parameter), the voxel -> material map (int32 [nz][ny][nx]) is
//     read through the read-only path (L2 resident: 201^3 = 32 MB of 126 MB),
//   * deposition goes to the 64-bit grid with RED.E.ADD.64.
#pragma once
#include "xo_core.cuh"
#include "xo_pf.cuh"
#include "xo_detectors.cuh"
#include "xo_fluence.cuh"
#include "mcvox_sources.cuh"

#include "mcvox_medium.cuh"

namespace xo {

typedef Detectors<XoDetTop, XoDetBottom, XoDetSpecular> XoDetectors;

#define XO_NEEDS_OPL (XO_TRACK_OPL || XoDetTop::needs_opl || XoDetBottom::needs_opl || \
	XoDetSpecular::needs_opl || XoFluence::needs_opl)

// Throughput mode (albedo weight / albedo rejection) runs the DDA formulation
// below; deterministic mode and microscopic Beer-Lambert keep the reference's
// iteration structure (one fresh step + three face divisions per iteration).
#ifndef XO_VOX_PACKED
#define XO_VOX_PACKED 0
#endif
#define XO_VOX_DDA (!XO_DETERMINISTIC && XO_METHOD != 2 && XO_VOX_PACKED)
#ifndef XO_VOX_POOL
#define XO_VOX_POOL 0               // slots of the per-warp packet pool (0: lane-resident packets)
#endif
#define XO_VOX_SENTINEL 255

// Per-material record of the throughput loop, derived once per CTA when the
// material table is staged in shared memory: exactly the register cache of the
// current material (two LDS.128 on a material change).
struct __align__(16) VoxHot { float step_k, absorb, mua, n; };
	// step_k = -ln2/mut: step = lg2(u)*step_k
struct __align__(16) VoxPfFast { XoPf::Fast v; };
struct VoxFastMat { VoxHot hot; VoxPfFast pf; };

struct VoxCtx {
	const VoxCfg &cfg;
	const VoxMaterial *materials;   // shared memory
	const i32 *voxels;              // global, read-only
	const float *lut = nullptr;     // float lookup-table pool (IsotropicVoxels)
	static constexpr bool has_specular = XoDetSpecular::active;
	__device__ __forceinline__ VoxCtx(const VoxCfg &c, const VoxMaterial *m, const i32 *v)
		: cfg(c), materials(m), voxels(v) {}
	__device__ __forceinline__ float material_n(i32 i) const { return materials[i].n; }
	__device__ __forceinline__ i32 voxel_material(i32 x, i32 y, i32 z) const {
		return __ldg(voxels + ((i64)z*cfg.ny + y)*cfg.nx + x);
	}
	__device__ __forceinline__ bool valid(i32 x, i32 y, i32 z) const {
		return x >= 0 && x < cfg.nx && y >= 0 && y < cfg.ny && z >= 0 && z < cfg.nz;
	}
	__device__ __forceinline__ bool box_contains(const P3 &p) const {
		return p.x >= cfg.top_left.x && p.x < cfg.bottom_right.x &&
			p.y >= cfg.top_left.y && p.y < cfg.bottom_right.y &&
			p.z >= cfg.top_left.z && p.z < cfg.bottom_right.z;
	}
	// truncating conversion as in mcvox.template.c:206-220
	__device__ __forceinline__ void position_to_voxel(const P3 &p, i32 *x, i32 *y, i32 *z) const {
		*x = f2i(M::div(p.x - cfg.top_left.x, cfg.size.x));
		*y = f2i(M::div(p.y - cfg.top_left.y, cfg.size.y));
		*z = f2i(M::div(p.z - cfg.top_left.z, cfg.size.z));
	}
	// slab test against the voxel box (mcvox.template.c:94-162)
	__device__ inline bool box_intersect(const P3 &pos, const P3 &dir, P3 *isect, P3 *normal) const {
		float ix = (dir.x != 0.0f) ? M::div(1.0f, dir.x) : XO_INF;
		float iy = (dir.y != 0.0f) ? M::div(1.0f, dir.y) : XO_INF;
		float iz = (dir.z != 0.0f) ? M::div(1.0f, dir.z) : XO_INF;
		float t1 = (cfg.top_left.x - pos.x)*ix, t2 = (cfg.bottom_right.x - pos.x)*ix;
		float t3 = (cfg.top_left.y - pos.y)*iy, t4 = (cfg.bottom_right.y - pos.y)*iy;
		float t5 = (cfg.top_left.z - pos.z)*iz, t6 = (cfg.bottom_right.z - pos.z)*iz;
		float txmin = fminf(t1, t2), txmax = fmaxf(t1, t2);
		float tymin = fminf(t3, t4), tymax = fmaxf(t3, t4);
		float tzmin = fminf(t5, t6), tzmax = fmaxf(t5, t6);
		float tmin = fmaxf(fmaxf(txmin, tymin), tzmin);
		float tmax = fminf(fminf(txmax, tymax), tzmax);
		if (tmin > tmax) return false;
		float nx = (txmin >= tymin && txmin >= tzmin) ? 1.0f : 0.0f;
		float ny = (nx == 0.0f && tymin >= tzmin) ? 1.0f : 0.0f;
		float nz = (nx == 0.0f && ny == 0.0f) ? 1.0f : 0.0f;
		normal->x = nx*signf(dir.x);
		normal->y = ny*signf(dir.y);
		normal->z = nz*signf(dir.z);
		float t = (tmin > 0.0f) ? tmin : tmax;
		isect->x = clipf(pos.x + t*dir.x, cfg.top_left.x, cfg.bottom_right.x - XO_FP_EPS);
		isect->y = clipf(pos.y + t*dir.y, cfg.top_left.y, cfg.bottom_right.y - XO_FP_EPS);
		isect->z = clipf(pos.z + t*dir.z, cfg.top_left.z, cfg.bottom_right.z - XO_FP_EPS);
		return tmax >= tmin && tmax >= 0.0f;
	}
};

}  // namespace xo

extern "C" __global__ void __launch_bounds__(XO_BLOCK, XO_MIN_BLOCKS)
McKernel(
	xo::u32 num_packets,
	xo::u32 *num_packets_done,
	xo::u32 *num_kernels,
	float rmax,
	xo::u64 *rng_state_x,
	const xo::u32 *rng_state_a,
	const __grid_constant__ xo::VoxCfg voxel_cfg,
	const xo::i32 *voxels,
	xo::u32 num_materials,
	const xo::VoxMaterial *materials,
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
	xo::u32 refill,             // throughput mode: waiting lanes per warp that trigger their joint handling
	const unsigned char *voxels8,   // throughput mode: padded compact map, 16-bit cells material | clearance << 8 (mcvox/mc.py)
	xo::u32 vox_bx,             // ... its index = (x+2) | (y+2) << vox_bx | (z+2) << (vox_bx + vox_by)
	xo::u32 vox_by)
{
	using namespace xo;
	extern __shared__ __align__(16) unsigned char xo_smem[];

	VoxMaterial *sh_mat = reinterpret_cast<VoxMaterial *>(xo_smem);
	u32 mat_words = num_materials*(u32)(sizeof(VoxMaterial)/4);
	{
		const u32 *src = reinterpret_cast<const u32 *>(materials);
		u32 *dst = reinterpret_cast<u32 *>(sh_mat);
		for (u32 i = threadIdx.x; i < mat_words; i += blockDim.x) dst[i] = src[i];
	}
	u32 off_words = (mat_words + 3u) & ~3u;
#if XO_VOX_DDA
	VoxFastMat *sh_fast = reinterpret_cast<VoxFastMat *>(reinterpret_cast<u32 *>(xo_smem) + off_words);
	for (u32 i = threadIdx.x; i < num_materials; i += blockDim.x) {
		const VoxMaterial &Mg = materials[i];
		VoxFastMat F;
#if XO_ANISO
		F.hot.step_k = 0.0f; F.hot.absorb = 0.0f; F.hot.mua = 0.0f;   // per ray (XO_DIR_CONSTS)
#else
		F.hot.step_k = -0.6931471805599453f*Mg.inv_mut;
		F.hot.absorb = Mg.mua_inv_mut;
		F.hot.mua = Mg.mua;
#endif
		F.hot.n = Mg.n;
		Mg.pf.prepare(F.pf.v);
		sh_fast[i] = F;
	}
	off_words += num_materials*(u32)(sizeof(VoxFastMat)/4);
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
	const u32 win_len = window.ext0*window.ext1*window.ext2;
	for (u32 i = threadIdx.x; i < win_len; i += blockDim.x) acc.win[i] = 0;
#if XO_VOX_DDA
	// per-warp launch queue: 32 slots of {pos, weight | dir, packet | trace count}
	off_words += 2*priv_len + win_len;
	off_words = (off_words + 3u) & ~3u;
#if XO_VOX_POOL
	// per-warp packet pool (mcvox_pool_loop.cuh), field by field: XO_VOX_POOL slots of
	// 4 x float4 (A, B, C, D), then the optical path length (float) - with a trace a fifth
	// quad T instead -, where the rmax sphere can be reached the ray parameter of its exit
	// (float), one state byte per slot, 32 bytes of gather indices
	const u32 pool_warps = blockDim.x >> 5, pool_warp = threadIdx.x >> 5;
	unsigned char *pool_next = reinterpret_cast<unsigned char *>(reinterpret_cast<u32 *>(xo_smem) + off_words);
#ifndef XO_POOL_SOA
#define XO_POOL_SOA 1
#endif
#define XO_POOL_FIELDS (16u + (XO_TRACE ? 4u : 1u) + (XO_USE_RMAX ? 1u : 0u))
#if XO_POOL_SOA
	// one array of XO_VOX_POOL floats per component and warp (mcvox_pool_loop.cuh)
	float *P_F = reinterpret_cast<float *>(pool_next) + pool_warp*(XO_POOL_FIELDS*XO_VOX_POOL);
	pool_next += pool_warps*XO_POOL_FIELDS*XO_VOX_POOL*4u;
#else
	float4 *P_A = reinterpret_cast<float4 *>(pool_next) + pool_warp*XO_VOX_POOL; pool_next += pool_warps*XO_VOX_POOL*16u;
	float4 *P_B = reinterpret_cast<float4 *>(pool_next) + pool_warp*XO_VOX_POOL; pool_next += pool_warps*XO_VOX_POOL*16u;
	float4 *P_C = reinterpret_cast<float4 *>(pool_next) + pool_warp*XO_VOX_POOL; pool_next += pool_warps*XO_VOX_POOL*16u;
	float4 *P_D = reinterpret_cast<float4 *>(pool_next) + pool_warp*XO_VOX_POOL; pool_next += pool_warps*XO_VOX_POOL*16u;
#if XO_TRACE
	float4 *P_T = reinterpret_cast<float4 *>(pool_next) + pool_warp*XO_VOX_POOL; pool_next += pool_warps*XO_VOX_POOL*16u;
#else
	float *P_E = reinterpret_cast<float *>(pool_next) + pool_warp*XO_VOX_POOL; pool_next += pool_warps*XO_VOX_POOL*4u;
#endif
#if XO_USE_RMAX
	float *P_R = reinterpret_cast<float *>(pool_next) + pool_warp*XO_VOX_POOL; pool_next += pool_warps*XO_VOX_POOL*4u;
#endif
#endif
	unsigned char *P_ST = pool_next + pool_warp*XO_VOX_POOL; pool_next += pool_warps*XO_VOX_POOL;
	unsigned char *P_IDX = pool_next + pool_warp*32u; pool_next += pool_warps*32u;
	// (XO_POOL_QUEUES: five rings of 64 slot numbers per warp)
	unsigned char *P_Q = pool_next + pool_warp*320u;
	(void)P_IDX; (void)P_Q;
#else
	float4 *q_a = reinterpret_cast<float4 *>(reinterpret_cast<u32 *>(xo_smem) + off_words) + (threadIdx.x & ~31u)*2u;
	float4 *q_b = q_a + 32;
	u32 *q_l = reinterpret_cast<u32 *>(xo_smem) + off_words + blockDim.x*8u + (threadIdx.x & ~31u);
	u32 *q_v = q_l + blockDim.x;    // voxel (low address word of the compact map) of the launch point
#endif
#endif
	__syncthreads();

	const u32 gid = blockIdx.x*blockDim.x + threadIdx.x;
	const u32 nthreads = gridDim.x*blockDim.x;
	Rng rng;
	rng.load(rng_state_x[gid]);
	rng.a = rng_state_a[gid];
	const VoxCfg &cfg = voxel_cfg;
	VoxCtx ctx(cfg, sh_mat, voxels);
	ctx.lut = lut;
	const P3 src_pos = source.origin();
	const float rmax2 = rmax*rmax;

	bool started = false;
	u32 iterations = 0;
#if XO_VOX_DDA && XO_VOX_POOL
#include "mcvox_pool_loop.cuh"
#elif XO_VOX_DDA
#include "mcvox_dda_loop.cuh"
#else
	(void)refill; (void)voxels8; (void)vox_bx; (void)vox_by;
	u32 pk_next, pk_end;
#if XO_DETERMINISTIC
	static_quota(num_packets, nthreads, gid, &pk_next, &pk_end);
	(void)chunk;
#else
	(void)nthreads;
	pk_next = atomicAdd(num_packets_done, chunk);
	pk_end = (pk_next < num_packets && num_packets - pk_next > chunk) ? pk_next + chunk : num_packets;
	if (pk_next >= num_packets) pk_end = pk_next;
#endif

	if (pk_next < pk_end) {
		started = true;
		P3 pos = { 0.0f, 0.0f, 0.0f }, dir;
		float weight;
		i32 vx, vy, vz, mat;
		float opl = 0.0f;
		u32 packet = 0, trace_count = 0, flags = 0;
		bool done = false;
		(void)opl; (void)packet; (void)trace_count; (void)flags;

#define XO_LAUNCH_PACKET() do { \
		Launch L_; \
		packet = pk_next++; \
		source.launch(rng, ctx, pos, L_); \
		pos = L_.pos; dir = L_.dir; weight = L_.weight; \
		if (XoDetSpecular::active && L_.spec_weight >= 0.0f) \
			detectors.specular.deposit(acc, L_.pos, L_.spec_dir, L_.spec_weight, 0.0f); \
		ctx.position_to_voxel(pos, &vx, &vy, &vz); \
		mat = ctx.voxel_material(vx, vy, vz); \
		flags |= EV_LAUNCH; \
		if (XO_TRACE & XO_TRACE_START) { \
			if (trace_event(*reinterpret_cast<const XoTraceCfg *>(&trace), float_buffer, packet, \
					trace_count, flags, pos, dir, weight, opl)) ++trace_count; \
		} \
	} while (0)

		XO_LAUNCH_PACKET();

		while (!done) {
			const VoxMaterial &Mt = sh_mat[mat];
			++iterations;
			float step;
#if XO_METHOD == 2
			step = M::div(-M::log(rng.next()), Mt.mus_at(dir));
#else
			step = -M::log(rng.next())*Mt.inv_mut_at(dir);
#endif
			step = fminf(step, XO_FLT_MAX);
			// distances to the exit faces of the current voxel (mcvox.template.c:173-196)
			float dx = cfg.top_left.x + (float)((dir.x >= 0.0f ? 1 : 0) + vx)*cfg.size.x - pos.x;
			float dy = cfg.top_left.y + (float)((dir.y >= 0.0f ? 1 : 0) + vy)*cfg.size.y - pos.y;
			float dz = cfg.top_left.z + (float)((dir.z >= 0.0f ? 1 : 0) + vz)*cfg.size.z - pos.z;
			dx = (dir.x != 0.0f) ? M::div(dx, dir.x) : XO_INF;
			dy = (dir.y != 0.0f) ? M::div(dy, dir.y) : XO_INF;
			dz = (dir.z != 0.0f) ? M::div(dz, dir.z) : XO_INF;
			float d = fminf(dx, fminf(dy, dz));
			float d_ok = fminf(d, step);
			pos.x = pos.x + d_ok*dir.x;
			pos.y = pos.y + d_ok*dir.y;
			pos.z = pos.z + d_ok*dir.z;
			if (XO_NEEDS_OPL) opl += Mt.n*d_ok;

#if XO_METHOD == 2
			{
				float mua = Mt.mua_at(dir);
				float frac = 1.0f - M::exp(-mua*d_ok);
				float deposit = frac*weight;
				weight -= deposit;
				flags |= EV_ABSORPTION;
				if (XoFluence::active) {
					float back = (mua != 0.0f) ?
						d_ok - M::div(-M::log(1.0f - rng.next()*frac), mua) : 0.0f;
					P3 dp = { pos.x - back*dir.x, pos.y - back*dir.y, pos.z - back*dir.z };
					fluence.deposit(acc, window, dp, deposit, mua, opl);
				}
			}
#endif
			if (d < step) {
				// ---- voxel boundary (mcvox.template.c:275-388) ----
				i32 nx = (dx <= dy && dx <= dz) ? 1 : 0;
				i32 ny = nx ? 0 : ((dy <= dx && dy <= dz) ? 1 : 0);
				i32 nz = !(nx + ny);
				nx = dir.x < 0.0f ? -nx : nx;
				ny = dir.y < 0.0f ? -ny : ny;
				nz = dir.z < 0.0f ? -nz : nz;
				i32 qx = vx + nx, qy = vy + ny, qz = vz + nz;
				bool escaping = !ctx.valid(qx, qy, qz);
				i32 next_mat = escaping ? 0 : ctx.voxel_material(qx, qy, qz);
				u32 bf;
				if (next_mat == mat && !escaping) {
					vx = qx; vy = qy; vz = qz;
					bf = EV_REFRACTION;
				} else {
					float n1 = Mt.n, n2 = sh_mat[next_mat].n;
					if (n1 == n2) {
						vx = qx; vy = qy; vz = qz; mat = next_mat;
						bf = EV_REFRACTION;
					} else {
						float cc = xo::cos_critical(n1, n2);
						float cos1 = (float)nx*dir.x + (float)ny*dir.y + (float)nz*dir.z;
						P3 fn = { (float)nx, (float)ny, (float)nz };
						bf = EV_REFLECTION;
						bool refracted = false;
						if (cos1 > cc) {
							float R = xo::reflectance(n1, n2, cos1, cc);
							if (R < rng.next()) {
								dir = refract3(dir, fn, n1, n2);
								vx = qx; vy = qy; vz = qz; mat = next_mat;
								bf = EV_REFRACTION;
								refracted = true;
							}
						}
						if (!refracted) dir = reflect3(dir, fn);
					}
				}
				flags |= bf | EV_BOUNDARY_HIT;
				if (!ctx.valid(vx, vy, vz)) {
					if (vz < 0) {
						if (XoDetTop::active) detectors.top.deposit(acc, pos, dir, weight, opl);
					} else if (vz >= cfg.nz) {
						if (XoDetBottom::active) detectors.bottom.deposit(acc, pos, dir, weight, opl);
					}
					done = true;
				}
			} else {
#if XO_METHOD == 1
				if (rng.next() < Mt.mua_inv_mut_at(dir)) {
					float deposit = weight;
					weight -= deposit;
					flags |= EV_ABSORPTION;
					done = true;
					if (XoFluence::active) fluence.deposit(acc, window, pos, deposit, Mt.mua_at(dir), opl);
				} else {
					pf_scatter(Mt.pf, rng, lut, dir);
					flags |= EV_SCATTERING;
				}
#else
#if XO_METHOD == 0
				{
					float deposit = weight*Mt.mua_inv_mut_at(dir);
					weight -= deposit;
					flags |= EV_ABSORPTION;
					if (XoFluence::active) fluence.deposit(acc, window, pos, deposit, Mt.mua_at(dir), opl);
				}
#endif
				pf_scatter(Mt.pf, rng, lut, dir);
				flags |= EV_SCATTERING;
#endif
			}
#if XO_METHOD != 1
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
				if (ex*ex + ey*ey + ez*ez > rmax2 || weight <= 0.0f) { done = true; flags |= EV_ESCAPED; }
			}
#if XO_TRACE
			flags |= done ? EV_TERMINATED : 0u;
			if (XO_TRACE == XO_TRACE_ALL || ((XO_TRACE & XO_TRACE_END) && done)) {
				if (trace_event(*reinterpret_cast<const XoTraceCfg *>(&trace), float_buffer, packet,
						trace_count, flags, pos, dir, weight, opl)) ++trace_count;
			}
#endif
			flags = 0;

			if (done) {
#if XO_TRACE
				trace_complete(*reinterpret_cast<const XoTraceCfg *>(&trace), int_buffer, packet, trace_count);
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
#endif  // XO_VOX_DDA
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
