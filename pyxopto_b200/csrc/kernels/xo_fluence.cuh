// xo_fluence.cuh -- fluence / deposition accumulators and the packet trace.
//
// Struct members = packed `McFluence` / `McTrace` of the reference plugins
// (xopto/mcbase/mcfluence/{fluence,fluencerz,fluencet}.py, mcbase/mctrace.py
// `cl_type`).  Deposits go straight to the 64-bit global grid with RED.E.ADD.64;
// the voxel index arithmetic restates `mcsim_fluence_deposit_at`.
#pragma once
#include "xo_core.cuh"

namespace xo {

struct FluNone {
	i32 dummy;
	static constexpr bool active = false;
	static constexpr bool needs_opl = false;
	static constexpr bool fixed_point = true;
	__device__ __forceinline__ void deposit(const Accu &, const FluWindow &, const P3 &, float, float, float) const {}
	__device__ __forceinline__ void deposit_fixed(const Accu &, const FluWindow &, const P3 &, u32, float) const {}
	__device__ __forceinline__ float fixed_scale(float) const { return 0.0f; }
	struct Prep { };
	struct Far { };
	__device__ __forceinline__ Prep prepare(const FluWindow &) const { return Prep(); }
	__device__ __forceinline__ Far prepare_far(const FluWindow &) const { return Far(); }
	__device__ __forceinline__ void deposit_prep(const Accu &, const Prep &, const FluWindow &, const P3 &, u32, float) const {}
	__device__ __forceinline__ u32 window_index(const FluWindow &, u32) const { return 0; }
};

// adds the CTA-private window back into the global grid (end of the kernel)
template <class Flu>
__device__ __forceinline__ void flush_window(const Flu &flu, const Accu &acc, const FluWindow &win) {
	const u32 n = win.ext0*win.ext1*win.ext2;
	for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
		u32 v = acc.win[i];
		if (v) atomicAdd(acc.global + flu.window_index(win, i), (u64)v);
	}
}

#ifndef XO_FLUENCE_RATE
#define XO_FLUENCE_RATE 0
#endif
// 0: the host passes an empty FluWindow (every deposit goes to the global grid)
#ifndef XO_FLU_WINDOW
#define XO_FLU_WINDOW 1
#endif

__device__ __forceinline__ u32 fluence_weight(float w, float mua, i32 k) {
#if XO_FLUENCE_RATE
	w *= (mua != 0.0f) ? M::div(1.0f, mua) : 0.0f;
#else
	(void)mua;
#endif
#if XO_DETERMINISTIC
	return f2u(__fadd_rn(__fmul_rn(w, (float)k), 0.5f));
#else
	return f2u(fmaf(w, (float)k, 0.5f));
#endif
}

// weight -> fixed point factor of a layer / material (throughput loops):
// fluence_weight(w, mua, k) == f2u(fmaf(w, fluence_scale(mua, k), 0.5f)) up to the
// rounding of the product
__device__ __forceinline__ float fluence_scale(float mua, i32 k) {
#if XO_FLUENCE_RATE
	return (mua != 0.0f) ? (float)k/mua : 0.0f;
#else
	(void)mua;
	return (float)k;
#endif
}

struct FluXyz {                     // mcfluence/fluence.py:57-63
	P3 inv_step, top_left; u32 nx, ny, nz, offset; i32 k;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	static constexpr bool fixed_point = true;
	// The reference tests 0 <= f < n in floating point and then truncates
	// (fluence.py:103-143).  floor-conversion + one unsigned compare per axis is
	// the same predicate for every non-NaN f: negative f floors to a negative
	// integer (huge as unsigned), f >= n converts to >= n or saturates.
	struct Prep { };
	struct Far { };
	__device__ __forceinline__ Prep prepare(const FluWindow &) const { return Prep(); }
	__device__ __forceinline__ Far prepare_far(const FluWindow &) const { return Far(); }
	__device__ __forceinline__ void deposit_prep(const Accu &acc, const Prep &, const FluWindow &win, const P3 &pos, u32 wfix, float opl) const {
		deposit_fixed(acc, win, pos, wfix, opl);
	}
	// deposit of a weight that is already in fixed point (throughput loops fold
	// the conversion constant into their per-layer constants, see fixed_scale)
	__device__ __forceinline__ void deposit(const Accu &acc, const FluWindow &win, const P3 &pos, float w, float mua, float opl) const {
		deposit_fixed(acc, win, pos, fluence_weight(w, mua, k), opl);
	}
	__device__ __forceinline__ float fixed_scale(float mua) const { return fluence_scale(mua, k); }
	__device__ __forceinline__ void deposit_fixed(const Accu &acc, const FluWindow &win, const P3 &pos, u32 wfix, float) const {
		u32 ix = (u32)__float2int_rd((pos.x - top_left.x)*inv_step.x);
		u32 iy = (u32)__float2int_rd((pos.y - top_left.y)*inv_step.y);
		u32 iz = (u32)__float2int_rd((pos.z - top_left.z)*inv_step.z);
		u32 lx = ix - win.org0, ly = iy - win.org1, lz = iz - win.org2;
		// the window lies inside the grid: a hit there needs no grid bounds test
		if (XO_FLU_WINDOW && lx < win.ext0 && ly < win.ext1 && lz < win.ext2) {
			if (acc.add_window((lz*win.ext1 + ly)*win.ext0 + lx, wfix))
				acc.carry_global(offset + (iz*ny + iy)*nx + ix);
		} else if (ix < nx && iy < ny && iz < nz) {
			acc.add_global(offset + (iz*ny + iy)*nx + ix, wfix);
		}
	}
	__device__ __forceinline__ u32 window_index(const FluWindow &win, u32 local) const {
		u32 lx = local % win.ext0, t = local/win.ext0;
		u32 ly = t % win.ext1, lz = t/win.ext1;
		return offset + ((lz + win.org2)*ny + (ly + win.org1))*nx + lx + win.org0;
	}
	// deposit into a cell whose (in-range) indices the caller already knows (mcvox
	// throughput loop with the fluence grid on the voxel grid)
	__device__ __forceinline__ void deposit_cell(const Accu &acc, u32 ix, u32 iy, u32 iz, u32 wfix) const {
		acc.add_global(offset + (iz*ny + iy)*nx + ix, wfix);
	}
};

struct FluRz {                      // mcfluence/fluencerz.py:64-72
	P3 center; float inv_dr, inv_dz; u32 n_r, n_z, offset; i32 k;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	static constexpr bool fixed_point = true;
	// bounds test in the integer domain, see FluXyz::deposit (fluencerz.py:112-160)
	// Throughput loops: the grid constants in the form the deposit uses them, held
	// in registers for the whole kernel (the kernel passes them through shared
	// memory once; read from the kernel parameters the loop reloads five uniform
	// registers from the constant bank per deposit).  The window origin is folded
	// into the index offsets: l = floor(x*inv_d + b) is the window-local index,
	// floor(f - n) == floor(f) - n for the integer n.
	struct Prep { float cx, cy, inv_dr, br, inv_dz, bz; u32 ext0, ext1; };
	// constants of the (rare, divergent) deposit outside the window, read from
	// shared memory with one LDS.128 (16 bytes in front of the window): grid cells left of / below the window
	// origin, row length, flat index of the window origin
	struct __align__(16) Far { u32 rem0, rem1, n_r, base; };
	__device__ __forceinline__ Prep prepare(const FluWindow &win) const {
		Prep p;
		p.cx = center.x; p.cy = center.y; p.inv_dr = inv_dr; p.inv_dz = inv_dz;
		p.br = XO_FLOOR_MAGIC - (float)win.org0;      // exact: both are integers
		p.bz = -center.z*inv_dz - (float)win.org1;
		p.ext0 = win.ext0; p.ext1 = win.ext1;
		return p;
	}
	__device__ __forceinline__ Far prepare_far(const FluWindow &win) const {
		Far f;
		f.rem0 = n_r - win.org0; f.rem1 = n_z - win.org1;
		f.n_r = n_r; f.base = offset + win.org1*n_r + win.org0;
		return f;
	}
	__device__ __forceinline__ void deposit_prep(const Accu &acc, const Prep &p, const FluWindow &win, const P3 &pos, u32 wfix, float opl) const {
		if (!XO_FLU_WINDOW) { deposit_fixed(acc, win, pos, wfix, opl); return; }
		float dx = pos.x - p.cx, dy = pos.y - p.cy;
		float r = FastMath::sqrt(fmaf(dx, dx, dy*dy));
		// floor without the conversion unit (F2I shares the XU pipe with MUFU, the
		// busiest pipe of this loop): adding 1.5 * 2^23 with round-down leaves
		// floor(x) in the low mantissa bits for |x| < 2^22, and the bit pattern is
		// monotonic in x beyond that, so the unsigned window test still rejects
		// every out-of-range / NaN coordinate (p.br carries the 1.5 * 2^23)
		u32 lr = __float_as_uint(__fmaf_rd(r, p.inv_dr, p.br)) - XO_FLOOR_MAGIC_BITS;
		u32 lz = __float_as_uint(__fadd_rd(fmaf(pos.z, p.inv_dz, p.bz), XO_FLOOR_MAGIC)) - XO_FLOOR_MAGIC_BITS;
		if (lr < p.ext0 && lz < p.ext1) {
			if (acc.add_window(lz*p.ext0 + lr, wfix))
				acc.carry_global(offset + (lz + win.org1)*n_r + lr + win.org0);
		} else {
			const uint4 q = acc.load_far();
			Far f; f.rem0 = q.x; f.rem1 = q.y; f.n_r = q.z; f.base = q.w;
			if (lr < f.rem0 && lz < f.rem1) acc.add_global(lz*f.n_r + lr + f.base, wfix);
		}
	}
	// deposit of a weight that is already in fixed point (throughput loops fold
	// the conversion constant into their per-layer constants, see fixed_scale)
	__device__ __forceinline__ void deposit(const Accu &acc, const FluWindow &win, const P3 &pos, float w, float mua, float opl) const {
		deposit_fixed(acc, win, pos, fluence_weight(w, mua, k), opl);
	}
	__device__ __forceinline__ float fixed_scale(float mua) const { return fluence_scale(mua, k); }
	__device__ __forceinline__ void deposit_fixed(const Accu &acc, const FluWindow &win, const P3 &pos, u32 wfix, float) const {
		float dx = pos.x - center.x, dy = pos.y - center.y;
		float r = M::sqrt(dx*dx + dy*dy);
		float dz = pos.z - center.z;
		u32 ir = (u32)__float2int_rd(r*inv_dr), iz = (u32)__float2int_rd(dz*inv_dz);
		u32 lr = ir - win.org0, lz = iz - win.org1;
		// the window lies inside the grid: a hit there needs no grid bounds test
		if (XO_FLU_WINDOW && lr < win.ext0 && lz < win.ext1) {
			if (acc.add_window(lz*win.ext0 + lr, wfix))
				acc.carry_global(offset + iz*n_r + ir);
		} else if (ir < n_r && iz < n_z) {
			acc.add_global(offset + iz*n_r + ir, wfix);
		}
	}
	__device__ __forceinline__ u32 window_index(const FluWindow &win, u32 local) const {
		u32 lr = local % win.ext0, lz = local/win.ext0;
		return offset + (lz + win.org1)*n_r + lr + win.org0;
	}
};

struct FluXyzt {                    // mcfluence/fluencet.py:57-63
	float inv_step[4], top_left[4]; u32 shape[4]; u32 offset; i32 k;
	static constexpr bool active = true;
	static constexpr bool needs_opl = true;
	static constexpr bool fixed_point = true;
	__device__ __forceinline__ u32 window_index(const FluWindow &, u32) const { return 0; }
	struct Prep { };
	struct Far { };
	__device__ __forceinline__ Prep prepare(const FluWindow &) const { return Prep(); }
	__device__ __forceinline__ Far prepare_far(const FluWindow &) const { return Far(); }
	__device__ __forceinline__ void deposit_prep(const Accu &acc, const Prep &, const FluWindow &win, const P3 &pos, u32 wfix, float opl) const {
		deposit_fixed(acc, win, pos, wfix, opl);
	}
	// deposit of a weight that is already in fixed point (throughput loops fold
	// the conversion constant into their per-layer constants, see fixed_scale)
	__device__ __forceinline__ void deposit(const Accu &acc, const FluWindow &win, const P3 &pos, float w, float mua, float opl) const {
		deposit_fixed(acc, win, pos, fluence_weight(w, mua, k), opl);
	}
	__device__ __forceinline__ float fixed_scale(float mua) const { return fluence_scale(mua, k); }
	__device__ __forceinline__ void deposit_fixed(const Accu &acc, const FluWindow &, const P3 &pos, u32 wfix, float opl) const {
		float fx = (pos.x - top_left[0])*inv_step[0];
		float fy = (pos.y - top_left[1])*inv_step[1];
		float fz = (pos.z - top_left[2])*inv_step[2];
		float ft = (opl*XO_FP_INV_C - top_left[3])*inv_step[3];
		if (ft >= 0.0f && fx >= 0.0f && fy >= 0.0f && fz >= 0.0f &&
				fx < (float)shape[0] && fy < (float)shape[1] &&
				fz < (float)shape[2] && ft < (float)shape[3]) {
			u32 index = ((f2u(fz)*shape[1] + f2u(fy))*shape[0] + f2u(fx))*shape[3] + f2u(ft);
			acc.add_global(offset + index, wfix);
		}
	}
};

struct FluRzt {                     // mcfluence/fluencerzt.py:54-66
	P3 center; float t_min, inv_dr, inv_dz, inv_dt; u32 n_r, n_z, n_t, offset; i32 k;
	static constexpr bool active = true;
	static constexpr bool needs_opl = true;
	static constexpr bool fixed_point = true;
	__device__ __forceinline__ u32 window_index(const FluWindow &, u32) const { return 0; }
	struct Prep { };
	struct Far { };
	__device__ __forceinline__ Prep prepare(const FluWindow &) const { return Prep(); }
	__device__ __forceinline__ Far prepare_far(const FluWindow &) const { return Far(); }
	__device__ __forceinline__ void deposit_prep(const Accu &acc, const Prep &, const FluWindow &win, const P3 &pos, u32 wfix, float opl) const {
		deposit_fixed(acc, win, pos, wfix, opl);
	}
	// deposit of a weight that is already in fixed point (throughput loops fold
	// the conversion constant into their per-layer constants, see fixed_scale)
	__device__ __forceinline__ void deposit(const Accu &acc, const FluWindow &win, const P3 &pos, float w, float mua, float opl) const {
		deposit_fixed(acc, win, pos, fluence_weight(w, mua, k), opl);
	}
	__device__ __forceinline__ float fixed_scale(float mua) const { return fluence_scale(mua, k); }
	__device__ __forceinline__ void deposit_fixed(const Accu &acc, const FluWindow &, const P3 &pos, u32 wfix, float opl) const {
		float dx = pos.x - center.x, dy = pos.y - center.y;
		float r = M::sqrt(dx*dx + dy*dy);
		float dz = pos.z - center.z;
		float dt = opl*XO_FP_INV_C - t_min;
		float fr = r*inv_dr, fz = dz*inv_dz, ft = dt*inv_dt;
		if (fr >= 0.0f && fz >= 0.0f && ft >= 0.0f &&
				fr < (float)n_r && fz < (float)n_z && ft < (float)n_t) {
			u32 index = (f2u(fz)*n_r + f2u(fr))*n_t + f2u(ft);
			acc.add_global(offset + index, wfix);
		}
	}
};

struct FluCyl {                     // mcfluence/fluencecyl.py:56-70
	P2 center; float r_min, fi_min, z_min, inv_dr, inv_dfi, inv_dz;
	u32 n_r, n_fi, n_z, offset; i32 k;
	static constexpr bool active = true;
	static constexpr bool needs_opl = false;
	static constexpr bool fixed_point = true;
	__device__ __forceinline__ u32 window_index(const FluWindow &, u32) const { return 0; }
	struct Prep { };
	struct Far { };
	__device__ __forceinline__ Prep prepare(const FluWindow &) const { return Prep(); }
	__device__ __forceinline__ Far prepare_far(const FluWindow &) const { return Far(); }
	__device__ __forceinline__ void deposit_prep(const Accu &acc, const Prep &, const FluWindow &win, const P3 &pos, u32 wfix, float opl) const {
		deposit_fixed(acc, win, pos, wfix, opl);
	}
	// deposit of a weight that is already in fixed point (throughput loops fold
	// the conversion constant into their per-layer constants, see fixed_scale)
	__device__ __forceinline__ void deposit(const Accu &acc, const FluWindow &win, const P3 &pos, float w, float mua, float opl) const {
		deposit_fixed(acc, win, pos, fluence_weight(w, mua, k), opl);
	}
	__device__ __forceinline__ float fixed_scale(float mua) const { return fluence_scale(mua, k); }
	__device__ __forceinline__ void deposit_fixed(const Accu &acc, const FluWindow &, const P3 &pos, u32 wfix, float) const {
		float dx = pos.x - center.x, dy = pos.y - center.y;
		float r = M::sqrt(dx*dx + dy*dy);
		float fi = M::atan2(dy, dx) + XO_FP_PI;
		float fr = (r - r_min)*inv_dr, fz = (pos.z - z_min)*inv_dz, ffi = (fi - fi_min)*inv_dfi;
		if (fr >= 0.0f && fz >= 0.0f && ffi >= 0.0f &&
				fr < (float)n_r && fz < (float)n_z && ffi < (float)n_fi) {
			u32 index = (f2u(fz)*n_fi + f2u(ffi))*n_r + f2u(fr);
			acc.add_global(offset + index, wfix);
		}
	}
};

struct FluCylt {                    // mcfluence/fluencecylt.py:79-95
	P2 center; float r_min, fi_min, z_min, t_min, inv_dr, inv_dfi, inv_dz, inv_dt;
	u32 n_r, n_fi, n_z, n_t, offset; i32 k;
	static constexpr bool active = true;
	static constexpr bool needs_opl = true;
	static constexpr bool fixed_point = true;
	__device__ __forceinline__ u32 window_index(const FluWindow &, u32) const { return 0; }
	struct Prep { };
	struct Far { };
	__device__ __forceinline__ Prep prepare(const FluWindow &) const { return Prep(); }
	__device__ __forceinline__ Far prepare_far(const FluWindow &) const { return Far(); }
	__device__ __forceinline__ void deposit_prep(const Accu &acc, const Prep &, const FluWindow &win, const P3 &pos, u32 wfix, float opl) const {
		deposit_fixed(acc, win, pos, wfix, opl);
	}
	__device__ __forceinline__ void deposit(const Accu &acc, const FluWindow &win, const P3 &pos, float w, float mua, float opl) const {
		deposit_fixed(acc, win, pos, fluence_weight(w, mua, k), opl);
	}
	__device__ __forceinline__ float fixed_scale(float mua) const { return fluence_scale(mua, k); }
	__device__ __forceinline__ void deposit_fixed(const Accu &acc, const FluWindow &, const P3 &pos, u32 wfix, float opl) const {
		float dx = pos.x - center.x, dy = pos.y - center.y;
		float r = M::sqrt(dx*dx + dy*dy);
		float fi = M::atan2(dy, dx) + XO_FP_PI;
		float dt = opl*XO_FP_INV_C - t_min;
		float fr = (r - r_min)*inv_dr, fz = (pos.z - z_min)*inv_dz;
		float ffi = (fi - fi_min)*inv_dfi, ft = dt*inv_dt;
		if (fr >= 0.0f && fz >= 0.0f && ffi >= 0.0f && ft >= 0.0f &&
				fr < (float)n_r && fz < (float)n_z && ffi < (float)n_fi && ft < (float)n_t) {
			u32 index = ((f2u(fz)*n_fi + f2u(ffi))*n_r + f2u(fr))*n_t + f2u(ft);
			acc.add_global(offset + index, wfix);
		}
	}
};

// ---- trace ------------------------------------------------------------------
#ifndef XO_TRACE
#define XO_TRACE 0
#endif
#ifndef XO_USE_EVENTS
#define XO_USE_EVENTS 0
#endif
#ifndef XO_TRACE_STORE_HINT
#define XO_TRACE_STORE_HINT 0
#endif
#define XO_TRACE_START 1
#define XO_TRACE_END 2
#define XO_TRACE_ALL 7

struct TraceCfg {                   // mcbase/mctrace.py:504-510
	i32 max_events; u32 data_off, count_off, event_mask;
};
struct TraceNone { i32 dummy; };
#ifndef XO_USER_TRACE
#define XO_USER_TRACE 0             // 1: the trace is a user-written fragment (xo_clcompat_slots.cuh)
#endif

// one event = 8 floats {x,y,z,px,py,pz,w,pl}; overflow keeps overwriting the
// last slot while the count keeps growing (mctrace.py:543-545,578-581)
__device__ __forceinline__ bool trace_event(const TraceCfg &t, float *fbuf, u32 packet,
		u32 count, u32 flags, const P3 &pos, const P3 &dir, float w, float opl) {
#if XO_USE_EVENTS
	if (!(t.event_mask & flags)) return false;
#else
	(void)flags;
#endif
	i32 slot = (i32)count < t.max_events - 1 ? (i32)count : t.max_events - 1;
	u32 p = (u32)slot*8u + packet*(u32)t.max_events*8u + t.data_off;
	float *dst = fbuf + p;
#if XO_TRACE_ALIGNED == 2
	// 32-byte aligned rows: the event is exactly one DRAM sector and leaves with ONE
	// 256-bit store (STG.E.256, sm_100+).  Measured on B200 with one 16 KB-strided
	// row per thread (tools/trace_store_probe.cu, profiles/trace_store_probe_r02.json):
	// 3.3-4.4 TB/s against 2.2 TB/s for two STG.128 and 3.0-4.2 TB/s for 128-byte
	// lines assembled by 8 lanes in shared memory (the round-1 path, ~100 extra
	// warp-instructions per trip); a per-lane cp.async.bulk serialises (UBLKCP takes
	// uniform registers: one lane at a time).  Parking the event in registers and storing
	// it one trip later (so that the warp does not wait behind the store until the LSU
	// has read the operands, 46 % of the stall samples) changed nothing: 2.75 vs 2.69 ms
	// per 1e6 packets of C4 - the SM's store path itself (~16 B/clk) is the limit.
#if XO_TRACE_STORE_HINT == 1
	// (experiment: the event stream marked evict-first in L2 and not allocated in L1, so
	// that it does not wash the voxel map / lookup tables out of the caches)
	{
		unsigned long long pol_;
		asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_));
		asm volatile("st.global.L1::no_allocate.L2::cache_hint.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8}, %9;"
			:: "l"(dst), "f"(pos.x), "f"(pos.y), "f"(pos.z), "f"(dir.x), "f"(dir.y),
			   "f"(dir.z), "f"(w), "f"(opl), "l"(pol_) : "memory");
	}
#elif XO_TRACE_STORE_HINT == 2
	asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
		:: "l"(dst), "f"(pos.x), "f"(pos.y), "f"(pos.z), "f"(dir.x), "f"(dir.y),
		   "f"(dir.z), "f"(w), "f"(opl) : "memory");
#else
	asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
		:: "l"(dst), "f"(pos.x), "f"(pos.y), "f"(pos.z), "f"(dir.x), "f"(dir.y),
		   "f"(dir.z), "f"(w), "f"(opl) : "memory");
#endif
#elif XO_TRACE_ALIGNED
	float4 *d4 = reinterpret_cast<float4 *>(dst);
	d4[0] = make_float4(pos.x, pos.y, pos.z, dir.x);
	d4[1] = make_float4(dir.y, dir.z, w, opl);
#else
	dst[0] = pos.x; dst[1] = pos.y; dst[2] = pos.z;
	dst[3] = dir.x; dst[4] = dir.y; dst[5] = dir.z;
	dst[6] = w; dst[7] = opl;
#endif
	return true;
}

// a packet is finished: its event count (mctrace.py:578-584)
__device__ __forceinline__ void trace_complete(const TraceCfg &t, i32 *ibuf, u32 packet, u32 count) {
	ibuf[t.count_off + packet] = (i32)count;
}

// XoTrace: the kernel parameter; XoTraceCfg: what the loops hand to trace_event /
// trace_complete (a kernel without a trace still names a TraceCfg it never touches)
#if XO_TRACE && XO_USER_TRACE
struct TraceUser;
typedef TraceUser XoTrace;
typedef TraceUser XoTraceCfg;
#elif XO_TRACE
typedef TraceCfg XoTrace;
typedef TraceCfg XoTraceCfg;
#else
typedef TraceNone XoTrace;
typedef TraceCfg XoTraceCfg;
#endif

}  // namespace xo
