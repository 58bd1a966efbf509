// xo_trace_kernels.cuh -- device-side trace filter + stable compaction.
//
// The reference stores the trace of EVERY packet (32 B x maxlen per packet),
// downloads all of it and filters on the host by the packet's terminal event
// (`Filter.__call__`, xopto/mcbase/mctrace.py:216-330).  Here the predicate is
// evaluated where the rows are, the accepted rows are compacted in packet order
// (stable: the result is byte-identical to the host filter) and only those rows
// ever cross PCIe - or none, when `Mc.sampling_volume` consumes them in place.
//
// Three launches: TraceFilterFlags (predicate + per-CTA counts),
// TraceFilterScan (one CTA: exclusive scan of the CTA counts), TraceCompact
// (per-CTA exclusive scan of the flags, then one warp copies one row with
// 128-bit loads/stores, 512 B per pass).
//
// Predicate arithmetic = the numpy float32 arithmetic of the host filter:
// constants are rounded to binary32 on the host, products and sums are rounded
// one by one (no FMA contraction).
#pragma once
#include "xo_core.cuh"
#include "xo_fluence.cuh"     // TraceCfg

namespace xo {

// number of (low, high) ranges per key; the range values live in one float
// array in this order: x, y, z, pz: (lo, hi); r: (r0^2, r1^2, x0, y0);
// dir: (c0, c1, px, py, pz); pl: (lo, hi)
struct FilterCfg { u32 nx, ny, nz, npz, nr, ndir, npl, plon; };

#define XO_FILTER_BLOCK 256

__device__ __forceinline__ bool filter_stage(const float *&g, u32 count, float v, bool &valid, bool &var, bool reset) {
	if (count == 0u) return valid;
	if (reset) var = false;
	for (u32 i = 0; i < count; ++i, g += 2)
		var |= (v >= g[0] && v <= g[1]) && valid;
	valid = valid && var;
	return valid;
}

// predicate of mctrace.py:216-330 on one terminal event {x,y,z,px,py,pz,w,pl}
__device__ inline bool filter_accepts(const FilterCfg &f, const float *ranges, const float *ev) {
	const float *g = ranges;
	bool valid = true, var = false;
	filter_stage(g, f.nx, ev[0], valid, var, true);
	filter_stage(g, f.ny, ev[1], valid, var, true);
	filter_stage(g, f.nz, ev[2], valid, var, true);
	filter_stage(g, f.npz, ev[5], valid, var, true);
	if (f.nr) {
		var = false;
		for (u32 i = 0; i < f.nr; ++i, g += 4) {
			float dx = __fsub_rn(ev[0], g[2]), dy = __fsub_rn(ev[1], g[3]);
			float rr = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
			var |= (rr >= g[0] && rr <= g[1]) && valid;
		}
		valid = valid && var;
	}
	if (f.ndir) {
		var = false;
		for (u32 i = 0; i < f.ndir; ++i, g += 5) {
			float ct = __fadd_rn(__fadd_rn(__fmul_rn(ev[3], g[2]), __fmul_rn(ev[4], g[3])),
				__fmul_rn(ev[5], g[4]));
			var |= (ct >= g[0] && ct <= g[1]) && valid;
		}
		valid = valid && var;
	}
	// the host filter does not clear its OR-mask before the pl stage
	// (mctrace.py:300-308); kept so that both paths select the same packets
	if (f.npl && f.plon) filter_stage(g, f.npl, ev[7], valid, var, false);
	return valid;
}

}  // namespace xo

// flags[i] = 1 when packet i passes the filter and its trace did not overflow;
// cta_counts[b] = number of flagged packets of CTA b; counters[0] += packets
// that pass the filter but overflowed (the host reports them as n_dropped).
extern "C" __global__ void __launch_bounds__(XO_FILTER_BLOCK)
TraceFilterFlags(
	xo::u32 npackets,
	const __grid_constant__ xo::TraceCfg trace,
	const __grid_constant__ xo::FilterCfg filter,
	const float *ranges,
	const xo::i32 *int_buffer,
	const float *fp_buffer,
	xo::u32 *flags,
	xo::u32 *cta_counts,
	xo::u32 *counters)
{
	using namespace xo;
	const u32 i = blockIdx.x*blockDim.x + threadIdx.x;
	u32 flag = 0, dropped = 0;
	if (i < npackets) {
		const i32 n = int_buffer[trace.count_off + i];
		i32 last = n - 1 < trace.max_events - 1 ? n - 1 : trace.max_events - 1;
		if (last < 0) last += trace.max_events;          // numpy negative index
		const float *row = fp_buffer + trace.data_off + (u64)i*(u64)trace.max_events*8u;
		float ev[8];
		for (int k = 0; k < 8; ++k) ev[k] = row[(u64)last*8u + k];
		const bool valid = filter_accepts(filter, ranges, ev);
		const bool overflow = n >= trace.max_events;
		flag = (valid && !overflow) ? 1u : 0u;
		dropped = (valid && overflow) ? 1u : 0u;
		flags[i] = flag;
	}
	const u32 cnt = __syncthreads_count((int)flag);
	const u32 drp = __syncthreads_count((int)dropped);
	if (threadIdx.x == 0) {
		cta_counts[blockIdx.x] = cnt;
		if (drp) atomicAdd(counters, drp);
	}
}

// exclusive scan of cta_counts[0..n) in place (one CTA); counters[1] = total
extern "C" __global__ void __launch_bounds__(1024)
TraceFilterScan(xo::u32 n, xo::u32 *cta_counts, xo::u32 *counters)
{
	using namespace xo;
	__shared__ u32 warp_sums[32];
	__shared__ u32 carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	for (u32 base = 0; base < n; base += blockDim.x) {
		const u32 i = base + threadIdx.x;
		const u32 v = i < n ? cta_counts[i] : 0u;
		u32 s = v;
		for (int o = 1; o < 32; o <<= 1) {
			u32 t = __shfl_up_sync(0xffffffffu, s, o);
			if (lane >= (u32)o) s += t;
		}
		if (lane == 31u) warp_sums[warp] = s;
		__syncthreads();
		if (warp == 0) {
			u32 w = warp_sums[lane];
			for (int o = 1; o < 32; o <<= 1) {
				u32 t = __shfl_up_sync(0xffffffffu, w, o);
				if (lane >= (u32)o) w += t;
			}
			warp_sums[lane] = w;
		}
		__syncthreads();
		const u32 before = carry + (warp ? warp_sums[warp - 1] : 0u) + s - v;
		if (i < n) cta_counts[i] = before;
		__syncthreads();
		if (threadIdx.x == blockDim.x - 1) carry = before + v;
		__syncthreads();
	}
	if (threadIdx.x == 0) counters[1] = carry;
}

// stable compaction: flagged packet i of CTA b goes to row cta_offsets[b] + (its
// rank among the flagged packets of the CTA).  Events beyond the packet's count
// are written as zeros (the trace buffer is zero-filled before every run).
extern "C" __global__ void __launch_bounds__(XO_FILTER_BLOCK)
TraceCompact(
	xo::u32 npackets,
	const __grid_constant__ xo::TraceCfg trace,
	const xo::u32 *flags,
	const xo::u32 *cta_offsets,
	const xo::i32 *int_buffer,
	const float *fp_buffer,
	xo::i32 *out_counts,
	float *out_rows)
{
	using namespace xo;
	__shared__ u32 dest[XO_FILTER_BLOCK];
	__shared__ u32 warp_sums[XO_FILTER_BLOCK/32];
	const u32 i = blockIdx.x*blockDim.x + threadIdx.x;
	const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const u32 flag = i < npackets ? flags[i] : 0u;
	const u32 ballot = __ballot_sync(0xffffffffu, flag != 0u);
	if (lane == 0u) warp_sums[warp] = (u32)__popc(ballot);
	__syncthreads();
	u32 before = cta_offsets[blockIdx.x];
	for (u32 w = 0; w < warp; ++w) before += warp_sums[w];
	before += (u32)__popc(ballot & ((1u << lane) - 1u));
	dest[threadIdx.x] = flag ? before : 0xffffffffu;
	if (flag) out_counts[before] = int_buffer[trace.count_off + i];
	__syncthreads();
	// one warp per row
	const u32 row_f4 = (u32)trace.max_events*2u;           // float4 per row
	const bool aligned = (trace.data_off & 3u) == 0u;
	for (u32 p = warp; p < blockDim.x; p += blockDim.x/32u) {
		const u32 d = dest[p];
		if (d == 0xffffffffu) continue;
		const u32 src_packet = blockIdx.x*blockDim.x + p;
		i32 n = int_buffer[trace.count_off + src_packet];
		n = n < trace.max_events ? n : trace.max_events;
		const float *src = fp_buffer + trace.data_off + (u64)src_packet*(u64)trace.max_events*8u;
		float *dst = out_rows + (u64)d*(u64)trace.max_events*8u;
		if (aligned) {
			const float4 *s4 = reinterpret_cast<const float4 *>(src);
			float4 *d4 = reinterpret_cast<float4 *>(dst);
			const u32 used = (u32)n*2u;
			for (u32 k = lane; k < row_f4; k += 32u)
				d4[k] = k < used ? s4[k] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		} else {
			const u32 used = (u32)n*8u;
			for (u32 k = lane; k < row_f4*4u; k += 32u) dst[k] = k < used ? src[k] : 0.0f;
		}
	}
}
