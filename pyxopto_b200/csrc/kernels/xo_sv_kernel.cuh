// xo_sv_kernel.cuh -- sampling-volume analysis of packet traces.
//
// B200 counterpart of the `SamplingVolume` kernel in
// xopto/mcbase/kernel/mcsv.template.c:236-420.  Every traced packet is walked
// event by event through a voxel grid; each voxel it crosses receives
// (terminal weight) x (path length inside the voxel) in 64-bit fixed point, the
// terminal weights are summed in `total_weight`.
//
// Machine mapping: one trace row per thread, rows claimed from a global counter
// (walk lengths are heavy-tailed); a row is 32 B x maxlen contiguous floats, read
// with two 128-bit loads per event (a full 32 B sector each, one event ahead of
// the walk); voxel deposits are RED.E.ADD.64 into the grid (L2); the terminal
// weights are summed per thread, reduced per warp and added with one RED per
// warp.  No random numbers and integer accumulation only: the result is
// independent of the row -> thread schedule, so in deterministic mode (IEEE
// division / square root, no FMA contraction) it is bit-identical to the
// reference kernel.
#pragma once
#include "xo_core.cuh"
#include "xo_fluence.cuh"     // TraceCfg

namespace xo {

struct SvCfg {                      // mcbase/mcsv.py:36-84
	P3 top_left, voxel_size; u32 nx, ny, nz; float multiplier; u32 offset; i32 k;
};
struct SvEvent { P3 pos, dir; float weight; };

__device__ __forceinline__ void sv_load_event(const float *row, bool aligned, i32 index, SvEvent &ev) {
	const float *e = row + (i64)index*8;
	if (aligned) {
		const float4 a = __ldg(reinterpret_cast<const float4 *>(e));
		const float4 b = __ldg(reinterpret_cast<const float4 *>(e) + 1);
		ev.pos.x = a.x; ev.pos.y = a.y; ev.pos.z = a.z;
		ev.dir.x = a.w; ev.dir.y = b.x; ev.dir.z = b.y;
		ev.weight = b.z;
	} else {
		ev.pos.x = e[0]; ev.pos.y = e[1]; ev.pos.z = e[2];
		ev.dir.x = e[3]; ev.dir.y = e[4]; ev.dir.z = e[5];
		ev.weight = e[6];
	}
}
__device__ __forceinline__ float sv_distance(const P3 &a, const P3 &b) {
	float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
	return M::sqrt(dx*dx + dy*dy + dz*dz);
}
__device__ __forceinline__ void sv_deposit(const SvCfg &sv, u64 *accu, i32 vx, i32 vy, i32 vz,
		float end_weight, float l_voxel) {
	if (vx >= 0 && vy >= 0 && vz >= 0 && (u32)vx < sv.nx && (u32)vy < sv.ny && (u32)vz < sv.nz) {
		u64 index = ((u64)vz*sv.ny + (u64)vy)*sv.nx + (u64)vx;
		u32 w = f2u(end_weight*l_voxel*sv.multiplier*(float)sv.k + 0.5f);
		atomicAdd(accu + sv.offset + index, (u64)w);
	}
}

}  // namespace xo

extern "C" __global__ void __launch_bounds__(256)
SamplingVolume(
	xo::u32 npackets,
	xo::u32 *npackets_processed,
	xo::u32 *num_kernels,
	const __grid_constant__ xo::TraceCfg trace,
	const __grid_constant__ xo::SvCfg sv,
	xo::u64 *total_weight,
	const xo::i32 *int_buffer,
	const float *fp_buffer,
	xo::u64 *accu_buffer)
{
	using namespace xo;
	// (128-bit loads of an event: single precision only - in binary64 `float` is double
	// and an event is 64 bytes)
	const bool aligned = !XO_DOUBLE && (trace.data_off & 3u) == 0u &&
		(reinterpret_cast<unsigned long long>(fp_buffer) & 15ull) == 0ull;
	u64 weight_sum = 0;
	u32 steps = 0;
	bool started = false;
	u32 packet;
	while ((packet = atomicAdd(npackets_processed, 1u)) < npackets) {
		started = true;
		const float *row = fp_buffer + trace.data_off + (u64)packet*(u64)trace.max_events*8u;
		i32 n = int_buffer[trace.count_off + packet];
		n = n < trace.max_events ? n : trace.max_events;
		// the reference reads row[-1] for an empty trace (never produced by Trace)
		if (n < 1) continue;
		// field 7 (optical path length) of the last event is what the reference
		// calls the terminal weight (mcsv.template.c:299 with mctrace.py:563-574)
		const float end_weight = row[(i64)n*8 - 1];
		weight_sum += f2u(end_weight*(float)sv.k + 0.5f);
		SvEvent ev1, ev2;
		sv_load_event(row, aligned, 0, ev1);
		sv_load_event(row, aligned, (1 < n - 1) ? 1 : n - 1, ev2);
		float d_ev = sv_distance(ev1.pos, ev2.pos);
		i32 vx = f2i(M::div(ev1.pos.x - sv.top_left.x, sv.voxel_size.x));
		i32 vy = f2i(M::div(ev1.pos.y - sv.top_left.y, sv.voxel_size.y));
		i32 vz = f2i(M::div(ev1.pos.z - sv.top_left.z, sv.voxel_size.z));
		float l_voxel = 0.0f;
		i32 event_index = 0;
		for (;;) {
			++steps;
			float dx = sv.top_left.x + (float)((ev1.dir.x >= 0.0f ? 1 : 0) + vx)*sv.voxel_size.x - ev1.pos.x;
			float dy = sv.top_left.y + (float)((ev1.dir.y >= 0.0f ? 1 : 0) + vy)*sv.voxel_size.y - ev1.pos.y;
			float dz = sv.top_left.z + (float)((ev1.dir.z >= 0.0f ? 1 : 0) + vz)*sv.voxel_size.z - ev1.pos.z;
			dx = (ev1.dir.x != 0.0f) ? M::div(dx, ev1.dir.x) : XO_INF;
			dy = (ev1.dir.y != 0.0f) ? M::div(dy, ev1.dir.y) : XO_INF;
			dz = (ev1.dir.z != 0.0f) ? M::div(dz, ev1.dir.z) : XO_INF;
			const float d_voxel = fminf(dz, fminf(dx, dy));
			const float d_move = fminf(d_voxel, d_ev);
			l_voxel += d_move;
			if (d_voxel < d_ev) {
				sv_deposit(sv, accu_buffer, vx, vy, vz, end_weight, l_voxel);
				ev1.pos.x += ev1.dir.x*d_move;
				ev1.pos.y += ev1.dir.y*d_move;
				ev1.pos.z += ev1.dir.z*d_move;
				d_ev -= d_move;
				const i32 nx = (dx <= dy && dx <= dz) ? 1 : 0;
				const i32 ny = nx ? 0 : ((dy <= dx && dy <= dz) ? 1 : 0);
				const i32 nz = !(nx + ny);
				vx += ev1.dir.x < 0.0f ? -nx : nx;
				vy += ev1.dir.y < 0.0f ? -ny : ny;
				vz += ev1.dir.z < 0.0f ? -nz : nz;
				l_voxel = 0.0f;
			} else {
				ev1 = ev2;
				event_index += 1;
				if (event_index < n - 1) {
					sv_load_event(row, aligned, event_index + 1, ev2);
					d_ev = sv_distance(ev1.pos, ev2.pos);
				} else {
					sv_deposit(sv, accu_buffer, vx, vy, vz, end_weight, l_voxel);
					break;
				}
			}
		}
	}
	if (started) atomicAdd(num_kernels, 1u);
	// exact 64-bit sums per warp (the block size is a multiple of 32 and every
	// lane reaches this point), then one RED per warp
	{
		u64 total = weight_sum;
		for (int o = 16; o > 0; o >>= 1) total += __shfl_down_sync(0xffffffffu, total, o);
		const u32 warp_steps = __reduce_add_sync(0xffffffffu, steps);
		if ((threadIdx.x & 31u) == 0u) {
			if (total) atomicAdd(total_weight, total);
			if (warp_steps) atomicAdd(reinterpret_cast<u64 *>(num_kernels + 1), (u64)warp_steps);
		}
	}
}

#if !XO_DETERMINISTIC && !XO_DOUBLE
// Throughput mode: one WARP per packet, one lane per segment (event i -> event i + 1).
//
// The kernel above follows the reference - one work-item walks all events of a packet, a
// serial chain of up to maxlen segments x voxel crossings, each step waiting for the event
// loads (config 4: 12 800 accepted packets keep 12 800 threads of the 300 000 the device
// holds busy, 1.5 ms for 6.3e6 steps).  Segments are independent: the voxel a segment
// starts in follows from its start point, and the path length a packet leaves in a voxel
// is the sum over its segments.  Here the 32 lanes of a warp take the segments of one
// packet side by side, each walks its own segment through the voxel grid and deposits
// end_weight x length per visited voxel.  Differences from the reference, both inside the
// fixed-point resolution: a voxel that several consecutive segments stay in receives one
// rounded deposit per segment instead of one for their sum, and a segment that starts
// exactly on a voxel face starts in the voxel behind the face and leaves it after a step
// of zero length.
extern "C" __global__ void __launch_bounds__(256)
SamplingVolumeWarp(
	xo::u32 npackets,
	xo::u32 *npackets_processed,
	xo::u32 *num_kernels,
	const __grid_constant__ xo::TraceCfg trace,
	const __grid_constant__ xo::SvCfg sv,
	xo::u64 *total_weight,
	const xo::i32 *int_buffer,
	const float *fp_buffer,
	xo::u64 *accu_buffer)
{
	using namespace xo;
	const bool aligned = (trace.data_off & 3u) == 0u &&
		(reinterpret_cast<unsigned long long>(fp_buffer) & 15ull) == 0ull;
	const u32 lane = threadIdx.x & 31u;
	const float inv_x = 1.0f/sv.voxel_size.x, inv_y = 1.0f/sv.voxel_size.y, inv_z = 1.0f/sv.voxel_size.z;
	u64 weight_sum = 0;
	u32 steps = 0;
	bool started = false;
	for (;;) {
		u32 packet = 0;
		if (lane == 0u) packet = atomicAdd(npackets_processed, 1u);
		packet = __shfl_sync(0xffffffffu, packet, 0);
		if (packet >= npackets) break;
		started = true;
		const float *row = fp_buffer + trace.data_off + (u64)packet*(u64)trace.max_events*8u;
		i32 n = int_buffer[trace.count_off + packet];
		n = n < trace.max_events ? n : trace.max_events;
		if (n < 1) continue;
		const float end_weight = row[(i64)n*8 - 1];
		if (lane == 0u) weight_sum += f2u(end_weight*(float)sv.k + 0.5f);
		const float dep_k = end_weight*sv.multiplier*(float)sv.k;
		for (i32 i = (i32)lane; i < n - 1; i += 32) {
			SvEvent ev1, ev2;
			sv_load_event(row, aligned, i, ev1);
			sv_load_event(row, aligned, i + 1, ev2);
			float d_ev = sv_distance(ev1.pos, ev2.pos);
			// voxel of the start point: floor - except for the first event, which the
			// reference truncates (mcsv.template.c:69-71, int conversion; differs only outside the grid)
			const float qx = (ev1.pos.x - sv.top_left.x)*inv_x, qy = (ev1.pos.y - sv.top_left.y)*inv_y,
				qz = (ev1.pos.z - sv.top_left.z)*inv_z;
			i32 vx = (i == 0) ? f2i(qx) : __float2int_rd(qx);
			i32 vy = (i == 0) ? f2i(qy) : __float2int_rd(qy);
			i32 vz = (i == 0) ? f2i(qz) : __float2int_rd(qz);
			const float rx = (ev1.dir.x != 0.0f) ? FastMath::rcp_approx(ev1.dir.x) : XO_INF;
			const float ry = (ev1.dir.y != 0.0f) ? FastMath::rcp_approx(ev1.dir.y) : XO_INF;
			const float rz = (ev1.dir.z != 0.0f) ? FastMath::rcp_approx(ev1.dir.z) : XO_INF;
			const i32 sx = ev1.dir.x < 0.0f ? -1 : 1, sy = ev1.dir.y < 0.0f ? -1 : 1, sz = ev1.dir.z < 0.0f ? -1 : 1;
			const i32 fx = ev1.dir.x >= 0.0f ? 1 : 0, fy = ev1.dir.y >= 0.0f ? 1 : 0, fz = ev1.dir.z >= 0.0f ? 1 : 0;
			for (u32 guard = 0; guard < (1u << 20); ++guard) {
				++steps;
				float dx = fmaf((float)(fx + vx), sv.voxel_size.x, sv.top_left.x) - ev1.pos.x;
				float dy = fmaf((float)(fy + vy), sv.voxel_size.y, sv.top_left.y) - ev1.pos.y;
				float dz = fmaf((float)(fz + vz), sv.voxel_size.z, sv.top_left.z) - ev1.pos.z;
				dx = (ev1.dir.x != 0.0f) ? dx*rx : XO_INF;
				dy = (ev1.dir.y != 0.0f) ? dy*ry : XO_INF;
				dz = (ev1.dir.z != 0.0f) ? dz*rz : XO_INF;
				const float d_voxel = fminf(dz, fminf(dx, dy));
				const bool crossing = d_voxel < d_ev;
				const float d_move = crossing ? d_voxel : d_ev;
				if (vx >= 0 && vy >= 0 && vz >= 0 && (u32)vx < sv.nx && (u32)vy < sv.ny && (u32)vz < sv.nz) {
					const u32 w = f2u(fmaf(dep_k, d_move, 0.5f));
					if (w) atomicAdd(accu_buffer + sv.offset +
						(((u64)vz*sv.ny + (u64)vy)*sv.nx + (u64)vx), (u64)w);
				}
				if (!crossing) break;
				ev1.pos.x = fmaf(ev1.dir.x, d_move, ev1.pos.x);
				ev1.pos.y = fmaf(ev1.dir.y, d_move, ev1.pos.y);
				ev1.pos.z = fmaf(ev1.dir.z, d_move, ev1.pos.z);
				d_ev -= d_move;
				const bool nx = (dx <= dy && dx <= dz);
				const bool ny = !nx && (dy <= dz);
				vx += nx ? sx : 0;
				vy += ny ? sy : 0;
				vz += (!nx && !ny) ? sz : 0;
			}
		}
	}
	if (started && lane == 0u) atomicAdd(num_kernels, 1u);
	{
		const u32 warp_steps = __reduce_add_sync(0xffffffffu, steps);
		if (lane == 0u) {
			if (weight_sum) atomicAdd(total_weight, weight_sum);
			if (warp_steps) atomicAdd(reinterpret_cast<u64 *>(num_kernels + 1), (u64)warp_steps);
		}
	}
}
#endif
