// mcvox_pool_loop.cuh -- throughput loop of the voxel kernel with a per-warp packet
// pool (body fragment, included inside McKernel of mcvox_kernel.cuh when XO_VOX_POOL).
//
// Same physics and the same ray formulation as mcvox_dda_loop.cuh (one ray per
// flight, incremental voxel walk over the compact 16-bit map, clearance shortcut,
// inline material change at equal refractive index; mcvox.template.c:667-1014), but
// the packets do not live in the registers of "their" lane.  In the lane-resident loop
// every phase - interaction, ray set-up, voxel walk, interface - runs with the lanes
// that happen to be in that state: 22 / 5 / 12 / 1.3 of 32 on the 201^3 skin model
// (profiles/r02s_*c3*).  Here a warp owns XO_VOX_POOL (64) packet slots in shared
// memory; every round it picks the phase most slots wait for, takes up to 32 slots of that
// class, loads exactly the fields the phase needs, runs the phase with (nearly) full lanes
// - several passes while enough lanes stay in the class - and stores the packets back.
// Which slots wait for what is kept in one ring of slot numbers per class with
// warp-uniform heads and counts (XO_POOL_QUEUES, default), or found by a census of the
// slot states (class counts in the bytes of one word, one REDUX) and a gather by rank:
//
//   INTERACT  slots NEW / RAY / SCAT / FAR: [voxel of the end point of a flight that
//             skipped the walk] deposit, scattering, lottery, then the next free path
//             and the clearance test -> FAR (again this phase) / DDA / EMPTY
//   WALK      slots DDA (first: walk set-up) and RUN: two crossings per trip until fewer
//             than XO_POOL_THR_W lanes still walk -> SCAT / BND / RUN
//   BOUNDARY  slots BND: interface physics at a face between materials of different
//             refractive index, or the face of the grid -> RAY / EMPTY
//   LAUNCH    EMPTY slots, while the packet budget lasts -> NEW
//
// A slot is 17 words: A = pos | weight, B = dir | free path (or the event parameter of
// a BND slot), C = next-face parameters | voxel address, D = per-voxel increments |
// material, clearance, axis of the last crossing, direction signs; E = optical path
// length.  The MWC stream belongs to the lane, not to the packet (throughput mode is
// not stream-aligned with the reference anyway).
//
// With a trace (XO_TRACE) a slot carries a fifth quad T = optical path length | packet
// index | recorded events | pending event flags, every phase records its events with
// trace_event (one 256-bit store per event) exactly where the lane-resident loop does, and
// for full traces the clearance shortcut is off (one event per voxel crossing is the
// product there): every flight walks.
//
// Where the rmax sphere around the source can be reached (XO_USE_RMAX) a slot also keeps the
// ray parameter at which its flight leaves the sphere (a ray leaves a convex body once): the
// first face beyond it is a BND event, the end-of-trip test of the reference follows every
// interaction and interface.
//
// Host conditions (mcvox/mc.py): compact map, throughput mode, albedo weight /
// albedo rejection, isotropic materials.
{
	// slot states: the class (INTERACT / WALK set-up / WALK / BOUNDARY) sits in bits 3-4
	enum : u32 { PS_EMPTY = 0, PS_NEW = 1, PS_RAY = 2, PS_SCAT = 3, PS_FAR = 4, PS_DDA = 8, PS_RUN = 16, PS_BND = 24 };
	enum : u32 { PH_INTERACT = 0, PH_WALK = 1, PH_BOUNDARY = 2, PH_LAUNCH = 3 };
#ifndef XO_POOL_THR_I
#define XO_POOL_THR_I 20        // INTERACT repeats while this many lanes skip the walk again
#endif
#ifndef XO_POOL_THR_W
#define XO_POOL_THR_W 12        // WALK goes on while this many lanes still walk
#endif
#ifndef XO_POOL_LAUNCH
#define XO_POOL_LAUNCH 16       // EMPTY slots that trigger a LAUNCH
#endif
#ifndef XO_POOL_BND
#define XO_POOL_BND 16          // BND slots that trigger a BOUNDARY phase
#endif
#ifndef XO_POOL_SOA
#define XO_POOL_SOA 1           // slot fields as one array per component (see below)
#endif
#ifndef XO_POOL_QUEUES
#define XO_POOL_QUEUES 1        // 1: rings of slot numbers per class; 0: census + gather by rank
#endif
	static_assert(XO_VOX_POOL == 64, "the census reads two slots per lane");
	constexpr u32 S = XO_VOX_POOL;
	const u32 lane = threadIdx.x & 31u;
	const u32 lanemask_lt = (1u << lane) - 1u;
	const u32 vox_bxy = vox_bx + vox_by;
	const u32 vox_mx = (1u << vox_bx) - 1u, vox_my = (1u << vox_by) - 1u;
	(void)chunk; (void)nthreads; (void)refill; (void)rmax2; (void)src_pos; (void)int_buffer; (void)float_buffer;
#if XO_USE_RMAX
#define XO_POOL_RMAX_TEST() do { \
		const float ex_ = pos.x - src_pos.x, ey_ = pos.y - src_pos.y, ez_ = pos.z - src_pos.z; \
		if (ex_*ex_ + ey_*ey_ + ez_*ez_ > rmax2) { done = true; flags |= EV_ESCAPED; } \
	} while (0)
#else
#define XO_POOL_RMAX_TEST() do { } while (0)
#endif
	u32 packet = 0, trace_count = 0, flags = 0;     // (trace builds: of the slot in hand)
	(void)packet; (void)trace_count; (void)flags;
	const u32 vbase_lo = (u32)reinterpret_cast<u64>(voxels8);
	const u32 vbase_hi = (u32)(reinterpret_cast<u64>(voxels8) >> 32);
	const float inv_sx = 1.0f/cfg.size.x, inv_sy = 1.0f/cfg.size.y, inv_sz = 1.0f/cfg.size.z;
#define XO_VOXEL(lo) ((u32)__ldg(reinterpret_cast<const unsigned short *>(((u64)vbase_hi << 32) | (u64)(lo))))
#define XO_PACK_VOXEL(ix_, iy_, iz_) (vbase_lo + 2u*((u32)((ix_) + 2) | ((u32)((iy_) + 2) << vox_bx) | ((u32)((iz_) + 2) << vox_bxy)))
#define XO_LOAD_MAT(idx) do { const VoxFastMat &F_ = sh_fast[idx]; c_hot = F_.hot; c_pf = F_.pf.v; } while (0)
#define XO_POOL_CLEARANCE (XO_TRACE != XO_TRACE_ALL)
	// Slot fields.  XO_POOL_SOA (default): one array of 64 floats per component - a lane
	// that holds slot s reads bank s mod 32 whatever the component, so the slots a phase
	// gathered (any 32 of 64) conflict two-way at most.  As 64 x float4 per quad the same
	// accesses are 128-bit wide and collide whenever two lanes of a quarter-warp hold slots
	// that are congruent mod 8: measured 9.8e8 bank-conflict cycles per 2e7 packets, the L1 /
	// shared-memory pipe 84 % busy and binding (profiles/r03s_*).
#if XO_POOL_SOA
#define XO_PF_A 0u
#define XO_PF_B 4u
#define XO_PF_C 8u
#define XO_PF_D 12u
#define XO_PF_T 16u
#define XO_PF(f_) P_F[(f_)*S + slot]
#define P4_LOAD(N_) make_float4(XO_PF(XO_PF_##N_), XO_PF(XO_PF_##N_ + 1u), XO_PF(XO_PF_##N_ + 2u), XO_PF(XO_PF_##N_ + 3u))
#define P4_STORE(N_, x_, y_, z_, w_) do { XO_PF(XO_PF_##N_) = (x_); XO_PF(XO_PF_##N_ + 1u) = (y_); \
		XO_PF(XO_PF_##N_ + 2u) = (z_); XO_PF(XO_PF_##N_ + 3u) = (w_); } while (0)
#define P3_STORE(N_, x_, y_, z_) do { XO_PF(XO_PF_##N_) = (x_); XO_PF(XO_PF_##N_ + 1u) = (y_); \
		XO_PF(XO_PF_##N_ + 2u) = (z_); } while (0)
#define PW_LOAD(N_) XO_PF(XO_PF_##N_ + 3u)
#define PW_STORE(N_, v_) (XO_PF(XO_PF_##N_ + 3u) = (v_))
#define P_E_REF XO_PF(16u)
#define P_R_REF XO_PF(XO_POOL_FIELDS - 1u)
#else
#define P4_LOAD(N_) (P_##N_[slot])
#define P4_STORE(N_, x_, y_, z_, w_) (P_##N_[slot] = make_float4(x_, y_, z_, w_))
#define P3_STORE(N_, x_, y_, z_) do { P_##N_[slot].x = (x_); P_##N_[slot].y = (y_); P_##N_[slot].z = (z_); } while (0)
#define PW_LOAD(N_) (P_##N_[slot].w)
#define PW_STORE(N_, v_) (P_##N_[slot].w = (v_))
#define P_E_REF P_E[slot]
#define P_R_REF P_R[slot]
#endif
#if XO_TRACE
	const XoTraceCfg &tcfg = *reinterpret_cast<const XoTraceCfg *>(&trace);
	// trace quad of a slot: optical path length | packet | recorded events | pending flags
#define XO_POOL_LOAD_T() do { const float4 t_ = P4_LOAD(T); opl = t_.x; packet = __float_as_uint(t_.y); \
		trace_count = __float_as_uint(t_.z); flags = __float_as_uint(t_.w); } while (0)
#define XO_POOL_STORE_T() do { P4_STORE(T, opl, __uint_as_float(packet), \
		__uint_as_float(trace_count), __uint_as_float(flags)); } while (0)
	// end of a loop trip of the reference (mcvox.template.c:983-1012): the event, the count
#define XO_POOL_TRACE_TRIP() do { \
		flags |= done ? EV_TERMINATED : 0u; \
		if (XO_TRACE == XO_TRACE_ALL || ((XO_TRACE & XO_TRACE_END) && done)) { \
			if (trace_event(tcfg, float_buffer, packet, trace_count, flags, pos, dir, weight, opl)) \
				++trace_count; \
		} \
		if (done) trace_complete(tcfg, int_buffer, packet, trace_count); \
		flags = 0u; \
	} while (0)
#define XO_POOL_OPL (true)
#else
#define XO_POOL_LOAD_T() do { if (XO_NEEDS_OPL) opl = P_E_REF; } while (0)
#define XO_POOL_STORE_T() do { if (XO_NEEDS_OPL) P_E_REF = opl; } while (0)
#define XO_POOL_TRACE_TRIP() do { } while (0)
#define XO_POOL_OPL (XO_NEEDS_OPL)
#endif
	// slots of this warp
	P_ST[lane] = (unsigned char)PS_EMPTY;
	P_ST[lane + 32u] = (unsigned char)PS_EMPTY;
	for (u32 slot = lane; slot < S; slot += 32u) P4_STORE(A, 0.0f, 0.0f, 0.0f, 0.0f);
#if XO_POOL_QUEUES
	// One ring of slot numbers per class (every slot is in exactly one ring: 64 entries
	// each never overflow); heads and counts are warp-uniform registers.  A phase pops up
	// to 32 slots from the ring(s) of its class and pushes every slot it held onto the ring
	// of its new class (ballot + rank + one byte store per class) - instead of a census of
	// all 64 slot states and a gather by rank every round.
	unsigned char *const Q_I = P_Q, *const Q_D = P_Q + 64, *const Q_W = P_Q + 128,
		*const Q_B = P_Q + 192, *const Q_E = P_Q + 256;
	Q_E[lane] = (unsigned char)lane;
	Q_E[lane + 32u] = (unsigned char)(lane + 32u);
	u32 hI = 0, hD = 0, hW = 0, hB = 0, hE = 0;
	u32 nI = 0, nD = 0, nW = 0, nB = 0, nE = S;
#define XO_Q_PUSH(Q_, h_, n_, pred_) do { \
		const bool p_ = (pred_); \
		const u32 m_ = __ballot_sync(0xffffffffu, p_); \
		if (p_) (Q_)[((h_) + (n_) + (u32)__popc(m_ & lanemask_lt)) & 63u] = (unsigned char)slot; \
		(n_) += (u32)__popc(m_); \
	} while (0)
#define XO_Q_POP(Q_, h_, n_, first_, take_) do { \
		if (lane >= (first_) && lane < (first_) + (take_)) slot = (Q_)[((h_) + lane - (first_)) & 63u]; \
		(h_) = ((h_) + (take_)) & 63u; (n_) -= (take_); \
	} while (0)
#else
#define XO_Q_PUSH(Q_, h_, n_, pred_) do { } while (0)
#endif
	__syncwarp();
	bool dry = false;               // warp-uniform: the packet budget is exhausted

	for (;;) {
#if XO_POOL_QUEUES
		u32 phase;
		if (!dry && nE >= XO_POOL_LAUNCH) phase = PH_LAUNCH;
		else if (nE == S) break;
		else if (nB >= XO_POOL_BND || (nB >= nI && nB >= nD + nW)) phase = PH_BOUNDARY;
		else phase = (nI >= nD + nW) ? PH_INTERACT : PH_WALK;
		u32 slot = 0;
		bool act;
		if (phase == PH_INTERACT) {
			const u32 n = nI < 32u ? nI : 32u;
			act = lane < n;
			XO_Q_POP(Q_I, hI, nI, 0u, n);
		} else if (phase == PH_WALK) {
			// (the slots that still need their walk set-up first)
			const u32 nd = nD < 32u ? nD : 32u;
			const u32 nw = nW < 32u - nd ? nW : 32u - nd;
			act = lane < nd + nw;
			XO_Q_POP(Q_D, hD, nD, 0u, nd);
			XO_Q_POP(Q_W, hW, nW, nd, nw);
		} else if (phase == PH_BOUNDARY) {
			const u32 n = nB < 32u ? nB : 32u;
			act = lane < n;
			XO_Q_POP(Q_B, hB, nB, 0u, n);
		} else {
			const u32 n = nE < 32u ? nE : 32u;
			act = lane < n;
			XO_Q_POP(Q_E, hE, nE, 0u, n);
		}
#else
		// ---- census of the slot states: class counts in the bytes of one word ------------
		u32 nI, nD, nW, nB, nE;
		const u32 s0 = P_ST[lane], s1 = P_ST[lane + 32u];
		{
			const u32 w0 = s0 ? (1u << (s0 & 24u)) : 0u;
			const u32 w1 = s1 ? (1u << (s1 & 24u)) : 0u;
			const u32 cen = __reduce_add_sync(0xffffffffu, w0 + w1);
			nI = cen & 0xffu; nD = (cen >> 8) & 0xffu; nW = (cen >> 16) & 0xffu; nB = cen >> 24;
			nE = S - (nI + nD + nW + nB);
		}
		u32 phase;
		if (!dry && nE >= XO_POOL_LAUNCH) phase = PH_LAUNCH;
		else if (nE == S) break;
		else if (nB >= XO_POOL_BND || (nB >= nI && nB >= nD + nW)) phase = PH_BOUNDARY;
		else phase = (nI >= nD + nW) ? PH_INTERACT : PH_WALK;

		// ---- gather up to 32 slots of the phase's class(es), first class first -----------
		u32 slot = 0;
		bool act;
		{
			// class key of a slot: 0 / 8 / 16 / 24 as in the census, 32 = EMPTY; a phase takes
			// its first class first (WALK: the slots that still need their walk set-up)
			const u32 k0 = s0 ? (s0 & 24u) : 32u, k1 = s1 ? (s1 & 24u) : 32u;
			const u32 key_f = (phase == PH_INTERACT) ? 0u : ((phase == PH_WALK) ? 8u :
				((phase == PH_BOUNDARY) ? 24u : 32u));
			const bool f0 = (k0 == key_f), f1 = (k1 == key_f);
			const u32 mf0 = __ballot_sync(0xffffffffu, f0), mf1 = __ballot_sync(0xffffffffu, f1);
			u32 n = (u32)__popc(mf0);
			if (f0) P_IDX[__popc(mf0 & lanemask_lt)] = (unsigned char)lane;
			u32 r = n + (u32)__popc(mf1 & lanemask_lt);
			if (f1 && r < 32u) P_IDX[r] = (unsigned char)(lane + 32u);
			n += (u32)__popc(mf1);
			if (phase == PH_WALK) {
				const bool g0 = (k0 == 16u), g1 = (k1 == 16u);
				const u32 mg0 = __ballot_sync(0xffffffffu, g0), mg1 = __ballot_sync(0xffffffffu, g1);
				r = n + (u32)__popc(mg0 & lanemask_lt);
				if (g0 && r < 32u) P_IDX[r] = (unsigned char)lane;
				n += (u32)__popc(mg0);
				r = n + (u32)__popc(mg1 & lanemask_lt);
				if (g1 && r < 32u) P_IDX[r] = (unsigned char)(lane + 32u);
				n += (u32)__popc(mg1);
			}
			__syncwarp();
			act = lane < n;         // (n may exceed 32: the rest waits for the next round)
			if (act) slot = P_IDX[lane];
			__syncwarp();
		}
#endif

		if (phase == PH_INTERACT) {
			// ======== interaction + next free path (mcvox.template.c:925-981, 667-700) ========
			u32 st = PS_EMPTY, vlo = vbase_lo, mat = 0, dcur = 1;
			P3 pos = { 0.0f, 0.0f, 0.0f }, dir = { 0.0f, 0.0f, 1.0f };
			float weight = 0.0f, t_s = 0.0f, opl = 0.0f;
			(void)opl;
#if XO_USE_RMAX
			float t_rmax = 0.0f;
#endif
			if (act) {
				st = P_ST[slot];
				const float4 a = P4_LOAD(A), b = P4_LOAD(B);
				pos.x = a.x; pos.y = a.y; pos.z = a.z; weight = a.w;
				dir.x = b.x; dir.y = b.y; dir.z = b.z; t_s = b.w;
				vlo = __float_as_uint(PW_LOAD(C));
				const u32 misc = __float_as_uint(PW_LOAD(D));
				mat = misc & 0xffu; dcur = (misc >> 8) & 0xffu;
				XO_POOL_LOAD_T();
				if (st == PS_NEW) { const u32 cell = XO_VOXEL(vlo); mat = cell & 0xffu; dcur = cell >> 8; }
			}
			VoxHot c_hot;
			XoPf::Fast c_pf;
			XO_LOAD_MAT(mat);
			for (;;) {
				if (st == PS_SCAT || st == PS_FAR) {
					bool done = false;
					++iterations;
					pos.x = fmaf(dir.x, t_s, pos.x);
					pos.y = fmaf(dir.y, t_s, pos.y);
					pos.z = fmaf(dir.z, t_s, pos.z);
					if (XO_POOL_OPL) opl = fmaf(c_hot.n, t_s, opl);
					u32 cell_far = 0x100u | XO_VOX_SENTINEL;      // (no look-up: keeps mat, clearance 1 unused)
					const bool was_far = (st == PS_FAR);
					if (st == PS_FAR) {
						// the flight skipped the walk: voxel of the interaction point from the
						// position, loop trips of the reference = faces crossed on the way
						const u32 idx0 = (vlo - vbase_lo) >> 1;
						i32 ix = __float2int_rd((pos.x - cfg.top_left.x)*inv_sx);
						i32 iy = __float2int_rd((pos.y - cfg.top_left.y)*inv_sy);
						i32 iz = __float2int_rd((pos.z - cfg.top_left.z)*inv_sz);
						ix = __vimin_s32_relu(ix, cfg.nx - 1);      // clip to [0, n - 1]: one VIMNMX.RELU
						iy = __vimin_s32_relu(iy, cfg.ny - 1);
						iz = __vimin_s32_relu(iz, cfg.nz - 1);
						iterations += (u32)(abs(ix - ((i32)(idx0 & vox_mx) - 2)) +
							abs(iy - ((i32)((idx0 >> vox_bx) & vox_my) - 2)) +
							abs(iz - ((i32)(idx0 >> vox_bxy) - 2)));
						vlo = XO_PACK_VOXEL(ix, iy, iz);
						// (the cell is looked up here and read after the deposit and the
						// scattering below: the L2 round trip overlaps them)
						cell_far = XO_VOXEL(vlo);
					}
#if XO_METHOD == 1
					if (rng.next() < c_hot.absorb) {
						float deposit = weight;
						done = true;
						weight = 0.0f;
						flags |= EV_ABSORPTION;
						if (XoFluence::active) fluence.deposit(acc, window, pos, deposit, c_hot.mua, opl);
					} else {
						pf_scatter(c_pf, rng, lut, dir);
						flags |= EV_SCATTERING;
					}
#else
					{
						float deposit = weight*c_hot.absorb;
						weight -= deposit;
#if XO_FLU_VOXGRID
						{   // the fluence grid is the voxel grid: the cell is the voxel of the packet
							const u32 idx = (vlo - vbase_lo) >> 1;
							fluence.deposit_cell(acc, (idx & vox_mx) - 2u, ((idx >> vox_bx) & vox_my) - 2u,
								(idx >> vox_bxy) - 2u, fluence_weight(deposit, c_hot.mua, fluence.k));
						}
#else
						if (XoFluence::active) fluence.deposit(acc, window, pos, deposit, c_hot.mua, opl);
#endif
					}
					pf_scatter(c_pf, rng, lut, dir);
					flags |= EV_ABSORPTION | EV_SCATTERING;
					if (weight < XO_WEIGHT_MIN) {
#if XO_USE_LOTTERY
						if (rng.next_raw() > XO_LOTTERY_CHANCE*4294967296.0f) done = true;
						else weight *= (1.0f/XO_LOTTERY_CHANCE);
#else
						done = true;
#endif
					}
#endif
					if (was_far) {
						dcur = cell_far >> 8;
						// (a rounding of the end point across a face of the clearance box can
						// land in another material: adopted after this interaction)
						const u32 mat_here = cell_far & 0xffu;
						if (mat_here != XO_VOX_SENTINEL && mat_here != mat) { mat = mat_here; XO_LOAD_MAT(mat); }
					}
					if (weight <= 0.0f) { done = true; flags |= EV_ESCAPED; }
					XO_POOL_RMAX_TEST();
					XO_POOL_TRACE_TRIP();
					st = done ? PS_EMPTY : PS_RAY;
				}
				if (st == PS_RAY || st == PS_NEW) {
					t_s = fminf((FastMath::lg2(rng.next_raw()) - 32.0f)*c_hot.step_k, XO_FLT_MAX);
					// extent of the flight in voxels along its longest axis against the clearance
					const float ext = t_s*fmaxf(fabsf(dir.x)*inv_sx, fmaxf(fabsf(dir.y)*inv_sy, fabsf(dir.z)*inv_sz));
					bool far = XO_POOL_CLEARANCE && ext < (float)dcur - 1.0f;
#if XO_USE_RMAX
					{   // parameter at which the ray leaves the rmax sphere around the source
						const float ex = pos.x - src_pos.x, ey = pos.y - src_pos.y, ez = pos.z - src_pos.z;
						const float b = ex*dir.x + ey*dir.y + ez*dir.z;
						const float c = ex*ex + ey*ey + ez*ez - rmax2;
						t_rmax = FastMath::sqrt(fmaxf(fmaf(b, b, -c), 0.0f)) - b;
						far = far && t_s <= t_rmax;
					}
#endif
					st = far ? PS_FAR : PS_DDA;
				}
				if ((u32)__popc(__ballot_sync(0xffffffffu, st == PS_FAR)) < XO_POOL_THR_I) break;
			}
			if (act) {
				P4_STORE(A, pos.x, pos.y, pos.z, weight);
				P4_STORE(B, dir.x, dir.y, dir.z, t_s);
				PW_STORE(C, __uint_as_float(vlo));
				PW_STORE(D, __uint_as_float(mat | (dcur << 8)));
				XO_POOL_STORE_T();
#if XO_USE_RMAX
				P_R_REF = t_rmax;
#endif
				P_ST[slot] = (unsigned char)st;
			}
			XO_Q_PUSH(Q_I, hI, nI, act && st == PS_FAR);
			XO_Q_PUSH(Q_D, hD, nD, act && st == PS_DDA);
			XO_Q_PUSH(Q_E, hE, nE, act && st == PS_EMPTY);
		} else if (phase == PH_WALK) {
			// ======== walk set-up (DDA slots) and voxel walk =====================================
			u32 st = PS_EMPTY, vlo = vbase_lo, mat = 0, dcur = 1, sg = 7u;
			i32 last_d = 0;
			float tmx = XO_INF, tmy = XO_INF, tmz = XO_INF, tdx = 0.0f, tdy = 0.0f, tdz = 0.0f;
			float t_s = 0.0f, t_evt = 0.0f;
			bool was_dda = false;
#if XO_USE_RMAX
			const float t_rmax = act ? P_R_REF : XO_INF;
#endif
#if XO_TRACE == XO_TRACE_ALL
			// (every crossing is an event: the ray itself, the weight and the trace quad)
			P3 pos = { 0.0f, 0.0f, 0.0f }, dir = { 0.0f, 0.0f, 1.0f };
			float weight = 0.0f, opl = 0.0f;
#endif
			if (act) {
				st = P_ST[slot];
				t_s = PW_LOAD(B);
#if XO_TRACE == XO_TRACE_ALL
				{
					const float4 a_ = P4_LOAD(A), b_ = P4_LOAD(B);
					pos.x = a_.x; pos.y = a_.y; pos.z = a_.z; weight = a_.w;
					dir.x = b_.x; dir.y = b_.y; dir.z = b_.z;
					XO_POOL_LOAD_T();
				}
#endif
				was_dda = (st == PS_DDA);
				if (st == PS_DDA) {
					const float4 a = P4_LOAD(A), b = P4_LOAD(B);
					vlo = __float_as_uint(PW_LOAD(C));
					const u32 misc = __float_as_uint(PW_LOAD(D));
					mat = misc & 0xffu; dcur = (misc >> 8) & 0xffu;
					const float rx = FastMath::rcp_approx(b.x), ry = FastMath::rcp_approx(b.y),
						rz = FastMath::rcp_approx(b.z);
					const bool fx = b.x >= 0.0f, fy = b.y >= 0.0f, fz = b.z >= 0.0f;
					sg = (fx ? 1u : 0u) | (fy ? 2u : 0u) | (fz ? 4u : 0u);
					// exit faces of the current voxel (mcvox.template.c:173-196); the packed
					// index holds coordinate + 2
					const u32 idx = (vlo - vbase_lo) >> 1;
					const float facex = fmaf((float)((i32)(idx & vox_mx) - (fx ? 1 : 2)), cfg.size.x, cfg.top_left.x);
					const float facey = fmaf((float)((i32)((idx >> vox_bx) & vox_my) - (fy ? 1 : 2)), cfg.size.y, cfg.top_left.y);
					const float facez = fmaf((float)((i32)(idx >> vox_bxy) - (fz ? 1 : 2)), cfg.size.z, cfg.top_left.z);
					tmx = (b.x != 0.0f) ? fmaxf((facex - a.x)*rx, 0.0f) : XO_INF;
					tmy = (b.y != 0.0f) ? fmaxf((facey - a.y)*ry, 0.0f) : XO_INF;
					tmz = (b.z != 0.0f) ? fmaxf((facez - a.z)*rz, 0.0f) : XO_INF;
					tdx = cfg.size.x*fabsf(rx);
					tdy = cfg.size.y*fabsf(ry);
					tdz = cfg.size.z*fabsf(rz);
					st = PS_RUN;
				} else {
					const float4 c = P4_LOAD(C), d = P4_LOAD(D);
					tmx = c.x; tmy = c.y; tmz = c.z; vlo = __float_as_uint(c.w);
					tdx = d.x; tdy = d.y; tdz = d.z;
					const u32 misc = __float_as_uint(d.w);
					mat = misc & 0xffu; dcur = (misc >> 8) & 0xffu; sg = (misc >> 18) & 7u;
				}
			}
			// address increment per crossing on each axis (pinned in registers over the walk)
			i32 stx = (sg & 1u) ? 2 : -2;
			i32 sty = ((sg & 2u) ? 2 : -2) << vox_bx;
			i32 stz = ((sg & 4u) ? 2 : -2) << vox_bxy;
			asm volatile("" : "+r"(stx), "+r"(sty), "+r"(stz));
			float step_k = sh_fast[mat].hot.step_k;
			(void)step_k;
#if XO_TRACE == XO_TRACE_ALL
			float n_mat = sh_fast[mat].hot.n;
#define XO_TRACE_CROSSING(tmin_) do { \
				P3 pc_ = { fmaf(dir.x, tmin_, pos.x), fmaf(dir.y, tmin_, pos.y), fmaf(dir.z, tmin_, pos.z) }; \
				if (trace_event(tcfg, float_buffer, packet, trace_count, \
						flags | EV_BOUNDARY_HIT | EV_REFRACTION, pc_, dir, weight, \
						fmaf(n_mat, tmin_, opl))) ++trace_count; \
				flags = 0u; \
			} while (0)
#else
#define XO_TRACE_CROSSING(tmin_) do { } while (0)
#endif
			// Two crossings per trip; the lookup of the second one is issued before the first
			// has returned, and committed only if the first one stayed inside the material.
			for (;;) {
				if (st == PS_RUN) {
					const float tmin_a = fminf(tmx, fminf(tmy, tmz));
					if (!(tmin_a < t_s)) {
						st = PS_SCAT;
					} else {
						++iterations;
						const bool px = (tmx == tmin_a);
						const bool py = !px && (tmy == tmin_a);
						const bool pz = !px && !py;
						if (px) tmx += tdx;
						if (py) tmy += tdy;
						if (pz) tmz += tdz;
						const i32 d_a = px ? stx : (py ? sty : stz);
						vlo += (u32)d_a;
						const u32 cell_a = XO_VOXEL(vlo);
						u32 m_a = cell_a & 0xffu;
						// second crossing, speculative
						const float tmin_b = fminf(tmx, fminf(tmy, tmz));
						const bool ok_b = tmin_b < t_s;
						const bool qx = (tmx == tmin_b);
						const bool qy = !qx && (tmy == tmin_b);
						const bool qz = !qx && !qy;
						const i32 d_b = qx ? stx : (qy ? sty : stz);
						const u32 vlo_b = vlo + (u32)d_b;
						u32 cell_b = 0x100u | mat;
						if (ok_b) cell_b = XO_VOXEL(vlo_b);
						u32 m_b = cell_b & 0xffu;
						dcur = cell_a >> 8;
#if XO_USE_RMAX
						// first face beyond rmax: handled as an event
						if (tmin_a > t_rmax) m_a = ~0u;
						if (tmin_b > t_rmax) m_b = ~0u;
#endif
#if XO_VOX_SAME_N
						// Equal refractive indices everywhere: a face between two materials is no
						// interface event, only the attenuation changes; the remaining optical depth
						// of the flight is still Exp(1) distributed, so the free path is rescaled to
						// the new material (the grid face stays an event).
#define XO_MATERIAL_CHANGE(m_, tmin_) do { \
							XO_TRACE_CROSSING(tmin_); \
							const float k_new_ = sh_fast[m_].hot.step_k; \
							t_s = fmaf(t_s - (tmin_), k_new_*FastMath::rcp_approx(step_k), (tmin_)); \
							step_k = k_new_; \
							mat = (m_); \
						} while (0)
#define XO_IS_PLAIN_CHANGE(m_) ((m_) < XO_VOX_SENTINEL && step_k > -XO_FLT_MAX && t_s < XO_FLT_MAX)
#else
#define XO_MATERIAL_CHANGE(m_, tmin_) do { } while (0)
#define XO_IS_PLAIN_CHANGE(m_) false
#endif
						if (m_a != mat) {
							if (XO_IS_PLAIN_CHANGE(m_a)) {
								XO_MATERIAL_CHANGE(m_a, tmin_a);
							} else {
								st = PS_BND; t_evt = tmin_a; last_d = d_a;
							}
						} else if (!ok_b) {
							XO_TRACE_CROSSING(tmin_a);      // the reference records every loop trip
							st = PS_SCAT;
						} else {
							XO_TRACE_CROSSING(tmin_a);
							++iterations;
							if (qx) tmx += tdx;
							if (qy) tmy += tdy;
							if (qz) tmz += tdz;
							vlo = vlo_b;
							dcur = cell_b >> 8;
							if (m_b != mat) {
								if (XO_IS_PLAIN_CHANGE(m_b)) {
									XO_MATERIAL_CHANGE(m_b, tmin_b);
								} else {
									st = PS_BND; t_evt = tmin_b; last_d = d_b;
								}
							} else {
								XO_TRACE_CROSSING(tmin_b);
							}
						}
#undef XO_MATERIAL_CHANGE
#undef XO_IS_PLAIN_CHANGE
					}
				}
				if ((u32)__popc(__ballot_sync(0xffffffffu, st == PS_RUN)) < XO_POOL_THR_W) break;
			}
			if (act) {
				const u32 axis = (last_d == stx) ? 0u : ((last_d == sty) ? 1u : 2u);
				PW_STORE(B, (st == PS_BND) ? t_evt : t_s);
				P4_STORE(C, tmx, tmy, tmz, __uint_as_float(vlo));
				// (the per-voxel increments of a ray never change: stored by its set-up only)
				if (was_dda) P3_STORE(D, tdx, tdy, tdz);
				PW_STORE(D, __uint_as_float(mat | (dcur << 8) | (axis << 16) | (sg << 18)));
#if XO_TRACE == XO_TRACE_ALL
				XO_POOL_STORE_T();
#endif
				P_ST[slot] = (unsigned char)st;
			}
			XO_Q_PUSH(Q_W, hW, nW, act && st == PS_RUN);
			XO_Q_PUSH(Q_I, hI, nI, act && st == PS_SCAT);
			XO_Q_PUSH(Q_B, hB, nB, act && st == PS_BND);
#undef XO_TRACE_CROSSING
		} else if (phase == PH_BOUNDARY) {
			// ======== a face between materials of different refractive index, or the face of
			// the grid (mcvox.template.c:275-388); the walk already moved the voxel address
			// across the face, along `axis` =========================================================
			bool survived = false;
			(void)survived;
			if (act) {
				const float4 a = P4_LOAD(A), b = P4_LOAD(B);
				P3 pos = { a.x, a.y, a.z }, dir = { b.x, b.y, b.z };
				float weight = a.w, opl = 0.0f;
				const float t_evt = b.w;
				u32 vlo = __float_as_uint(PW_LOAD(C));
				const u32 misc = __float_as_uint(PW_LOAD(D));
				u32 mat = misc & 0xffu;
				const u32 axis = (misc >> 16) & 3u, sg = (misc >> 18) & 7u;
				XO_POOL_LOAD_T();
				const VoxHot hot = sh_fast[mat].hot;
				bool done = false;
				pos.x = fmaf(dir.x, t_evt, pos.x);
				pos.y = fmaf(dir.y, t_evt, pos.y);
				pos.z = fmaf(dir.z, t_evt, pos.z);
				if (XO_POOL_OPL) opl = fmaf(hot.n, t_evt, opl);
				const u32 entered = XO_VOXEL(vlo) & 0xffu;
				const bool escaping = (entered == XO_VOX_SENTINEL);
				const u32 next_mat = escaping ? 0u : entered;
				const float n1 = hot.n, n2 = sh_fast[next_mat].hot.n;
				bool through = true;
				if (n1 != n2) {
					const float n12 = n1*FastMath::rcp_approx(n2);
					const float n21 = n2*FastMath::rcp_approx(n1);
					const float cc = (n1 > n2) ? FastMath::sqrt(fmaxf(fmaf(-n21, n21, 1.0f), 0.0f)) : 0.0f;
					if (axis == 0u) through = fresnel_axis_fast(n12, cc, dir.x, dir.y, dir.z, rng);
					else if (axis == 1u) through = fresnel_axis_fast(n12, cc, dir.y, dir.x, dir.z, rng);
					else through = fresnel_axis_fast(n12, cc, dir.z, dir.x, dir.y, rng);
				}
				flags |= EV_BOUNDARY_HIT | (through ? EV_REFRACTION : EV_REFLECTION);
				if (through) {
					if (escaping) {
						const i32 iz = (i32)(((vlo - vbase_lo) >> 1) >> vox_bxy) - 2;
						if (iz < 0) {
							if (XoDetTop::active) detectors.top.deposit(acc, pos, dir, weight, opl);
						} else if (iz >= cfg.nz) {
							if (XoDetBottom::active) detectors.bottom.deposit(acc, pos, dir, weight, opl);
						}
						done = true;
					} else {
						mat = next_mat;
					}
				} else {
					// reflected: back into the voxel the packet came from
					const i32 d = (axis == 0u) ? ((sg & 1u) ? 2 : -2) :
						((axis == 1u) ? (((sg & 2u) ? 2 : -2) << vox_bx) : (((sg & 4u) ? 2 : -2) << vox_bxy));
					vlo -= (u32)d;
				}
				if (weight <= 0.0f) { done = true; flags |= EV_ESCAPED; }
				XO_POOL_RMAX_TEST();
				XO_POOL_TRACE_TRIP();
				P4_STORE(A, pos.x, pos.y, pos.z, weight);
				P4_STORE(B, dir.x, dir.y, dir.z, 0.0f);
				PW_STORE(C, __uint_as_float(vlo));
				PW_STORE(D, __uint_as_float(mat | (1u << 8)));    // on a face: clearance 1
				XO_POOL_STORE_T();
				P_ST[slot] = (unsigned char)(done ? PS_EMPTY : PS_RAY);
				survived = !done;
			}
			XO_Q_PUSH(Q_I, hI, nI, act && survived);
			XO_Q_PUSH(Q_E, hE, nE, act && !survived);
		} else {
			// ======== new packets into the EMPTY slots =============================================
			const u32 want = (u32)__popc(__ballot_sync(0xffffffffu, act));
			u32 base = 0;
			if (lane == 0u) base = atomicAdd(num_packets_done, want);
			base = __shfl_sync(0xffffffffu, base, 0);
			const u32 n_new = base < num_packets ?
				(num_packets - base < want ? num_packets - base : want) : 0u;
			dry = n_new < want;
			if (lane < n_new) {
				const float4 prev = P4_LOAD(A);     // (last position of the packet that died here)
				P3 prev_pos = { prev.x, prev.y, prev.z };
				Launch L_;
				source.launch(rng, ctx, prev_pos, L_);
				if (XoDetSpecular::active && L_.spec_weight >= 0.0f)
					detectors.specular.deposit(acc, L_.pos, L_.spec_dir, L_.spec_weight, 0.0f);
				// voxel under the launch point (mcvox.template.c:206-220), kept inside the grid
				i32 ix, iy, iz;
				ctx.position_to_voxel(L_.pos, &ix, &iy, &iz);
				ix = clipi(ix, 0, cfg.nx - 1);
				iy = clipi(iy, 0, cfg.ny - 1);
				iz = clipi(iz, 0, cfg.nz - 1);
				P4_STORE(A, L_.pos.x, L_.pos.y, L_.pos.z, L_.weight);
				P4_STORE(B, L_.dir.x, L_.dir.y, L_.dir.z, 0.0f);
				PW_STORE(C, __uint_as_float(XO_PACK_VOXEL(ix, iy, iz)));
				PW_STORE(D, __uint_as_float(0u));
				{
					float opl = 0.0f;
					(void)opl;
					packet = base + lane;
					trace_count = 0u;
					flags = EV_LAUNCH;
#if XO_TRACE & XO_TRACE_START
					if (trace_event(tcfg, float_buffer, packet, 0u, EV_LAUNCH,
							L_.pos, L_.dir, L_.weight, 0.0f)) trace_count = 1u;
#endif
					XO_POOL_STORE_T();
				}
				P_ST[slot] = (unsigned char)PS_NEW;
				started = true;
			}
			// (slots taken off the ring of empty slots: launched, or back when the budget ran out)
			XO_Q_PUSH(Q_I, hI, nI, act && lane < n_new);
			XO_Q_PUSH(Q_E, hE, nE, act && lane >= n_new);
		}
		__syncwarp();
	}
#undef XO_Q_PUSH
#if XO_POOL_QUEUES
#undef XO_Q_POP
#endif
#undef XO_LOAD_MAT
#undef XO_VOXEL
#undef XO_PACK_VOXEL
#undef XO_POOL_CLEARANCE
#undef XO_POOL_LOAD_T
#undef XO_POOL_STORE_T
#undef XO_POOL_TRACE_TRIP
#undef XO_POOL_OPL
#undef P4_LOAD
#undef P4_STORE
#undef P3_STORE
#undef PW_LOAD
#undef PW_STORE
#undef P_E_REF
#undef P_R_REF
#undef XO_POOL_RMAX_TEST
	// every lane drew from its stream: all states go back
	rng_state_x[gid] = rng.state();
}
