// mcvox_dda_loop.cuh -- throughput loop of the voxel kernel (body fragment,
// included inside McKernel of mcvox_kernel.cuh when XO_VOX_DDA).
//
// Same physics as one iteration of the reference loop
// (xopto/mcvox/kernel/mcvox.template.c:667-1014), re-expressed for the SM:
//
//   * The reference draws a fresh exponential step in every iteration and
//     divides three face distances by the direction (3 divisions + 1 log per
//     voxel crossing).  Here a straight flight is ONE ray: origin `pos`, the
//     free path `t_s` drawn once, and an incremental voxel walk (Amanatides &
//     Woo): per axis the ray parameter of the next face `tm` and the parameter
//     increment per voxel `td`.  A crossing is min3 + one predicated add per
//     axis + the material lookup of the entered voxel.  Statistically identical:
//     the exponential distribution is memoryless, so "fresh step after every
//     crossing inside one material" == "one step for the whole flight"; a fresh
//     step IS drawn whenever the material changes or the packet is reflected
//     (the reference's behaviour).  The iteration counter
//     counts crossings + interactions, i.e. the reference's loop trips.
//   * 85 % of the trips are crossings (~25 instructions), 15 % interactions
//     (~110: deposit, phase function, rotation, new ray).  Executed as one
//     divergent loop a warp would run both paths every trip.  Instead each lane
//     is a small state machine and the warp runs the crossing step for the lanes
//     in RUN state until `refill` lanes wait for something else; then the
//     waiting lanes execute their interaction / interface / relaunch jointly.
//   * the voxel -> material map is read from a compact copy (uint8, one voxel of
//     sentinel padding on every side, power-of-two strides; mcvox/mc.py): the
//     linear index is the packed coordinate triple, one add per crossing moves
//     it, leaving the grid is a "material change" to the sentinel, and the few
//     hundred KB around the beam stay L1-resident (the int32 map of the
//     reference layout is 4x larger and every lookup went to L2).
//   * every cell of the compact map also carries the CLEARANCE of its voxel: the
//     chessboard distance (in voxels, capped at 255) to the nearest voxel of another
//     material or outside the grid.  A flight whose extent along every axis stays
//     below clearance - 1 voxels cannot leave the material, whatever the start point
//     inside the voxel: it skips the walk altogether (no DDA set-up, no per-voxel
//     lookups) and goes straight to its interaction; the loop-trip count of the
//     reference is restored as the Manhattan distance between the start and the end
//     voxel (a straight ray is monotone on every axis).  Half of the voxels of the
//     201^3 skin + vessel model have a clearance above 10 voxels, the mean free path
//     is 5.4.  Disabled for full traces (one event per crossing is the product there).
//   * new packets come from a per-warp launch queue filled by all 32 lanes
//     together (as in mcml_kernel.cuh).
//   * rmax: the reference tests |pos - source| > rmax after every trip.  A ray
//     leaves the (convex) sphere once, at parameter `t_rmax` computed per ray,
//     so the test of a crossing is one compare (compiled out by the host when
//     the voxel box lies inside the sphere).
{
	enum : u32 { ST_RUN = 0, ST_SCAT = 1, ST_BND = 2, ST_SETUP = 3, ST_DEAD = 4, ST_DRY = 5, ST_FAR = 6 };
#define XO_VOX_CLEARANCE (XO_TRACE != XO_TRACE_ALL)
	const u32 lane = threadIdx.x & 31u;
	const u32 lanemask_lt = (1u << lane) - 1u;
	const u32 vox_bxy = vox_bx + vox_by;
	const u32 vox_mx = (1u << vox_bx) - 1u, vox_my = (1u << vox_by) - 1u;
	const XoTraceCfg &tcfg = *reinterpret_cast<const XoTraceCfg *>(&trace);
	(void)tcfg; (void)chunk; (void)nthreads;

	u32 state = ST_DEAD;
	u32 n_dry = 0, q_count = 0;     // warp-uniform
	bool q_dry = false;             // warp-uniform: the packet budget is exhausted
	// packet
	P3 pos = { 0.0f, 0.0f, 0.0f }, dir = { 0.0f, 0.0f, 1.0f };
	float weight = 0.0f, opl = 0.0f;
	u32 packet = 0, trace_count = 0, flags = 0;
	(void)opl; (void)packet; (void)trace_count; (void)flags;
	// ray: voxel walk state
	// The walk keeps the low address word of the current voxel in the compact
	// map (16-bit cells: material | clearance << 8),
	// vlo = lo32(voxels8) + 2*packed index, packed index = (x+2) | (y+2) << bx | (z+2) << bxy
	// (two voxels of padding: the speculative second crossing of a trip may look
	// one voxel beyond the sentinel layer);
	// the host guarantees that the map does not straddle a 4 GB boundary, so a
	// crossing is one 32-bit add and the high word is a constant.
	const u32 vbase_lo = (u32)reinterpret_cast<u64>(voxels8);
	const u32 vbase_hi = (u32)(reinterpret_cast<u64>(voxels8) >> 32);
	u32 vlo = vbase_lo;
	u32 dcur = 1;                   // clearance of the current voxel
	i32 stx = 2, sty = 2, stz = 2;  // address increment per crossing on each axis
	const float inv_sx = 1.0f/cfg.size.x, inv_sy = 1.0f/cfg.size.y, inv_sz = 1.0f/cfg.size.z;
	(void)dcur; (void)inv_sx; (void)inv_sy; (void)inv_sz;
	i32 last_d = 0;                 // increment of the last crossing (names its axis)
	u32 mat = 0;
	float tmx = 0.0f, tmy = 0.0f, tmz = 0.0f, tdx = 0.0f, tdy = 0.0f, tdz = 0.0f;
	float t_s = 0.0f, t_evt = 0.0f;
#if XO_USE_RMAX
	float t_rmax = 0.0f;
#endif
	// constants of the current material (registers; reloaded on material change)
	VoxHot c_hot = { 0.0f, 0.0f, 0.0f, 1.0f };
	XoPf::Fast c_pf;
#define XO_LOAD_MAT(idx) do { const VoxFastMat &F_ = sh_fast[idx]; c_hot = F_.hot; c_pf = F_.pf.v; } while (0)
#if XO_ANISO
	// anisotropic materials: step / absorption constants of a ray = the tensors projected
	// on its direction, derived when the ray is set up
#define XO_DIR_CONSTS() do { \
		const VoxMaterial &M_ = sh_mat[mat]; \
		const float mut_ = tensor_project(M_.mut_t, dir), mua_ = tensor_project(M_.mua_t, dir); \
		const float inv_ = (mut_ != 0.0f) ? FastMath::rcp_approx(mut_) : XO_INF; \
		c_hot.step_k = -0.6931471805599453f*inv_; \
		c_hot.absorb = (mua_ != 0.0f) ? mua_*inv_ : 0.0f; \
		c_hot.mua = mua_; \
	} while (0)
#else
#define XO_DIR_CONSTS() do { } while (0)
#endif
#define XO_VOXEL(lo) ((u32)__ldg(reinterpret_cast<const unsigned short *>(((u64)vbase_hi << 32) | (u64)(lo))))
#define XO_PACK_VOXEL(ix_, iy_, iz_) (vbase_lo + 2u*((u32)((ix_) + 2) | ((u32)((iy_) + 2) << vox_bx) | ((u32)((iz_) + 2) << vox_bxy)))
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
#define XO_END_TRIP() do { \
		if (weight <= 0.0f) { done = true; flags |= EV_ESCAPED; } \
		XO_RMAX_TEST(); \
		XO_TRACE_TRIP(); \
		flags = 0; \
		state = done ? ST_DEAD : ST_SETUP; \
	} while (0)

	for (;;) {
		// ---- hand new packets to the lanes that need one -------------------------
		const u32 dead_mask = __ballot_sync(0xffffffffu, state == ST_DEAD);
		if (dead_mask != 0u) {
			if (q_count == 0u && !q_dry) {
				// queue empty: all 32 lanes launch one packet each into the queue
				u32 base = 0;
				if (lane == 0u) base = atomicAdd(num_packets_done, 32u);
				base = __shfl_sync(0xffffffffu, base, 0);
				const u32 n_new = base < num_packets ?
					(num_packets - base < 32u ? num_packets - base : 32u) : 0u;
				q_dry = n_new < 32u;
				if (lane < n_new) {
					Launch L_;
					source.launch(rng, ctx, pos, L_);
					if (XoDetSpecular::active && L_.spec_weight >= 0.0f)
						detectors.specular.deposit(acc, L_.pos, L_.spec_dir, L_.spec_weight, 0.0f);
					u32 tc = 0;
					if (XO_TRACE & XO_TRACE_START) {
						if (trace_event(tcfg, float_buffer, base + lane, 0u, EV_LAUNCH,
								L_.pos, L_.dir, L_.weight, 0.0f)) tc = 1u;
					}
					q_a[lane] = make_float4(L_.pos.x, L_.pos.y, L_.pos.z, L_.weight);
					q_b[lane] = make_float4(L_.dir.x, L_.dir.y, L_.dir.z, __uint_as_float(base + lane));
					q_l[lane] = tc;
					{   // voxel under the launch point (mcvox.template.c:206-220), kept inside
						// the grid; found here by the whole warp, not by the 1-2 lanes of a pop
						i32 ix, iy, iz;
						ctx.position_to_voxel(L_.pos, &ix, &iy, &iz);
						ix = clipi(ix, 0, cfg.nx - 1);
						iy = clipi(iy, 0, cfg.ny - 1);
						iz = clipi(iz, 0, cfg.nz - 1);
						q_v[lane] = XO_PACK_VOXEL(ix, iy, iz);
					}
				}
				__syncwarp();
				q_count = n_new;
			}
			if (state == ST_DEAD) {
				const u32 rank = (u32)__popc(dead_mask & lanemask_lt);
				if (rank < q_count) {
					const u32 slot = q_count - 1u - rank;
					const float4 a = q_a[slot], b = q_b[slot];
					trace_count = q_l[slot];
					pos.x = a.x; pos.y = a.y; pos.z = a.z; weight = a.w;
					dir.x = b.x; dir.y = b.y; dir.z = b.z; packet = __float_as_uint(b.w);
					vlo = q_v[slot];
					{ const u32 cell = XO_VOXEL(vlo); mat = cell & 0xffu; dcur = cell >> 8; }
					XO_LOAD_MAT(mat);
					opl = 0.0f;
					flags = EV_LAUNCH;
					state = ST_SETUP;
					started = true;
				} else if (q_dry) {
					state = ST_DRY;
				}
			}
			const u32 n_dead = (u32)__popc(dead_mask);
			q_count -= (n_dead < q_count) ? n_dead : q_count;
			__syncwarp();
			if (q_dry) {
				n_dry = (u32)__popc(__ballot_sync(0xffffffffu, state == ST_DRY));
				if (n_dry == 32u) break;
			}
		}

		// ---- a face between different materials, or the face of the grid -----------
		// (mcvox.template.c:275-388).  The crossing step already moved the voxel
		// address across the face, by `last_d`.
		if (state == ST_BND) {
			const u32 axis = (last_d == stx) ? 0u : ((last_d == sty) ? 1u : 2u);
			bool done = false;
			pos.x = fmaf(dir.x, t_evt, pos.x);
			pos.y = fmaf(dir.y, t_evt, pos.y);
			pos.z = fmaf(dir.z, t_evt, pos.z);
			if (XO_NEEDS_OPL) opl = fmaf(c_hot.n, t_evt, opl);
			const u32 entered = XO_VOXEL(vlo) & 0xffu;
			const bool escaping = (entered == XO_VOX_SENTINEL);
			dcur = 1;               // on a face between materials / of the grid
			const u32 next_mat = escaping ? 0u : entered;
			const float n1 = c_hot.n, n2 = sh_fast[next_mat].hot.n;
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
				} else if (next_mat != mat) {
					mat = next_mat;
					XO_LOAD_MAT(mat);
				}
			} else {
				// reflected: back into the voxel the packet came from
				vlo -= (u32)last_d;
			}
			XO_END_TRIP();
		}

		// ---- interaction: absorb, scatter, lottery (mcvox.template.c:925-981) -------
		// (convergence barrier: without it the compiler dispatches the phases through
		// jump tables on `state`, and the lanes that walked (SCAT) and the lanes that
		// skipped the walk (FAR) reach this block at different times and run it twice)
		__syncwarp();
		if (state == ST_SCAT || state == ST_FAR) {
			bool done = false;
			++iterations;
			pos.x = fmaf(dir.x, t_s, pos.x);
			pos.y = fmaf(dir.y, t_s, pos.y);
			pos.z = fmaf(dir.z, t_s, pos.z);
			if (XO_NEEDS_OPL) opl = fmaf(c_hot.n, t_s, opl);
#if XO_VOX_CLEARANCE
			u32 mat_here = mat;
			if (state == ST_FAR) {
				// the flight skipped the walk: voxel of the interaction point from the
				// position, loop trips of the reference = faces crossed on the way
				const u32 idx0 = (vlo - vbase_lo) >> 1;
				i32 ix = __float2int_rd((pos.x - cfg.top_left.x)*inv_sx);
				i32 iy = __float2int_rd((pos.y - cfg.top_left.y)*inv_sy);
				i32 iz = __float2int_rd((pos.z - cfg.top_left.z)*inv_sz);
				ix = clipi(ix, 0, cfg.nx - 1);
				iy = clipi(iy, 0, cfg.ny - 1);
				iz = clipi(iz, 0, cfg.nz - 1);
				iterations += (u32)(abs(ix - ((i32)(idx0 & vox_mx) - 2)) +
					abs(iy - ((i32)((idx0 >> vox_bx) & vox_my) - 2)) +
					abs(iz - ((i32)(idx0 >> vox_bxy) - 2)));
				vlo = XO_PACK_VOXEL(ix, iy, iz);
				const u32 cell = XO_VOXEL(vlo);
				dcur = cell >> 8;
				// (a rounding of the end point across a face of the clearance box can
				// land in another material: adopted after this interaction)
				if ((cell & 0xffu) != XO_VOX_SENTINEL) mat_here = cell & 0xffu;
			}
#endif
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
				flags |= EV_ABSORPTION;
#if XO_FLU_VOXGRID
				{   // the fluence grid is the voxel grid: the cell is the voxel of the walk
					const u32 idx = (vlo - vbase_lo) >> 1;
					fluence.deposit_cell(acc, (idx & vox_mx) - 2u, ((idx >> vox_bx) & vox_my) - 2u,
						(idx >> vox_bxy) - 2u, fluence_weight(deposit, c_hot.mua, fluence.k));
				}
#else
				if (XoFluence::active) fluence.deposit(acc, window, pos, deposit, c_hot.mua, opl);
#endif
			}
			pf_scatter(c_pf, rng, lut, dir);
			flags |= EV_SCATTERING;
			if (weight < XO_WEIGHT_MIN) {
#if XO_USE_LOTTERY
				if (rng.next_raw() > XO_LOTTERY_CHANCE*4294967296.0f) done = true;
				else weight *= (1.0f/XO_LOTTERY_CHANCE);
#else
				done = true;
#endif
			}
#endif
#if XO_VOX_CLEARANCE
			if (mat_here != mat) { mat = mat_here; XO_LOAD_MAT(mat); }
#endif
			XO_END_TRIP();
		}

		// ---- new ray from `pos` along `dir` in voxel (ix, iy, iz) -----------------------
		__syncwarp();
		if (state == ST_SETUP) {
			XO_DIR_CONSTS();
			t_s = fminf((FastMath::lg2(rng.next_raw()) - 32.0f)*c_hot.step_k, XO_FLT_MAX);
#if XO_USE_RMAX
			{   // parameter at which the ray leaves the rmax sphere around the source
				const float ex = pos.x - src_pos.x, ey = pos.y - src_pos.y, ez = pos.z - src_pos.z;
				const float b = ex*dir.x + ey*dir.y + ez*dir.z;
				const float c = ex*ex + ey*ey + ez*ez - rmax2;
				t_rmax = FastMath::sqrt(fmaxf(fmaf(b, b, -c), 0.0f)) - b;
			}
#endif
			state = ST_RUN;
#if XO_VOX_CLEARANCE
			{   // extent of the flight in voxels along its longest axis against the clearance
				const float ext = t_s*fmaxf(fabsf(dir.x)*inv_sx, fmaxf(fabsf(dir.y)*inv_sy, fabsf(dir.z)*inv_sz));
				bool far = ext < (float)dcur - 1.0f;
#if XO_USE_RMAX
				far = far && t_s <= t_rmax;
#endif
				if (far) state = ST_FAR;
			}
#endif
			if (state == ST_RUN) {
				const float rx = FastMath::rcp_approx(dir.x), ry = FastMath::rcp_approx(dir.y),
					rz = FastMath::rcp_approx(dir.z);
				const bool fx = dir.x >= 0.0f, fy = dir.y >= 0.0f, fz = dir.z >= 0.0f;
				stx = fx ? 2 : -2;
				sty = (fy ? 2 : -2) << vox_bx;
				stz = (fz ? 2 : -2) << vox_bxy;
				// exit faces of the current voxel (mcvox.template.c:173-196); the packed
				// index holds coordinate + 2
				const u32 idx = (vlo - vbase_lo) >> 1;
				const float facex = fmaf((float)((i32)(idx & vox_mx) - (fx ? 1 : 2)), cfg.size.x, cfg.top_left.x);
				const float facey = fmaf((float)((i32)((idx >> vox_bx) & vox_my) - (fy ? 1 : 2)), cfg.size.y, cfg.top_left.y);
				const float facez = fmaf((float)((i32)(idx >> vox_bxy) - (fz ? 1 : 2)), cfg.size.z, cfg.top_left.z);
				tmx = (dir.x != 0.0f) ? fmaxf((facex - pos.x)*rx, 0.0f) : XO_INF;
				tmy = (dir.y != 0.0f) ? fmaxf((facey - pos.y)*ry, 0.0f) : XO_INF;
				tmz = (dir.z != 0.0f) ? fmaxf((facez - pos.z)*rz, 0.0f) : XO_INF;
				tdx = cfg.size.x*fabsf(rx);
				tdy = cfg.size.y*fabsf(ry);
				tdz = cfg.size.z*fabsf(rz);
			}
		}

		// ---- voxel walk: lanes in RUN state cross faces until enough lanes wait ------
		// Two crossings per trip.  The lookup of the second one is issued before
		// the first has returned (its address needs only the face parameters):
		// a warp waits for L2 once per two crossings, since with 20+ lanes per
		// lookup some lane always misses L1.  The second crossing is committed
		// only if the first one stayed inside the material.
		const u32 wake = (refill + n_dry < 32u) ? refill + n_dry : 32u;
		for (;;) {
			if (state == ST_RUN) {
				const float tmin_a = fminf(tmx, fminf(tmy, tmz));
				if (!(tmin_a < t_s)) {
					state = ST_SCAT;
				} else {
					++iterations;
					const bool px = (tmx == tmin_a);
					const bool py = !px && (tmy == tmin_a);
					const bool pz = !px && !py;
					if (px) tmx += tdx;
					if (py) tmy += tdy;
					if (pz) tmz += tdz;
					last_d = px ? stx : (py ? sty : stz);
					vlo += (u32)last_d;
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
#if XO_TRACE == XO_TRACE_ALL
#define XO_TRACE_CROSSING(tmin_) do { \
						P3 pc_ = { fmaf(dir.x, tmin_, pos.x), fmaf(dir.y, tmin_, pos.y), fmaf(dir.z, tmin_, pos.z) }; \
						if (trace_event(tcfg, float_buffer, packet, trace_count, \
								flags | EV_BOUNDARY_HIT | EV_REFRACTION, pc_, dir, weight, \
								XO_NEEDS_OPL ? fmaf(c_hot.n, tmin_, opl) : 0.0f)) ++trace_count; \
						flags = 0; \
					} while (0)
#else
#define XO_TRACE_CROSSING(tmin_) do { } while (0)
#endif
#if XO_VOX_SAME_N
					// Equal refractive indices everywhere: a face between two materials is
					// no interface event, only the attenuation changes.  The remaining
					// optical depth of the flight is still Exp(1) distributed (memoryless),
					// so the free path is rescaled to the new material instead of stopping
					// the ray for a fresh draw (what the reference does, statistically the
					// same; the grid face and rmax stay events).
#define XO_MATERIAL_CHANGE(m_, tmin_) do { \
						const float k_old_ = c_hot.step_k; \
						XO_LOAD_MAT(m_); \
						mat = (m_); \
						t_s = fmaf(t_s - (tmin_), c_hot.step_k*FastMath::rcp_approx(k_old_), (tmin_)); \
					} while (0)
#define XO_IS_PLAIN_CHANGE(m_) ((m_) < XO_VOX_SENTINEL && c_hot.step_k > -XO_FLT_MAX && t_s < XO_FLT_MAX)
#else
#define XO_MATERIAL_CHANGE(m_, tmin_) do { } while (0)
#define XO_IS_PLAIN_CHANGE(m_) false
#endif
					if (m_a != mat) {
						if (XO_IS_PLAIN_CHANGE(m_a)) {
							XO_TRACE_CROSSING(tmin_a);
							XO_MATERIAL_CHANGE(m_a, tmin_a);
						} else {
							state = ST_BND;
							t_evt = tmin_a;
						}
					} else {
						XO_TRACE_CROSSING(tmin_a);      // the reference records every loop trip
						if (!ok_b) {
							state = ST_SCAT;
						} else {
							++iterations;
							if (qx) tmx += tdx;
							if (qy) tmy += tdy;
							if (qz) tmz += tdz;
							vlo = vlo_b;
							last_d = d_b;
							dcur = cell_b >> 8;
							if (m_b != mat) {
								if (XO_IS_PLAIN_CHANGE(m_b)) {
									XO_TRACE_CROSSING(tmin_b);
									XO_MATERIAL_CHANGE(m_b, tmin_b);
								} else {
									state = ST_BND;
									t_evt = tmin_b;
								}
							} else {
								XO_TRACE_CROSSING(tmin_b);
							}
						}
					}
#undef XO_TRACE_CROSSING
#undef XO_MATERIAL_CHANGE
#undef XO_IS_PLAIN_CHANGE
				}
			}
			if ((u32)__popc(__ballot_sync(0xffffffffu, state != ST_RUN)) >= wake) break;
		}
	}
#undef XO_LOAD_MAT
#undef XO_DIR_CONSTS
#undef XO_VOXEL
#undef XO_PACK_VOXEL
#undef XO_VOX_CLEARANCE
#undef XO_RMAX_TEST
#undef XO_TRACE_TRIP
#undef XO_END_TRIP
	// every lane drew from its stream (queue refills), whether or not it ever
	// carried a packet: all states go back
	rng_state_x[gid] = rng.state();
}
