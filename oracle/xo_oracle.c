/* oracle/xo_oracle.c -- TEST INFRASTRUCTURE: CPU restatement of the reference's
 * photon-packet kernels.  Never linked into, imported by or shipped with the
 * product (pyxopto_b200); see oracle/xo_oracle.h.
 *
 * Each function cites the reference text it restates (paths relative to
 * /root/reference/xopto).  Floating-point expressions keep the reference's
 * operand order, so with -ffp-contract=off and math=XO_MATH_LIBM this file is
 * bit-identical to the reference kernel compiled by gcc behind clshim.h
 * (oracle/_ref): its outputs on every case of tests/cases.py are committed as
 * tests/golden/<case>.npz (tests/golden/make_golden.py, run in the container that holds
 * the reference) and tests/test_oracle_golden.py holds this file to them bit for bit.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "xo_oracle.h"
#include "xo_detmath.h"

#define FP_0 0.0f
#define FP_1 1.0f
#define FP_2 2.0f
#define FP_0p5 0.5f
#define FP_2PI 6.283185307179586f
#define FP_PI 3.141592653589793f
#define FP_COS_90 0.0f
#define FP_COS_0 (1.0f - FP_COS_90)
#define FP_COS_30 0.8660254037844386f
#define FP_INV_C 3.3356409519815204e-09f
#define FP_RMIN 1e-12f
#define FP_PLMIN 1e-12f
#define ACCU_K 0x7FFFFF

/* event flags, mcbase.template.h:664-681 */
#define EV_REFLECTION 1u
#define EV_REFRACTION 2u
#define EV_BOUNDARY_HIT 4u
#define EV_LAUNCH 8u
#define EV_ABSORPTION 16u
#define EV_SCATTERING 32u
#define EV_TERMINATED 64u
#define EV_ESCAPED 128u

typedef struct { float x, y, z; } p3f;
typedef struct { float x, y; } p2f;
typedef struct { float a11, a12, a13, a21, a22, a23, a31, a32, a33; } m3f;

/* ---- packed plugin structs (ctypes layouts of the reference) -------------- */
/* mcml/mclayer/layer.py:57-69 */
typedef struct {
	float thickness, top, bottom, n, cc_top, cc_bottom, mus, mua, inv_mut, mua_inv_mut;
	/* McPf follows */
} ml_layer;
/* mcml/mclayer/layer.py:412-424 (AnisotropicLayer): same head, tensors instead of
 * the scalar coefficients */
typedef struct {
	float thickness, top, bottom, n, cc_top, cc_bottom;
	m3f mus, mua, mut;
} ml_aniso_layer;
/* mccyl/mclayer/layer.py:119-130 */
typedef struct {
	float r_inner, r_outer, n, cc_inner, cc_outer, mus, mua, inv_mut, mua_inv_mut;
} cyl_layer;
/* mccyl/mclayer/layer.py:477-491 (AnisotropicLayer) */
typedef struct {
	float r_inner, r_outer, n, cc_inner, cc_outer;
	m3f mus, mua, mut;
} cyl_aniso_layer;
/* mcbase/mcmaterial.py:52-62 */
typedef struct { float n, mus, mua, inv_mut, mua_inv_mut; } vox_material;
/* mcbase/mcmaterial.py:330-341 (AnisotropicMaterial) */
typedef struct { float n; m3f mus, mua, mut; } vox_aniso_material;
/* mcvox/mcgeometry/voxel.py:96-121 */
typedef struct { p3f top_left, bottom_right, size; int32_t nx, ny, nz; } vox_cfg;

typedef struct { float g; } pf_hg;                          /* mcpf/hg.py:49 */
typedef struct { float g, beta; } pf_mhg;                   /* mcpf/mhg.py:52 */
typedef struct { float g, a, inv_a, a1, a2; } pf_gk;        /* mcpf/gk.py:58 */
typedef struct { float a, b, c; uint32_t offset, size; } pf_lut; /* mcpf/lut.py:78 */
typedef struct { float g1, g2, b; } pf_hg2;                  /* mcpf/hg2.py:58 */
typedef struct { pf_gk gk_1, gk_2; float b; } pf_gk2;        /* mcpf/gk2.py:61 */
typedef struct { float g, a, beta, inv_a, a1, a2; } pf_mgk;  /* mcpf/mgk.py:64 */
typedef struct { float n; } pf_pc;                           /* mcpf/pc.py:50 */
typedef struct { float gamma, a, b; } pf_rayleigh;           /* mcpf/rayleigh.py:54 */
typedef struct { float n, beta; } pf_mpc;
typedef struct { p3f direction; float g, p; } pf_hgdir;      /* mcpf/hgdir.py:62-66 */                    /* mcpf/mpc.py:54 */

typedef struct { p3f position, dir_medium, dir_sample, dir_reflected; float reflectance; } src_line;
typedef struct { m3f T; p3f position, direction; p2f sigma; float clip, reflectance; } src_gauss_ml;
typedef struct { m3f T; p3f position, direction; float radius, cos_min, n; } src_ufiber;
typedef struct { m3f T; p3f position, direction; p2f radius; float reflectance; } src_ubeam_ml; /* mcsource/uniformbeam.py:36-47 */
typedef struct { p3f position; uint32_t layer_index; } src_isopoint;
/* mcsource/rectangular.py:56-63 (cos_min) / :343-350 (na) */
typedef struct { p3f position; p2f size; float n, cos_critical, aperture; uint32_t layer_index; } src_rect;
typedef struct { m3f T; p3f position, direction; p2f sigma; float clip; } src_gauss_vox;  /* mcvox/mcsource/gaussianbeam.py:71-77 */
typedef struct { p3f position; uint32_t n, offset; } src_isovoxels;                    /* mcvox/mcsource/voxel.py:205-209 */
typedef struct { p3f position; int32_t vx, vy, vz; } src_isovoxel;                    /* mcvox/mcsource/voxel.py:44-47 */
typedef struct { p3f position; } src_isopoint_vox;                                  /* mcvox/mcsource/point.py:44-46 */

typedef struct { p3f direction; float cos_min; uint32_t offset; } det_total;
typedef struct { p3f direction; p2f position; float r_min, inv_dr, cos_min;
	uint32_t n, offset; int32_t log_scale; } det_radial;
typedef struct { p3f direction; float x_min, inv_dx, y_min, inv_dy, cos_min;
	uint32_t n_x, n_y, offset; } det_cartesian;
typedef struct { m3f T; p2f position; float core_r_squared, core_spacing, cos_min;
	uint32_t offset; } det_six;
/* mcdetector/probe/lineararray.py:58-66 */
typedef struct { m3f T; p2f first_position, delta_position; float core_r_squared, cos_min;
	uint32_t offset; } det_linarray;
/* mcdetector/probe/lineararraypl.py:66-78 */
typedef struct { m3f T; p2f first_position, delta_position; float core_r_squared, cos_min,
	pl_min, inv_dpl; uint32_t n_pl, offset; int32_t pl_log_scale; } det_linarraypl;
typedef struct { p3f direction; p2f position; float r_min, inv_dr, pl_min, inv_dpl, cos_min;
	uint32_t n_r, n_pl, offset; int32_t r_log_scale, pl_log_scale; } det_radialpl;
typedef struct { p3f direction; float cos_min, pl_min, inv_dpl;
	uint32_t n_pl, offset; int32_t pl_log_scale; } det_totalpl;
/* mcdetector/symmetric.py:45-56 */
typedef struct { p3f direction; float position_x, x_offset, inv_step, cos_min;
	uint32_t n_half; int32_t log_scale; uint32_t offset; } det_symx;
/* mcdetector/cartesianpl.py, mcdetector/probe/sixaroundonepl.py */
typedef struct { p3f direction; float x_min, inv_dx, y_min, inv_dy, pl_min, inv_dpl, cos_min;
	uint32_t n_x, n_y, n_pl, offset; int32_t pl_log_scale; } det_cartesianpl;
typedef struct { m3f T; p2f position; float core_r_squared, core_spacing, pl_min, inv_dpl,
	cos_min; uint32_t n_pl, offset; int32_t pl_log_scale; } det_sixpl;
/* mccyl/mcdetector/fiz.py:44-56 (pack=1), mccyl/mcdetector/total.py:44-50 */
typedef struct { float fi_min, inv_dfi, z_min, inv_dz, cos_min;
	uint32_t n_fi, n_z, offset; } det_fiz;
typedef struct { float cos_min; uint32_t offset; } det_total_cyl;

typedef struct { p3f inv_step, top_left; uint32_t nx, ny, nz, offset; int32_t k; } flu_xyz;
typedef struct { p3f center; float inv_dr, inv_dz; uint32_t n_r, n_z, offset; int32_t k; } flu_rz;
typedef struct { float inv_step[4], top_left[4]; uint32_t shape[4]; uint32_t offset; int32_t k; } flu_xyzt;

typedef struct { p3f center; float t_min, inv_dr, inv_dz, inv_dt;
	uint32_t n_r, n_z, n_t, offset; int32_t k; } flu_rzt;       /* mcfluence/fluencerzt.py:54 */
typedef struct { p2f center; float r_min, fi_min, z_min, inv_dr, inv_dfi, inv_dz;
	uint32_t n_r, n_fi, n_z, offset; int32_t k; } flu_cyl;       /* mcfluence/fluencecyl.py:56 */

typedef struct { p2f center; float r_min, fi_min, z_min, t_min, inv_dr, inv_dfi, inv_dz, inv_dt;
	uint32_t n_r, n_fi, n_z, n_t, offset; int32_t k; } flu_cylt;  /* mcfluence/fluencecylt.py:79 */

typedef struct { int32_t max_events; uint32_t data_off, count_off, event_mask; } trace_cfg;

/* ---- simulator state (mcml.template.h McSimState) ------------------------- */
typedef struct {
	const xo_oracle_job *job;
	p3f pos, dir;
	uint64_t rng_x;
	uint32_t rng_a;
	float weight;
	uint32_t photon_index;
	int32_t layer_index;      /* layer (mcml/mccyl) or material index (mcvox) */
	uint32_t event_flags;
	float opl;
	uint32_t trace_count;
	int32_t vx, vy, vz;       /* mcvox voxel index */
	uint64_t iterations;
	volatile uint32_t *dyn_counter;   /* dynamic schedule packet counter */
} sim_t;

/* ---- math binding ------------------------------------------------------- */
#define PORTABLE(s) ((s)->job->math == XO_MATH_PORTABLE)
static inline float m_log(const sim_t *s, float x) { return PORTABLE(s) ? xo_logf(x) : logf(x); }
static inline float m_exp(const sim_t *s, float x) { return PORTABLE(s) ? xo_expf(x) : expf(x); }
static inline float m_cbrt(const sim_t *s, float x) { return PORTABLE(s) ? xo_cbrtf(x) : cbrtf(x); }
static inline float m_pow(const sim_t *s, float x, float y) { return PORTABLE(s) ? xo_powf(x, y) : powf(x, y); }
static inline float m_atan2(const sim_t *s, float y, float x) { return PORTABLE(s) ? xo_atan2f(y, x) : atan2f(y, x); }
static inline void m_sincos(const sim_t *s, float x, float *sn, float *cs) {
	if (PORTABLE(s)) { *sn = xo_sincosf(x, cs); }
	else { *cs = cosf(x); *sn = sinf(x); }
}
#define m_sqrt(x) sqrtf(x)
#define m_div(a, b) ((float)(a)/(float)(b))
static inline float fclip(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline int32_t iclip(int32_t x, int32_t lo, int32_t hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline int fsign(float x) { return (x >= FP_0) ? 1 : -1; }

/* ---- RNG: mcbase.template.c:1576-1594 -------------------------------------- */
static inline float rng_next(uint64_t *x, uint32_t a) {
	*x = (*x & 0xFFFFFFFFull)*a + (*x >> 32);
	return (float)(uint32_t)*x / (float)0xFFFFFFFFu;
}
/* MC_USE_ENHANCED_RNG (mcbase.template.c:1577-1586): two steps, 64 random bits,
 * (float)u64 / (float)0xFFFFFFFFFFFFFFFF (the divisor rounds to 2^64) */
static inline float rng_next_enhanced(uint64_t *x, uint32_t a) {
	*x = (*x & 0xFFFFFFFFull)*a + (*x >> 32);
	uint32_t high = (uint32_t)*x;
	*x = (*x & 0xFFFFFFFFull)*a + (*x >> 32);
	uint32_t low = (uint32_t)*x;
	return (float)((((uint64_t)high) << 32) + low) / (float)0xFFFFFFFFFFFFFFFFull;
}
static inline float sim_random(sim_t *s) {
	return s->job->enhanced_rng ? rng_next_enhanced(&s->rng_x, s->rng_a)
		: rng_next(&s->rng_x, s->rng_a);
}

void xo_oracle_rng_test(uint64_t x, uint32_t a, uint32_t n, float *out) {
	for (uint32_t i = 0; i < n; ++i) out[i] = rng_next(&x, a);
}

/* ---- seeds: src/rng/rng.cpp:64-103 ---------------------------------------- */
int xo_oracle_init_rng(uint64_t *x, uint32_t *a, const uint32_t *fora,
		uint32_t n_rng, uint64_t xinit) {
	uint32_t begin = fora[0];
	if (xinit == 0ull || (uint32_t)(xinit >> 32) >= begin - 1 ||
			(uint32_t)xinit >= 0xffffffffu)
		return 1;
	for (uint32_t i = 0; i < n_rng; ++i) {
		a[i] = fora[i + 1];
		x[i] = 0;
		while (x[i] == 0 || (uint32_t)(x[i] >> 32) >= fora[i + 1] - 1 ||
				(uint32_t)x[i] >= 0xffffffffu) {
			xinit = (xinit & 0xffffffffull)*begin + (xinit >> 32);
			x[i] = (uint32_t)floor(((double)(uint32_t)xinit/4294967296.0)*fora[i + 1]);
			x[i] <<= 32;
			xinit = (xinit & 0xffffffffull)*begin + (xinit >> 32);
			x[i] += (uint32_t)xinit;
		}
	}
	return 0;
}

/* ---- accumulators: mcbase.template.c:58-61 --------------------------------- */
static inline void accu_deposit(sim_t *s, uint32_t index, uint32_t w) {
	__atomic_fetch_add(&s->job->accumulator_buffer[index], (uint64_t)w, __ATOMIC_RELAXED);
}
/* mcml.template.h:606 weight_to_int + the `*(cond)` idiom of every detector */
static inline uint32_t weight_to_u32(float weight, int accept) {
	return (uint32_t)((weight*ACCU_K + FP_0p5)*accept);
}

/* ---- vector helpers: mcbase.template.c:693-700,868-990 --------------------- */
static inline float dot3(const p3f *a, const p3f *b) { return a->x*b->x + a->y*b->y + a->z*b->z; }
static inline void transform3(const m3f *m, const p3f *v, p3f *r) {
	r->x = m->a11*v->x + m->a12*v->y + m->a13*v->z;
	r->y = m->a21*v->x + m->a22*v->y + m->a23*v->z;
	r->z = m->a31*v->x + m->a32*v->y + m->a33*v->z;
}
static inline void normalize3(p3f *a) {
	float k = m_div(FP_1, m_sqrt(a->x*a->x + a->y*a->y + a->z*a->z));
	a->x = a->x*k; a->y = a->y*k; a->z = a->z*k;
}

/* ---- boundary physics: mcbase.template.c:1229-1507 ------------------------- */
static inline float cos_critical(float n1, float n2) {
	return (n1 > n2) ? m_sqrt(FP_1 - m_div(n2*n2, n1*n1)) : FP_0;
}
static float reflectance(float n1, float n2, float cos1, float cos_crit) {
	float Rp, Rs, R = FP_1, n1_d_n2, sin1, sin2, cos2, n_cos1, n_cos2;
	cos1 = fabsf(cos1);
	if (n1 == n2) return FP_0;
	if (cos1 > cos_crit) {
		n1_d_n2 = m_div(n1, n2);
		sin1 = m_sqrt(FP_1 - cos1*cos1);
		if (cos1 >= FP_COS_0) sin1 = FP_0;
		sin2 = fminf(FP_1, n1_d_n2*sin1);
		cos2 = m_sqrt(FP_1 - sin2*sin2);
		n_cos1 = n1_d_n2*cos1;
		n_cos2 = n1_d_n2*cos2;
		Rs = m_div(n_cos1 - cos2, n_cos1 + cos2); Rs *= Rs;
		Rp = m_div(n_cos2 - cos1, n_cos2 + cos1); Rp *= Rp;
		R = FP_0p5*(Rp + Rs);
		if (cos1 <= FP_COS_90 || sin2 == FP_1) return FP_1;
	}
	return R;
}
static inline void reflect3(const p3f *p, const p3f *n, p3f *r) {
	float p_n_2 = FP_2*dot3(p, n);
	r->x = p->x - n->x*p_n_2; r->y = p->y - n->y*p_n_2; r->z = p->z - n->z*p_n_2;
}
/* mcbase.template.c:1303-1348 */
static inline float reflectance_cos2(float n1, float n2, float cos2) {
	float Rp, Rs, R = FP_1, n1_d_n2, sin1, sin2, cos1, n_cos1, n_cos2;
	cos2 = fabsf(cos2);
	if (n1 == n2) return FP_0;
	sin2 = m_sqrt(FP_1 - cos2*cos2);
	if (cos2 >= FP_COS_0) sin2 = FP_0;
	sin1 = m_div(n2, n1)*sin2;
	if (sin1 < FP_1) {
		cos1 = m_sqrt(FP_1 - sin1*sin1);
		n1_d_n2 = m_div(n1, n2);
		n_cos1 = n1_d_n2*cos1;
		n_cos2 = n1_d_n2*cos2;
		Rs = m_div(n_cos1 - cos2, n_cos1 + cos2); Rs *= Rs;
		Rp = m_div(n_cos2 - cos1, n_cos2 + cos1); Rp *= Rp;
		R = FP_0p5*(Rp + Rs);
		if (cos1 <= FP_COS_90 || sin2 == FP_1) return FP_1;
	}
	return R;
}

static inline void refract3(const p3f *p, const p3f *n, float n1, float n2, p3f *r) {
	float cos1 = dot3(p, n);
	float n1_d_n2 = m_div(n1, n2);
	float sin2_squared = n1_d_n2*n1_d_n2*(FP_1 - cos1*cos1);
	float k = fsign(cos1)*(n1_d_n2*fabsf(cos1) - m_sqrt(FP_1 - sin2_squared));
	r->x = n1_d_n2*p->x - k*n->x; r->y = n1_d_n2*p->y - k*n->y; r->z = n1_d_n2*p->z - k*n->z;
}
static inline int refract3_safe(const p3f *p, const p3f *n, float n1, float n2, p3f *r) {
	float cos1 = dot3(p, n);
	float n1_d_n2 = m_div(n1, n2);
	float sin2_squared = n1_d_n2*n1_d_n2*(FP_1 - cos1*cos1);
	if (sin2_squared > FP_1) return 1;
	float k = fsign(cos1)*(n1_d_n2*fabsf(cos1) - m_sqrt(FP_1 - sin2_squared));
	r->x = n1_d_n2*p->x - k*n->x; r->y = n1_d_n2*p->y - k*n->y; r->z = n1_d_n2*p->z - k*n->z;
	return 0;
}

/* mcbase.template.c:1522-1555 (renormalisation is unconditional) */
static void scatter_direction(const sim_t *s, p3f *dir, float cos_theta, float fi) {
	float sin_fi, cos_fi, sin_theta, px, k, stcf, stsf;
	sin_theta = m_sqrt(FP_1 - cos_theta*cos_theta);
	m_sincos(s, fi, &sin_fi, &cos_fi);
	stcf = sin_theta*cos_fi;
	stsf = sin_theta*sin_fi;
	px = dir->x;
	if (fabsf(dir->z) >= FP_COS_0) {
		dir->x = stcf;
		dir->y = stsf;
		dir->z = copysignf(cos_theta, dir->z*cos_theta);
	} else {
		k = m_sqrt(FP_1 - dir->z*dir->z);
		dir->x = m_div(stcf*px*dir->z - stsf*dir->y, k) + px*cos_theta;
		dir->y = m_div(stcf*dir->y*dir->z + stsf*px, k) + dir->y*cos_theta;
		dir->z = (-stcf)*k + dir->z*cos_theta;
	}
	normalize3(dir);
}

/* ---- layer / material access --------------------------------------------- */
static inline size_t layer_stride(const xo_oracle_job *j) {
	switch (j->geometry) {
		case XO_GEOM_MCML: return (j->anisotropic ? sizeof(ml_aniso_layer) : sizeof(ml_layer)) +
			(size_t)j->pf_size;
		case XO_GEOM_MCCYL: return (j->anisotropic ? sizeof(cyl_aniso_layer) : sizeof(cyl_layer)) +
			(size_t)j->pf_size;
		default: return (j->anisotropic ? sizeof(vox_aniso_material) : sizeof(vox_material)) +
			(size_t)j->pf_size;
	}
}
static inline const ml_layer *ml_layer_at(const xo_oracle_job *j, int32_t i) {
	return (const ml_layer *)((const char *)j->layers + (size_t)i*layer_stride(j));
}
static inline const cyl_layer *cyl_layer_at(const xo_oracle_job *j, int32_t i) {
	return (const cyl_layer *)((const char *)j->layers + (size_t)i*layer_stride(j));
}
static inline const vox_material *vox_material_at(const xo_oracle_job *j, int32_t i) {
	return (const vox_material *)((const char *)j->layers + (size_t)i*layer_stride(j));
}
static inline const void *current_pf(const sim_t *s) {
	const xo_oracle_job *j = s->job;
	const char *base = (const char *)j->layers + (size_t)s->layer_index*layer_stride(j);
	return base + (layer_stride(j) - (size_t)j->pf_size);
}
/* refractive index of layer/material i in any geometry */
static inline float medium_n(const xo_oracle_job *j, int32_t i) {
	switch (j->geometry) {
		case XO_GEOM_MCML: return ml_layer_at(j, i)->n;
		case XO_GEOM_MCCYL: return cyl_layer_at(j, i)->n;
		default: return vox_material_at(j, i)->n;
	}
}
/* mcbase.template.h:2227-2230 tensor3f_project: p T p' */
static inline float tensor_project(const m3f *T, const p3f *p) {
	return p->x*(T->a11*p->x + T->a12*p->y + T->a13*p->z) +
		p->y*(T->a21*p->x + T->a22*p->y + T->a23*p->z) +
		p->z*(T->a31*p->x + T->a32*p->y + T->a33*p->z);
}
/* Optical properties of the current medium along the propagation direction.  Isotropic
 * media: the packed scalars (mclayer/layer.py:153-216, mcmaterial.py:94-134); anisotropic
 * media: the projected tensors with the reference's guards (mc_layer_inv_mut /
 * mc_layer_mua_inv_mut, mclayer/layer.py:497-551; mc_material_*, mcmaterial.py:390-455) */
typedef struct { float mus, mua, inv_mut, mua_inv_mut; } medium_scalars;
static inline void aniso_tensors(const sim_t *s, const void *medium,
		const m3f **mus, const m3f **mua, const m3f **mut) {
	switch (s->job->geometry) {
		case XO_GEOM_MCML: { const ml_aniso_layer *L = (const ml_aniso_layer *)medium;
			*mus = &L->mus; *mua = &L->mua; *mut = &L->mut; break; }
		case XO_GEOM_MCCYL: { const cyl_aniso_layer *L = (const cyl_aniso_layer *)medium;
			*mus = &L->mus; *mua = &L->mua; *mut = &L->mut; break; }
		default: { const vox_aniso_material *M = (const vox_aniso_material *)medium;
			*mus = &M->mus; *mua = &M->mua; *mut = &M->mut; break; }
	}
}
static inline float aniso_mus(const sim_t *s, const void *medium) {
	const m3f *mus, *mua, *mut; aniso_tensors(s, medium, &mus, &mua, &mut);
	return tensor_project(mus, &s->dir);
}
static inline float aniso_mua(const sim_t *s, const void *medium) {
	const m3f *mus, *mua, *mut; aniso_tensors(s, medium, &mus, &mua, &mut);
	return tensor_project(mua, &s->dir);
}
static inline float aniso_inv_mut(const sim_t *s, const void *medium) {
	const m3f *mus, *mua, *mut; aniso_tensors(s, medium, &mus, &mua, &mut);
	float m = tensor_project(mut, &s->dir);
	return (m != FP_0) ? m_div(FP_1, m) : INFINITY;
}
static inline float aniso_mua_inv_mut(const sim_t *s, const void *medium) {
	const m3f *mus, *mua, *mut; aniso_tensors(s, medium, &mus, &mua, &mut);
	float a = tensor_project(mua, &s->dir);
	float m = tensor_project(mut, &s->dir);
	return (a != FP_0) ? ((m != FP_0) ? m_div(a, m) : INFINITY) : FP_0;
}
#define MEDIUM_ACCESSORS(prefix, type) \
	static inline float prefix##_mus(const sim_t *s, const type *L) { \
		return s->job->anisotropic ? aniso_mus(s, L) : L->mus; } \
	static inline float prefix##_mua(const sim_t *s, const type *L) { \
		return s->job->anisotropic ? aniso_mua(s, L) : L->mua; } \
	static inline float prefix##_inv_mut(const sim_t *s, const type *L) { \
		return s->job->anisotropic ? aniso_inv_mut(s, L) : L->inv_mut; } \
	static inline float prefix##_mua_inv_mut(const sim_t *s, const type *L) { \
		return s->job->anisotropic ? aniso_mua_inv_mut(s, L) : L->mua_inv_mut; }
MEDIUM_ACCESSORS(ml, ml_layer)
MEDIUM_ACCESSORS(cyl, cyl_layer)
MEDIUM_ACCESSORS(vox, vox_material)

/* ---- phase functions ------------------------------------------------------- */
static float pf_sample_angles(sim_t *s, float *azimuth) {
	const void *pf = current_pf(s);
	float cos_theta;
	switch (s->job->pf_kind) {
	case XO_PF_HG: {                                   /* mcpf/hg.py:74-90 */
		float g = ((const pf_hg *)pf)->g, k;
		*azimuth = FP_2PI*sim_random(s);
		k = m_div(FP_1 - g*g, FP_1 + g*(FP_2*sim_random(s) - FP_1));
		cos_theta = m_div(FP_1 + g*g - k*k, FP_2*g);
		if (g == FP_0) cos_theta = FP_1 - FP_2*sim_random(s);
		return fmaxf(fminf(cos_theta, FP_1), -FP_1);
	}
	case XO_PF_MHG: {                                  /* mcpf/mhg.py:84-105 */
		float g = ((const pf_mhg *)pf)->g, beta = ((const pf_mhg *)pf)->beta, k;
		*azimuth = FP_2PI*sim_random(s);
		if (sim_random(s) <= beta) {
			k = m_div(FP_1 - g*g, FP_1 + g*(FP_2*sim_random(s) - FP_1));
			cos_theta = m_div(FP_1 + g*g - k*k, FP_2*g);
			if (g == FP_0) cos_theta = FP_1 - FP_2*sim_random(s);
		} else {
			cos_theta = m_cbrt(s, FP_2*sim_random(s) - FP_1);
		}
		return fclip(cos_theta, -FP_1, FP_1);
	}
	case XO_PF_GK: {                                   /* mcpf/gk.py:99-132 */
		const pf_gk *p = (const pf_gk *)pf;
		float g = p->g, a = p->a, inv_a = p->inv_a, a1 = p->a1, a2 = p->a2, tmp;
		*azimuth = FP_2PI*sim_random(s);
		if (g == FP_0) {
			cos_theta = FP_1 - FP_2*sim_random(s);
		} else if (a == FP_0) {
			cos_theta = a1 + m_pow(s, m_div(FP_1 - g, FP_1 + g), FP_2*sim_random(s))*a2;
		} else {
			tmp = a1*sim_random(s) + a2;
			tmp = FP_1 + g*g - m_pow(s, tmp, -inv_a);
			cos_theta = m_div(tmp, FP_2*g);
		}
		return fclip(cos_theta, -FP_1, FP_1);
	}
	case XO_PF_HG2: {                                  /* mcpf/hg2.py:100-128 */
		const pf_hg2 *p = (const pf_hg2 *)pf;
		float g, k;
		*azimuth = FP_2PI*sim_random(s);
		if (sim_random(s) >= p->b) g = p->g1; else g = p->g2;
		k = m_div(FP_1 - g*g, FP_1 + g*(FP_2*sim_random(s) - FP_1));
		cos_theta = m_div(FP_1 + g*g - k*k, FP_2*g);
		if (g == FP_0) cos_theta = FP_1 - FP_2*sim_random(s);
		return fmaxf(fminf(cos_theta, FP_1), -FP_1);
	}
	case XO_PF_GK2:                                    /* mcpf/gk2.py:119-160 */
	case XO_PF_MGK: {                                  /* mcpf/mgk.py:108-140 */
		float g, a, inv_a, a1, a2, tmp;
		*azimuth = FP_2PI*sim_random(s);
		if (s->job->pf_kind == XO_PF_GK2) {
			const pf_gk2 *p = (const pf_gk2 *)pf;
			const pf_gk *q = (sim_random(s) >= p->b) ? &p->gk_1 : &p->gk_2;
			g = q->g; a = q->a; inv_a = q->inv_a; a1 = q->a1; a2 = q->a2;
		} else {
			const pf_mgk *p = (const pf_mgk *)pf;
			g = p->g; a = p->a; inv_a = p->inv_a; a1 = p->a1; a2 = p->a2;
			if (!(sim_random(s) <= p->beta)) {
				cos_theta = m_cbrt(s, FP_2*sim_random(s) - FP_1);
				return fclip(cos_theta, -FP_1, FP_1);
			}
		}
		if (g == FP_0) {
			cos_theta = FP_1 - FP_2*sim_random(s);
		} else if (a == FP_0) {
			cos_theta = a1 + m_pow(s, m_div(FP_1 - g, FP_1 + g), FP_2*sim_random(s))*a2;
		} else {
			tmp = a1*sim_random(s) + a2;
			tmp = FP_1 + g*g - m_pow(s, tmp, -inv_a);
			cos_theta = m_div(tmp, FP_2*g);
		}
		return fclip(cos_theta, -FP_1, FP_1);
	}
	case XO_PF_RAYLEIGH: {
		/* mcpf/rayleigh.py:86-104 with its two syntax slips (:96 missing ';', :99 ');')
		   repaired - the text does not compile as shipped */
		const pf_rayleigh *p = (const pf_rayleigh *)pf;
		float b = p->b, a = p->a, tmp;
		*azimuth = FP_2PI*sim_random(s);
		if (p->gamma == FP_1) {
			cos_theta = FP_2*sim_random(s) - FP_1;
		} else {
			b = b*(FP_1 - FP_2*sim_random(s));
			tmp = m_sqrt(b*b*0.25f + a*a*a*0.037037037037037035f);
			cos_theta = m_cbrt(s, -0.5f*b + tmp) + m_cbrt(s, -0.5f*b - tmp);
		}
		return fclip(cos_theta, -FP_1, FP_1);
	}
	case XO_PF_PC: {                                   /* mcpf/pc.py:82-95 */
		float n = ((const pf_pc *)pf)->n, r;
		*azimuth = FP_2PI*sim_random(s);
		r = sim_random(s);
		cos_theta = FP_2*m_pow(s, r, m_div(FP_1, n + FP_1)) - FP_1;
		return fclip(cos_theta, -FP_1, FP_1);
	}
	case XO_PF_MPC: {                                  /* mcpf/mpc.py:91-110 */
		const pf_mpc *p = (const pf_mpc *)pf;
		*azimuth = FP_2PI*sim_random(s);
		if (sim_random(s) <= p->beta) {
			float r = sim_random(s);
			cos_theta = FP_2*m_pow(s, r, m_div(FP_1, p->n + FP_1)) - FP_1;
		} else {
			cos_theta = m_cbrt(s, FP_2*sim_random(s) - FP_1);
		}
		return fclip(cos_theta, -FP_1, FP_1);
	}
	case XO_PF_LUT: {                                  /* mcpf/lut.py:125-158 */
		const pf_lut *p = (const pf_lut *)pf;
		float a = p->a, b = p->b, c = p->c, fp_index, fp_index_floor, d;
		size_t offset = p->offset, lut_size = p->size - 1, index;
		const float *lut = s->job->fp_lut;
		*azimuth = FP_2PI*sim_random(s);
		fp_index = (m_div(a, sim_random(s) - c) - b + FP_1)*lut_size*FP_0p5;
		fp_index_floor = floorf(fp_index);
		d = fp_index - fp_index_floor;
		index = (size_t)(int32_t)fp_index_floor;
		{
			/* mc_min(index + 1, lut_size) casts both to mc_int_t */
			int32_t i1 = (int32_t)(index + 1), ls = (int32_t)lut_size;
			int32_t i2 = i1 < ls ? i1 : ls;
			cos_theta = lut[offset + index]*(FP_1 - d) + lut[offset + (size_t)i2]*d;
		}
		return cos_theta;
	}
	}
	return FP_1;
}

/* ---- linear lookup tables: mcbase.template.h:2494-2548 ---------------------- */
typedef struct { float first, inv_span; uint32_t n, offset; } fp_lut_t;
typedef struct { fp_lut_t lut; p3f direction; uint32_t offset; } det_totallut;   /* total.py:262-266 */
typedef struct { fp_lut_t lut; p3f direction; float pl_min, inv_dpl; uint32_t n_pl, offset;
	int32_t pl_log_scale; } det_totallutpl;                                      /* totalpl.py:340-348 */
/* fp_linear_lut_sample: rounded first index, fractional weight, untouched when
 * the position is outside the table */
static inline void fp_lut_sample(const float *buffer, const fp_lut_t *lut, float where, float *value) {
	float fp_index = (where - lut->first)*lut->inv_span*(lut->n - 1);
	if (fp_index >= FP_0 && fp_index <= lut->n - 1) {
		uint32_t index1 = (uint32_t)(fp_index + FP_0p5);
		float w2 = fp_index - floorf(fp_index);
		uint32_t index2 = (uint32_t)iclip((int32_t)(index1 + 1), 0, (int32_t)(lut->n - 1));
		*value = buffer[lut->offset + index1]*(FP_1 - w2) + buffer[lut->offset + index2]*w2;
	}
}

typedef struct { m3f T; p3f position, direction; float radius, n; fp_lut_t lut; } src_ufiberlut;  /* mcsource/fiber.py:719-727 */
/* mcsource/fiberni.py:210-215 (aperture = cos_min), :453-458 (aperture = na), :611-616 */
typedef struct { p3f position; float radius, aperture, n; } src_fiberni;
typedef struct { p3f position; float radius, n; fp_lut_t lut; } src_fiberlutni;
/* mcsource/rectangular.py:553-561 */
typedef struct { p3f position; p2f size; float n, cos_critical; fp_lut_t lut; uint32_t layer_index; } src_rectlut;
/* fp_linear_lut_rel_sample (mcbase.template.h:2517-2527) */
static inline void fp_lut_rel_sample(const float *buffer, const fp_lut_t *lut, float where, float *value) {
	float fp_index = where*(lut->n - 1);
	uint32_t index1 = (uint32_t)(fp_index + FP_0p5);
	if (index1 < lut->n) {
		float w2 = fp_index - floorf(fp_index);
		uint32_t index2 = (uint32_t)iclip((int32_t)(index1 + 1), 0, (int32_t)(lut->n - 1));
		*value = buffer[lut->offset + index1]*(FP_1 - w2) + buffer[lut->offset + index2]*w2;
	}
}

/* ---- detectors ------------------------------------------------------------- */
enum { LOC_TOP = 0, LOC_BOTTOM = 1, LOC_SPECULAR = 2 };

static void detector_deposit(sim_t *s, int loc, const p3f *pos, const p3f *dir, float weight) {
	const xo_oracle_job *j = s->job;
	const char *base = (const char *)j->detectors + j->det_offset[loc];
	const int32_t param = j->det_param[loc];
	(void)param;
	switch (j->det_kind[loc]) {
	case XO_DET_TOTAL: {                               /* mcdetector/total.py:86-110 */
		const det_total *d = (const det_total *)base;
		p3f dd = d->direction;
		uint32_t w = weight_to_u32(weight, d->cos_min <= fabsf(dot3(dir, &dd)));
		if (w > 0) accu_deposit(s, d->offset, w);
		break;
	}
	case XO_DET_TOTALLUT: {                            /* mcdetector/total.py:290-326 */
		const det_totallut *d = (const det_totallut *)base;
		float sensitivity = FP_0;
		p3f dd = d->direction;
		fp_lut_sample(j->fp_lut, &d->lut, fabsf(dot3(dir, &dd)), &sensitivity);
		uint32_t w = (uint32_t)(weight*sensitivity*ACCU_K + FP_0p5);
		if (w > 0) accu_deposit(s, d->offset, w);
		break;
	}
	case XO_DET_TOTALLUTPL: {                          /* mcdetector/totalpl.py:380-415 */
		const det_totallutpl *d = (const det_totallutpl *)base;
		float pl = s->opl;
		if (d->pl_log_scale) pl = m_log(s, fmaxf(pl, FP_PLMIN));
		int32_t pl_index = (int32_t)((pl - d->pl_min)*d->inv_dpl);
		pl_index = iclip(pl_index, 0, (int32_t)d->n_pl - 1);
		float sensitivity = FP_0;
		p3f dd = d->direction;
		fp_lut_sample(j->fp_lut, &d->lut, fabsf(dot3(dir, &dd)), &sensitivity);
		uint32_t w = (uint32_t)(weight*sensitivity*ACCU_K + FP_0p5);
		if (w > 0) accu_deposit(s, d->offset + (uint32_t)pl_index, w);
		break;
	}
	case XO_DET_FIBERLUTARRAY: {                       /* mcdetector/probe/fiberlutarray.py:100-170 */
		/* packed: m3f T[n]; p2f core_position[n]; float core_r_squared[n]; fp_lut_t lut[n]; u32 offset */
		uint32_t n = (uint32_t)param, fiber_index = n;
		const m3f *Ts = (const m3f *)base;
		const p2f *cp = (const p2f *)(Ts + n);
		const float *r2s = (const float *)(cp + n);
		const fp_lut_t *luts = (const fp_lut_t *)(r2s + n);
		uint32_t offset = *(const uint32_t *)(luts + n);
		p3f mc_pos, dp; float dx, dy, r2;
		for (uint32_t index = 0; index < n; ++index) {
			mc_pos.x = pos->x - cp[index].x; mc_pos.y = pos->y - cp[index].y; mc_pos.z = FP_0;
			m3f T = Ts[index];
			transform3(&T, &mc_pos, &dp);
			dx = dp.x; dy = dp.y; r2 = dx*dx + dy*dy;
			if (r2 <= r2s[index]) { fiber_index = index; break; }
		}
		if (fiber_index >= n) return;
		const m3f *Tf = &Ts[fiber_index];
		float pz = Tf->a31*dir->x + Tf->a32*dir->y + Tf->a33*dir->z;
		float sensitivity = FP_0;
		fp_lut_sample(j->fp_lut, &luts[fiber_index], fabsf(pz), &sensitivity);
		uint32_t w = (uint32_t)(weight*sensitivity*ACCU_K + FP_0p5);
		if (w > 0) accu_deposit(s, offset + fiber_index, w);
		break;
	}
	case XO_DET_RADIAL: {                              /* mcdetector/radial.py:117-150 */
		const det_radial *d = (const det_radial *)base;
		float dx = pos->x - d->position.x, dy = pos->y - d->position.y;
		float r = m_sqrt(dx*dx + dy*dy);
		if (d->log_scale) r = m_log(s, fmaxf(r, FP_RMIN));
		int32_t ri = (int32_t)((r - d->r_min)*d->inv_dr);
		uint32_t index = (uint32_t)iclip(ri, 0, (int32_t)(d->n - 1));
		p3f dd = d->direction;
		uint32_t w = weight_to_u32(weight, d->cos_min <= fabsf(dot3(dir, &dd)));
		if (w > 0) accu_deposit(s, d->offset + index, w);
		break;
	}
	case XO_DET_CARTESIAN: {                           /* mcdetector/cartesian.py:124-157 */
		const det_cartesian *d = (const det_cartesian *)base;
		int32_t ix = (int32_t)((pos->x - d->x_min)*d->inv_dx);
		ix = iclip(ix, 0, (int32_t)(d->n_x - 1));
		int32_t iy = (int32_t)((pos->y - d->y_min)*d->inv_dy);
		iy = iclip(iy, 0, (int32_t)(d->n_y - 1));
		uint32_t index = (uint32_t)iy*d->n_x + (uint32_t)ix;
		p3f dd = d->direction;
		uint32_t w = weight_to_u32(weight, d->cos_min <= fabsf(dot3(dir, &dd)));
		if (w > 0) accu_deposit(s, index + d->offset, w);
		break;
	}
	case XO_DET_SIXAROUNDONE: {                        /* mcdetector/probe/sixaroundone.py:103-192 */
		const det_six *d = (const det_six *)base;
		uint32_t fiber_index = 7;
		p3f rel = { pos->x - d->position.x, pos->y - d->position.y, FP_0 };
		p3f mc_pos, dp; float dx, dy, r2;
		m3f T = d->T;
		mc_pos.x = rel.x; mc_pos.y = rel.y; mc_pos.z = FP_0;
		transform3(&T, &mc_pos, &dp);
		dx = dp.x; dy = dp.y; r2 = dx*dx + dy*dy;
		if (r2 <= d->core_r_squared) fiber_index = 0;
		mc_pos.x = fabsf(rel.x) - d->core_spacing; mc_pos.y = rel.y;
		transform3(&T, &mc_pos, &dp);
		dx = dp.x; dy = dp.y; r2 = dx*dx + dy*dy;
		if (r2 <= d->core_r_squared) fiber_index = (rel.x >= FP_0) ? 1 : 4;
		mc_pos.x = fabsf(rel.x) - d->core_spacing*FP_0p5;
		mc_pos.y = fabsf(rel.y) - d->core_spacing*FP_COS_30;
		transform3(&T, &mc_pos, &dp);
		dx = dp.x; dy = dp.y; r2 = dx*dx + dy*dy;
		if (r2 <= d->core_r_squared)
			fiber_index = (rel.x >= FP_0) ? ((rel.y >= FP_0) ? 2 : 6) : ((rel.y >= FP_0) ? 3 : 5);
		if (fiber_index > 6) return;
		float pz = T.a31*dir->x + T.a32*dir->y + T.a33*dir->z;
		uint32_t w = weight_to_u32(weight, d->cos_min <= fabsf(pz));
		if (w > 0) accu_deposit(s, d->offset + fiber_index, w);
		break;
	}
	case XO_DET_LINEARARRAY: {                         /* mcdetector/probe/lineararray.py:105-167 */
		const det_linarray *d = (const det_linarray *)base;
		uint32_t n = (uint32_t)param, fiber_index = n;
		float fiber_x = d->first_position.x, fiber_y = d->first_position.y;
		p3f mc_pos, dp; float dx, dy, r2;
		for (uint32_t index = 0; index < n; ++index) {
			mc_pos.x = pos->x - fiber_x; mc_pos.y = pos->y - fiber_y; mc_pos.z = FP_0;
			m3f T = d->T;
			transform3(&T, &mc_pos, &dp);
			dx = dp.x; dy = dp.y; r2 = dx*dx + dy*dy;
			if (r2 <= d->core_r_squared) { fiber_index = index; break; }
			fiber_x += d->delta_position.x;
			fiber_y += d->delta_position.y;
		}
		if (fiber_index >= n) return;
		float pz = d->T.a31*dir->x + d->T.a32*dir->y + d->T.a33*dir->z;
		uint32_t w = weight_to_u32(weight, d->cos_min <= fabsf(pz));
		if (w > 0) accu_deposit(s, d->offset + fiber_index, w);
		break;
	}
	case XO_DET_FIBERARRAY: {                          /* mcdetector/probe/fiberarray.py:101-160 */
		/* packed: m3f T[n]; p2f core_position[n]; float core_r_squared[n], cos_min[n]; u32 offset */
		uint32_t n = (uint32_t)param, fiber_index = n;
		const m3f *Ts = (const m3f *)base;
		const p2f *cp = (const p2f *)(Ts + n);
		const float *r2s = (const float *)(cp + n);
		const float *cmin = r2s + n;
		uint32_t offset = *(const uint32_t *)(cmin + n);
		p3f mc_pos, dp; float dx, dy, r2;
		for (uint32_t index = 0; index < n; ++index) {
			mc_pos.x = pos->x - cp[index].x; mc_pos.y = pos->y - cp[index].y; mc_pos.z = FP_0;
			m3f T = Ts[index];
			transform3(&T, &mc_pos, &dp);
			dx = dp.x; dy = dp.y; r2 = dx*dx + dy*dy;
			if (r2 <= r2s[index]) { fiber_index = index; break; }
		}
		if (fiber_index >= n) return;
		const m3f *Tf = &Ts[fiber_index];
		float pz = Tf->a31*dir->x + Tf->a32*dir->y + Tf->a33*dir->z;
		uint32_t w = weight_to_u32(weight, cmin[fiber_index] <= fabsf(pz));
		if (w > 0) accu_deposit(s, offset + fiber_index, w);
		break;
	}
	case XO_DET_LINEARARRAYPL: {                       /* mcdetector/probe/lineararraypl.py:118-190 */
		const det_linarraypl *d = (const det_linarraypl *)base;
		uint32_t n = (uint32_t)param, fiber_index = n;
		float fiber_x = d->first_position.x, fiber_y = d->first_position.y;
		p3f mc_pos, dp; float dx, dy, r2;
		for (uint32_t index = 0; index < n; ++index) {
			mc_pos.x = pos->x - fiber_x; mc_pos.y = pos->y - fiber_y; mc_pos.z = FP_0;
			m3f T = d->T;
			transform3(&T, &mc_pos, &dp);
			dx = dp.x; dy = dp.y; r2 = dx*dx + dy*dy;
			if (r2 <= d->core_r_squared) { fiber_index = index; break; }
			fiber_x += d->delta_position.x;
			fiber_y += d->delta_position.y;
		}
		if (fiber_index >= n) return;
		float pl = s->opl;
		if (d->pl_log_scale) pl = m_log(s, fmaxf(pl, FP_PLMIN));
		int32_t pl_index = (int32_t)((pl - d->pl_min)*d->inv_dpl);
		pl_index = iclip(pl_index, 0, (int32_t)d->n_pl - 1);
		float pz = d->T.a31*dir->x + d->T.a32*dir->y + d->T.a33*dir->z;
		uint32_t w = weight_to_u32(weight, d->cos_min <= fabsf(pz));
		if (w > 0) accu_deposit(s, d->offset + (size_t)pl_index*n + fiber_index, w);
		break;
	}
	case XO_DET_FIBERARRAYPL: {                        /* mcdetector/probe/fiberarraypl.py:110-185 */
		/* packed: m3f T[n]; p2f core_position[n]; float core_r_squared[n], cos_min[n];
		 * float pl_min, inv_dpl; u32 n_pl, offset; i32 pl_log_scale */
		uint32_t n = (uint32_t)param, fiber_index = n;
		const m3f *Ts = (const m3f *)base;
		const p2f *cp = (const p2f *)(Ts + n);
		const float *r2s = (const float *)(cp + n);
		const float *cmin = r2s + n;
		const float *tail = cmin + n;
		float pl_min = tail[0], inv_dpl = tail[1];
		uint32_t n_pl = ((const uint32_t *)tail)[2], offset = ((const uint32_t *)tail)[3];
		int32_t pl_log_scale = ((const int32_t *)tail)[4];
		p3f mc_pos, dp; float dx, dy, r2;
		for (uint32_t index = 0; index < n; ++index) {
			mc_pos.x = pos->x - cp[index].x; mc_pos.y = pos->y - cp[index].y; mc_pos.z = FP_0;
			m3f T = Ts[index];
			transform3(&T, &mc_pos, &dp);
			dx = dp.x; dy = dp.y; r2 = dx*dx + dy*dy;
			if (r2 <= r2s[index]) { fiber_index = index; break; }
		}
		if (fiber_index >= n) return;
		float pl = s->opl;
		if (pl_log_scale) pl = m_log(s, fmaxf(pl, FP_PLMIN));
		int32_t pl_index = (int32_t)((pl - pl_min)*inv_dpl);
		pl_index = iclip(pl_index, 0, (int32_t)n_pl - 1);
		const m3f *Tf = &Ts[fiber_index];
		float pz = Tf->a31*dir->x + Tf->a32*dir->y + Tf->a33*dir->z;
		uint32_t w = weight_to_u32(weight, cmin[fiber_index] <= fabsf(pz));
		if (w > 0) accu_deposit(s, offset + (size_t)pl_index*n + fiber_index, w);
		break;
	}
	case XO_DET_RADIALPL: {                            /* mcdetector/radialpl.py:138-180 */
		const det_radialpl *d = (const det_radialpl *)base;
		float dx = pos->x - d->position.x, dy = pos->y - d->position.y;
		float r = m_sqrt(dx*dx + dy*dy);
		if (d->r_log_scale) r = m_log(s, fmaxf(r, FP_RMIN));
		int32_t ri = iclip((int32_t)((r - d->r_min)*d->inv_dr), 0, (int32_t)(d->n_r - 1));
		float pl = s->opl;
		if (d->pl_log_scale) pl = m_log(s, fmaxf(pl, FP_PLMIN));
		int32_t pi = iclip((int32_t)((pl - d->pl_min)*d->inv_dpl), 0, (int32_t)(d->n_pl - 1));
		uint32_t index = (uint32_t)pi*d->n_r + (uint32_t)ri;
		p3f dd = d->direction;
		uint32_t w = weight_to_u32(weight, d->cos_min <= fabsf(dot3(dir, &dd)));
		if (w > 0) accu_deposit(s, d->offset + index, w);
		break;
	}
	case XO_DET_TOTALPL: {                             /* mcdetector/totalpl.py:101-121 */
		const det_totalpl *d = (const det_totalpl *)base;
		float pl = s->opl;
		if (d->pl_log_scale) pl = m_log(s, fmaxf(pl, FP_PLMIN));
		int32_t pi = iclip((int32_t)((pl - d->pl_min)*d->inv_dpl), 0, (int32_t)(d->n_pl - 1));
		p3f dd = d->direction;
		uint32_t w = weight_to_u32(weight, d->cos_min <= fabsf(dot3(dir, &dd)));
		if (w > 0) accu_deposit(s, d->offset + (uint32_t)pi, w);
		break;
	}
	case XO_DET_SYMMETRICX: {                          /* mcdetector/symmetric.py:106-140 */
		const det_symx *d = (const det_symx *)base;
		float x = fabsf(pos->x - d->position_x);
		int32_t index_x;
		size_t accu_index;
		p3f dd = d->direction;
		if (d->log_scale) x = m_log(s, fmaxf(x, FP_RMIN));
		index_x = (int32_t)((x - d->x_offset)*d->inv_step);
		index_x = iclip(index_x, 0, (int32_t)d->n_half - 1);
		/* the reference reads the simulator position here (== pos at top/bottom) */
		accu_index = (s->pos.x - d->position_x >= FP_0) ?
			(size_t)index_x + d->n_half : (size_t)d->n_half - index_x - 1;
		{
			uint32_t w = weight_to_u32(weight, d->cos_min <= fabsf(dot3(dir, &dd)));
			if (w > 0) accu_deposit(s, d->offset + accu_index, w);
		}
		break;
	}
	case XO_DET_CARTESIANPL: {                         /* mcdetector/cartesianpl.py:131-175 */
		const det_cartesianpl *d = (const det_cartesianpl *)base;
		int32_t index_x, index_y, index_pl;
		size_t index;
		float pl;
		p3f dd = d->direction;
		index_x = (int32_t)((pos->x - d->x_min)*d->inv_dx);
		index_x = iclip(index_x, 0, (int32_t)d->n_x - 1);
		index_y = (int32_t)((pos->y - d->y_min)*d->inv_dy);
		index_y = iclip(index_y, 0, (int32_t)d->n_y - 1);
		pl = s->opl;
		if (d->pl_log_scale) pl = m_log(s, fmaxf(pl, FP_PLMIN));
		index_pl = (int32_t)((pl - d->pl_min)*d->inv_dpl);
		index_pl = iclip(index_pl, 0, (int32_t)d->n_pl - 1);
		index = ((size_t)index_pl*d->n_y + (size_t)index_y)*d->n_x + (size_t)index_x;
		{
			uint32_t w = weight_to_u32(weight, d->cos_min <= fabsf(dot3(dir, &dd)));
			if (w > 0) accu_deposit(s, d->offset + index, w);
		}
		break;
	}
	case XO_DET_SIXAROUNDONEPL: {                      /* mcdetector/probe/sixaroundonepl.py:128-225 */
		const det_sixpl *d = (const det_sixpl *)base;
		uint32_t fiber_index = 7;
		p3f rel = { pos->x - d->position.x, pos->y - d->position.y, FP_0 };
		p3f mc_pos, dp; float dx, dy, r2, pl;
		int32_t pl_index;
		m3f T = d->T;
		mc_pos.x = rel.x; mc_pos.y = rel.y; mc_pos.z = FP_0;
		transform3(&T, &mc_pos, &dp);
		dx = dp.x; dy = dp.y; r2 = dx*dx + dy*dy;
		if (r2 <= d->core_r_squared) fiber_index = 0;
		mc_pos.x = fabsf(rel.x) - d->core_spacing; mc_pos.y = rel.y;
		transform3(&T, &mc_pos, &dp);
		dx = dp.x; dy = dp.y; r2 = dx*dx + dy*dy;
		if (r2 <= d->core_r_squared) fiber_index = (rel.x >= FP_0) ? 1 : 4;
		mc_pos.x = fabsf(rel.x) - d->core_spacing*FP_0p5;
		mc_pos.y = fabsf(rel.y) - d->core_spacing*FP_COS_30;
		transform3(&T, &mc_pos, &dp);
		dx = dp.x; dy = dp.y; r2 = dx*dx + dy*dy;
		if (r2 <= d->core_r_squared)
			fiber_index = (rel.x >= FP_0) ? ((rel.y >= FP_0) ? 2 : 6) : ((rel.y >= FP_0) ? 3 : 5);
		if (fiber_index > 6) return;
		pl = s->opl;
		if (d->pl_log_scale) pl = m_log(s, fmaxf(pl, FP_PLMIN));
		pl_index = (int32_t)((pl - d->pl_min)*d->inv_dpl);
		pl_index = iclip(pl_index, 0, (int32_t)d->n_pl - 1);
		{
			float pz = T.a31*dir->x + T.a32*dir->y + T.a33*dir->z;
			uint32_t w = weight_to_u32(weight, d->cos_min <= fabsf(pz));
			if (w > 0) accu_deposit(s, d->offset + (size_t)pl_index*7 + fiber_index, w);
		}
		break;
	}
	case XO_DET_FIZ: {                                 /* mccyl/mcdetector/fiz.py:86-120 */
		const det_fiz *d = (const det_fiz *)base;
		float fi = m_atan2(s, pos->y, pos->x);
		int32_t index_fi = iclip((int32_t)((fi - d->fi_min)*d->inv_dfi), 0, (int32_t)(d->n_fi - 1));
		int32_t index_z = iclip((int32_t)((pos->z - d->z_min)*d->inv_dz), 0, (int32_t)(d->n_z - 1));
		uint32_t index = (uint32_t)index_z*d->n_fi + (uint32_t)index_fi;
		float k = m_sqrt(pos->x*pos->x + pos->y*pos->y);
		k = (k > FP_0) ? m_div(FP_1, k) : FP_0;
		p3f normal = { pos->x*k, pos->y*k, FP_0 };
		uint32_t w = weight_to_u32(weight, d->cos_min <= fabsf(dot3(dir, &normal)));
		if (w > 0) accu_deposit(s, index + d->offset, w);
		break;
	}
	case XO_DET_TOTAL_CYL: {                           /* mccyl/mcdetector/total.py:75-100 */
		const det_total_cyl *d = (const det_total_cyl *)base;
		float r = m_sqrt(pos->x*pos->x + pos->y*pos->y);
		float k = (r > FP_0) ? m_div(FP_1, r) : FP_0;
		p3f normal = { pos->x*k, pos->y*k, FP_0 };
		uint32_t w = weight_to_u32(weight, d->cos_min <= fabsf(dot3(dir, &normal)));
		if (w > 0) accu_deposit(s, d->offset, w);
		break;
	}
	default: break;
	}
}

/* ---- fluence --------------------------------------------------------------- */
static void fluence_deposit_at(sim_t *s, const p3f *pos, float weight, float mua) {
	const xo_oracle_job *j = s->job;
	switch (j->fluence_kind) {
	case XO_FLU_XYZ: {                                 /* mcfluence/fluence.py:103-143 */
		const flu_xyz *f = (const flu_xyz *)j->fluence;
		float fx = (pos->x - f->top_left.x)*f->inv_step.x;
		float fy = (pos->y - f->top_left.y)*f->inv_step.y;
		float fz = (pos->z - f->top_left.z)*f->inv_step.z;
		if (fx >= 0 && fy >= 0 && fz >= 0 && fx < f->nx && fy < f->ny && fz < f->nz) {
			uint32_t ix = (uint32_t)fx, iy = (uint32_t)fy, iz = (uint32_t)fz;
			uint32_t index = (iz*f->ny + iy)*f->nx + ix;
			if (j->fluence_rate) weight *= (mua != FP_0) ? m_div(FP_1, mua) : FP_0;
			uint32_t w = (uint32_t)(weight*f->k + FP_0p5);
			accu_deposit(s, f->offset + index, w);
		}
		break;
	}
	case XO_FLU_RZ: {                                  /* mcfluence/fluencerz.py:112-160 */
		const flu_rz *f = (const flu_rz *)j->fluence;
		float dx = pos->x - f->center.x, dy = pos->y - f->center.y;
		float r = m_sqrt(dx*dx + dy*dy);
		float dz = pos->z - f->center.z;
		float fr = r*f->inv_dr, fz = dz*f->inv_dz;
		if (fr >= 0 && fz >= 0 && fr < f->n_r && fz < f->n_z) {
			uint32_t ir = (uint32_t)fr, iz = (uint32_t)fz;
			uint32_t index = iz*f->n_r + ir;
			if (j->fluence_rate) weight *= (mua != FP_0) ? m_div(FP_1, mua) : FP_0;
			uint32_t w = (uint32_t)(weight*f->k + FP_0p5);
			accu_deposit(s, f->offset + index, w);
		}
		break;
	}
	case XO_FLU_XYZT: {                                /* mcfluence/fluencet.py:107-158 */
		const flu_xyzt *f = (const flu_xyzt *)j->fluence;
		float fx = (pos->x - f->top_left[0])*f->inv_step[0];
		float fy = (pos->y - f->top_left[1])*f->inv_step[1];
		float fz = (pos->z - f->top_left[2])*f->inv_step[2];
		float ft = (s->opl*FP_INV_C - f->top_left[3])*f->inv_step[3];
		if (ft >= FP_0 && fx >= FP_0 && fy >= FP_0 && fz >= FP_0 &&
				fx < f->shape[0] && fy < f->shape[1] && fz < f->shape[2] && ft < f->shape[3]) {
			uint32_t ix = (uint32_t)fx, iy = (uint32_t)fy, iz = (uint32_t)fz, it = (uint32_t)ft;
			uint32_t index = ((iz*f->shape[1] + iy)*f->shape[0] + ix)*f->shape[3] + it;
			if (j->fluence_rate) weight *= (mua != FP_0) ? m_div(FP_1, mua) : FP_0;
			uint32_t w = (uint32_t)(weight*f->k + FP_0p5);
			accu_deposit(s, f->offset + index, w);
		}
		break;
	}
	case XO_FLU_RZT: {                                 /* mcfluence/fluencerzt.py:104-150 */
		const flu_rzt *f = (const flu_rzt *)j->fluence;
		float dx = pos->x - f->center.x, dy = pos->y - f->center.y;
		float r = m_sqrt(dx*dx + dy*dy);
		float dz = pos->z - f->center.z;
		float dt = s->opl*FP_INV_C - f->t_min;
		float fr = r*f->inv_dr, fz = dz*f->inv_dz, ft = dt*f->inv_dt;
		if (fr >= 0 && fz >= 0 && ft >= 0 && fr < f->n_r && fz < f->n_z && ft < f->n_t) {
			uint32_t ir = (uint32_t)fr, iz = (uint32_t)fz, it = (uint32_t)ft;
			uint32_t index = (iz*f->n_r + ir)*f->n_t + it;
			if (j->fluence_rate) weight *= (mua != FP_0) ? m_div(FP_1, mua) : FP_0;
			uint32_t w = (uint32_t)(weight*f->k + FP_0p5);
			accu_deposit(s, f->offset + index, w);
		}
		break;
	}
	case XO_FLU_CYL: {                                 /* mcfluence/fluencecyl.py:108-152 */
		const flu_cyl *f = (const flu_cyl *)j->fluence;
		/* (the reference reads the simulator position here, which is the
		 * `position` argument at every call site of AW / AR) */
		float dx = pos->x - f->center.x, dy = pos->y - f->center.y;
		float r = m_sqrt(dx*dx + dy*dy);
		float fi = m_atan2(s, dy, dx) + FP_PI;
		float fr = (r - f->r_min)*f->inv_dr, fz = (pos->z - f->z_min)*f->inv_dz;
		float ffi = (fi - f->fi_min)*f->inv_dfi;
		if (fr >= 0 && fz >= 0 && ffi >= 0 && fr < f->n_r && fz < f->n_z && ffi < f->n_fi) {
			uint32_t ir = (uint32_t)fr, iz = (uint32_t)fz, ifi = (uint32_t)ffi;
			uint32_t index = (iz*f->n_fi + ifi)*f->n_r + ir;
			if (j->fluence_rate) weight *= (mua != FP_0) ? m_div(FP_1, mua) : FP_0;
			uint32_t w = (uint32_t)(weight*f->k + FP_0p5);
			accu_deposit(s, f->offset + index, w);
		}
		break;
	}
	case XO_FLU_CYLT: {                                /* mcfluence/fluencecylt.py:151-204 */
		const flu_cylt *f = (const flu_cylt *)j->fluence;
		float dx = pos->x - f->center.x, dy = pos->y - f->center.y;
		float r = m_sqrt(dx*dx + dy*dy);
		float fi = m_atan2(s, dy, dx) + FP_PI;
		float dt = s->opl*FP_INV_C - f->t_min;
		float fr = (r - f->r_min)*f->inv_dr, fz = (pos->z - f->z_min)*f->inv_dz;
		float ffi = (fi - f->fi_min)*f->inv_dfi, ft = dt*f->inv_dt;
		if (fr >= 0 && fz >= 0 && ffi >= 0 && ft >= 0 && fr < f->n_r && fz < f->n_z &&
				ffi < f->n_fi && ft < f->n_t) {
			uint32_t ir = (uint32_t)fr, iz = (uint32_t)fz, ifi = (uint32_t)ffi, it = (uint32_t)ft;
			uint32_t index = ((iz*f->n_fi + ifi)*f->n_r + ir)*f->n_t + it;
			if (j->fluence_rate) weight *= (mua != FP_0) ? m_div(FP_1, mua) : FP_0;
			uint32_t w = (uint32_t)(weight*f->k + FP_0p5);
			accu_deposit(s, f->offset + index, w);
		}
		break;
	}
	default: break;
	}
}

/* ---- trace: mcbase/mctrace.py:541-585 -------------------------------------- */
static int trace_event(sim_t *s, uint32_t event_count) {
	const trace_cfg *t = (const trace_cfg *)s->job->trace;
	/* mc_min(event_count, max_events - 1) on mc_int_t */
	int32_t ec = (int32_t)event_count, me = t->max_events - 1;
	uint32_t pos = (uint32_t)(ec < me ? ec : me)*8u +
		s->photon_index*(uint32_t)t->max_events*8u + t->data_off;
	if (s->job->use_events && !(t->event_mask & s->event_flags))
		return 0;
	float *fb = s->job->float_buffer;
	fb[pos++] = s->pos.x; fb[pos++] = s->pos.y; fb[pos++] = s->pos.z;
	fb[pos++] = s->dir.x; fb[pos++] = s->dir.y; fb[pos++] = s->dir.z;
	fb[pos++] = s->weight;
	fb[pos++] = s->job->track_opl ? s->opl : FP_0;
	return 1;
}
static inline void trace_this_event(sim_t *s) {          /* mcml.template.c:255-258 */
	if (trace_event(s, s->trace_count)) s->trace_count++;
}
static inline void trace_finalize(sim_t *s) {            /* mcml.template.c:263-265 */
	const trace_cfg *t = (const trace_cfg *)s->job->trace;
	s->job->int_buffer[t->count_off + s->photon_index] = (int32_t)s->trace_count;
}

/* ---- sources (mcml) -------------------------------------------------------- */
static void launch_mcml(sim_t *s) {
	const xo_oracle_job *j = s->job;
	switch (j->src_kind) {
	case XO_SRC_LINE: {                                /* mcsource/line.py:95-110 */
		const src_line *src = (const src_line *)j->source;
		s->weight = FP_1 - src->reflectance;
		s->pos = src->position;
		s->dir = src->dir_sample;
		if (j->det_kind[LOC_SPECULAR]) {
			p3f d = src->dir_reflected;
			detector_deposit(s, LOC_SPECULAR, &s->pos, &d, src->reflectance);
		}
		s->layer_index = 1;
		break;
	}
	case XO_SRC_GAUSSIANBEAM: {                        /* mcsource/gaussianbeam.py:119-165 */
		const src_gauss_ml *src = (const src_gauss_ml *)j->source;
		float cos_fi, sin_fi, r; p3f pt_src, pt_mc;
		r = m_sqrt(-FP_2*m_log(s, FP_1 - sim_random(s)));
		r = fminf(r, src->clip);
		m_sincos(s, FP_2PI*sim_random(s), &sin_fi, &cos_fi);
		pt_src.x = r*cos_fi*src->sigma.x;
		pt_src.y = r*sin_fi*src->sigma.y;
		pt_src.z = FP_0;
		m3f T = src->T;
		transform3(&T, &pt_src, &pt_mc);
		float k = m_div(FP_0 - pt_mc.z, src->direction.z);
		pt_mc.x += k*src->direction.x;
		pt_mc.y += k*src->direction.y;
		pt_mc.z = FP_0;
		s->pos.x = src->position.x + pt_mc.x;
		s->pos.y = src->position.y + pt_mc.y;
		s->pos.z = FP_0;
		s->dir = src->direction;
		s->layer_index = 1;
		s->weight = FP_1 - src->reflectance;
		if (j->det_kind[LOC_SPECULAR]) {
			p3f dir_in = { s->dir.x, s->dir.y, -s->dir.z }, dir;
			p3f normal = { FP_0, FP_0, -FP_1 };
			refract3(&dir_in, &normal, medium_n(j, 1), medium_n(j, 0), &dir);
			detector_deposit(s, LOC_SPECULAR, &s->pos, &dir, src->reflectance);
		}
		break;
	}
	case XO_SRC_UNIFORMBEAM: {                         /* mcsource/uniformbeam.py:75-112 */
		const src_ubeam_ml *src = (const src_ubeam_ml *)j->source;
		float cos_fi, sin_fi, rand_sqrt, fi; p3f pt_src, pt_mc;
		rand_sqrt = m_sqrt(sim_random(s));
		fi = FP_2PI*sim_random(s);
		m_sincos(s, fi, &sin_fi, &cos_fi);     /* mc_cos(fi), mc_sin(fi) */
		pt_src.x = rand_sqrt*cos_fi*src->radius.x;
		pt_src.y = rand_sqrt*sin_fi*src->radius.y;
		pt_src.z = FP_0;
		{
			m3f T = src->T;
			float k;
			transform3(&T, &pt_src, &pt_mc);
			k = m_div(FP_0 - pt_mc.z, src->direction.z);
			pt_mc.x += k*src->direction.x;
			pt_mc.y += k*src->direction.y;
			pt_mc.z = FP_0;
		}
		s->pos.x = src->position.x + pt_mc.x;
		s->pos.y = src->position.y + pt_mc.y;
		s->pos.z = FP_0;
		s->dir = src->direction;
		s->weight = FP_1 - src->reflectance;
		s->layer_index = 1;
		if (j->det_kind[LOC_SPECULAR]) {
			p3f dir_in = { s->dir.x, s->dir.y, -s->dir.z }, dir;
			p3f normal = { FP_0, FP_0, -FP_1 };
			refract3(&dir_in, &normal, medium_n(j, 1), medium_n(j, 0), &dir);
			detector_deposit(s, LOC_SPECULAR, &s->pos, &dir, src->reflectance);
		}
		break;
	}
	case XO_SRC_UNIFORMFIBER: {                        /* mcsource/fiber.py:273-331 */
		const src_ufiber *src = (const src_ufiber *)j->source;
		float sin_fi, cos_fi, sin_theta, cos_theta; p3f pt_src, pt_mc;
		float r = m_sqrt(sim_random(s))*src->radius;
		m_sincos(s, sim_random(s)*FP_2PI, &sin_fi, &cos_fi);
		pt_src.x = r*cos_fi; pt_src.y = r*sin_fi; pt_src.z = FP_0;
		m3f T = src->T;
		transform3(&T, &pt_src, &pt_mc);
		float k = m_div(FP_0 - pt_mc.z, src->direction.z);
		pt_mc.x += k*src->direction.x;
		pt_mc.y += k*src->direction.y;
		pt_mc.z = FP_0;
		s->pos.x = src->position.x + pt_mc.x;
		s->pos.y = src->position.y + pt_mc.y;
		s->pos.z = FP_0;
		m_sincos(s, sim_random(s)*FP_2PI, &sin_fi, &cos_fi);
		cos_theta = FP_1 - sim_random(s)*(FP_1 - src->cos_min);
		sin_theta = m_sqrt(FP_1 - cos_theta*cos_theta);
		sin_theta = m_div(sin_theta, src->n);
		cos_theta = m_sqrt(FP_1 - sin_theta*sin_theta);
		pt_src.x = cos_fi*sin_theta; pt_src.y = sin_fi*sin_theta; pt_src.z = cos_theta;
		p3f direction;
		transform3(&T, &pt_src, &direction);
		float cc = cos_critical(src->n, medium_n(j, 1));
		p3f normal = { FP_0, FP_0, FP_1 };
		p3f refracted = direction;
		if (pt_mc.z > cc)             /* never true: pt_mc.z == 0 */
			refract3(&pt_mc, &normal, src->n, medium_n(j, 1), &refracted);
		s->dir = refracted;
		float specular_r = reflectance(src->n, medium_n(j, 1), direction.z, cc);
		s->weight = FP_1 - specular_r;
		if (j->det_kind[LOC_SPECULAR])
			detector_deposit(s, LOC_SPECULAR, &s->pos, &direction, specular_r);
		s->layer_index = 1;
		break;
	}
	case XO_SRC_LAMBERTIANFIBER: {                     /* mcsource/fiber.py:567-632 */
		const src_ufiber *src = (const src_ufiber *)j->source;    /* cos_min slot holds na */
		float sin_fi, cos_fi, sin_theta, cos_theta; p3f pt_src, pt_mc;
		float r = m_sqrt(sim_random(s))*src->radius;
		m_sincos(s, sim_random(s)*FP_2PI, &sin_fi, &cos_fi);
		pt_src.x = r*cos_fi; pt_src.y = r*sin_fi; pt_src.z = FP_0;
		m3f T = src->T;
		transform3(&T, &pt_src, &pt_mc);
		float k = m_div(FP_0 - pt_mc.z, src->direction.z);
		pt_mc.x += k*src->direction.x;
		pt_mc.y += k*src->direction.y;
		pt_mc.z = FP_0;
		s->pos.x = src->position.x + pt_mc.x;
		s->pos.y = src->position.y + pt_mc.y;
		s->pos.z = FP_0;
		m_sincos(s, sim_random(s)*FP_2PI, &sin_fi, &cos_fi);
		sin_theta = m_sqrt(sim_random(s))*src->cos_min;
		sin_theta = m_div(sin_theta, src->n);
		cos_theta = m_sqrt(FP_1 - sin_theta*sin_theta);
		pt_src.x = cos_fi*sin_theta; pt_src.y = sin_fi*sin_theta; pt_src.z = cos_theta;
		p3f direction;
		transform3(&T, &pt_src, &direction);
		float cc = cos_critical(src->n, medium_n(j, 1));
		p3f normal = { FP_0, FP_0, FP_1 };
		p3f refracted = direction;
		if (pt_mc.z > cc)             /* never true: pt_mc.z == 0 */
			refract3(&pt_mc, &normal, src->n, medium_n(j, 1), &refracted);
		s->dir = refracted;
		float specular_r = reflectance(src->n, medium_n(j, 1), direction.z, cc);
		s->weight = FP_1 - specular_r;
		if (j->det_kind[LOC_SPECULAR])
			detector_deposit(s, LOC_SPECULAR, &s->pos, &direction, specular_r);
		s->layer_index = 1;
		break;
	}
	case XO_SRC_UNIFORMFIBERLUT: {                     /* mcsource/fiber.py:766-830 */
		const src_ufiberlut *src = (const src_ufiberlut *)j->source;
		float sin_fi, cos_fi, sin_theta, cos_theta = FP_0; p3f pt_src, pt_mc;
		float r = m_sqrt(sim_random(s))*src->radius;
		m_sincos(s, sim_random(s)*FP_2PI, &sin_fi, &cos_fi);
		pt_src.x = r*cos_fi; pt_src.y = r*sin_fi; pt_src.z = FP_0;
		m3f T = src->T;
		transform3(&T, &pt_src, &pt_mc);
		float k = m_div(FP_0 - pt_mc.z, src->direction.z);
		pt_mc.x += k*src->direction.x;
		pt_mc.y += k*src->direction.y;
		pt_mc.z = FP_0;
		s->pos.x = src->position.x + pt_mc.x;
		s->pos.y = src->position.y + pt_mc.y;
		s->pos.z = FP_0;
		m_sincos(s, sim_random(s)*FP_2PI, &sin_fi, &cos_fi);
		fp_lut_rel_sample(j->fp_lut, &src->lut, sim_random(s), &cos_theta);
		sin_theta = m_sqrt(FP_1 - cos_theta*cos_theta);
		sin_theta = m_div(sin_theta, src->n);
		cos_theta = m_sqrt(FP_1 - sin_theta*sin_theta);
		pt_src.x = cos_fi*sin_theta; pt_src.y = sin_fi*sin_theta; pt_src.z = cos_theta;
		p3f direction;
		transform3(&T, &pt_src, &direction);
		float cc = cos_critical(src->n, medium_n(j, 1));
		p3f normal = { FP_0, FP_0, FP_1 };
		p3f refracted = direction;
		if (direction.z > cc)
			refract3(&direction, &normal, src->n, medium_n(j, 1), &refracted);
		s->dir = refracted;
		float specular_r = reflectance(src->n, medium_n(j, 1), direction.z, cc);
		s->weight = FP_1 - specular_r;
		if (j->det_kind[LOC_SPECULAR])
			detector_deposit(s, LOC_SPECULAR, &s->pos, &direction, specular_r);
		s->layer_index = 1;
		break;
	}
	case XO_SRC_UNIFORMFIBERNI:                        /* mcsource/fiberni.py:242-293 */
	case XO_SRC_LAMBERTIANFIBERNI:                     /* mcsource/fiberni.py:489-534 */
	case XO_SRC_UNIFORMFIBERLUTNI: {                   /* mcsource/fiberni.py:642-696 */
		const src_fiberni *src = (const src_fiberni *)j->source;
		const src_fiberlutni *srcl = (const src_fiberlutni *)j->source;
		const int is_lut = j->src_kind == XO_SRC_UNIFORMFIBERLUTNI;
		const float n_core = is_lut ? srcl->n : src->n;
		float sin_fi, cos_fi, sin_theta, cos_theta = FP_0;
		float r = m_sqrt(sim_random(s))*src->radius;
		m_sincos(s, sim_random(s)*FP_2PI, &sin_fi, &cos_fi);
		s->pos.x = src->position.x + r*cos_fi;
		s->pos.y = src->position.y + r*sin_fi;
		s->pos.z = FP_0;
		m_sincos(s, sim_random(s)*FP_2PI, &sin_fi, &cos_fi);
		if (is_lut) {
			fp_lut_rel_sample(j->fp_lut, &srcl->lut, sim_random(s), &cos_theta);
			sin_theta = m_sqrt(FP_1 - cos_theta*cos_theta);
		} else if (j->src_kind == XO_SRC_LAMBERTIANFIBERNI) {
			sin_theta = m_sqrt(sim_random(s))*src->aperture;
		} else {
			cos_theta = FP_1 - sim_random(s)*(FP_1 - src->aperture);
			sin_theta = m_sqrt(FP_1 - cos_theta*cos_theta);
		}
		/* emission angle adjusted to the refractive index of the sample */
		sin_theta = m_div(sin_theta, medium_n(j, 1));
		cos_theta = m_sqrt(FP_1 - sin_theta*sin_theta);
		s->dir.x = cos_fi*sin_theta; s->dir.y = sin_fi*sin_theta; s->dir.z = cos_theta;
		r = reflectance_cos2(n_core, medium_n(j, 1), cos_theta);
		s->weight = FP_1 - r;
		if (j->det_kind[LOC_SPECULAR]) {
			p3f dir_in = { s->dir.x, s->dir.y, -s->dir.z }, dir;
			p3f normal = { FP_0, FP_0, -FP_1 };
			refract3(&dir_in, &normal, medium_n(j, 1), n_core, &dir);
			detector_deposit(s, LOC_SPECULAR, &s->pos, &dir, r);
		}
		s->layer_index = 1;
		break;
	}
	case XO_SRC_UNIFORMRECTANGULAR:                    /* mcsource/rectangular.py:91-146 */
	case XO_SRC_LAMBERTIANRECTANGULAR: {               /* mcsource/rectangular.py:376-430 */
		const src_rect *src = (const src_rect *)j->source;
		float sin_fi, cos_fi, sin_theta, cos_theta;
		s->pos.x = src->position.x + (sim_random(s) - FP_0p5)*src->size.x;
		s->pos.y = src->position.y + (sim_random(s) - FP_0p5)*src->size.y;
		s->pos.z = src->position.z;
		m_sincos(s, sim_random(s)*FP_2PI, &sin_fi, &cos_fi);
		if (j->src_kind == XO_SRC_LAMBERTIANRECTANGULAR) {
			sin_theta = m_sqrt(sim_random(s))*src->aperture;
		} else {
			cos_theta = FP_1 - sim_random(s)*(FP_1 - src->aperture);
			sin_theta = m_sqrt(FP_1 - cos_theta*cos_theta);
		}
		sin_theta = m_div(sin_theta, medium_n(j, (int32_t)src->layer_index));
		cos_theta = m_sqrt(FP_1 - sin_theta*sin_theta);
		s->dir.x = cos_fi*sin_theta; s->dir.y = sin_fi*sin_theta; s->dir.z = cos_theta;
		float r = reflectance_cos2(src->n, medium_n(j, (int32_t)src->layer_index), cos_theta);
		s->weight = FP_1 - r;
		/* (the specular branch of the reference does not compile: no such case) */
		s->layer_index = (int32_t)src->layer_index;
		break;
	}
	case XO_SRC_UNIFORMRECTANGULARLUT: {               /* mcsource/rectangular.py:600-660 */
		const src_rectlut *src = (const src_rectlut *)j->source;
		float sin_fi, cos_fi, sin_theta, cos_theta = FP_0;
		s->pos.x = src->position.x + (sim_random(s) - FP_0p5)*src->size.x;
		s->pos.y = src->position.y + (sim_random(s) - FP_0p5)*src->size.y;
		s->pos.z = src->position.z;
		m_sincos(s, sim_random(s)*FP_2PI, &sin_fi, &cos_fi);
		fp_lut_rel_sample(j->fp_lut, &src->lut, sim_random(s), &cos_theta);
		sin_theta = m_div(m_sqrt(FP_1 - cos_theta*cos_theta),
			medium_n(j, (int32_t)src->layer_index));
		cos_theta = m_sqrt(FP_1 - sin_theta*sin_theta);
		s->dir.x = cos_fi*sin_theta; s->dir.y = sin_fi*sin_theta; s->dir.z = cos_theta;
		float r = reflectance_cos2(src->n, medium_n(j, (int32_t)src->layer_index), cos_theta);
		s->weight = FP_1 - r;
		/* (the specular branch of the reference does not compile: no such case) */
		s->layer_index = (int32_t)src->layer_index;
		break;
	}
	case XO_SRC_ISOTROPICPOINT: {                      /* mcsource/point.py:76-135 */
		const src_isopoint *src = (const src_isopoint *)j->source;
		float sin_fi, cos_fi, sin_theta, cos_theta, specular_r = FP_0;
		p3f position = src->position, direction;
		m_sincos(s, sim_random(s)*FP_2PI, &sin_fi, &cos_fi);
		cos_theta = FP_1 - FP_2*sim_random(s);
		sin_theta = m_sqrt(FP_1 - cos_theta*cos_theta);
		direction.x = cos_fi*sin_theta; direction.y = sin_fi*sin_theta; direction.z = cos_theta;
		p3f refracted = direction;
		if (position.z <= FP_0) {
			float cc = ml_layer_at(j, 0)->cc_bottom;
			if (direction.z > cc) {
				p3f normal = { FP_0, FP_0, FP_1 };
				float n_sample = medium_n(j, 1), n_medium = medium_n(j, 0);
				refract3(&direction, &normal, n_sample, n_medium, &refracted);
				specular_r = reflectance(n_medium, n_sample, direction.z, cc);
				float t = m_div(-position.z, direction.z);
				position.x += direction.x*t;
				position.y += direction.y*t;
			} else {
				specular_r = FP_1;
			}
			position.z = FP_0;
		}
		s->dir = refracted;
		s->pos = position;
		if (j->det_kind[LOC_SPECULAR])
			detector_deposit(s, LOC_SPECULAR, &s->pos, &direction, specular_r);
		s->weight = FP_1 - specular_r;
		s->layer_index = (int32_t)src->layer_index;
		break;
	}
	default: break;
	}
}

/* `source->position` (mcml.template.c:380): byte offset depends on the source struct */
static inline p3f source_position(const xo_oracle_job *j) {
	size_t off = 0;
	switch (j->src_kind) {
		case XO_SRC_GAUSSIANBEAM: off = sizeof(m3f); break;
		case XO_SRC_UNIFORMFIBER: off = sizeof(m3f); break;
		case XO_SRC_LAMBERTIANFIBER: off = sizeof(m3f); break;
		case XO_SRC_UNIFORMFIBERLUT: off = sizeof(m3f); break;
		case XO_SRC_UNIFORMBEAM: off = sizeof(m3f); break;
		default: off = 0; break;
	}
	return *(const p3f *)((const char *)j->source + off);
}

/* ---- mcml boundary: mcml.template.c:80-203 --------------------------------- */
/* ---- sample surface layouts (mcml) ----------------------------------------- */
/* mcsurface/lambertian.py: struct {reflectance, specular} */
typedef struct { float reflectance, specular; } surf_lambertian;
/* mcsurface/probe/sixaroundone.py:87-103 */
typedef struct { m3f T; p2f position; float core_spacing;
	float cladding_r_squared, cladding_n, cladding_cc;
	float core_r_squared, core_n, core_cc;
	float cutout_r_squared, cutout_n, cutout_cc;
	float probe_r_squared, probe_reflectivity; } surf_six;
/* mcsurface/probe/lineararray.py:44-64 */
typedef struct { m3f T; float c11, c12, c21, c22; p2f position, first_position, delta_position;
	float core_spacing;
	float cladding_r_squared, cladding_n, cladding_cc;
	float core_r_squared, core_n, core_cc;
	float cutout_width_half, cutout_height_half, cutout_n, cutout_cc;
	float probe_r_squared, probe_reflectivity; } surf_linarray;
enum { SURF_CONTINUE = 0, SURF_REFLECTED = 1 };

static int surf_six_fiber(const surf_six *l, float r2, float *n2, float *cc) {
	if (r2 <= l->cladding_r_squared) {
		if (r2 <= l->core_r_squared) { *n2 = l->core_n; *cc = l->core_cc; return 1; }
		*n2 = l->cladding_n; *cc = l->cladding_cc;
		return 1;
	}
	return 0;
}

/* mcsim_{top,bottom}_surface_layout_handler of the plugin at surface `which` */
static int surface_layout_handler(sim_t *s, int which, float *n2, float *cc) {
	const xo_oracle_job *j = s->job;
	const char *base = (const char *)j->surface + j->surf_offset[which];
	switch (j->surf_kind[which]) {
	case XO_SURF_LAMBERTIAN: {                         /* mcsurface/lambertian.py:78-113 */
		const surf_lambertian *l = (const surf_lambertian *)base;
		float sin_fi, cos_fi, sin_theta, cos_theta;
		if (sim_random(s) > l->specular) {
			sin_theta = m_sqrt(sim_random(s));
			cos_theta = m_sqrt(FP_1 - sin_theta*sin_theta);
			m_sincos(s, sim_random(s)*FP_2PI, &sin_fi, &cos_fi);
			{
				float z = fsign(-s->dir.z)*cos_theta;
				s->dir.x = cos_fi*sin_theta;
				s->dir.y = sin_fi*sin_theta;
				s->dir.z = z;
			}
		} else {
			s->dir.z = -s->dir.z;
		}
		s->weight = s->weight*l->reflectance;
		return SURF_REFLECTED;
	}
	case XO_SURF_SIXAROUNDONE: {                       /* mcsurface/probe/sixaroundone.py:156-290 */
		const surf_six *l = (const surf_six *)base;
		p3f rel = { s->pos.x - l->position.x, s->pos.y - l->position.y, FP_0 };
		p3f mc_pos, lp; float dx, dy, r2;
		m3f T = l->T;
		mc_pos.x = rel.x; mc_pos.y = rel.y; mc_pos.z = FP_0;
		transform3(&T, &mc_pos, &lp);
		dx = lp.x; dy = lp.y; r2 = dx*dx + dy*dy;
		if (surf_six_fiber(l, r2, n2, cc)) return SURF_CONTINUE;
		mc_pos.x = fabsf(rel.x) - l->core_spacing; mc_pos.y = rel.y;
		transform3(&T, &mc_pos, &lp);
		dx = lp.x; dy = lp.y; r2 = dx*dx + dy*dy;
		if (surf_six_fiber(l, r2, n2, cc)) return SURF_CONTINUE;
		mc_pos.x = fabsf(rel.x) - l->core_spacing*FP_0p5;
		mc_pos.y = fabsf(rel.y) - l->core_spacing*FP_COS_30;
		transform3(&T, &mc_pos, &lp);
		dx = lp.x; dy = lp.y; r2 = dx*dx + dy*dy;
		if (surf_six_fiber(l, r2, n2, cc)) return SURF_CONTINUE;
		dx = rel.x; dy = rel.y; r2 = dx*dx + dy*dy;
		if (r2 <= l->cutout_r_squared) { *n2 = l->cutout_n; *cc = l->cutout_cc; return SURF_CONTINUE; }
		if (r2 <= l->probe_r_squared) {
			s->dir.z = -s->dir.z;
			s->weight = s->weight*l->probe_reflectivity;
			return SURF_REFLECTED;
		}
		return SURF_CONTINUE;
	}
	case XO_SURF_LINEARARRAY: {                        /* mcsurface/probe/lineararray.py:152-228 */
		const surf_linarray *l = (const surf_linarray *)base;
		uint32_t n = (uint32_t)j->surf_param[which];
		float dx, dy, r2;
		p3f mc_pos = { FP_0, FP_0, FP_0 }, lp;
		float fiber_x = l->first_position.x, fiber_y = l->first_position.y;
		for (uint32_t index = 0; index < n; ++index) {
			mc_pos.x = s->pos.x - fiber_x; mc_pos.y = s->pos.y - fiber_y; mc_pos.z = FP_0;
			m3f T = l->T;
			transform3(&T, &mc_pos, &lp);
			dx = lp.x; dy = lp.y; r2 = dx*dx + dy*dy;
			if (r2 <= l->cladding_r_squared) {
				if (r2 <= l->core_r_squared) { *n2 = l->core_n; *cc = l->core_cc; return SURF_CONTINUE; }
				*n2 = l->cladding_n; *cc = l->cladding_cc;
				return SURF_CONTINUE;
			}
			fiber_x += l->delta_position.x;
			fiber_y += l->delta_position.y;
		}
		/* the cut-out of the probe, in the frame of the array */
		{
			float cx = s->pos.x - l->position.x, cy = s->pos.y - l->position.y;
			float ox = l->c11*cx + l->c12*cy, oy = l->c21*cx + l->c22*cy;
			dx = fabsf(ox); dy = fabsf(oy);
			if (dx <= l->cutout_width_half && dy < l->cutout_height_half) {
				*n2 = l->cutout_n; *cc = l->cutout_cc;
				return SURF_CONTINUE;
			}
		}
		/* the probe tip: relative to the LAST fiber (mc_pos survives the loop) */
		dx = mc_pos.x; dy = mc_pos.y; r2 = dx*dx + dy*dy;
		if (r2 <= l->probe_r_squared) {
			s->dir.z = -s->dir.z;
			s->weight = s->weight*l->probe_reflectivity;
			return SURF_REFLECTED;
		}
		return SURF_CONTINUE;
	}
	case XO_SURF_FIBERARRAY: {                         /* mcsurface/probe/fiberarray.py:133-183 */
		/* packed: m3f T[n]; p2f fiber_position[n]; float cladding_r2[n], cladding_n[n],
		 * cladding_cc[n], core_r2[n], core_n[n], core_cc[n]; p2f probe_position;
		 * float probe_r_squared, probe_reflectivity */
		uint32_t n = (uint32_t)j->surf_param[which];
		const m3f *Ts = (const m3f *)base;
		const p2f *fp = (const p2f *)(Ts + n);
		const float *clad_r2 = (const float *)(fp + n), *clad_n = clad_r2 + n, *clad_cc = clad_n + n;
		const float *core_r2 = clad_cc + n, *core_n = core_r2 + n, *core_cc = core_n + n;
		const float *tail = core_cc + n;   /* probe_position.x, .y, probe_r_squared, reflectivity */
		float dx, dy, r2; p3f mc_pos, lp;
		for (uint32_t index = 0; index < n; ++index) {
			mc_pos.x = s->pos.x - fp[index].x; mc_pos.y = s->pos.y - fp[index].y; mc_pos.z = FP_0;
			m3f T = Ts[index];
			transform3(&T, &mc_pos, &lp);
			dx = lp.x; dy = lp.y; r2 = dx*dx + dy*dy;
			if (r2 <= clad_r2[index]) {
				if (r2 <= core_r2[index]) { *n2 = core_n[index]; *cc = core_cc[index]; return SURF_CONTINUE; }
				*n2 = clad_n[index]; *cc = clad_cc[index];
				return SURF_CONTINUE;
			}
		}
		dx = s->pos.x - tail[0]; dy = s->pos.y - tail[1]; r2 = dx*dx + dy*dy;
		if (r2 <= tail[2]) {
			s->dir.z = -s->dir.z;
			s->weight = s->weight*tail[3];
			return SURF_REFLECTED;
		}
		return SURF_CONTINUE;
	}
	}
	return SURF_CONTINUE;
}

static uint32_t boundary_mcml(sim_t *s, int32_t next_index) {
	const xo_oracle_job *j = s->job;
	const ml_layer *cur = ml_layer_at(j, s->layer_index);
	const ml_layer *nxt = ml_layer_at(j, next_index);
	float cos_crit, cos1, sin1, cos2, sin2, n_cos1, n_cos2, n2, n1_d_n2, R, Rs, Rp;
	p3f *dir = &s->dir;
	cos_crit = (dir->z < FP_0) ? cur->cc_top : cur->cc_bottom;
	n2 = nxt->n;
	/* surface layouts (mcml.template.c:100-126) */
	if (j->surf_kind[0] != XO_SURF_NONE && next_index == 0) {
		if (surface_layout_handler(s, 0, &n2, &cos_crit) != SURF_CONTINUE) return EV_REFLECTION;
	} else if (j->surf_kind[1] != XO_SURF_NONE && next_index == (int32_t)j->num_layers - 1) {
		if (surface_layout_handler(s, 1, &n2, &cos_crit) != SURF_CONTINUE) return EV_REFLECTION;
	}
	if (cur->n == n2) {
		s->layer_index = next_index;
		return EV_REFRACTION;
	}
	dir->z = -dir->z;
	cos1 = fabsf(dir->z);
	if (cos1 > cos_crit) {
		if (cos1 > FP_COS_0) {                 /* unreachable: FP_COS_0 == 1 */
			R = m_div(cur->n - n2, cur->n + n2);
			if (R*R < sim_random(s)) {
				s->layer_index = next_index;
				dir->z = -dir->z;
				return EV_REFRACTION;
			} else {
				return EV_REFLECTION;
			}
		}
		n1_d_n2 = m_div(cur->n, n2);
		sin1 = m_sqrt(FP_1 - cos1*cos1);
		if (cos1 >= FP_COS_0) sin1 = FP_0;
		sin2 = fminf(FP_1, n1_d_n2*sin1);
		cos2 = m_sqrt(FP_1 - sin2*sin2);
		n_cos1 = n1_d_n2*cos1;
		n_cos2 = n1_d_n2*cos2;
		Rs = m_div(n_cos1 - cos2, n_cos1 + cos2); Rs *= Rs;
		Rp = m_div(n_cos2 - cos1, n_cos2 + cos1); Rp *= Rp;
		R = FP_0p5*(Rs + Rp);
		if (cos1 <= FP_COS_90 || sin2 == FP_1) R = FP_1;
		if (R < sim_random(s)) {
			s->layer_index = next_index;
			dir->x *= n1_d_n2;
			dir->y *= n1_d_n2;
			dir->z = -copysignf(cos2, dir->z);
			return EV_REFRACTION;
		}
	}
	return EV_REFLECTION;
}

static inline void sim_scatter(sim_t *s) {               /* mcml.template.c:277-290 */
	if (s->job->pf_kind == XO_PF_HGDIR) {                /* MC_PF_SAMPLE_DIRECTION: mcpf/hgdir.py:95-118 */
		const pf_hgdir *pf = (const pf_hgdir *)current_pf(s);
		float g = pf->g, k, cos_theta;
		p3f out_dir;
		k = m_div(FP_1 - g*g, FP_1 + g*(FP_2*sim_random(s) - FP_1));
		cos_theta = m_div(FP_1 + g*g - k*k, FP_2*g);
		if (g == FP_0) cos_theta = FP_1 - FP_2*sim_random(s);
		cos_theta = fmaxf(fminf(cos_theta, FP_1), -FP_1);
		out_dir = (sim_random(s) < pf->p) ? pf->direction : s->dir;
		cos_theta = (dot3(&s->dir, &out_dir) < FP_0) ? -cos_theta : cos_theta;
		scatter_direction(s, &out_dir, cos_theta, FP_2PI*sim_random(s));
		s->dir = out_dir;
		return;
	}
	float fi, cos_theta = pf_sample_angles(s, &fi);
	scatter_direction(s, &s->dir, cos_theta, fi);
}

/* survival lottery, mcml.template.c:737-753 */
static inline void lottery(sim_t *s, int *done) {
	const xo_oracle_job *j = s->job;
	if (s->weight < j->weight_min) {
		if (j->use_lottery) {
			if (sim_random(s) > j->lottery_chance) *done = 1;
			else s->weight = m_div(s->weight, j->lottery_chance);
		} else {
			*done = 1;
		}
	}
}

/* next packet index for this work-item, or 0 when exhausted */
typedef struct { uint32_t next, end; } quota_t;
static inline int next_packet(sim_t *s, quota_t *q) {
	if (s->dyn_counter) {
		s->photon_index = __atomic_fetch_add(s->dyn_counter, 1u, __ATOMIC_RELAXED);
		return s->photon_index < s->job->num_packets;
	}
	if (q->next >= q->end) return 0;
	s->photon_index = q->next++;
	return 1;
}

static void fluence_deposit_weight(sim_t *s, const p3f *pos, float deposit, float mua) {
	if (s->job->fluence_kind) fluence_deposit_at(s, pos, deposit, mua);
}

/* ---- mcml work-item: mcml.template.c:346-824 -------------------------------- */
static void workitem_mcml(sim_t *s, quota_t *q) {
	const xo_oracle_job *j = s->job;
	const p3f src_pos = source_position(j);
	float step, deposit, rmax = j->rmax;
	int32_t next_index;
	int done = 0;
	const int tr = j->trace_flags;

	if (!next_packet(s, q)) return;
	s->opl = FP_0; s->trace_count = 0; s->event_flags = 0;
	launch_mcml(s);
	s->event_flags |= EV_LAUNCH;
	if (tr & XO_TRACE_START) trace_this_event(s);

	while (!done) {
		const ml_layer *L = ml_layer_at(j, s->layer_index);
		s->iterations++;
		if (j->method == XO_METHOD_MBL)
			step = m_div(-m_log(s, sim_random(s)), ml_mus(s, L));
		else
			step = -m_log(s, sim_random(s))*ml_inv_mut(s, L);
		step = fminf(step, FLT_MAX);
		next_index = s->layer_index;
		if (s->pos.z + step*s->dir.z < L->top) {
			--next_index;
			if (fabsf(s->dir.z) != FP_0) step = m_div(L->top - s->pos.z, s->dir.z);
		}
		if (s->pos.z + step*s->dir.z >= L->bottom) {
			++next_index;
			if (fabsf(s->dir.z) != FP_0) step = m_div(L->bottom - s->pos.z, s->dir.z);
		}
		s->pos.x = s->pos.x + s->dir.x*step;
		s->pos.y = s->pos.y + s->dir.y*step;
		s->pos.z = s->pos.z + s->dir.z*step;
		if (j->track_opl) s->opl += L->n*step;
		if (s->layer_index < next_index) s->pos.z = L->bottom;
		if (s->layer_index > next_index) s->pos.z = L->top;

		if (j->method == XO_METHOD_MBL) {              /* mcml.template.c:584-666 */
			float mua = ml_mua(s, L);
			float deposit_fraction = FP_1 - m_exp(s, -mua*step);
			deposit = deposit_fraction*s->weight;
			s->weight -= deposit;
			s->event_flags |= EV_ABSORPTION;
			if (j->fluence_kind) {
				float step_back = (mua != FP_0) ?
					step - m_div(-m_log(s, FP_1 - sim_random(s)*deposit_fraction), mua) : FP_0;
				p3f dp = { s->pos.x - step_back*s->dir.x, s->pos.y - step_back*s->dir.y,
					s->pos.z - step_back*s->dir.z };
				fluence_deposit_weight(s, &dp, deposit, mua);
			}
		}

		if (next_index != s->layer_index) {
			uint32_t bf = boundary_mcml(s, next_index);
			s->event_flags |= bf | EV_BOUNDARY_HIT;
			if (s->layer_index <= 0 || s->layer_index >= (int32_t)j->num_layers - 1) {
				if (s->layer_index <= 0) {
					if (j->det_kind[LOC_TOP])
						detector_deposit(s, LOC_TOP, &s->pos, &s->dir, s->weight);
				} else if (j->det_kind[LOC_BOTTOM]) {
					detector_deposit(s, LOC_BOTTOM, &s->pos, &s->dir, s->weight);
				}
				done = 1;
			}
			if (j->method == XO_METHOD_MBL) lottery(s, &done);
		} else if (j->method == XO_METHOD_MBL) {
			sim_scatter(s);
			s->event_flags |= EV_SCATTERING;
			lottery(s, &done);
		} else if (j->method == XO_METHOD_AR) {        /* mcml.template.c:705-721 */
			L = ml_layer_at(j, s->layer_index);
			if (sim_random(s) < ml_mua_inv_mut(s, L)) {
				deposit = s->weight;
				done = 1;
				s->weight -= deposit;
				s->event_flags |= EV_ABSORPTION;
				fluence_deposit_weight(s, &s->pos, deposit, ml_mua(s, L));
			} else {
				sim_scatter(s);
				s->event_flags |= EV_SCATTERING;
			}
		} else {                                       /* AW, mcml.template.c:722-754 */
			deposit = s->weight*ml_mua_inv_mut(s, L);
			s->weight -= deposit;
			s->event_flags |= EV_ABSORPTION;
			fluence_deposit_weight(s, &s->pos, deposit, ml_mua(s, L));
			sim_scatter(s);
			s->event_flags |= EV_SCATTERING;
			lottery(s, &done);
		}

		{   /* mcml.template.c:759-763 */
			p3f d = { s->pos.x - src_pos.x, s->pos.y - src_pos.y, s->pos.z - src_pos.z };
			if (dot3(&d, &d) > rmax*rmax) { done = 1; s->event_flags |= EV_ESCAPED; }
		}
		s->event_flags |= done ? EV_TERMINATED : 0;
		if (tr == XO_TRACE_ALL) trace_this_event(s);
		else if ((tr & XO_TRACE_END) && done) trace_this_event(s);
		s->event_flags = 0;

		if (done) {
			if (tr) trace_finalize(s);
			if (next_packet(s, q)) {
				s->trace_count = 0;
				s->opl = FP_0;
				launch_mcml(s);
				s->event_flags |= EV_LAUNCH;
				if (tr & XO_TRACE_START) trace_this_event(s);
				done = 0;
			}
		}
	}
}

#include "xo_oracle_vox.inc"
#include "xo_oracle_cyl.inc"
#include "xo_oracle_sv.inc"

static void run_workitem(xo_oracle_job *job, uint32_t t, quota_t *q, volatile uint32_t *dyn,
		uint32_t *num_kernels, uint64_t *iterations) {
	sim_t s;
	memset(&s, 0, sizeof(s));
	s.job = job;
	s.rng_x = job->rng_x[t];
	s.rng_a = job->rng_a[t];
	s.weight = FP_1;
	s.dyn_counter = dyn;
	uint32_t before = q ? q->next : 0;
	switch (job->geometry) {
		case XO_GEOM_MCML: workitem_mcml(&s, q); break;
		case XO_GEOM_MCVOX: workitem_mcvox(&s, q); break;
		case XO_GEOM_MCCYL: workitem_mccyl(&s, q); break;
	}
	(void)before;
	if (s.iterations > 0) {
		__atomic_fetch_add(num_kernels, 1u, __ATOMIC_RELAXED);
		job->rng_x[t] = s.rng_x;     /* mcml.template.c:821 (only if a packet was taken) */
	}
	__atomic_fetch_add(iterations, s.iterations, __ATOMIC_RELAXED);
}

int xo_oracle_run(xo_oracle_job *job) {
	uint32_t N = job->num_packets, T = job->num_threads;
	if (T == 0) return 1;
	uint32_t qn = N/T, r = N % T, base = 0;
	job->num_kernels = 0; job->num_iterations = 0;
	for (uint32_t t = 0; t < T; ++t) {
		uint32_t n_t = qn + (t < r ? 1u : 0u);
		quota_t q = { base, base + n_t };
		if (n_t) run_workitem(job, t, &q, NULL, &job->num_kernels, &job->num_iterations);
		base += n_t;
	}
	job->num_packets_done = N + T;
	return 0;
}

typedef struct { xo_oracle_job *job; uint32_t t; volatile uint32_t *counter; } dyn_arg;
static void *dyn_worker(void *p) {
	dyn_arg *d = (dyn_arg *)p;
	run_workitem(d->job, d->t, NULL, d->counter, &d->job->num_kernels, &d->job->num_iterations);
	return NULL;
}
int xo_oracle_run_dynamic(xo_oracle_job *job, uint32_t ncpu) {
	if (ncpu == 0) return 1;
	volatile uint32_t counter = 0;
	pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t)*ncpu);
	dyn_arg *da = (dyn_arg *)malloc(sizeof(dyn_arg)*ncpu);
	job->num_kernels = 0; job->num_iterations = 0;
	for (uint32_t t = 0; t < ncpu; ++t) {
		da[t].job = job; da[t].t = t; da[t].counter = &counter;
		pthread_create(&th[t], NULL, dyn_worker, &da[t]);
	}
	for (uint32_t t = 0; t < ncpu; ++t) pthread_join(th[t], NULL);
	job->num_packets_done = counter;
	free(th); free(da);
	return 0;
}

void xo_oracle_math_probe(int32_t fn, int32_t math, uint32_t n,
		const float *in0, const float *in1, float *out0, float *out1) {
	xo_oracle_job j; memset(&j, 0, sizeof(j)); j.math = math;
	sim_t s; memset(&s, 0, sizeof(s)); s.job = &j;
	for (uint32_t i = 0; i < n; ++i) {
		switch (fn) {
			case 0: out0[i] = m_log(&s, in0[i]); break;
			case 1: m_sincos(&s, in0[i], &out0[i], &out1[i]); break;
			case 2: out0[i] = m_cbrt(&s, in0[i]); break;
			case 3: out0[i] = m_pow(&s, in0[i], in1[i]); break;
			case 4: out0[i] = m_exp(&s, in0[i]); break;
			case 5: out0[i] = m_atan2(&s, in0[i], in1[i]); break;
			case 6: out0[i] = m_sqrt(in0[i]); break;
			case 7: out0[i] = m_div(in0[i], in1[i]); break;
		}
	}
}
