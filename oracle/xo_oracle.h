/* oracle/xo_oracle.h -- TEST INFRASTRUCTURE: job descriptor of the CPU oracle.
 *
 * The oracle is a plain-C restatement of the reference's photon-packet kernels
 * (xopto/mc{ml,vox,cyl}/kernel/*.template.c + the built-in plugin fragments).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it; the product path (pyxopto_b200) never does.
 *
 * All plugin parameter blocks are the reference's *packed ctypes structs* (same
 * bytes the OpenCL kernel receives); plugin kinds that the reference selects at
 * compile time through #defines are selected here at run time.
 */
#ifndef XO_ORACLE_H
#define XO_ORACLE_H
#include <stdint.h>

enum { XO_GEOM_MCML = 0, XO_GEOM_MCVOX = 1, XO_GEOM_MCCYL = 2 };
enum { XO_METHOD_AW = 0, XO_METHOD_AR = 1, XO_METHOD_MBL = 2 };
enum { XO_MATH_LIBM = 0, XO_MATH_PORTABLE = 1 };
enum { XO_PF_HG = 1, XO_PF_MHG = 2, XO_PF_GK = 3, XO_PF_LUT = 4, XO_PF_HG2 = 5,
	XO_PF_GK2 = 6, XO_PF_MGK = 7, XO_PF_PC = 8, XO_PF_MPC = 9, XO_PF_HGDIR = 10,
	XO_PF_RAYLEIGH = 11 };
enum {
	XO_SRC_LINE = 1, XO_SRC_GAUSSIANBEAM = 2, XO_SRC_UNIFORMFIBER = 3,
	XO_SRC_ISOTROPICPOINT = 4, XO_SRC_UNIFORMBEAM = 5, XO_SRC_LAMBERTIANFIBER = 6,
	XO_SRC_ISOTROPICVOXEL = 7, XO_SRC_UNIFORMFIBERLUT = 8,
	XO_SRC_UNIFORMRECTANGULAR = 9, XO_SRC_LAMBERTIANRECTANGULAR = 10,
	XO_SRC_ISOTROPICVOXELS = 11,
	XO_SRC_UNIFORMFIBERNI = 12, XO_SRC_LAMBERTIANFIBERNI = 13, XO_SRC_UNIFORMFIBERLUTNI = 14,
	XO_SRC_UNIFORMRECTANGULARLUT = 15
};
enum {
	XO_DET_NONE = 0, XO_DET_TOTAL = 1, XO_DET_RADIAL = 2, XO_DET_CARTESIAN = 3,
	XO_DET_SIXAROUNDONE = 4, XO_DET_RADIALPL = 5, XO_DET_TOTALPL = 6,
	XO_DET_SYMMETRICX = 7, XO_DET_FIZ = 8, XO_DET_CARTESIANPL = 9,
	XO_DET_SIXAROUNDONEPL = 10, XO_DET_TOTAL_CYL = 11, XO_DET_LINEARARRAY = 12,
	XO_DET_FIBERARRAY = 13, XO_DET_LINEARARRAYPL = 14, XO_DET_FIBERARRAYPL = 15,
	XO_DET_TOTALLUT = 16, XO_DET_TOTALLUTPL = 17,
	XO_DET_FIBERLUTARRAY = 18
};
enum {
	XO_FLU_NONE = 0, XO_FLU_XYZ = 1, XO_FLU_RZ = 2, XO_FLU_XYZT = 3,
	XO_FLU_RZT = 4, XO_FLU_CYL = 5, XO_FLU_CYLT = 6
};
enum { XO_SURF_NONE = 0, XO_SURF_LAMBERTIAN = 1, XO_SURF_SIXAROUNDONE = 2,
	XO_SURF_LINEARARRAY = 3, XO_SURF_FIBERARRAY = 4 };
enum { XO_TRACE_NONE = 0, XO_TRACE_START = 1, XO_TRACE_END = 2, XO_TRACE_ALL = 7 };

typedef struct xo_oracle_job {
	/* compile-time options of the reference, chosen at run time here */
	int32_t geometry;
	int32_t method;
	int32_t math;
	int32_t use_lottery;
	float weight_min;
	float lottery_chance;
	int32_t pf_kind;
	int32_t pf_size;           /* sizeof(McPf) in the packed layer/material */
	int32_t src_kind;
	int32_t det_kind[3];       /* top, bottom, specular */
	int32_t det_offset[3];     /* byte offsets inside the packed McDetectors */
	int32_t det_param[3];      /* compile-time parameter of the plugin text (fiber count) */
	int32_t fluence_kind;
	int32_t fluence_rate;      /* MC_FLUENCE_MODE_RATE */
	int32_t trace_flags;       /* MC_USE_TRACE value */
	int32_t use_events;        /* MC_USE_EVENTS (trace event mask active) */
	int32_t track_opl;         /* MC_TRACK_OPTICAL_PATHLENGTH */
	int32_t surf_kind[2];      /* top, bottom surface layout (mcml) */
	int32_t surf_offset[2];    /* byte offsets inside the packed McSurfaceLayouts */
	int32_t surf_param[2];     /* compile-time parameter of the layout text (fiber count) */
	int32_t enhanced_rng;      /* MC_USE_ENHANCED_RNG: two MWC steps per draw (mcbase.template.c:1577-1586) */
	int32_t anisotropic;       /* layers / materials carry mus, mua, mut tensors (AnisotropicLayer / AnisotropicMaterial) */

	/* run-time kernel arguments (mcml.template.c:346-376, mcvox.template.c:548) */
	uint32_t num_packets;
	uint32_t num_threads;      /* work-items of the static block schedule */
	float rmax;
	uint32_t num_layers;       /* layers (mcml/mccyl) or materials (mcvox) */
	const void *layers;        /* packed McLayer[] or McMaterial[] */
	const void *voxel_cfg;     /* packed McVoxelConfig (mcvox) */
	const int32_t *voxels;     /* McVoxel[nz][ny][nx] (mcvox) */
	const void *source;
	const void *detectors;
	const void *fluence;
	const void *trace;
	const void *surface;       /* packed McSurfaceLayouts (mcml; may be NULL) */
	const float *fp_lut;
	uint64_t *rng_x;           /* in/out, one per work-item */
	const uint32_t *rng_a;
	int32_t *int_buffer;
	float *float_buffer;
	uint64_t *accumulator_buffer;
	/* outputs */
	uint32_t num_kernels;
	uint32_t num_packets_done;
	uint64_t num_iterations;   /* total loop iterations (roofline unit count) */
} xo_oracle_job;

/* static block schedule: work-item t simulates packets [base_t, base_t + n_t) */
int xo_oracle_run(xo_oracle_job *job);
/* dynamic schedule on `ncpu` host threads (CPU baseline "port") */
int xo_oracle_run_dynamic(xo_oracle_job *job, uint32_t ncpu);
/* n draws of fp_random_single from (x, a) -- mcbase.template.c:1640-1646 */
void xo_oracle_rng_test(uint64_t x, uint32_t a, uint32_t n, float *out);
/* seed derivation restated from xopto/src/rng/rng.cpp:64-103 */
int xo_oracle_init_rng(uint64_t *x, uint32_t *a, const uint32_t *fora,
	uint32_t n_rng, uint64_t xinit);
/* SamplingVolume kernel restated (mcsv.template.c:236-420): packed McTrace and
 * McSamplingVolume, the int / float / accumulator flat buffers; returns the
 * number of loop trips */
uint64_t xo_oracle_sampling_volume(uint32_t npackets, const void *trace, const void *sv,
	uint64_t *total_weight, const int32_t *int_buffer, const float *fp_buffer,
	uint64_t *accu_buffer);
/* elementary function probes for the GPU math parity test */
void xo_oracle_math_probe(int32_t fn, int32_t math, uint32_t n,
	const float *in0, const float *in1, float *out0, float *out1);
#endif
