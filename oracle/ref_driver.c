/* oracle/ref_driver.c -- TEST INFRASTRUCTURE.
 *
 * Drives the reference's McKernel (rendered by the reference's own Python and
 * compiled unchanged behind clshim.h) on CPU cores.  Compiled once per rendered
 * kernel by oracle/refkernel.py with -DXO_REF_GEOMETRY=0|1|2 (mcml|mcvox|mccyl).
 *
 * Two schedules:
 *  - xo_ref_run_static : the deterministic "block" schedule of DESIGN.md
 *    (work-item t simulates packets [base_t, base_t+n_t) with its own MWC
 *    stream) -- a legal outcome of the reference's racing atomic packet counter
 *    (mcml.template.c:460,790).  Implemented without touching the kernel: the
 *    packet counter is reset per work-item and the int/float buffers are shifted
 *    so Trace rows land at row base_t + k.
 *  - xo_ref_run_dynamic: P host threads, work-item p on thread p, shared atomic
 *    counter -- what an OpenCL CPU runtime does; used as the CPU baseline.
 */
#include <stdint.h>
#include <stddef.h>
#include <pthread.h>
#include <stdlib.h>

_Thread_local size_t xo_ref_global_id;

#ifndef XO_REF_GEOMETRY
#define XO_REF_GEOMETRY 0
#endif

/* floating-point type of the rendered kernel: -DXO_REF_DOUBLE for McDataTypesDouble */
#ifdef XO_REF_DOUBLE
typedef double xo_ref_fp;
#else
typedef float xo_ref_fp;
#endif

typedef struct {
	uint32_t num_packets;
	uint32_t *num_packets_done;
	uint32_t *num_kernels;
	double rmax;                /* converted to the kernel's precision at the call */
	uint64_t *rng_x;
	const uint32_t *rng_a;
	/* geometry: mcml/mccyl: g0=num_layers(u32 by value), g1=layers
	 *           mcvox: g1=voxel_cfg, g2=voxel data, g0=num_materials, g3=materials */
	uint32_t g0;
	const void *g1, *g2, *g3;
	const void *source, *surface, *trace, *fluence, *detectors;
	const xo_ref_fp *fp_lut;
	int32_t *int_buffer;
	xo_ref_fp *float_buffer;
	uint64_t *accumulator_buffer;
	/* trace row strides (0 when no trace): floats per packet, ints per packet */
	uint64_t trace_float_stride, trace_int_stride;
} xo_ref_args;

#if XO_REF_GEOMETRY == 1
void McKernel(uint32_t, uint32_t *, uint32_t *, xo_ref_fp, uint64_t *, const uint32_t *,
	const void *, const void *, uint32_t, const void *,
	const void *, const void *, const void *, const void *, const void *, const xo_ref_fp *,
	int32_t *, xo_ref_fp *, uint64_t *);
static void call_kernel(const xo_ref_args *a, uint32_t n, uint32_t *done,
		int32_t *ibuf, xo_ref_fp *fbuf) {
	McKernel(n, done, a->num_kernels, (xo_ref_fp)a->rmax, a->rng_x, a->rng_a,
		a->g1, a->g2, a->g0, a->g3,
		a->source, a->surface, a->trace, a->fluence, a->detectors, a->fp_lut,
		ibuf, fbuf, a->accumulator_buffer);
}
#else
void McKernel(uint32_t, uint32_t *, uint32_t *, xo_ref_fp, uint64_t *, const uint32_t *,
	uint32_t, const void *,
	const void *, const void *, const void *, const void *, const void *, const xo_ref_fp *,
	int32_t *, xo_ref_fp *, uint64_t *);
static void call_kernel(const xo_ref_args *a, uint32_t n, uint32_t *done,
		int32_t *ibuf, xo_ref_fp *fbuf) {
	McKernel(n, done, a->num_kernels, (xo_ref_fp)a->rmax, a->rng_x, a->rng_a,
		a->g0, a->g1,
		a->source, a->surface, a->trace, a->fluence, a->detectors, a->fp_lut,
		ibuf, fbuf, a->accumulator_buffer);
}
#endif

void xo_ref_run_static(const xo_ref_args *a, uint32_t nthreads) {
	uint32_t N = a->num_packets, q = N / nthreads, r = N % nthreads;
	uint64_t base = 0;
	for (uint32_t t = 0; t < nthreads; ++t) {
		uint32_t n_t = q + (t < r ? 1u : 0u);
		uint32_t done = 0;
		xo_ref_global_id = t;
		if (n_t)
			call_kernel(a, n_t, &done,
				a->int_buffer + base*a->trace_int_stride,
				a->float_buffer + base*a->trace_float_stride);
		base += n_t;
	}
	*a->num_packets_done = N + nthreads;
}

typedef struct { const xo_ref_args *a; uint32_t gid; } worker_arg;
static void *worker(void *p) {
	worker_arg *w = (worker_arg *)p;
	xo_ref_global_id = w->gid;
	call_kernel(w->a, w->a->num_packets, w->a->num_packets_done,
		w->a->int_buffer, w->a->float_buffer);
	return NULL;
}

void xo_ref_run_dynamic(const xo_ref_args *a, uint32_t nthreads) {
	pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t)*nthreads);
	worker_arg *wa = (worker_arg *)malloc(sizeof(worker_arg)*nthreads);
	for (uint32_t t = 0; t < nthreads; ++t) {
		wa[t].a = a; wa[t].gid = t;
		pthread_create(&th[t], NULL, worker, &wa[t]);
	}
	for (uint32_t t = 0; t < nthreads; ++t)
		pthread_join(th[t], NULL);
	free(th); free(wa);
}

/* P host threads executing `nitems` >= P work-items (thread p runs work-items
 * p, p+P, ...): what an OpenCL CPU runtime does with a global size larger than
 * the core count.  Needed for mccyl, whose reference kernel gives every
 * work-item a budget of 1e6 loop trips (mccyl.template.c:681). */
typedef struct { const xo_ref_args *a; uint32_t first, stride, nitems; } items_arg;
static void *items_worker(void *p) {
	items_arg *w = (items_arg *)p;
	for (uint32_t gid = w->first; gid < w->nitems; gid += w->stride) {
		if (*(volatile uint32_t *)w->a->num_packets_done >= w->a->num_packets) break;
		xo_ref_global_id = gid;
		call_kernel(w->a, w->a->num_packets, w->a->num_packets_done,
			w->a->int_buffer, w->a->float_buffer);
	}
	return NULL;
}

void xo_ref_run_dynamic_items(const xo_ref_args *a, uint32_t nthreads, uint32_t nitems) {
	pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t)*nthreads);
	items_arg *wa = (items_arg *)malloc(sizeof(items_arg)*nthreads);
	for (uint32_t t = 0; t < nthreads; ++t) {
		wa[t].a = a; wa[t].first = t; wa[t].stride = nthreads; wa[t].nitems = nitems;
		pthread_create(&th[t], NULL, items_worker, &wa[t]);
	}
	for (uint32_t t = 0; t < nthreads; ++t)
		pthread_join(th[t], NULL);
	free(th); free(wa);
}

#ifdef XO_REF_HAS_SV
/* the reference's SamplingVolume kernel (mcsv.template.c:236), one work-item:
 * no random numbers, integer accumulation -> schedule independent */
void SamplingVolume(uint32_t, uint32_t *, uint32_t *, const void *, const void *,
	uint64_t *, int32_t *, xo_ref_fp *, uint64_t *);
void xo_ref_run_sv(uint32_t npackets, const void *trace, const void *sv,
		uint64_t *total_weight, int32_t *ibuf, xo_ref_fp *fbuf, uint64_t *abuf) {
	uint32_t processed = 0, kernels = 0;
	xo_ref_global_id = 0;
	SamplingVolume(npackets, &processed, &kernels, trace, sv, total_weight, ibuf, fbuf, abuf);
}
#endif
