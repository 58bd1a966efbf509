/* include/xopto_b200.h -- C ABI of libxopto_b200.so
 *
 * Drop-in boundary of the B200-native photon-packet engine.  The reference
 * (xopto/pyxopto) reaches its device exclusively through pyopencl from
 * xopto/mcbase/mcworker.py (ClWorker), xopto/mc{ml,vox,cyl}/mc.py (Mc.run) and
 * xopto/cl/clrng.py (seed derivation, optionally via the C library built from
 * xopto/src/rng/rng.cpp).  Every entry point below replaces one of those
 * pyopencl (or rng.cpp) interfaces; the reference file:line each one stands in
 * for is cited at its declaration.  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions: plain C, no CUDA/torch types; handles are opaque 64-bit values;
 * host pointers are borrowed only for the duration of the call (NumPy buffers);
 * every function returns 0 on success or a negative xo_status, never throws;
 * xo_last_error() returns the message of the last failure on the calling thread
 * (the Python layer turns it into RuntimeError, mirroring pyopencl.RuntimeError).
 * The library has no link-time dependency on the CUDA driver or NVRTC: both are
 * loaded lazily, so it loads (and exports every symbol) on a machine without a
 * GPU, where device calls fail loudly with XO_ERR_NO_DRIVER -- there is no CPU
 * fallback of any kind.
 */
#ifndef XOPTO_B200_H
#define XOPTO_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t xo_handle;

typedef enum xo_status {
	XO_OK = 0,
	XO_ERR_NO_DRIVER = -1,     /* libcuda.so.1 / libnvrtc missing or no device */
	XO_ERR_CUDA = -2,          /* a CUDA driver call failed */
	XO_ERR_COMPILE = -3,       /* NVRTC compilation failed (log returned) */
	XO_ERR_INVALID = -4,       /* bad handle / argument */
	XO_ERR_NOT_FOUND = -5      /* kernel name not in module */
} xo_status;

/* pyopencl.Device attributes read by xopto/cl/clinfo.py:283-357 (device_info) */
typedef struct xo_device_info {
	char name[256];
	int32_t cc_major, cc_minor;
	int32_t multiprocessor_count;
	int32_t max_threads_per_block;
	int32_t max_threads_per_multiprocessor;
	int32_t max_shared_per_block_optin;
	int32_t regs_per_multiprocessor;
	int32_t clock_rate_khz;
	int32_t l2_cache_bytes;
	int32_t reserved;
	uint64_t total_global_mem;
} xo_device_info;

/* one kernel argument: replaces the NumPy-scalar / cl.Buffer positional
 * arguments of `McKernel(queue, global, local, *args)` (mcml/mc.py:922-953) */
typedef struct xo_arg {
	int32_t kind;              /* 0: by-value bytes (scalar or packed struct); 1: buffer */
	int32_t size;              /* by-value: number of bytes */
	const void *value;         /* by-value: host pointer to the bytes */
	xo_handle buffer;          /* buffer: handle from xo_buffer_alloc */
	uint64_t offset;           /* buffer: byte offset added to the device pointer */
} xo_arg;

/* message of the last failure on this thread (empty string if none) */
const char *xo_last_error(void);

/* pyopencl.get_platforms()/get_devices(): xopto/cl/clinfo.py:39-281 */
int xo_device_count(int32_t *count);
int xo_device_get_info(int32_t ordinal, xo_device_info *info);

/* cl.Context(devices): xopto/mcbase/mcworker.py:168-199 */
int xo_ctx_create(int32_t ordinal, xo_handle *ctx);
int xo_ctx_destroy(xo_handle ctx);

/* cl.CommandQueue(ctx, properties=PROFILING_ENABLE): mcworker.py:183-199;
 * a second queue on the same context is used by mcprogress.py:72 */
int xo_stream_create(xo_handle ctx, xo_handle *stream);
int xo_stream_destroy(xo_handle stream);
int xo_stream_sync(xo_handle stream);
/* raw CUstream value, so torch.cuda.ExternalStream / NCCL can share it */
int xo_stream_native(xo_handle stream, uint64_t *custream);

/* cl.Program(ctx, src).build(options): mcworker.py:243-286 (cl_build).
 * `src` is CUDA C++ text, `headers`/`header_names` are in-memory include files,
 * `options` NVRTC flags (the architecture flag is appended by the library from
 * the context's device, sm_100a on B200).  `log`/`log_cap` receive the NVRTC
 * log (PYOPENCL_COMPILER_OUTPUT equivalent). */
int xo_module_build(xo_handle ctx, const char *src, const char *name,
	const char *const *options, int32_t n_options,
	const char *const *headers, const char *const *header_names, int32_t n_headers,
	xo_handle *module, char *log, size_t log_cap);
/* compile only (no device needed): returns the cubin for `arch`
 * (e.g. "sm_100a") in a library-owned blob; used by the CPU build check and to
 * pre-populate the in-tree kernel cache */
int xo_compile(const char *src, const char *name, const char *arch,
	const char *const *options, int32_t n_options,
	const char *const *headers, const char *const *header_names, int32_t n_headers,
	xo_handle *blob, char *log, size_t log_cap);
int xo_blob_size(xo_handle blob, size_t *size);
int xo_blob_copy(xo_handle blob, void *dst, size_t cap);
int xo_blob_free(xo_handle blob);
/* load a cubin produced by xo_compile (kernel cache hit) */
int xo_module_load(xo_handle ctx, const void *image, size_t size, xo_handle *module);
int xo_module_unload(xo_handle module);
/* attribute access by kernel name on cl.Program: mc.py:922 `_cl_exec.McKernel` */
int xo_module_get_kernel(xo_handle module, const char *name, xo_handle *kernel);
/* registers, static shared memory, max threads per block, local (spill) bytes */
int xo_kernel_get_attributes(xo_handle kernel, int32_t *num_regs,
	int32_t *static_shared, int32_t *max_threads, int32_t *local_bytes);
/* active blocks per SM for a block size / dynamic shared memory */
int xo_kernel_occupancy(xo_handle kernel, int32_t block, size_t dynamic_shared,
	int32_t *blocks_per_sm);

/* cl.Buffer(ctx, flags, size=|hostbuf=): mcworker.py:288-393, 395-430 */
int xo_buffer_alloc(xo_handle ctx, size_t size, xo_handle *buffer);
int xo_buffer_free(xo_handle buffer);
int xo_buffer_size(xo_handle buffer, size_t *size);
int xo_buffer_device_ptr(xo_handle buffer, uint64_t *dptr);

/* cl.enqueue_copy(queue, dst, src, device_offset=): mcworker.py:790-899;
 * blocking when `blocking` != 0 (the reference always .wait()s) */
int xo_copy_h2d(xo_handle stream, xo_handle buffer, size_t offset,
	const void *host, size_t size, int32_t blocking);
int xo_copy_d2h(xo_handle stream, void *host, xo_handle buffer, size_t offset,
	size_t size, int32_t blocking);
/* fill_<type> kernels / cl.enqueue_fill_buffer: mcworker.py:432-496,
 * mcbase.template.c:1655-1748.  elem_size in {1,2,4,8}. */
int xo_fill(xo_handle stream, xo_handle buffer, size_t offset, size_t count,
	int32_t elem_size, const void *pattern);

/* page-locked host staging memory (bench e2e path; no reference equivalent) */
int xo_host_alloc(xo_handle ctx, size_t size, void **ptr);
int xo_host_free(xo_handle ctx, void *ptr);

/* kernel call `k(queue, global_size, local_size, *args)`: mc.py:904-953.
 * grid/block are CUDA launch dimensions (1-D), dynamic_shared in bytes. */
int xo_launch(xo_handle stream, xo_handle kernel, uint32_t grid, uint32_t block,
	uint32_t dynamic_shared, const xo_arg *args, int32_t n_args);

/* pyopencl event .wait()/.profile: mcworker.py:212-241 (event_timing) */
int xo_event_create(xo_handle ctx, xo_handle *event);
int xo_event_destroy(xo_handle event);
int xo_event_record(xo_handle event, xo_handle stream);
int xo_event_sync(xo_handle event);
int xo_event_elapsed_ms(xo_handle start, xo_handle stop, float *ms);

/* seed derivation: same signature and semantics as the reference's native
 * library, xopto/src/rng/rng.cpp:64-103 (bound by xopto/cl/clrng.py:333-369,
 * 513-525).  Returns 1 for an invalid xinit, 0 on success.  Host only. */
int init_RNG(uint64_t *x, uint32_t *a, uint32_t *fora, const uint32_t n_rng,
	uint64_t xinit);

/* library version (major*10000 + minor*100 + patch) */
int xo_version(void);

#ifdef __cplusplus
}
#endif
#endif /* XOPTO_B200_H */
