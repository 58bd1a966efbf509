// libxopto_b200.so -- C ABI over the CUDA driver API + NVRTC (include/xopto_b200.h).
//
// Replaces the pyopencl surface used by the reference's ClWorker / Mc.run
// (xopto/mcbase/mcworker.py, xopto/mcml/mc.py:797-1018) and the native seed
// library xopto/src/rng/rng.cpp.  libcuda.so.1 and libnvrtc are dlopen'ed on
// first use, so the library loads without a GPU; device calls then fail with
// XO_ERR_NO_DRIVER (there is deliberately no CPU fallback).
#include "../../include/xopto_b200.h"

#include <cuda.h>
#include <nvrtc.h>
#include <dlfcn.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace {

thread_local std::string g_error;

int fail(int code, const char *fmt, ...) {
	char buf[2048];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_error = buf;
	return code;
}

// ---- lazily bound driver / NVRTC entry points ------------------------------
#define XO_CU_FUNCS(X) \
	X(cuInit) X(cuDeviceGetCount) X(cuDeviceGet) X(cuDeviceGetName) \
	X(cuDeviceGetAttribute) X(cuDeviceTotalMem_v2) X(cuDevicePrimaryCtxRetain) \
	X(cuDevicePrimaryCtxRelease_v2) X(cuCtxPushCurrent_v2) X(cuCtxPopCurrent_v2) \
	X(cuStreamCreate) X(cuStreamDestroy_v2) X(cuStreamSynchronize) \
	X(cuModuleLoadData) X(cuModuleUnload) X(cuModuleGetFunction) \
	X(cuFuncGetAttribute) X(cuFuncSetAttribute) \
	X(cuOccupancyMaxActiveBlocksPerMultiprocessor) \
	X(cuMemAlloc_v2) X(cuMemFree_v2) X(cuMemcpyHtoDAsync_v2) X(cuMemcpyDtoHAsync_v2) \
	X(cuMemsetD8Async) X(cuMemsetD16Async) X(cuMemsetD32Async) X(cuMemsetD2D32Async) \
	X(cuMemAllocHost_v2) X(cuMemFreeHost) X(cuLaunchKernel) \
	X(cuEventCreate) X(cuEventDestroy_v2) X(cuEventRecord) X(cuEventSynchronize) \
	X(cuEventElapsedTime) X(cuGetErrorString) X(cuGetErrorName)

struct CudaApi {
	void *handle = nullptr;
	bool ok = false;
#define X(name) decltype(&::name) name = nullptr;
	XO_CU_FUNCS(X)
#undef X
};

// cuda.h maps the unsuffixed names to _v2 with macros; take addresses through
// the real exported symbol names.
#define XO_STR2(x) #x
#define XO_STR(x) XO_STR2(x)

CudaApi g_cu;
std::once_flag g_cu_once;
std::string g_cu_error;

void load_cuda() {
	const char *names[] = {"libcuda.so.1", "libcuda.so"};
	for (const char *n : names) {
		g_cu.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
		if (g_cu.handle) break;
	}
	if (!g_cu.handle) {
		g_cu_error = "CUDA driver library libcuda.so.1 not found (no GPU driver)";
		return;
	}
#define X(name) \
	g_cu.name = reinterpret_cast<decltype(g_cu.name)>(dlsym(g_cu.handle, XO_STR(name))); \
	if (!g_cu.name) { g_cu_error = std::string("missing driver symbol ") + XO_STR(name); return; }
	XO_CU_FUNCS(X)
#undef X
	CUresult r = g_cu.cuInit(0);
	if (r != CUDA_SUCCESS) {
		g_cu_error = "cuInit failed with code " + std::to_string((int)r) + " (no usable GPU)";
		return;
	}
	g_cu.ok = true;
}

int need_cuda() {
	std::call_once(g_cu_once, load_cuda);
	if (!g_cu.ok) return fail(XO_ERR_NO_DRIVER, "%s", g_cu_error.c_str());
	return XO_OK;
}

int cu_check(CUresult r, const char *what) {
	if (r == CUDA_SUCCESS) return XO_OK;
	const char *name = nullptr, *msg = nullptr;
	g_cu.cuGetErrorName(r, &name);
	g_cu.cuGetErrorString(r, &msg);
	return fail(XO_ERR_CUDA, "%s failed: %s (%s)", what, name ? name : "?", msg ? msg : "?");
}
#define CU(call) do { int rc_ = cu_check(g_cu.call, #call); if (rc_) return rc_; } while (0)

#define XO_NVRTC_FUNCS(X) \
	X(nvrtcCreateProgram) X(nvrtcDestroyProgram) X(nvrtcCompileProgram) \
	X(nvrtcGetProgramLogSize) X(nvrtcGetProgramLog) X(nvrtcGetCUBINSize) \
	X(nvrtcGetCUBIN) X(nvrtcGetErrorString) X(nvrtcVersion)

struct NvrtcApi {
	void *handle = nullptr;
	bool ok = false;
#define X(name) decltype(&::name) name = nullptr;
	XO_NVRTC_FUNCS(X)
#undef X
};
NvrtcApi g_rtc;
std::once_flag g_rtc_once;
std::string g_rtc_error;

void load_nvrtc() {
	// The newest NVRTC in reach wins: a host process may already have mapped an older
	// libnvrtc.so.12 under the same soname (PyTorch bundles the one of its own CUDA
	// build), and the kernels use PTX of the toolkit this library was built against
	// (256-bit global stores need PTX ISA 8.8).  Absolute paths are separate objects
	// for the dynamic loader, so every candidate is opened and asked for its version.
	std::vector<std::string> names;
	if (const char *env = getenv("XOPTO_NVRTC")) names.push_back(env);
	for (const char *var : {"CUDA_HOME", "CUDA_PATH"})
		if (const char *home = getenv(var)) names.push_back(std::string(home) + "/lib64/libnvrtc.so.12");
	names.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
	names.push_back("/usr/local/cuda/targets/x86_64-linux/lib/libnvrtc.so.12");
	names.push_back("/usr/local/cuda/lib64/libnvrtc.so");
	names.push_back("libnvrtc.so.12");
	names.push_back("libnvrtc.so");
	int best = -1;
	for (size_t i = 0; i < names.size(); ++i) {
		void *h = dlopen(names[i].c_str(), RTLD_NOW | RTLD_LOCAL);
		if (!h) continue;
		int major = 0, minor = 0;
		auto version = reinterpret_cast<decltype(&::nvrtcVersion)>(dlsym(h, "nvrtcVersion"));
		const int v = (version && version(&major, &minor) == NVRTC_SUCCESS) ? major*1000 + minor : 0;
		if (i == 0 && getenv("XOPTO_NVRTC")) { g_rtc.handle = h; break; }   // explicit choice
		if (v > best) {
			if (g_rtc.handle) dlclose(g_rtc.handle);
			g_rtc.handle = h;
			best = v;
		} else {
			dlclose(h);
		}
	}
	if (!g_rtc.handle) {
		g_rtc_error = "NVRTC library libnvrtc.so.12 not found";
		return;
	}
#define X(name) \
	g_rtc.name = reinterpret_cast<decltype(g_rtc.name)>(dlsym(g_rtc.handle, #name)); \
	if (!g_rtc.name) { g_rtc_error = std::string("missing NVRTC symbol ") + #name; return; }
	XO_NVRTC_FUNCS(X)
#undef X
	g_rtc.ok = true;
}

int need_nvrtc() {
	std::call_once(g_rtc_once, load_nvrtc);
	if (!g_rtc.ok) return fail(XO_ERR_NO_DRIVER, "%s", g_rtc_error.c_str());
	return XO_OK;
}

// ---- objects ---------------------------------------------------------------
enum class Kind : uint32_t { Ctx = 0x58430001, Stream, Module, Kernel, Buffer, Event, Blob };

struct Object { Kind kind; };
// A context stays alive until the last object created from it is destroyed, so
// host-language finalisers may run in any order (Python GC does not order them).
struct Ctx : Object { CUdevice dev; CUcontext ctx; int ordinal; int cc_major, cc_minor;
	std::atomic<int> refs{1}; };
struct Stream : Object { Ctx *ctx; CUstream s; };
struct Kernel;
struct Module : Object { Ctx *ctx; CUmodule m; std::vector<Kernel *> kernels; };
struct Kernel : Object { Module *mod; CUfunction f; };
struct Buffer : Object { Ctx *ctx; CUdeviceptr p; size_t size; };
struct Event : Object { Ctx *ctx; CUevent e; };
struct Blob : Object { std::vector<char> data; };

template <class T> T *as(xo_handle h, Kind k) {
	Object *o = reinterpret_cast<Object *>(static_cast<uintptr_t>(h));
	if (!o || o->kind != k) return nullptr;
	return static_cast<T *>(o);
}
template <class T> xo_handle to_handle(T *p) { return static_cast<xo_handle>(reinterpret_cast<uintptr_t>(p)); }

struct CtxScope {
	bool pushed = false;
	explicit CtxScope(Ctx *c) { if (c && g_cu.cuCtxPushCurrent_v2(c->ctx) == CUDA_SUCCESS) pushed = true; }
	~CtxScope() { if (pushed) { CUcontext old; g_cu.cuCtxPopCurrent_v2(&old); } }
};

void ctx_retain(Ctx *c) { c->refs.fetch_add(1); }
void ctx_release(Ctx *c) {
	if (c->refs.fetch_sub(1) == 1) {
		g_cu.cuDevicePrimaryCtxRelease_v2(c->dev);
		c->kind = Kind(0);
		delete c;
	}
}

int compile_to_cubin(const char *src, const char *name, const char *arch,
		const char *const *options, int32_t n_options,
		const char *const *headers, const char *const *header_names, int32_t n_headers,
		std::vector<char> &cubin, char *log, size_t log_cap) {
	int rc = need_nvrtc();
	if (rc) return rc;
	if (log && log_cap) log[0] = 0;
	nvrtcProgram prog;
	nvrtcResult r = g_rtc.nvrtcCreateProgram(&prog, src, name ? name : "xo_kernel.cu",
		n_headers, headers, header_names);
	if (r != NVRTC_SUCCESS)
		return fail(XO_ERR_COMPILE, "nvrtcCreateProgram: %s", g_rtc.nvrtcGetErrorString(r));
	std::vector<const char *> opts;
	std::string archopt = std::string("--gpu-architecture=") + arch;
	opts.push_back(archopt.c_str());
	for (int i = 0; i < n_options; ++i) opts.push_back(options[i]);
	r = g_rtc.nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
	size_t log_size = 0;
	g_rtc.nvrtcGetProgramLogSize(prog, &log_size);
	std::string full_log(log_size, '\0');
	if (log_size > 1) g_rtc.nvrtcGetProgramLog(prog, &full_log[0]);
	if (log && log_cap) {
		size_t n = full_log.size() < log_cap - 1 ? full_log.size() : log_cap - 1;
		memcpy(log, full_log.data(), n);
		log[n] = 0;
	}
	if (r != NVRTC_SUCCESS) {
		g_rtc.nvrtcDestroyProgram(&prog);
		std::string tail = full_log.size() > 1500 ? full_log.substr(0, 1500) : full_log;
		return fail(XO_ERR_COMPILE, "NVRTC compilation failed (%s):\n%s",
			g_rtc.nvrtcGetErrorString(r), tail.c_str());
	}
	size_t size = 0;
	r = g_rtc.nvrtcGetCUBINSize(prog, &size);
	if (r != NVRTC_SUCCESS || size == 0) {
		g_rtc.nvrtcDestroyProgram(&prog);
		return fail(XO_ERR_COMPILE, "nvrtcGetCUBINSize: %s", g_rtc.nvrtcGetErrorString(r));
	}
	cubin.resize(size);
	g_rtc.nvrtcGetCUBIN(prog, cubin.data());
	g_rtc.nvrtcDestroyProgram(&prog);
	return XO_OK;
}

}  // namespace

extern "C" {

const char *xo_last_error(void) { return g_error.c_str(); }

int xo_version(void) { return 100; }

int xo_device_count(int32_t *count) {
	if (!count) return fail(XO_ERR_INVALID, "count is NULL");
	*count = 0;
	int rc = need_cuda();
	if (rc) return rc;
	int n = 0;
	CU(cuDeviceGetCount(&n));
	*count = n;
	return XO_OK;
}

int xo_device_get_info(int32_t ordinal, xo_device_info *info) {
	if (!info) return fail(XO_ERR_INVALID, "info is NULL");
	int rc = need_cuda();
	if (rc) return rc;
	CUdevice dev;
	CU(cuDeviceGet(&dev, ordinal));
	memset(info, 0, sizeof(*info));
	CU(cuDeviceGetName(info->name, sizeof(info->name), dev));
	auto attr = [&](CUdevice_attribute a, int32_t *out) {
		int v = 0;
		g_cu.cuDeviceGetAttribute(&v, a, dev);
		*out = v;
	};
	attr(CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR, &info->cc_major);
	attr(CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR, &info->cc_minor);
	attr(CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, &info->multiprocessor_count);
	attr(CU_DEVICE_ATTRIBUTE_MAX_THREADS_PER_BLOCK, &info->max_threads_per_block);
	attr(CU_DEVICE_ATTRIBUTE_MAX_THREADS_PER_MULTIPROCESSOR, &info->max_threads_per_multiprocessor);
	attr(CU_DEVICE_ATTRIBUTE_MAX_SHARED_MEMORY_PER_BLOCK_OPTIN, &info->max_shared_per_block_optin);
	attr(CU_DEVICE_ATTRIBUTE_MAX_REGISTERS_PER_MULTIPROCESSOR, &info->regs_per_multiprocessor);
	attr(CU_DEVICE_ATTRIBUTE_CLOCK_RATE, &info->clock_rate_khz);
	attr(CU_DEVICE_ATTRIBUTE_L2_CACHE_SIZE, &info->l2_cache_bytes);
	size_t total = 0;
	CU(cuDeviceTotalMem_v2(&total, dev));
	info->total_global_mem = total;
	return XO_OK;
}

int xo_ctx_create(int32_t ordinal, xo_handle *ctx) {
	if (!ctx) return fail(XO_ERR_INVALID, "ctx is NULL");
	*ctx = 0;
	int rc = need_cuda();
	if (rc) return rc;
	Ctx *c = new Ctx();
	c->kind = Kind::Ctx;
	c->ordinal = ordinal;
	rc = cu_check(g_cu.cuDeviceGet(&c->dev, ordinal), "cuDeviceGet");
	if (!rc) rc = cu_check(g_cu.cuDevicePrimaryCtxRetain(&c->ctx, c->dev), "cuDevicePrimaryCtxRetain");
	if (rc) { delete c; return rc; }
	g_cu.cuDeviceGetAttribute(&c->cc_major, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR, c->dev);
	g_cu.cuDeviceGetAttribute(&c->cc_minor, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR, c->dev);
	*ctx = to_handle(c);
	return XO_OK;
}

int xo_ctx_destroy(xo_handle h) {
	Ctx *c = as<Ctx>(h, Kind::Ctx);
	if (!c) return fail(XO_ERR_INVALID, "invalid context handle");
	ctx_release(c);
	return XO_OK;
}

int xo_stream_create(xo_handle hctx, xo_handle *stream) {
	Ctx *c = as<Ctx>(hctx, Kind::Ctx);
	if (!c || !stream) return fail(XO_ERR_INVALID, "invalid context handle");
	CtxScope scope(c);
	CUstream s;
	CU(cuStreamCreate(&s, CU_STREAM_NON_BLOCKING));
	Stream *st = new Stream();
	st->kind = Kind::Stream; st->ctx = c; st->s = s;
	ctx_retain(c);
	*stream = to_handle(st);
	return XO_OK;
}

int xo_stream_destroy(xo_handle h) {
	Stream *s = as<Stream>(h, Kind::Stream);
	if (!s) return fail(XO_ERR_INVALID, "invalid stream handle");
	CtxScope scope(s->ctx);
	g_cu.cuStreamDestroy_v2(s->s);
	s->kind = Kind(0);
	Ctx *owner = s->ctx;
	delete s;
	{ CUcontext old; if (scope.pushed) { g_cu.cuCtxPopCurrent_v2(&old); scope.pushed = false; } }
	ctx_release(owner);
	return XO_OK;
}

int xo_stream_sync(xo_handle h) {
	Stream *s = as<Stream>(h, Kind::Stream);
	if (!s) return fail(XO_ERR_INVALID, "invalid stream handle");
	CtxScope scope(s->ctx);
	CU(cuStreamSynchronize(s->s));
	return XO_OK;
}

int xo_stream_native(xo_handle h, uint64_t *custream) {
	Stream *s = as<Stream>(h, Kind::Stream);
	if (!s || !custream) return fail(XO_ERR_INVALID, "invalid stream handle");
	*custream = (uint64_t)(uintptr_t)s->s;
	return XO_OK;
}

int xo_compile(const char *src, const char *name, const char *arch,
		const char *const *options, int32_t n_options,
		const char *const *headers, const char *const *header_names, int32_t n_headers,
		xo_handle *blob, char *log, size_t log_cap) {
	if (!src || !arch || !blob) return fail(XO_ERR_INVALID, "NULL argument");
	*blob = 0;
	Blob *b = new Blob();
	b->kind = Kind::Blob;
	int rc = compile_to_cubin(src, name, arch, options, n_options, headers, header_names,
		n_headers, b->data, log, log_cap);
	if (rc) { delete b; return rc; }
	*blob = to_handle(b);
	return XO_OK;
}

int xo_blob_size(xo_handle h, size_t *size) {
	Blob *b = as<Blob>(h, Kind::Blob);
	if (!b || !size) return fail(XO_ERR_INVALID, "invalid blob handle");
	*size = b->data.size();
	return XO_OK;
}

int xo_blob_copy(xo_handle h, void *dst, size_t cap) {
	Blob *b = as<Blob>(h, Kind::Blob);
	if (!b || !dst || cap < b->data.size()) return fail(XO_ERR_INVALID, "invalid blob copy");
	memcpy(dst, b->data.data(), b->data.size());
	return XO_OK;
}

int xo_blob_free(xo_handle h) {
	Blob *b = as<Blob>(h, Kind::Blob);
	if (!b) return fail(XO_ERR_INVALID, "invalid blob handle");
	b->kind = Kind(0);
	delete b;
	return XO_OK;
}

int xo_module_load(xo_handle hctx, const void *image, size_t size, xo_handle *module) {
	Ctx *c = as<Ctx>(hctx, Kind::Ctx);
	if (!c || !image || !module) return fail(XO_ERR_INVALID, "invalid argument");
	(void)size;
	CtxScope scope(c);
	CUmodule m;
	CU(cuModuleLoadData(&m, image));
	Module *mod = new Module();
	mod->kind = Kind::Module; mod->ctx = c; mod->m = m;
	ctx_retain(c);
	*module = to_handle(mod);
	return XO_OK;
}

int xo_module_build(xo_handle hctx, const char *src, const char *name,
		const char *const *options, int32_t n_options,
		const char *const *headers, const char *const *header_names, int32_t n_headers,
		xo_handle *module, char *log, size_t log_cap) {
	Ctx *c = as<Ctx>(hctx, Kind::Ctx);
	if (!c || !src || !module) return fail(XO_ERR_INVALID, "invalid argument");
	char arch[32];
	// architecture-specific target ("a" suffix) for Hopper/Blackwell class parts
	snprintf(arch, sizeof(arch), "sm_%d%d%s", c->cc_major, c->cc_minor,
		c->cc_major >= 9 ? "a" : "");
	std::vector<char> cubin;
	int rc = compile_to_cubin(src, name, arch, options, n_options, headers, header_names,
		n_headers, cubin, log, log_cap);
	if (rc) return rc;
	return xo_module_load(hctx, cubin.data(), cubin.size(), module);
}

int xo_module_unload(xo_handle h) {
	Module *m = as<Module>(h, Kind::Module);
	if (!m) return fail(XO_ERR_INVALID, "invalid module handle");
	CtxScope scope(m->ctx);
	g_cu.cuModuleUnload(m->m);
	for (Kernel *k : m->kernels) { k->kind = Kind(0); delete k; }
	m->kind = Kind(0);
	Ctx *owner = m->ctx;
	delete m;
	{ CUcontext old; if (scope.pushed) { g_cu.cuCtxPopCurrent_v2(&old); scope.pushed = false; } }
	ctx_release(owner);
	return XO_OK;
}

int xo_module_get_kernel(xo_handle h, const char *name, xo_handle *kernel) {
	Module *m = as<Module>(h, Kind::Module);
	if (!m || !name || !kernel) return fail(XO_ERR_INVALID, "invalid module handle");
	CtxScope scope(m->ctx);
	CUfunction f;
	CUresult r = g_cu.cuModuleGetFunction(&f, m->m, name);
	if (r == CUDA_ERROR_NOT_FOUND) return fail(XO_ERR_NOT_FOUND, "kernel '%s' not found", name);
	int rc = cu_check(r, "cuModuleGetFunction");
	if (rc) return rc;
	Kernel *k = new Kernel();
	k->kind = Kind::Kernel; k->mod = m; k->f = f;
	m->kernels.push_back(k);
	*kernel = to_handle(k);
	return XO_OK;
}

int xo_kernel_get_attributes(xo_handle h, int32_t *num_regs, int32_t *static_shared,
		int32_t *max_threads, int32_t *local_bytes) {
	Kernel *k = as<Kernel>(h, Kind::Kernel);
	if (!k) return fail(XO_ERR_INVALID, "invalid kernel handle");
	CtxScope scope(k->mod->ctx);
	int v;
	if (num_regs) { CU(cuFuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_NUM_REGS, k->f)); *num_regs = v; }
	if (static_shared) { CU(cuFuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, k->f)); *static_shared = v; }
	if (max_threads) { CU(cuFuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK, k->f)); *max_threads = v; }
	if (local_bytes) { CU(cuFuncGetAttribute(&v, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, k->f)); *local_bytes = v; }
	return XO_OK;
}

int xo_kernel_occupancy(xo_handle h, int32_t block, size_t dynamic_shared, int32_t *blocks_per_sm) {
	Kernel *k = as<Kernel>(h, Kind::Kernel);
	if (!k || !blocks_per_sm) return fail(XO_ERR_INVALID, "invalid kernel handle");
	CtxScope scope(k->mod->ctx);
	if (dynamic_shared > 48*1024)
		CU(cuFuncSetAttribute(k->f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)dynamic_shared));
	int n = 0;
	CU(cuOccupancyMaxActiveBlocksPerMultiprocessor(&n, k->f, block, dynamic_shared));
	*blocks_per_sm = n;
	return XO_OK;
}

int xo_buffer_alloc(xo_handle hctx, size_t size, xo_handle *buffer) {
	Ctx *c = as<Ctx>(hctx, Kind::Ctx);
	if (!c || !buffer) return fail(XO_ERR_INVALID, "invalid context handle");
	CtxScope scope(c);
	CUdeviceptr p = 0;
	CU(cuMemAlloc_v2(&p, size ? size : 1));
	Buffer *b = new Buffer();
	b->kind = Kind::Buffer; b->ctx = c; b->p = p; b->size = size;
	ctx_retain(c);
	*buffer = to_handle(b);
	return XO_OK;
}

int xo_buffer_free(xo_handle h) {
	Buffer *b = as<Buffer>(h, Kind::Buffer);
	if (!b) return fail(XO_ERR_INVALID, "invalid buffer handle");
	CtxScope scope(b->ctx);
	g_cu.cuMemFree_v2(b->p);
	b->kind = Kind(0);
	Ctx *owner = b->ctx;
	delete b;
	{ CUcontext old; if (scope.pushed) { g_cu.cuCtxPopCurrent_v2(&old); scope.pushed = false; } }
	ctx_release(owner);
	return XO_OK;
}

int xo_buffer_size(xo_handle h, size_t *size) {
	Buffer *b = as<Buffer>(h, Kind::Buffer);
	if (!b || !size) return fail(XO_ERR_INVALID, "invalid buffer handle");
	*size = b->size;
	return XO_OK;
}

int xo_buffer_device_ptr(xo_handle h, uint64_t *dptr) {
	Buffer *b = as<Buffer>(h, Kind::Buffer);
	if (!b || !dptr) return fail(XO_ERR_INVALID, "invalid buffer handle");
	*dptr = (uint64_t)b->p;
	return XO_OK;
}

int xo_copy_h2d(xo_handle hs, xo_handle hb, size_t offset, const void *host, size_t size,
		int32_t blocking) {
	Stream *s = as<Stream>(hs, Kind::Stream);
	Buffer *b = as<Buffer>(hb, Kind::Buffer);
	if (!s || !b || (!host && size)) return fail(XO_ERR_INVALID, "invalid handle in xo_copy_h2d");
	if (offset + size > b->size) return fail(XO_ERR_INVALID, "h2d copy out of range (%zu+%zu > %zu)", offset, size, b->size);
	if (!size) return XO_OK;
	CtxScope scope(s->ctx);
	CU(cuMemcpyHtoDAsync_v2(b->p + offset, host, size, s->s));
	if (blocking) CU(cuStreamSynchronize(s->s));
	return XO_OK;
}

int xo_copy_d2h(xo_handle hs, void *host, xo_handle hb, size_t offset, size_t size,
		int32_t blocking) {
	Stream *s = as<Stream>(hs, Kind::Stream);
	Buffer *b = as<Buffer>(hb, Kind::Buffer);
	if (!s || !b || (!host && size)) return fail(XO_ERR_INVALID, "invalid handle in xo_copy_d2h");
	if (offset + size > b->size) return fail(XO_ERR_INVALID, "d2h copy out of range (%zu+%zu > %zu)", offset, size, b->size);
	if (!size) return XO_OK;
	CtxScope scope(s->ctx);
	CU(cuMemcpyDtoHAsync_v2(host, b->p + offset, size, s->s));
	if (blocking) CU(cuStreamSynchronize(s->s));
	return XO_OK;
}

int xo_fill(xo_handle hs, xo_handle hb, size_t offset, size_t count, int32_t elem_size,
		const void *pattern) {
	Stream *s = as<Stream>(hs, Kind::Stream);
	Buffer *b = as<Buffer>(hb, Kind::Buffer);
	if (!s || !b || !pattern) return fail(XO_ERR_INVALID, "invalid handle in xo_fill");
	if (offset + count*(size_t)elem_size > b->size) return fail(XO_ERR_INVALID, "fill out of range");
	if (!count) return XO_OK;
	CtxScope scope(s->ctx);
	CUdeviceptr p = b->p + offset;
	switch (elem_size) {
	case 1: CU(cuMemsetD8Async(p, *(const uint8_t *)pattern, count, s->s)); break;
	case 2: CU(cuMemsetD16Async(p, *(const uint16_t *)pattern, count, s->s)); break;
	case 4: CU(cuMemsetD32Async(p, *(const uint32_t *)pattern, count, s->s)); break;
	case 8: {
		uint32_t lo = ((const uint32_t *)pattern)[0], hi = ((const uint32_t *)pattern)[1];
		if (lo == hi) {
			CU(cuMemsetD32Async(p, lo, count*2, s->s));
		} else {
			// two strided 2-D memsets: column 0 = low words, column 1 = high words
			CU(cuMemsetD2D32Async(p, 8, lo, 1, count, s->s));
			CU(cuMemsetD2D32Async(p + 4, 8, hi, 1, count, s->s));
		}
		break;
	}
	default: return fail(XO_ERR_INVALID, "unsupported fill element size %d", elem_size);
	}
	return XO_OK;
}

int xo_host_alloc(xo_handle hctx, size_t size, void **ptr) {
	Ctx *c = as<Ctx>(hctx, Kind::Ctx);
	if (!c || !ptr) return fail(XO_ERR_INVALID, "invalid context handle");
	CtxScope scope(c);
	CU(cuMemAllocHost_v2(ptr, size ? size : 1));
	return XO_OK;
}

int xo_host_free(xo_handle hctx, void *ptr) {
	Ctx *c = as<Ctx>(hctx, Kind::Ctx);
	if (!c) return fail(XO_ERR_INVALID, "invalid context handle");
	CtxScope scope(c);
	CU(cuMemFreeHost(ptr));
	return XO_OK;
}

int xo_launch(xo_handle hs, xo_handle hk, uint32_t grid, uint32_t block,
		uint32_t dynamic_shared, const xo_arg *args, int32_t n_args) {
	Stream *s = as<Stream>(hs, Kind::Stream);
	Kernel *k = as<Kernel>(hk, Kind::Kernel);
	if (!s || !k || (n_args && !args)) return fail(XO_ERR_INVALID, "invalid handle in xo_launch");
	if (grid == 0 || block == 0) return fail(XO_ERR_INVALID, "empty launch");
	CtxScope scope(s->ctx);
	std::vector<CUdeviceptr> ptrs(n_args);
	std::vector<void *> params(n_args);
	for (int i = 0; i < n_args; ++i) {
		if (args[i].kind == 1) {
			Buffer *b = as<Buffer>(args[i].buffer, Kind::Buffer);
			if (!b) return fail(XO_ERR_INVALID, "argument %d is not a valid buffer", i);
			ptrs[i] = b->p + args[i].offset;
			params[i] = &ptrs[i];
		} else {
			params[i] = const_cast<void *>(args[i].value);
		}
	}
	if (dynamic_shared > 48*1024)
		CU(cuFuncSetAttribute(k->f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)dynamic_shared));
	CU(cuLaunchKernel(k->f, grid, 1, 1, block, 1, 1, dynamic_shared, s->s, params.data(), nullptr));
	return XO_OK;
}

int xo_event_create(xo_handle hctx, xo_handle *event) {
	Ctx *c = as<Ctx>(hctx, Kind::Ctx);
	if (!c || !event) return fail(XO_ERR_INVALID, "invalid context handle");
	CtxScope scope(c);
	CUevent e;
	CU(cuEventCreate(&e, CU_EVENT_DEFAULT));
	Event *ev = new Event();
	ev->kind = Kind::Event; ev->ctx = c; ev->e = e;
	ctx_retain(c);
	*event = to_handle(ev);
	return XO_OK;
}

int xo_event_destroy(xo_handle h) {
	Event *e = as<Event>(h, Kind::Event);
	if (!e) return fail(XO_ERR_INVALID, "invalid event handle");
	CtxScope scope(e->ctx);
	g_cu.cuEventDestroy_v2(e->e);
	e->kind = Kind(0);
	Ctx *owner = e->ctx;
	delete e;
	{ CUcontext old; if (scope.pushed) { g_cu.cuCtxPopCurrent_v2(&old); scope.pushed = false; } }
	ctx_release(owner);
	return XO_OK;
}

int xo_event_record(xo_handle he, xo_handle hs) {
	Event *e = as<Event>(he, Kind::Event);
	Stream *s = as<Stream>(hs, Kind::Stream);
	if (!e || !s) return fail(XO_ERR_INVALID, "invalid handle in xo_event_record");
	CtxScope scope(s->ctx);
	CU(cuEventRecord(e->e, s->s));
	return XO_OK;
}

int xo_event_sync(xo_handle h) {
	Event *e = as<Event>(h, Kind::Event);
	if (!e) return fail(XO_ERR_INVALID, "invalid event handle");
	CtxScope scope(e->ctx);
	CU(cuEventSynchronize(e->e));
	return XO_OK;
}

int xo_event_elapsed_ms(xo_handle hstart, xo_handle hstop, float *ms) {
	Event *a = as<Event>(hstart, Kind::Event);
	Event *b = as<Event>(hstop, Kind::Event);
	if (!a || !b || !ms) return fail(XO_ERR_INVALID, "invalid event handle");
	CtxScope scope(a->ctx);
	CU(cuEventElapsedTime(ms, a->e, b->e));
	return XO_OK;
}

// Seed derivation for the per-work-item MWC generators.  Restated from the
// published algorithm of the reference's native library (rng.cpp:64-103): the
// first multiplier drives a 64-bit MWC stream that produces, for every
// generator i, a carry c in [0, a_i) and a start value x; degenerate states
// are rejected and redrawn.
int init_RNG(uint64_t *x, uint32_t *a, uint32_t *fora, const uint32_t n_rng, uint64_t xinit) {
	const uint64_t mult = fora[0];
	const uint32_t seed_hi = (uint32_t)(xinit >> 32), seed_lo = (uint32_t)xinit;
	if (xinit == 0 || seed_hi >= fora[0] - 1 || seed_lo == 0xffffffffu)
		return 1;
	uint64_t state = xinit;
	auto advance = [&]() -> uint32_t {
		state = (state & 0xffffffffull)*mult + (state >> 32);
		return (uint32_t)state;
	};
	for (uint32_t i = 0; i < n_rng; ++i) {
		const uint32_t ai = fora[i + 1];
		a[i] = ai;
		uint64_t xi;
		do {
			const double u = (double)advance()/4294967296.0;
			const uint32_t carry = (uint32_t)std::floor(u*(double)ai);
			xi = ((uint64_t)carry << 32) + advance();
		} while (xi == 0 || (uint32_t)(xi >> 32) >= ai - 1 || (uint32_t)xi == 0xffffffffu);
		x[i] = xi;
	}
	return 0;
}

}  // extern "C"
