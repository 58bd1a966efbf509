// atomic_probe.cu -- measured atomic "roofline" of the accumulator paths:
// how many 64-bit fixed-point deposits per second a B200 sustains through
//   redg64 : RED.E.ADD.64 straight to L2 (what the fluence grids use)
//   atoms  : shared-memory lo/hi 32-bit ATOMS (what the detector windows use)
// for uniform and for source-peaked bin distributions.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/atomic_probe.bin tools/atomic_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned int u32; typedef unsigned long long u64;

__device__ __forceinline__ u32 lcg(u32 &s) { s = s*1664525u + 1013904223u; return s; }
__device__ __forceinline__ u32 pick(u32 &s, u32 nbins, int peaked) {
	u32 r = lcg(s) >> 8;                       // 24 bits
	if (!peaked) return (u32)(((u64)r*nbins) >> 24);
	// peaked: index ~ nbins * u^4 (most deposits in a few % of the bins)
	float u = r*(1.0f/16777216.0f); u *= u; u *= u;
	return (u32)(u*(float)(nbins - 1));
}

template <int FILL>
__global__ void k_redg64(u64 *grid, u32 nbins, int iters, int peaked, float *sink) {
	u32 s = (blockIdx.x*blockDim.x + threadIdx.x)*2654435761u + 12345u;
	float f = s*1e-9f;
	for (int i = 0; i < iters; ++i) {
		u32 idx = pick(s, nbins, peaked);
#pragma unroll
		for (int j = 0; j < FILL; ++j) f = fmaf(f, 1.0000001f, 1e-7f);
		atomicAdd(grid + idx, (u64)(s & 0x7fffff));
	}
	if (f == 123.456f) *sink = f;
}

template <int FILL>
__global__ void k_atoms(u64 *grid, u32 nbins, int iters, int peaked, float *sink) {
	extern __shared__ u32 sh[];                // 2*nbins words (lo, hi)
	for (u32 i = threadIdx.x; i < 2*nbins; i += blockDim.x) sh[i] = 0;
	__syncthreads();
	u32 s = (blockIdx.x*blockDim.x + threadIdx.x)*2654435761u + 12345u;
	float f = s*1e-9f;
	for (int i = 0; i < iters; ++i) {
		u32 idx = pick(s, nbins, peaked);
#pragma unroll
		for (int j = 0; j < FILL; ++j) f = fmaf(f, 1.0000001f, 1e-7f);
		u32 w = s & 0x7fffff;
		u32 old = atomicAdd(sh + 2*idx, w);
		if (old + w < old) atomicAdd(sh + 2*idx + 1, 1u);
	}
	__syncthreads();
	for (u32 i = threadIdx.x; i < nbins; i += blockDim.x) {
		u64 v = ((u64)sh[2*i + 1] << 32) | sh[2*i];
		if (v) atomicAdd(grid + i, v);
	}
	if (f == 123.456f) *sink = f;
}

template <class F> float timeit(F launch) {
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	launch(); cudaDeviceSynchronize();
	cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
	float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	int sms = p.multiProcessorCount;
	printf("{\"device\": \"%s\", \"sms\": %d, \"results\": [\n", p.name, sms);
	u64 *grid; float *sink;
	size_t maxbins = 201u*201u*201u;
	cudaMalloc(&grid, maxbins*8); cudaMemset(grid, 0, maxbins*8); cudaMalloc(&sink, 4);
	const int block = 256, iters = 2000;
	bool first = true;
	for (int ctas_per_sm = 4; ctas_per_sm <= 8; ctas_per_sm += 4) {
		int gridsz = sms*ctas_per_sm;
		double total = (double)gridsz*block*iters;
		u32 sizes[3] = {125000u, 8120601u, 7u};
		for (int si = 0; si < 3; ++si) for (int peaked = 0; peaked < 2; ++peaked) {
			u32 nb = sizes[si];
			float ms0 = timeit([&] { k_redg64<0><<<gridsz, block>>>(grid, nb, iters, peaked, sink); });
			float ms1 = timeit([&] { k_redg64<100><<<gridsz, block>>>(grid, nb, iters, peaked, sink); });
			printf("%s{\"path\": \"redg64\", \"bins\": %u, \"peaked\": %d, \"ctas_per_sm\": %d, \"G_deposits_per_s\": %.2f, \"with_100_fma\": %.2f}",
				first ? "" : ",\n", nb, peaked, ctas_per_sm, total/ms0*1e-6, total/ms1*1e-6);
			first = false;
		}
		u32 ssizes[3] = {1024u, 4096u, 6000u};
		for (int si = 0; si < 3; ++si) for (int peaked = 0; peaked < 2; ++peaked) {
			u32 nb = ssizes[si];
			size_t shb = (size_t)nb*8;
			if (shb*ctas_per_sm > 200*1024) continue;
			cudaFuncSetAttribute(k_atoms<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shb);
			cudaFuncSetAttribute(k_atoms<100>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shb);
			float ms0 = timeit([&] { k_atoms<0><<<gridsz, block, shb>>>(grid, nb, iters, peaked, sink); });
			float ms1 = timeit([&] { k_atoms<100><<<gridsz, block, shb>>>(grid, nb, iters, peaked, sink); });
			printf(",\n{\"path\": \"atoms_lohi\", \"bins\": %u, \"peaked\": %d, \"ctas_per_sm\": %d, \"G_deposits_per_s\": %.2f, \"with_100_fma\": %.2f}",
				nb, peaked, ctas_per_sm, total/ms0*1e-6, total/ms1*1e-6);
		}
	}
	printf("\n]}\n");
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
	return 0;
}
