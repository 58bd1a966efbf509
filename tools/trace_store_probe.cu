// trace_store_probe.cu -- which store path carries the Trace event stream (config 4)?
// Every thread owns one packet row (row stride = maxlen x 32 B = 16 KB at maxlen 512)
// and emits one 32-byte event per loop trip, with FILL dependent FMAs of "physics"
// between two events.  Paths:
//   direct128 : two STG.128 per event and lane (the un-staged kernel)
//   direct256 : one 256-bit store per event and lane (st.global.v8.f32, sm_100+)
//   coop128   : events staged in shared memory, whole 128-byte lines written by 8 lanes
//               each (the round-1 staged kernel; lanes in phase here: its best case)
//   bulkN     : N events staged per lane in a lane-private, bank-skewed slab of shared
//               memory; the lane itself emits the slab with ONE cp.async.bulk
//               (shared::cta -> global, UBLKCP), N = 4 / 8 / 16 (128 / 256 / 512 B)
// Output: JSON lines with the achieved GB/s of the event stream.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/trace_store_probe.bin tools/trace_store_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned int u32; typedef unsigned long long u64;

#define ROW_F4 1024            // float4 per row: 512 events x 2

__device__ __forceinline__ float work(float f, int fill) {
	for (int j = 0; j < fill; ++j) f = fmaf(f, 1.0000001f, 1e-7f);
	return f;
}

__global__ void k_direct128(float4 *rows, int events, int fill) {
	const u64 t = blockIdx.x*(u64)blockDim.x + threadIdx.x;
	float4 *row = rows + t*ROW_F4;
	float f = t*1e-9f;
	for (int e = 0; e < events; ++e) {
		f = work(f, fill);
		row[2*e] = make_float4(f, f + 1.0f, f + 2.0f, f + 3.0f);
		row[2*e + 1] = make_float4(f + 4.0f, f + 5.0f, f + 6.0f, f + 7.0f);
	}
}

__global__ void k_direct256(float4 *rows, int events, int fill) {
	const u64 t = blockIdx.x*(u64)blockDim.x + threadIdx.x;
	float4 *row = rows + t*ROW_F4;
	float f = t*1e-9f;
	for (int e = 0; e < events; ++e) {
		f = work(f, fill);
		asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
			:: "l"(row + 2*e), "f"(f), "f"(f + 1.0f), "f"(f + 2.0f), "f"(f + 3.0f),
			   "f"(f + 4.0f), "f"(f + 5.0f), "f"(f + 6.0f), "f"(f + 7.0f) : "memory");
	}
}

// 8 float4 columns x (32 lanes + 1 pad) per warp; a line = 4 events of one lane
__global__ void k_coop128(float4 *rows, int events, int fill) {
	extern __shared__ float4 sh4[];
	float4 *stage = sh4 + (threadIdx.x >> 5)*264;
	const u32 lane = threadIdx.x & 31u;
	const u64 t = blockIdx.x*(u64)blockDim.x + threadIdx.x;
	const u64 warp_row0 = t - lane;
	float f = t*1e-9f;
	for (int e = 0; e < events; ++e) {
		f = work(f, fill);
		const u32 col = (e & 3)*2;
		stage[col*33 + lane] = make_float4(f, f + 1.0f, f + 2.0f, f + 3.0f);
		stage[(col + 1)*33 + lane] = make_float4(f + 4.0f, f + 5.0f, f + 6.0f, f + 7.0f);
		if ((e & 3) == 3) {
			__syncwarp();
			const u32 k = lane & 7u;
			for (u32 src = lane >> 3; src < 32u; src += 4u) {
				float4 *line = rows + (warp_row0 + src)*ROW_F4 + (u64)(e >> 2)*8u;
				line[k] = stage[k*33 + src];
			}
			__syncwarp();
		}
	}
}

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

// lane-private slab of N events (+16 B of skew so that the 8 lanes of a quarter warp
// hit 8 different 16-byte bank groups), emitted by the lane with one bulk copy
template <int N>
__global__ void k_bulk(float4 *rows, int events, int fill) {
	extern __shared__ float4 sh4[];
	constexpr int SLAB = 2*N + 1;                       // float4 per lane
	float4 *slab = sh4 + threadIdx.x*SLAB;
	const u64 t = blockIdx.x*(u64)blockDim.x + threadIdx.x;
	float4 *row = rows + t*ROW_F4;
	float f = t*1e-9f;
	for (int e = 0; e < events; ++e) {
		f = work(f, fill);
		const int s = e % N;
		if (s == 0)     // the slab is free again once the previous copy has read it
			asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
		slab[2*s] = make_float4(f, f + 1.0f, f + 2.0f, f + 3.0f);
		slab[2*s + 1] = make_float4(f + 4.0f, f + 5.0f, f + 6.0f, f + 7.0f);
		if (s == N - 1) {
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
				:: "l"(row + (u64)(e/N)*(2*N)), "r"(smem_u32(slab)), "n"(32*N) : "memory");
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
		}
	}
	asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class F> float timeit(F launch) {
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	launch(); cudaDeviceSynchronize();
	cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
	float ms = 0; cudaEventElapsedTime(&ms, a, b);
	cudaError_t err = cudaGetLastError();
	if (err != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(err)); return -1.0f; }
	return ms;
}

int main() {
	cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
	const int sms = prop.multiProcessorCount;
	const int events = 256;
	float4 *rows = nullptr;
	printf("{\"device\": \"%s\", \"sms\": %d, \"row_bytes\": %d, \"events_per_row\": %d, \"results\": [\n",
		prop.name, sms, ROW_F4*16, events);
	bool first = true;
	const int blocks_per_sm_list[] = {2, 3, 4};
	const int fills[] = {0, 100, 200};
	for (int bi = 0; bi < 3; ++bi) {
		const int block = 256, bps = blocks_per_sm_list[bi];
		const int grid = sms*bps;
		const size_t bytes = (size_t)grid*block*ROW_F4*16;
		if (cudaMalloc(&rows, bytes) != cudaSuccess) { fprintf(stderr, "alloc failed\n"); return 1; }
		cudaMemset(rows, 0, bytes);
		const double gb = (double)grid*block*events*32.0/1e9;
		for (int fi = 0; fi < 3; ++fi) {
			const int fill = fills[fi];
			struct { const char *name; float ms; } res[6];
			int n = 0;
			res[n++] = { "direct128", timeit([&] { k_direct128<<<grid, block>>>(rows, events, fill); }) };
			res[n++] = { "direct256", timeit([&] { k_direct256<<<grid, block>>>(rows, events, fill); }) };
			{
				size_t sh = (block/32)*264*16;
				cudaFuncSetAttribute(k_coop128, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
				res[n++] = { "coop128", timeit([&] { k_coop128<<<grid, block, sh>>>(rows, events, fill); }) };
			}
#define BULK(N) { \
				size_t sh = (size_t)block*(2*N + 1)*16; \
				if (sh*bps <= 220*1024) { \
					cudaFuncSetAttribute(k_bulk<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh); \
					res[n++] = { "bulk" #N, timeit([&] { k_bulk<N><<<grid, block, sh>>>(rows, events, fill); }) }; \
				} }
			BULK(4) BULK(8) BULK(16)
			for (int i = 0; i < n; ++i) {
				printf("%s{\"path\": \"%s\", \"ctas_per_sm\": %d, \"block\": %d, \"fill_fma\": %d, \"ms\": %.4f, \"GBps\": %.1f}",
					first ? "" : ",\n", res[i].name, bps, block, fill, res[i].ms, gb/(res[i].ms*1e-3));
				first = false;
			}
		}
		cudaFree(rows);
	}
	printf("\n]}\n");
	return 0;
}
