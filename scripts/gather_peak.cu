/*
 * gather_peak.cu — what this GPU sustains for the access patterns that bound the hash kernels of the path, measured the way
 * MEASURED_PEAKS.json measures the streaming copy: a kernel that does nothing else.
 *
 *   gather   every thread loads 16 bytes (header + first key, as the join probe does) from a random 32-byte entry of a table
 *            much larger than L2; UNROLL independent loads in flight per thread            -> G sectors/s
 *   cas      every thread claims a random 32-byte entry with atomicCAS on its first word and writes the other 24 bytes
 *            (hash build: nodeHash.c ExecHashTableInsert's device counterpart)              -> G inserts/s
 *   atomic   every thread atomicAdds two 64-bit words of a random 32-byte entry (general HashAggregate transition)
 *                                                                                           -> G rows/s
 *
 * The roofline of a kernel that streams S bytes and makes R such random accesses is  S / copy_bw + R / rate  when the two do
 * not overlap and max(...) when they overlap perfectly; bench.py reports both bounds for the join and hash-aggregate kernels.
 *
 * build:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/gather_peak scripts/gather_peak.cu
 * run:    build/gather_peak [table MiB = 2048] [accesses = 2e8]        -> one JSON line
 */
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t h)
{
	h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
	return h;
}

template <int UNROLL>
__global__ void __launch_bounds__(256) gather_kernel(const ulonglong2 *tab, uint64_t mask, uint64_t n, unsigned long long *sink)
{
	const uint64_t tid = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x, nt = gridDim.x * (uint64_t) blockDim.x;
	unsigned long long acc = 0;
	for (uint64_t i = tid; i < n; i += nt * UNROLL)
	{
		ulonglong2 v[UNROLL];
#pragma unroll
		for (int u = 0; u < UNROLL; u++)
			v[u] = __ldg(tab + 2 * (mix(i + u * nt) & mask));
#pragma unroll
		for (int u = 0; u < UNROLL; u++)
			acc += v[u].x ^ v[u].y;
	}
	if (acc == 0x1234567890abcdefull) *sink = acc;
}

__global__ void __launch_bounds__(256) cas_kernel(unsigned long long *tab, uint64_t mask, uint64_t n, unsigned long long *sink)
{
	const uint64_t tid = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x, nt = gridDim.x * (uint64_t) blockDim.x;
	unsigned long long fails = 0;
	for (uint64_t i = tid; i < n; i += nt)
	{
		uint64_t slot = mix(i) & mask;
		for (int tries = 0; tries < 64; tries++)
		{
			unsigned long long *e = tab + 4 * slot;
			if (atomicCAS(e, 0ull, 0x8000000000000000ull | i) == 0ull) { e[1] = i; e[2] = i + 1; e[3] = i + 2; break; }
			slot = (slot + 1) & mask;
			fails++;
		}
	}
	if (fails == 0xffffffffffffffffull) *sink = fails;
}

__global__ void __launch_bounds__(256) atomic_kernel(unsigned long long *tab, uint64_t mask, uint64_t n)
{
	const uint64_t tid = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x, nt = gridDim.x * (uint64_t) blockDim.x;
	for (uint64_t i = tid; i < n; i += nt)
	{
		unsigned long long *e = tab + 4 * (mix(i) & mask);
		atomicAdd(e + 2, 1ull);
		atomicAdd((double *) (e + 3), 1.5);
	}
}

int main(int argc, char **argv)
{
	const size_t mib = argc > 1 ? (size_t) atoll(argv[1]) : 2048;
	const uint64_t n = argc > 2 ? (uint64_t) atof(argv[2]) : 200000000ull;
	uint64_t entries = 1;
	while (entries * 2 * 32 <= mib << 20) entries *= 2;
	const uint64_t mask = entries - 1;
	unsigned long long *tab, *sink;
	CK(cudaMalloc(&tab, entries * 32));
	CK(cudaMalloc(&sink, 8));
	CK(cudaMemset(tab, 0x5a, entries * 32));
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, 0));
	const int grid = prop.multiProcessorCount * 8;
	cudaEvent_t a, b;
	CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	float ms, best[5] = { 1e30f, 1e30f, 1e30f, 1e30f, 1e30f };
	for (int rep = 0; rep < 4; rep++)
	{
		CK(cudaEventRecord(a)); gather_kernel<1><<<grid, 256>>>((const ulonglong2 *) tab, mask, n, sink); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
		CK(cudaEventElapsedTime(&ms, a, b)); if (rep && ms < best[0]) best[0] = ms;
		CK(cudaEventRecord(a)); gather_kernel<4><<<grid, 256>>>((const ulonglong2 *) tab, mask, n, sink); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
		CK(cudaEventElapsedTime(&ms, a, b)); if (rep && ms < best[1]) best[1] = ms;
		CK(cudaEventRecord(a)); gather_kernel<8><<<grid, 256>>>((const ulonglong2 *) tab, mask, n, sink); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
		CK(cudaEventElapsedTime(&ms, a, b)); if (rep && ms < best[2]) best[2] = ms;
	}
	const uint64_t nins = entries / 2 < n ? entries / 2 : n;          /* load factor 0.5, like the join table */
	for (int rep = 0; rep < 3; rep++)
	{
		CK(cudaMemset(tab, 0, entries * 32));
		CK(cudaEventRecord(a)); cas_kernel<<<grid, 256>>>(tab, mask, nins, sink); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
		CK(cudaEventElapsedTime(&ms, a, b)); if (rep && ms < best[3]) best[3] = ms;
	}
	for (int rep = 0; rep < 3; rep++)
	{
		CK(cudaEventRecord(a)); atomic_kernel<<<grid, 256>>>(tab, mask, n); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
		CK(cudaEventElapsedTime(&ms, a, b)); if (rep && ms < best[4]) best[4] = ms;
	}
	CK(cudaGetLastError());
	printf("{\"table_bytes\": %llu, \"accesses\": %llu, \"gather_g_per_s\": {\"ilp1\": %.3f, \"ilp4\": %.3f, \"ilp8\": %.3f}, "
	       "\"cas_insert_g_per_s\": %.3f, \"cas_inserts\": %llu, \"atomic_pair_g_per_s\": %.3f, \"sms\": %d}\n",
	       (unsigned long long) (entries * 32), (unsigned long long) n, n / best[0] / 1e6, n / best[1] / 1e6, n / best[2] / 1e6,
	       nins / best[3] / 1e6, (unsigned long long) nins, n / best[4] / 1e6, prop.multiProcessorCount);
	return 0;
}
