// Micro-benchmark (B200): instruction-fetch cost of a single warp running straight-line code of growing footprint.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache icache.cu && ./icache
#include <cstdio>
#include <cuda_runtime.h>
template <int N>
__device__ __forceinline__ float body(float x, float y) {
#pragma unroll
	for (int i = 0; i < N; i++) x = fmaf(x, y, (float)(i & 7));   // N dependent FFMA (4 cycles each when fetch keeps up)
	return x;
}
template <int N>
__global__ void k(float *out, long long *cyc, float y, int iters) {
	float x = threadIdx.x;
	x = body<N>(x, y);  // warm
	long long t0 = clock64();
	for (int it = 0; it < iters; it++) x = body<N>(x, y);
	long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = x;
	if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int N>
void run(float *out, long long *cyc, int blocks, int threads) {
	const int iters = 64;
	k<N><<<blocks, threads>>>(out, cyc, 0.999f, iters);
	cudaDeviceSynchronize();
	k<N><<<blocks, threads>>>(out, cyc, 0.999f, iters);
	cudaDeviceSynchronize();
	long long h[1024];
	cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
	double s = 0; for (int i = 0; i < blocks; i++) s += h[i];
	printf("  code %7.1f KB : %.2f cycles/instr (blocks=%d threads=%d)\n", N * 16 / 1024.0, s / blocks / iters / N, blocks, threads);
}
int main() {
	float *out; long long *cyc;
	cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8192);
	for (int cfg = 0; cfg < 3; cfg++) {
		const int blocks = cfg == 0 ? 1 : 296, threads = cfg == 2 ? 192 : 32;
		run<256>(out, cyc, blocks, threads); run<1024>(out, cyc, blocks, threads); run<2048>(out, cyc, blocks, threads); run<4096>(out, cyc, blocks, threads);
		run<8192>(out, cyc, blocks, threads); run<16384>(out, cyc, blocks, threads); run<32768>(out, cyc, blocks, threads);
	}
	return 0;
}
