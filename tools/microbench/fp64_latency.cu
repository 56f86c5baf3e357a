// Micro-benchmark (B200): dependent-issue latencies that bound the critical path of the register-resident solver.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double seed, int mode) {
	__shared__ double sm[256];
	double x = seed + threadIdx.x * 1e-9, y = seed * 0.5;
	sm[threadIdx.x] = x;
	__syncthreads();
	long long t0 = clock64();
	const int N = 1024;
	if (mode == 0) { for (int i = 0; i < N; i++) x = fma(x, y, 1e-3); }                       // dependent DFMA
	else if (mode == 1) { for (int i = 0; i < N; i++) x = rsqrt(x) + 1.0; }                    // rsqrt + DADD
	else if (mode == 2) { for (int i = 0; i < N; i++) x = 1.0 / x + 1.0; }                     // division
	else if (mode == 3) { for (int i = 0; i < N; i++) x = sqrt(x) + 1.0; }                     // sqrt
	else if (mode == 4) { int j = threadIdx.x; for (int i = 0; i < N; i++) { j = (int)sm[j & 255] & 255; sm[(j + 7) & 255] = j; } x = j; }  // LDS->STS chain
	else if (mode == 5) { for (int i = 0; i < N; i++) { __syncthreads(); } }                   // barrier
	else if (mode == 6) { for (int i = 0; i < N; i++) { double s, c; sincos(x, &s, &c); x = s + c; } }
	else if (mode == 7) { for (int i = 0; i < N; i++) { x = fmod(x + 7.0, 6.283185307179586) ; } }
	else if (mode == 8) { for (int i = 0; i < N; i++) { sm[threadIdx.x] = x; __syncthreads(); x = sm[(threadIdx.x + 1) & 63] + 1.0; __syncthreads(); } }  // publish/consume
	else if (mode == 9) { float f = (float)x; for (int i = 0; i < N; i++) f = fmaf(f, 0.5f, 1e-3f); x = f; }  // dependent FFMA
	else if (mode == 10) { double a = x, b = x + 1, c = x + 2, d = x + 3; for (int i = 0; i < N; i++) { a = fma(a, y, 1e-3); b = fma(b, y, 1e-3); c = fma(c, y, 1e-3); d = fma(d, y, 1e-3); } x = a + b + c + d; }  // 4 independent chains
	long long t1 = clock64();
	out[blockIdx.x * blockDim.x + threadIdx.x] = x;
	if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0);
}
int main() {
	double *out; long long *cyc;
	cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096 * 8);
	const char *names[] = {"dependent DFMA", "rsqrt+DADD", "div+DADD", "sqrt+DADD", "LDS->STS chain", "__syncthreads", "sincos+DADD", "fmod(2pi)+DADD", "STS,bar,LDS,bar", "dependent FFMA", "4 indep DFMA chains (per 4)"};
	for (int threads : {32, 192}) for (int blocks : {1, 296}) {
		printf("threads=%d blocks=%d\n", threads, blocks);
		for (int mode = 0; mode <= 10; mode++) {
			k<<<blocks, threads>>>(out, cyc, 1.2345, mode);
			cudaDeviceSynchronize();
			k<<<blocks, threads>>>(out, cyc, 1.2345, mode);
			cudaDeviceSynchronize();
			long long h[4096];
			cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
			double s = 0; for (int i = 0; i < blocks; i++) s += h[i];
			printf("  %-28s %7.1f cycles/iter\n", names[mode], s / blocks / 1024.0);
		}
	}
	return 0;
}
