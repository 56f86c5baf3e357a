// DMMA trial for the Schur block products (VERDICT r01 N1 / BASELINE north_star: "tensor cores only for the dense per-landmark 3x6/6x6
// block products where they genuinely are small dense contractions ... evidenced by tensor-pipe utilisation").
//
// The contraction (impl/schur.h:237-240): every reduced block  S_ij -= sum_k W_ik (6x3) * HApf_jk^t (3x6)  over the landmarks k that
// both edges i and j observe.  Per block that is ONE small GEMM  [W_i1 W_i2 ...] (6 x 3K) * [H_j1 H_j2 ...]^t (3K x 6)  with a
// gathered inner dimension 3K (K ~ 33 on the depth-5 areas of config #5).  Two kernels on the same gathered operands:
//   fma  : the product path's shape -- one thread per block ROW (6 outputs), W row in registers, scalar DFMA     (lm_kernel.cu stage_schur)
//   dmma : one warp per block, mma.sync.aligned.m8n8k4.row.col.f64 -- the 6x6 output padded to 8x8, 4 landmark columns... i.e. k = 4
//          scalar inner indices per instruction (3K padded to a multiple of 4), A/B fragments gathered straight from the blocks.
// Both walk the same term lists in the same order per output element up to the association inside one MMA (k = 4 partial sums), so the
// results agree to rounding (checked).  Build: nvcc -arch=sm_100a -O3 -lineinfo -o dmma_schur dmma_schur.cu ; run: ./dmma_schur [blocks] [terms]
// ncu: --metrics sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,
//                sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// blocks of W / H: 6x3 row-major (18 doubles).  term lists: for block b, terms [ptr[b], ptr[b+1]) with (ia[t], ib[t]) = indices of the
// W block (rows of the output) and of the H block (columns of the output).
__global__ void __launch_bounds__(256) schur_fma(const double *__restrict__ W, const double *__restrict__ H, const int *__restrict__ ptr, const int *__restrict__ ia,
                                                 const int *__restrict__ ib, double *__restrict__ S, int nblk) {
	const int it = blockIdx.x * blockDim.x + threadIdx.x;  // (block, row)
	if (it >= nblk * 6) return;
	const int b = it / 6, r = it - 6 * b;
	double acc[6] = {0, 0, 0, 0, 0, 0};
	for (int t = ptr[b]; t < ptr[b + 1]; t++) {
		const double *w = W + (size_t)ia[t] * 18 + r * 3, *h = H + (size_t)ib[t] * 18;
		const double w0 = w[0], w1 = w[1], w2 = w[2];
#pragma unroll
		for (int c = 0; c < 6; c++) acc[c] += w0 * h[c * 3] + w1 * h[c * 3 + 1] + w2 * h[c * 3 + 2];
	}
#pragma unroll
	for (int c = 0; c < 6; c++) S[(size_t)b * 36 + r * 6 + c] -= acc[c];
}

// One warp per block.  m8n8k4 f64 fragments: A (8x4, row): lane l holds A[l/4][l%4];  B (4x8, col): lane l holds B[l%4][l/4];
// C (8x8): lane l holds C[l/4][2*(l%4)], C[l/4][2*(l%4)+1].  Inner index q = 0..3K-1 -> (term q/3, component q%3), four per instruction.
__global__ void __launch_bounds__(256) schur_dmma(const double *__restrict__ W, const double *__restrict__ H, const int *__restrict__ ptr, const int *__restrict__ ia,
                                                  const int *__restrict__ ib, double *__restrict__ S, int nblk) {
	const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (warp >= nblk) return;
	const int t0 = ptr[warp], nq = 3 * (ptr[warp + 1] - t0);
	const int row = lane >> 2, kk = lane & 3;   // A: output row / inner offset;  B: output column = lane >> 2, inner offset = lane & 3
	double c0 = 0, c1 = 0;
	for (int q0 = 0; q0 < nq; q0 += 4) {
		const int q = q0 + kk;
		double a = 0, b = 0;
		if (q < nq) {
			const int t = t0 + q / 3, comp = q - 3 * (q / 3);
			if (row < 6) {
				a = W[(size_t)ia[t] * 18 + row * 3 + comp];
				b = H[(size_t)ib[t] * 18 + row * 3 + comp];   // (for B, `row` is the output COLUMN)
			}
		}
		asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
	}
	const int col = 2 * kk;
	if (row < 6 && col < 6) {
		S[(size_t)warp * 36 + row * 6 + col] -= c0;
		S[(size_t)warp * 36 + row * 6 + col + 1] -= c1;
	}
}

int main(int argc, char **argv) {
	const int nblk = argc > 1 ? atoi(argv[1]) : 400000, K = argc > 2 ? atoi(argv[2]) : 33;
	const int nW = 200000;   // 200k W blocks + 200k H blocks x 144 B = 58 MB: L2-resident, gathered at random like the term lists
	std::vector<double> hW((size_t)nW * 18), hH((size_t)nW * 18);
	srand(1);
	for (auto &v : hW) v = rand() / (double)RAND_MAX - 0.5;
	for (auto &v : hH) v = rand() / (double)RAND_MAX - 0.5;
	std::vector<int> ptr(nblk + 1), ia((size_t)nblk * K), ib((size_t)nblk * K);
	for (int b = 0; b <= nblk; b++) ptr[b] = b * K;
	for (size_t t = 0; t < ia.size(); t++) { ia[t] = rand() % nW; ib[t] = rand() % nW; }
	double *dW, *dH, *dS1, *dS2;
	int *dptr, *dia, *dib;
	CK(cudaMalloc(&dW, hW.size() * 8)); CK(cudaMalloc(&dH, hH.size() * 8)); CK(cudaMalloc(&dS1, (size_t)nblk * 36 * 8)); CK(cudaMalloc(&dS2, (size_t)nblk * 36 * 8));
	CK(cudaMalloc(&dptr, ptr.size() * 4)); CK(cudaMalloc(&dia, ia.size() * 4)); CK(cudaMalloc(&dib, ib.size() * 4));
	CK(cudaMemcpy(dW, hW.data(), hW.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dH, hH.data(), hH.size() * 8, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(dptr, ptr.data(), ptr.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dia, ia.data(), ia.size() * 4, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(dib, ib.data(), ib.size() * 4, cudaMemcpyHostToDevice));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	float ms_fma = 0, ms_dmma = 0;
	const int reps = 10;
	for (int pass = 0; pass < 2; pass++) {  // pass 0 = warm-up
		CK(cudaMemset(dS1, 0, (size_t)nblk * 36 * 8)); CK(cudaMemset(dS2, 0, (size_t)nblk * 36 * 8));
		CK(cudaEventRecord(e0));
		for (int r = 0; r < (pass ? reps : 1); r++) schur_fma<<<(nblk * 6 + 255) / 256, 256>>>(dW, dH, dptr, dia, dib, dS1, nblk);
		CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms_fma, e0, e1));
		CK(cudaEventRecord(e0));
		for (int r = 0; r < (pass ? reps : 1); r++) schur_dmma<<<(nblk * 32 + 255) / 256, 256>>>(dW, dH, dptr, dia, dib, dS2, nblk);
		CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms_dmma, e0, e1));
	}
	CK(cudaGetLastError());
	std::vector<double> s1((size_t)nblk * 36), s2((size_t)nblk * 36);
	CK(cudaMemcpy(s1.data(), dS1, s1.size() * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(s2.data(), dS2, s2.size() * 8, cudaMemcpyDeviceToHost));
	double worst = 0, big = 0;
	for (size_t i = 0; i < s1.size(); i++) { worst = fmax(worst, fabs(s1[i] - s2[i])); big = fmax(big, fabs(s1[i])); }
	const double trip = (double)nblk * K, flop = trip * 2 * 6 * 6 * 3;
	printf("{\"blocks\": %d, \"terms_per_block\": %d, \"triplets\": %.0f, \"fma_ms\": %.4f, \"dmma_ms\": %.4f, \"fma_ns_per_triplet\": %.3f, \"dmma_ns_per_triplet\": %.3f, "
	       "\"fma_gflops\": %.1f, \"dmma_gflops_useful\": %.1f, \"dmma_over_fma_speedup\": %.3f, \"max_abs_diff\": %.3e, \"max_abs\": %.3e}\n",
	       nblk, K, trip, ms_fma / reps, ms_dmma / reps, 1e6 * ms_fma / reps / trip, 1e6 * ms_dmma / reps / trip, flop / (ms_fma / reps) / 1e6, flop / (ms_dmma / reps) / 1e6,
	       ms_fma / ms_dmma, worst, big);
	return 0;
}
