// B200-native (sm_100a) Levenberg–Marquardt local-area optimiser: one persistent CTA ("team") per flattened
// optimize_edges() problem, the WHOLE LM loop of MRPT/srba's impl/optimize_edges.h:245-751 on the device:
//   numeric spanning tree (impl/spantree_update_numeric.h:19-81) -> Jacobians (impl/jacobians.h:207-354,363-494,
//   644-744,888-1013; models/sensors.h) -> block-sparse Hessian H_Ap/H_f/H_Apf (impl/sparse_hessian_update_numeric.h:
//   22-60) -> gradient (impl/compute_minus_gradient.h:20-91) -> Schur complement (impl/schur.h:180-311) -> dense
//   Cholesky + back-substitution (impl/lev-marq_solvers.h:474-568) -> manifold update / accept / reject / relinearise
//   (impl/optimize_edges.h:454-696).
// Inputs are read once from the packed image in HBM, every intermediate lives in a per-problem workspace that stays
// L1/L2-resident for the kernel's lifetime, results are written once.  All arithmetic is FP64 (the reference's).
// Hessian blocks are GATHERED: each (block,element) is owned by one thread that walks the block's term list in
// ascending observation order == the reference's summation order, so results are deterministic and differ from the
// CPU path only by FMA contraction.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>
#include <stdint.h>

#include "dev_types.h"

#ifndef SRBA_TMA_STAGING
#define SRBA_TMA_STAGING 1   // arena staging by cp.async.bulk + mbarrier (0: 16-byte load/store loop)
#endif
#ifndef SRBA_SCHUR_DMMA
// SE(3) Schur block products on the FP64 tensor pipe (mma.sync.m8n8k4.f64, one warp per block) instead of one thread per block row with
// scalar DFMA.  Measured (profiles/r02_dmma_schur.txt): in isolation at full occupancy the DMMA kernel is 1.9x faster (tensor pipe 28 %
// active); INSIDE the product kernels it is slower -- config #5, 1 000 areas: 373 vs 345 ms (one CTA per area), 603 vs 540 ms (wave) --
// because the stage is bound by dependent gathers at low occupancy, where 1 024 independent row-threads keep more loads in flight than
// 32 warps.  Parity tests pass with either setting.  Default: off.
#define SRBA_SCHUR_DMMA 0
#endif

// One TEAM solves one problem.  Normally the team is one CTA; lm_kernel_cluster.cu compiles this file again with
// SRBA_CLUSTER = 8 so that a thread-block CLUSTER of 8 CTAs (8 SMs) shares a large problem: stage loops stride over the
// whole cluster, barriers are cluster barriers, reductions are exchanged through distributed shared memory, and every
// mutable array lives in the (L2-resident) global workspace.  CTA size is a launch parameter (srba_kernel_cfg).
#ifdef SRBA_CLUSTER
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define LM_RANK ((int)cg::this_cluster().block_rank())
#define LM_NCTA SRBA_CLUSTER
#define LM_TID ((int)(cg::this_cluster().block_rank() * blockDim.x + threadIdx.x))
#define LM_THREADS ((int)(blockDim.x * SRBA_CLUSTER))
#define LM_BARRIER() cg::this_cluster().sync()
#elif defined(SRBA_WAVE)
// lm_kernel_wave.cu compiles this file a third time: the team of a stage is the ROW OF CTAs (blockIdx.x) a grid-wide stage kernel
// assigns to one area (blockIdx.y); barrier-separated phases of a stage are separate launches (lm_wave.cuh), so LM_BARRIER is
// only ever reached by the one-CTA-per-area kernels (solver, list compaction).
#include <cooperative_groups.h>
__shared__ int g_wave_gx;  // CTAs of this area's team in the running stage kernel (wave_team, lm_wave.cuh)
#define LM_RANK 0
#define LM_NCTA 1
#define LM_TID ((int)(blockIdx.x * blockDim.x + threadIdx.x))
#define LM_THREADS ((int)blockDim.x * g_wave_gx)
#define LM_BARRIER() __syncthreads()
#define LM_STAGE __device__ __forceinline__
#else
#define LM_RANK 0
#define LM_NCTA 1
#define LM_TID ((int)threadIdx.x)
#define LM_THREADS ((int)blockDim.x)
#define LM_BARRIER() __syncthreads()
#endif
#define LM_MAX_WARPS 32
#define MAXO 4

namespace {

// ------------------------------------------------------------------------------------------------ pose algebra
// The LM program is latency-bound on a handful of warps, and ncu shows the instruction cache as a first-order limiter
// (icc hit rate 65 %, "no instruction" the second largest stall): the transcendental helpers are therefore kept OUT of
// line so that the kernel holds one copy of each instead of one per call site.
__device__ __noinline__ void sincos_ni(double a, double *s, double *c) { sincos(a, s, c); }
__device__ __noinline__ double fmod_2pi_ni(double a) { return fmod(a, 2.0 * M_PI); }
__device__ __forceinline__ double wrapTo2Pi(double a) {  // mrpt::math::wrapTo2Pi
	// fmod(a, 2pi) costs ~330 cycles on B200 (tools/microbench); for |a| < 4pi -- every angle this path produces -- the
	// same value comes from at most one subtraction, which is EXACT (Sterbenz), so the result is bit-identical to fmod's.
	const double T = 2.0 * M_PI;
	const bool was_neg = a < 0;
	const double m = fabs(a);
	double r;
	if (m < T) r = a;
	else if (m < 2.0 * T) r = was_neg ? a + T : a - T;
	else r = fmod_2pi_ni(a);
	if (was_neg) r += T;
	return r;
}
__device__ __forceinline__ double wrapToPi(double a) { return wrapTo2Pi(a + M_PI) - M_PI; }

template <int P>
struct PoseOps;

template <>
struct PoseOps<3> {  // CPose2D: x,y,phi
	static constexpr int S = 3;
	__device__ static void identity(double *p) { p[0] = 0; p[1] = 0; p[2] = 0; }
	__device__ static void compose(double *o, const double *a, const double *b) {
		double s, c;
		sincos(a[2], &s, &c);
		const double nx = a[0] + b[0] * c - b[1] * s;
		const double ny = a[1] + b[0] * s + b[1] * c;
		const double nphi = wrapToPi(a[2] + b[2]);
		o[0] = nx; o[1] = ny; o[2] = nphi;
	}
	// compose(identity, b) without the sincos(0): 0 + b.x*1 - b.y*0 is exactly b.x; the angle still goes through the wrap
	__device__ static void compose_from_identity(double *o, const double *b) { o[0] = b[0]; o[1] = b[1]; o[2] = wrapToPi(0.0 + b[2]); }
	__device__ static void inverse(double *o, const double *a) {
		double s, c;
		sincos(a[2], &s, &c);
		const double x = a[0], y = a[1];
		o[0] = -x * c - y * s;
		o[1] = x * s - y * c;
		o[2] = -a[2];
	}
	__device__ static void pseudo_exp(double *o, const double *v) { o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; }
};

template <>
struct PoseOps<6> {  // CPose3D: R row-major (9) + t (3)
	static constexpr int S = 12;
	__device__ static void identity(double *p) {
		p[0] = 1; p[1] = 0; p[2] = 0; p[3] = 0; p[4] = 1; p[5] = 0; p[6] = 0; p[7] = 0; p[8] = 1; p[9] = 0; p[10] = 0; p[11] = 0;
	}
	__device__ static void compose(double *o, const double *a, const double *b) {
		double t[12];
#pragma unroll
		for (int r = 0; r < 3; r++) t[9 + r] = a[9 + r] + a[3 * r + 0] * b[9] + a[3 * r + 1] * b[10] + a[3 * r + 2] * b[11];
#pragma unroll
		for (int r = 0; r < 3; r++)
#pragma unroll
			for (int c = 0; c < 3; c++) t[3 * r + c] = a[3 * r + 0] * b[c] + a[3 * r + 1] * b[3 + c] + a[3 * r + 2] * b[6 + c];
#pragma unroll
		for (int i = 0; i < 12; i++) o[i] = t[i];
	}
	__device__ static void compose_from_identity(double *o, const double *b) {
		double id[12];
		identity(id);
		compose(o, id, b);  // I*R and 0 + I*t reproduce the reference's products exactly; no shortcut needed for parity
	}
	__device__ static void inverse(double *o, const double *a) {
		double t[12];
#pragma unroll
		for (int r = 0; r < 3; r++)
#pragma unroll
			for (int c = 0; c < 3; c++) t[3 * r + c] = a[3 * c + r];
#pragma unroll
		for (int r = 0; r < 3; r++) t[9 + r] = -(t[3 * r + 0] * a[9] + t[3 * r + 1] * a[10] + t[3 * r + 2] * a[11]);
#pragma unroll
		for (int i = 0; i < 12; i++) o[i] = t[i];
	}
	// CPose3D::exp(v, pseudo_exponential=true): t copied, R = Rodrigues(w) with MRPT's small-angle series.
	__device__ static void pseudo_exp(double *o, const double *v) {
		o[9] = v[0]; o[10] = v[1]; o[11] = v[2];
		const double w0 = v[3], w1 = v[4], w2 = v[5];
		const double theta_sq = w0 * w0 + w1 * w1 + w2 * w2;
		const double theta = sqrt(theta_sq);
		double A, B;
		if (theta_sq < 1e-8) { A = 1.0 - theta_sq / 6.0; B = 0.5; }
		else if (theta_sq < 1e-6) { B = 0.5 - 0.25 * (1.0 / 6.0) * theta_sq; A = 1.0 - theta_sq * (1.0 / 6.0) * (1.0 - (1.0 / 20.0) * theta_sq); }
		else { const double inv_theta = 1.0 / theta; A = sin(theta) * inv_theta; B = (1 - cos(theta)) * (inv_theta * inv_theta); }
		const double wx2 = w0 * w0, wy2 = w1 * w1, wz2 = w2 * w2;
		o[0] = 1.0 - B * (wy2 + wz2);
		o[4] = 1.0 - B * (wx2 + wz2);
		o[8] = 1.0 - B * (wx2 + wy2);
		{ const double a = A * w2, b = B * (w0 * w1); o[1] = b - a; o[3] = b + a; }
		{ const double a = A * w1, b = B * (w0 * w2); o[2] = b + a; o[6] = b - a; }
		{ const double a = A * w0, b = B * (w1 * w2); o[5] = b - a; o[7] = b + a; }
	}
};

__device__ __forceinline__ void compose_point3(const double *P, const double *l, double *g) {
	const double x = P[0] * l[0] + P[1] * l[1] + P[2] * l[2] + P[9];
	const double y = P[3] * l[0] + P[4] * l[1] + P[5] * l[2] + P[10];
	const double z = P[6] * l[0] + P[7] * l[1] + P[8] * l[2] + P[11];
	g[0] = x; g[1] = y; g[2] = z;
}
__device__ __forceinline__ void inverse_compose_point3(const double *P, const double *g, double *l) {
	const double dx = g[0] - P[9], dy = g[1] - P[10], dz = g[2] - P[11];
	const double x = P[0] * dx + P[3] * dy + P[6] * dz;
	const double y = P[1] * dx + P[4] * dy + P[7] * dz;
	const double z = P[2] * dx + P[5] * dy + P[8] * dz;
	l[0] = x; l[1] = y; l[2] = z;
}
// C(n x m) = A(n x k) * B(k x m), row-major, left-to-right accumulation like the CPU restatement
template <int N, int K, int M>
__device__ __forceinline__ void matmul(double *C, const double *A, const double *B) {
#pragma unroll
	for (int r = 0; r < N; r++)
#pragma unroll
		for (int c = 0; c < M; c++) {
			double s = 0;
#pragma unroll
			for (int t = 0; t < K; t++) s += A[r * K + t] * B[t * M + c];
			C[r * M + c] = s;
		}
}

template <int OBSK> struct ObsDims;
template <> struct ObsDims<SRBA_OBS_RELPOSES2D> { static constexpr int O = 3; };
template <> struct ObsDims<SRBA_OBS_CARTESIAN3D> { static constexpr int O = 3; };
template <> struct ObsDims<SRBA_OBS_MONO> { static constexpr int O = 2; };
template <> struct ObsDims<SRBA_OBS_STEREO> { static constexpr int O = 4; };

// ------------------------------------------------------------------------------------------------ views
// Every array is reached through a generic pointer that targets either the CTA's shared-memory arena (when the packer
// placed it there, dev_types.h) or the problem's blob in global memory.
struct View {
	const DevProblem *h;
	// inputs
#define X(T, name) const T *name;
	SRBA_IMG_ARRAYS(X)
#undef X
	// workspace
#define X(T, name) T *name;
	SRBA_WS_ARRAYS(X)
#undef X
	double *resid[2];
	const double *S;  // reduced-system blocks: Sblk, or HAp itself when there is no Schur term (h->s_alias_hap)
	// results
	srba_stats_t *stats;
	double *r_edge_pose, *r_lm_pos, *r_slot_pose, *r_Hf_diag, *r_dense, *r_trace, *r_trace_delta;
	uint8_t *r_obs_valid, *r_lm_inv;
};

__device__ void make_view(View &v, const DevProblem *h, const char *img, char *ws, char *res, char *arena) {
	v.h = h;
#define X(T, name) v.name = reinterpret_cast<const T *>(h->img_sm[SRBA_I_##name] != SRBA_NOT_IN_SMEM ? arena + h->img_sm[SRBA_I_##name] : img + h->img_off[SRBA_I_##name]);
	SRBA_IMG_ARRAYS(X)
#undef X
#define X(T, name) v.name = reinterpret_cast<T *>(h->ws_sm[SRBA_W_##name] != SRBA_NOT_IN_SMEM ? arena + h->ws_sm[SRBA_W_##name] : ws + h->ws_off[SRBA_W_##name]);
	SRBA_WS_ARRAYS(X)
#undef X
	v.resid[0] = v.resid0;
	v.resid[1] = v.resid1;
	v.S = h->s_alias_hap ? v.HAp : v.Sblk;
	v.stats = reinterpret_cast<srba_stats_t *>(res + h->res_stats);
	v.r_edge_pose = reinterpret_cast<double *>(res + h->res_edge_pose);
	v.r_lm_pos = reinterpret_cast<double *>(res + h->res_lm_pos);
	v.r_slot_pose = reinterpret_cast<double *>(res + h->res_slot_pose);
	v.r_Hf_diag = reinterpret_cast<double *>(res + h->res_Hf_diag);
	v.r_dense = reinterpret_cast<double *>(res + h->res_dense);
	v.r_trace = reinterpret_cast<double *>(res + h->res_trace);
	v.r_trace_delta = reinterpret_cast<double *>(res + h->res_trace_delta);
	v.r_obs_valid = reinterpret_cast<uint8_t *>(res + h->res_obs_valid);
	v.r_lm_inv = reinterpret_cast<uint8_t *>(res + h->res_lm_inv);
}

// ------------------------------------------------------------------------------------------------ team (one CTA)
enum { ST_INIT = 0, ST_SCHUR, ST_CHOL, ST_BACKSUB, ST_FEATURES, ST_UPDATE, ST_CHAIN, ST_RESID, ST_RHO, ST_JACOB, ST_HESS, ST_GRAD, ST_CTRL, ST_OUT, ST_COUNT };

struct Shared {
	DevProblem hdr;         // on-chip copy of the problem header (model parameters, counts, placement tables)
	View view;              // resolved array pointers (arena or global), one copy per CTA
	double red[3][2 * LM_MAX_WARPS];
	unsigned long long stage_cycles[ST_COUNT];
	unsigned long long dbg[8];  // solver sub-timers: gather, factor loop, back-substitution loop, critical-thread busy time
	unsigned long long t_last;
	double U11[64];         // diagonal block of the blocked Cholesky
	double xblk[8];
	double bcast;
	// LM state (impl/optimize_edges.h:361-363,452-459)
	double lambda, nu, err, rmse, rho, new_err;
	int iter, stop, trials, accepted, relin, cholfail, cur, chol_ok, status, n_listed, n_refresh, need_outer, go;
	double sensor_inv[12];  // (-sensor_pose) hoisted: pose_robot2sensor (srba_options_sensor_pose.h:108-113)
};

// Team-wide reductions with ONE barrier: warp partials go to one of three scratch rows (a row is reused only after at
// least one other barrier), then every thread combines them in the same fixed order -- identical result on all threads.
// In a cluster every CTA publishes its warp partials, one cluster barrier, then every thread of every CTA combines all
// partials (CTA 0 first) -- the same bits everywhere, so all CTAs take the same LM decisions.
// two sums at once (second value in the upper half of the row)
__device__ __forceinline__ double team_sum2(double v, double w, double &wsum, Shared &sh, int buf) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) { v += __shfl_down_sync(0xffffffffu, v, o); w += __shfl_down_sync(0xffffffffu, w, o); }
	if ((threadIdx.x & 31) == 0) { sh.red[buf][threadIdx.x >> 5] = v; sh.red[buf][LM_MAX_WARPS + (threadIdx.x >> 5)] = w; }
	LM_BARRIER();
	double s = 0, t = 0;
	for (int c = 0; c < LM_NCTA; c++) {
#ifdef SRBA_CLUSTER
		const double *red = cg::this_cluster().map_shared_rank(&sh.red[buf][0], c);
#else
		const double *red = &sh.red[buf][0];
#endif
		for (int i = 0; i < (int)blockDim.x / 32; i++) { s += red[i]; t += red[LM_MAX_WARPS + i]; }
	}
	wsum = t;
	return s;
}
__device__ __forceinline__ double team_max(double v, Shared &sh, int buf) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
	if ((threadIdx.x & 31) == 0) sh.red[buf][threadIdx.x >> 5] = v;
	LM_BARRIER();
	double s = 0;
	for (int c = 0; c < LM_NCTA; c++) {
#ifdef SRBA_CLUSTER
		const double *red = cg::this_cluster().map_shared_rank(&sh.red[buf][0], c);
#else
		const double *red = &sh.red[buf][0];
#endif
		for (int i = 0; i < (int)blockDim.x / 32; i++) s = (c == 0 && i == 0) ? red[0] : fmax(s, red[i]);
	}
	return s;
}

// solver_engine<false,false> (impl/lev-marq_solvers.h:27-209): the whole system is factorised and solve() leaves the
// caller's gradient untouched, whereas the Schur variants reduce it in place (schur.h:248-265,283-296).  The device program
// is the same (landmark blocks eliminated first); it restarts from a pristine copy of the gradient at every trial.
__device__ __forceinline__ bool no_schur_solver(const DevProblem *h) {
	return h->model.solver_kind == SRBA_SOLVER_NO_SCHUR_SPARSE_CHOLESKY && h->n_unk_lms > 0;
}

// ------------------------------------------------------------------------------------------------ stages
// Stage functions are OUT OF LINE in the monolithic one-CTA / cluster program (its hot loop must fit the instruction cache) and
// INLINED into the single-stage kernels of the wave pipeline (each kernel then loads only the View fields its stage uses).
#ifndef LM_STAGE
#define LM_STAGE __device__ __noinline__
#endif
// K1: numeric spanning tree.  update_numeric_only_all_from_node (impl/spantree_update_numeric.h:19-81)
template <int P>
LM_STAGE void stage_chain(const View &vs, int n_refresh) {
	const View v = vs;  // field-wise copy into registers: no aliasing with the stores below
	constexpr int S = PoseOps<P>::S;
	// n_refresh < 0: every pair (update_numeric(set,false)); otherwise the compact list of pairs holding a listed pose
	const int n = n_refresh < 0 ? v.h->n_pairs : n_refresh;
	for (int q = LM_TID; q < n; q += LM_THREADS) {
		const int t = n_refresh < 0 ? q : v.refresh_idx[q];
		double accum[S], tmp[S];
		const int k0 = v.pair_path_ptr[t], k1 = v.pair_path_ptr[t + 1];
		if (k0 == k1) PoseOps<P>::identity(accum);
		for (int k = k0; k < k1; k++) {
			const double *ep = v.edge_cur + (size_t)v.path_edge[k] * S;
			const double *rhs = ep;
			if (!v.path_dir[k]) { PoseOps<P>::inverse(tmp, ep); rhs = tmp; }
			if (k == k0) PoseOps<P>::compose_from_identity(accum, rhs);
			else PoseOps<P>::compose(accum, accum, rhs);
		}
		double *o = v.slot + (size_t)(2 * t) * S;
#pragma unroll
		for (int i = 0; i < S; i++) o[i] = accum[i];
		PoseOps<P>::inverse(tmp, accum);
#pragma unroll
		for (int i = 0; i < S; i++) o[S + i] = tmp[i];
	}
}

// sensor_model<>::eval_jacob_dh_dx on the sensor-frame point (models/sensors.h:85-110,230-280,383-393,804-814)
template <int OBSK>
__device__ __forceinline__ bool eval_dh_dx(const DevProblem *h, double *dh_dx, const double *x) {
	constexpr int O = ObsDims<OBSK>::O;
	if (OBSK == SRBA_OBS_RELPOSES2D || OBSK == SRBA_OBS_CARTESIAN3D) {
#pragma unroll
		for (int i = 0; i < O * 3; i++) dh_dx[i] = 0;
		dh_dx[0] = 1; dh_dx[4] = 1; dh_dx[8] = 1;
		return true;
	}
	if (x[2] <= 0) return false;
	{
		const double pz_inv = 1.0 / x[2], pz_inv2 = pz_inv * pz_inv;
		const double fx = h->model.cam_left[0], fy = h->model.cam_left[1];
		dh_dx[0] = fx * pz_inv; dh_dx[1] = 0; dh_dx[2] = -fx * x[0] * pz_inv2;
		dh_dx[3] = 0; dh_dx[4] = fy * pz_inv; dh_dx[5] = -fy * x[1] * pz_inv2;
	}
	if (OBSK == SRBA_OBS_STEREO) {
		double r[3];
		compose_point3(h->model.right_cam_R2L, x, r);
		const double pz_inv = 1.0 / r[2], pz_inv2 = pz_inv * pz_inv;
		const double fx = h->model.cam_right[0], fy = h->model.cam_right[1];
		dh_dx[6] = fx * pz_inv; dh_dx[7] = 0; dh_dx[8] = -fx * r[0] * pz_inv2;
		dh_dx[9] = 0; dh_dx[10] = fy * pz_inv; dh_dx[11] = -fy * r[1] * pz_inv2;
	}
	return true;
}
// jacob_dh_dx_rotate (srba_options_sensor_pose.h:124-129): dh_dx := dh_dx * R_s^T
template <int O>
__device__ __forceinline__ void dh_dx_rotate(const DevProblem *h, double *dh_dx) {
	if (h->model.sensor_pose_kind != SRBA_SENSORPOSE_SE3) return;
	const double *R = h->model.sensor_pose;
	double t[O * 3];
#pragma unroll
	for (int r = 0; r < O; r++)
#pragma unroll
		for (int c = 0; c < 3; c++) t[3 * r + c] = dh_dx[3 * r + 0] * R[3 * c + 0] + dh_dx[3 * r + 1] * R[3 * c + 1] + dh_dx[3 * r + 2] * R[3 * c + 2];
#pragma unroll
	for (int i = 0; i < O * 3; i++) dh_dx[i] = t[i];
}

// One dh_dAp block: compute_jacobian_dh_dp (impl/jacobians.h:207-354) + compute_jacobian_dAepsDx_deps
// (SE3 x point :363-494, SE2 x relpose :644-744).  Returns false when the sensor model rejects the point.
template <int P, int OBSK>
__device__ bool jacob_dh_dAp(const View &v, int b, double *J) {
	constexpr int S = PoseOps<P>::S;
	constexpr int O = ObsDims<OBSK>::O;
	const int obs = v.blkA_obs[b];
	const int sA = v.blkA_slotA[b];
	const double *D = v.slot + (size_t)v.blkA_slotD[b] * S;
	const double *A = sA >= 0 ? v.slot + (size_t)sA * S : nullptr;
	const bool inverse_dir = !v.blkA_normal_dir[b];
	const double *pdd1 = v.edge_cur + (size_t)v.unk_edge[v.blkA_col[b]] * S;  // k2k_edges[k2k_edge_id].inv_pose
	if (P == 3) {
		double dh_dx[9];
		eval_dh_dx<OBSK>(v.h, dh_dx, nullptr);
		double Xd, Yd, PHIa, bwo[3];
		if (!inverse_dir) {
			Xd = D[0]; Yd = D[1]; PHIa = 0;
			if (A) { PHIa = A[2]; PoseOps<3>::compose(bwo, A, D); }
			else { bwo[0] = D[0]; bwo[1] = D[1]; bwo[2] = D[2]; }
		} else {
			double Dp[3], pinv[3], Ap[3];
			PoseOps<3>::compose(Dp, pdd1, D);
			PoseOps<3>::inverse(pinv, pdd1);
			if (A) PoseOps<3>::compose(Ap, A, pinv);
			else { Ap[0] = pinv[0]; Ap[1] = pinv[1]; Ap[2] = pinv[2]; }
			Xd = Dp[0]; Yd = Dp[1]; PHIa = Ap[2];
			PoseOps<3>::compose(bwo, Ap, Dp);
		}
		double s_ad, c_ad, s_a, c_a;
		sincos_ni(bwo[2], &s_ad, &c_ad);
		sincos_ni(PHIa, &s_a, &c_a);
		const double J0[9] = {c_ad, s_ad, 0, -s_ad, c_ad, 0, 0, 0, 1};
		const double J1[9] = {1, 0, -Xd * s_a - Yd * c_a, 0, 1, Xd * c_a - Yd * s_a, 0, 0, 1};
		const double J2[9] = {c_a, -s_a, 0, s_a, c_a, 0, 0, 0, 1};
		double T0[9], T1[9];
		matmul<3, 3, 3>(T0, dh_dx, J0);
		matmul<3, 3, 3>(T1, T0, J1);
		matmul<3, 3, 3>(J, T1, J2);
		if (inverse_dir) {
#pragma unroll
			for (int i = 0; i < 9; i++) J[i] = -J[i];
		}
		return true;
	} else {
		const double *xji_i = v.lm_cur + (size_t)v.obs_lm[obs] * 3;
		double piwl[12];
		if (A) PoseOps<6>::compose(piwl, A, D);
		else {
#pragma unroll
			for (int i = 0; i < 12; i++) piwl[i] = D[i];
		}
		double xji_l[3];
		compose_point3(piwl, xji_i, xji_l);
		if (v.h->model.sensor_pose_kind == SRBA_SENSORPOSE_SE3) inverse_compose_point3(v.h->model.sensor_pose, xji_l, xji_l);
		double dh_dx[O * 3];
		if (!eval_dh_dx<OBSK>(v.h, dh_dx, xji_l)) return false;
		dh_dx_rotate<O>(v.h, dh_dx);
		double H_ROTA[O * 3], Dq[12];
		if (!inverse_dir) {
#pragma unroll
			for (int i = 0; i < 12; i++) Dq[i] = D[i];
			if (A) matmul<O, 3, 3>(H_ROTA, dh_dx, A);
			else {
#pragma unroll
				for (int i = 0; i < O * 3; i++) H_ROTA[i] = dh_dx[i];
			}
		} else {
			PoseOps<6>::compose(Dq, pdd1, D);
			double Rt[9];
#pragma unroll
			for (int r = 0; r < 3; r++)
#pragma unroll
				for (int c = 0; c < 3; c++) Rt[3 * r + c] = pdd1[3 * c + r];
			if (A) { double T[O * 3]; matmul<O, 3, 3>(T, dh_dx, A); matmul<O, 3, 3>(H_ROTA, T, Rt); }
			else matmul<O, 3, 3>(H_ROTA, dh_dx, Rt);
		}
		double vv[3];
#pragma unroll
		for (int r = 0; r < 3; r++) vv[r] = -Dq[9 + r] - xji_i[0] * Dq[3 * r + 0] - xji_i[1] * Dq[3 * r + 1] - xji_i[2] * Dq[3 * r + 2];
		const double aux[9] = {0, -vv[2], vv[1], vv[2], 0, -vv[0], -vv[1], vv[0], 0};
		double second[O * 3];
		matmul<O, 3, 3>(second, H_ROTA, aux);
#pragma unroll
		for (int r = 0; r < O; r++)
#pragma unroll
			for (int c = 0; c < 3; c++) {
				J[r * 6 + c] = inverse_dir ? -H_ROTA[3 * r + c] : H_ROTA[3 * r + c];
				J[r * 6 + 3 + c] = inverse_dir ? -second[3 * r + c] : second[3 * r + c];
			}
		return true;
	}
}

// One dh_df block: compute_jacobian_dh_df (impl/jacobians.h:888-1013).  Only for point landmarks (SE3 models here).
template <int OBSK>
__device__ bool jacob_dh_df(const View &v, int b, double *J) {
	constexpr int O = ObsDims<OBSK>::O;
	const int obs = v.blkF_obs[b];
	const int s = v.blkF_slot[b];
	const double *Pp = s >= 0 ? v.slot + (size_t)s * 12 : nullptr;
	const double *xji_i = v.lm_cur + (size_t)v.obs_lm[obs] * 3;
	double xji_l[3] = {xji_i[0], xji_i[1], xji_i[2]};
	if (Pp) compose_point3(Pp, xji_i, xji_l);
	if (v.h->model.sensor_pose_kind == SRBA_SENSORPOSE_SE3) inverse_compose_point3(v.h->model.sensor_pose, xji_l, xji_l);
	double dh_dx[O * 3];
	if (!eval_dh_dx<OBSK>(v.h, dh_dx, xji_l)) return false;
	dh_dx_rotate<O>(v.h, dh_dx);
	if (Pp) matmul<O, 3, 3>(J, dh_dx, Pp);
	else {
#pragma unroll
		for (int i = 0; i < O * 3; i++) J[i] = dh_dx[i];
	}
	return true;
}

// K2a: recompute_all_Jacobians (impl/jacobians.h:1083-1117) with the validity protocol of SURVEY.md §9-5:
// the first block of a row (in the reference's column-major walk) whose sensor model fails zeroes itself and clears the
// row's validity byte; blocks walked before it keep their fresh value, blocks walked after it keep their PREVIOUS value.
// (PREVIOUS = previous linearisation of the same call; at the first linearisation of a call they are zero, see below.)
// `first` additionally records which numeric-ST entries the walk registers in list_of_required_num_poses (:226-230,:902-903).
// `phase`: 0 = the whole stage with team barriers between its phases (one CTA / one cluster per area); 1, 2, 3 = only that
// barrier-separated phase (grid-wide wave mode, lm_wave.cuh: the launch boundary is the barrier).
template <int P, int OBSK>
LM_STAGE void stage_jacobians(const View &vs, bool first, int phase = 0) {
	const View v = vs;  // field-wise copy into registers: no aliasing with the stores below
	constexpr int O = ObsDims<OBSK>::O;
	const int nA = v.h->n_blkA, nF = (OBSK == SRBA_OBS_RELPOSES2D) ? 0 : v.h->n_blkF;
	const bool can_fail = (OBSK == SRBA_OBS_MONO || OBSK == SRBA_OBS_STEREO);
	if (phase == 0 || phase == 1)
		for (int i = LM_TID; i < v.h->n_obs; i += LM_THREADS) v.valid[i] = 1;
	if (phase == 0) LM_BARRIER();
	// pass 1: evaluate every block into its slot of JA/JF scratch (fresh), remember ok flags
	// (two loops, not one loop with a branch on the block kind: warps stay convergent)
	if (phase == 0 || phase == 2) {
		for (int b = LM_TID; b < nA; b += LM_THREADS) {
			double J[O * P];
			const bool ok = jacob_dh_dAp<P, OBSK>(v, b, J);
			v.blk_ok[b] = ok;
			if (!can_fail) {
				double *dst = v.JA + (size_t)b * O * P;
#pragma unroll
				for (int i = 0; i < O * P; i++) dst[i] = J[i];
			}
		}
		if (P == 6)
			for (int f = LM_TID; f < nF; f += LM_THREADS) {
				double J[O * 3];
				const bool ok = jacob_dh_df<OBSK>(v, f, J);
				v.blk_ok[nA + f] = ok;
				if (!can_fail) {
					double *dst = v.JF + (size_t)f * O * 3;
#pragma unroll
					for (int i = 0; i < O * 3; i++) dst[i] = J[i];
				}
			}
	}
	if (phase == 2) return;
	if (can_fail) {
		if (phase == 0) LM_BARRIER();
		// pass 2: commit according to the walk order of each row
		if (phase == 0 || phase == 3)
		for (int part = 0; part < 2; part++)   // dh_dAp blocks, then dh_df blocks: no warp mixes the two kinds
		for (int b = (part ? nA : 0) + LM_TID; b < (part ? nA + nF : nA); b += LM_THREADS) {
			const int obs = b < nA ? v.blkA_obs[b] : v.blkF_obs[b - nA];
			int first_bad = -1, my_pos = -1;
			for (int k = v.row_ptr[obs]; k < v.row_ptr[obs + 1]; k++) {
				const int rb = v.row_blk[k];
				if (rb == b) my_pos = k;
				if (first_bad < 0 && !v.blk_ok[rb]) first_bad = k;
			}
			const bool visited_valid = (first_bad < 0 || my_pos <= first_bad);
			if (b < nA) {
				double *dst = v.JA + (size_t)b * O * P;
				if (first_bad < 0 || my_pos < first_bad) {
					double J[O * P];
					jacob_dh_dAp<P, OBSK>(v, b, J);
#pragma unroll
					for (int i = 0; i < O * P; i++) dst[i] = J[i];
				} else if (my_pos == first_bad || first) {
					// the deciding block is zero-filled (:324-326).  Later blocks of the row keep their value from the previous
					// linearisation; on the FIRST linearisation of a call that would be the value an earlier optimize_edges call
					// left in lin_system (or an uninitialised Eigen matrix): not carried across calls here -- zero, deterministic
#pragma unroll
					for (int i = 0; i < O * P; i++) dst[i] = 0;
					if (my_pos == first_bad) v.valid[obs] = 0;
				}
			} else if (P == 6) {
				const int f = b - nA;
				double *dst = v.JF + (size_t)f * O * 3;
				if (first_bad < 0 || my_pos < first_bad) {
					double J[O * 3];
					jacob_dh_df<OBSK>(v, f, J);
#pragma unroll
					for (int i = 0; i < O * 3; i++) dst[i] = J[i];
				} else if (my_pos == first_bad || first) {
#pragma unroll
					for (int i = 0; i < O * 3; i++) dst[i] = 0;
					if (my_pos == first_bad) v.valid[obs] = 0;
				}
			}
			if (first && visited_valid) {
				if (b < nA) { if (v.blkA_slotA[b] >= 0) v.slot_listed[v.blkA_slotA[b]] = 1; v.slot_listed[v.blkA_slotD[b]] = 1; }
				else if (v.blkF_slot[b - nA] >= 0) v.slot_listed[v.blkF_slot[b - nA]] = 1;
			}
		}
	} else if (first && (phase == 0 || phase == 3)) {
		for (int b = LM_TID; b < nA + nF; b += LM_THREADS) {
			if (b < nA) { if (v.blkA_slotA[b] >= 0) v.slot_listed[v.blkA_slotA[b]] = 1; v.slot_listed[v.blkA_slotD[b]] = 1; }
			else if (v.blkF_slot[b - nA] >= 0) v.slot_listed[v.blkF_slot[b - nA]] = 1;
		}
	}
}

// accum_JtJ element (a,b) of J1^t Lambda J2 (srba_options_noise.h:45-50,103-109)
template <int O>
__device__ __forceinline__ double jtj_elem(const DevProblem *h, const double *J1, int m1, int a, const double *J2, int m2, int b) {
	double s = 0;
	if (h->model.noise_kind == SRBA_NOISE_IDENTITY) {
#pragma unroll
		for (int k = 0; k < O; k++) s += J1[k * m1 + a] * J2[k * m2 + b];
	} else {
#pragma unroll
		for (int k = 0; k < O; k++) {
			double t = 0;
#pragma unroll
			for (int l = 0; l < O; l++) t += J1[l * m1 + a] * h->model.noise_lambda[l * O + k];
			s += t * J2[k * m2 + b];
		}
	}
	return s;
}

// One ROW of a J1^t Lambda J2 term: term[b] = sum_k c_k J2[k][b] with c = column a of J1 (times Lambda); the same products in
// the same order as jtj_elem, with the J1 column loaded once for the M2 outputs.
template <int O, int M2>
__device__ __forceinline__ void jtj_row(const DevProblem *h, const double *J1, int m1, int a, const double *J2, double *term) {
	double c[O];
	if (h->model.noise_kind == SRBA_NOISE_IDENTITY) {
#pragma unroll
		for (int k = 0; k < O; k++) c[k] = J1[k * m1 + a];
	} else {
#pragma unroll
		for (int k = 0; k < O; k++) {
			double t = 0;
#pragma unroll
			for (int l = 0; l < O; l++) t += J1[l * m1 + a] * h->model.noise_lambda[l * O + k];
			c[k] = t;
		}
	}
#pragma unroll
	for (int b = 0; b < M2; b++) {
		double s2 = 0;
#pragma unroll
		for (int k = 0; k < O; k++) s2 += c[k] * J2[k * M2 + b];
		term[b] = s2;
	}
}

// K2b: sparse_hessian_update_numeric x3 (impl/sparse_hessian_update_numeric.h:22-60): terms in ascending observation order,
// invalid rows skipped, identity noise scaled by 1/sigma afterwards.  SE(2): one thread per block element (everything on chip);
// SE(3): one thread per block ROW -- the landmark areas are load-bound (J lives in L2), and a row shares the J1 column and the
// index / validity loads between its 6 (or 3) outputs: 2.4x fewer load instructions for the same sums in the same order.
template <int P, int OBSK>
LM_STAGE void stage_hessian(const View &vs) {
	const View v = vs;  // field-wise copy into registers: no aliasing with the stores below
	constexpr int O = ObsDims<OBSK>::O;
	const DevProblem *h = v.h;
	const double scale = (h->model.noise_kind == SRBA_NOISE_IDENTITY) ? 1.0 / h->model.noise_std : 1.0;
	const bool do_scale = (h->model.noise_kind == SRBA_NOISE_IDENTITY);
	if constexpr (P == 6) {
		const int nApR = h->n_HAp * 6, nApfR = h->n_HApf * 6, nfR = h->n_unk_lms * 3;
		for (int it = LM_TID; it < nApR + nApfR + nfR; it += LM_THREADS) {
			if (it < nApR) {
				const int blk = it / 6, a = it % 6;
				double s[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll 2
				for (int t = v.HAp_term_ptr[blk]; t < v.HAp_term_ptr[blk + 1]; t++) {
					const int ba = v.HAp_term_a[t], bb = v.HAp_term_b[t];
					double term[6];
					jtj_row<O, 6>(h, v.JA + (size_t)ba * O * 6, 6, a, v.JA + (size_t)bb * O * 6, term);
					const bool on = v.valid[v.blkA_obs[ba]];
#pragma unroll
					for (int b = 0; b < 6; b++) s[b] += on ? term[b] : 0.0;
				}
#pragma unroll
				for (int b = 0; b < 6; b++) v.HAp[(size_t)it * 6 + b] = do_scale ? s[b] * scale : s[b];
			} else if (it < nApR + nApfR) {
				const int j = it - nApR;
				const int blk = j / 6, a = j % 6;
				double s[3] = {0, 0, 0};
#pragma unroll 2
				for (int t = v.HApf_term_ptr[blk]; t < v.HApf_term_ptr[blk + 1]; t++) {
					const int ba = v.HApf_term_a[t], bf = v.HApf_term_f[t];
					double term[3];
					jtj_row<O, 3>(h, v.JA + (size_t)ba * O * 6, 6, a, v.JF + (size_t)bf * O * 3, term);
					const bool on = v.valid[v.blkA_obs[ba]];
#pragma unroll
					for (int b = 0; b < 3; b++) s[b] += on ? term[b] : 0.0;
				}
#pragma unroll
				for (int b = 0; b < 3; b++) v.HApf[(size_t)j * 3 + b] = do_scale ? s[b] * scale : s[b];
			} else {
				const int j = it - nApR - nApfR;
				const int k = j / 3, a = j % 3;
				double s[3] = {0, 0, 0};
#pragma unroll 2
				for (int bf = v.colF_ptr[k]; bf < v.colF_ptr[k + 1]; bf++) {
					double term[3];
					jtj_row<O, 3>(h, v.JF + (size_t)bf * O * 3, 3, a, v.JF + (size_t)bf * O * 3, term);
					const bool on = v.valid[v.blkF_obs[bf]];
#pragma unroll
					for (int b = 0; b < 3; b++) s[b] += on ? term[b] : 0.0;
				}
#pragma unroll
				for (int b = 0; b < 3; b++) v.Hf[(size_t)j * 3 + b] = do_scale ? s[b] * scale : s[b];
			}
		}
		return;
	}
	const int nAp = h->n_HAp * P * P, nApf = h->n_HApf * P * 3, nf = h->n_unk_lms * 9;
	for (int it = LM_TID; it < nAp + nApf + nf; it += LM_THREADS) {
		if (it < nAp) {
			const int blk = it / (P * P), e = it % (P * P), a = e / P, b = e % P;
			double s = 0;
			// terms are added in list order (== the reference's); an invalid row contributes an exact +0.0, so the loads of
			// several terms can be in flight at once (for large areas J lives in L2)
#pragma unroll 4
			for (int t = v.HAp_term_ptr[blk]; t < v.HAp_term_ptr[blk + 1]; t++) {
				const int ba = v.HAp_term_a[t], bb = v.HAp_term_b[t];
				const double term = jtj_elem<O>(h, v.JA + (size_t)ba * O * P, P, a, v.JA + (size_t)bb * O * P, P, b);
				s += v.valid[v.blkA_obs[ba]] ? term : 0.0;
			}
			v.HAp[it] = do_scale ? s * scale : s;
		} else if (it < nAp + nApf) {
			const int j = it - nAp;
			const int blk = j / (P * 3), e = j % (P * 3), a = e / 3, b = e % 3;
			double s = 0;
#pragma unroll 4
			for (int t = v.HApf_term_ptr[blk]; t < v.HApf_term_ptr[blk + 1]; t++) {
				const int ba = v.HApf_term_a[t], bf = v.HApf_term_f[t];
				const double term = jtj_elem<O>(h, v.JA + (size_t)ba * O * P, P, a, v.JF + (size_t)bf * O * 3, 3, b);
				s += v.valid[v.blkA_obs[ba]] ? term : 0.0;
			}
			v.HApf[j] = do_scale ? s * scale : s;
		} else {
			const int j = it - nAp - nApf;
			const int k = j / 9, e = j % 9, a = e / 3, b = e % 3;
			double s = 0;
#pragma unroll 4
			for (int bf = v.colF_ptr[k]; bf < v.colF_ptr[k + 1]; bf++) {
				const double term = jtj_elem<O>(h, v.JF + (size_t)bf * O * 3, 3, a, v.JF + (size_t)bf * O * 3, 3, b);
				s += v.valid[v.blkF_obs[bf]] ? term : 0.0;
			}
			v.Hf[j] = do_scale ? s * scale : s;
		}
	}
}

// reprojection_residuals (impl/reprojection_residuals.h:16-81) + sensor_model<>::observe_error; returns the team-wide
// total squared error INCLUDING the duplicated rows of optimize_edges.h:177-193 (obs_mult).
// Returns THIS THREAD's share of the total; the caller reduces it over the team (team_sum2 / per-CTA partials in wave mode).
template <int P, int OBSK>
LM_STAGE double stage_residuals(const View &vs, double *resid, const double *sensor_inv) {
	const View v = vs;  // field-wise copy into registers: no aliasing with the stores below
	constexpr int S = PoseOps<P>::S;
	constexpr int O = ObsDims<OBSK>::O;
	const DevProblem *h = v.h;
	double local = 0;
	for (int i = LM_TID; i < h->n_obs; i += LM_THREADS) {
		const int s = v.obs_slot[i];
		const double *z = v.obs_z + (size_t)i * O;
		double r[O];
		if (P == 3) {
			double Pp[3] = {0, 0, 0};
			if (s >= 0) { Pp[0] = v.slot[(size_t)s * S]; Pp[1] = v.slot[(size_t)s * S + 1]; Pp[2] = v.slot[(size_t)s * S + 2]; }
			double sn, cs;
			sincos_ni(Pp[2], &sn, &cs);
			r[0] = (z[0] - Pp[0]) * cs + (z[1] - Pp[1]) * sn;
			r[1] = -(z[0] - Pp[0]) * sn + (z[1] - Pp[1]) * cs;
			r[2] = wrapToPi(z[2] - Pp[2]);
		} else {
			double Pp[12];
			if (s >= 0) {
#pragma unroll
				for (int k = 0; k < 12; k++) Pp[k] = v.slot[(size_t)s * 12 + k];
			} else PoseOps<6>::identity(Pp);
			if (h->model.sensor_pose_kind == SRBA_SENSORPOSE_SE3) PoseOps<6>::compose(Pp, sensor_inv, Pp);
			double x[3];
			compose_point3(Pp, v.lm_cur + (size_t)v.obs_lm[i] * 3, x);
			double pred[O];
			if (OBSK == SRBA_OBS_CARTESIAN3D) { pred[0] = x[0]; pred[1] = x[1]; pred[2] = x[2]; }
			else {
				pred[0] = h->model.cam_left[2] + h->model.cam_left[0] * x[0] / x[2];
				pred[1] = h->model.cam_left[3] + h->model.cam_left[1] * x[1] / x[2];
				if (OBSK == SRBA_OBS_STEREO) {
					double rr[3];
					compose_point3(h->model.right_cam_R2L, x, rr);
					pred[2] = h->model.cam_right[2] + h->model.cam_right[0] * rr[0] / rr[2];
					pred[O - 1] = h->model.cam_right[3] + h->model.cam_right[1] * rr[1] / rr[2];
				}
			}
#pragma unroll
			for (int k = 0; k < O; k++) r[k] = z[k] - pred[k];
		}
		double sum_2 = 0;
#pragma unroll
		for (int k = 0; k < O; k++) sum_2 += r[k] * r[k];
		double contrib = sum_2;
		if (h->model.use_robust_kernel) {
			const double nrm = fmax(1e-11, sqrt(sum_2));
			const double kp = h->model.kernel_param;
			const double hub = fabs(2 * (kp * kp) * (sqrt(1 + (nrm / kp) * (nrm / kp)) - 1));  // huber_kernel RbaEngine.h:810-813
			const double w = sqrt(hub) / nrm;
#pragma unroll
			for (int k = 0; k < O; k++) r[k] *= w;
			contrib = (w * w) * sum_2;
		}
#pragma unroll
		for (int k = 0; k < O; k++) resid[(size_t)i * O + k] = r[k];
		local += contrib * (double)v.obs_mult[i];
	}
	return local;
}

// compute_minus_gradient (impl/compute_minus_gradient.h:20-91): one thread per scalar unknown, blocks in ascending obs.
template <int P, int OBSK>
LM_STAGE double stage_gradient(const View &vs, const double *resid) {
	const View v = vs;  // field-wise copy into registers: no aliasing with the stores below
	constexpr int O = ObsDims<OBSK>::O;
	const DevProblem *h = v.h;
	const bool ident = (h->model.noise_kind == SRBA_NOISE_IDENTITY);
	const double scale = ident ? 1.0 / h->model.noise_std : 1.0;
	const int nAp = h->n_unk_edges * P;
	const bool keep = (OBSK != SRBA_OBS_RELPOSES2D) && no_schur_solver(h);
	double gmax = 0;
	for (int it = LM_TID; it < h->n_scalars; it += LM_THREADS) {
		double g = 0;
		if (it < nAp) {
			const int col = it / P, a = it % P;
			for (int b = v.colA_ptr[col]; b < v.colA_ptr[col + 1]; b++) {
				const double *J = v.JA + (size_t)b * O * P;
				const double *r = resid + (size_t)v.blkA_obs[b] * O;
				double s = 0;
				if (ident) {
#pragma unroll
					for (int k = 0; k < O; k++) s += J[k * P + a] * r[k];
				} else {
#pragma unroll
					for (int k = 0; k < O; k++) {
						double t = 0;
#pragma unroll
						for (int l = 0; l < O; l++) t += J[l * P + a] * h->model.noise_lambda[l * O + k];
						s += t * r[k];
					}
				}
				g += s;
			}
		} else {
			const int j = it - nAp, col = j / 3, a = j % 3;
			for (int b = v.colF_ptr[col]; b < v.colF_ptr[col + 1]; b++) {
				const double *J = v.JF + (size_t)b * O * 3;
				const double *r = resid + (size_t)v.blkF_obs[b] * O;
				double s = 0;
				if (ident) {
#pragma unroll
					for (int k = 0; k < O; k++) s += J[k * 3 + a] * r[k];
				} else {
#pragma unroll
					for (int k = 0; k < O; k++) {
						double t = 0;
#pragma unroll
						for (int l = 0; l < O; l++) t += J[l * 3 + a] * h->model.noise_lambda[l * O + k];
						s += t * r[k];
					}
				}
				g += s;
			}
		}
		const double gv = ident ? g * scale : g;
		v.grad[it] = gv;
		if (keep) v.grad0[it] = gv;
		gmax = fmax(gmax, fabs(gv));
	}
	return gmax;  // this thread's share of max |g| (optimize_edges.h:636-640)
}

// Eigen::FullPivLU<3x3>: isInvertible() + inverse()  (impl/schur.h:193-208)
__device__ bool fullpiv_lu_inverse3(const double *Ain, double *inv) {
	double lu[9];
#pragma unroll
	for (int i = 0; i < 9; i++) lu[i] = Ain[i];
	int rowsT[3] = {0, 1, 2}, colsT[3] = {0, 1, 2};
	int nonzero_pivots = 3;
	double maxpivot = 0;
	for (int k = 0; k < 3; k++) {
		int rbig = k, cbig = k;
		double big = 0;
		for (int r = k; r < 3; r++)
			for (int c = k; c < 3; c++)
				if (fabs(lu[3 * r + c]) > big) { big = fabs(lu[3 * r + c]); rbig = r; cbig = c; }
		if (big == 0) { nonzero_pivots = k; for (int i = k; i < 3; i++) { rowsT[i] = i; colsT[i] = i; } break; }
		if (big > maxpivot) maxpivot = big;
		rowsT[k] = rbig; colsT[k] = cbig;
		if (k != rbig) for (int c = 0; c < 3; c++) { const double t = lu[3 * k + c]; lu[3 * k + c] = lu[3 * rbig + c]; lu[3 * rbig + c] = t; }
		if (k != cbig) for (int r = 0; r < 3; r++) { const double t = lu[3 * r + k]; lu[3 * r + k] = lu[3 * r + cbig]; lu[3 * r + cbig] = t; }
		if (k < 2) {
			for (int r = k + 1; r < 3; r++) lu[3 * r + k] /= lu[3 * k + k];
			for (int r = k + 1; r < 3; r++)
				for (int c = k + 1; c < 3; c++) lu[3 * r + c] -= lu[3 * r + k] * lu[3 * k + c];
		}
	}
	const double thr = fabs(maxpivot) * (2.220446049250313e-16 * 3);
	int rank = 0;
	for (int i = 0; i < nonzero_pivots; i++) rank += (fabs(lu[3 * i + i]) > thr) ? 1 : 0;
	if (rank != 3) return false;
	int p[3] = {0, 1, 2}, q[3] = {0, 1, 2};
	for (int k = 0; k < 3; k++) { const int t = p[k]; p[k] = p[rowsT[k]]; p[rowsT[k]] = t; }
	for (int k = 0; k < 3; k++) { const int t = q[k]; q[k] = q[colsT[k]]; q[colsT[k]] = t; }
	for (int col = 0; col < 3; col++) {
		double b[3];
		for (int i = 0; i < 3; i++) b[i] = (p[i] == col) ? 1.0 : 0.0;
		for (int i = 0; i < 3; i++)
			for (int j = 0; j < i; j++) b[i] -= lu[3 * i + j] * b[j];
		for (int i = 2; i >= 0; i--) {
			for (int j = i + 1; j < 3; j++) b[i] -= lu[3 * i + j] * b[j];
			b[i] /= lu[3 * i + i];
		}
		for (int i = 0; i < 3; i++) inv[3 * q[i] + col] = b[i];
	}
	return true;
}

// K3: SchurComplement::numeric_build_reduced_system(lambda) (impl/schur.h:180-268).  Writes the reduced blocks
// S_ji = HAp_ji - sum_k W_jk HApf_ik^T into Sblk (lambda is added when the dense system is filled, as the reference does,
// lev-marq_solvers.h:506-509) and reduces the gradient head IN PLACE (quirk SURVEY.md §9-1).  Without landmark unknowns
// (relative graph-SLAM) the reduced system IS H_Ap and this stage is empty.
template <int P>
LM_STAGE void stage_schur(const View &vs, double lambda, int phase = 0) {  // `phase` as in stage_jacobians
	const View v = vs;  // field-wise copy into registers: no aliasing with the stores below
	const DevProblem *h = v.h;
	const int nf = h->n_unk_lms, n = h->n_dense;
	if (nf == 0 && h->s_alias_hap) return;
	// (1) inv(Hf_kk + lambda I)
	if (phase == 0 || phase == 1)
	for (int k = LM_TID; k < nf; k += LM_THREADS) {
		double M[9], inv[9];
#pragma unroll
		for (int i = 0; i < 9; i++) M[i] = v.Hf[(size_t)k * 9 + i];
		M[0] += lambda; M[4] += lambda; M[8] += lambda;
		const bool ok = fullpiv_lu_inverse3(M, inv);
		v.lm_inv[k] = ok;
		if (ok) {
#pragma unroll
			for (int i = 0; i < 9; i++) v.Hf_inv[(size_t)k * 9 + i] = inv[i];
		}
	}
	if (phase == 0) LM_BARRIER();
	// (2) W_ik = HApf_ik * inv_k
	if (phase == 0 || phase == 2)
	for (int it = LM_TID; it < h->n_HApf * P * 3; it += LM_THREADS) {
		const int blk = it / (P * 3), e = it % (P * 3), a = e / 3, b = e % 3;
		const int k = v.HApf_col[blk];
		if (!v.lm_inv[k]) continue;
		const double *Hb = v.HApf + (size_t)blk * P * 3;
		const double *inv = v.Hf_inv + (size_t)k * 9;
		v.W[it] = Hb[a * 3 + 0] * inv[0 * 3 + b] + Hb[a * 3 + 1] * inv[1 * 3 + b] + Hb[a * 3 + 2] * inv[2 * 3 + b];
	}
	if (phase == 0) LM_BARRIER();
	if (phase == 1 || phase == 2) return;
	// (3) reduced blocks; (4) reduced gradient.  SE(3): one thread per block ROW (W row loaded once for the 6 outputs, index and
	// invertibility loads shared) -- same sums, same order, 2.4x fewer load instructions in the stage that dominates landmark areas
#if SRBA_SCHUR_DMMA
	// SE(3): the reduced blocks on the FP64 TENSOR pipe.  S_ij = H_ij - sum_k W_ik (6x3) HApf_jk^t (3x6) is one small GEMM per block with a
	// gathered inner dimension 3K (K = landmarks seen from both edges, ~33 on depth-5 areas): one warp per block, mma.sync.m8n8k4.f64 with the
	// 6x6 output padded to 8x8 and four inner indices per instruction; A = -W (zero for non-invertible landmarks, schur.h:228), C starts
	// at the H_Ap block.  tools/microbench/dmma_schur.cu + profiles/r02_dmma_schur.txt: 1.9x the one-thread-per-row DFMA version on
	// L2-resident operands (tensor pipe 28 % active vs FP64 pipe 11 %).  The association of the sum inside one MMA (4 terms) is the
	// hardware's: results differ from the scalar order at rounding level.
	if constexpr (P == 6) {
		const int lane = LM_TID & 31, warp = LM_TID >> 5, nwarp = LM_THREADS >> 5;
		const int row = lane >> 2, kk = lane & 3, col = 2 * kk;   // A: output row / inner offset; B: output column = lane >> 2; C: (row, col), (row, col + 1)
		const bool live = row < 6;
		if (!h->s_alias_hap)
			for (int blk = warp; blk < h->n_S; blk += nwarp) {
				const int hb = v.S_HAp[blk];
				double c0 = 0.0, c1 = 0.0;
				if (live && col < 6 && hb >= 0) { c0 = v.HAp[(size_t)hb * 36 + row * 6 + col]; c1 = v.HAp[(size_t)hb * 36 + row * 6 + col + 1]; }
				const int t0 = v.S_term_ptr[blk], nq = 3 * (v.S_term_ptr[blk + 1] - t0);
				for (int q0 = 0; q0 < nq; q0 += 4) {
					const int q = q0 + kk;
					double a = 0.0, b = 0.0;
					if (live && q < nq) {
						const int t = t0 + q / 3, comp = q - 3 * (q / 3);
						const int rb = v.S_term_rowblk[t], cb = v.S_term_colblk[t];
						a = v.lm_inv[v.HApf_col[rb]] ? -v.W[(size_t)rb * 18 + row * 3 + comp] : 0.0;
						b = v.HApf[(size_t)cb * 18 + row * 3 + comp];   // (for B, `row` is the output column)
					}
					asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
				}
				if (live && col < 6) { v.Sblk[(size_t)blk * 36 + row * 6 + col] = c0; v.Sblk[(size_t)blk * 36 + row * 6 + col + 1] = c1; }
			}
		for (int i = LM_TID; i < n; i += LM_THREADS) {  // (4) reduced gradient: scalar of the pose part
			const int ucol = i / 6, a = i % 6;
			double g = v.grad[i];
			for (int blk = v.rowApf_ptr[ucol]; blk < v.rowApf_ptr[ucol + 1]; blk++) {
				const int k = v.HApf_col[blk];
				if (!v.lm_inv[k]) continue;
				const double *Wr = v.W + (size_t)blk * 18;
				const double *gf = v.grad + n + (size_t)k * 3;
				g -= Wr[a * 3 + 0] * gf[0] + Wr[a * 3 + 1] * gf[1] + Wr[a * 3 + 2] * gf[2];
			}
			v.grad[i] = g;
		}
		return;
	}
#else
	if constexpr (P == 6) {
		const int nSR = h->s_alias_hap ? 0 : h->n_S * 6;
		for (int it = LM_TID; it < nSR + n; it += LM_THREADS) {
			if (it < nSR) {
				const int blk = it / 6, a = it % 6;
				const int hb = v.S_HAp[blk];
				double s[6];
#pragma unroll
				for (int b = 0; b < 6; b++) s[b] = hb >= 0 ? v.HAp[(size_t)hb * 36 + a * 6 + b] : 0.0;
#pragma unroll 2
				for (int t = v.S_term_ptr[blk]; t < v.S_term_ptr[blk + 1]; t++) {
					const int rb = v.S_term_rowblk[t], cb = v.S_term_colblk[t];
					const double *Wr = v.W + (size_t)rb * 18 + a * 3;
					const double *Hc = v.HApf + (size_t)cb * 18;
					const double w0 = Wr[0], w1 = Wr[1], w2 = Wr[2];
					const bool on = v.lm_inv[v.HApf_col[rb]];  // non-invertible landmarks are skipped (schur.h:228): exact -0.0
#pragma unroll
					for (int b = 0; b < 6; b++) {
						const double term = w0 * Hc[b * 3 + 0] + w1 * Hc[b * 3 + 1] + w2 * Hc[b * 3 + 2];
						s[b] -= on ? term : 0.0;
					}
				}
#pragma unroll
				for (int b = 0; b < 6; b++) v.Sblk[(size_t)it * 6 + b] = s[b];
			} else {
				const int i = it - nSR;  // scalar of the pose part
				const int ucol = i / 6, a = i % 6;
				double g = v.grad[i];
				for (int blk = v.rowApf_ptr[ucol]; blk < v.rowApf_ptr[ucol + 1]; blk++) {
					const int k = v.HApf_col[blk];
					if (!v.lm_inv[k]) continue;
					const double *Wr = v.W + (size_t)blk * 18;
					const double *gf = v.grad + n + (size_t)k * 3;
					g -= Wr[a * 3 + 0] * gf[0] + Wr[a * 3 + 1] * gf[1] + Wr[a * 3 + 2] * gf[2];
				}
				v.grad[i] = g;
			}
		}
		return;
	}
#endif
	const int nS = h->s_alias_hap ? 0 : h->n_S * P * P;
	for (int it = LM_TID; it < nS + n; it += LM_THREADS) {
		if (it < nS) {
			const int blk = it / (P * P), e = it % (P * P), a = e / P, b = e % P;
			const int hb = v.S_HAp[blk];
			double s = hb >= 0 ? v.HAp[(size_t)hb * P * P + e] : 0.0;
#pragma unroll 4
			for (int t = v.S_term_ptr[blk]; t < v.S_term_ptr[blk + 1]; t++) {
				const int rb = v.S_term_rowblk[t], cb = v.S_term_colblk[t];
				const double *Wr = v.W + (size_t)rb * P * 3;
				const double *Hc = v.HApf + (size_t)cb * P * 3;
				const double term = Wr[a * 3 + 0] * Hc[b * 3 + 0] + Wr[a * 3 + 1] * Hc[b * 3 + 1] + Wr[a * 3 + 2] * Hc[b * 3 + 2];
				s -= v.lm_inv[v.HApf_col[rb]] ? term : 0.0;  // non-invertible landmarks are skipped (schur.h:228): exact -0.0
			}
			v.Sblk[it] = s;
		} else {
			const int i = it - nS;  // scalar of the pose part
			const int ucol = i / P, a = i % P;
			double g = v.grad[i];
			for (int blk = v.rowApf_ptr[ucol]; blk < v.rowApf_ptr[ucol + 1]; blk++) {
				const int k = v.HApf_col[blk];
				if (!v.lm_inv[k]) continue;
				const double *Wr = v.W + (size_t)blk * P * 3;
				const double *gf = v.grad + n + (size_t)k * 3;
				g -= Wr[a * 3 + 0] * gf[0] + Wr[a * 3 + 1] * gf[1] + Wr[a * 3 + 2] * gf[2];
			}
			v.grad[i] = g;
		}
	}
}

// Dense fill of solver_engine<true,true>::solve (impl/lev-marq_solvers.h:492-519) for the blocked solver and for
// extra_results.hessian: denseH := 0, upper blocks of the reduced system, +lambda on the diagonal, and the reduced
// gradient as an augmented column (forward substitution then comes for free).
template <int P>
__device__ __noinline__ void stage_dense_fill(const View &vs, double lambda) {
	const View v = vs;  // field-wise copy into registers: no aliasing with the stores below
	const DevProblem *h = v.h;
	const int n = h->n_dense, ld = h->ld_dense;
	for (size_t i = LM_TID; i < (size_t)n * ld; i += LM_THREADS) v.dense[i] = 0;
	LM_BARRIER();
	const int nS = h->n_S * P * P;
	for (int it = LM_TID; it < nS + n; it += LM_THREADS) {
		if (it < nS) {
			const int blk = it / (P * P), e = it % (P * P), a = e / P, b = e % P;
			const int row = v.S_row[blk], col = v.S_col[blk];
			double s = v.S[it];
			if (row == col && a == b) s += lambda;
			v.dense[(size_t)(P * row + a) * ld + P * col + b] = s;
		} else {
			const int i = it - nS;
			v.dense[(size_t)i * ld + n] = v.grad[i];
		}
	}
}

// K4: dense Cholesky A = U^T U on the upper triangle + forward substitution (augmented column n) + back substitution
// (Eigen::LLT<Upper>, impl/lev-marq_solvers.h:527-549).  Blocked right-looking, NB = 8:
//   (a) warp 0 factors the 8x8 diagonal block (already updated by earlier block columns) and keeps U11 in shared memory,
//   (b) every thread solves one column of the 8-row panel  U[kb:kb+8, j] = U11^-T A[kb:kb+8, j]  (coalesced row reads),
//       the panel is staged in shared memory,
//   (c) the trailing trapezoid gets the rank-8 update from the staged panel with 4x4 register tiles (one global RMW per
//       element per block column instead of one per column).
// 3 barriers per block column instead of 3 per column.  Returns false on a non-positive pivot (Eigen: "if (x<=0) return k").
#define CH_NB 8
__device__ __noinline__ bool stage_cholesky(const View &vs, Shared &sh, double *panel, int panel_cap) {
	const View v = vs;  // field-wise copy into registers: no aliasing with the stores below
	const int n = v.h->n_dense, ld = v.h->ld_dense;
	double *A = v.dense;
	const int ncols = n + 1;  // including the augmented column
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = LM_THREADS / 32;
	if (threadIdx.x == 0) sh.chol_ok = 1;
	__syncthreads();
	for (int kb = 0; kb < n; kb += CH_NB) {
		const int w = min(CH_NB, n - kb);
		// (a) diagonal block
		if (warp == 0) {
			for (int e = lane; e < 64; e += 32) {
				const int r = e >> 3, c = e & 7;
				sh.U11[e] = (r < w && c < w && c >= r) ? A[(size_t)(kb + r) * ld + kb + c] : 0.0;
			}
			__syncwarp();
			for (int r = 0; r < w; r++) {
				const double d = sh.U11[r * 8 + r];
				if (d <= 0) { if (lane == 0) sh.chol_ok = 0; break; }
				const double sd = sqrt(d);
				__syncwarp();
				if (lane < 8 && lane >= r && lane < w) sh.U11[r * 8 + lane] = (lane == r) ? sd : sh.U11[r * 8 + lane] / sd;
				__syncwarp();
				for (int e = lane; e < 64; e += 32) {
					const int i = e >> 3, c = e & 7;
					if (i > r && i < w && c >= i && c < w) sh.U11[e] -= sh.U11[r * 8 + i] * sh.U11[r * 8 + c];
				}
				__syncwarp();
			}
			if (sh.chol_ok)
				for (int e = lane; e < 64; e += 32) {
					const int r = e >> 3, c = e & 7;
					if (r < w && c < w && c >= r) A[(size_t)(kb + r) * ld + kb + c] = sh.U11[e];
				}
		}
		__syncthreads();
		if (!sh.chol_ok) return false;
		const int j0 = kb + w;           // first trailing column
		const int m = ncols - j0;        // trailing columns incl. the augmented one
		if (m <= 0) break;
		const bool staged = (m <= panel_cap);
		// (b) panel solve: thread per column
		for (int jj = threadIdx.x; jj < m; jj += LM_THREADS) {
			double x[CH_NB];
#pragma unroll
			for (int r = 0; r < CH_NB; r++) {
				if (r < w) {
					double s = A[(size_t)(kb + r) * ld + j0 + jj];
#pragma unroll
					for (int q = 0; q < CH_NB; q++)
						if (q < r) s -= sh.U11[q * 8 + r] * x[q];
					x[r] = s / sh.U11[r * 8 + r];
					A[(size_t)(kb + r) * ld + j0 + jj] = x[r];
					if (staged) panel[r * panel_cap + jj] = x[r];
				} else if (staged) panel[r * panel_cap + jj] = 0.0;
			}
		}
		__syncthreads();
		// (c) trailing update on rows i in [j0, n), columns j in [i, n]  (local indices i' = i - j0, j' = j - j0)
		const int mr = n - j0;  // trailing rows
		if (mr > 0) {
			const int MT = (m + 3) >> 2, MR = (mr + 3) >> 2;
			for (int ti = warp; ti < MR; ti += nw) {
				for (int tj = ti + lane; tj < MT; tj += 32) {
					double acc[4][4];
#pragma unroll
					for (int a = 0; a < 4; a++)
#pragma unroll
						for (int b = 0; b < 4; b++) acc[a][b] = 0;
#pragma unroll
					for (int r = 0; r < CH_NB; r++) {
						double pi[4], pj[4];
#pragma unroll
						for (int a = 0; a < 4; a++) {
							const int ii = 4 * ti + a, jj = 4 * tj + a;
							if (staged) {
								pi[a] = ii < m ? panel[r * panel_cap + ii] : 0.0;
								pj[a] = jj < m ? panel[r * panel_cap + jj] : 0.0;
							} else {
								pi[a] = (r < w && ii < m) ? A[(size_t)(kb + r) * ld + j0 + ii] : 0.0;
								pj[a] = (r < w && jj < m) ? A[(size_t)(kb + r) * ld + j0 + jj] : 0.0;
							}
						}
#pragma unroll
						for (int a = 0; a < 4; a++)
#pragma unroll
							for (int b = 0; b < 4; b++) acc[a][b] += pi[a] * pj[b];
					}
#pragma unroll
					for (int a = 0; a < 4; a++) {
						const int ii = 4 * ti + a;
						if (ii >= mr) continue;
						double *row = A + (size_t)(j0 + ii) * ld + j0;
#pragma unroll
						for (int b = 0; b < 4; b++) {
							const int jj = 4 * tj + b;
							if (jj >= ii && jj < m) row[jj] -= acc[a][b];
						}
					}
				}
			}
		}
		__syncthreads();
	}
	return true;
}

// K4 (small/medium systems): ON-CHIP BLOCK-SPARSE  L D L^T  solve of the reduced system, one scalar element per lane.
// Solves the system that Eigen::LLT<Upper> + solve see after the dense fill (impl/lev-marq_solvers.h:492-549), fed
// DIRECTLY from the reduced blocks: the dense matrix is never materialised, and structural zeros are never touched
// (a relative graph-SLAM area is block-banded: ~5 % of the dense flops and no fill-in in the reference's own ordering).
// The factorisation is the block form of the same elimination (3x3 pivot blocks inverted explicitly by cofactors, no
// square roots): a pivot block is rejected exactly when one of its leading minors is <= 0, i.e. when LLT would meet a
// non-positive scalar pivot.
//
// Why this shape (measured on B200, tools/microbench + in-kernel phase clocks): the elimination is a chain of dependent
// 3x3 products, one link per block row, and a lone warp retires ~1 instruction per 5-6 cycles -- so the cost of a block
// row is the number of instructions (and barriers) on its critical path, not its flops.  One ELEMENT per lane makes every
// 3x3 product 3 DFMA per lane, and the row is organised so that ONE barrier separates consecutive rows:
//   U  every pair (i,j) of tiles of row k:  T(i,j) -= T(k,i)^t inv(D_k) T(k,j)   -- 9 lanes per update, host-made list;
//      the lanes form inv(D_k) T(k,j) themselves (9 extra DFMA) instead of waiting for a separate "W" step,
//   P  meanwhile 9 lanes of the last warp apply row k to the NEXT pivot block and invert it straight away,
//   W  meanwhile the tiles of row k-1 (no longer read by anybody) are overwritten by W = inv(D_{k-1}) T for the backward
//      pass (inverse pivot blocks are triple-buffered so that U, P and W never meet).
// The filled upper pattern, the tile numbering (row-major, diagonal first, right-hand side as tile column NTR) and the
// per-row update lists come from the packer's symbolic factorisation (capi.cpp: symbolic_tiles).  Tiles live in shared
// memory, column-major 3x3 padded to 3x4 so that a column is one 16-byte + one 8-byte load.
// The backward pass  x_j = z_j - sum W(j,.) x  runs column by column (one barrier each), again one element per lane.
// Returns false when a pivot block is not positive definite (Eigen::LLT "if (x<=0) return k", lev-marq_solvers.h:527-534).
// Backward pass of the block-sparse solver as a REGISTER WAVEFRONT carried by one warp: lane l owns rows l, l+32, ...
// (R rows per lane).  Step j: the owner of x_j broadcasts it by shuffle, every lane whose rows hold a tile in column j
// folds it into that row's accumulator (the W tile was fetched while waiting).  No barrier, no shared-memory round trip.
template <int R>
__device__ __forceinline__ void ldl_backward_wavefront(const double *tile, const int4 *rowinfo, const int32_t *tij, double *sX, int NTR, int lane) {
	int r0[R], m[R], nxt_col[R];
	double z[R][3], w[R][9];
#pragma unroll
	for (int s2 = 0; s2 < R; s2++) {
		const int k = lane + 32 * s2;
		r0[s2] = 0; m[s2] = 0; nxt_col[s2] = -1;
		z[s2][0] = z[s2][1] = z[s2][2] = 0.0;
#pragma unroll
		for (int e = 0; e < 9; e++) w[s2][e] = 0.0;
		if (k < NTR) {
			const int4 rk = rowinfo[k];
			const int r1 = rk.y & 0x3fffffff;
			r0[s2] = rk.x; m[s2] = r1 - rk.x - 2;  // off-diagonal tiles without the right-hand side
			const double *zz = tile + (r1 - 1) * 12;
			z[s2][0] = zz[0]; z[s2][1] = zz[1]; z[s2][2] = zz[2];
		}
	}
	auto fetch = [&](int s2) {  // the row's next tile, walking it from the last column down
		if (m[s2] > 0) {
			const int tl = r0[s2] + m[s2];
			nxt_col[s2] = tij[tl] & 0xffff;
			const double *Wt = tile + tl * 12;
#pragma unroll
			for (int b = 0; b < 3; b++)
#pragma unroll
				for (int a = 0; a < 3; a++) w[s2][b * 3 + a] = Wt[b * 4 + a];
			m[s2]--;
		} else nxt_col[s2] = -1;
	};
#pragma unroll
	for (int s2 = 0; s2 < R; s2++) fetch(s2);
#pragma unroll 1
	for (int j = NTR - 1; j >= 0; j--) {
		const int js = j >> 5, jl = j & 31;
		double y0 = z[0][0], y1 = z[0][1], y2 = z[0][2];
#pragma unroll
		for (int s2 = 1; s2 < R; s2++)
			if (js == s2) { y0 = z[s2][0]; y1 = z[s2][1]; y2 = z[s2][2]; }
		const double x0 = __shfl_sync(0xffffffffu, y0, jl), x1 = __shfl_sync(0xffffffffu, y1, jl), x2 = __shfl_sync(0xffffffffu, y2, jl);
#pragma unroll
		for (int s2 = 0; s2 < R; s2++)
			if (nxt_col[s2] == j) {
				z[s2][0] -= w[s2][0] * x0 + w[s2][3] * x1 + w[s2][6] * x2;
				z[s2][1] -= w[s2][1] * x0 + w[s2][4] * x1 + w[s2][7] * x2;
				z[s2][2] -= w[s2][2] * x0 + w[s2][5] * x1 + w[s2][8] * x2;
				fetch(s2);
			}
	}
#pragma unroll
	for (int s2 = 0; s2 < R; s2++) {
		const int k = lane + 32 * s2;
		if (k < NTR) { sX[3 * k] = z[s2][0]; sX[3 * k + 1] = z[s2][1]; sX[3 * k + 2] = z[s2][2]; }
	}
}

template <int P>
__device__ __noinline__ bool stage_ldl_solve_smem(const View &vs, Shared &sh, double *scr, double lambda) {
	static_assert(P % 3 == 0, "tiles are 3x3");
#ifdef SRBA_SOLVER_TIMERS
	const long long c_entry = clock64();
#endif
	const View v = vs;
	const int n = v.h->n_dense, ntiles = v.h->ldl_n_tiles;
	const int NTR = n / 3;
	const bool tiles_in_ws = v.h->ldl_tiles_in_ws != 0;
	double *tile = tiles_in_ws ? v.ldl_tiles : scr;                    // ntiles x 12 (on chip unless the pattern is too large)
	double *sInv = tiles_in_ws ? scr : scr + 12 * (size_t)ntiles;      // 6 x 12 : inverse pivot blocks of the previous, current and next phase (slot = row % 6)
	int *chol_ok = &sh.chol_ok;
#ifdef SRBA_CLUSTER
	// the pivot chain and the backward pass run on CTA 0: its scratch and flag are read by the other CTAs through
	// distributed shared memory; the tiles themselves live in the global workspace (the packer guarantees tiles_in_ws)
	sInv = cg::this_cluster().map_shared_rank(sInv, 0);
	chol_ok = cg::this_cluster().map_shared_rank(chol_ok, 0);
#endif
	double *sPiv = sInv + 72;                  // 2 x 12 : the two pivot blocks being updated + inverted (kept on chip even when the tiles live in L2)
	double *sAcc = sPiv + 24;                  // 3*NTR : backward-pass accumulators
	const int4 *rowinfo = reinterpret_cast<const int4 *>(v.ldl_row);  // {first tile, end tile, first update, end update}
	const int32_t *tij = v.ldl_tile_ij, *upd = v.ldl_upd, *src = v.ldl_src, *orow = v.ldl_orow;
	const int tid = LM_TID, NT = LM_THREADS;   // team-wide thread id / size (one CTA, or the 8 CTAs of a cluster)
	const int lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
	const int iwarp = ((int)blockDim.x >> 5) - 1;  // the warp that carries the pivot chain (its lanes 0..17): last warp of CTA 0
	const int utid = warp < iwarp ? tid : tid - 32; // thread id among the warps that take update items (all but the pivot warp)
	const int wslot = warp < iwarp ? iwarp - 1 - warp : warp - 1;  // same for the W step
	// coordinates of this lane inside a 9-lane tile group: element (wa,wb) at column-major padded offset eo
	const int wg = lane / 9, we = lane - wg * 9, wb = we / 3, wa = we - wb * 3, eo = wb * 4 + wa;
	// cyclic cofactor of a symmetric 3x3: C(a,b) = d(a1,b1) d(a2,b2) - d(a1,b2) d(a2,b1)
	const int a1 = wa == 2 ? 0 : wa + 1, a2 = wa == 0 ? 2 : wa - 1, b1 = wb == 2 ? 0 : wb + 1, b2 = wb == 0 ? 2 : wb - 1;
	const int o11 = b1 * 4 + a1, o22 = b2 * 4 + a2, o12 = b2 * 4 + a1, o21 = b1 * 4 + a2;
	// phase records (4 x int4): {first row, rows, first update, end of row 0's updates | end update, pivots of the next phase,
	// pivot tile 0, pivot tile 1 | link 0, link 1, first tile, first tile of the second row | end tile, -, -, -}
	const int4 *phase = reinterpret_cast<const int4 *>(v.ldl_phase);
	const int nph = v.h->ldl_n_phases;
	int4 pa = phase[0], pb = phase[1], pc = phase[2], pd = phase[3];
	// pivot-block inversion by 9 lanes: own cofactor, determinant, one reciprocal
	auto invert_pivot = [&](const double *D, double *inv) {
		const double d00 = D[0], d01 = D[4], d02 = D[8], d11 = D[5], d12 = D[9], d22 = D[10];
		const double c = D[o11] * D[o22] - D[o12] * D[o21];
		const double det = d00 * (d11 * d22 - d12 * d12) + d01 * (d02 * d12 - d01 * d22) + d02 * (d01 * d12 - d02 * d11);
		inv[eo] = c / det;
		// leading minors == LLT's scalar pivots up to positive factors; garbage after a failure is harmless and is
		// discarded by the single check after the loop
		if (d00 <= 0 || d00 * d11 - d01 * d01 <= 0 || det <= 0) *chol_ok = 0;
	};
	// the pivot blocks of the first phase (one or two rows) do not wait for the gather: their 9-lane groups fetch their own
	// elements and invert them while the other warps gather the remaining tiles
	if (warp == iwarp) {
		if (lane == 0) *chol_ok = 1;
		__syncwarp();
		const bool mine = lane < 9 * pa.y;   // group g handles row g of phase 0
		const int pt = mine ? rowinfo[wg].x : 0;
		if (mine) sPiv[wg * 12 + eo] = v.S[(src[pt] & 0x1fffffff) + wa * P + wb] + (wa == wb ? lambda : 0.0);
		__syncwarp();
		if (mine) invert_pivot(sPiv + wg * 12, sInv + wg * 12);
	}
	// gather: tile element <- reduced block element (+lambda on the diagonal); tile column NTR = reduced gradient.
	// 9 lanes per tile (3 tiles per warp and round); ldl_src: >= 0 offset of the tile's (0,0) in S (bit 30: diagonal tile, bit 29: the tile is the transpose),
	// -1 fill-in, <= -2 right-hand side of row -src-2
	if (lane < 27) {
		const double dl = wa == wb ? lambda : 0.0;
#pragma unroll 4
		for (int tl = 1 + 3 * warp + wg; tl < ntiles; tl += 3 * nwarps) {
			const int so = src[tl];
			double val = 0.0;
			if (so >= 0) val = v.S[(so & 0x1fffffff) + ((so >> 29) & 1 ? wb * P + wa : wa * P + wb)] + ((so >> 30) ? dl : 0.0);
			else if (so <= -2 && wb == 0) val = v.grad[3 * (-so - 2) + wa];
			tile[tl * 12 + eo] = val;
		}
	}
	int2 u_first = make_int2(0, 0);
	const bool wide = v.h->ldl_wide != 0;
	if (!wide && warp != iwarp && utid < (pb.x - pa.z) * 9) u_first = *reinterpret_cast<const int2 *>(upd + 2 * (pa.z + utid / 9));
	LM_BARRIER();
#ifdef SRBA_SOLVER_TIMERS
	const long long c_loop = clock64();
#endif
	int prev_t0 = 0, prev_tm = 0, prev_t1 = 0;  // tiles of the previous phase's rows, still to be turned into W
	int slot0 = 0, prev_slot0 = 0;              // inverse-pivot slot (row % 6) of the first row of the current / previous phase
	const int NTU = NT - 32;                    // the warp that carries the pivot chain takes no update items
#pragma unroll 1
	for (int m = 0; m <= nph; m++) {
		const int nrows = pa.y, u0 = pa.z, umid = pa.w, nitems = (pb.x - pa.z) * 9, nnext = pb.y;
		const int mn = 4 * (m < nph ? m + 1 : nph);
		const int4 na = phase[mn], nb = phase[mn + 1], nc = phase[mn + 2], nd = phase[mn + 3];
		const int slot1 = slot0 + 1 >= 6 ? slot0 - 5 : slot0 + 1;
		const double *Ik0 = sInv + slot0 * 12, *Ik1 = sInv + slot1 * 12;
		int nslot0 = slot0 + nrows;
		if (nslot0 >= 6) nslot0 -= 6;
		if (warp != iwarp) {
			// ---- U: T(i,j) -= T(k,i)^t inv(D_k) T(k,j) for the pairs of tiles of the phase's rows (x: dst | tile(k,i) << 16,
			// y: tile(k,j)).  The first descriptor of a phase is fetched before the previous barrier, and two items are in
			// flight per thread: for large areas the lists and tiles live in L2.
#pragma unroll 2
			for (int q = utid; q < nitems; q += NTU) {
				const int ui = q / 9, e = q - ui * 9, b = e / 3, a = e - b * 3;
				int dsti, tii, tji;
				if (wide) { const int32_t *w3 = upd + 3 * (size_t)(u0 + ui); dsti = w3[0]; tii = w3[1]; tji = w3[2]; }
				else {
					const int2 u = q == utid ? u_first : *reinterpret_cast<const int2 *>(upd + 2 * (u0 + ui));
					dsti = u.x & 0xffff; tii = u.x >> 16; tji = u.y;
				}
				const double *Ik = (u0 + ui < umid) ? Ik0 : Ik1;
				const double *Ti = tile + (size_t)tii * 12 + a * 4;
				const double *Tj = tile + (size_t)tji * 12 + b * 4;
				const double2 p01 = *reinterpret_cast<const double2 *>(Ti), t01 = *reinterpret_cast<const double2 *>(Tj);
				const double p2 = Ti[2], t2 = Tj[2];
				const double2 i0 = *reinterpret_cast<const double2 *>(Ik), i1 = *reinterpret_cast<const double2 *>(Ik + 4);
				const double i02 = Ik[2], i12 = Ik[6], i22 = Ik[10];
				const double w0 = i0.x * t01.x + i0.y * t01.y + i02 * t2;
				const double w1 = i0.y * t01.x + i1.y * t01.y + i12 * t2;
				const double w2 = i02 * t01.x + i12 * t01.y + i22 * t2;
				double *dst = tile + (size_t)dsti * 12 + b * 4 + a;
				*dst = *dst - p01.x * w0 - p01.y * w1 - p2 * w2;
			}
			// first update descriptor of the NEXT phase, in the shadow of this phase's barrier
			u_first = make_int2(0, 0);
			if (!wide && utid < (nb.x - na.z) * 9) u_first = *reinterpret_cast<const int2 *>(upd + 2 * (na.z + utid / 9));
			// ---- W: the off-diagonal tiles of the previous phase's rows become W = inv(D) T in place (3 tiles per warp and round)
			if (m > 0) {
#pragma unroll 1
				for (int base = prev_t0 + 3 * wslot; base < prev_t1; base += 3 * (nwarps - 1)) {
					const int tl = base + wg;
					const bool second = tl >= prev_tm;
					const bool act = lane < 27 && tl < prev_t1 && tl != prev_t0 && tl != prev_tm;  // not the diagonal tiles
					double w = 0.0;
					if (act) {
						int sl = prev_slot0 + (second ? 1 : 0);
						if (sl >= 6) sl -= 6;
						const double *Ip = sInv + sl * 12;
						const double *Tc = tile + tl * 12 + wb * 4;
						const double2 t01 = *reinterpret_cast<const double2 *>(Tc);
						const double2 i01 = *reinterpret_cast<const double2 *>(Ip + wa * 4);
						w = i01.x * t01.x + i01.y * t01.y + Ip[wa * 4 + 2] * Tc[2];
					}
					__syncwarp();
					if (act) tile[tl * 12 + eo] = w;
				}
			}
		} else {
			// ---- P: the pivot blocks of the next phase (one 9-lane group each) get this phase's update through their link
			// tile (bit 30: it belongs to the phase's second row) and are inverted right away
			const bool mine = lane < 9 * nnext;
			const int pt = wg == 0 ? pb.z : pb.w, link = wg == 0 ? pc.x : pc.y;
			const double *Dn = tile + pt * 12;
			double *Dp = sPiv + wg * 12;
			if (mine) {
				double dv = Dn[eo];
				if (link >= 0) {
					const double *Ts = tile + (link & 0x3fffffff) * 12;
					const double *Ik = (link >> 30) ? Ik1 : Ik0;
					const double2 p01 = *reinterpret_cast<const double2 *>(Ts + wa * 4), t01 = *reinterpret_cast<const double2 *>(Ts + wb * 4);
					const double p2 = Ts[wa * 4 + 2], t2 = Ts[wb * 4 + 2];
					const double2 i0 = *reinterpret_cast<const double2 *>(Ik), i1 = *reinterpret_cast<const double2 *>(Ik + 4);
					const double i02 = Ik[2], i12 = Ik[6], i22 = Ik[10];
					const double w0 = i0.x * t01.x + i0.y * t01.y + i02 * t2;
					const double w1 = i0.y * t01.x + i1.y * t01.y + i12 * t2;
					const double w2 = i02 * t01.x + i12 * t01.y + i22 * t2;
					dv = dv - p01.x * w0 - p01.y * w1 - p2 * w2;
				}
				Dp[eo] = dv;
			}
			__syncwarp();
			int ns = nslot0 + wg;
			if (ns >= 6) ns -= 6;
			if (mine) invert_pivot(Dp, sInv + ns * 12);
		}
		prev_t0 = pc.z; prev_tm = nrows > 1 ? pc.w : pd.x; prev_t1 = pd.x;
		prev_slot0 = slot0; slot0 = nslot0;
		pa = na; pb = nb; pc = nc; pd = nd;
		LM_BARRIER();
	}
	if (!*chol_ok) return false;
#ifdef SRBA_SOLVER_TIMERS
	const long long c_back = clock64();
#endif
	// backward pass  x_k = z_k - sum_j W(k,j) x_j  (z_k = column 0 of the last tile of row k): a recurrence over the rows,
	// carried by ONE warp with warp-level synchronisation only (a CTA barrier per row costs ~400 cycles here).
	if (warp == 0) {
		double *sX = sAcc;
		if (NTR <= 32) ldl_backward_wavefront<1>(tile, rowinfo, tij, sX, NTR, lane);
		else if (NTR <= 64) ldl_backward_wavefront<2>(tile, rowinfo, tij, sX, NTR, lane);
		else if (NTR <= 128) ldl_backward_wavefront<4>(tile, rowinfo, tij, sX, NTR, lane);
#ifdef SRBA_CLUSTER  /* 255 registers per thread: room for 8 rows per lane */
		else if (NTR <= 256) ldl_backward_wavefront<8>(tile, rowinfo, tij, sX, NTR, lane);
#endif
		else {
			// general size: lane (a, t) multiplies row a of the t-th tile of the row, a 3-step butterfly sums over the tiles;
			// the next row's record, tile columns and W entries do not depend on x and are fetched one row ahead
			const int a = lane >> 3, t8 = lane & 7;
			int4 rk = rowinfo[NTR - 1];
			auto prefetch = [&](const int4 &r, int &col, double &w0, double &w1, double &w2) {
				const int r0 = r.x, m = (r.y & 0x3fffffff) - r0 - 2;
				col = -1; w0 = w1 = w2 = 0.0;
				if (a < 3 && t8 < m) {
					const int tl = r0 + 1 + t8;
					col = tij[tl] & 0xffff;
					const double *Wt = tile + tl * 12 + a;
					w0 = Wt[0]; w1 = Wt[4]; w2 = Wt[8];
				}
			};
			int col; double w0, w1, w2;
			prefetch(rk, col, w0, w1, w2);
#pragma unroll 1
			for (int k = NTR - 1; k >= 0; k--) {
				const int r0 = rk.x, r1 = rk.y & 0x3fffffff, m = r1 - r0 - 2;
				const double zk = a < 3 ? tile[(r1 - 1) * 12 + a] : 0.0;
				const int4 rnx = rowinfo[k > 0 ? k - 1 : 0];
				int ncol; double n0, n1, n2;
				prefetch(rnx, ncol, n0, n1, n2);
				double p = 0.0;
				if (col >= 0) { const double *x = sX + 3 * col; p = w0 * x[0] + w1 * x[1] + w2 * x[2]; }
#pragma unroll 1
				for (int base = 8; base < m; base += 8) {  // rows with more than 8 off-diagonal tiles
					const int t = base + t8;
					if (a < 3 && t < m) {
						const int tl = r0 + 1 + t;
						const double *x = sX + 3 * (tij[tl] & 0xffff);
						const double *Wt = tile + tl * 12 + a;
						p += Wt[0] * x[0] + Wt[4] * x[1] + Wt[8] * x[2];
					}
				}
				p += __shfl_xor_sync(0xffffffffu, p, 4);
				p += __shfl_xor_sync(0xffffffffu, p, 2);
				p += __shfl_xor_sync(0xffffffffu, p, 1);
				if (a < 3 && t8 == 0) sX[3 * k + a] = zk - p;
				__syncwarp();
				rk = rnx; col = ncol; w0 = n0; w1 = n1; w2 = n2;
			}
		}
	}
	LM_BARRIER();
#pragma unroll 1
	for (int i = tid; i < n; i += NT) { const int r = i / 3; v.delta[3 * orow[r] + (i - 3 * r)] = sAcc[i]; }  // back to the reference's unknown order
	LM_BARRIER();
#ifdef SRBA_SOLVER_TIMERS
	if (tid == 0) {  // team thread 0 == thread 0 of CTA 0
		const long long c_end = clock64();
		sh.dbg[0] += c_loop - c_entry; sh.dbg[1] += c_back - c_loop; sh.dbg[2] += c_end - c_back;
	}
#endif
	return true;
}

// Back substitution U x = y (y = augmented column), blocked by 8 from the bottom: warp 0 solves the diagonal block,
// then every thread updates one earlier row with the 8 new unknowns.
__device__ __noinline__ void stage_backsub(const View &vs, Shared &sh) {
	const View v = vs;  // field-wise copy into registers: no aliasing with the stores below
	const int n = v.h->n_dense, ld = v.h->ld_dense;
	const double *A = v.dense;
	double *x = v.delta;
	for (int i = threadIdx.x; i < n; i += LM_THREADS) x[i] = A[(size_t)i * ld + n];
	__syncthreads();
	const int nb = (n + CH_NB - 1) / CH_NB;
	for (int b = nb - 1; b >= 0; b--) {
		const int kb = b * CH_NB, w = min(CH_NB, n - kb);
		if (threadIdx.x < 32) {
			for (int e = threadIdx.x; e < 64; e += 32) {
				const int r = e >> 3, c = e & 7;
				sh.U11[e] = (r < w && c < w && c >= r) ? A[(size_t)(kb + r) * ld + kb + c] : 0.0;
			}
			if (threadIdx.x < 8) sh.xblk[threadIdx.x] = threadIdx.x < w ? x[kb + threadIdx.x] : 0.0;
			__syncwarp();
			if (threadIdx.x == 0) {
				for (int r = w - 1; r >= 0; r--) {
					double s = sh.xblk[r];
					for (int c = r + 1; c < w; c++) s -= sh.U11[r * 8 + c] * sh.xblk[c];
					sh.xblk[r] = s / sh.U11[r * 8 + r];
				}
			}
			__syncwarp();
			if (threadIdx.x < w) x[kb + threadIdx.x] = sh.xblk[threadIdx.x];
		}
		__syncthreads();
		for (int i = threadIdx.x; i < kb; i += LM_THREADS) {
			const double *row = A + (size_t)i * ld + kb;
			double s = 0;
#pragma unroll
			for (int c = 0; c < CH_NB; c++)
				if (c < w) s += row[c] * sh.xblk[c];
			x[i] -= s;
		}
		__syncthreads();
	}
}

// K5a: SchurComplement::numeric_solve_for_features (impl/schur.h:271-311): mutates the gradient tail in place.
template <int P>
LM_STAGE void stage_features(const View &vs) {
	const View v = vs;  // field-wise copy into registers: no aliasing with the stores below
	const DevProblem *h = v.h;
	const int n = h->n_dense;
	for (int k = LM_TID; k < h->n_unk_lms; k += LM_THREADS) {
		double *d = v.delta + n + (size_t)k * 3;
		if (!v.lm_inv[k]) { d[0] = 0; d[1] = 0; d[2] = 0; continue; }  // delta_eps.tail().setZero()
		double *gf = v.grad + n + (size_t)k * 3;
		double g0 = gf[0], g1 = gf[1], g2 = gf[2];
		for (int q = v.colApf_ptr[k]; q < v.colApf_ptr[k + 1]; q++) {
			const int blk = v.colApf_idx[q];
			const double *Hb = v.HApf + (size_t)blk * P * 3;
			const double *dp = v.delta + (size_t)v.HApf_row[blk] * P;
			double s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
			for (int a = 0; a < P; a++) { s0 += Hb[a * 3 + 0] * dp[a]; s1 += Hb[a * 3 + 1] * dp[a]; s2 += Hb[a * 3 + 2] * dp[a]; }
			g0 -= s0; g1 -= s1; g2 -= s2;
		}
		gf[0] = g0; gf[1] = g1; gf[2] = g2;
		const double *inv = v.Hf_inv + (size_t)k * 9;
		d[0] = inv[0] * g0 + inv[1] * g1 + inv[2] * g2;
		d[1] = inv[3] * g0 + inv[4] * g1 + inv[5] * g2;
		d[2] = inv[6] * g0 + inv[7] * g1 + inv[8] * g2;
	}
}

// K5b: back-up + manifold update (impl/optimize_edges.h:491-557)
template <int P>
LM_STAGE void stage_update(const View &vs, int n_listed) {
	const View v = vs;  // field-wise copy into registers: no aliasing with the stores below
	constexpr int S = PoseOps<P>::S;
	const DevProblem *h = v.h;
	for (int i = LM_TID; i < h->n_unk_edges; i += LM_THREADS) {
		double *ep = v.edge_cur + (size_t)v.unk_edge[i] * S;
		double oldp[S], incr[S], np[S];
#pragma unroll
		for (int k = 0; k < S; k++) { oldp[k] = ep[k]; v.edge_bak[(size_t)i * S + k] = oldp[k]; }
		PoseOps<P>::pseudo_exp(incr, v.delta + (size_t)i * P);
		PoseOps<P>::compose(np, incr, oldp);
#pragma unroll
		for (int k = 0; k < S; k++) ep[k] = np[k];
	}
	for (int i = LM_TID; i < h->n_unk_lms; i += LM_THREADS) {
		double *lp = v.lm_cur + (size_t)v.unk_lm[i] * 3;
		const double *d = v.delta + h->n_dense + (size_t)i * 3;
#pragma unroll
		for (int k = 0; k < 3; k++) { v.lm_bak[(size_t)i * 3 + k] = lp[k]; lp[k] += d[k]; }
	}
	// only the LISTED directed entries are ever restored (optimize_edges.h:664-670): back up just those, compactly
	for (int q = LM_TID; q < n_listed * S; q += LM_THREADS) v.slot_bak[q] = v.slot[(size_t)v.listed_idx[q / S] * S + (q % S)];
}
// rejected step (impl/optimize_edges.h:660-680): only the LISTED directed entries of the numeric ST are restored.
template <int P>
LM_STAGE void stage_restore(const View &vs, int n_listed) {
	const View v = vs;  // field-wise copy into registers: no aliasing with the stores below
	constexpr int S = PoseOps<P>::S;
	const DevProblem *h = v.h;
	for (int i = LM_TID; i < h->n_unk_edges * S; i += LM_THREADS) v.edge_cur[(size_t)v.unk_edge[i / S] * S + (i % S)] = v.edge_bak[i];
	for (int i = LM_TID; i < h->n_unk_lms * 3; i += LM_THREADS) v.lm_cur[(size_t)v.unk_lm[i / 3] * 3 + (i % 3)] = v.lm_bak[i];
	for (int q = LM_TID; q < n_listed * S; q += LM_THREADS) v.slot[(size_t)v.listed_idx[q / S] * S + (q % S)] = v.slot_bak[q];
}

// ------------------------------------------------------------------------------------------------ the LM program
// Per-stage cycle counters (tools/stage_breakdown.py) are compiled in only with -DSRBA_SOLVER_TIMERS: each probe costs
// the lone active warp ~250 cycles.
#ifdef SRBA_SOLVER_TIMERS
#define STAGE(id)                                            \
	do {                                                     \
		if (LM_TID == 0) {                              \
			const unsigned long long t_ = clock64();         \
			sh.stage_cycles[id] += t_ - sh.t_last;           \
			sh.t_last = t_;                                  \
		}                                                    \
	} while (0)
#else
#define STAGE(id) do { } while (0)
#endif

template <int P, int OBSK, int TPT>
__device__ void lm_program(const DevProblem *gh, const char *img, char *ws, char *res, Shared &sh, char *dyn) {
	constexpr int S = PoseOps<P>::S;
	if (threadIdx.x == 0) {
		for (int i = 0; i < ST_COUNT; i++) sh.stage_cycles[i] = 0;
		for (int i = 0; i < 8; i++) sh.dbg[i] = 0;
		sh.t_last = clock64();
	}
	// ---- problem header on chip, then the hot read-only index arrays the packer placed in the arena
	{
		const uint32_t *src = reinterpret_cast<const uint32_t *>(gh);
		uint32_t *dst = reinterpret_cast<uint32_t *>(&sh.hdr);
		for (int i = threadIdx.x; i < (int)(sizeof(DevProblem) / 4); i += blockDim.x) dst[i] = src[i];  // every CTA of the team keeps its own copy
	}
	__syncthreads();
	const DevProblem *h = &sh.hdr;
	if (threadIdx.x == 0) make_view(sh.view, h, img, ws, res, dyn);
	__syncthreads();
	const View &v = sh.view;
	double *scratch = reinterpret_cast<double *>(dyn + h->sm_scratch);
#if SRBA_TMA_STAGING
	// Staging of the read-only index arrays the packer placed in the arena: ONE thread issues a bulk asynchronous copy (TMA, cp.async.bulk
	// global -> shared, SASS UBLKCP) per array, all completing on one mbarrier whose transaction count is the total byte count; the other
	// threads go straight to the barrier wait instead of executing a 16-byte load/store loop.
	{
		__shared__ __align__(8) unsigned long long stage_bar;
		const unsigned bar = (unsigned)__cvta_generic_to_shared(&stage_bar);
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
			unsigned total = 0;
			for (int a = 0; a < SRBA_I_COUNT; a++)
				if (h->img_sm[a] != SRBA_NOT_IN_SMEM) total += (unsigned)(((h->img_len[a] + 15) >> 4) << 4);
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
			for (int a = 0; a < SRBA_I_COUNT; a++) {
				if (h->img_sm[a] == SRBA_NOT_IN_SMEM) continue;
				const unsigned bytes = (unsigned)(((h->img_len[a] + 15) >> 4) << 4);
				if (!bytes) continue;
				const unsigned dst = (unsigned)__cvta_generic_to_shared(dyn + h->img_sm[a]);
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(img + h->img_off[a]), "r"(bytes), "r"(bar)
				             : "memory");
			}
		}
		__syncthreads();
		unsigned done = 0;
		while (!done)
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
	}
#else
	for (int a = 0; a < SRBA_I_COUNT; a++) {
		if (h->img_sm[a] == SRBA_NOT_IN_SMEM) continue;
		const uint4 *src = reinterpret_cast<const uint4 *>(img + h->img_off[a]);
		uint4 *dst = reinterpret_cast<uint4 *>(dyn + h->img_sm[a]);
		const int n16 = (int)((h->img_len[a] + 15) >> 4);
		for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
	}
#endif
	const int n = h->n_dense, nsc = h->n_scalars;
	const int nObsRows = h->n_total_obs_rows;

	// ---- load inputs into the workspace (edges and landmarks are the only mutable inputs)
	{
		const double *ge = reinterpret_cast<const double *>(img + h->img_off[SRBA_I_edge_pose]);
		const double *gl = reinterpret_cast<const double *>(img + h->img_off[SRBA_I_lm_pos]);
		for (size_t i = LM_TID; i < (size_t)h->n_edges * S; i += LM_THREADS) v.edge_cur[i] = ge[i];
		for (size_t i = LM_TID; i < (size_t)h->n_lms * 3; i += LM_THREADS) v.lm_cur[i] = gl[i];
	}
	for (int i = LM_TID; i < 2 * h->n_pairs; i += LM_THREADS) v.slot_listed[i] = 0;
	if (threadIdx.x == 0) {
		if (h->model.sensor_pose_kind == SRBA_SENSORPOSE_SE3) PoseOps<6>::inverse(sh.sensor_inv, h->model.sensor_pose);
		sh.status = SRBA_OK;
	}
	LM_BARRIER();
	// E. update_numeric(set,false)  (optimize_edges.h:256)
	stage_chain<P>(v, -1);
	LM_BARRIER();
	// F/G. validity reset + recompute_all_Jacobians (:269-276)
	stage_jacobians<P, OBSK>(v, true);
	LM_BARRIER();
	// (in a cluster every CTA builds the same two lists: identical values to the same addresses, and its own counts)
	if (threadIdx.x < 32) {  // compact list of the listed directed entries (ascending), built once by one warp
		int cnt = 0;
		for (int base = 0; base < 2 * h->n_pairs; base += 32) {
			const int sidx = base + threadIdx.x;
			const bool on = sidx < 2 * h->n_pairs && v.slot_listed[sidx];
			const unsigned m = __ballot_sync(0xffffffffu, on);
			if (on) v.listed_idx[cnt + __popc(m & ((1u << threadIdx.x) - 1))] = sidx;
			cnt += __popc(m);
		}
		if (threadIdx.x == 0) sh.n_listed = cnt;
	} else if (threadIdx.x < 64) {  // compact list of the pairs holding a listed pose, by the second warp
		const int l = threadIdx.x - 32;
		int cnt = 0;
		for (int base = 0; base < h->n_pairs; base += 32) {
			const int t = base + l;
			const bool on = t < h->n_pairs && (v.slot_listed[2 * t] | v.slot_listed[2 * t + 1]);
			const unsigned m = __ballot_sync(0xffffffffu, on);
			if (on) v.refresh_idx[cnt + __popc(m & ((1u << l) - 1))] = t;
			cnt += __popc(m);
		}
		if (l == 0) sh.n_refresh = cnt;
	}
	// J. Hessian (:329-331)
	stage_hessian<P, OBSK>(v);
	LM_BARRIER();
	// L. lambda0 = tau * max diag (:365-390)
	double dmax = 0;
	for (int b = LM_TID; b < h->n_HAp; b += LM_THREADS)
		if (v.HAp_row[b] == v.HAp_col[b]) {
			double m = v.HAp[(size_t)b * P * P];
			for (int k = 1; k < P; k++) m = fmax(m, v.HAp[(size_t)b * P * P + k * P + k]);
			dmax = fmax(dmax, m);
		}
	for (int k = LM_TID; k < h->n_unk_lms; k += LM_THREADS)
		dmax = fmax(dmax, fmax(v.Hf[(size_t)k * 9], fmax(v.Hf[(size_t)k * 9 + 4], v.Hf[(size_t)k * 9 + 8])));
	dmax = team_max(dmax, sh, 1);
	// M. residuals (:395-404)
	double unused_sum;
	const double err0 = team_sum2(stage_residuals<P, OBSK>(v, v.resid[0], sh.sensor_inv), 0.0, unused_sum, sh, 0);
	// N. gradient (:422-426)
	stage_gradient<P, OBSK>(v, v.resid[0]);
	if (threadIdx.x == 0) {
		sh.nu = 2;
		sh.lambda = 1e-3 * dmax;
		sh.err = err0;
		sh.rmse = sqrt(err0 / nObsRows);
		sh.iter = 0; sh.stop = 0; sh.trials = 0; sh.accepted = 0; sh.relin = 0; sh.cholfail = 0; sh.cur = 0;
		sh.need_outer = 1; sh.rho = 0;
		v.stats->total_sqr_error_init = err0;
		if (!isfinite(err0) || !isfinite(dmax)) sh.status = SRBA_ERR_NON_FINITE;  // non-finite inputs: reported, the loop below still mirrors the reference
	}
	const double MAX_LAMBDA = h->model.max_lambda;
	const bool keep_grad = (OBSK != SRBA_OBS_RELPOSES2D) && no_schur_solver(h);  // relative graph-SLAM has no landmark unknowns: compiled out
	const double max_gradient_to_stop = 1e-15;
	// P. LM iterations (:454-696).  The reference's two nested loops
	//     while (iter < max_iters && !stop) { rho = 0; stop tests; while (rho <= 0 && !stop) { TRIAL } iter++; }
	// are flattened: after every trial thread 0 walks the scalar control flow up to the next trial (or the exit) and
	// publishes the decision in sh.go, so a trial costs one control barrier instead of one per loop level.
	auto advance = [&]() {
		for (;;) {
			if (sh.need_outer) {
				if (!(sh.iter < h->model.max_iters && !sh.stop)) { sh.go = 0; return; }
				sh.rho = 0;
				if (sh.lambda >= MAX_LAMBDA) sh.stop = 1;
				if (sh.rmse < h->model.max_error_per_obs_to_stop) sh.stop = 1;
				sh.need_outer = 0;
			}
			if (sh.rho <= 0 && !sh.stop) { sh.go = 1; return; }
			sh.iter++;
			sh.need_outer = 1;
		}
	};
	if (threadIdx.x == 0) advance();
	LM_BARRIER();
	STAGE(ST_INIT);

	{
		while (sh.go) {
			const double lambda = sh.lambda;
			const int trial = sh.trials;
			bool lm_singular = false;
			if (!(h->n_unk_lms == 0 && h->s_alias_hap)) {  // relative graph-SLAM: the reduced system IS H_Ap
				if (keep_grad)  // (read again only after the two barriers inside stage_schur)
					for (int i = LM_TID; i < nsc; i += LM_THREADS) v.grad[i] = v.grad0[i];
				stage_schur<P>(v, lambda);
				LM_BARRIER();
				if (keep_grad) {  // a landmark block that cannot be eliminated = the whole-system Cholesky fails (:171-175)
					double bad = 0;
					for (int k = LM_TID; k < h->n_unk_lms; k += LM_THREADS) bad = fmax(bad, v.lm_inv[k] ? 0.0 : 1.0);
					lm_singular = team_max(bad, sh, 2) > 0;
				}
			}
			const bool reg = (h->chol_mode == SRBA_CHOL_REG);
			if (!reg || h->want_dense) {
				stage_dense_fill<P>(v, lambda);
				LM_BARRIER();
				if (h->want_dense) {  // extra_results.hessian = denseH of the last solve (lev-marq_solvers.h:582-586)
					for (size_t i = LM_TID; i < (size_t)n * n; i += LM_THREADS) v.r_dense[i] = v.dense[(i / n) * h->ld_dense + (i % n)];
					LM_BARRIER();
				}
			}
			STAGE(ST_SCHUR);
			// dense solve: register-resident pipelined tiles when the system fits the CTA's register file, blocked global/shared otherwise
			bool ok;
			if (reg) ok = stage_ldl_solve_smem<P>(v, sh, scratch, lambda);
			else ok = stage_cholesky(v, sh, scratch, h->panel_cap);
			STAGE(ST_CHOL);
			if (!ok || lm_singular) {  // not positive definite (:476-485)
				if (threadIdx.x == 0) {
					if (trial < h->trace_cap) {
						double *tr = v.r_trace + (size_t)trial * 6;
						tr[0] = lambda; tr[1] = 0; tr[2] = 0; tr[3] = 0; tr[4] = 0; tr[5] = 0;
					}
					sh.trials++; sh.cholfail++;
					sh.lambda *= sh.nu;
					sh.nu *= 2.;
					sh.stop = (sh.lambda > MAX_LAMBDA);
					advance();
				}
				if (trial < h->trace_cap)
					for (int i = LM_TID; i < nsc; i += LM_THREADS) v.r_trace_delta[(size_t)trial * nsc + i] = 0;
				LM_BARRIER();
				STAGE(ST_CTRL);
				continue;
			}
			if (!reg) stage_backsub(v, sh);
			STAGE(ST_BACKSUB);
			if (h->n_unk_lms > 0) {
				stage_features<P>(v);
				LM_BARRIER();
			}
			STAGE(ST_FEATURES);
			stage_update<P>(v, sh.n_listed);
			LM_BARRIER();
			STAGE(ST_UPDATE);
			stage_chain<P>(v, sh.n_refresh);  // update_numeric(set, true) (:565)
			LM_BARRIER();
			STAGE(ST_CHAIN);
			double *new_resid = v.resid[sh.cur ^ 1];
			// rho (:585): its denominator rides along in the residual reduction
			double dl = 0;
			const double *gr = keep_grad ? v.grad0 : v.grad;  // the Schur variants see their own in-place reduced gradient here
			for (int i = LM_TID; i < nsc; i += LM_THREADS) dl += v.delta[i] * (lambda * v.delta[i] + gr[i]);
			double denom;  // `dl` rides along in the same reduction (the denominator of rho, optimize_edges.h:585)
			const double new_err = team_sum2(stage_residuals<P, OBSK>(v, new_resid, sh.sensor_inv), dl, denom, sh, 0);
			STAGE(ST_RESID);
			const double err = sh.err;
			const double rho = (err - new_err) / denom;
			const double error_reduction_ratio = err > 0 ? (err - new_err) / err : 0;
			const bool accept = rho > 0;
			const bool do_relinearize = accept && (error_reduction_ratio < 0 || error_reduction_ratio > h->model.min_error_reduction_ratio_to_relinearize);
			if (trial < h->trace_cap) {
				if (threadIdx.x == 0) {
					double *tr = v.r_trace + (size_t)trial * 6;
					tr[0] = lambda; tr[1] = 1; tr[2] = rho; tr[3] = new_err; tr[4] = accept ? 1 : 0; tr[5] = do_relinearize ? 1 : 0;
				}
				for (int i = LM_TID; i < nsc; i += LM_THREADS) v.r_trace_delta[(size_t)trial * nsc + i] = v.delta[i];
			}
			STAGE(ST_RHO);
			if (accept) {
				if (do_relinearize) {  // :606-629
					stage_jacobians<P, OBSK>(v, false);
					LM_BARRIER();
					STAGE(ST_JACOB);
					stage_hessian<P, OBSK>(v);
					LM_BARRIER();
					STAGE(ST_HESS);
				}
				double gm = stage_gradient<P, OBSK>(v, new_resid);  // :633
				gm = team_max(gm, sh, 1);
				STAGE(ST_GRAD);
				if (threadIdx.x == 0) {
					sh.cur ^= 1;
					sh.err = new_err;
					sh.rmse = sqrt(new_err / nObsRows);
					sh.trials++; sh.accepted++; sh.relin += do_relinearize ? 1 : 0;
					if (gm <= max_gradient_to_stop) sh.stop = 1;
					if (sh.rmse < h->model.max_error_per_obs_to_stop) sh.stop = 1;
					if (rho > h->model.max_rho) sh.stop = 1;
					sh.lambda *= 1.0 / 3.0;
					sh.nu = 2.0;
					sh.rho = rho;
					advance();
				}
			} else {
				stage_restore<P>(v, sh.n_listed);
				if (threadIdx.x == 0) {
					sh.trials++;
					sh.lambda *= sh.nu;
					sh.nu *= 2.0;
					sh.stop = (sh.lambda > MAX_LAMBDA);
					sh.rho = rho;  // <=0 or NaN: NaN also leaves the while in the reference ("rho<=0" false)
					if (!isfinite(new_err) || !isfinite(rho)) sh.status = SRBA_ERR_NON_FINITE;  // a non-finite iterate
					advance();
				}
			}
			LM_BARRIER();
			STAGE(ST_CTRL);
		}
	}
	LM_BARRIER();
	// Q/S. outputs
	for (size_t i = LM_TID; i < (size_t)h->n_unk_edges * S; i += LM_THREADS) v.r_edge_pose[i] = v.edge_cur[(size_t)v.unk_edge[i / S] * S + (i % S)];
	for (int i = LM_TID; i < h->n_unk_lms * 3; i += LM_THREADS) v.r_lm_pos[i] = v.lm_cur[(size_t)v.unk_lm[i / 3] * 3 + (i % 3)];
	for (size_t i = LM_TID; i < (size_t)2 * h->n_pairs * S; i += LM_THREADS) v.r_slot_pose[i] = v.slot[i];
	for (int i = LM_TID; i < h->n_obs; i += LM_THREADS) v.r_obs_valid[i] = v.valid[i];
	for (int i = LM_TID; i < h->n_unk_lms; i += LM_THREADS) v.r_lm_inv[i] = keep_grad ? 1 : (sh.trials > 0 ? v.lm_inv[i] : 0);  // was_ith_feature_invertible (:196-200,386-389,587-590)
	for (int i = LM_TID; i < h->n_unk_lms * 9; i += LM_THREADS) v.r_Hf_diag[i] = v.Hf[i];
	if (threadIdx.x == 0) {
		srba_stats_t *st = v.stats;
		st->num_iters = sh.iter;
		st->num_trials = sh.trials;
		st->num_accepted = sh.accepted;
		st->num_relinearized = sh.relin;
		st->num_chol_failures = sh.cholfail;
		st->num_jacobians = h->n_blkA + ((OBSK == SRBA_OBS_RELPOSES2D) ? 0 : h->n_blkF);
		st->hessian_valid = sh.trials > 0;
		st->status = sh.status;
		st->total_sqr_error_final = sh.err;
		st->obs_rmse = sh.rmse;
		st->lambda_final = sh.lambda;
		unsigned long long *sc = reinterpret_cast<unsigned long long *>(res + h->res_stage_cycles);
		const unsigned long long t_ = clock64();
		sh.stage_cycles[ST_OUT] += t_ - sh.t_last;
		if (LM_RANK == 0) {  // in a cluster every CTA keeps the same stage clocks; CTA 0 (which carries the pivot chain) reports
			for (int i = 0; i < ST_COUNT; i++) sc[i] = sh.stage_cycles[i];
			for (int i = 0; i < 8; i++) sc[16 + i] = sh.dbg[i];
		}
	}
}

// One kernel per model combination so that each gets its own register budget (the SE(2) graph-SLAM program needs
// far fewer registers than the SE(3) camera models and keeps 2 CTAs per SM resident).
template <int P, int OBSK, int NT, int MINB, int TPT>
__global__ void __launch_bounds__(NT, MINB) srba_lm_batch_kernel(const char *img, const uint64_t *img_off, char *ws, const uint64_t *ws_off, char *res,
                                                                          const uint64_t *res_off, const int32_t *order, int n_problems) {
	__shared__ Shared sh;
	extern __shared__ __align__(16) char dyn_smem[];
	for (int q = blockIdx.x / LM_NCTA; q < n_problems; q += gridDim.x / LM_NCTA) {
		const int pi = order[q];
		const char *pimg = img + img_off[pi];
		const DevProblem *h = reinterpret_cast<const DevProblem *>(pimg);
		lm_program<P, OBSK, TPT>(h, pimg, ws + ws_off[pi], res + res_off[pi], sh, dyn_smem);
		LM_BARRIER();
	}
}

}  // namespace

#if defined(SRBA_WAVE)
#include "lm_wave.cuh"
#elif !defined(SRBA_CLUSTER)
// Per-model kernel configuration: CTA size, register tiles per thread, dynamic shared memory per CTA.
//   SE2+RelativePoses2D: 168 regs x 192 threads -> 2 CTAs/SM, 2 x (108 KB dynamic + ~5 KB static + 1 KB reserved) of the SM's
//   228 KB;  SE3 models: 1024 threads x 64 regs -> 1 CTA/SM, 220 KB.
//   Models 4..7 are the CLUSTER variants of 0..3 (lm_kernel_cluster.cu): 8 CTAs x 256 threads per problem, every mutable
//   array in the global workspace, only the solver's pivot scratch on chip.
#define SE2_NT 192
#ifndef SE3_NT   /* compile-time experiment hook: threads per CTA of the SE(3) kernels (registers per thread = 65536 / SE3_NT) */
#define SE3_NT 1024
#endif
#define SE2_TPT 2
#define SE3_TPT 4
extern "C" SrbaKernelCfg srba_kernel_cfg(int model) {
	SrbaKernelCfg c;
	if (model >= 4) { c.threads = 256; c.tpt = 0; c.smem_bytes = 160 * 1024; return c; }
	c.threads = model == 0 ? SE2_NT : SE3_NT;
	c.tpt = model == 0 ? SE2_TPT : SE3_TPT;
	c.smem_bytes = model == 0 ? 108 * 1024 : 220 * 1024;
	return c;
}
// Doubles of shared-memory scratch of the block-sparse solver with `n_tiles_on_chip` 3x3 tiles kept in shared memory (0:
// the tiles live in the L2-resident workspace), or -1 when even that does not fit.
extern "C" int srba_chol_reg_scratch_doubles(int model, int n_dense, int n_tiles_on_chip) {
	const SrbaKernelCfg c = srba_kernel_cfg(model);
	const int NTR = n_dense / 3;
	const long need = 12L * n_tiles_on_chip + 96 + 3L * NTR + 1;
	if (need * 8 > (long)c.smem_bytes * 7 / 8) return -1;
	return (int)need;
}

extern "C" int srba_launch_lm_batch_cluster(int model, const char *img, const uint64_t *img_off, char *ws, const uint64_t *ws_off, char *res,
                                            const uint64_t *res_off, const int32_t *order, int n_problems, int smem_bytes, void *stream);

// model: 0 = SE2+RelativePoses2D, 1 = SE3+Cartesian3D, 2 = SE3+Mono, 3 = SE3+Stereo; 4..7 = the same on 8-CTA clusters.
// `order` lists the problems of that group.
extern "C" int srba_launch_lm_batch(int model, const char *img, const uint64_t *img_off, char *ws, const uint64_t *ws_off, char *res,
                                    const uint64_t *res_off, const int32_t *order, int n_problems, int grid, int smem_bytes, void *stream) {
	if (model >= 4) return srba_launch_lm_batch_cluster(model - 4, img, img_off, ws, ws_off, res, res_off, order, n_problems, smem_bytes, stream);
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	// dynamic shared memory = the largest arena any problem of this group was planned with (<= the model's budget);
	// what the batch does not need stays available to L1
	const size_t smem = (size_t)std::min(std::max(smem_bytes, 1024), srba_kernel_cfg(model).smem_bytes);
#define LAUNCH(P, OBSK, NT, MINB, TPT)                                                                                                    \
	do {                                                                                                                                  \
		auto kern = srba_lm_batch_kernel<P, OBSK, NT, MINB, TPT>;                                                                         \
		cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                                               \
		kern<<<grid, NT, smem, st>>>(img, img_off, ws, ws_off, res, res_off, order, n_problems);                                          \
	} while (0)
	switch (model) {
		case 0: LAUNCH(3, SRBA_OBS_RELPOSES2D, SE2_NT, 2, SE2_TPT); break;
#ifndef SRBA_ONLY_SE2  /* compile-time experiments on the graph-SLAM kernel only */
		case 1: LAUNCH(6, SRBA_OBS_CARTESIAN3D, SE3_NT, 1, SE3_TPT); break;
		case 2: LAUNCH(6, SRBA_OBS_MONO, SE3_NT, 1, SE3_TPT); break;
		case 3: LAUNCH(6, SRBA_OBS_STEREO, SE3_NT, 1, SE3_TPT); break;
#endif
		default: return (int)cudaErrorInvalidValue;
	}
#undef LAUNCH
	return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ global diagnostics
// RbaEngine::eval_overall_squared_error (impl/eval_overall_error.h:15-137).  Streaming kernels over the whole map (every
// observation once, every needed pair once): HBM-bound by construction, grid = a multiple of the SM count.
struct EvalArgs {
	srba_model_t model;
	int n_pairs, n_obs;
	const double *edge_pose;
	const int32_t *pair_path_ptr, *path_edge;
	const uint8_t *path_dir;
	const double *obs_z;
	const int32_t *obs_pair;
	const uint8_t *obs_inverse;
	const int32_t *obs_lm;
	const double *lm_pos;
	double *pair_pose;   // n_pairs x S (device)
	double *partial;     // one partial sum per CTA of the residual kernel
	double *total;
};

// numeric half of create_complete_spanning_tree (impl/spantree_create_complete.h:91-125): one thread per target
template <int P>
__global__ void __launch_bounds__(256) srba_eval_chain_kernel(const EvalArgs a) {
	constexpr int S = PoseOps<P>::S;
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < a.n_pairs; t += gridDim.x * blockDim.x) {
		double accum[S], tmp[S];
		PoseOps<P>::identity(accum);  // span_tree[root] = pose_t()  (:108)
		for (int k = a.pair_path_ptr[t]; k < a.pair_path_ptr[t + 1]; k++) {
			const double *ep = a.edge_pose + (size_t)a.path_edge[k] * S;
			const double *rhs = ep;
			if (!a.path_dir[k]) { PoseOps<P>::inverse(tmp, ep); rhs = tmp; }  // edge.to == kf: parent (+) (-inv_pose)  (:117-118)
			PoseOps<P>::compose(accum, accum, rhs);
		}
#pragma unroll
		for (int i = 0; i < S; i++) a.pair_pose[(size_t)t * S + i] = accum[i];
	}
}

// sensor_model<>::observe_error on one observation (models/sensors.h:49-66,174-211,348-364,769-785), no robust kernel
template <int P, int OBSK>
__device__ __forceinline__ double eval_obs_sqerr(const EvalArgs &a, const double *sensor_inv, int i) {
	constexpr int S = PoseOps<P>::S;
	constexpr int O = ObsDims<OBSK>::O;
	const double *z = a.obs_z + (size_t)i * O;
	double Pp[S];
	const int t = a.obs_pair[i];
	if (t < 0) PoseOps<P>::identity(Pp);  // base == observer: aux_null_pose (:98-101)
	else {
#pragma unroll
		for (int k = 0; k < S; k++) Pp[k] = a.pair_pose[(size_t)t * S + k];
		if (a.obs_inverse[i]) PoseOps<P>::inverse(Pp, Pp);  // all_rel_poses[*it_want][root] = -pose  (:75)
	}
	double r[O];
	if constexpr (P == 3) {
		double sn, cs;
		sincos(Pp[2], &sn, &cs);
		r[0] = (z[0] - Pp[0]) * cs + (z[1] - Pp[1]) * sn;
		r[1] = -(z[0] - Pp[0]) * sn + (z[1] - Pp[1]) * cs;
		r[2] = wrapToPi(z[2] - Pp[2]);
	} else {
		if (a.model.sensor_pose_kind == SRBA_SENSORPOSE_SE3) PoseOps<6>::compose(Pp, sensor_inv, Pp);
		double x[3];
		compose_point3(Pp, a.lm_pos + (size_t)a.obs_lm[i] * 3, x);
		double pred[O];
		if (OBSK == SRBA_OBS_CARTESIAN3D) { pred[0] = x[0]; pred[1] = x[1]; pred[2] = x[2]; }
		else {
			pred[0] = a.model.cam_left[2] + a.model.cam_left[0] * x[0] / x[2];
			pred[1] = a.model.cam_left[3] + a.model.cam_left[1] * x[1] / x[2];
			if (OBSK == SRBA_OBS_STEREO) {
				double rr[3];
				compose_point3(a.model.right_cam_R2L, x, rr);
				pred[2] = a.model.cam_right[2] + a.model.cam_right[0] * rr[0] / rr[2];
				pred[O - 1] = a.model.cam_right[3] + a.model.cam_right[1] * rr[1] / rr[2];
			}
		}
#pragma unroll
		for (int k = 0; k < O; k++) r[k] = z[k] - pred[k];
	}
	double s2 = 0;
#pragma unroll
	for (int k = 0; k < O; k++) s2 += r[k] * r[k];
	return s2;
}

template <int P, int OBSK>
__global__ void __launch_bounds__(256) srba_eval_residual_kernel(const EvalArgs a) {
	__shared__ double sensor_inv[12];
	__shared__ double wsum[8];
	if (threadIdx.x == 0 && P == 6 && a.model.sensor_pose_kind == SRBA_SENSORPOSE_SE3) PoseOps<6>::inverse(sensor_inv, a.model.sensor_pose);
	__syncthreads();
	double local = 0;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_obs; i += gridDim.x * blockDim.x) local += eval_obs_sqerr<P, OBSK>(a, sensor_inv, i);
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
	if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
	__syncthreads();
	if (threadIdx.x == 0) {
		double s = 0;
		for (int w = 0; w < (int)blockDim.x / 32; w++) s += wsum[w];
		a.partial[blockIdx.x] = s;
	}
}
// fixed-order final sum of the CTA partials (deterministic for a given grid)
__global__ void srba_eval_sum_kernel(const double *partial, int n, double *total) {
	if (threadIdx.x == 0 && blockIdx.x == 0) {
		double s = 0;
		for (int i = 0; i < n; i++) s += partial[i];
		*total = s;
	}
}

// grid_resid = CTAs of the residual kernel (= length of `partial`).
extern "C" int srba_launch_eval_error(const srba_model_t *model, int n_pairs, int n_obs, const double *edge_pose, const int32_t *pair_path_ptr,
                                      const int32_t *path_edge, const uint8_t *path_dir, const double *obs_z, const int32_t *obs_pair,
                                      const uint8_t *obs_inverse, const int32_t *obs_lm, const double *lm_pos, double *pair_pose, double *partial,
                                      double *total, int grid_chain, int grid_resid, void *stream) {
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	EvalArgs a;
	a.model = *model; a.n_pairs = n_pairs; a.n_obs = n_obs; a.edge_pose = edge_pose; a.pair_path_ptr = pair_path_ptr; a.path_edge = path_edge;
	a.path_dir = path_dir; a.obs_z = obs_z; a.obs_pair = obs_pair; a.obs_inverse = obs_inverse; a.obs_lm = obs_lm; a.lm_pos = lm_pos;
	a.pair_pose = pair_pose; a.partial = partial; a.total = total;
	const bool se2 = model->pose_kind == SRBA_POSE_SE2;
	if (n_pairs > 0) {
		if (se2) srba_eval_chain_kernel<3><<<grid_chain, 256, 0, st>>>(a);
		else srba_eval_chain_kernel<6><<<grid_chain, 256, 0, st>>>(a);
	}
	if (se2) srba_eval_residual_kernel<3, SRBA_OBS_RELPOSES2D><<<grid_resid, 256, 0, st>>>(a);
	else if (model->obs_kind == SRBA_OBS_CARTESIAN3D) srba_eval_residual_kernel<6, SRBA_OBS_CARTESIAN3D><<<grid_resid, 256, 0, st>>>(a);
	else if (model->obs_kind == SRBA_OBS_MONO) srba_eval_residual_kernel<6, SRBA_OBS_MONO><<<grid_resid, 256, 0, st>>>(a);
	else if (model->obs_kind == SRBA_OBS_STEREO) srba_eval_residual_kernel<6, SRBA_OBS_STEREO><<<grid_resid, 256, 0, st>>>(a);
	else return (int)cudaErrorInvalidValue;
	srba_eval_sum_kernel<<<1, 32, 0, st>>>(partial, grid_resid, total);
	return (int)cudaGetLastError();
}
#else   // SRBA_CLUSTER: the same program on clusters of SRBA_CLUSTER CTAs (one cluster per problem)
extern "C" int srba_launch_lm_batch_cluster(int model, const char *img, const uint64_t *img_off, char *ws, const uint64_t *ws_off, char *res,
                                            const uint64_t *res_off, const int32_t *order, int n_problems, int smem_bytes, void *stream) {
	constexpr int NT = 256;
	const size_t smem = (size_t)std::min(std::max(smem_bytes, 1024), 160 * 1024);
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)(n_problems * SRBA_CLUSTER), 1, 1);
	cfg.blockDim = dim3(NT, 1, 1);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = static_cast<cudaStream_t>(stream);
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = SRBA_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
#define LAUNCH(P, OBSK)                                                                                                       \
	do {                                                                                                                      \
		auto kern = srba_lm_batch_kernel<P, OBSK, NT, 1, 0>;                                                                  \
		cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                                   \
		const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, img, img_off, ws, ws_off, res, res_off, order, n_problems);      \
		if (e != cudaSuccess) return (int)e;                                                                                  \
	} while (0)
	switch (model) {
		case 0: LAUNCH(3, SRBA_OBS_RELPOSES2D); break;
		case 1: LAUNCH(6, SRBA_OBS_CARTESIAN3D); break;
		case 2: LAUNCH(6, SRBA_OBS_MONO); break;
		case 3: LAUNCH(6, SRBA_OBS_STEREO); break;
		default: return (int)cudaErrorInvalidValue;
	}
#undef LAUNCH
	return (int)cudaGetLastError();
}
#endif
