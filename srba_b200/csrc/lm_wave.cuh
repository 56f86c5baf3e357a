// Grid-wide ("wave") execution of the LM program for batches of HEAVY areas: the same stages as lm_program (lm_kernel.cu), but
// every barrier-separated stage is ONE KERNEL LAUNCH over all areas of the batch -- blockIdx.y = area, blockIdx.x = the CTAs
// that share the area's items -- and the areas advance in lock step, trial by trial, each with its own accept / reject /
// relinearise / stop decisions kept in a small per-area state record in global memory (regime 2):
//
//   per trial:  [schur 1|2|3] -> solve (one CTA per area: the block-sparse L D L^t is a sequential recurrence) -> [features] ->
//               update -> chain (listed pairs) -> residuals (+ rho denominator, per-CTA partial sums) -> ctrl1 (rho, decisions,
//               trace) -> accepted & relinearise: jacobians (2 launches), hessian -> accepted: gradient | rejected: restore ->
//               ctrl2 (lambda / nu schedule, stop rules, next trial or done)
//
// Why: with one CTA per heavy area (lm_program) 148 co-resident areas keep 148 x 8 MB of workspace live -- 10x the L2 -- and
// every stage runs at the occupancy the WHOLE program's register / shared-memory footprint allows (1024 threads x 64 registers
// with 2 KB spill frames).  Here a launch touches only the arrays of its stage, each stage kernel gets its own register
// budget (no spills) and as many CTAs per area as it has items, so the streams over Jacobian blocks / Hessian terms / Schur
// triplets run at full occupancy across all SMs.  Reductions (sum r^2, rho denominator, max |g|, max diag) are per-CTA
// partials combined by the control kernels in a fixed order: deterministic for a given launch configuration.
// The arithmetic of every stage is the code of lm_kernel.cu (same functions, LM_TID / LM_THREADS = position in the area's
// row of CTAs), so parity arguments carry over; only the grouping of the reductions differs.
#pragma once

namespace {

enum { WP_ERR = 0, WP_DENOM, WP_GMAX, WP_KINDS };  // per-CTA partials
#define WAVE_PART_CTAS 64                           // at most this many CTAs per area in a reducing stage

struct WaveArea {  // what `Shared` holds in the one-CTA program (impl/optimize_edges.h:361-363,452-459)
	double lambda, nu, err, rmse, rho, new_err, trial_lambda;
	int iter, stop, trials, accepted, relin, cholfail, cur, status, n_listed, n_refresh, need_outer, go;
	int trial_ok, accept, do_relin, chol_fail;
	double sensor_inv[12];
};

struct WaveArgs {
	const char *img;
	const uint64_t *img_off;
	char *ws;
	const uint64_t *ws_off;
	char *res;
	const uint64_t *res_off;
	const int32_t *order;  // the problems of the wave group
	int n;                 // number of areas
	View *views;           // n
	WaveArea *st;          // n
	double *partials;      // n x WP_KINDS x WAVE_PART_CTAS
	int *active;           // [0]: areas still iterating after the last ctrl2
	int *big, *solo;       // areas (positions in `order`) that have elimination levels below the tail: cluster teams / one CTA; [n] = fill counter
};

// Team of an area in a stage kernel: as many CTAs as the AREA has items (a function of the area alone, so that the grouping of
// its reductions -- hence every bit of its result -- does not depend on which other areas share the batch); the launch grid is
// sized for the largest area of the group and the surplus CTAs of smaller areas leave at once.
__device__ __forceinline__ int wave_team_ctas(int64_t items, int threads, int cap) { return (int)max((int64_t)1, min((items + threads - 1) / threads, (int64_t)cap)); }
__device__ __forceinline__ bool wave_team(int64_t items, int cap) {
	// (never more CTAs than the launch holds: a grid sized too small on the host must cost time, not silently drop items)
	const int gx = min(wave_team_ctas(items, (int)blockDim.x, cap), (int)gridDim.x);
	if ((int)blockIdx.x >= gx) return false;
	if (threadIdx.x == 0) g_wave_gx = gx;
	__syncthreads();
	return true;
}
#define WAVE_CAP 256  // CTAs per area in a non-reducing stage

__device__ __forceinline__ double *wave_part(const WaveArgs &a, int area, int kind) { return a.partials + ((size_t)area * WP_KINDS + kind) * WAVE_PART_CTAS; }

// sum / max of one value per thread over the CTA; result valid in thread 0
__device__ __forceinline__ double cta_sum(double v, double *scr) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	if ((threadIdx.x & 31) == 0) scr[threadIdx.x >> 5] = v;
	__syncthreads();
	double s = 0;
	if (threadIdx.x == 0)
		for (int i = 0; i < (int)blockDim.x / 32; i++) s += scr[i];
	__syncthreads();
	return s;
}
__device__ __forceinline__ double cta_max(double v, double *scr) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
	if ((threadIdx.x & 31) == 0) scr[threadIdx.x >> 5] = v;
	__syncthreads();
	double s = 0;
	if (threadIdx.x == 0)
		for (int i = 0; i < (int)blockDim.x / 32; i++) s = i == 0 ? scr[0] : fmax(s, scr[i]);
	__syncthreads();
	return s;
}

// the scalar control flow between two trials (lm_program::advance)
__device__ void wave_advance(WaveArea &s, const DevProblem *h) {
	for (;;) {
		if (s.need_outer) {
			if (!(s.iter < h->model.max_iters && !s.stop)) { s.go = 0; return; }
			s.rho = 0;
			if (s.lambda >= h->model.max_lambda) s.stop = 1;
			if (s.rmse < h->model.max_error_per_obs_to_stop) s.stop = 1;
			s.need_outer = 0;
		}
		if (s.rho <= 0 && !s.stop) { s.go = 1; return; }
		s.iter++;
		s.need_outer = 1;
	}
}

// ------------------------------------------------------------------------------------------------ set-up
__global__ void wave_views_kernel(const WaveArgs a) {
	const int area = blockIdx.x * blockDim.x + threadIdx.x;
	if (area >= a.n) return;
	const int pi = a.order[area];
	const char *pimg = a.img + a.img_off[pi];
	make_view(a.views[area], reinterpret_cast<const DevProblem *>(pimg), pimg, a.ws + a.ws_off[pi], a.res + a.res_off[pi], nullptr);
	const DevProblem *h = reinterpret_cast<const DevProblem *>(pimg);
	if (h->ldl_tail_level > 0) {  // levels below the tail: forward / backward kernels, as a cluster team or solo
		if (h->ldl_n_dst >= SRBA_WAVE_TEAM_DST_PER_LEVEL * h->ldl_tail_level) a.big[atomicAdd(&a.big[a.n], 1)] = area;
		else a.solo[atomicAdd(&a.solo[a.n], 1)] = area;
	}
}

__global__ void wave_dump_kernel(const WaveArgs a, int tag) {
	const View &v = a.views[0];
	const DevProblem *h = v.h;
	printf("[wave dump %d] n=%d views=%p h=%p wave=%d n_lms=%d n_pairs=%d nA=%d nF=%d ws_sm[lm_cur]=%u ws_off[lm_cur]=%llu lm_cur=%p slot=%p blkF_slot=%p obs_lm=%p blkF_obs=%p JF=%p ws=%p\n", tag, a.n, a.views, h,
	       h->wave, h->n_lms, h->n_pairs, h->n_blkA, h->n_blkF, h->ws_sm[SRBA_W_lm_cur], (unsigned long long)h->ws_off[SRBA_W_lm_cur], v.lm_cur, v.slot, v.blkF_slot, v.obs_lm, v.blkF_obs, v.JF, a.ws);
}

template <int P>
__global__ void __launch_bounds__(256) wave_init_kernel(const WaveArgs a) {
	constexpr int S = PoseOps<P>::S;
	const int area = blockIdx.y;
	const View v = a.views[area];
	const DevProblem *h = v.h;
	if (!wave_team(max((int64_t)h->n_edges * S, (int64_t)2 * h->n_pairs), WAVE_CAP)) return;
	const char *pimg = reinterpret_cast<const char *>(h);
	const double *ge = reinterpret_cast<const double *>(pimg + h->img_off[SRBA_I_edge_pose]);
	const double *gl = reinterpret_cast<const double *>(pimg + h->img_off[SRBA_I_lm_pos]);
	for (size_t i = LM_TID; i < (size_t)h->n_edges * S; i += LM_THREADS) v.edge_cur[i] = ge[i];
	for (size_t i = LM_TID; i < (size_t)h->n_lms * 3; i += LM_THREADS) v.lm_cur[i] = gl[i];
	for (int i = LM_TID; i < 2 * h->n_pairs; i += LM_THREADS) v.slot_listed[i] = 0;
	if (LM_TID == 0) {
		WaveArea &s = a.st[area];
		if (h->model.sensor_pose_kind == SRBA_SENSORPOSE_SE3) PoseOps<6>::inverse(s.sensor_inv, h->model.sensor_pose);
		s.status = SRBA_OK;
		s.go = 1;  // the set-up stages below run for every area
		s.trial_ok = 0; s.accept = 0; s.do_relin = 0; s.cur = 0; s.chol_fail = 0;
	}
}

// compact lists of the listed directed entries and of the pairs that hold one (lm_program, after the first linearisation)
__global__ void __launch_bounds__(64) wave_lists_kernel(const WaveArgs a) {
	const int area = blockIdx.y;
	const View v = a.views[area];
	const DevProblem *h = v.h;
	WaveArea &s = a.st[area];
	if (threadIdx.x < 32) {
		int cnt = 0;
		for (int base = 0; base < 2 * h->n_pairs; base += 32) {
			const int sidx = base + threadIdx.x;
			const bool on = sidx < 2 * h->n_pairs && v.slot_listed[sidx];
			const unsigned m = __ballot_sync(0xffffffffu, on);
			if (on) v.listed_idx[cnt + __popc(m & ((1u << threadIdx.x) - 1))] = sidx;
			cnt += __popc(m);
		}
		if (threadIdx.x == 0) s.n_listed = cnt;
	} else {
		const int l = threadIdx.x - 32;
		int cnt = 0;
		for (int base = 0; base < h->n_pairs; base += 32) {
			const int t = base + l;
			const bool on = t < h->n_pairs && (v.slot_listed[2 * t] | v.slot_listed[2 * t + 1]);
			const unsigned m = __ballot_sync(0xffffffffu, on);
			if (on) v.refresh_idx[cnt + __popc(m & ((1u << l) - 1))] = t;
			cnt += __popc(m);
		}
		if (l == 0) s.n_refresh = cnt;
	}
}

// ------------------------------------------------------------------------------------------------ stage kernels
// `when`: which areas take part -- 0: every area still iterating (go); 1: go and the solve of this trial succeeded;
// 2: ... and the step was accepted and relinearises; 3: ... accepted; 4: ... rejected; 5: set-up (all areas)
enum { WHEN_GO = 0, WHEN_TRIAL_OK, WHEN_RELIN, WHEN_ACCEPT, WHEN_REJECT, WHEN_SETUP };
__device__ __forceinline__ bool wave_on(const WaveArea &s, int when) {
	if (when == WHEN_SETUP) return true;
	if (!s.go) return false;
	if (when == WHEN_GO) return true;
	if (!s.trial_ok) return false;
	if (when == WHEN_TRIAL_OK) return true;
	if (when == WHEN_RELIN) return s.accept && s.do_relin;
	if (when == WHEN_ACCEPT) return s.accept;
	return !s.accept;
}

template <int P>
__global__ void __launch_bounds__(256) wave_chain_kernel(const WaveArgs a, int when) {
	const int area = blockIdx.y;
	const WaveArea &s = a.st[area];
	if (!wave_on(s, when)) return;
	if (!wave_team(when == WHEN_SETUP ? a.views[area].h->n_pairs : s.n_refresh, WAVE_CAP)) return;
	stage_chain<P>(a.views[area], when == WHEN_SETUP ? -1 : s.n_refresh);
}

// launch A: validity reset + evaluation (phases 1, 2); launch B: commit by walk order / listed marks (phase 3)
template <int P, int OBSK>
__global__ void __launch_bounds__(128) wave_jacobians_kernel(const WaveArgs a, int when, int second) {
	const int area = blockIdx.y;
	if (!wave_on(a.st[area], when)) return;
	{
		const DevProblem *h = a.views[area].h;
		if (!wave_team((int64_t)h->n_blkA + h->n_blkF, WAVE_CAP)) return;
	}
	const bool first = when == WHEN_SETUP;
	if (!second) {
		stage_jacobians<P, OBSK>(a.views[area], first, 1);
		stage_jacobians<P, OBSK>(a.views[area], first, 2);
	} else
		stage_jacobians<P, OBSK>(a.views[area], first, 3);
}

template <int P, int OBSK>
__global__ void __launch_bounds__(128) wave_hessian_kernel(const WaveArgs a, int when) {
	const int area = blockIdx.y;
	if (!wave_on(a.st[area], when)) return;
	{
		const DevProblem *h = a.views[area].h;
		if (!wave_team(P == 6 ? (int64_t)6 * h->n_HAp + 6 * h->n_HApf + 3 * h->n_unk_lms : (int64_t)9 * h->n_HAp, WAVE_CAP)) return;
	}
	stage_hessian<P, OBSK>(a.views[area]);
}

// residuals into resid[cur] (set-up) or resid[cur ^ 1] (trial) + this CTA's partial of the total and of rho's denominator
template <int P, int OBSK>
__global__ void __launch_bounds__(256) wave_residuals_kernel(const WaveArgs a, int when) {
	__shared__ double scr[8];
	const int area = blockIdx.y;
	const WaveArea &s = a.st[area];
	if (!wave_on(s, when)) return;
	const View &v = a.views[area];
	const DevProblem *h = v.h;
	if (!wave_team(h->n_obs, WAVE_PART_CTAS)) return;
	const bool setup = when == WHEN_SETUP;
	double *resid = v.resid[setup ? 0 : (s.cur ^ 1)];
	double dl = 0;
	if (!setup) {  // rho (:585)
		const bool keep_grad = (OBSK != SRBA_OBS_RELPOSES2D) && no_schur_solver(h);
		const double *gr = keep_grad ? v.grad0 : v.grad;
		const double lambda = s.trial_lambda;
		for (int i = LM_TID; i < h->n_scalars; i += LM_THREADS) dl += v.delta[i] * (lambda * v.delta[i] + gr[i]);
	}
	const double local = stage_residuals<P, OBSK>(a.views[area], resid, s.sensor_inv);
	const double se = cta_sum(local, scr), sd = cta_sum(dl, scr);
	if (threadIdx.x == 0) { wave_part(a, area, WP_ERR)[blockIdx.x] = se; wave_part(a, area, WP_DENOM)[blockIdx.x] = sd; }
}

// accepted (or set-up): gradient from the new residuals + partial of max |g|;  rejected: restore
template <int P, int OBSK>
__global__ void __launch_bounds__(128) wave_gradient_or_restore_kernel(const WaveArgs a, int setup) {
	__shared__ double scr[8];
	const int area = blockIdx.y;
	const WaveArea &s = a.st[area];
	if (!setup && (!s.go || !s.trial_ok)) return;
	const View &v = a.views[area];
	if (!wave_team(v.h->n_scalars, WAVE_PART_CTAS)) return;  // (restore: the same team strides over its three lists)
	if (setup || s.accept) {
		const double gm = stage_gradient<P, OBSK>(v, v.resid[setup ? 0 : (s.cur ^ 1)]);
		const double m = cta_max(gm, scr);
		if (threadIdx.x == 0) wave_part(a, area, WP_GMAX)[blockIdx.x] = m;
	} else
		stage_restore<P>(v, s.n_listed);
}

template <int P, int OBSK>
__global__ void __launch_bounds__(128) wave_schur_kernel(const WaveArgs a, int phase) {
	const int area = blockIdx.y;
	const WaveArea &s = a.st[area];
	if (!s.go) return;
	const View &v = a.views[area];
	const DevProblem *h = v.h;
	if (h->n_unk_lms == 0 && h->s_alias_hap) return;
	if (!wave_team(phase == 1 ? (int64_t)h->n_unk_lms : phase == 2 ? (int64_t)h->n_HApf * P * 3 : (int64_t)h->n_S * (P == 6 ? (SRBA_SCHUR_DMMA ? 32 : 6) : 9) + h->n_dense, WAVE_CAP)) return;
	if (phase == 1 && (OBSK != SRBA_OBS_RELPOSES2D) && no_schur_solver(h))  // (grad is next read by phase 3)
		for (int i = LM_TID; i < h->n_scalars; i += LM_THREADS) v.grad[i] = v.grad0[i];
	stage_schur<P>(v, s.lambda, phase);
}

// ------------------------------------------------------------------------------------------------ the reduced system
// Block-sparse L D L^t on 3x3 tiles with the LEVEL schedule of the packer (capi.cpp: symbolic_levels), solving what Eigen::LLT<Upper> +
// solve see after the dense fill (impl/lev-marq_solvers.h:492-549).  Four launches per trial:
//   fwd    (areas with levels below the tail; a CLUSTER of WAVE_SOLVE_CTAS CTAs each)  gather + the levels below the tail, tiles in the
//          global workspace (L2):  A  W(k,.) = inv(D_k) T(k,.), one 9-lane group per tile (pivot block inverted by cofactors; rejected
//          exactly when a leading minor is <= 0, i.e. when Eigen::LLT would meet a non-positive pivot, lev-marq_solvers.h:527-534);
//          B  one 9-lane group per destination tile: T(a,b) -= sum_k T(k,a)^t W(k,b) over the contributions of this level, in ascending
//          source row (deterministic, no atomics); two cluster barriers per level;
//   tail   (every area; one CTA)  the last <= SRBA_WAVE_TAIL_ROWS tile rows -- the dense trailing front of a minimum-degree order, one
//          row per level -- as a DENSE upper triangle in shared memory: one thread per destination tile, two CTA barriers per row,
//          then their back-substitution in one warp;
//   bwd    (as fwd)  x_k = y_k - sum_j W(k,j) x_j down the levels below the tail, one warp per row;
//   finish (every area; one CTA)  increments back in the reference's unknown order; a failed factorisation (not positive definite,
//          optimize_edges.h:476-485) is booked -- the later stages of this trial skip the area (trial_ok = 0).
#ifndef WAVE_SOLVE_CTAS
#define WAVE_SOLVE_CTAS 8      // > 8: non-portable cluster size (tools/variant_bench.py experiments)
#endif
#ifndef WAVE_SOLVE_NT
#define WAVE_SOLVE_NT 1024
#endif
#ifndef WAVE_TEAM_MIN_CTAS
#define WAVE_TEAM_MIN_CTAS (1024 / WAVE_SOLVE_NT)   // resident team CTAs per SM the register budget is set for
#endif
#ifndef WAVE_FWD_UNROLL
#define WAVE_FWD_UNROLL 1      // U > 1: U destination entries per thread and pass in the forward levels (independent load chains)
#endif
#define WAVE_TAIL_NT 1024

// Team of an area in the forward / backward kernels: a cluster of WAVE_SOLVE_CTAS x WAVE_SOLVE_NT threads when its levels are wide (many
// destination tiles per level, list a.big), otherwise one CTA of WAVE_SOLO_NT threads (list a.solo; several per SM, CTA barriers).
// The same predicate on the host (capi.cpp: batch_create_impl) sizes the two launches.
#define WAVE_SOLO_NT 512
#define WAVE_TEAM_SYNC()                                       \
	do {                                                       \
		if (TEAM) cooperative_groups::this_cluster().sync();   \
		else __syncthreads();                                  \
	} while (0)

#ifdef SRBA_WAVE_TIMERS
__device__ __forceinline__ unsigned long long wave_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define WT_DECL unsigned long long wt_ = wave_now(); unsigned long long *wsc_ = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(vg.stats) - h->res_stats + h->res_stage_cycles)
#define WT_ADD(slot) do { if (tid == 0) { const unsigned long long n_ = wave_now(); wsc_[slot] += n_ - wt_; wt_ = n_; } } while (0)
#else
#define WT_DECL
#define WT_ADD(slot)
#endif

// pivot block (upper part of a symmetric 3x3) -> cofactors / determinant; `bad` = a leading minor is <= 0
struct Piv3 { double c00, c01, c02, c11, c12, c22, idet; bool bad; };
__device__ __forceinline__ Piv3 piv3(double d00, double d01, double d02, double d11, double d12, double d22) {
	Piv3 p;
	p.c00 = d11 * d22 - d12 * d12; p.c01 = d02 * d12 - d01 * d22; p.c02 = d01 * d12 - d02 * d11;
	p.c11 = d00 * d22 - d02 * d02; p.c12 = d01 * d02 - d00 * d12; p.c22 = d00 * d11 - d01 * d01;
	const double det = d00 * p.c00 + d01 * p.c01 + d02 * p.c02;
	p.bad = d00 <= 0 || p.c22 <= 0 || det <= 0;
	p.idet = 1.0 / det;
	return p;
}
// element (wa, wb) of tile `tl` of the reduced system: reduced block element (+lambda on the diagonal) / reduced gradient / fill-in
template <int P>
__device__ __forceinline__ double wave_tile_src(int so, int wa, int wb, const double *S, const double *grad, double lambda) {
	if (so >= 0) return S[(so & 0x1fffffff) + ((so >> 29) & 1 ? wb * P + wa : wa * P + wb)] + (((so >> 30) && wa == wb) ? lambda : 0.0);
	if (so <= -2 && wb == 0) return grad[3 * (-so - 2) + wa];
	return 0.0;
}

// One contribution T(k,a)^t inv(D_k) T(k,b) to element (wa, wb) of a destination tile: D_k is final when the level of row k runs, so
// inv(D_k) T(k,b) is formed on the fly (no separate W pass, no barrier for it).
__device__ __forceinline__ double wave_contrib(const double *T, int ta, int tb, int td, int wa, int wb) {
	const double *D = T + (size_t)td * 12, *Ta = T + (size_t)ta * 12 + wa * 4, *Tb = T + (size_t)tb * 12 + wb * 4;
	const double d00 = D[0], d01 = D[4], d02 = D[8], d11 = D[5], d12 = D[9], d22 = D[10];
	const double a0 = Ta[0], a1 = Ta[1], a2 = Ta[2], b0 = Tb[0], b1 = Tb[1], b2 = Tb[2];
	const Piv3 pv = piv3(d00, d01, d02, d11, d12, d22);
	const double w0 = pv.c00 * b0 + pv.c01 * b1 + pv.c02 * b2, w1 = pv.c01 * b0 + pv.c11 * b1 + pv.c12 * b2, w2 = pv.c02 * b0 + pv.c12 * b1 + pv.c22 * b2;
	return (a0 * w0 + a1 * w1 + a2 * w2) * pv.idet;
}

// TEAM: a cluster of WAVE_SOLVE_CTAS CTAs per area (areas with wide levels, list a.big); otherwise one CTA per area (list a.solo)
template <int P, int OBSK, bool TEAM>
__global__ void __launch_bounds__(TEAM ? WAVE_SOLVE_NT : WAVE_SOLO_NT, TEAM ? WAVE_TEAM_MIN_CTAS : 1) wave_solve_fwd_kernel(const WaveArgs a) {
	const int area = TEAM ? a.big[blockIdx.y] : a.solo[blockIdx.y];
	WaveArea &s = a.st[area];
	if (!s.go) return;  // (the same decision in every CTA of the cluster)
	const View &vg = a.views[area];
	const DevProblem *h = vg.h;
	const double lambda = s.lambda;
	const int ntiles = h->ldl_n_tiles, L0 = h->ldl_tail_level, ndst = h->ldl_n_dst;
	double *T = vg.ldl_tiles;
	const int32_t *src = vg.ldl_src;
	const int2 *lvl = reinterpret_cast<const int2 *>(vg.ldl_phase);      // {first row, first destination} per level
	const int4 *drec = reinterpret_cast<const int4 *>(vg.ldl_upd);       // {destination, first contribution (a, b, diagonal)}
	const int32_t *xptr = vg.ldl_upd + 4 * (size_t)ndst;                 // further contributions of a destination (rare)
	const int4 *xrec = reinterpret_cast<const int4 *>(xptr + ((ndst + 1 + 3) & ~3));
	const double *S = vg.S, *grad = vg.grad;
	const int tid = (int)(blockIdx.x * blockDim.x + threadIdx.x), NTH = (int)(gridDim.x * blockDim.x);
	const int lane = tid & 31, gwarp = tid >> 5, NW = NTH >> 5;
	const int wg = lane / 9, we = lane - wg * 9, wb = we / 3, wa = we - wb * 3, eo = wb * 4 + wa;  // 9-lane tile group: element (wa,wb)
	const bool glane = lane < 27;
	WT_DECL;
	// gather: every tile (the tail kernel reads its rows from here too)
	if (glane)
		for (int tl = 3 * gwarp + wg; tl < ntiles; tl += 3 * NW) T[(size_t)tl * 12 + eo] = wave_tile_src<P>(src[tl], wa, wb, S, grad, lambda);
	WAVE_TEAM_SYNC();
	WT_ADD(0);
	for (int L = 0; L < L0; L++) {
		const int e1 = lvl[L + 1].y;
#if WAVE_FWD_UNROLL > 1
		if (glane)
			for (int e = lvl[L].y + gwarp * 3 + wg; e < e1; e += NW * 3 * WAVE_FWD_UNROLL) {
				int4 rc[WAVE_FWD_UNROLL];
				int xa[WAVE_FWD_UNROLL], xb[WAVE_FWD_UNROLL];
				double cur[WAVE_FWD_UNROLL], acc[WAVE_FWD_UNROLL];
#pragma unroll
				for (int u = 0; u < WAVE_FWD_UNROLL; u++) {
					const int f = e + u * NW * 3, g = f < e1 ? f : e;
					rc[u] = drec[g]; xa[u] = xptr[g]; xb[u] = f < e1 ? xptr[g + 1] : xa[u];
				}
#pragma unroll
				for (int u = 0; u < WAVE_FWD_UNROLL; u++) cur[u] = T[(size_t)rc[u].x * 12 + eo];
#pragma unroll
				for (int u = 0; u < WAVE_FWD_UNROLL; u++) acc[u] = wave_contrib(T, rc[u].y, rc[u].z, rc[u].w, wa, wb);
#pragma unroll
				for (int u = 0; u < WAVE_FWD_UNROLL; u++)
					for (int c = xa[u]; c < xb[u]; c++) { const int4 xr = xrec[c]; acc[u] += wave_contrib(T, xr.x, xr.y, xr.z, wa, wb); }
#pragma unroll
				for (int u = 0; u < WAVE_FWD_UNROLL; u++)
					if (e + u * NW * 3 < e1) T[(size_t)rc[u].x * 12 + eo] = cur[u] - acc[u];
			}
#else
		if (glane)
			for (int e = lvl[L].y + gwarp * 3 + wg; e < e1; e += NW * 3) {
				const int4 rec = drec[e];
				const int x0 = xptr[e], x1 = xptr[e + 1];
				double *dp = T + (size_t)rec.x * 12 + eo;
				const double cur = *dp;
				double acc = wave_contrib(T, rec.y, rec.z, rec.w, wa, wb);
				for (int c = x0; c < x1; c++) {
					const int4 xr = xrec[c];
					acc += wave_contrib(T, xr.x, xr.y, xr.z, wa, wb);
				}
				*dp = cur - acc;
			}
#endif
		WT_ADD(3);
		WAVE_TEAM_SYNC();
		WT_ADD(4);
	}
}

// local tile index of (i, j), j in i..r (r = right-hand side), in the dense upper triangle of r rows
__device__ __forceinline__ int tail_base(int i, int r) { return i * (r + 1) - (i * (i - 1)) / 2; }
__host__ __device__ inline size_t wave_tail_smem_bytes(int r) {
	const size_t ntl = (size_t)r * (r + 3) / 2;
	return ntl * 72 + 2 * (size_t)(r + 1) * 72 + ntl * 4 + (size_t)3 * r * 8 + 64;
}

template <int P, int OBSK>
__global__ void __launch_bounds__(WAVE_TAIL_NT) wave_solve_tail_kernel(const WaveArgs a) {
	extern __shared__ double tail_sm[];
	const int area = blockIdx.x;
	WaveArea &s = a.st[area];
	if (!s.go) return;
	const View &vg = a.views[area];
	const DevProblem *h = vg.h;
	const double lambda = s.lambda;
	const int ntiles = h->ldl_n_tiles, L0 = h->ldl_tail_level, ntr = h->n_dense / 3;
	const int2 *lvl = reinterpret_cast<const int2 *>(vg.ldl_phase);
	const int2 *rowinfo = reinterpret_cast<const int2 *>(vg.ldl_row);
	const int row0 = lvl[L0].x, r = ntr - row0;
	if (r <= 0) return;
	const int ntl = r * (r + 3) / 2;
	double *Tl = tail_sm;                                  // ntl tiles of 9 doubles, row-major M[row * 3 + col]
	double *Wbuf = Tl + (size_t)ntl * 9;                   // 2 x (r + 1) tiles: W(k, .) of the current / previous pivot row
	uint32_t *ij = reinterpret_cast<uint32_t *>(Wbuf + (size_t)2 * (r + 1) * 9);  // local tile -> (i << 16 | j)
	double *xs = reinterpret_cast<double *>(ij + ntl + (ntl & 1));
	double *T = vg.ldl_tiles, *x = T + (size_t)ntiles * 24;
	const int32_t *tij = vg.ldl_tile_ij, *src = vg.ldl_src;
	const int nt = (int)blockDim.x, t_ = (int)threadIdx.x;
#ifdef SRBA_WAVE_TIMERS
	const int tid = t_;
#endif
	WT_DECL;
	for (int i = t_; i < ntl * 9; i += nt) Tl[i] = 0.0;
	for (int i = t_; i < r; i += nt) {
		const int b0 = tail_base(i, r);
		for (int j = i; j <= r; j++) ij[b0 + j - i] = ((uint32_t)i << 16) | (uint32_t)j;
	}
	__syncthreads();
	{  // the rows of the tail: from the global tiles (areas with levels below the tail: updated by them) or straight from the reduced blocks
		const bool from_T = L0 > 0;
		const int g0 = rowinfo[2 * row0].x;
		for (int q = t_; q < (ntiles - g0) * 9; q += nt) {
			const int tl = g0 + q / 9, e = q % 9, wa = e / 3, wb = e % 3;
			const int gij = tij[tl], i = (gij >> 16) - row0, gj = gij & 0xffff, j = gj == ntr ? r : gj - row0;
			Tl[(size_t)(tail_base(i, r) + j - i) * 9 + e] = from_T ? T[(size_t)tl * 12 + wb * 4 + wa] : wave_tile_src<P>(src[tl], wa, wb, vg.S, vg.grad, lambda);
		}
	}
	__syncthreads();
	for (int k = 0; k < r; k++) {
		const int bk = tail_base(k, r);
		double *buf = Wbuf + (size_t)(k & 1) * (r + 1) * 9;
		const double *prev = Wbuf + (size_t)((k & 1) ^ 1) * (r + 1) * 9;
		// ---- 1: W(k, j) = inv(D_k) T(k, j), j = k+1..r; the previous row's W goes to its place (read by the back-substitution)
		{
			const double *D = Tl + (size_t)bk * 9;
			const Piv3 pv = piv3(D[0], D[1], D[2], D[4], D[5], D[8]);
			if (t_ == 0 && pv.bad) s.chol_fail = 1;
			for (int q = t_; q < (r - k) * 9; q += nt) {
				const int jj = 1 + q / 9, e = q % 9, wa = e / 3, wb = e % 3;
				const double i0 = (wa == 0 ? pv.c00 : wa == 1 ? pv.c01 : pv.c02) * pv.idet, i1 = (wa == 0 ? pv.c01 : wa == 1 ? pv.c11 : pv.c12) * pv.idet,
				             i2 = (wa == 0 ? pv.c02 : wa == 1 ? pv.c12 : pv.c22) * pv.idet;
				const double *Tc = Tl + (size_t)(bk + jj) * 9 + wb;
				buf[jj * 9 + e] = i0 * Tc[0] + i1 * Tc[3] + i2 * Tc[6];
			}
			if (k > 0) {
				const int bp = tail_base(k - 1, r);
				for (int q = t_; q < (r - k + 1) * 9; q += nt) Tl[(size_t)(bp + 1) * 9 + q] = prev[9 + q];
			}
		}
		__syncthreads();
		// ---- 2: T(i, j) -= T(k, i)^t W(k, j), k < i <= j: the tiles of the rows below k are one contiguous range; one thread per tile
		for (int u = tail_base(k + 1, r) + t_; u < ntl; u += nt) {
			const uint32_t c = ij[u];
			const int i = (int)(c >> 16), j = (int)(c & 0xffffu);
			const double *A = Tl + (size_t)(bk + i - k) * 9, *B = buf + (size_t)(j - k) * 9;
			double *Cc = Tl + (size_t)u * 9;
			const double a0 = A[0], a1 = A[1], a2 = A[2], a3 = A[3], a4 = A[4], a5 = A[5], a6 = A[6], a7 = A[7], a8 = A[8];
			const double b0 = B[0], b1 = B[1], b2 = B[2], b3 = B[3], b4 = B[4], b5 = B[5], b6 = B[6], b7 = B[7], b8 = B[8];
			Cc[0] -= a0 * b0 + a3 * b3 + a6 * b6; Cc[1] -= a0 * b1 + a3 * b4 + a6 * b7; Cc[2] -= a0 * b2 + a3 * b5 + a6 * b8;
			Cc[3] -= a1 * b0 + a4 * b3 + a7 * b6; Cc[4] -= a1 * b1 + a4 * b4 + a7 * b7; Cc[5] -= a1 * b2 + a4 * b5 + a7 * b8;
			Cc[6] -= a2 * b0 + a5 * b3 + a8 * b6; Cc[7] -= a2 * b1 + a5 * b4 + a8 * b7; Cc[8] -= a2 * b2 + a5 * b5 + a8 * b8;
		}
		__syncthreads();
	}
	WT_ADD(7);
	{  // W of the last row (its right-hand-side tile only)
		const double *last = Wbuf + (size_t)((r - 1) & 1) * (r + 1) * 9;
		if (t_ < 9) Tl[(size_t)(tail_base(r - 1, r) + 1) * 9 + t_] = last[9 + t_];
	}
	__syncthreads();
	// ---- back-substitution inside the tail: x_i = W(i, rhs) - sum_{j > i} W(i, j) x_j, one warp
	if (t_ < 32) {
		for (int i = r - 1; i >= 0; i--) {
			const int bi = tail_base(i, r);
			double p0 = 0, p1 = 0, p2 = 0;
			for (int j = i + 1 + t_; j < r; j += 32) {
				const double *Wij = Tl + (size_t)(bi + j - i) * 9, *xj = xs + 3 * j;
				const double x0 = xj[0], x1 = xj[1], x2 = xj[2];
				p0 += Wij[0] * x0 + Wij[1] * x1 + Wij[2] * x2;
				p1 += Wij[3] * x0 + Wij[4] * x1 + Wij[5] * x2;
				p2 += Wij[6] * x0 + Wij[7] * x1 + Wij[8] * x2;
			}
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) {
				p0 += __shfl_xor_sync(0xffffffffu, p0, o); p1 += __shfl_xor_sync(0xffffffffu, p1, o); p2 += __shfl_xor_sync(0xffffffffu, p2, o);
			}
			if (t_ < 3) xs[3 * i + t_] = Tl[(size_t)(bi + r - i) * 9 + 3 * t_] - (t_ == 0 ? p0 : t_ == 1 ? p1 : p2);
			__syncwarp();
		}
	}
	__syncthreads();
	WT_ADD(8);
	for (int i = t_; i < 3 * r; i += nt) x[3 * row0 + i] = xs[i];
}

// x_k = inv(D_k) (y_k - sum_j T(k,j) x_j) down the levels below the tail, one warp per row; the pivot blocks are final here: a block with a
// non-positive leading minor books the failed factorisation.  ONE CTA per area (a level has a handful of rows: a cluster would only add
// barrier latency).  The chain of a level is  level bounds -> row bounds -> column indices -> x  -- four dependent memory round trips when
// everything sits in global memory, times 80-240 levels.  So the CTA first stages the level bounds, the row bounds and x (the tail's
// solution; then every level's) in shared memory, and each warp loads the pivot block, the right-hand side and its first 32 tiles + column
// indices of its row of the NEXT level while the current level is being solved: after a barrier only shared-memory reads of x remain.
// Levels with few rows (the chain of a dense front: ONE row of 40-70 tiles per level) split each row over 2-4 warps.
// The levels of such a chain rotate over four groups of warps, each loading its next row four levels ahead (registers).
// `smem_bytes` = the launch's dynamic shared memory; an area that needs more runs the same loop on the global arrays.
struct BwdRow {
	int r, t0, t1, j;
	double d00, d01, d02, d11, d12, d22, y0, y1, y2, m0, m1, m2, m3, m4, m5, m6, m7, m8;
};
template <int P, int OBSK>
__global__ void __launch_bounds__(WAVE_SOLO_NT) wave_solve_bwd_kernel(const WaveArgs a, int big_list, int smem_bytes) {
	extern __shared__ double bwd_sm[];
	__shared__ double bwd_part[3 * (WAVE_SOLO_NT / 32)];
	const int area = big_list ? a.big[blockIdx.y] : a.solo[blockIdx.y];
	WaveArea &s = a.st[area];
	if (!s.go) return;
	const View &vg = a.views[area];
	const DevProblem *h = vg.h;
	const int ntiles = h->ldl_n_tiles, L0 = h->ldl_tail_level, nx = h->n_dense;
	const double *T = vg.ldl_tiles;
	double *xg = vg.ldl_tiles + (size_t)ntiles * 24;
	const int2 *rowinfo = reinterpret_cast<const int2 *>(vg.ldl_row);
	const int32_t *tij = vg.ldl_tile_ij;
	const int2 *lvl = reinterpret_cast<const int2 *>(vg.ldl_phase);
	const int tid = (int)threadIdx.x, NTH = (int)blockDim.x, lane = tid & 31, warp = tid >> 5, NW = NTH >> 5;
	if (L0 <= 0) return;
	const int row0 = lvl[L0].x;   // rows below the tail
	const bool staged = (size_t)nx * 8 + (size_t)row0 * 8 + (size_t)(L0 + 1) * 4 + 16 <= (size_t)smem_bytes;
	double *x = staged ? bwd_sm : xg;
	int2 *srow = reinterpret_cast<int2 *>(bwd_sm + nx);
	int *slvl = reinterpret_cast<int *>(srow + row0);
	if (staged) {
		for (int i = 3 * row0 + tid; i < nx; i += NTH) x[i] = xg[i];
		for (int i = tid; i < row0; i += NTH) srow[i] = rowinfo[2 * i];
		for (int i = tid; i <= L0; i += NTH) slvl[i] = lvl[i].x;
		__syncthreads();
	}
	auto level_row = [&](int L) -> int { return staged ? slvl[L] : lvl[L].x; };
	// A level with few rows (the chain of a dense front: one row of 40-70 tiles per level) gives each row to a group of G warps
	auto group_of = [&](int n_rows) -> int { return 4 * n_rows <= NW ? 4 : (2 * n_rows <= NW ? 2 : 1); };
	// everything of this warp's share of its first row in level L that does not depend on x
	auto fetch_row = [&](int r, int G, BwdRow &q) {
		const int2 ri = staged ? srow[r] : rowinfo[2 * r];
		q.r = r; q.t0 = ri.x; q.t1 = ri.y;
		const double *D = T + (size_t)ri.x * 12, *y = T + (size_t)(ri.y - 1) * 12;
		q.d00 = D[0]; q.d01 = D[4]; q.d02 = D[8]; q.d11 = D[5]; q.d12 = D[9]; q.d22 = D[10];
		q.y0 = y[0]; q.y1 = y[1]; q.y2 = y[2];
		const int tl = ri.x + 1 + 32 * (warp % G) + lane;
		q.j = -1;
		if (tl < ri.y - 1) {
			const double *Tr = T + (size_t)tl * 12;
			q.j = tij[tl] & 0xffff;
			q.m0 = Tr[0]; q.m1 = Tr[1]; q.m2 = Tr[2]; q.m3 = Tr[4]; q.m4 = Tr[5]; q.m5 = Tr[6]; q.m6 = Tr[8]; q.m7 = Tr[9]; q.m8 = Tr[10];
		}
	};
	// Chain levels (ONE row) rotate over the NW / 4 groups of four warps: level L belongs to group L mod (NW / 4), which loads its row of
	// level L - NW / 4 as soon as it has finished level L -- several levels of lead over the memory round trip.
	const int NG = NW / 4, group = warp / 4;
	BwdRow pre;
	int pre_L = -1;
	auto prefetch_chain = [&](int L) {
		if (L < 0 || L % NG != group) return;
		const int r0 = level_row(L);
		if (level_row(L + 1) - r0 != 1) return;
		fetch_row(r0, 4, pre);
		pre_L = L;
	};
	double *part = reinterpret_cast<double *>(bwd_part);
	// partial sum_j T(k,j) x_j of this warp's share of the row (all lanes hold it)
	auto row_partial = [&](const BwdRow &q, int G, double &p0, double &p1, double &p2) {
		p0 = p1 = p2 = 0;
		if (q.j >= 0) {
			const double x0 = x[3 * q.j], x1 = x[3 * q.j + 1], x2 = x[3 * q.j + 2];
			p0 = q.m0 * x0 + q.m3 * x1 + q.m6 * x2;
			p1 = q.m1 * x0 + q.m4 * x1 + q.m7 * x2;
			p2 = q.m2 * x0 + q.m5 * x1 + q.m8 * x2;
		}
		for (int tl = q.t0 + 1 + 32 * (G + warp % G) + lane; tl < q.t1 - 1; tl += 32 * G) {   // rows longer than one pass of the group
			const double *xj = x + 3 * (tij[tl] & 0xffff);
			const double *Tr = T + (size_t)tl * 12;
			const double x0 = xj[0], x1 = xj[1], x2 = xj[2];
			p0 += Tr[0] * x0 + Tr[4] * x1 + Tr[8] * x2;
			p1 += Tr[1] * x0 + Tr[5] * x1 + Tr[9] * x2;
			p2 += Tr[2] * x0 + Tr[6] * x1 + Tr[10] * x2;
		}
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			p0 += __shfl_xor_sync(0xffffffffu, p0, o); p1 += __shfl_xor_sync(0xffffffffu, p1, o); p2 += __shfl_xor_sync(0xffffffffu, p2, o);
		}
	};
	auto finish_row = [&](const BwdRow &q, double p0, double p1, double p2) {   // one lane
		const Piv3 pv = piv3(q.d00, q.d01, q.d02, q.d11, q.d12, q.d22);
		if (pv.bad) s.chol_fail = 1;
		const double y0 = q.y0 - p0, y1 = q.y1 - p1, y2 = q.y2 - p2;
		x[3 * q.r + 0] = (pv.c00 * y0 + pv.c01 * y1 + pv.c02 * y2) * pv.idet;
		x[3 * q.r + 1] = (pv.c01 * y0 + pv.c11 * y1 + pv.c12 * y2) * pv.idet;
		x[3 * q.r + 2] = (pv.c02 * y0 + pv.c12 * y1 + pv.c22 * y2) * pv.idet;
	};
	for (int d = 0; d < NG; d++) prefetch_chain(L0 - 1 - d);
	WT_DECL;
	for (int L = L0 - 1; L >= 0; L--) {
		const int r0 = level_row(L), r1 = level_row(L + 1), n_rows = r1 - r0;
		double p0, p1, p2;
		if (n_rows == 1) {
			const bool mine = (L % NG == group);
			if (mine) {
				if (pre_L != L) fetch_row(r0, 4, pre);
				row_partial(pre, 4, p0, p1, p2);
				if (lane == 0) { part[3 * warp] = p0; part[3 * warp + 1] = p1; part[3 * warp + 2] = p2; }
			}
			__syncthreads();
			if (mine) {
				if (warp % 4 == 0 && lane == 0) {
					p0 = part[3 * warp]; p1 = part[3 * warp + 1]; p2 = part[3 * warp + 2];
					for (int g = 1; g < 4; g++) { p0 += part[3 * (warp + g)]; p1 += part[3 * (warp + g) + 1]; p2 += part[3 * (warp + g) + 2]; }
					finish_row(pre, p0, p1, p2);
				}
				prefetch_chain(L - NG);
			}
		} else {
			const int G = group_of(n_rows);
			BwdRow q;
			q.r = -1;
			if (r0 + warp / G < r1) {
				fetch_row(r0 + warp / G, G, q);
				row_partial(q, G, p0, p1, p2);
			}
			if (G == 1) {
				if (q.r >= 0) {
					if (lane == 0) finish_row(q, p0, p1, p2);
					for (int r = q.r + NW; r < r1; r += NW) {   // more rows than warps in this level (the wide first levels)
						fetch_row(r, 1, q);
						row_partial(q, 1, p0, p1, p2);
						if (lane == 0) finish_row(q, p0, p1, p2);
					}
				}
			} else {
				if (q.r >= 0 && lane == 0) { part[3 * warp] = p0; part[3 * warp + 1] = p1; part[3 * warp + 2] = p2; }
				__syncthreads();
				if (q.r >= 0 && warp % G == 0 && lane == 0) {
					p0 = part[3 * warp]; p1 = part[3 * warp + 1]; p2 = part[3 * warp + 2];
					for (int g = 1; g < G; g++) { p0 += part[3 * (warp + g)]; p1 += part[3 * (warp + g) + 1]; p2 += part[3 * (warp + g) + 2]; }
					finish_row(q, p0, p1, p2);
				}
			}
		}
		WT_ADD(5);
		__syncthreads();
		WT_ADD(6);
	}
	if (staged)
		for (int i = tid; i < 3 * row0; i += NTH) xg[i] = x[i];
}

template <int P, int OBSK>
__global__ void __launch_bounds__(256) wave_solve_finish_kernel(const WaveArgs a) {
	__shared__ int sh_fail;
	const int area = blockIdx.x;
	WaveArea &s = a.st[area];
	if (!s.go) return;
	const View &vg = a.views[area];
	const DevProblem *h = vg.h;
	const int n = h->n_dense;
	const double *x = vg.ldl_tiles + (size_t)h->ldl_n_tiles * 24;
	const int32_t *orow = vg.ldl_orow;
	const bool keep_grad = (OBSK != SRBA_OBS_RELPOSES2D) && no_schur_solver(h);
	if (threadIdx.x == 0) sh_fail = s.chol_fail;
	__syncthreads();
	if (keep_grad) {  // a landmark block that cannot be eliminated = the whole-system Cholesky fails (lev-marq_solvers.h:171-175)
		for (int k = threadIdx.x; k < h->n_unk_lms; k += blockDim.x)
			if (!vg.lm_inv[k]) sh_fail = 1;
	}
	for (int i = threadIdx.x; i < n; i += blockDim.x) { const int r = i / 3; vg.delta[3 * orow[r] + (i - 3 * r)] = x[i]; }  // back to the reference's unknown order
	__syncthreads();
	const bool failed = sh_fail != 0;
	const int trial = s.trials;
	const int nsc = h->n_scalars;
	const double lambda = s.lambda;
	if (failed) {
		if (trial < h->trace_cap)
			for (int i = threadIdx.x; i < nsc; i += blockDim.x) vg.r_trace_delta[(size_t)trial * nsc + i] = 0;
		if (threadIdx.x == 0) {
			if (trial < h->trace_cap) {
				double *tr = vg.r_trace + (size_t)trial * 6;
				tr[0] = lambda; tr[1] = 0; tr[2] = 0; tr[3] = 0; tr[4] = 0; tr[5] = 0;
			}
			s.trials++; s.cholfail++;
			s.lambda *= s.nu;
			s.nu *= 2.;
			s.stop = (s.lambda > h->model.max_lambda);
			s.trial_ok = 0;
			s.chol_fail = 0;
			wave_advance(s, h);
			if (s.go) atomicAdd(a.active, 1);
		}
	} else if (threadIdx.x == 0) {
		s.trial_ok = 1;
		s.trial_lambda = lambda;
		s.chol_fail = 0;
	}
}

template <int P>
__global__ void __launch_bounds__(128) wave_features_kernel(const WaveArgs a) {
	const int area = blockIdx.y;
	const WaveArea &s = a.st[area];
	if (!s.go || !s.trial_ok) return;
	if (!wave_team(a.views[area].h->n_unk_lms, WAVE_CAP)) return;
	stage_features<P>(a.views[area]);
}

template <int P>
__global__ void __launch_bounds__(128) wave_update_kernel(const WaveArgs a) {
	const int area = blockIdx.y;
	const WaveArea &s = a.st[area];
	if (!s.go || !s.trial_ok) return;
	{
		const DevProblem *h = a.views[area].h;
		if (!wave_team(max(max((int64_t)h->n_unk_edges, (int64_t)h->n_unk_lms), (int64_t)s.n_listed * PoseOps<P>::S), WAVE_CAP)) return;
	}
	stage_update<P>(a.views[area], s.n_listed);
}

// ------------------------------------------------------------------------------------------------ control kernels (one CTA per area)
// After the set-up stages: lambda0 = tau * max diag (optimize_edges.h:365-390), initial error, first advance().
template <int P, int OBSK>
__global__ void __launch_bounds__(128) wave_ctrl_init_kernel(const WaveArgs a) {
	__shared__ double scr[8];
	const int area = blockIdx.y;
	const View v = a.views[area];
	const DevProblem *h = v.h;
	WaveArea &s = a.st[area];
	double dmax = 0;
	for (int b = threadIdx.x; b < h->n_HAp; b += blockDim.x)
		if (v.HAp_row[b] == v.HAp_col[b]) {
			double m = v.HAp[(size_t)b * P * P];
			for (int k = 1; k < P; k++) m = fmax(m, v.HAp[(size_t)b * P * P + k * P + k]);
			dmax = fmax(dmax, m);
		}
	for (int k = threadIdx.x; k < h->n_unk_lms; k += blockDim.x)
		dmax = fmax(dmax, fmax(v.Hf[(size_t)k * 9], fmax(v.Hf[(size_t)k * 9 + 4], v.Hf[(size_t)k * 9 + 8])));
	dmax = cta_max(dmax, scr);
	if (threadIdx.x == 0) {
		double err0 = 0;
		const double *pe = wave_part(a, area, WP_ERR);
		const int gx_resid = wave_team_ctas(h->n_obs, 256, WAVE_PART_CTAS);  // the residual kernel's team (256-thread CTAs)
		for (int i = 0; i < gx_resid; i++) err0 += pe[i];
		s.nu = 2;
		s.lambda = 1e-3 * dmax;
		s.err = err0;
		s.rmse = sqrt(err0 / h->n_total_obs_rows);
		s.iter = 0; s.stop = 0; s.trials = 0; s.accepted = 0; s.relin = 0; s.cholfail = 0; s.cur = 0;
		s.need_outer = 1; s.rho = 0; s.trial_ok = 0; s.accept = 0; s.do_relin = 0;
		v.stats->total_sqr_error_init = err0;
		if (!isfinite(err0) || !isfinite(dmax)) s.status = SRBA_ERR_NON_FINITE;
		wave_advance(s, h);
		if (s.go) atomicAdd(a.active, 1);
	}
}

// rho, accept / relinearise decisions (optimize_edges.h:585-605), trace
template <int P, int OBSK>
__global__ void __launch_bounds__(128) wave_ctrl1_kernel(const WaveArgs a) {
	const int area = blockIdx.y;
	WaveArea &s = a.st[area];
	if (!s.go || !s.trial_ok) return;
	const View v = a.views[area];
	const DevProblem *h = v.h;
	__shared__ int sh_trial;
	if (threadIdx.x == 0) {
		double new_err = 0, denom = 0;
		const double *pe = wave_part(a, area, WP_ERR), *pd = wave_part(a, area, WP_DENOM);
		const int gx_resid = wave_team_ctas(h->n_obs, 256, WAVE_PART_CTAS);
		for (int i = 0; i < gx_resid; i++) { new_err += pe[i]; denom += pd[i]; }
		const double err = s.err;
		const double rho = (err - new_err) / denom;
		const double error_reduction_ratio = err > 0 ? (err - new_err) / err : 0;
		const bool accept = rho > 0;
		const bool do_relinearize = accept && (error_reduction_ratio < 0 || error_reduction_ratio > h->model.min_error_reduction_ratio_to_relinearize);
		s.new_err = new_err; s.rho = rho; s.accept = accept ? 1 : 0; s.do_relin = do_relinearize ? 1 : 0;
		const int trial = s.trials;
		sh_trial = trial;
		if (trial < h->trace_cap) {
			double *tr = v.r_trace + (size_t)trial * 6;
			tr[0] = s.trial_lambda; tr[1] = 1; tr[2] = rho; tr[3] = new_err; tr[4] = accept ? 1 : 0; tr[5] = do_relinearize ? 1 : 0;
		}
	}
	__syncthreads();
	const int trial = sh_trial, nsc = h->n_scalars;
	if (trial < h->trace_cap)
		for (int i = threadIdx.x; i < nsc; i += blockDim.x) v.r_trace_delta[(size_t)trial * nsc + i] = v.delta[i];
}

// end of a trial: state update after an accepted (optimize_edges.h:633-656) or rejected (:658-690) step, next trial or done
template <int P, int OBSK>
__global__ void __launch_bounds__(32) wave_ctrl2_kernel(const WaveArgs a) {
	const int area = blockIdx.x * blockDim.x + threadIdx.x;
	if (area >= a.n) return;
	WaveArea &s = a.st[area];
	if (!s.go || !s.trial_ok) return;   // (a failed factorisation was already booked by the solve kernel)
	const DevProblem *h = a.views[area].h;
	const double rho = s.rho, new_err = s.new_err;
	if (s.accept) {
		double gm = 0;
		const double *pg = wave_part(a, area, WP_GMAX);
		const int gx_grad = wave_team_ctas(h->n_scalars, 128, WAVE_PART_CTAS);  // the gradient kernel's team (128-thread CTAs)
		for (int i = 0; i < gx_grad; i++) gm = i == 0 ? pg[0] : fmax(gm, pg[i]);
		s.cur ^= 1;
		s.err = new_err;
		s.rmse = sqrt(new_err / h->n_total_obs_rows);
		s.trials++; s.accepted++; s.relin += s.do_relin;
		if (gm <= 1e-15) s.stop = 1;  // max_gradient_to_stop (optimize_edges.h:362)
		if (s.rmse < h->model.max_error_per_obs_to_stop) s.stop = 1;
		if (rho > h->model.max_rho) s.stop = 1;
		s.lambda *= 1.0 / 3.0;
		s.nu = 2.0;
	} else {
		s.trials++;
		s.lambda *= s.nu;
		s.nu *= 2.0;
		s.stop = (s.lambda > h->model.max_lambda);
		if (!isfinite(new_err) || !isfinite(rho)) s.status = SRBA_ERR_NON_FINITE;
	}
	s.trial_ok = 0;
	wave_advance(s, h);
	if (s.go) atomicAdd(a.active, 1);
}

// Q/S. outputs
template <int P, int OBSK>
__global__ void __launch_bounds__(256) wave_output_kernel(const WaveArgs a) {
	constexpr int S = PoseOps<P>::S;
	const int area = blockIdx.y;
	const View v = a.views[area];
	const DevProblem *h = v.h;
	const WaveArea &s = a.st[area];
	if (!wave_team((int64_t)2 * h->n_pairs * S, WAVE_CAP)) return;
	const bool keep_grad = (OBSK != SRBA_OBS_RELPOSES2D) && no_schur_solver(h);
	for (size_t i = LM_TID; i < (size_t)h->n_unk_edges * S; i += LM_THREADS) v.r_edge_pose[i] = v.edge_cur[(size_t)v.unk_edge[i / S] * S + (i % S)];
	for (int i = LM_TID; i < h->n_unk_lms * 3; i += LM_THREADS) v.r_lm_pos[i] = v.lm_cur[(size_t)v.unk_lm[i / 3] * 3 + (i % 3)];
	for (size_t i = LM_TID; i < (size_t)2 * h->n_pairs * S; i += LM_THREADS) v.r_slot_pose[i] = v.slot[i];
	for (int i = LM_TID; i < h->n_obs; i += LM_THREADS) v.r_obs_valid[i] = v.valid[i];
	for (int i = LM_TID; i < h->n_unk_lms; i += LM_THREADS) v.r_lm_inv[i] = keep_grad ? 1 : (s.trials > 0 ? v.lm_inv[i] : 0);
	for (int i = LM_TID; i < h->n_unk_lms * 9; i += LM_THREADS) v.r_Hf_diag[i] = v.Hf[i];
	if (LM_TID == 0) {
		srba_stats_t *st = v.stats;
		st->num_iters = s.iter;
		st->num_trials = s.trials;
		st->num_accepted = s.accepted;
		st->num_relinearized = s.relin;
		st->num_chol_failures = s.cholfail;
		st->num_jacobians = h->n_blkA + ((OBSK == SRBA_OBS_RELPOSES2D) ? 0 : h->n_blkF);
		st->hessian_valid = s.trials > 0;
		st->status = s.status;
		st->total_sqr_error_final = s.err;
		st->obs_rmse = s.rmse;
		st->lambda_final = s.lambda;
		unsigned long long *sc = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(v.stats) - h->res_stats + h->res_stage_cycles);
#ifndef SRBA_WAVE_TIMERS
		for (int i = 0; i < 24; i++) sc[i] = 0;
#endif
	}
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host driver
// Per-stage maxima of the item counts over the areas of the group (the grid of a stage kernel is sized for the largest area;
// the CTAs of smaller areas beyond their items find their loops empty).
struct SrbaWaveDims {
	int n_areas;
	int max_pairs, max_blocks, max_hess_items, max_obs, max_scalars, max_unk, max_schur2, max_schur3, max_lms, max_edges_S, max_out;
	int has_lms, n_big, n_solo, max_tail_rows;  // n_big / n_solo: areas with levels below the tail (cluster team / one CTA); largest dense tail (tile rows)
	int max_bwd_smem;                           // shared memory of the backward kernel (x, row bounds, level bounds), largest area
};

// Optional CUDA-event timing of the dominant kernel (the cluster-team forward solver) for bench.py's roofline line: a pair of events around
// every launch of it, read after the run.  Off by default (srba_b200_kernel_timing).
static int g_wave_time_kernels = 0;
static double g_wave_fwd_team_ms = 0.0;
static long g_wave_fwd_team_launches = 0;
extern "C" void srba_wave_timing_enable(int on) { g_wave_time_kernels = on; g_wave_fwd_team_ms = 0.0; g_wave_fwd_team_launches = 0; }
extern "C" void srba_wave_timing_get(double *ms, long *launches) { *ms = g_wave_fwd_team_ms; *launches = g_wave_fwd_team_launches; }

namespace {
inline int wave_gx(int items, int threads, int cap = WAVE_CAP) { return std::max(1, std::min((items + threads - 1) / threads, cap)); }

template <int P, int OBSK>
int wave_run(const WaveArgs &a, const SrbaWaveDims &d, int *h_active_pinned, cudaStream_t st, int64_t *launches) {
	constexpr bool can_fail = (OBSK == SRBA_OBS_MONO || OBSK == SRBA_OBS_STEREO);
	const int n = d.n_areas;
	auto grid = [&](int gx) { return dim3((unsigned)gx, (unsigned)n, 1); };
	const int gx_chain = wave_gx(d.max_pairs, 256), gx_jac = wave_gx(d.max_blocks, 128), gx_hess = wave_gx(d.max_hess_items, 128);
	const int gx_resid = wave_gx(d.max_obs, 256, WAVE_PART_CTAS), gx_grad = wave_gx(d.max_scalars, 128, WAVE_PART_CTAS), gx_upd = wave_gx(d.max_unk, 128);
	const int gx_s1 = wave_gx(d.max_lms, 128), gx_s2 = wave_gx(d.max_schur2, 128), gx_s3 = wave_gx(d.max_schur3, 128);
	const int gx_init = wave_gx(d.max_edges_S, 256), gx_out = wave_gx(d.max_out, 256);
	int64_t nl = 0;
	cudaLaunchConfig_t scfg = {};  // forward / backward levels below the tail: one cluster of WAVE_SOLVE_CTAS CTAs per area that has such levels
	scfg.gridDim = dim3(WAVE_SOLVE_CTAS, (unsigned)std::max(1, d.n_big), 1);
	scfg.blockDim = dim3(WAVE_SOLVE_NT, 1, 1);
	scfg.dynamicSmemBytes = 0;
	scfg.stream = st;
	cudaLaunchAttribute sattr[1];
	sattr[0].id = cudaLaunchAttributeClusterDimension;
	sattr[0].val.clusterDim.x = WAVE_SOLVE_CTAS; sattr[0].val.clusterDim.y = 1; sattr[0].val.clusterDim.z = 1;
	scfg.attrs = sattr;
	scfg.numAttrs = 1;
	auto solve_fwd = wave_solve_fwd_kernel<P, OBSK, true>;
	if (WAVE_SOLVE_CTAS > 8 && cudaFuncSetAttribute(solve_fwd, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return (int)cudaGetLastError();
	auto solve_tail = wave_solve_tail_kernel<P, OBSK>;
	const size_t tail_smem = wave_tail_smem_bytes(std::max(1, d.max_tail_rows));
	auto solve_bwd = wave_solve_bwd_kernel<P, OBSK>;
	const size_t bwd_smem = (size_t)std::min(std::max(d.max_bwd_smem, 16), 200 * 1024);   // larger areas run the backward levels on the global arrays
	if (cudaFuncSetAttribute(solve_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem) != cudaSuccess) return (int)cudaGetLastError();
	if (cudaFuncSetAttribute(solve_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tail_smem) != cudaSuccess) return (int)cudaGetLastError();
	cudaMemsetAsync(a.big + n, 0, sizeof(int), st);
	cudaMemsetAsync(a.solo + n, 0, sizeof(int), st);
	const bool both = d.n_big > 0 && d.n_solo > 0;
	const bool timing = g_wave_time_kernels != 0;
	std::vector<cudaEvent_t> tev;
	cudaStream_t side = st;
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
	if (both) {
		if (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess) return (int)cudaGetLastError();
		cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
		cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming);
		scfg.stream = side;
	}
	cudaMemsetAsync(a.active, 0, sizeof(int), st);
	// ---- set-up: E (numeric spanning tree), F/G (Jacobians), lists, J (Hessian), M (residuals), N (gradient), L (lambda0)
	wave_views_kernel<<<(n + 127) / 128, 128, 0, st>>>(a);
	static const bool dbg = std::getenv("SRBA_WAVE_DEBUG") != nullptr;
	if (dbg) { wave_dump_kernel<<<1, 1, 0, st>>>(a, 0); cudaStreamSynchronize(st); fprintf(stderr, "[wave] after views: %s\n", cudaGetErrorString(cudaGetLastError())); }
	wave_init_kernel<P><<<grid(gx_init), 256, 0, st>>>(a);
	if (dbg) { wave_dump_kernel<<<1, 1, 0, st>>>(a, 1); cudaStreamSynchronize(st); fprintf(stderr, "[wave] after init: %s\n", cudaGetErrorString(cudaGetLastError())); }
	wave_chain_kernel<P><<<grid(gx_chain), 256, 0, st>>>(a, WHEN_SETUP);
	if (dbg) { wave_dump_kernel<<<1, 1, 0, st>>>(a, 2); cudaStreamSynchronize(st); fprintf(stderr, "[wave] after chain: %s\n", cudaGetErrorString(cudaGetLastError())); }
	wave_jacobians_kernel<P, OBSK><<<grid(gx_jac), 128, 0, st>>>(a, WHEN_SETUP, 0);
	wave_jacobians_kernel<P, OBSK><<<grid(gx_jac), 128, 0, st>>>(a, WHEN_SETUP, 1);
	wave_lists_kernel<<<grid(1), 64, 0, st>>>(a);
	wave_hessian_kernel<P, OBSK><<<grid(gx_hess), 128, 0, st>>>(a, WHEN_SETUP);
	wave_residuals_kernel<P, OBSK><<<grid(gx_resid), 256, 0, st>>>(a, WHEN_SETUP);
	wave_gradient_or_restore_kernel<P, OBSK><<<grid(gx_grad), 128, 0, st>>>(a, 1);
	wave_ctrl_init_kernel<P, OBSK><<<grid(1), 128, 0, st>>>(a);
	nl += 10;
	// ---- trials, all areas in lock step.  The count of areas that still iterate is read back with one trial of lag, so the
	// next trial is already queued while the host waits (a trial queued after the last area finished finds every area idle).
	cudaEvent_t ev[2];
	cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
	cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
	cudaMemcpyAsync(&h_active_pinned[0], a.active, sizeof(int), cudaMemcpyDeviceToHost, st);
	cudaEventRecord(ev[0], st);
	int rc = 0;
	for (int t = 0;; t++) {
		cudaMemsetAsync(a.active, 0, sizeof(int), st);
		if (d.has_lms) {
			wave_schur_kernel<P, OBSK><<<grid(gx_s1), 128, 0, st>>>(a, 1);
			wave_schur_kernel<P, OBSK><<<grid(gx_s2), 128, 0, st>>>(a, 2);
			wave_schur_kernel<P, OBSK><<<grid(gx_s3), 128, 0, st>>>(a, 3);
			nl += 3;
		}
		// the cluster teams of the few wide areas run on a side stream next to the one-CTA areas (fork / join per phase)
		if (both) { cudaEventRecord(ev_fork, st); cudaStreamWaitEvent(side, ev_fork, 0); }
		if (d.n_solo) { wave_solve_fwd_kernel<P, OBSK, false><<<dim3(1, (unsigned)d.n_solo, 1), WAVE_SOLO_NT, 0, st>>>(a); nl++; }
		if (d.n_big) {
			cudaEvent_t t0 = nullptr, t1 = nullptr;
			if (timing) { cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventRecord(t0, scfg.stream); }
			const cudaError_t e = cudaLaunchKernelEx(&scfg, solve_fwd, a);
			if (e != cudaSuccess) { rc = (int)e; break; }
			nl++;
			if (timing) { cudaEventRecord(t1, scfg.stream); tev.push_back(t0); tev.push_back(t1); }
		}
		if (both) { cudaEventRecord(ev_join, side); cudaStreamWaitEvent(st, ev_join, 0); }
		solve_tail<<<n, WAVE_TAIL_NT, tail_smem, st>>>(a);
		if (both) { cudaEventRecord(ev_fork, st); cudaStreamWaitEvent(side, ev_fork, 0); }
		// The back-substitution has at most a handful of rows per level (one warp each): a cluster only adds barrier latency there, so the
		// wide areas run it with ONE CTA as well (CTA barriers, all of them resident at once) -- on the side stream, next to the solo areas.
		if (d.n_solo) { solve_bwd<<<dim3(1, (unsigned)d.n_solo, 1), WAVE_SOLO_NT, bwd_smem, st>>>(a, 0, (int)bwd_smem); nl++; }
		if (d.n_big) { solve_bwd<<<dim3(1, (unsigned)d.n_big, 1), WAVE_SOLO_NT, bwd_smem, scfg.stream>>>(a, 1, (int)bwd_smem); nl++; }
		if (both) { cudaEventRecord(ev_join, side); cudaStreamWaitEvent(st, ev_join, 0); }
		wave_solve_finish_kernel<P, OBSK><<<n, 256, 0, st>>>(a);
		nl++;
		if (d.has_lms) { wave_features_kernel<P><<<grid(gx_s1), 128, 0, st>>>(a); nl++; }
		wave_update_kernel<P><<<grid(gx_upd), 128, 0, st>>>(a);
		wave_chain_kernel<P><<<grid(gx_chain), 256, 0, st>>>(a, WHEN_TRIAL_OK);
		wave_residuals_kernel<P, OBSK><<<grid(gx_resid), 256, 0, st>>>(a, WHEN_TRIAL_OK);
		wave_ctrl1_kernel<P, OBSK><<<grid(1), 128, 0, st>>>(a);
		wave_jacobians_kernel<P, OBSK><<<grid(gx_jac), 128, 0, st>>>(a, WHEN_RELIN, 0);
		if (can_fail) { wave_jacobians_kernel<P, OBSK><<<grid(gx_jac), 128, 0, st>>>(a, WHEN_RELIN, 1); nl++; }
		wave_hessian_kernel<P, OBSK><<<grid(gx_hess), 128, 0, st>>>(a, WHEN_RELIN);
		wave_gradient_or_restore_kernel<P, OBSK><<<grid(gx_grad), 128, 0, st>>>(a, 0);
		wave_ctrl2_kernel<P, OBSK><<<(n + 31) / 32, 32, 0, st>>>(a);
		nl += 9;
		cudaMemcpyAsync(&h_active_pinned[(t + 1) & 1], a.active, sizeof(int), cudaMemcpyDeviceToHost, st);
		cudaEventRecord(ev[(t + 1) & 1], st);
		// wait for the count that PRECEDES the trial just queued
		if (cudaEventSynchronize(ev[t & 1]) != cudaSuccess) { rc = (int)cudaGetLastError(); break; }
		if (h_active_pinned[t & 1] == 0) break;  // nobody was iterating when trial t was queued: it was idle, done
		if (t > 100000) { rc = (int)cudaErrorUnknown; break; }
	}
	wave_output_kernel<P, OBSK><<<grid(gx_out), 256, 0, st>>>(a);
	nl++;
	cudaEventDestroy(ev[0]);
	cudaEventDestroy(ev[1]);
	if (timing) {
		cudaStreamSynchronize(st);
		if (both) cudaStreamSynchronize(side);
		for (size_t i = 0; i + 1 < tev.size(); i += 2) {
			float ms = 0;
			if (cudaEventElapsedTime(&ms, tev[i], tev[i + 1]) == cudaSuccess) { g_wave_fwd_team_ms += ms; g_wave_fwd_team_launches++; }
			cudaEventDestroy(tev[i]); cudaEventDestroy(tev[i + 1]);
		}
	}
	if (both) {  // (every side-stream launch was joined back into `st`)
		cudaStreamSynchronize(side);
		cudaEventDestroy(ev_fork);
		cudaEventDestroy(ev_join);
		cudaStreamDestroy(side);
	}
	if (launches) *launches += nl;
	if (rc) return rc;
	return (int)cudaGetLastError();
}
}  // namespace

// model: 0 = SE2+RelativePoses2D, 1 = SE3+Cartesian3D, 2 = SE3+Mono, 3 = SE3+Stereo.  scratch: device memory of
// srba_wave_scratch_bytes(n) bytes; h_active_pinned: 2 ints of pinned host memory.  Blocks the host until every area is done.
extern "C" size_t srba_wave_scratch_bytes(int n_areas) {
	auto al = [](size_t x) { return (x + 255) / 256 * 256; };
	return al(sizeof(View) * (size_t)n_areas) + al(sizeof(WaveArea) * (size_t)n_areas) + al(sizeof(double) * (size_t)n_areas * WP_KINDS * WAVE_PART_CTAS) + 256 +
	       2 * al(sizeof(int) * ((size_t)n_areas + 1));
}
extern "C" int srba_launch_lm_wave(int model, const char *img, const uint64_t *img_off, char *ws, const uint64_t *ws_off, char *res, const uint64_t *res_off,
                                   const int32_t *order, const SrbaWaveDims *dims, char *scratch, int *h_active_pinned, void *stream, int64_t *launches) {
	auto al = [](size_t x) { return (x + 255) / 256 * 256; };
	WaveArgs a;
	a.img = img; a.img_off = img_off; a.ws = ws; a.ws_off = ws_off; a.res = res; a.res_off = res_off; a.order = order; a.n = dims->n_areas;
	size_t off = 0;
	a.views = reinterpret_cast<View *>(scratch + off); off += al(sizeof(View) * (size_t)a.n);
	a.st = reinterpret_cast<WaveArea *>(scratch + off); off += al(sizeof(WaveArea) * (size_t)a.n);
	a.partials = reinterpret_cast<double *>(scratch + off); off += al(sizeof(double) * (size_t)a.n * WP_KINDS * WAVE_PART_CTAS);
	a.active = reinterpret_cast<int *>(scratch + off); off += 256;
	a.big = reinterpret_cast<int *>(scratch + off); off += al(sizeof(int) * ((size_t)a.n + 1));
	a.solo = reinterpret_cast<int *>(scratch + off);
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	switch (model) {
		case 0: return wave_run<3, SRBA_OBS_RELPOSES2D>(a, *dims, h_active_pinned, st, launches);
		case 1: return wave_run<6, SRBA_OBS_CARTESIAN3D>(a, *dims, h_active_pinned, st, launches);
		case 2: return wave_run<6, SRBA_OBS_MONO>(a, *dims, h_active_pinned, st, launches);
		case 3: return wave_run<6, SRBA_OBS_STEREO>(a, *dims, h_active_pinned, st, launches);
	}
	return (int)cudaErrorInvalidValue;
}
