// C ABI of the B200 hot path (include/srba_b200.h): packs flattened optimize_edges() problems into one pinned host
// image, moves it to HBM with ONE copy, launches the persistent LM kernel (lm_kernel.cu) and brings one packed result
// image back.  No CPU fallback: without a CUDA device every entry point fails with SRBA_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>

#include <pthread.h>
#include <sched.h>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "dev_types.h"

extern "C" int srba_launch_eval_error(const srba_model_t *model, int n_pairs, int n_obs, const double *edge_pose, const int32_t *pair_path_ptr,
                                      const int32_t *path_edge, const uint8_t *path_dir, const double *obs_z, const int32_t *obs_pair,
                                      const uint8_t *obs_inverse, const int32_t *obs_lm, const double *lm_pos, double *pair_pose, double *partial,
                                      double *total, int grid_chain, int grid_resid, void *stream);
extern "C" int srba_launch_lm_batch(int model, const char *img, const uint64_t *img_off, char *ws, const uint64_t *ws_off, char *res,
                                    const uint64_t *res_off, const int32_t *order, int n_problems, int grid, int smem_bytes, void *stream);

// grid-wide lock-step pipeline for heavy areas (lm_wave.cuh)
struct SrbaWaveDims {
	int n_areas;
	int max_pairs, max_blocks, max_hess_items, max_obs, max_scalars, max_unk, max_schur2, max_schur3, max_lms, max_edges_S, max_out;
	int has_lms, n_big, n_solo, max_tail_rows;
	int max_bwd_smem;
};
extern "C" size_t srba_wave_scratch_bytes(int n_areas);
extern "C" int srba_launch_lm_wave(int model, const char *img, const uint64_t *img_off, char *ws, const uint64_t *ws_off, char *res, const uint64_t *res_off,
                                   const int32_t *order, const SrbaWaveDims *dims, char *scratch, int *h_active_pinned, void *stream, int64_t *launches);
extern "C" void srba_wave_timing_enable(int on);
extern "C" void srba_wave_timing_get(double *ms, long *launches);

namespace {
thread_local std::string g_err;
thread_local int g_code = SRBA_OK;

int fail(int code, const std::string &msg) {
	g_err = msg;
	g_code = code;
	return code;
}
#define CUDA_TRY(x)                                                                                           \
	do {                                                                                                      \
		cudaError_t e_ = (x);                                                                                 \
		if (e_ != cudaSuccess) return fail(SRBA_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_));   \
	} while (0)

struct Layout {
	uint64_t cur = 0;
	uint64_t add(uint64_t bytes) {
		cur = (cur + 15) / 16 * 16;
		const uint64_t o = cur;
		cur += bytes;
		return o;
	}
	uint64_t total() { return (cur + 15) / 16 * 16; }
};

int obs_dims(int k) { return k == SRBA_OBS_MONO ? 2 : (k == SRBA_OBS_STEREO ? 4 : 3); }

// Run-time execution options (srba_b200_set_option): test / experiment hooks, initialised once from the environment.
//   cluster  -1 auto | 0 never | 1 always : 8-CTA cluster team per heavy area          (env SRBA_B200_CLUSTER)
//   blocked   0 auto | 1 : dense blocked Cholesky instead of the block-sparse L D L^t   (env SRBA_B200_SOLVER=blocked)
//   wide      0 auto | 1 : 3-int update descriptors                                     (env SRBA_B200_WIDE)
//   chunks    1..8 pipelined chunks of srba_b200_optimize_batch                          (env SRBA_B200_CHUNKS, default 4)
//   wave     -1 auto | 0 never | 1 always : grid-wide lock-step pipeline (lm_wave.cuh) for areas that overflow one CTA's arena
enum { OPT_CLUSTER = 0, OPT_BLOCKED, OPT_WIDE, OPT_CHUNKS, OPT_WAVE, OPT_COUNT };
const char *const k_opt_names[OPT_COUNT] = {"cluster", "blocked", "wide", "chunks", "wave"};
std::atomic<int> g_opt[OPT_COUNT];
std::once_flag g_opt_once;
void init_options() {
	std::call_once(g_opt_once, [] {
		const char *e;
		g_opt[OPT_CLUSTER].store((e = std::getenv("SRBA_B200_CLUSTER")) ? std::atoi(e) : -1);
		g_opt[OPT_BLOCKED].store(((e = std::getenv("SRBA_B200_SOLVER")) && std::string(e) == "blocked") ? 1 : 0);
		g_opt[OPT_WIDE].store(std::getenv("SRBA_B200_WIDE") ? 1 : 0);
		g_opt[OPT_CHUNKS].store((e = std::getenv("SRBA_B200_CHUNKS")) ? std::max(1, std::min(std::atoi(e), 8)) : 4);
		g_opt[OPT_WAVE].store((e = std::getenv("SRBA_B200_WAVE")) ? std::atoi(e) : -1);
	});
}
int option(int id) { init_options(); return g_opt[id].load(std::memory_order_relaxed); }

// Structural validation of one caller-owned problem: the packer, the symbolic factorisation and the kernel all index
// with these arrays, so a malformed problem must be rejected here, not found as a fault of the whole batch.  O(size);
// returns NULL when the problem is well formed, else a static message (callable from worker threads).
const char *validate_structure(const srba_problem_t &p) {
	const int32_t counts[] = {p.n_edges, p.n_unk_edges, p.n_lms, p.n_unk_lms, p.n_pairs, p.n_path, p.n_obs, p.n_blkA, p.n_blkF, p.n_HAp, p.n_HAp_terms,
	                          p.n_HApf, p.n_HApf_terms, p.n_S, p.n_S_terms, p.n_total_obs_rows};
	for (int32_t c : counts)
		if (c < 0) return "negative count";
	if ((int64_t)2 * p.n_pairs > INT32_MAX) return "too many pose pairs";
	struct Arr { const void *ptr; int64_t n; };
	const Arr arrs[] = {{p.edge_pose, p.n_edges}, {p.unk_edge, p.n_unk_edges}, {p.lm_pos, p.n_lms}, {p.unk_lm, p.n_unk_lms}, {p.pair_path_ptr, p.n_pairs + 1},
	                    {p.path_edge, p.n_path}, {p.path_dir, p.n_path}, {p.obs_z, p.n_obs}, {p.obs_slot, p.n_obs}, {p.obs_lm, p.n_obs}, {p.obs_mult, p.n_obs},
	                    {p.colA_ptr, p.n_unk_edges + 1}, {p.blkA_obs, p.n_blkA}, {p.blkA_slotA, p.n_blkA}, {p.blkA_slotD, p.n_blkA}, {p.blkA_normal_dir, p.n_blkA},
	                    {p.colF_ptr, p.n_unk_lms + 1}, {p.blkF_obs, p.n_blkF}, {p.blkF_slot, p.n_blkF}, {p.row_ptr, p.n_obs + 1}, {p.row_blk, p.n_blkA + p.n_blkF},
	                    {p.HAp_row, p.n_HAp}, {p.HAp_col, p.n_HAp}, {p.HAp_term_ptr, p.n_HAp + 1}, {p.HAp_term_a, p.n_HAp_terms}, {p.HAp_term_b, p.n_HAp_terms},
	                    {p.HApf_row, p.n_HApf}, {p.HApf_col, p.n_HApf}, {p.HApf_term_ptr, p.n_HApf + 1}, {p.HApf_term_a, p.n_HApf_terms}, {p.HApf_term_f, p.n_HApf_terms},
	                    {p.rowApf_ptr, p.n_unk_edges + 1}, {p.colApf_ptr, p.n_unk_lms + 1}, {p.colApf_idx, p.n_HApf},
	                    {p.S_row, p.n_S}, {p.S_col, p.n_S}, {p.S_HAp, p.n_S}, {p.S_term_ptr, p.n_S + 1}, {p.S_term_rowblk, p.n_S_terms}, {p.S_term_colblk, p.n_S_terms}};
	for (const Arr &a : arrs)
		if (a.n > 0 && !a.ptr) return "NULL array with a non-zero count";
	auto csr = [](const int32_t *ptr, int n, int total) {
		if (ptr[0] != 0 || ptr[n] != total) return false;
		for (int i = 0; i < n; i++)
			if (ptr[i] > ptr[i + 1]) return false;
		return true;
	};
	auto in_range = [](const int32_t *a, int n, int lo, int hi) {  // every a[i] in [lo, hi)
		for (int i = 0; i < n; i++)
			if (a[i] < lo || a[i] >= hi) return false;
		return true;
	};
	const int n_slots = 2 * p.n_pairs, n_blk = p.n_blkA + p.n_blkF;
	if (!csr(p.pair_path_ptr, p.n_pairs, p.n_path)) return "pair_path_ptr is not a CSR row pointer over n_path";
	if (!csr(p.colA_ptr, p.n_unk_edges, p.n_blkA)) return "colA_ptr is not a CSR row pointer over n_blkA";
	if (!csr(p.colF_ptr, p.n_unk_lms, p.n_blkF)) return "colF_ptr is not a CSR row pointer over n_blkF";
	if (!csr(p.row_ptr, p.n_obs, n_blk)) return "row_ptr is not a CSR row pointer over n_blkA + n_blkF";
	if (!csr(p.HAp_term_ptr, p.n_HAp, p.n_HAp_terms)) return "HAp_term_ptr is not a CSR row pointer over n_HAp_terms";
	if (!csr(p.HApf_term_ptr, p.n_HApf, p.n_HApf_terms)) return "HApf_term_ptr is not a CSR row pointer over n_HApf_terms";
	if (!csr(p.rowApf_ptr, p.n_unk_edges, p.n_HApf)) return "rowApf_ptr is not a CSR row pointer over n_HApf";
	if (!csr(p.colApf_ptr, p.n_unk_lms, p.n_HApf)) return "colApf_ptr is not a CSR row pointer over n_HApf";
	if (!csr(p.S_term_ptr, p.n_S, p.n_S_terms)) return "S_term_ptr is not a CSR row pointer over n_S_terms";
	if (!in_range(p.unk_edge, p.n_unk_edges, 0, p.n_edges)) return "unk_edge out of range";
	if (!in_range(p.unk_lm, p.n_unk_lms, 0, p.n_lms)) return "unk_lm out of range";
	if (!in_range(p.path_edge, p.n_path, 0, p.n_edges)) return "path_edge out of range";
	if (!in_range(p.obs_slot, p.n_obs, -1, n_slots)) return "obs_slot out of range";
	if (p.model.pose_kind == SRBA_POSE_SE3 && !in_range(p.obs_lm, p.n_obs, 0, p.n_lms)) return "obs_lm out of range";
	if (!in_range(p.obs_mult, p.n_obs, 0, INT32_MAX)) return "obs_mult negative";
	if (!in_range(p.blkA_obs, p.n_blkA, 0, p.n_obs)) return "blkA_obs out of range";
	if (!in_range(p.blkA_slotA, p.n_blkA, -1, n_slots)) return "blkA_slotA out of range";
	if (!in_range(p.blkA_slotD, p.n_blkA, 0, n_slots)) return "blkA_slotD out of range";
	if (!in_range(p.blkF_obs, p.n_blkF, 0, p.n_obs)) return "blkF_obs out of range";
	if (!in_range(p.blkF_slot, p.n_blkF, -1, n_slots)) return "blkF_slot out of range";
	if (!in_range(p.row_blk, n_blk, 0, n_blk)) return "row_blk out of range";
	if (!in_range(p.HAp_row, p.n_HAp, 0, p.n_unk_edges) || !in_range(p.HAp_col, p.n_HAp, 0, p.n_unk_edges)) return "HAp_row / HAp_col out of range";
	if (!in_range(p.HAp_term_a, p.n_HAp_terms, 0, p.n_blkA) || !in_range(p.HAp_term_b, p.n_HAp_terms, 0, p.n_blkA)) return "HAp_term_a / HAp_term_b out of range";
	if (!in_range(p.HApf_row, p.n_HApf, 0, p.n_unk_edges) || !in_range(p.HApf_col, p.n_HApf, 0, p.n_unk_lms)) return "HApf_row / HApf_col out of range";
	if (!in_range(p.HApf_term_a, p.n_HApf_terms, 0, p.n_blkA) || !in_range(p.HApf_term_f, p.n_HApf_terms, 0, p.n_blkF)) return "HApf_term_a / HApf_term_f out of range";
	if (!in_range(p.colApf_idx, p.n_HApf, 0, p.n_HApf)) return "colApf_idx out of range";
	if (!in_range(p.S_row, p.n_S, 0, p.n_unk_edges) || !in_range(p.S_col, p.n_S, 0, p.n_unk_edges)) return "S_row / S_col out of range";
	if (!in_range(p.S_HAp, p.n_S, -1, p.n_HAp)) return "S_HAp out of range";
	if (!in_range(p.S_term_rowblk, p.n_S_terms, 0, p.n_HApf) || !in_range(p.S_term_colblk, p.n_S_terms, 0, p.n_HApf)) return "S_term_rowblk / S_term_colblk out of range";
	for (int b = 0; b < p.n_S; b++)
		if (p.S_row[b] > p.S_col[b]) return "reduced-system block below the diagonal (row > col)";
	for (int b = 0; b < p.n_HAp; b++)
		if (p.HAp_row[b] > p.HAp_col[b]) return "H_Ap block below the diagonal (row > col)";
	return nullptr;
}

int validate(const srba_problem_t &p, bool deep = true) {
	const srba_model_t &m = p.model;
	if (m.pose_kind == SRBA_POSE_SE2) {
		if (m.obs_kind != SRBA_OBS_RELPOSES2D || m.sensor_pose_kind != SRBA_SENSORPOSE_NONE)
			return fail(SRBA_ERR_UNSUPPORTED_MODEL, "SE2 is only built with RelativePoses_2D observations and no sensor offset");
		if (p.n_unk_lms != 0) return fail(SRBA_ERR_UNSUPPORTED_MODEL, "relative graph-SLAM has no landmark unknowns (SURVEY.md §3.6)");
	} else if (m.pose_kind == SRBA_POSE_SE3) {
		if (m.obs_kind != SRBA_OBS_CARTESIAN3D && m.obs_kind != SRBA_OBS_MONO && m.obs_kind != SRBA_OBS_STEREO)
			return fail(SRBA_ERR_UNSUPPORTED_MODEL, "SE3 is built with Cartesian_3D / MonocularCamera / StereoCamera observations");
	} else
		return fail(SRBA_ERR_UNSUPPORTED_MODEL, "unknown pose_kind");
	if (p.n_unk_edges <= 0)
		return fail(SRBA_ERR_BAD_ARGUMENT, "n_unk_edges must be >0 (the reference dereferences an empty Schur table when only landmarks are optimised, schur.h:34,288)");
	if (p.n_obs <= 0 || p.n_blkA <= 0) return fail(SRBA_ERR_BAD_ARGUMENT, "empty problem");
	if (m.sensor_pose_kind != SRBA_SENSORPOSE_NONE && m.sensor_pose_kind != SRBA_SENSORPOSE_SE3) return fail(SRBA_ERR_BAD_ARGUMENT, "unknown sensor_pose_kind");
	if (m.noise_kind != SRBA_NOISE_IDENTITY && m.noise_kind != SRBA_NOISE_CONST_MATRIX) return fail(SRBA_ERR_BAD_ARGUMENT, "unknown noise_kind");
	const int P = m.pose_kind == SRBA_POSE_SE2 ? 3 : 6;
	const int o = obs_dims(m.obs_kind);
	if (m.noise_kind == SRBA_NOISE_IDENTITY && !(m.noise_std > 0)) return fail(SRBA_ERR_BAD_ARGUMENT, "noise_std must be > 0");
	if (m.solver_kind < SRBA_SOLVER_SCHUR_DENSE_CHOLESKY || m.solver_kind > SRBA_SOLVER_NO_SCHUR_SPARSE_CHOLESKY)
		return fail(SRBA_ERR_BAD_ARGUMENT, "solver_kind must be one of SRBA_SOLVER_* (srba_options_solver.h:21-77)");
	if (deep) {
		if (const char *msg = validate_structure(p)) return fail(SRBA_ERR_BAD_ARGUMENT, msg);
	} else {
		if (p.n_total_obs_rows < 0 || p.n_unk_lms < 0) return fail(SRBA_ERR_BAD_ARGUMENT, "negative count");
	}
	// ASSERT_ABOVEEQ_(OBS_DIMS*nObs, POSE_DIMS*nUnknowns_k2k+LM_DIMS*nUnknowns_k2f)  optimize_edges.h:355
	if ((int64_t)o * p.n_total_obs_rows < (int64_t)P * p.n_unk_edges + 3LL * p.n_unk_lms)
		return fail(SRBA_ERR_RANK_DEFICIENT, "OBS_DIMS*nObs < number of scalar unknowns (optimize_edges.h:355)");
	return SRBA_OK;
}

template <class T>
void put(char *base, uint64_t off, const T *src, size_t n) {
	if (n) std::memcpy(base + off, src, n * sizeof(T));
}

inline const void *src_of_edge_pose(const srba_problem_t &p) { return p.edge_pose; }
inline const void *src_of_unk_edge(const srba_problem_t &p) { return p.unk_edge; }
inline const void *src_of_lm_pos(const srba_problem_t &p) { return p.lm_pos; }
inline const void *src_of_unk_lm(const srba_problem_t &p) { return p.unk_lm; }
inline const void *src_of_pair_path_ptr(const srba_problem_t &p) { return p.pair_path_ptr; }
inline const void *src_of_path_edge(const srba_problem_t &p) { return p.path_edge; }
inline const void *src_of_path_dir(const srba_problem_t &p) { return p.path_dir; }
inline const void *src_of_obs_z(const srba_problem_t &p) { return p.obs_z; }
inline const void *src_of_obs_slot(const srba_problem_t &p) { return p.obs_slot; }
inline const void *src_of_obs_lm(const srba_problem_t &p) { return p.obs_lm; }
inline const void *src_of_obs_mult(const srba_problem_t &p) { return p.obs_mult; }
inline const void *src_of_colA_ptr(const srba_problem_t &p) { return p.colA_ptr; }
inline const void *src_of_blkA_obs(const srba_problem_t &p) { return p.blkA_obs; }
inline const void *src_of_blkA_slotA(const srba_problem_t &p) { return p.blkA_slotA; }
inline const void *src_of_blkA_slotD(const srba_problem_t &p) { return p.blkA_slotD; }
inline const void *src_of_blkA_normal_dir(const srba_problem_t &p) { return p.blkA_normal_dir; }
inline const void *src_of_colF_ptr(const srba_problem_t &p) { return p.colF_ptr; }
inline const void *src_of_blkF_obs(const srba_problem_t &p) { return p.blkF_obs; }
inline const void *src_of_blkF_slot(const srba_problem_t &p) { return p.blkF_slot; }
inline const void *src_of_row_ptr(const srba_problem_t &p) { return p.row_ptr; }
inline const void *src_of_row_blk(const srba_problem_t &p) { return p.row_blk; }
inline const void *src_of_HAp_row(const srba_problem_t &p) { return p.HAp_row; }
inline const void *src_of_HAp_col(const srba_problem_t &p) { return p.HAp_col; }
inline const void *src_of_HAp_term_ptr(const srba_problem_t &p) { return p.HAp_term_ptr; }
inline const void *src_of_HAp_term_a(const srba_problem_t &p) { return p.HAp_term_a; }
inline const void *src_of_HAp_term_b(const srba_problem_t &p) { return p.HAp_term_b; }
inline const void *src_of_HApf_row(const srba_problem_t &p) { return p.HApf_row; }
inline const void *src_of_HApf_col(const srba_problem_t &p) { return p.HApf_col; }
inline const void *src_of_HApf_term_ptr(const srba_problem_t &p) { return p.HApf_term_ptr; }
inline const void *src_of_HApf_term_a(const srba_problem_t &p) { return p.HApf_term_a; }
inline const void *src_of_HApf_term_f(const srba_problem_t &p) { return p.HApf_term_f; }
inline const void *src_of_rowApf_ptr(const srba_problem_t &p) { return p.rowApf_ptr; }
inline const void *src_of_colApf_ptr(const srba_problem_t &p) { return p.colApf_ptr; }
inline const void *src_of_colApf_idx(const srba_problem_t &p) { return p.colApf_idx; }
inline const void *src_of_S_row(const srba_problem_t &p) { return p.S_row; }
inline const void *src_of_S_col(const srba_problem_t &p) { return p.S_col; }
inline const void *src_of_S_HAp(const srba_problem_t &p) { return p.S_HAp; }
inline const void *src_of_S_term_ptr(const srba_problem_t &p) { return p.S_term_ptr; }
inline const void *src_of_S_term_rowblk(const srba_problem_t &p) { return p.S_term_rowblk; }
inline const void *src_of_S_term_colblk(const srba_problem_t &p) { return p.S_term_colblk; }
inline const void *src_of_blkA_col(const srba_problem_t &) { return nullptr; }
inline const void *src_of_blkF_col(const srba_problem_t &) { return nullptr; }
inline const void *src_of_ldl_row(const srba_problem_t &) { return nullptr; }
inline const void *src_of_ldl_tile_ij(const srba_problem_t &) { return nullptr; }
inline const void *src_of_ldl_src(const srba_problem_t &) { return nullptr; }
inline const void *src_of_ldl_orow(const srba_problem_t &) { return nullptr; }
inline const void *src_of_ldl_phase(const srba_problem_t &) { return nullptr; }
inline const void *src_of_ldl_upd(const srba_problem_t &) { return nullptr; }
// pointer to the caller's array behind image slot `id` (NULL for the derived ones)
const void *src_ptr(const srba_problem_t &p, int id) {
	switch (id) {
#define X(T, name) case SRBA_I_##name: return src_of_##name(p);
		SRBA_IMG_ARRAYS(X)
#undef X
	}
	return nullptr;
}

// Symbolic factorisation of the reduced system at 3x3-tile granularity (SE(2) block = 1 tile, SE(3) block = 2x2 tiles):
// filled upper pattern of U (no reordering: the reference's unknown order) + the right-hand side as tile column NTR, the
// tile numbering and the per-row update schedule consumed by stage_ldl_solve_smem (lm_kernel.cu).
struct TileSymbolic {
	int ntr = 0;
	std::vector<int32_t> rs;        // NTR+1: first tile of each row (row-major numbering, diagonal first, RHS last)
	std::vector<int32_t> row;       // 4 per row (+ sentinel): {first tile, end tile | has_superdiag<<30, first update, end update}
	std::vector<int32_t> tile_ij;   // ntiles: (ti<<16)|tj
	std::vector<int32_t> src;       // ntiles: offset of the tile's (0,0) in the reduced blocks (bit 30: diagonal tile, bit 29: transposed), -1 fill-in, -2-row right-hand side
	std::vector<int32_t> phase;     // 16 per phase (+ sentinel), see symbolic_tiles
	int n_phases = 0;
	std::vector<int32_t> orow;      // NTR: tile row in elimination order -> tile row in the reference's unknown order
	bool on_chip = false;           // the tiles fit the shared-memory budget (else they live in the L2-resident workspace)
	bool wide = false;              // more than 32767 tiles: update descriptors are {dst, tile(k,i), tile(k,j)} (3 ints)
	std::vector<int32_t> upd_ptr;   // NTR+1
	std::vector<int32_t> upd;       // 2 per update: {dst tile | tile(k,i)<<16, tile(k,j)}; the update of the next pivot block is implicit
	bool ok = false;                // false: too large for the on-chip solver (schedule not built)
	// level schedule (symbolic_levels, wave mode): `phase` holds 2 ints per level (+ sentinel) {first row, first dst entry}, `upd` holds
	// [dst tile (n_dst) | contribution pointer (n_dst + 1) | 1 pad | (tile(k,a), tile(k,b)) x n_contrib]
	bool levels = false;
	int n_dst = 0, n_contrib = 0, tail_level = 0;
	void clear() { rs.clear(); row.clear(); phase.clear(); n_phases = 0; src.clear(); orow.clear(); on_chip = false; wide = false; tile_ij.clear(); upd_ptr.clear(); upd.clear(); ok = false; levels = false; n_dst = n_contrib = 0; }
};
// Elimination order of the unknown edges (block level).  The reference's own order is kept whenever its fill fits on chip
// (a relative graph-SLAM chain is block-banded in that order: no fill at all); areas with loop closures fill the whole band
// between the two ends of every loop, so they are eliminated in greedy MINIMUM-DEGREE order instead.  Only the increment
// must match the reference (any elimination order of an SPD system gives the same solution up to rounding, and meets a
// non-positive pivot iff the matrix is not positive definite).
// Returns the number of 3x3 tiles of the filled pattern (right-hand side column included) for `order`.
long elimination_fill(const std::vector<uint64_t> &adj0, int np, int Wb, int q, const std::vector<int> *order, std::vector<int> *md_order) {
	std::vector<uint64_t> adj(adj0);
	std::vector<uint64_t> alive(Wb, 0);
	for (int i = 0; i < np; i++) alive[i >> 6] |= 1ull << (i & 63);
	long tiles = (long)q * np;  // right-hand side
	std::vector<uint64_t> nb(Wb);
	for (int step = 0; step < np; step++) {
		int pick = -1;
		if (order) pick = (*order)[step];
		else {
			int best = 1 << 30;
			for (int i = 0; i < np; i++) {
				if (!(alive[i >> 6] >> (i & 63) & 1)) continue;
				int d = 0;
				for (int w = 0; w < Wb; w++) d += __builtin_popcountll(adj[(size_t)i * Wb + w] & alive[w]);
				if (d < best) { best = d; pick = i; }
			}
			md_order->push_back(pick);
		}
		alive[pick >> 6] &= ~(1ull << (pick & 63));
		int deg = 0;
		for (int w = 0; w < Wb; w++) { nb[w] = adj[(size_t)pick * Wb + w] & alive[w]; deg += __builtin_popcountll(nb[w]); }
		tiles += (long)q * (q + 1) / 2 + (long)q * q * deg;
		for (int w = 0; w < Wb; w++) {
			uint64_t m = nb[w];
			while (m) {
				const int j = w * 64 + __builtin_ctzll(m);
				m &= m - 1;
				for (int w2 = 0; w2 < Wb; w2++) adj[(size_t)j * Wb + w2] |= nb[w2];
				adj[(size_t)j * Wb + (j >> 6)] &= ~(1ull << (j & 63));
			}
		}
	}
	return tiles;
}

void symbolic_tiles(const srba_problem_t &p, int P, long max_tiles_on_chip, TileSymbolic &sy, bool two_front = true) {
	const int q = P / 3, np = p.n_unk_edges, ntr = q * np, ntc = ntr + 1;
	sy.clear();
	sy.ntr = ntr;
	if (ntr <= 0 || ntr > 16384) return;  // tile coordinates are packed (ti << 16) | tj; the pivot scratch (3 doubles per tile row) bounds what actually runs on chip
	// ---- elimination order
	const int Wb = (np + 63) / 64;
	std::vector<uint64_t> adj((size_t)np * Wb, 0);
	for (int b = 0; b < p.n_S; b++) {
		const int r = p.S_row[b], c = p.S_col[b];
		if (r == c) continue;
		adj[(size_t)r * Wb + (c >> 6)] |= 1ull << (c & 63);
		adj[(size_t)c * Wb + (r >> 6)] |= 1ull << (r & 63);
	}
	std::vector<int> perm(np);  // position in the elimination order -> unknown edge
	std::iota(perm.begin(), perm.end(), 0);
	long tiles = elimination_fill(adj, np, Wb, q, &perm, nullptr);
	if (q == 1 && np >= 8 && tiles <= max_tiles_on_chip && two_front) {
		// chain-like area (block-banded in the reference's order): eliminate from BOTH ends at once -- order 0, n-1, 1, n-2, ...
		// The two fronts are independent until they meet, so the solver retires two block rows per barrier.
		std::vector<int> tw(np);
		for (int i = 0; i < np; i++) tw[i] = (i & 1) ? np - 1 - i / 2 : i / 2;
		const long tiles_tw = elimination_fill(adj, np, Wb, q, &tw, nullptr);
		if (tiles_tw <= max_tiles_on_chip && tiles_tw <= tiles + tiles / 4) { perm = tw; tiles = tiles_tw; }
	}
	if (tiles > max_tiles_on_chip) {
		std::vector<int> md;
		const long tiles_md = elimination_fill(adj, np, Wb, q, nullptr, &md);
		if (tiles_md < tiles) { perm = md; tiles = tiles_md; }
	}
	if (tiles > (1L << 20)) return;  // > 96 MB of tiles: such systems use the blocked dense solver
	sy.wide = tiles > 32767 || option(OPT_WIDE) != 0;  // (test hook)  // tile ids no longer fit 15/16 bits: 3-int update descriptors
	sy.on_chip = tiles <= max_tiles_on_chip;
	std::vector<int> inv(np);
	for (int i = 0; i < np; i++) inv[perm[i]] = i;
	sy.orow.resize(ntr);
	for (int t = 0; t < ntr; t++) sy.orow[t] = q * perm[t / q] + t % q;
	// ---- initial tile pattern in elimination order + where every tile comes from
	const int W = (ntc + 63) / 64;
	std::vector<uint64_t> rows((size_t)ntr * W, 0);
	std::vector<int32_t> srcmap((size_t)ntr * ntc, -1);
	auto setbit = [&](int r, int c) { rows[(size_t)r * W + (c >> 6)] |= 1ull << (c & 63); };
	for (int r = 0; r < ntr; r++) { setbit(r, r); setbit(r, ntr); srcmap[(size_t)r * ntc + ntr] = -2 - sy.orow[r]; }
	for (int b = 0; b < p.n_S; b++) {
		const int pr = inv[p.S_row[b]], pc = inv[p.S_col[b]];
		for (int a = 0; a < q; a++)
			for (int c = 0; c < q; c++) {
				const int32_t off = b * P * P + a * 3 * P + c * 3;  // sub-block (a,c) of the stored block (row <= col in the reference's order)
				int ti = q * pr + a, tj = q * pc + c, flags = 0;
				if (pr == pc) { if (a > c) continue; if (a == c) flags = 1 << 30; }
				else if (pr > pc) { std::swap(ti, tj); flags = 1 << 29; }  // the tile is the TRANSPOSE of the stored sub-block
				setbit(ti, tj);
				srcmap[(size_t)ti * ntc + tj] = off | flags;
			}
	}
	long total = 0;
	for (int k = 0; k < ntr; k++) {
		const uint64_t *rk = &rows[(size_t)k * W];
		// every row i > k present in row k inherits row k's columns >= i
		for (int w = k >> 6; w < W; w++) {
			uint64_t m = rk[w];
			if (w == (k >> 6)) m &= ~((2ull << (k & 63)) - 1);  // bits > k
			while (m) {
				const int i = w * 64 + __builtin_ctzll(m);
				m &= m - 1;
				if (i >= ntr) continue;  // right-hand-side column
				uint64_t *ri = &rows[(size_t)i * W];
				for (int w2 = i >> 6; w2 < W; w2++) {
					uint64_t add = rk[w2];
					if (w2 == (i >> 6)) add &= ~((1ull << (i & 63)) - 1);  // bits >= i
					ri[w2] |= add;
				}
			}
		}
		for (int w = 0; w < W; w++) total += __builtin_popcountll(rk[w]);
	}
	// numbering
	sy.rs.assign(ntr + 1, 0);
	sy.tile_ij.reserve(total);
	for (int r = 0; r < ntr; r++) {
		sy.rs[r] = (int32_t)sy.tile_ij.size();
		const uint64_t *rr = &rows[(size_t)r * W];
		for (int w = r >> 6; w < W; w++) {
			uint64_t m = rr[w];
			if (w == (r >> 6)) m &= ~((1ull << (r & 63)) - 1);
			while (m) { const int c = w * 64 + __builtin_ctzll(m); m &= m - 1; sy.tile_ij.push_back((r << 16) | c); }
		}
	}
	sy.rs[ntr] = (int32_t)sy.tile_ij.size();
	// id of tile (i,j): small systems keep a dense (row, column) -> id table, large ones search the row (tiles of a row are
	// contiguous, ascending column)
	const bool dense_ids = (size_t)ntr * ntc <= (1u << 20);
	std::vector<int32_t> tid;
	if (dense_ids) {
		tid.assign((size_t)ntr * ntc, -1);
		for (size_t t = 0; t < sy.tile_ij.size(); t++) tid[(size_t)(sy.tile_ij[t] >> 16) * ntc + (sy.tile_ij[t] & 0xffff)] = (int32_t)t;
	}
	auto find_tile = [&](int i, int j) {
		if (dense_ids) return tid[(size_t)i * ntc + j];
		const int32_t key = (i << 16) | j;
		const auto it = std::lower_bound(sy.tile_ij.begin() + sy.rs[i], sy.tile_ij.begin() + sy.rs[i + 1], key);
		return (int32_t)(it - sy.tile_ij.begin());
	};
	// phases: consecutive rows p, p+1 are eliminated together when they are independent -- no tile (p,p+1) and no common
	// column (then their update targets are disjoint and each pivot of the next phase is touched by at most one of them)
	auto has = [&](int r, int c) { return (rows[(size_t)r * W + (c >> 6)] >> (c & 63)) & 1; };
	std::vector<int> phase_row;  // first row of each phase (+ sentinel)
	for (int pz = 0; pz < ntr;) {
		phase_row.push_back(pz);
		bool pair = pz + 1 < ntr && !has(pz, pz + 1);
		if (pair) {
			const uint64_t *ra = &rows[(size_t)pz * W], *rb = &rows[(size_t)(pz + 1) * W];
			for (int w = 0; w < W && pair; w++) {
				uint64_t m = ra[w] & rb[w];
				if (w == (ntr >> 6)) m &= ~(1ull << (ntr & 63));         // the right-hand-side column is shared by construction
				if (w == ((pz + 1) >> 6)) m &= ~(1ull << ((pz + 1) & 63));  // row p+1's own diagonal is not a column of row p
				if (m) pair = false;
			}
		}
		pz += pair ? 2 : 1;
	}
	const int nph = (int)phase_row.size();
	phase_row.push_back(ntr);
	std::vector<int> phase_of(ntr + 1, nph);
	for (int m = 0; m < nph; m++)
		for (int r = phase_row[m]; r < phase_row[m + 1]; r++) phase_of[r] = m;
	// update schedule: row k's pairs of tiles; the updates of the NEXT phase's pivot blocks are implicit (done by the lanes
	// that invert those pivots)
	sy.upd_ptr.assign(ntr + 1, 0);
	sy.upd.reserve(4 * sy.tile_ij.size());
	int32_t n_upd = 0;
	for (int k = 0; k < ntr; k++) {
		sy.upd_ptr[k] = n_upd;
		for (int a = sy.rs[k] + 1; a < sy.rs[k + 1]; a++) {
			const int i = sy.tile_ij[a] & 0xffff;
			if (i >= ntr) break;  // the right-hand side is never a row
			for (int b = a; b < sy.rs[k + 1]; b++) {
				const int j = sy.tile_ij[b] & 0xffff;
				if (i == j && phase_of[i] == phase_of[k] + 1) continue;  // next pivot block
				if (sy.wide) { sy.upd.push_back(find_tile(i, j)); sy.upd.push_back(a); sy.upd.push_back(b); }
				else {
					sy.upd.push_back(find_tile(i, j) | (a << 16));  // dst | tile(k,i) << 16
					sy.upd.push_back(b);                             // tile(k,j)
				}
				n_upd++;
			}
		}
	}
	sy.upd_ptr[ntr] = n_upd;
	// per-row records
	sy.row.assign(4 * (size_t)(ntr + 1), 0);
	for (int k = 0; k <= ntr; k++) {
		int32_t *r = &sy.row[4 * (size_t)k];
		if (k == ntr) { r[0] = r[1] = sy.rs[ntr]; r[2] = r[3] = sy.upd_ptr[ntr]; break; }
		r[0] = sy.rs[k]; r[1] = sy.rs[k + 1]; r[2] = sy.upd_ptr[k]; r[3] = sy.upd_ptr[k + 1];
	}
	// per-phase records (16 ints, + sentinel): {first row, rows, first update, end of the first row's updates | end update,
	// pivots in the next phase, pivot tile 0, pivot tile 1 | link 0, link 1, first tile, first tile of the second row |
	// end tile, 0, 0, 0}; link g = the tile (r, q_g) through which a row r of this phase updates pivot q_g of the next one
	// (bit 30: r is the phase's second row), or -1
	sy.phase.assign(16 * (size_t)(nph + 1), 0);
	for (int m = 0; m <= nph; m++) {
		int32_t *r = &sy.phase[16 * (size_t)m];
		const int row0 = phase_row[std::min(m, nph)], nrows = m < nph ? phase_row[m + 1] - row0 : 0;
		r[0] = row0; r[1] = nrows; r[2] = sy.upd_ptr[row0]; r[3] = sy.upd_ptr[std::min(row0 + 1, ntr)]; r[4] = sy.upd_ptr[std::min(row0 + nrows, ntr)];
		const int nnext = m + 1 < nph ? phase_row[m + 2] - phase_row[m + 1] : 0;
		r[5] = nnext; r[6] = r[7] = 0; r[8] = r[9] = -1;
		for (int g = 0; g < nnext; g++) {
			const int qrow = phase_row[m + 1] + g;
			r[6 + g] = sy.rs[qrow];
			for (int rr = row0; rr < row0 + nrows; rr++)
				if (has(rr, qrow)) r[8 + g] = find_tile(rr, qrow) | (rr > row0 ? 1 << 30 : 0);
		}
		r[10] = sy.rs[row0]; r[11] = sy.rs[std::min(row0 + 1, ntr)]; r[12] = sy.rs[std::min(row0 + nrows, ntr)];
	}
	sy.n_phases = nph;
	sy.src.resize(sy.tile_ij.size());
	for (size_t t = 0; t < sy.tile_ij.size(); t++) sy.src[t] = srcmap[(size_t)(sy.tile_ij[t] >> 16) * ntc + (sy.tile_ij[t] & 0xffff)];
	sy.ok = true;
}

// LEVEL schedule of the block-sparse L D L^t for the wave pipeline's multi-CTA solver (lm_wave.cuh: wave_solve_kernel).
// Elimination order: minimum degree, eliminating a maximal independent set of minimum-degree nodes per round (same fill as greedy
// minimum degree, shorter elimination trees).  The tile rows are then SORTED BY LEVEL -- level(j) = 1 + max level(i) over the rows
// i that hold a tile in column j, i.e. the rows whose elimination updates row j -- which is a topological order of the same
// elimination (same filled pattern), so that all rows of a level are consecutive and can be eliminated together:
//   per level:  A  every row: invert its 3x3 pivot block, W(k,.) = inv(D_k) T(k,.)               (one warp per row)
//               B  every tile (a,b) that some row k of the level updates: T(a,b) -= sum_k T(k,a)^t W(k,b), contributions of one
//                  destination grouped (one 9-lane group owns the destination: no atomics, fixed summation order)
// and the backward pass walks the levels downwards, one warp per row.  Two team barriers per level instead of one per block row:
// a 641-edge loop-closure area has 150 levels for 641 rows, and every phase has enough independent work for thousands of threads.
void symbolic_levels(const srba_problem_t &p, int P, TileSymbolic &sy, bool debug_format = false) {
	static const bool sym_timing = std::getenv("SRBA_B200_SYM_TIMING") != nullptr;
	auto sym_t0 = std::chrono::steady_clock::now();
#define SYM_T(name) do { if (sym_timing) { auto n_ = std::chrono::steady_clock::now(); std::fprintf(stderr, "[sym] %s %.2f ms\n", name, std::chrono::duration<double, std::milli>(n_ - sym_t0).count()); sym_t0 = n_; } } while (0)
	const int q = P / 3, np = p.n_unk_edges, ntr = q * np, ntc = ntr + 1;
	sy.clear();
	sy.ntr = ntr;
	if (ntr <= 0 || ntr > 32000) return;
	// ---- elimination order (block level): rounds of independent minimum-degree nodes
	const int Wb = (np + 63) / 64;
	std::vector<uint64_t> adj((size_t)np * Wb, 0), alive(Wb, 0), blocked(Wb), nb(Wb);
	for (int b = 0; b < p.n_S; b++) {
		const int r = p.S_row[b], c = p.S_col[b];
		if (r == c) continue;
		adj[(size_t)r * Wb + (c >> 6)] |= 1ull << (c & 63);
		adj[(size_t)c * Wb + (r >> 6)] |= 1ull << (r & 63);
	}
	for (int i = 0; i < np; i++) alive[i >> 6] |= 1ull << (i & 63);
	std::vector<int> perm, deg(np);
	perm.reserve(np);
	while ((int)perm.size() < np) {
		int dmin = 1 << 30;
		for (int i = 0; i < np; i++) {
			if (!(alive[i >> 6] >> (i & 63) & 1)) { deg[i] = -1; continue; }
			int d = 0;
			for (int w = 0; w < Wb; w++) d += __builtin_popcountll(adj[(size_t)i * Wb + w] & alive[w]);
			deg[i] = d;
			dmin = std::min(dmin, d);
		}
		std::fill(blocked.begin(), blocked.end(), 0);
		for (int i = 0; i < np; i++) {
			if (deg[i] != dmin || (blocked[i >> 6] >> (i & 63) & 1)) continue;
			// eliminate i: its alive neighbours become a clique and are blocked for this round
			alive[i >> 6] &= ~(1ull << (i & 63));
			for (int w = 0; w < Wb; w++) { nb[w] = adj[(size_t)i * Wb + w] & alive[w]; blocked[w] |= nb[w]; }
			for (int w = 0; w < Wb; w++) {
				uint64_t m = nb[w];
				while (m) {
					const int j = w * 64 + __builtin_ctzll(m);
					m &= m - 1;
					for (int w2 = 0; w2 < Wb; w2++) adj[(size_t)j * Wb + w2] |= nb[w2];
					adj[(size_t)j * Wb + (j >> 6)] &= ~(1ull << (j & 63));
				}
			}
			perm.push_back(i);
		}
	}
	SYM_T("order");
	std::vector<int> inv(np);
	for (int i = 0; i < np; i++) inv[perm[i]] = i;
	// ---- filled pattern at tile level in that order (row bitsets, column ntr = right-hand side) + where every tile comes from
	const int W = (ntc + 63) / 64;
	std::vector<uint64_t> rows((size_t)ntr * W, 0);
	auto setbit = [&](int r, int c) { rows[(size_t)r * W + (c >> 6)] |= 1ull << (c & 63); };
	struct Src { int64_t key; int32_t val; };
	std::vector<Src> srcs;
	srcs.reserve((size_t)p.n_S * q * q);
	for (int r = 0; r < ntr; r++) { setbit(r, r); setbit(r, ntr); }
	for (int b = 0; b < p.n_S; b++) {
		const int pr = inv[p.S_row[b]], pc = inv[p.S_col[b]];
		for (int a = 0; a < q; a++)
			for (int c = 0; c < q; c++) {
				const int32_t off = b * P * P + a * 3 * P + c * 3;  // sub-block (a,c) of the stored block (row <= col in the reference's order)
				int ti = q * pr + a, tj = q * pc + c, flags = 0;
				if (pr == pc) { if (a > c) continue; if (a == c) flags = 1 << 30; }
				else if (pr > pc) { std::swap(ti, tj); flags = 1 << 29; }  // the tile is the TRANSPOSE of the stored sub-block
				setbit(ti, tj);
				srcs.push_back(Src{((int64_t)ti << 32) | tj, off | flags});
			}
	}
	for (int k = 0; k < ntr; k++) {
		const uint64_t *rk = &rows[(size_t)k * W];
		for (int w = k >> 6; w < W; w++) {
			uint64_t m = rk[w];
			if (w == (k >> 6)) m &= ~((2ull << (k & 63)) - 1);  // bits > k
			while (m) {
				const int i = w * 64 + __builtin_ctzll(m);
				m &= m - 1;
				if (i >= ntr) continue;
				uint64_t *ri = &rows[(size_t)i * W];
				for (int w2 = i >> 6; w2 < W; w2++) {
					uint64_t add = rk[w2];
					if (w2 == (i >> 6)) add &= ~((1ull << (i & 63)) - 1);  // bits >= i
					ri[w2] |= add;
				}
			}
		}
	}
	SYM_T("fill");
	// ---- levels, rows sorted by level (stable): rank[old row] = new row
	std::vector<int> level(ntr, 0), order(ntr), rank(ntr);
	for (int k = 0; k < ntr; k++) {
		const uint64_t *rk = &rows[(size_t)k * W];
		for (int w = k >> 6; w < W; w++) {
			uint64_t m = rk[w];
			if (w == (k >> 6)) m &= ~((2ull << (k & 63)) - 1);
			while (m) {
				const int j = w * 64 + __builtin_ctzll(m);
				m &= m - 1;
				if (j < ntr) level[j] = std::max(level[j], level[k] + 1);
			}
		}
	}
	std::iota(order.begin(), order.end(), 0);
	std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return level[a] < level[b]; });
	for (int r = 0; r < ntr; r++) rank[order[r]] = r;
	const int nlev = level[order[ntr - 1]] + 1;
	sy.orow.resize(ntr);
	for (int r = 0; r < ntr; r++) { const int o = order[r]; sy.orow[r] = q * perm[o / q] + o % q; }
	SYM_T("levels");
	// ---- tile numbering in the new order: row-major, diagonal first, ascending column, right-hand side last
	std::sort(srcs.begin(), srcs.end(), [](const Src &a, const Src &b) { return a.key < b.key; });
	sy.rs.assign(ntr + 1, 0);
	std::vector<int> cols;
	for (int r = 0; r < ntr; r++) {
		sy.rs[r] = (int32_t)sy.tile_ij.size();
		const int o = order[r];
		const uint64_t *ro = &rows[(size_t)o * W];
		cols.clear();
		for (int w = o >> 6; w < W; w++) {
			uint64_t m = ro[w];
			if (w == (o >> 6)) m &= ~((1ull << (o & 63)) - 1);
			while (m) { const int c = w * 64 + __builtin_ctzll(m); m &= m - 1; cols.push_back(c < ntr ? rank[c] : ntr); }
		}
		std::sort(cols.begin(), cols.end());  // (every column of old row o ranks above r: its level is higher)
		for (int c : cols) {
			sy.tile_ij.push_back((r << 16) | c);
			int32_t sv = -1;
			if (c == ntr) sv = -2 - sy.orow[r];
			else {
				const int oc = order[c];
				const int64_t key = ((int64_t)std::min(o, oc) << 32) | std::max(o, oc);  // stored with old ti < tj
				auto it = std::lower_bound(srcs.begin(), srcs.end(), key, [](const Src &a, int64_t k) { return a.key < k; });
				if (it != srcs.end() && it->key == key) sv = it->val;
			}
			sy.src.push_back(sv);
		}
	}
	sy.rs[ntr] = (int32_t)sy.tile_ij.size();
	const int ntiles = (int)sy.tile_ij.size();
	sy.row.assign(4 * (size_t)(ntr + 1), 0);
	for (int k = 0; k <= ntr; k++) { sy.row[4 * (size_t)k] = sy.rs[std::min(k, ntr)]; sy.row[4 * (size_t)k + 1] = sy.rs[std::min(k + 1, ntr)]; }
	auto find_tile = [&](int i, int j) {
		const int32_t key = (i << 16) | j;
		return (int32_t)(std::lower_bound(sy.tile_ij.begin() + sy.rs[i], sy.tile_ij.begin() + sy.rs[i + 1], key) - sy.tile_ij.begin());
	};
	SYM_T("tiles");
	// ---- per level: destinations with their contributions (ascending source row).  Built directly in the device format (lm_wave.cuh:
	// wave_solve_fwd_kernel): 16-byte records so that a destination with ONE contribution -- almost all of them -- costs two dependent loads:
	//   [n_dst x {dst, tile(k,a), tile(k,b), diagonal tile of row k} : the FIRST contribution of each destination]
	//   [n_dst + 1 pointers into the further contributions, padded to 4 ints] [{tile(k,a), tile(k,b), diagonal tile of row k, 0} ...]
	// One pass: the first contribution of a destination writes its record; the rare further ones go to a side list that is bucketed by
	// destination entry afterwards (stable: ascending source row is kept).
	// (the records are written straight into sy.upd, the pointer and side lists are appended behind them: no second copy of the 16 n_dst bytes)
	std::vector<int32_t> rec_dbg, ex, slot(ntiles, -1), touched;
	std::vector<int32_t> &rec = debug_format ? rec_dbg : sy.upd;
	rec.clear();
	{  // exact number of destination entries is not known yet: an upper bound on the contributions, capped, avoids most regrowth
		size_t bound = 0;
		for (int k = 0; k < ntr; k++) { const size_t m = (size_t)(sy.rs[k + 1] - sy.rs[k]) - 1; bound += m * (m + 1) / 2; }
		rec.reserve(5 * std::min<size_t>(bound, (size_t)4 << 20) + 64);   // records + pointers + side list
	}
	// Destination look-up.  Row i holds every column j >= i of row k (fill).  Small systems: a dense (row, column) -> tile table, one
	// look-up per contribution; large ones: both lists ascend, so a merge walk along row i finds the tiles.
	const bool dense_map = (size_t)ntr * ntc <= ((size_t)16 << 20);
	std::vector<int32_t> tile_of;
	if (dense_map) {
		tile_of.resize((size_t)ntr * ntc);   // only entries of stored tiles are ever read
		for (int t = 0; t < ntiles; t++) tile_of[(size_t)(sy.tile_ij[t] >> 16) * ntc + (sy.tile_ij[t] & 0xffff)] = t;
	}
	sy.phase.assign(2 * (size_t)(nlev + 1), 0);
	int r = 0, ne = 0;
	for (int L = 0; L < nlev; L++) {
		sy.phase[2 * (size_t)L] = r;
		sy.phase[2 * (size_t)L + 1] = ne;
		const int r0 = r;
		while (r < ntr && level[order[r]] == L) r++;
		touched.clear();
		for (int k = r0; k < r; k++) {
			const int32_t td = sy.rs[k], k_end = sy.rs[k + 1];
			for (int a = td + 1; a < k_end - 1; a++) {  // tile (k,i): off-diagonal, not the right-hand side
				const int i = sy.tile_ij[a] & 0xffff;
				int32_t d = sy.rs[i];
				const int32_t *row_i = dense_map ? &tile_of[(size_t)i * ntc] : nullptr;
				for (int b = a; b < k_end; b++) {           // tile (k,j), j >= i, right-hand side included
					if (dense_map) d = row_i[sy.tile_ij[b] & 0xffff];
					else {
						const int32_t key = (i << 16) | (sy.tile_ij[b] & 0xffff);
						while (sy.tile_ij[d] != key) d++;
					}
					if (slot[d] < 0) {
						slot[d] = ne++;
						touched.push_back(d);
						rec.push_back(d); rec.push_back(a); rec.push_back(b); rec.push_back(td);
					} else { ex.push_back(slot[d]); ex.push_back(a); ex.push_back(b); ex.push_back(td); }
				}
			}
		}
		for (int32_t d : touched) slot[d] = -1;
	}
	SYM_T("dests");
	sy.phase[2 * (size_t)nlev] = ntr;
	sy.phase[2 * (size_t)nlev + 1] = ne;
	sy.n_phases = nlev;
	// the tail: the last levels holding at most SRBA_WAVE_TAIL_ROWS tile rows in total (dense elimination in shared memory)
	sy.tail_level = nlev;
	while (sy.tail_level > 0 && ntr - sy.phase[2 * (size_t)(sy.tail_level - 1)] <= SRBA_WAVE_TAIL_ROWS) sy.tail_level--;
	const size_t nd = (size_t)ne, n_extra = ex.size() / 4, xlen = (nd + 1 + 3) & ~(size_t)3;
	sy.n_dst = ne;
	sy.n_contrib = (int)(nd + n_extra);
	// further contributions bucketed by destination entry (counting sort, stable)
	std::vector<int32_t> xptr(xlen, 0), extra(4 * n_extra);
	for (size_t q = 0; q < n_extra; q++) xptr[(size_t)ex[4 * q] + 1]++;
	for (size_t e = 0; e < nd; e++) xptr[e + 1] += xptr[e];
	for (size_t e = nd + 1; e < xlen; e++) xptr[e] = xptr[nd];
	{
		std::vector<int32_t> at(xptr.begin(), xptr.begin() + nd);
		for (size_t q = 0; q < n_extra; q++) {
			const size_t w = (size_t)at[ex[4 * q]]++;
			extra[4 * w] = ex[4 * q + 1]; extra[4 * w + 1] = ex[4 * q + 2]; extra[4 * w + 2] = ex[4 * q + 3]; extra[4 * w + 3] = 0;
		}
	}
	if (debug_format) {  // srba_b200_debug_level_schedule: [dst | contribution pointer | 1 pad | (tile(k,a), tile(k,b)) x n_contrib]
		sy.upd.clear();
		sy.upd.reserve(2 * nd + 2 + 2 * (nd + n_extra));
		for (size_t e = 0; e < nd; e++) sy.upd.push_back(rec[4 * e]);
		for (size_t e = 0; e <= nd; e++) sy.upd.push_back((int32_t)(e + (size_t)xptr[e]));   // first + further contributions before entry e
		sy.upd.push_back(0);
		for (size_t e = 0; e < nd; e++) {
			sy.upd.push_back(rec[4 * e + 1]); sy.upd.push_back(rec[4 * e + 2]);
			for (int32_t c = xptr[e]; c < xptr[e + 1]; c++) { sy.upd.push_back(extra[4 * (size_t)c]); sy.upd.push_back(extra[4 * (size_t)c + 1]); }
		}
	} else {
		sy.upd.resize(4 * nd + xlen + 4 * n_extra);   // (the 4 nd record ints are already in place)
		std::memcpy(sy.upd.data() + 4 * nd, xptr.data(), xlen * sizeof(int32_t));
		if (n_extra) std::memcpy(sy.upd.data() + 4 * nd + xlen, extra.data(), 4 * n_extra * sizeof(int32_t));
	}
	SYM_T("format");
	sy.levels = true;
	sy.ok = true;
}

int model_index(const srba_model_t &m) {
	return m.pose_kind == SRBA_POSE_SE2 ? 0 : (m.obs_kind == SRBA_OBS_CARTESIAN3D ? 1 : (m.obs_kind == SRBA_OBS_MONO ? 2 : 3));
}

// FMAs of one linearisation + Schur reduction: areas above ~1e6 are throughput-bound on one SM (SE(3) landmark areas),
// below they are latency-bound (graph-SLAM areas, even with loop closures) and gain nothing from more SMs.
bool heavy_problem(const srba_problem_t &p) {
	const int P = p.model.pose_kind == SRBA_POSE_SE2 ? 3 : 6, o = obs_dims(p.model.obs_kind);
	const double w = (double)p.n_blkA * o * P * 10 + (double)p.n_blkF * o * 3 * 10 + ((double)p.n_HAp_terms * P * P + (double)p.n_HApf_terms * P * 3) * o +
	                 (double)p.n_S_terms * P * P * 3;
	return w > 1e6;
}
// Would "auto" send this problem to the grid-wide wave pipeline?  (pack_problem's rule on the array sizes alone: an SE(2) area whose core
// arrays overflow one CTA's shared-memory arena; areas that overflow only through their tiles are found at pack time.)
bool wave_candidate(const srba_problem_t &p) {
	const int fw = option(OPT_WAVE);
	if (fw == 0 || option(OPT_CLUSTER) == 1 || option(OPT_BLOCKED) != 0) return false;
	if (fw == 1) return true;
	if (p.model.pose_kind != SRBA_POSE_SE2) return p.n_unk_edges >= SRBA_WAVE_SE3_MIN_EDGES;
	const uint64_t demand = (uint64_t)2 * p.n_pairs * 3 * 8 + (uint64_t)p.n_blkA * 3 * 3 * 8 + (uint64_t)p.n_HAp * 9 * 8;
	return demand > (uint64_t)srba_kernel_cfg(0).smem_bytes;
}
// A cluster takes 8 SMs: it shortens the LATENCY of a heavy area, while a batch with many heavy areas already fills the
// GPU with one CTA per area.
bool clusters_pay(int n_heavy) { return n_heavy > 0 && n_heavy <= 36; }

// Fills the header's layout fields; when `base` != NULL also copies the arrays.
void pack_problem(const srba_problem_t &p, DevProblem &h, char *base, int trace_cap, int want_dense, TileSymbolic &sy, bool allow_cluster) {
	std::memset(&h, 0, sizeof(h));
	h.model = p.model;
	h.n_edges = p.n_edges; h.n_unk_edges = p.n_unk_edges; h.n_lms = p.n_lms; h.n_unk_lms = p.n_unk_lms; h.n_pairs = p.n_pairs;
	h.n_path = p.n_path; h.n_obs = p.n_obs; h.n_blkA = p.n_blkA; h.n_blkF = p.n_blkF; h.n_HAp = p.n_HAp; h.n_HAp_terms = p.n_HAp_terms;
	h.n_HApf = p.n_HApf; h.n_HApf_terms = p.n_HApf_terms; h.n_S = p.n_S; h.n_S_terms = p.n_S_terms; h.n_total_obs_rows = p.n_total_obs_rows;
	const int P = p.model.pose_kind == SRBA_POSE_SE2 ? 3 : 6;
	const int S = P == 3 ? 3 : 12;
	const int o = obs_dims(p.model.obs_kind);
	h.n_dense = P * p.n_unk_edges;
	h.n_scalars = h.n_dense + 3 * p.n_unk_lms;
	h.ld_dense = (h.n_dense + 1 + 1) / 2 * 2;
	h.trace_cap = trace_cap;
	h.want_dense = want_dense;
	// ---- dense solver: register-resident block-sparse tiles when the filled pattern fits the model kernel's register file
	const int model = model_index(p.model);
	const SrbaKernelCfg cfg = srba_kernel_cfg(model);
	const bool force_blocked = option(OPT_BLOCKED) != 0;  // tests
	// ---- execution mode.  An area whose core arrays fit one CTA's shared-memory arena runs the persistent one-CTA LM program.
	// Larger areas (SE(3) landmark areas, SE(2) areas with loop closures) keep every array in the global workspace and run
	//   WAVE    : the grid-wide lock-step pipeline (lm_wave.cuh) with the level-scheduled multi-CTA solver -- the default for such areas;
	//   CLUSTER : an 8-CTA cluster team (lm_kernel_cluster.cu) -- when the wave mode is switched off and few heavy areas share the batch;
	//   else one CTA with the overflow in the global workspace.  Options "wave" / "cluster" force the choice (tests).
	const int force_cluster = option(OPT_CLUSTER), force_wave = option(OPT_WAVE);
	const uint64_t demand_no_tiles = (uint64_t)2 * p.n_pairs * S * 8 + (uint64_t)p.n_blkA * o * P * 8 + (uint64_t)p.n_blkF * o * 3 * 8 + (uint64_t)p.n_HAp * P * P * 8 +
	                                 (uint64_t)2 * p.n_HApf * P * 3 * 8;
	// Measured (profiles/r02_*): the wave pipeline wins on SE(2) loop-closure areas (hundreds of unknown edges, 100-240 elimination levels:
	// 109 ms vs 172 ms per 256-area batch); on SE(3) landmark areas (n_p ~ 90, reduced system ~30 % dense) the level solver's tile
	// traffic through L2 makes it slower than one CTA per area (540 vs 345 ms per 1 000 areas), so "auto" keeps those on the one-CTA /
	// cluster-team program and the wave mode is opt-in there (option "wave" = 1) -- except whole-map-scale SE(3) problems (hundreds of unknown
	// edges in ONE area: BASELINE configs #2 / #3), where the level schedule + cluster team beat the sequential block-row elimination of the
	// 8-CTA cluster program 2.9x / 4.1x (9.5 -> 3.3 ms and 52.9 -> 12.9 ms per LM iteration, profiles/r02_bench_config{2,3}*.json).
	const bool wave_allowed = !want_dense && !force_blocked && force_cluster != 1 && force_wave != 0 && (force_wave == 1 || model == 0 || p.n_unk_edges >= SRBA_WAVE_SE3_MIN_EDGES);
	bool cluster = false, wave = false;
	if (!base) {  // layout pass: symbolic factorisation (reused by the copy pass)
		if (wave_allowed && (force_wave == 1 || demand_no_tiles > (uint64_t)cfg.smem_bytes)) { symbolic_levels(p, P, sy); wave = sy.ok; }
		if (!wave) {
			// tiles may take up to 3/4 of the CTA's shared memory (other arrays then spill to the L2-resident workspace)
			const long max_tiles = std::min(32767L, ((long)cfg.smem_bytes * 3 / 4 / 8 - 97 - h.n_dense) / 12);
			if (max_tiles > 0) symbolic_tiles(p, P, max_tiles, sy);
			else sy.clear();
			if (wave_allowed && sy.ok && demand_no_tiles + (uint64_t)sy.tile_ij.size() * 96 > (uint64_t)cfg.smem_bytes) {
				TileSymbolic lv;
				symbolic_levels(p, P, lv);
				if (lv.ok) { sy = std::move(lv); wave = true; }
			}
		}
	} else
		wave = sy.levels;
	if (sy.ok && !force_blocked && !wave) {
		const bool overflows = demand_no_tiles + (uint64_t)sy.tile_ij.size() * 96 > (uint64_t)cfg.smem_bytes;
		if (force_cluster == 1) cluster = true;
		else if (force_cluster < 0) cluster = allow_cluster && heavy_problem(p) && overflows;
	}
	h.cluster = cluster ? 1 : 0;
	h.wave = wave ? 1 : 0;
	const SrbaKernelCfg kcfg = cluster ? srba_kernel_cfg(model + 4) : cfg;
	const bool tiles_on_chip = sy.on_chip && !cluster && !wave;
	int reg_scratch = (sy.ok && !force_blocked) ? srba_chol_reg_scratch_doubles(cluster ? model + 4 : model, h.n_dense, tiles_on_chip ? (int)sy.tile_ij.size() : 0) : -1;
	if (wave) reg_scratch = 0;  // the wave solver keeps tiles, W tiles and the solution in the global workspace
	uint64_t scratch_bytes;
	if (reg_scratch >= 0) {
		h.chol_mode = SRBA_CHOL_REG;
		h.ldl_n_tiles = (int)sy.tile_ij.size();
		h.ldl_tiles_in_ws = tiles_on_chip ? 0 : 1;
		h.ldl_n_phases = sy.n_phases;
		h.ldl_n_dst = sy.n_dst;
		h.ldl_tail_level = sy.tail_level;
		h.ldl_n_contrib = sy.n_contrib;
		h.ldl_wide = sy.wide ? 1 : 0;
		scratch_bytes = (uint64_t)reg_scratch * 8;
	} else {
		h.chol_mode = SRBA_CHOL_BLOCKED;
		h.panel_cap = std::min(std::max((h.n_dense + 1 + 3) / 4 * 4, 24), 1024);
		scratch_bytes = (uint64_t)8 * h.panel_cap * 8;
	}
	// ---- input image
	uint64_t img_count[SRBA_I_COUNT];
	{
		uint64_t *c = img_count;
		c[SRBA_I_edge_pose] = (uint64_t)p.n_edges * S; c[SRBA_I_unk_edge] = p.n_unk_edges; c[SRBA_I_lm_pos] = (uint64_t)p.n_lms * 3; c[SRBA_I_unk_lm] = p.n_unk_lms;
		c[SRBA_I_pair_path_ptr] = p.n_pairs + 1; c[SRBA_I_path_edge] = p.n_path; c[SRBA_I_path_dir] = p.n_path;
		c[SRBA_I_obs_z] = (uint64_t)p.n_obs * o; c[SRBA_I_obs_slot] = p.n_obs; c[SRBA_I_obs_lm] = p.n_obs; c[SRBA_I_obs_mult] = p.n_obs;
		c[SRBA_I_colA_ptr] = p.n_unk_edges + 1; c[SRBA_I_blkA_obs] = p.n_blkA; c[SRBA_I_blkA_slotA] = p.n_blkA; c[SRBA_I_blkA_slotD] = p.n_blkA; c[SRBA_I_blkA_normal_dir] = p.n_blkA;
		c[SRBA_I_colF_ptr] = p.n_unk_lms + 1; c[SRBA_I_blkF_obs] = p.n_blkF; c[SRBA_I_blkF_slot] = p.n_blkF;
		c[SRBA_I_row_ptr] = p.n_obs + 1; c[SRBA_I_row_blk] = p.n_blkA + p.n_blkF;
		c[SRBA_I_HAp_row] = p.n_HAp; c[SRBA_I_HAp_col] = p.n_HAp; c[SRBA_I_HAp_term_ptr] = p.n_HAp + 1; c[SRBA_I_HAp_term_a] = p.n_HAp_terms; c[SRBA_I_HAp_term_b] = p.n_HAp_terms;
		c[SRBA_I_HApf_row] = p.n_HApf; c[SRBA_I_HApf_col] = p.n_HApf; c[SRBA_I_HApf_term_ptr] = p.n_HApf + 1; c[SRBA_I_HApf_term_a] = p.n_HApf_terms; c[SRBA_I_HApf_term_f] = p.n_HApf_terms;
		c[SRBA_I_rowApf_ptr] = p.n_unk_edges + 1; c[SRBA_I_colApf_ptr] = p.n_unk_lms + 1; c[SRBA_I_colApf_idx] = p.n_HApf;
		c[SRBA_I_S_row] = p.n_S; c[SRBA_I_S_col] = p.n_S; c[SRBA_I_S_HAp] = p.n_S; c[SRBA_I_S_term_ptr] = p.n_S + 1; c[SRBA_I_S_term_rowblk] = p.n_S_terms; c[SRBA_I_S_term_colblk] = p.n_S_terms;
		c[SRBA_I_blkA_col] = p.n_blkA; c[SRBA_I_blkF_col] = p.n_blkF;
		const bool reg = h.chol_mode == SRBA_CHOL_REG;
		c[SRBA_I_ldl_row] = reg ? sy.row.size() : 0; c[SRBA_I_ldl_tile_ij] = reg ? sy.tile_ij.size() : 0;
		c[SRBA_I_ldl_src] = reg ? sy.src.size() : 0; c[SRBA_I_ldl_upd] = reg ? sy.upd.size() : 0; c[SRBA_I_ldl_orow] = reg ? sy.orow.size() : 0; c[SRBA_I_ldl_phase] = reg ? sy.phase.size() : 0;
	}
	Layout L;
	L.add(sizeof(DevProblem));
#define X(T, name)                                                        \
	h.img_len[SRBA_I_##name] = (uint32_t)(img_count[SRBA_I_##name] * sizeof(T)); \
	h.img_off[SRBA_I_##name] = L.add(img_count[SRBA_I_##name] * sizeof(T));      \
	h.img_sm[SRBA_I_##name] = SRBA_NOT_IN_SMEM;
	SRBA_IMG_ARRAYS(X)
#undef X
	h.img_bytes = L.total();
	if (base) {
#define X(T, name) \
	if (src_ptr(p, SRBA_I_##name)) put<T>(base, h.img_off[SRBA_I_##name], reinterpret_cast<const T *>(src_ptr(p, SRBA_I_##name)), (size_t)img_count[SRBA_I_##name]);
		SRBA_IMG_ARRAYS(X)
#undef X
		// derived: column of every block
		int32_t *ca = reinterpret_cast<int32_t *>(base + h.img_off[SRBA_I_blkA_col]);
		for (int c = 0; c < p.n_unk_edges; c++)
			for (int b = p.colA_ptr[c]; b < p.colA_ptr[c + 1]; b++) ca[b] = c;
		int32_t *cf = reinterpret_cast<int32_t *>(base + h.img_off[SRBA_I_blkF_col]);
		for (int c = 0; c < p.n_unk_lms; c++)
			for (int b = p.colF_ptr[c]; b < p.colF_ptr[c + 1]; b++) cf[b] = c;
		if (h.chol_mode == SRBA_CHOL_REG) {
			put<int32_t>(base, h.img_off[SRBA_I_ldl_row], sy.row.data(), sy.row.size());
			put<int32_t>(base, h.img_off[SRBA_I_ldl_tile_ij], sy.tile_ij.data(), sy.tile_ij.size());
			put<int32_t>(base, h.img_off[SRBA_I_ldl_src], sy.src.data(), sy.src.size());
			put<int32_t>(base, h.img_off[SRBA_I_ldl_orow], sy.orow.data(), sy.orow.size());
			put<int32_t>(base, h.img_off[SRBA_I_ldl_phase], sy.phase.data(), sy.phase.size());
			put<int32_t>(base, h.img_off[SRBA_I_ldl_upd], sy.upd.data(), sy.upd.size());
		}
	}

	// ---- reduced system == H_Ap when there is no Schur term (relative graph-SLAM): no copy, no Sblk
	h.s_alias_hap = (p.n_S_terms == 0 && p.n_S == p.n_HAp) ? 1 : 0;
	if (h.s_alias_hap)
		for (int b = 0; b < p.n_S; b++)
			if (p.S_HAp[b] != b) { h.s_alias_hap = 0; break; }

	// ---- workspace
	uint64_t ws_bytes[SRBA_W_COUNT];
	{
		uint64_t *w = ws_bytes;
		w[SRBA_W_edge_cur] = (uint64_t)p.n_edges * S * 8; w[SRBA_W_edge_bak] = (uint64_t)p.n_unk_edges * S * 8;
		w[SRBA_W_lm_cur] = (uint64_t)p.n_lms * 3 * 8; w[SRBA_W_lm_bak] = (uint64_t)p.n_unk_lms * 3 * 8;
		w[SRBA_W_slot] = (uint64_t)2 * p.n_pairs * S * 8; w[SRBA_W_slot_bak] = w[SRBA_W_slot];
		w[SRBA_W_JA] = (uint64_t)p.n_blkA * o * P * 8; w[SRBA_W_JF] = (uint64_t)p.n_blkF * o * 3 * 8;
		w[SRBA_W_resid0] = (uint64_t)p.n_obs * o * 8; w[SRBA_W_resid1] = w[SRBA_W_resid0];
		w[SRBA_W_HAp] = (uint64_t)p.n_HAp * P * P * 8; w[SRBA_W_HApf] = (uint64_t)p.n_HApf * P * 3 * 8;
		w[SRBA_W_Hf] = (uint64_t)p.n_unk_lms * 9 * 8; w[SRBA_W_Hf_inv] = w[SRBA_W_Hf]; w[SRBA_W_W] = w[SRBA_W_HApf];
		w[SRBA_W_Sblk] = h.s_alias_hap ? 0 : (uint64_t)p.n_S * P * P * 8;
		w[SRBA_W_grad] = (uint64_t)h.n_scalars * 8; w[SRBA_W_delta] = w[SRBA_W_grad];
		// solver_engine<false,false> never touches the caller's gradient (lev-marq_solvers.h:182): the pristine copy the
		// in-place Schur reduction starts from at every trial
		w[SRBA_W_grad0] = (p.model.solver_kind == SRBA_SOLVER_NO_SCHUR_SPARSE_CHOLESKY && p.n_unk_lms > 0) ? w[SRBA_W_grad] : 0;
		w[SRBA_W_dense] = (h.chol_mode == SRBA_CHOL_BLOCKED || want_dense) ? (uint64_t)h.n_dense * h.ld_dense * 8 : 0;
		// (wave: T tiles, W tiles and the solution x of the level solver)
		w[SRBA_W_ldl_tiles] = (h.chol_mode == SRBA_CHOL_REG && h.ldl_tiles_in_ws) ? (uint64_t)h.ldl_n_tiles * 12 * 8 * (wave ? 2 : 1) + (wave ? (uint64_t)h.n_dense * 8 + 64 : 0) : 0;
		w[SRBA_W_valid] = p.n_obs; w[SRBA_W_blk_ok] = p.n_blkA + p.n_blkF; w[SRBA_W_lm_inv] = p.n_unk_lms;
		w[SRBA_W_slot_listed] = 2 * p.n_pairs; w[SRBA_W_listed_idx] = (uint64_t)2 * p.n_pairs * 4; w[SRBA_W_refresh_idx] = (uint64_t)p.n_pairs * 4;
	}
	Layout W;
	for (int a = 0; a < SRBA_W_COUNT; a++) { h.ws_off[a] = W.add(ws_bytes[a]); h.ws_sm[a] = SRBA_NOT_IN_SMEM; }
	h.ws_bytes = W.total();

	// ---- on-chip placement: scratch first, then arrays in priority order (most latency-critical per byte first) while they fit
	{
		Layout A;
		h.sm_scratch = (uint32_t)A.add(scratch_bytes);
		const uint64_t budget = (cluster || wave) ? 0 : (uint64_t)kcfg.smem_bytes;  // clusters and waves keep every array in the global workspace
		struct Cand { int is_img, id; };
		static const Cand order[] = {
			{0, SRBA_W_edge_cur}, {0, SRBA_W_edge_bak}, {0, SRBA_W_grad}, {0, SRBA_W_delta}, {0, SRBA_W_lm_cur}, {0, SRBA_W_lm_bak},
			{0, SRBA_W_valid}, {0, SRBA_W_lm_inv}, {0, SRBA_W_slot_listed}, {0, SRBA_W_listed_idx}, {0, SRBA_W_refresh_idx},
			{0, SRBA_W_resid0}, {0, SRBA_W_resid1}, {0, SRBA_W_slot}, {0, SRBA_W_HAp}, {0, SRBA_W_Hf}, {0, SRBA_W_Hf_inv},
			{1, SRBA_I_ldl_phase}, {1, SRBA_I_ldl_row}, {1, SRBA_I_ldl_tile_ij}, {1, SRBA_I_ldl_upd}, {1, SRBA_I_ldl_src}, {1, SRBA_I_ldl_orow}, {1, SRBA_I_unk_edge}, {1, SRBA_I_unk_lm}, {1, SRBA_I_obs_slot}, {1, SRBA_I_obs_z}, {1, SRBA_I_obs_mult}, {1, SRBA_I_obs_lm},
			{1, SRBA_I_pair_path_ptr}, {1, SRBA_I_path_edge}, {1, SRBA_I_path_dir},
			{0, SRBA_W_Sblk}, {0, SRBA_W_HApf}, {0, SRBA_W_W}, {0, SRBA_W_JA}, {0, SRBA_W_JF}, {0, SRBA_W_slot_bak}, {0, SRBA_W_blk_ok},
			{1, SRBA_I_colA_ptr}, {1, SRBA_I_blkA_obs}, {1, SRBA_I_blkA_col}, {1, SRBA_I_blkA_slotA}, {1, SRBA_I_blkA_slotD}, {1, SRBA_I_blkA_normal_dir},
			{1, SRBA_I_HAp_term_ptr}, {1, SRBA_I_HAp_term_a}, {1, SRBA_I_HAp_term_b}, {1, SRBA_I_HAp_row}, {1, SRBA_I_HAp_col},
			{1, SRBA_I_S_row}, {1, SRBA_I_S_col},
		};
		for (const Cand &c : order) {
			const uint64_t bytes = c.is_img ? h.img_len[c.id] : ws_bytes[c.id];
			if (!bytes) continue;
			const uint64_t start = (A.cur + 15) / 16 * 16;
			if (start + (bytes + 15) / 16 * 16 > budget) continue;
			const uint32_t off = (uint32_t)A.add((bytes + 15) / 16 * 16);
			if (c.is_img) h.img_sm[c.id] = off; else h.ws_sm[c.id] = off;
		}
		h.sm_bytes = (uint32_t)A.total();
	}

	Layout R;
	h.res_stats = R.add(sizeof(srba_stats_t));
	h.res_edge_pose = R.add((size_t)p.n_unk_edges * S * 8);
	h.res_lm_pos = R.add((size_t)p.n_unk_lms * 3 * 8);
	h.res_slot_pose = R.add((size_t)2 * p.n_pairs * S * 8);
	h.res_obs_valid = R.add(p.n_obs);
	h.res_lm_inv = R.add(p.n_unk_lms);
	h.res_Hf_diag = R.add((size_t)p.n_unk_lms * 9 * 8);
	h.res_dense = R.add(want_dense ? (size_t)h.n_dense * h.n_dense * 8 : 0);
	h.res_trace = R.add((size_t)trace_cap * 6 * 8);
	h.res_trace_delta = R.add((size_t)trace_cap * h.n_scalars * 8);
	h.res_stage_cycles = R.add(24 * 8);
	h.res_bytes = R.total();
	if (base) std::memcpy(base, &h, sizeof(h));
}

// Small persistent worker pool for the host side of a batch call (packing / unpacking 256 problems is as long as the
// kernel itself when done on one core).  parallel_for(n, f) runs f(0..n-1) on the caller + workers (dynamic chunks).
// SRBA_B200_THREADS overrides the thread count (default: the process's CPUs / LOCAL_WORLD_SIZE, at most 16).
std::atomic<bool> g_forked{false};  // worker threads do not survive fork(): a forked child packs serially
class WorkerPool {
	// Workers SPIN on an atomic generation counter for a short while after their last job (a batch call issues eight
	// parallel_for rounds within ~1 ms, and a condition-variable wake-up costs as much as packing a chunk), then sleep.
	std::vector<std::thread> th_;
	std::mutex m_, job_m_;
	std::condition_variable cv_;
	const std::function<void(int)> *fn_ = nullptr;
	int n_ = 0;
	std::atomic<int> next_{0}, active_{0}, sleepers_{0};
	std::atomic<uint64_t> gen_{0};
	std::atomic<bool> stop_{false};
	void work() {
		for (;;) {
			const int i = next_.fetch_add(1, std::memory_order_relaxed);
			if (i >= n_) break;
			(*fn_)(i);
		}
	}
	void loop() {
		uint64_t seen = 0;
		for (;;) {
			const auto t0 = std::chrono::steady_clock::now();
			int spins = 0;
			while (gen_.load(std::memory_order_acquire) == seen && !stop_.load(std::memory_order_relaxed)) {
				if (++spins < 2048) continue;
				spins = 0;
				if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(300)) {
					std::unique_lock<std::mutex> lk(m_);
					sleepers_.fetch_add(1);
					cv_.wait(lk, [&] { return stop_.load() || gen_.load() != seen; });
					sleepers_.fetch_sub(1);
				}
			}
			if (stop_.load()) return;
			seen = gen_.load(std::memory_order_acquire);
			work();
			active_.fetch_sub(1, std::memory_order_acq_rel);
		}
	}

public:
	WorkerPool() {
		// default: the CPUs this process may run on, shared evenly between the ranks of the node when launched one process
		// per GPU (torchrun exports LOCAL_WORLD_SIZE): spinning workers of 8 ranks must not oversubscribe the host cores
		int k = (int)std::thread::hardware_concurrency();
		cpu_set_t cs;
		CPU_ZERO(&cs);
		if (sched_getaffinity(0, sizeof(cs), &cs) == 0 && CPU_COUNT(&cs) > 0) k = CPU_COUNT(&cs);
		if (const char *lw = std::getenv("LOCAL_WORLD_SIZE")) { const int w = std::atoi(lw); if (w > 1) k = std::max(2, k / w); }
		if (const char *e = std::getenv("SRBA_B200_THREADS")) k = std::atoi(e);
		k = std::max(1, std::min(k, 16));
		for (int i = 1; i < k; i++) th_.emplace_back([this] { loop(); });
		pthread_atfork(nullptr, nullptr, [] { g_forked.store(true); });
	}
	~WorkerPool() {
		if (g_forked.load()) { for (auto &t : th_) t.detach(); return; }
		{
			std::lock_guard<std::mutex> lk(m_);
			stop_.store(true);
		}
		cv_.notify_all();
		for (auto &t : th_) t.join();
	}
	int threads() const { return (int)th_.size() + 1; }
	// Safe for concurrent callers: one job at a time owns the workers; a caller that finds the pool busy (another host
	// thread is inside a batch call) runs its own loop serially instead of waiting.
	void parallel_for(int n, const std::function<void(int)> &f) {
		if (n < 16 || th_.empty() || g_forked.load()) { for (int i = 0; i < n; i++) f(i); return; }
		std::unique_lock<std::mutex> job(job_m_, std::try_to_lock);
		if (!job.owns_lock()) { for (int i = 0; i < n; i++) f(i); return; }
		fn_ = &f; n_ = n; next_.store(0);
		active_.store((int)th_.size(), std::memory_order_release);
		gen_.fetch_add(1, std::memory_order_acq_rel);
		if (sleepers_.load() > 0) {
			std::lock_guard<std::mutex> lk(m_);
			cv_.notify_all();
		}
		work();
		while (active_.load(std::memory_order_acquire) != 0) { /* the stragglers finish their last item */ }
	}
};
WorkerPool &workers() {
	static WorkerPool p;
	return p;
}

double problem_cost(const DevProblem &h) {
	const double n = h.n_dense;
	return n * n * n / 3.0 + 40.0 * (h.n_blkA + h.n_blkF) + 100.0 * h.n_S_terms + 1000.0;
}
}  // namespace

// Grow-only buffer pool reused by the synchronous entry points (srba_b200_optimize / _optimize_batch): the library keeps
// no PROBLEM state between calls, only this reusable workspace (SURVEY.md §8b "optional reusable device workspace"),
// one set per calling host thread and CUDA device.  It is rebuilt when the thread's current device changes, released by
// srba_b200_release_workspace() and when the thread exits.
struct Pool {
	char *h_img = nullptr, *h_res = nullptr, *d_img = nullptr, *d_ws = nullptr, *d_res = nullptr;
	uint64_t h_img_cap = 0, h_res_cap = 0, d_img_cap = 0, d_ws_cap = 0, d_res_cap = 0;
	cudaStream_t stream = nullptr;
	int device = -1;
	void release() {
		if (device >= 0) {
			int cur = -1;
			const bool sw = cudaGetDevice(&cur) == cudaSuccess && cur != device && cudaSetDevice(device) == cudaSuccess;
			if (h_img) cudaFreeHost(h_img);
			if (h_res) cudaFreeHost(h_res);
			if (d_img) cudaFree(d_img);
			if (d_ws) cudaFree(d_ws);
			if (d_res) cudaFree(d_res);
			if (stream) cudaStreamDestroy(stream);
			if (sw) cudaSetDevice(cur);
			cudaGetLastError();  // a context already torn down at process exit is not an error worth keeping
		}
		h_img = h_res = d_img = d_ws = d_res = nullptr;
		h_img_cap = h_res_cap = d_img_cap = d_ws_cap = d_res_cap = 0;
		stream = nullptr;
		device = -1;
	}
};
#define SRBA_MAX_CHUNKS 8
struct PoolSet {
	Pool p[SRBA_MAX_CHUNKS];  // one buffer set per pipelined chunk of a batch call
	~PoolSet() { if (!g_forked.load()) for (Pool &q : p) q.release(); }
};
static thread_local PoolSet g_pools;
// the calling thread's pool for chunk c on its CURRENT device (buffers of another device are released first)
static Pool &pool_for(int c) {
	Pool &q = g_pools.p[c];
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
	if (q.device != dev) { q.release(); q.device = dev; }
	return q;
}

static bool pool_reserve_host(char **p, uint64_t *cap, uint64_t need) {
	if (*cap >= need) return true;
	if (*p) cudaFreeHost(*p);
	*p = nullptr; *cap = 0;
	const uint64_t want = need + need / 4 + 4096;
	if (cudaMallocHost(p, want) != cudaSuccess) return false;
	*cap = want;
	return true;
}
static bool pool_reserve_dev(char **p, uint64_t *cap, uint64_t need) {
	if (*cap >= need) return true;
	if (*p) cudaFree(*p);
	*p = nullptr; *cap = 0;
	const uint64_t want = need + need / 4 + 4096;
	if (cudaMalloc(p, want) != cudaSuccess) return false;
	*cap = want;
	return true;
}

struct srba_batch {
	int n = 0;
	bool pooled = false;
	std::vector<DevProblem> hdr;
	std::vector<TileSymbolic> sym;
	std::vector<uint64_t> img_off, ws_off, res_off;
	std::vector<int32_t> dirty_values;  // problems whose numeric inputs changed in the pinned image since the last upload (batch_update_values)
	std::vector<int32_t> order;      // grouped by model, cost-descending inside each group
	int group_begin[13] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // 4 models x {one CTA, 8-CTA cluster, grid-wide wave} per problem
	int group_max_n[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
	int group_smem[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // largest planned shared-memory arena of each group
	SrbaWaveDims wave_dims[4];      // per model: stage grids of the wave group
	uint64_t wave_scratch_off[4] = {0, 0, 0, 0};  // offset of the wave group's state / views / partials inside d_ws
	int *h_wave_active = nullptr;   // 2 pinned ints: areas still iterating (read back with one trial of lag)
	cudaStream_t side_stream = nullptr;  // one-CTA / cluster groups run here while the wave groups use the caller's stream
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
	uint64_t tab_img_off = 0, tab_ws_off = 0, tab_res_off = 0, tab_order = 0;
	uint64_t img_bytes = 0, ws_bytes = 0, res_bytes = 0;
	char *h_img = nullptr, *h_res = nullptr;
	char *d_img = nullptr, *d_ws = nullptr, *d_res = nullptr;
	int64_t launches = 0;
	bool uploaded = false, downloaded = false;
};

extern "C" {

const char *srba_b200_last_error(void) { return g_err.c_str(); }
int srba_b200_set_option(const char *name, int32_t value) {
	init_options();
	for (int i = 0; name && i < OPT_COUNT; i++)
		if (std::strcmp(name, k_opt_names[i]) == 0) { g_opt[i].store(value); return SRBA_OK; }
	return fail(SRBA_ERR_BAD_ARGUMENT, "unknown option");
}
void srba_b200_release_workspace(void) {
	for (Pool &q : g_pools.p) q.release();
}
int32_t srba_b200_get_option(const char *name) {
	init_options();
	for (int i = 0; name && i < OPT_COUNT; i++)
		if (std::strcmp(name, k_opt_names[i]) == 0) return g_opt[i].load();
	return INT32_MIN;
}
const char *srba_b200_version(void) { return "srba_b200 0.1 (sm_100a)"; }
int srba_b200_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

// Global diagnostics: RbaEngine::eval_overall_squared_error (impl/eval_overall_error.h:15-137), see include/srba_b200.h.
int srba_b200_eval_overall_squared_error(const srba_eval_problem_t *in, double *out_sqerr, double *out_pair_pose) {
	if (!in || !out_sqerr) return fail(SRBA_ERR_BAD_ARGUMENT, "null argument");
	const srba_eval_problem_t &p = *in;
	const srba_model_t &m = p.model;
	const bool se2 = m.pose_kind == SRBA_POSE_SE2;
	if (se2 ? (m.obs_kind != SRBA_OBS_RELPOSES2D || m.sensor_pose_kind != SRBA_SENSORPOSE_NONE)
	        : (m.pose_kind != SRBA_POSE_SE3 || m.obs_kind < SRBA_OBS_CARTESIAN3D || m.obs_kind > SRBA_OBS_STEREO))
		return fail(SRBA_ERR_UNSUPPORTED_MODEL, "model combination not built (SE2+RelativePoses_2D, SE3+Cartesian_3D/Monocular/Stereo)");
	if (p.n_edges < 0 || p.n_pairs < 0 || p.n_path < 0 || p.n_obs < 0 || p.n_lms < 0) return fail(SRBA_ERR_BAD_ARGUMENT, "negative count");
	if ((p.n_pairs > 0 && (!p.pair_path_ptr || !p.edge_pose)) || (p.n_path > 0 && (!p.path_edge || !p.path_dir)) ||
	    (p.n_obs > 0 && (!p.obs_z || !p.obs_pair || !p.obs_inverse)) || (!se2 && p.n_obs > 0 && (!p.obs_lm || !p.lm_pos)))
		return fail(SRBA_ERR_BAD_ARGUMENT, "null array");
	for (int t = 0; t < p.n_pairs; t++)
		if (p.pair_path_ptr[t] > p.pair_path_ptr[t + 1] || p.pair_path_ptr[t] < 0 || p.pair_path_ptr[t + 1] > p.n_path) return fail(SRBA_ERR_BAD_ARGUMENT, "pair_path_ptr is not a CSR row pointer");
	for (int k = 0; k < p.n_path; k++)
		if (p.path_edge[k] < 0 || p.path_edge[k] >= p.n_edges) return fail(SRBA_ERR_BAD_ARGUMENT, "path_edge out of range");
	for (int i = 0; i < p.n_obs; i++) {
		if (p.obs_pair[i] < -1 || p.obs_pair[i] >= p.n_pairs) return fail(SRBA_ERR_BAD_ARGUMENT, "obs_pair out of range");
		if (!se2 && (p.obs_lm[i] < 0 || p.obs_lm[i] >= p.n_lms)) return fail(SRBA_ERR_BAD_ARGUMENT, "obs_lm out of range");
	}
	if (srba_b200_device_count() <= 0) return fail(SRBA_ERR_NO_DEVICE, "no CUDA device (srba_b200 has no CPU fallback)");
	const int S = se2 ? 3 : 12, o = obs_dims(m.obs_kind);
	int dev = 0, sms = 148;
	CUDA_TRY(cudaGetDevice(&dev));
	CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	// grids: whole waves of the SM count, 256-thread CTAs, at most 8 CTAs per SM (grid-stride loops take the rest)
	auto grid_for = [&](int n) { const int need = (n + 255) / 256; return std::max(1, std::min(need, 8 * sms)); };
	const int grid_chain = grid_for(p.n_pairs), grid_resid = grid_for(p.n_obs);
	// one packed image, one H2D copy
	Layout L;
	const uint64_t o_edge = L.add((uint64_t)p.n_edges * S * 8), o_ptr = L.add((uint64_t)(p.n_pairs + 1) * 4), o_pe = L.add((uint64_t)p.n_path * 4),
	               o_pd = L.add(p.n_path), o_z = L.add((uint64_t)p.n_obs * o * 8), o_op = L.add((uint64_t)p.n_obs * 4), o_oi = L.add(p.n_obs),
	               o_ol = L.add((uint64_t)p.n_obs * 4), o_lm = L.add((uint64_t)p.n_lms * 3 * 8);
	const uint64_t in_bytes = L.total();
	const uint64_t o_pose = L.add((uint64_t)p.n_pairs * S * 8), o_part = L.add((uint64_t)grid_resid * 8), o_tot = L.add(8);
	const uint64_t all_bytes = L.total();
	char *h = nullptr, *d = nullptr;
	if (cudaMallocHost(&h, all_bytes) != cudaSuccess) return fail(SRBA_ERR_CUDA, "cudaMallocHost failed");
	if (cudaMalloc(&d, all_bytes) != cudaSuccess) { cudaFreeHost(h); return fail(SRBA_ERR_CUDA, "cudaMalloc failed"); }
	auto put = [&](uint64_t off, const void *src, uint64_t bytes) { if (bytes && src) std::memcpy(h + off, src, bytes); };
	put(o_edge, p.edge_pose, (uint64_t)p.n_edges * S * 8);
	if (p.n_pairs > 0) put(o_ptr, p.pair_path_ptr, (uint64_t)(p.n_pairs + 1) * 4);
	put(o_pe, p.path_edge, (uint64_t)p.n_path * 4); put(o_pd, p.path_dir, p.n_path);
	put(o_z, p.obs_z, (uint64_t)p.n_obs * o * 8); put(o_op, p.obs_pair, (uint64_t)p.n_obs * 4); put(o_oi, p.obs_inverse, p.n_obs);
	if (!se2) { put(o_ol, p.obs_lm, (uint64_t)p.n_obs * 4); put(o_lm, p.lm_pos, (uint64_t)p.n_lms * 3 * 8); }
	int rc = SRBA_OK;
	cudaError_t e = cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, 0);
	if (e == cudaSuccess)
		e = (cudaError_t)srba_launch_eval_error(&m, p.n_pairs, p.n_obs, (const double *)(d + o_edge), (const int32_t *)(d + o_ptr), (const int32_t *)(d + o_pe),
		                                        (const uint8_t *)(d + o_pd), (const double *)(d + o_z), (const int32_t *)(d + o_op), (const uint8_t *)(d + o_oi),
		                                        (const int32_t *)(d + o_ol), (const double *)(d + o_lm), (double *)(d + o_pose), (double *)(d + o_part),
		                                        (double *)(d + o_tot), grid_chain, grid_resid, nullptr);
	if (e == cudaSuccess) e = cudaMemcpyAsync(h + o_pose, d + o_pose, all_bytes - o_pose, cudaMemcpyDeviceToHost, 0);
	if (e == cudaSuccess) e = cudaStreamSynchronize(0);
	if (e != cudaSuccess) rc = fail(SRBA_ERR_CUDA, std::string("eval_overall_squared_error: ") + cudaGetErrorString(e));
	else {
		*out_sqerr = *reinterpret_cast<double *>(h + o_tot);
		if (out_pair_pose && p.n_pairs > 0) std::memcpy(out_pair_pose, h + o_pose, (uint64_t)p.n_pairs * S * 8);
	}
	cudaFree(d);
	cudaFreeHost(h);
	return rc;
}

void srba_b200_batch_destroy(srba_batch_t *b) {
	if (!b) return;
	if (b->h_wave_active) cudaFreeHost(b->h_wave_active);
	if (b->side_stream) cudaStreamDestroy(b->side_stream);
	if (b->ev_fork) cudaEventDestroy(b->ev_fork);
	if (b->ev_join) cudaEventDestroy(b->ev_join);
	if (b->pooled) { delete b; return; }
	if (b->h_img) cudaFreeHost(b->h_img);
	if (b->h_res) cudaFreeHost(b->h_res);
	if (b->d_img) cudaFree(b->d_img);
	if (b->d_ws) cudaFree(b->d_ws);
	if (b->d_res) cudaFree(b->d_res);
	delete b;
}

static srba_batch_t *batch_create_impl(int32_t n, const srba_problem_t *const *in, int32_t trace_cap, Pool *pool, int allow_cluster = -1);
srba_batch_t *srba_b200_batch_create(int32_t n, const srba_problem_t *const *in, int32_t trace_cap) { return batch_create_impl(n, in, trace_cap, nullptr); }
static srba_batch_t *batch_create_impl(int32_t n, const srba_problem_t *const *in, int32_t trace_cap, Pool *pool, int allow_cluster) {
	if (n <= 0 || !in) { fail(SRBA_ERR_BAD_ARGUMENT, "empty batch"); return nullptr; }
	if (srba_b200_device_count() <= 0) { fail(SRBA_ERR_NO_DEVICE, "no CUDA device: the srba_b200 hot path has no CPU fallback"); return nullptr; }
	for (int i = 0; i < n; i++)
		if (!in[i] || validate(*in[i], false) != SRBA_OK) {
			if (!in[i]) fail(SRBA_ERR_BAD_ARGUMENT, "null problem");
			return nullptr;
		}
	{  // structural checks (O(size) per problem) on the worker pool, before anything dereferences the caller's index arrays
		std::atomic<const char *> bad{nullptr};
		workers().parallel_for(n, [&](int i) {
			if (bad.load(std::memory_order_relaxed)) return;
			if (const char *msg = validate_structure(*in[i])) { const char *expect = nullptr; bad.compare_exchange_strong(expect, msg); }
		});
		if (bad.load()) { fail(SRBA_ERR_BAD_ARGUMENT, bad.load()); return nullptr; }
	}
	srba_batch_t *b = new srba_batch();
	b->n = n;
	b->hdr.resize(n);
	b->sym.resize(n);
	b->img_off.resize(n); b->ws_off.resize(n); b->res_off.resize(n); b->order.resize(n);
	const int want_dense = trace_cap < 0 ? 1 : 0;  // negative trace_cap: |trace_cap| trials of trace AND the dense Hessian
	const int tcap = trace_cap < 0 ? -trace_cap : trace_cap;
	Layout I;
	b->tab_img_off = I.add((uint64_t)n * 8);
	b->tab_ws_off = I.add((uint64_t)n * 8);
	b->tab_res_off = I.add((uint64_t)n * 8);
	b->tab_order = I.add((uint64_t)n * 4);
	uint64_t ws = 0, rs = 0;
	if (allow_cluster < 0) {  // decided for this batch alone
		int n_heavy = 0;
		for (int i = 0; i < n; i++) n_heavy += heavy_problem(*in[i]) ? 1 : 0;
		allow_cluster = clusters_pay(n_heavy) ? 1 : 0;
	}
	const bool ac = allow_cluster != 0;
	// (largest problems first: the symbolic schedule of a 600-edge loop-closure area takes tens of ms, started last it would be the tail of the pool)
	std::vector<int32_t> by_cost(n);
	std::iota(by_cost.begin(), by_cost.end(), 0);
	std::stable_sort(by_cost.begin(), by_cost.end(), [&](int a, int c) {
		return (int64_t)in[a]->n_unk_edges * in[a]->n_unk_edges + in[a]->n_blkA > (int64_t)in[c]->n_unk_edges * in[c]->n_unk_edges + in[c]->n_blkA;
	});
	workers().parallel_for(n, [&](int q) { const int i = by_cost[q]; pack_problem(*in[i], b->hdr[i], nullptr, tcap, want_dense, b->sym[i], ac); });  // layout + symbolic
	for (int i = 0; i < n; i++) {
		b->img_off[i] = I.add(b->hdr[i].img_bytes);
		b->ws_off[i] = ws; ws += b->hdr[i].ws_bytes;
		b->res_off[i] = rs; rs += b->hdr[i].res_bytes;
	}
	std::iota(b->order.begin(), b->order.end(), 0);
	auto model_of = [&](int i) { return model_index(b->hdr[i].model) + 4 * b->hdr[i].cluster + 8 * b->hdr[i].wave; };
	std::stable_sort(b->order.begin(), b->order.end(), [&](int a, int c) {
		const int ma = model_of(a), mc = model_of(c);
		if (ma != mc) return ma < mc;
		return problem_cost(b->hdr[a]) > problem_cost(b->hdr[c]);
	});
	{
		int cnt[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
		for (int i = 0; i < n; i++) { const int g = model_of(i); cnt[g]++; b->group_max_n[g] = std::max(b->group_max_n[g], b->hdr[i].n_dense); b->group_smem[g] = std::max(b->group_smem[g], (int)b->hdr[i].sm_bytes); }
		for (int g = 0; g < 12; g++) b->group_begin[g + 1] = b->group_begin[g] + cnt[g];
		// wave groups: stage grids sized for the largest area, scratch (views, LM state, partials) appended to the workspace
		ws = (ws + 255) / 256 * 256;
		for (int m = 0; m < 4; m++) {
			SrbaWaveDims &d = b->wave_dims[m];
			std::memset(&d, 0, sizeof(d));
			d.n_areas = cnt[8 + m];
			if (!d.n_areas) continue;
			for (int q = b->group_begin[8 + m]; q < b->group_begin[9 + m]; q++) {
				const DevProblem &h = b->hdr[b->order[q]];
				const int P = m == 0 ? 3 : 6, S = m == 0 ? 3 : 12;
				auto mx = [](int &a, int64_t v) { a = (int)std::max<int64_t>(a, std::min<int64_t>(v, INT32_MAX)); };
				mx(d.max_pairs, h.n_pairs); mx(d.max_blocks, (int64_t)h.n_blkA + h.n_blkF);
				mx(d.max_hess_items, P == 6 ? (int64_t)6 * h.n_HAp + 6 * h.n_HApf + 3 * h.n_unk_lms : (int64_t)9 * h.n_HAp);
				mx(d.max_obs, h.n_obs); mx(d.max_scalars, h.n_scalars);
				mx(d.max_unk, std::max<int64_t>(std::max(h.n_unk_edges, h.n_unk_lms), (int64_t)2 * h.n_pairs * S));
				mx(d.max_schur2, (int64_t)h.n_HApf * P * 3); mx(d.max_schur3, (int64_t)h.n_S * (P == 6 ? 32 : 9) + h.n_dense); mx(d.max_lms, h.n_unk_lms);  // (x 32: room for the one-warp-per-block DMMA build, SRBA_SCHUR_DMMA)
				mx(d.max_edges_S, std::max<int64_t>((int64_t)h.n_edges * S, 2 * (int64_t)h.n_pairs));
				mx(d.max_out, (int64_t)2 * h.n_pairs * S);
				if (h.n_unk_lms > 0) d.has_lms = 1;
				const TileSymbolic &sy = b->sym[b->order[q]];
				if (sy.tail_level > 0) { if (sy.n_dst >= SRBA_WAVE_TEAM_DST_PER_LEVEL * sy.tail_level) d.n_big++; else d.n_solo++; }
				mx(d.max_tail_rows, (int64_t)sy.ntr - sy.phase[2 * (size_t)sy.tail_level]);
				if (sy.tail_level > 0) mx(d.max_bwd_smem, (int64_t)h.n_dense * 8 + (int64_t)sy.phase[2 * (size_t)sy.tail_level] * 8 + ((int64_t)sy.tail_level + 1) * 4 + 16);
			}
			b->wave_scratch_off[m] = ws;
			ws += (srba_wave_scratch_bytes(d.n_areas) + 255) / 256 * 256;
		}
	}
	b->img_bytes = I.total(); b->ws_bytes = std::max<uint64_t>(ws, 16); b->res_bytes = std::max<uint64_t>(rs, 16);
	bool ok;
	if (pool) {
		b->pooled = true;
		ok = pool_reserve_host(&pool->h_img, &pool->h_img_cap, b->img_bytes) && pool_reserve_host(&pool->h_res, &pool->h_res_cap, b->res_bytes) &&
		     pool_reserve_dev(&pool->d_img, &pool->d_img_cap, b->img_bytes) && pool_reserve_dev(&pool->d_ws, &pool->d_ws_cap, b->ws_bytes) &&
		     pool_reserve_dev(&pool->d_res, &pool->d_res_cap, b->res_bytes);
		b->h_img = pool->h_img; b->h_res = pool->h_res; b->d_img = pool->d_img; b->d_ws = pool->d_ws; b->d_res = pool->d_res;
	} else {
		ok = cudaMallocHost(&b->h_img, b->img_bytes) == cudaSuccess && cudaMallocHost(&b->h_res, b->res_bytes) == cudaSuccess &&
		     cudaMalloc(&b->d_img, b->img_bytes) == cudaSuccess && cudaMalloc(&b->d_ws, b->ws_bytes) == cudaSuccess &&
		     cudaMalloc(&b->d_res, b->res_bytes) == cudaSuccess;
	}
	if (!ok) {
		fail(SRBA_ERR_CUDA, std::string("allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
		srba_b200_batch_destroy(b);
		return nullptr;
	}
	std::memcpy(b->h_img + b->tab_img_off, b->img_off.data(), (size_t)n * 8);
	std::memcpy(b->h_img + b->tab_ws_off, b->ws_off.data(), (size_t)n * 8);
	std::memcpy(b->h_img + b->tab_res_off, b->res_off.data(), (size_t)n * 8);
	std::memcpy(b->h_img + b->tab_order, b->order.data(), (size_t)n * 4);
	workers().parallel_for(n, [&](int q) { const int i = by_cost[q]; pack_problem(*in[i], b->hdr[i], b->h_img + b->img_off[i], tcap, want_dense, b->sym[i], ac); });  // copy
	return b;
}

int srba_b200_batch_upload(srba_batch_t *b, void *stream) {
	if (!b) return fail(SRBA_ERR_BAD_ARGUMENT, "null batch");
	CUDA_TRY(cudaMemcpyAsync(b->d_img, b->h_img, b->img_bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
	b->uploaded = true;
	return SRBA_OK;
}

int srba_b200_batch_update_values(srba_batch_t *b, int32_t i, const double *edge_pose, const double *lm_pos, const double *obs_z) {
	if (!b || i < 0 || i >= b->n) return fail(SRBA_ERR_BAD_ARGUMENT, "bad batch_update_values arguments");
	const DevProblem &h = b->hdr[i];
	char *base = b->h_img + b->img_off[i];
	const double *srcs[3] = {edge_pose, lm_pos, obs_z};
	const int ids[3] = {SRBA_I_edge_pose, SRBA_I_lm_pos, SRBA_I_obs_z};
	for (int k = 0; k < 3; k++) {
		if (!srcs[k] || !h.img_len[ids[k]]) continue;
		const size_t n = h.img_len[ids[k]] / 8;
		for (size_t j = 0; j < n; j++)
			if (!std::isfinite(srcs[k][j])) return fail(SRBA_ERR_NON_FINITE, "batch_update_values: non-finite input value");
		std::memcpy(base + h.img_off[ids[k]], srcs[k], h.img_len[ids[k]]);
		b->dirty_values.push_back(i);
	}
	return SRBA_OK;
}

int srba_b200_batch_upload_values(srba_batch_t *b, void *stream) {
	if (!b) return fail(SRBA_ERR_BAD_ARGUMENT, "null batch");
	if (!b->uploaded) return srba_b200_batch_upload(b, stream);   // nothing resident yet: the whole image
	std::sort(b->dirty_values.begin(), b->dirty_values.end());
	b->dirty_values.erase(std::unique(b->dirty_values.begin(), b->dirty_values.end()), b->dirty_values.end());
	if (b->dirty_values.size() > 2048) {  // (thousands of small copies cost more than one large one)
		b->dirty_values.clear();
		return srba_b200_batch_upload(b, stream);
	}
	const int ids[3] = {SRBA_I_edge_pose, SRBA_I_lm_pos, SRBA_I_obs_z};
	for (int32_t i : b->dirty_values) {
		const DevProblem &h = b->hdr[i];
		for (int k = 0; k < 3; k++) {
			if (!h.img_len[ids[k]]) continue;
			const uint64_t off = b->img_off[i] + h.img_off[ids[k]];
			CUDA_TRY(cudaMemcpyAsync(b->d_img + off, b->h_img + off, h.img_len[ids[k]], cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
		}
	}
	b->dirty_values.clear();
	return SRBA_OK;
}

int srba_b200_batch_run(srba_batch_t *b, void *stream) {
	if (!b) return fail(SRBA_ERR_BAD_ARGUMENT, "null batch");
	if (!b->uploaded) return fail(SRBA_ERR_BAD_ARGUMENT, "batch_run before batch_upload");
	b->launches = 0;
	cudaStream_t main_st = static_cast<cudaStream_t>(stream);
	const bool any_wave = b->group_begin[12] > b->group_begin[8], any_cta = b->group_begin[8] > 0;
	// The wave groups are driven from the host (one launch sequence per LM trial) on the caller's stream; the one-CTA / cluster
	// groups of the same batch run meanwhile on a side stream, forked from and joined back into the caller's stream.
	cudaStream_t cta_st = main_st;
	if (any_wave && any_cta) {
		if (!b->side_stream) {
			CUDA_TRY(cudaStreamCreateWithFlags(&b->side_stream, cudaStreamNonBlocking));
			CUDA_TRY(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
			CUDA_TRY(cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming));
		}
		cta_st = b->side_stream;
		CUDA_TRY(cudaEventRecord(b->ev_fork, main_st));
		CUDA_TRY(cudaStreamWaitEvent(cta_st, b->ev_fork, 0));
	}
	const uint64_t *t_img = reinterpret_cast<const uint64_t *>(b->d_img + b->tab_img_off), *t_ws = reinterpret_cast<const uint64_t *>(b->d_img + b->tab_ws_off),
	               *t_res = reinterpret_cast<const uint64_t *>(b->d_img + b->tab_res_off);
	const int32_t *t_order = reinterpret_cast<const int32_t *>(b->d_img + b->tab_order);
	for (int g = 0; g < 8; g++) {
		const int cnt = b->group_begin[g + 1] - b->group_begin[g];
		if (!cnt) continue;
		// one CTA (or one cluster) per problem, dispatched in cost-descending order (largest first)
		const int rc = srba_launch_lm_batch(g, b->d_img, t_img, b->d_ws, t_ws, b->d_res, t_res, t_order + b->group_begin[g], cnt, cnt, b->group_smem[g], cta_st);
		if (rc != 0) return fail(SRBA_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString((cudaError_t)rc));
		b->launches++;
	}
	if (any_wave && any_cta) CUDA_TRY(cudaEventRecord(b->ev_join, cta_st));
	for (int m = 0; m < 4; m++) {
		if (!b->wave_dims[m].n_areas) continue;
		if (!b->h_wave_active) CUDA_TRY(cudaMallocHost(&b->h_wave_active, 2 * sizeof(int)));
		const int rc = srba_launch_lm_wave(m, b->d_img, t_img, b->d_ws, t_ws, b->d_res, t_res, t_order + b->group_begin[8 + m], &b->wave_dims[m],
		                                   b->d_ws + b->wave_scratch_off[m], b->h_wave_active, main_st, &b->launches);
		if (rc != 0) return fail(SRBA_ERR_CUDA, std::string("wave pipeline: ") + cudaGetErrorString((cudaError_t)rc));
	}
	if (any_wave && any_cta) CUDA_TRY(cudaStreamWaitEvent(main_st, b->ev_join, 0));
	return SRBA_OK;
}

int srba_b200_batch_download(srba_batch_t *b, void *stream) {
	if (!b) return fail(SRBA_ERR_BAD_ARGUMENT, "null batch");
	CUDA_TRY(cudaMemcpyAsync(b->h_res, b->d_res, b->res_bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
	CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
	b->downloaded = true;
	return SRBA_OK;
}

int srba_b200_batch_unpack(srba_batch_t *b, int32_t i, srba_result_t *out) {
	if (!b || !out || i < 0 || i >= b->n) return fail(SRBA_ERR_BAD_ARGUMENT, "bad unpack arguments");
	if (!b->downloaded) return fail(SRBA_ERR_BAD_ARGUMENT, "batch_unpack before batch_download");
	const DevProblem &h = b->hdr[i];
	const char *r = b->h_res + b->res_off[i];
	const int S = h.model.pose_kind == SRBA_POSE_SE2 ? 3 : 12;
	std::memcpy(&out->stats, r + h.res_stats, sizeof(srba_stats_t));
	if (out->edge_pose) std::memcpy(out->edge_pose, r + h.res_edge_pose, (size_t)h.n_unk_edges * S * 8);
	if (out->lm_pos) std::memcpy(out->lm_pos, r + h.res_lm_pos, (size_t)h.n_unk_lms * 3 * 8);
	if (out->slot_pose) std::memcpy(out->slot_pose, r + h.res_slot_pose, (size_t)2 * h.n_pairs * S * 8);
	if (out->obs_valid) std::memcpy(out->obs_valid, r + h.res_obs_valid, h.n_obs);
	if (out->lm_invertible) std::memcpy(out->lm_invertible, r + h.res_lm_inv, h.n_unk_lms);
	if (out->Hf_diag) std::memcpy(out->Hf_diag, r + h.res_Hf_diag, (size_t)h.n_unk_lms * 9 * 8);
	if (out->dense_hessian && h.want_dense) std::memcpy(out->dense_hessian, r + h.res_dense, (size_t)h.n_dense * h.n_dense * 8);
	const int tl = std::min<int>(std::min<int>(out->stats.num_trials, h.trace_cap), out->trace_cap);
	out->trace_len = tl > 0 ? tl : 0;
	if (out->trace && tl > 0) std::memcpy(out->trace, r + h.res_trace, (size_t)tl * 6 * 8);
	if (out->trace_delta && tl > 0) std::memcpy(out->trace_delta, r + h.res_trace_delta, (size_t)tl * h.n_scalars * 8);
	if (out->stats.status != SRBA_OK) g_err = "device-side error status (srba_stats_t::status)";
	return out->stats.status;
}

int srba_b200_batch_stage_cycles(srba_batch_t *b, int32_t i, uint64_t out[24]) {
	if (!b || i < 0 || i >= b->n || !b->downloaded) return fail(SRBA_ERR_BAD_ARGUMENT, "bad stage_cycles arguments");
	std::memcpy(out, b->h_res + b->res_off[i] + b->hdr[i].res_stage_cycles, 24 * 8);
	return SRBA_OK;
}
int64_t srba_b200_batch_h2d_bytes(const srba_batch_t *b) { return b ? (int64_t)b->img_bytes : 0; }
int64_t srba_b200_batch_d2h_bytes(const srba_batch_t *b) { return b ? (int64_t)b->res_bytes : 0; }
int64_t srba_b200_batch_kernel_launches(const srba_batch_t *b) { return b ? b->launches : 0; }

void srba_b200_batch_algorithmic_bytes(const srba_batch_t *b, double out[5]) {
	for (int k = 0; k < 5; k++) out[k] = 0;
	if (!b) return;
	for (int i = 0; i < b->n; i++) {
		const DevProblem &h = b->hdr[i];
		const double P = h.model.pose_kind == SRBA_POSE_SE2 ? 3 : 6, S = P == 3 ? 24 : 96, f = 3;
		const double o = obs_dims(h.model.obs_kind);
		const double NT = h.n_pairs, Lbar_NT = h.n_path;
		// unique-touch model, FP64 values + 4-byte indices
		out[0] += NT * 4 + Lbar_NT * 5 + h.n_edges * S + 2 * NT * S;
		out[1] += h.n_obs * (o * 8 + 8) + h.n_blkA * 20.0 + h.n_blkF * 8.0 + 2 * NT * S + h.n_lms * f * 8 +
		          (h.n_HAp * P * P + h.n_HApf * P * f + h.n_unk_lms * f * f) * 8 + (h.n_unk_edges * P + h.n_unk_lms * f) * 8 + h.n_obs * (o * 8 + 1);
		out[2] += (h.n_HApf * P * f + 2 * h.n_unk_lms * f * f + h.n_HAp * P * P) * 8 + 2.0 * h.n_S * P * P * 8 + (h.n_unk_edges * P + h.n_unk_lms * f) * 8 * 2;
		const double n = h.n_dense;
		out[3] += n * n * 8;
		out[4] += n * n * n / 3.0 + 2 * n * n;
	}
}

int srba_b200_debug_ldl_schedule(const srba_problem_t *in, int32_t counts[6], int32_t *row, int32_t *tile_ij, int32_t *src, int32_t *upd, int32_t *orow,
                                 int32_t *phase, int32_t cap_tiles, int32_t cap_upd, int32_t max_tiles_on_chip) {
	if (!in || !counts) return fail(SRBA_ERR_BAD_ARGUMENT, "null argument");
	const int rc = validate(*in);
	if (rc != SRBA_OK) return rc;
	const int P = in->model.pose_kind == SRBA_POSE_SE2 ? 3 : 6;
	TileSymbolic sy;
	symbolic_tiles(*in, P, max_tiles_on_chip, sy);
	counts[0] = sy.ntr; counts[1] = (int32_t)sy.tile_ij.size(); counts[2] = (int32_t)(sy.upd.size() / (sy.wide ? 3 : 2)); counts[3] = (sy.ok && !sy.wide) ? 1 : 0;
	counts[4] = sy.n_phases; counts[5] = sy.on_chip ? 1 : 0;
	if (!sy.ok) return SRBA_OK;
	if (row && tile_ij && src && upd) {
		if (cap_tiles < counts[1] || cap_upd < counts[2]) return fail(SRBA_ERR_BAD_ARGUMENT, "schedule buffers too small");
		std::memcpy(row, sy.row.data(), sy.row.size() * 4);
		std::memcpy(tile_ij, sy.tile_ij.data(), sy.tile_ij.size() * 4);
		std::memcpy(src, sy.src.data(), sy.src.size() * 4);
		std::memcpy(upd, sy.upd.data(), sy.upd.size() * 4);
		if (orow) std::memcpy(orow, sy.orow.data(), sy.orow.size() * 4);
		if (phase) std::memcpy(phase, sy.phase.data(), sy.phase.size() * 4);
	}
	return SRBA_OK;
}

int srba_b200_debug_level_schedule(const srba_problem_t *in, int32_t counts[6], int32_t *row2, int32_t *tile_ij, int32_t *src, int32_t *orow, int32_t *lv,
                                   int32_t *gu) {
	if (!in || !counts) return fail(SRBA_ERR_BAD_ARGUMENT, "null argument");
	const int rc = validate(*in);
	if (rc != SRBA_OK) return rc;
	const int P = in->model.pose_kind == SRBA_POSE_SE2 ? 3 : 6;
	TileSymbolic sy;
	symbolic_levels(*in, P, sy, true);
	counts[0] = sy.ntr; counts[1] = (int32_t)sy.tile_ij.size(); counts[2] = sy.n_phases; counts[3] = sy.n_dst; counts[4] = sy.n_contrib; counts[5] = sy.ok ? 1 : 0;
	if (!sy.ok) return SRBA_OK;
	if (row2) for (int k = 0; k <= sy.ntr; k++) { row2[2 * k] = sy.row[4 * (size_t)k]; row2[2 * k + 1] = sy.row[4 * (size_t)k + 1]; }
	if (tile_ij) std::memcpy(tile_ij, sy.tile_ij.data(), sy.tile_ij.size() * 4);
	if (src) std::memcpy(src, sy.src.data(), sy.src.size() * 4);
	if (orow) std::memcpy(orow, sy.orow.data(), sy.orow.size() * 4);
	if (lv) std::memcpy(lv, sy.phase.data(), sy.phase.size() * 4);
	if (gu) {  // without the alignment pad of the device image
		const size_t head = 2 * (size_t)sy.n_dst + 1;
		std::memcpy(gu, sy.upd.data(), head * 4);
		std::memcpy(gu + head, sy.upd.data() + head + 1, 2 * (size_t)sy.n_contrib * 4);
	}
	return SRBA_OK;
}

void srba_b200_kernel_timing(int32_t enable) { srba_wave_timing_enable(enable); }
void srba_b200_kernel_timing_get(double out[2]) {
	long n = 0;
	srba_wave_timing_get(&out[0], &n);
	out[1] = (double)n;
}

int srba_b200_problem_plan(const srba_problem_t *in, int64_t out[8]) {
	if (!in || !out) return fail(SRBA_ERR_BAD_ARGUMENT, "null argument");
	const int rc = validate(*in);
	if (rc != SRBA_OK) return rc;
	DevProblem h;
	TileSymbolic sy;
	pack_problem(*in, h, nullptr, 0, 0, sy, true);
	int n_img = 0, n_ws = 0;
	for (int a = 0; a < SRBA_I_COUNT; a++) n_img += (h.img_sm[a] != SRBA_NOT_IN_SMEM && h.img_len[a]) ? 1 : 0;
	for (int a = 0; a < SRBA_W_COUNT; a++) n_ws += h.ws_sm[a] != SRBA_NOT_IN_SMEM ? 1 : 0;
	out[0] = (int64_t)h.img_bytes; out[1] = (int64_t)h.ws_bytes; out[2] = (int64_t)h.res_bytes; out[3] = h.sm_bytes;
	out[4] = h.chol_mode + 2 * h.cluster + 4 * h.wave; out[5] = n_img; out[6] = n_ws; out[7] = srba_kernel_cfg(model_index(in->model) + 4 * h.cluster).smem_bytes;
	return SRBA_OK;
}

int srba_b200_batch_problem_plan(const srba_batch_t *b, int32_t i, int64_t out[8]) {
	if (!b || !out || i < 0 || i >= b->n) return fail(SRBA_ERR_BAD_ARGUMENT, "bad arguments");
	const DevProblem &h = b->hdr[i];
	int n_img = 0, n_ws = 0;
	for (int a = 0; a < SRBA_I_COUNT; a++) n_img += (h.img_sm[a] != SRBA_NOT_IN_SMEM && h.img_len[a]) ? 1 : 0;
	for (int a = 0; a < SRBA_W_COUNT; a++) n_ws += h.ws_sm[a] != SRBA_NOT_IN_SMEM ? 1 : 0;
	out[0] = (int64_t)h.img_bytes; out[1] = (int64_t)h.ws_bytes; out[2] = (int64_t)h.res_bytes; out[3] = h.sm_bytes;
	out[4] = h.chol_mode + 2 * h.cluster + 4 * h.wave; out[5] = n_img; out[6] = n_ws; out[7] = srba_kernel_cfg(model_index(h.model) + 4 * h.cluster).smem_bytes;
	return SRBA_OK;
}

int srba_b200_problem_ldl_counts(const srba_problem_t *in, int64_t out[4]) {
	if (!in || !out) return fail(SRBA_ERR_BAD_ARGUMENT, "null argument");
	const int rc = validate(*in);
	if (rc != SRBA_OK) return rc;
	DevProblem h;
	TileSymbolic sy;
	pack_problem(*in, h, nullptr, 0, 0, sy, true);
	const bool reg = h.chol_mode == SRBA_CHOL_REG;
	out[0] = reg ? (int64_t)sy.tile_ij.size() : 0;
	out[1] = reg ? (sy.levels ? (int64_t)sy.n_contrib : (int64_t)(sy.upd.size() / (sy.wide ? 3 : 2))) : 0;
	out[2] = reg ? sy.n_phases : 0;
	out[3] = (reg && !h.ldl_tiles_in_ws) ? 1 : 0;
	return SRBA_OK;
}

int srba_b200_problem_wave_counts(const srba_problem_t *in, int64_t out[8]) {
	if (!in || !out) return fail(SRBA_ERR_BAD_ARGUMENT, "null argument");
	const int rc = validate(*in);
	if (rc != SRBA_OK) return rc;
	DevProblem h;
	TileSymbolic sy;
	pack_problem(*in, h, nullptr, 0, 0, sy, true);
	for (int k = 0; k < 8; k++) out[k] = 0;
	if (!h.wave || !sy.levels) return SRBA_OK;
	const int tail_row = sy.phase[2 * (size_t)sy.tail_level];
	out[0] = (int64_t)sy.tile_ij.size(); out[1] = sy.n_dst; out[2] = sy.n_contrib; out[3] = sy.n_phases; out[4] = sy.tail_level;
	out[5] = (sy.tail_level > 0 && sy.n_dst >= SRBA_WAVE_TEAM_DST_PER_LEVEL * sy.tail_level) ? 1 : 0;
	out[6] = sy.rs[tail_row];                       // tiles of the rows below the tail
	out[7] = sy.phase[2 * (size_t)sy.tail_level + 1];  // destination entries of the levels below the tail
	return SRBA_OK;
}

int srba_b200_optimize_batch(int32_t n, const srba_problem_t *const *in, srba_result_t *const *out) {
	if (!out) return fail(SRBA_ERR_BAD_ARGUMENT, "null out");
	int tcap = 0, want_dense = 0;
	for (int i = 0; i < n; i++) {
		if (!out[i]) return fail(SRBA_ERR_BAD_ARGUMENT, "null result");
		tcap = std::max(tcap, out[i]->trace ? out[i]->trace_cap : 0);
		if (out[i]->dense_hessian) want_dense = 1;
	}
	if (want_dense && tcap == 0) tcap = 1;
	static const bool timing = std::getenv("SRBA_B200_TIMING") != nullptr;  // host-side phase timings on stderr
	auto now = [] { return std::chrono::steady_clock::now(); };
	auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point c) { return std::chrono::duration<double, std::milli>(c - a).count(); };
	const auto t0 = now();
	// A large batch is cut into up to SRBA_MAX_CHUNKS chunks, each with its own stream and buffers: while chunk c is copied
	// and solved, chunk c+1 is being packed on the host cores, and results come back chunk by chunk.  All chunks fit the
	// GPU together (2 CTAs/SM), so the kernels overlap; only the first pack and the last copy-back stay exposed.
	const int k_chunks = std::max(1, std::min(option(OPT_CHUNKS), SRBA_MAX_CHUNKS));  // default 4 pipelined chunks
	int n_heavy = 0, n_wave = 0;
	for (int i = 0; i < n; i++) {
		n_heavy += (in[i] && heavy_problem(*in[i])) ? 1 : 0;
		n_wave += (in[i] && wave_candidate(*in[i])) ? 1 : 0;
	}
	// Areas of the grid-wide wave pipeline advance in lock step, one host-driven launch sequence per LM trial for the whole group: cutting
	// such a batch into chunks would repeat that sequence per chunk (and the host blocks inside each), so it stays one piece.
	const int K = (n >= 64 && n_wave == 0) ? k_chunks : 1;
	const int allow_cluster = clusters_pay(n_heavy) ? 1 : 0;
	srba_batch_t *bt[SRBA_MAX_CHUNKS] = {};
	int begin[SRBA_MAX_CHUNKS + 1];
	for (int c = 0; c <= K; c++) begin[c] = (int)((int64_t)n * c / K);
	int rc = SRBA_OK;
	double t_pack = 0;
	for (int c = 0; c < K && rc == SRBA_OK; c++) {
		Pool &pool = pool_for(c);
		const auto tp = now();
		bt[c] = batch_create_impl(begin[c + 1] - begin[c], in + begin[c], want_dense ? -tcap : tcap, &pool, allow_cluster);
		t_pack += ms(tp, now());
		if (!bt[c]) { rc = g_code != SRBA_OK ? g_code : SRBA_ERR_INTERNAL; break; }
		if (!pool.stream && cudaStreamCreateWithFlags(&pool.stream, cudaStreamNonBlocking) != cudaSuccess) { rc = fail(SRBA_ERR_CUDA, "cudaStreamCreate failed"); break; }
		rc = srba_b200_batch_upload(bt[c], pool.stream);
		if (rc == SRBA_OK) rc = srba_b200_batch_run(bt[c], pool.stream);
		if (rc == SRBA_OK && cudaMemcpyAsync(bt[c]->h_res, bt[c]->d_res, bt[c]->res_bytes, cudaMemcpyDeviceToHost, pool.stream) != cudaSuccess)
			rc = fail(SRBA_ERR_CUDA, "cudaMemcpyAsync (results) failed");
	}
	const auto t2 = now();
	double t_wait = 0, t_unpack = 0;
	for (int c = 0; c < K; c++) {
		if (!bt[c]) continue;
		const auto tw = now();
		if (cudaStreamSynchronize(g_pools.p[c].stream) != cudaSuccess && rc == SRBA_OK) rc = fail(SRBA_ERR_CUDA, std::string("kernel execution: ") + cudaGetErrorString(cudaGetLastError()));
		const auto tu = now();
		t_wait += ms(tw, tu);
		if (rc == SRBA_OK) {
			bt[c]->downloaded = true;
			std::atomic<int> first_bad{SRBA_OK};
			srba_batch_t *b = bt[c];
			srba_result_t *const *o = out + begin[c];
			workers().parallel_for(b->n, [&](int i) {
				const int r = srba_b200_batch_unpack(b, i, o[i]);
				if (r != SRBA_OK) { int expect = SRBA_OK; first_bad.compare_exchange_strong(expect, r); }
			});
			if (first_bad.load() != SRBA_OK)
				rc = fail(first_bad.load(), "a problem of the batch ended with a device-side error (srba_stats_t::status; SRBA_ERR_NON_FINITE: NaN/Inf inputs or a diverged iterate)");
		}
		t_unpack += ms(tu, now());
		srba_b200_batch_destroy(bt[c]);
	}
	if (timing)
		std::fprintf(stderr, "[srba_b200] optimize_batch n=%d in %d chunk(s): total %.3f ms (pack %.3f ms, enqueue window %.3f ms, waits %.3f ms, unpack %.3f ms)\n", n, K,
		             ms(t0, now()), t_pack, ms(t0, t2), t_wait, t_unpack);
	return rc;
}

int srba_b200_optimize(const srba_problem_t *in, srba_result_t *out) { return srba_b200_optimize_batch(1, &in, &out); }

}  // extern "C"
