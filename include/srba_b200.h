/* srba_b200.h — C ABI of the B200-native SRBA local Levenberg–Marquardt hot path.
 *
 * The reference (MRPT/srba) has no FFI: its optimiser is the protected member template
 * RbaEngine<>::optimize_edges (include/srba/RbaEngine.h:533-538, defined in include/srba/impl/optimize_edges.h:44-793)
 * reached from define_new_keyframe (impl/define_new_keyframe.h:80,97) and optimize_local_area
 * (impl/optimize_local_area.h:54).  The drop-in seam is therefore: a replacement `srba/impl/optimize_edges.h` that
 * flattens the selected sub-problem into the index-addressed arrays below, calls srba_b200_optimize(), and scatters the
 * result back.  One call == one optimize_edges() == steps D..S of SURVEY.md §3.3.
 *
 * Conventions: caller owns every buffer; plain pointers + sizes; no exceptions cross the boundary (negative return
 * codes, message via srba_b200_last_error()).  All reals are IEEE double, all local indices int32.
 * State: the library keeps no PROBLEM state between calls except what a batch / area handle owns.  The synchronous entry
 * points reuse a grow-only workspace (pinned + device buffers, streams) per calling host thread and CUDA device; it is
 * rebuilt when the thread's current device changes and freed by srba_b200_release_workspace() or at thread exit.
 * Threading: every entry point may be called concurrently from several host threads (the reference is single-threaded
 * per engine, RbaEngine.h:623,628; distinct engines are independent).  Error strings and codes are per thread.  The
 * internal pack/unpack worker pool serves one call at a time; a concurrent caller packs on its own thread.
 * Arguments are validated structurally (counts, NULL arrays, CSR pointers, index ranges) before anything is dereferenced
 * on the device: a malformed problem yields SRBA_ERR_BAD_ARGUMENT for the call, never a fault.
 *
 * Pose layout ("S" doubles):  SE(2): x,y,phi (S=3)  |  SE(3): R row-major (9) then t (3) (S=12).
 */
#ifndef SRBA_B200_H
#define SRBA_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Model tags == the reference's compile-time template arguments (RbaEngine.h:66, RBA_OPTIONS :39-45). */
enum { SRBA_POSE_SE2 = 0, SRBA_POSE_SE3 = 1 };                                   /* models/kf2kf_poses.h:26-38 */
enum { SRBA_OBS_RELPOSES2D = 0, SRBA_OBS_CARTESIAN3D = 1, SRBA_OBS_MONO = 2, SRBA_OBS_STEREO = 3 }; /* models/sensors.h */
enum { SRBA_SENSORPOSE_NONE = 0, SRBA_SENSORPOSE_SE3 = 1 };                     /* srba_options_sensor_pose.h:32,92 */
enum { SRBA_NOISE_IDENTITY = 0, SRBA_NOISE_CONST_MATRIX = 1 };                  /* srba_options_noise.h:25,79 */
/* RBA_OPTIONS::solver_t (srba_options_solver.h:21-77), a compile-time choice in the reference (optimize_edges.h:54).
 * All three run the same device program -- block-sparse L D L^t of the reduced system after eliminating the landmark
 * blocks first (which IS a sparse Cholesky of the whole system in a landmarks-first order) -- and differ only where the
 * reference's solver_engine<> specialisations differ observably (impl/lev-marq_solvers.h):
 *   SCHUR_*          : solve() reduces the caller's gradient IN PLACE (schur.h:248-265,283-296), landmarks whose
 *                      H_f + lambda I is not invertible are skipped (was_ith_feature_invertible);
 *   NO_SCHUR_SPARSE  : the gradient is left untouched by solve() (:182), a non-invertible landmark block makes the
 *                      factorisation fail (solve() returns false -> lambda is raised), every landmark counts as invertible (:196-200). */
enum { SRBA_SOLVER_SCHUR_DENSE_CHOLESKY = 0, SRBA_SOLVER_SCHUR_SPARSE_CHOLESKY = 1, SRBA_SOLVER_NO_SCHUR_SPARSE_CHOLESKY = 2 };

/* error codes */
enum {
	SRBA_OK = 0,
	SRBA_ERR_BAD_ARGUMENT = -1,
	SRBA_ERR_UNSUPPORTED_MODEL = -2,
	SRBA_ERR_RANK_DEFICIENT = -3, /* ASSERT_ABOVEEQ_(OBS_DIMS*nObs, nUnknowns)  optimize_edges.h:355 */
	SRBA_ERR_CUDA = -4,
	SRBA_ERR_NO_DEVICE = -5,
	SRBA_ERR_INTERNAL = -6,
	SRBA_ERR_NON_FINITE = -7 /* raised on the DEVICE (srba_stats_t::status): the initial error / Hessian diagonal, or the error of a
	                            trial, is not finite (NaN/Inf in the inputs or a diverged iterate).  The outputs still hold what
	                            the reference's arithmetic yields (NaN propagates; "rho<=0" is false for NaN, optimize_edges.h:476) */
};

/* Sensor / noise / LM parameters read by the hot path (TAllParameters, RbaEngine.h:463-472; TSRBAParameters :424-460). */
typedef struct {
	int32_t pose_kind, obs_kind, sensor_pose_kind, noise_kind;
	double sensor_pose[12];   /* parameters.sensor_pose.relative_pose as R|t (identity when sensor_pose_kind==NONE) */
	double noise_std;         /* observation_noise_identity: H,g scaled by 1/std (sic, srba_options_noise.h:55,70) */
	double noise_lambda[16];  /* observation_noise_constant_matrix::lambda, row-major o x o */
	double cam_left[4];       /* fx,fy,cx,cy */
	double cam_right[4];
	double right_cam_R2L[12]; /* (-rightCameraPose) as R|t, hoisted from models/sensors.h:193,259 */
	int32_t max_iters;        /* parameters.srba.max_iters */
	int32_t use_robust_kernel;
	double kernel_param;
	double max_error_per_obs_to_stop, max_rho, max_lambda, min_error_reduction_ratio_to_relinearize;
	int32_t solver_kind;      /* SRBA_SOLVER_* */
	int32_t reserved_;
} srba_model_t;

/* One flattened optimize_edges() problem.  Index spaces:
 *   edge  : k2k edges touched (unknown, or on any needed spanning-tree path)         [0,n_edges)
 *   ucol  : unknown k2k edges in optimisation order (optimize_edges.h:141-148)        [0,n_unk_edges)
 *   lm    : landmarks touched                                                         [0,n_lms)
 *   fcol  : unknown landmarks in optimisation order (optimize_edges.h:151-163)        [0,n_unk_lms)
 *   pair  : (from>to) keyframe pairs of the touched numeric-spanning-tree rows        [0,n_pairs)
 *   slot  : directed pose entries; slot 2t = num[from][to], slot 2t+1 = num[to][from] [0,2*n_pairs)
 *   obs   : distinct involved observations                                            [0,n_obs)
 */
typedef struct {
	srba_model_t model;
	int32_t n_edges, n_unk_edges, n_lms, n_unk_lms, n_pairs, n_path, n_obs;
	int32_t n_blkA, n_blkF;       /* dh_dAp / dh_df blocks of the unknown columns, reference walk order (jacobians.h:1083-1117) */
	int32_t n_HAp, n_HAp_terms;   /* H_Ap upper blocks before Schur + their (J1,J2) term lists (sparse_hessian_build_symbolic.h:38-105) */
	int32_t n_HApf, n_HApf_terms; /* H_Apf blocks (row-compressed) (:182-235) */
	int32_t n_S, n_S_terms;       /* blocks of the reduced system incl. Schur fill-in + triplets (schur.h:56-157) */
	int32_t n_total_obs_rows;     /* nObs of the reference INCLUDING the duplicates of optimize_edges.h:177-193 (= sum obs_mult) */
	int32_t reserved_;

	const double *edge_pose;        /* n_edges x S   current k2k_edge_t::inv_pose */
	const int32_t *unk_edge;        /* n_unk_edges   -> edge */
	const double *lm_pos;           /* n_lms x 3     TRelativeLandmarkPos::pos */
	const int32_t *unk_lm;          /* n_unk_lms     -> lm */

	const int32_t *pair_path_ptr;   /* n_pairs+1     CSR over path entries (sym.all_edges[from][to]) */
	const int32_t *path_edge;       /* n_path        -> edge */
	const uint8_t *path_dir;        /* n_path        1: compose inv_pose (edge.to==cur), 0: compose its inverse (spantree_update_numeric.h:47-65) */

	const double *obs_z;            /* n_obs x o     k2f_edge_t::obs.obs_arr */
	const int32_t *obs_slot;        /* n_obs         slot of num[obs_kf][base] or -1 when base==obs_kf (reprojection_residuals.h:41-52) */
	const int32_t *obs_lm;          /* n_obs         -> lm */
	const int32_t *obs_mult;        /* n_obs         multiplicity of this observation in involved_obs (>=1) */

	const int32_t *colA_ptr;        /* n_unk_edges+1 blocks of unknown column i are [colA_ptr[i],colA_ptr[i+1]) ascending obs */
	const int32_t *blkA_obs;        /* n_blkA        -> obs */
	const int32_t *blkA_slotA;      /* n_blkA        slot of num[obs_kf][d+1] ("A") or -1 == identity (add-observations.h:199-204) */
	const int32_t *blkA_slotD;      /* n_blkA        slot of num[d+1][base] ("D") */
	const uint8_t *blkA_normal_dir; /* n_blkA        TJacobianSymbolicInfo_dh_dAp::edge_normal_dir */

	const int32_t *colF_ptr;        /* n_unk_lms+1 */
	const int32_t *blkF_obs;        /* n_blkF        -> obs */
	const int32_t *blkF_slot;       /* n_blkF        slot of num[obs_kf][base] or -1 (first observation, add-observations.h:252-256) */

	const int32_t *row_ptr;         /* n_obs+1       blocks of one observation row in reference walk order ... */
	const int32_t *row_blk;         /* ... A block b as b, F block b as n_blkA+b (validity quirk) */

	const int32_t *HAp_row, *HAp_col;     /* n_HAp   block (row<=col) in ucol indices, sorted by (col,row) */
	const int32_t *HAp_term_ptr;          /* n_HAp+1 */
	const int32_t *HAp_term_a, *HAp_term_b; /* n_HAp_terms  blkA of J1 (row's column) and J2 (col's column), ascending obs */

	const int32_t *HApf_row, *HApf_col;   /* n_HApf  (ucol, fcol), sorted by (row,col) */
	const int32_t *HApf_term_ptr;         /* n_HApf+1 */
	const int32_t *HApf_term_a, *HApf_term_f; /* n_HApf_terms  blkA, blkF */
	const int32_t *rowApf_ptr;            /* n_unk_edges+1  H_Apf blocks of row i are contiguous */
	const int32_t *colApf_ptr;            /* n_unk_lms+1 */
	const int32_t *colApf_idx;            /* n_HApf  H_Apf block ids of landmark column k, ascending row */

	const int32_t *S_row, *S_col;         /* n_S     reduced-system blocks (row<=col), sorted by (col,row) */
	const int32_t *S_HAp;                 /* n_S     index of the pre-Schur H_Ap block or -1 for pure fill-in */
	const int32_t *S_term_ptr;            /* n_S+1 */
	const int32_t *S_term_rowblk, *S_term_colblk; /* n_S_terms  H_Apf block ids (row,k) and (col,k), ascending k */
} srba_problem_t;

typedef struct {
	int32_t num_iters;             /* final `iter` of optimize_edges.h:452 */
	int32_t num_trials;            /* number of solve(lambda) attempts == the BASELINE "LM iteration" unit */
	int32_t num_accepted, num_relinearized, num_chol_failures;
	int32_t num_jacobians;         /* count_jacobians (optimize_edges.h:276) */
	int32_t hessian_valid;
	int32_t status;                /* SRBA_OK or error raised on the device side */
	double total_sqr_error_init, total_sqr_error_final, obs_rmse, lambda_final;
} srba_stats_t;

/* Caller-allocated outputs; any pointer may be NULL to skip that output. */
typedef struct {
	srba_stats_t stats;
	double *edge_pose;      /* n_unk_edges x S   new inv_pose of the unknown edges (ucol order) */
	double *lm_pos;         /* n_unk_lms x 3 */
	double *slot_pose;      /* 2*n_pairs x S     final numeric spanning-tree entries, incl. the stale-entry semantics of
	                           optimize_edges.h:562-565,664-670 */
	uint8_t *obs_valid;     /* n_obs             all_observations_Jacob_validity after the last linearisation */
	uint8_t *lm_invertible; /* n_unk_lms         was_ith_feature_invertible (schur.h:313) */
	double *Hf_diag;        /* n_unk_lms x 9     final H_f diagonal blocks (optimize_edges.h:734-745) */
	double *dense_hessian;  /* (p n_unk_edges)^2 row-major, extra_results.hessian (lev-marq_solvers.h:582-586); upper triangle valid */
	int32_t trace_cap;      /* capacity (in trials) of trace / trace_delta */
	int32_t trace_len;      /* out: trials recorded */
	double *trace;          /* trace_cap x 6: lambda, solved_ok, rho, new_total_sq_err, accepted, relinearized */
	double *trace_delta;    /* trace_cap x (p n_unk_edges + 3 n_unk_lms): delta_eps of each trial */
} srba_result_t;

/* ---- the two entry points a reference-side shim binds (replace optimize_edges.h:44-793) ---- */
int srba_b200_optimize(const srba_problem_t *in, srba_result_t *out);
int srba_b200_optimize_batch(int32_t n, const srba_problem_t *const *in, srba_result_t *const *out);

/* ---- staged batch API: same work with explicit residency, used for measurement and pipelining ---- */
typedef struct srba_batch srba_batch_t;
/* Packs n problems into one pinned host image + sizes the device image/workspace.  Problems are copied. */
srba_batch_t *srba_b200_batch_create(int32_t n, const srba_problem_t *const *in, int32_t trace_cap);
int srba_b200_batch_upload(srba_batch_t *b, void *cuda_stream);    /* one H2D copy of the packed image */
/* Persistent problem handles: a batch keeps its SYMBOLIC structure -- CSR index arrays, Hessian / Schur term lists, the
 * elimination schedule of the reduced system, the shared-memory plan -- resident on the device between runs.  When only the NUMERIC
 * inputs of a problem change (the same unknowns and observations re-optimised from new edge poses / landmark positions / measurements,
 * e.g. after neighbouring areas moved), batch_update_values overwrites those arrays in the pinned image (counts as in the problem the
 * batch was created from; a NULL array keeps its values) and batch_upload_values copies just them: no symbolic work, no re-pack.
 * Replaces the per-call rebuilds of impl/sparse_hessian_build_symbolic.h:22-237 and impl/schur.h:25-159 for repeated solves. */
int srba_b200_batch_update_values(srba_batch_t *b, int32_t i, const double *edge_pose, const double *lm_pos, const double *obs_z);
int srba_b200_batch_upload_values(srba_batch_t *b, void *cuda_stream);
int srba_b200_batch_run(srba_batch_t *b, void *cuda_stream);       /* device-resident: whole LM loop for every problem */
int srba_b200_batch_download(srba_batch_t *b, void *cuda_stream);  /* one D2H copy of the packed results + stream sync */
int srba_b200_batch_unpack(srba_batch_t *b, int32_t i, srba_result_t *out); /* host: copy problem i's results out */
void srba_b200_batch_destroy(srba_batch_t *b);
int64_t srba_b200_batch_h2d_bytes(const srba_batch_t *b);
int64_t srba_b200_batch_d2h_bytes(const srba_batch_t *b);
int64_t srba_b200_batch_kernel_launches(const srba_batch_t *b); /* launches issued by the last batch_run */
/* SM clock cycles problem i spent in each stage of the LM program (measured by the kernel itself with clock64):
 * out[0..13] = init, schur+fill, cholesky, backsub, features, update, chain, residuals, rho, jacobians, hessian, gradient,
 * control, output; out[16..20] = sub-timers of the register-resident solver: tile gather + first pivot block, factorisation
 * loop, back-substitution loop, busy time of the critical-path threads inside the factorisation loop, call prologue */
int srba_b200_batch_stage_cycles(srba_batch_t *b, int32_t i, uint64_t out[24]);
/* Algorithmic bytes of one accepted+relinearised LM iteration of the whole batch (SURVEY.md §8d formulas), per stage:
 * out[0]=B_chain out[1]=B_assem out[2]=B_schur out[3]=B_chol out[4]=F_chol(flop) */
void srba_b200_batch_algorithmic_bytes(const srba_batch_t *b, double out[5]);

/* Measurement hook (bench.py's roofline line): when enabled, batch_run brackets every launch of the dominant kernel of the wave pipeline (the
 * cluster-team forward solver, wave_solve_fwd_kernel) with CUDA events; get: out[0] = summed milliseconds, out[1] = launches since enabling. */
void srba_b200_kernel_timing(int32_t enable);
void srba_b200_kernel_timing_get(double out[2]);

/* Host-only: how the packer would lay problem `in` out.  out[0..7] = image bytes, global workspace bytes, result bytes,
 * shared-memory bytes planned (solver scratch + on-chip arena), solver (0 block-sparse L D L^t, 1 dense blocked; +2 when
 * the problem is solved by an 8-CTA cluster, +4 by the grid-wide wave pipeline),
 * image arrays on chip, workspace arrays on chip, shared-memory budget of the model's kernel.  No device needed. */
int srba_b200_problem_plan(const srba_problem_t *in, int64_t out[8]);
/* The same for problem i of a batch, as the batch really planned it (execution modes depend on the batch: how many heavy
 * areas it holds).  out[4] = solver + 2 * (8-CTA cluster team) + 4 * (grid-wide wave pipeline). */
int srba_b200_batch_problem_plan(const srba_batch_t *b, int32_t i, int64_t out[8]);

/* Host-only: size of the block-sparse L D L^t schedule the packer plans for `in` (its own elimination order and on-chip budget):
 * out[0..3] = 3x3 tiles of the filled pattern incl. the right-hand side, tile updates per factorisation, barrier-separated phases,
 * 1 when the tiles stay in shared memory.  All 0 when the problem falls back to the dense blocked solver. */
int srba_b200_problem_ldl_counts(const srba_problem_t *in, int64_t out[4]);
/* Host-only: the level schedule of a problem that "auto" runs in the wave pipeline (all 0 otherwise): out[0..7] = tiles, destination entries,
 * tile updates (contributions), elimination levels, first level of the dense tail, 1 when its forward / backward levels run as a cluster
 * team, tiles of the rows below the tail, destination entries of the levels below the tail. */
int srba_b200_problem_wave_counts(const srba_problem_t *in, int64_t out[8]);

/* Host-only (tests): the symbolic factorisation the packer ships with a problem for the on-chip block-sparse L D L^t
 * solver -- 3x3 tiles numbered row-major (diagonal first, right-hand side as tile column NTR).  counts[6] = {NTR, tiles,
 * updates, ok, phases, tiles fit on chip}; row[4*(NTR+1)] = {first tile, end tile, first update, end update} per row;
 * phase[16*(phases+1)] = {first row, rows (1 or 2 independent rows eliminated together), first update, end of the first
 * row's updates, end update, pivots of the next phase, their tiles (2), the tiles through which this phase updates them
 * (2; bit 30: from the second row; -1 = none), first tile, first tile of the second row, end tile, 0, 0, 0};
 * tile_ij[tiles] = ti<<16|tj; src[tiles] = offset into the reduced blocks (bit 30: diagonal, bit 29: transposed), -1
 * fill-in, -2-row right-hand side; upd[2*updates] = {dst tile | tile(k,i)<<16, tile(k,j)}; orow[NTR] = tile row in
 * elimination order -> tile row in the caller's unknown order (identity unless the reference's order would overflow
 * `max_tiles_on_chip`, then greedy minimum degree; chain-like SE(2) areas are eliminated from both ends: 0, n-1, 1, ...).
 * Call with NULL arrays to get the counts first. */
int srba_b200_debug_ldl_schedule(const srba_problem_t *in, int32_t counts[6], int32_t *row, int32_t *tile_ij, int32_t *src, int32_t *upd,
                                 int32_t *orow, int32_t *phase, int32_t cap_tiles, int32_t cap_upd, int32_t max_tiles_on_chip);

/* Host-only (tests): the LEVEL schedule the packer ships with problems that run in the grid-wide wave pipeline (multi-CTA solver).
 * Tile rows sorted by elimination level.  counts[6] = {NTR, tiles, levels, destinations, contributions, ok};
 * row2[2*(NTR+1)] = {first tile, end tile} per row (+ sentinel); tile_ij / src / orow as in srba_b200_debug_ldl_schedule;
 * lv[2*(levels+1)] = {first row, first destination} per level (+ sentinel); gu = [destination tile (n_dst) | contribution pointer
 * (n_dst+1) | (tile(k,a), tile(k,b)) per contribution]: destination T(a,b) -= sum T(k,a)^t inv(D_k) T(k,b) over its contributions.
 * Call with NULL arrays to get the counts first. */
int srba_b200_debug_level_schedule(const srba_problem_t *in, int32_t counts[6], int32_t *row2, int32_t *tile_ij, int32_t *src, int32_t *orow, int32_t *lv,
                                   int32_t *gu);

/* ---- Global diagnostics: RbaEngine::eval_overall_squared_error (impl/eval_overall_error.h:15-137) and the
 * numeric half of create_complete_spanning_tree (impl/spantree_create_complete.h:91-125).
 * The caller (host engine) runs the reference's BFS (adjacency-order tie-breaking, :35-83) and hands over, per needed
 * keyframe pair, the CHAIN of k2k edges from the pair's root to its target as the BFS tree found it; the device composes
 *   pose[target wrt root] = I (+) e_1 (+) e_2 ...   with e_k = inv_pose when path_dir[k] else (-inv_pose)   (:111-122)
 * then evaluates every observation's error with the pose of its base keyframe as seen from its observer (the pair pose
 * or its inverse, eval_overall_error.h:72-76; identity when observer == base, :98-101), sensor pose and sensor model as in
 * the LM path, NO robust kernel (:129), and sums the squared norms in a fixed order (deterministic). */
typedef struct {
	srba_model_t model;             /* pose_kind, obs_kind, sensor pose, cameras (LM parameters ignored) */
	int32_t n_edges, n_pairs, n_path, n_obs, n_lms, reserved_;
	const double *edge_pose;        /* n_edges x S  k2k_edge_t::inv_pose */
	const int32_t *pair_path_ptr;   /* n_pairs+1   CSR into path_edge/path_dir, root -> target order */
	const int32_t *path_edge;       /* n_path      -> edge */
	const uint8_t *path_dir;        /* n_path      1: compose inv_pose (edge.from is the new node), 0: compose its inverse (edge.to is) */
	const double *obs_z;            /* n_obs x o   k2f_edge_t::obs.obs_arr */
	const int32_t *obs_pair;        /* n_obs       -> pair, or -1 when observer == base */
	const uint8_t *obs_inverse;     /* n_obs       1: the observer is the pair's target (use the inverse of the pair pose) */
	const int32_t *obs_lm;          /* n_obs       -> lm (ignored for relative-pose observations) */
	const double *lm_pos;           /* n_lms x 3   TRelativeLandmarkPos::pos */
} srba_eval_problem_t;
/* out_sqerr: the total; out_pair_pose (optional, n_pairs x S): the composed poses (what create_complete_spanning_tree
 * returns for the targets).  n_obs may be 0 (poses only).  Two kernels + a fixed-order final sum, all on the device. */
int srba_b200_eval_overall_squared_error(const srba_eval_problem_t *in, double *out_sqerr, double *out_pair_pose);

const char *srba_b200_last_error(void);
/* Frees the calling thread's reusable workspace of the synchronous entry points (see "State" above). */
void srba_b200_release_workspace(void);
/* Execution options (tests / experiments; process-wide, initialised from the SRBA_B200_* environment variables):
 *   "cluster"  -1 auto | 0 never | 1 always  an 8-CTA cluster team per heavy area
 *   "blocked"   0 auto | 1                   dense blocked Cholesky instead of the block-sparse L D L^t
 *   "wide"      0 auto | 1                   3-int solver update descriptors
 *   "chunks"    1..8                         pipelined chunks of srba_b200_optimize_batch (default 4)
 *   "wave"     -1 auto | 0 never | 1 always  the grid-wide lock-step pipeline: one kernel launch per LM stage over ALL areas
 *                                            (auto: areas whose arrays overflow one CTA's shared-memory arena)
 * set returns SRBA_OK or SRBA_ERR_BAD_ARGUMENT (unknown name); get returns INT32_MIN for an unknown name. */
int srba_b200_set_option(const char *name, int32_t value);
int32_t srba_b200_get_option(const char *name);
int srba_b200_device_count(void);
const char *srba_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
