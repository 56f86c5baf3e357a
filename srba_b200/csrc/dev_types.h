// Packed device image of one flattened optimize_edges() problem (host/device shared POD).
// The host packer (capi.cpp) lays every array of srba_problem_t (include/srba_b200.h) out in ONE contiguous,
// 16-byte-aligned blob per problem, all problems of a batch back to back, so a batch is one H2D copy.
// Offsets are in BYTES: img_off[] relative to the problem's blob in the input image, ws_off[] relative to the problem's
// workspace, res_* relative to the problem's result blob.
//
// On-chip residency: the LM kernel (lm_kernel.cu) owns a shared-memory ARENA per CTA.  The packer decides, per problem
// and in priority order, which mutable workspace arrays (and which hot read-only index arrays of the image) live in that
// arena for the whole LM loop (sm offset != SRBA_NOT_IN_SMEM) and which stay in the L2-resident global workspace.
#pragma once
#include <stdint.h>

#include "../../include/srba_b200.h"

#define SRBA_NOT_IN_SMEM 0xFFFFFFFFu
// wave solver: tile rows of the trailing dense triangle that one CTA keeps in shared memory ((R+1)(R+2)/2 tiles of 72 B + row buffers)
#define SRBA_WAVE_TAIL_ROWS 70
// wave solver: an area whose levels below the tail hold at least this many destination tiles on average runs them as an 8-CTA cluster team
#ifndef SRBA_WAVE_TEAM_DST_PER_LEVEL
#define SRBA_WAVE_TEAM_DST_PER_LEVEL 600
#endif
// "auto" execution mode: SE(3) problems with at least this many unknown edges (whole-map scale) take the wave pipeline
#define SRBA_WAVE_SE3_MIN_EDGES 256

// (type, name) of every array of the input image, in srba_problem_t order + derived ones: block -> column, and the
// symbolic factorisation of the reduced system at 3x3-tile granularity (tiles numbered row-major, diagonal first, the
// right-hand side as tile column NTR):  ldl_row[4*(NTR+1)] per row {first tile, end tile, first update, end update}
// (+ sentinel);  ldl_phase[16*(phases+1)]: the barrier-separated elimination phases of 1 or 2 independent rows (capi.cpp);  ldl_tile_ij (ti<<16|tj);  ldl_src: where the tile comes from (offset into the reduced
// blocks, -1 fill-in, -2-row right-hand side);  ldl_upd: per eliminated row the updates {dst tile | tile(k,i)<<16,
// tile(k,j)} (the one that hits the next pivot block is implicit);  ldl_orow: tile row in elimination order -> tile row in the
// reference's unknown order (the packer re-orders areas with loop closures by minimum degree).
#define SRBA_IMG_ARRAYS(X)                                                                                              \
	X(double, edge_pose) X(int32_t, unk_edge) X(double, lm_pos) X(int32_t, unk_lm)                                       \
	X(int32_t, pair_path_ptr) X(int32_t, path_edge) X(uint8_t, path_dir)                                                 \
	X(double, obs_z) X(int32_t, obs_slot) X(int32_t, obs_lm) X(int32_t, obs_mult)                                        \
	X(int32_t, colA_ptr) X(int32_t, blkA_obs) X(int32_t, blkA_slotA) X(int32_t, blkA_slotD) X(uint8_t, blkA_normal_dir)  \
	X(int32_t, colF_ptr) X(int32_t, blkF_obs) X(int32_t, blkF_slot)                                                      \
	X(int32_t, row_ptr) X(int32_t, row_blk)                                                                              \
	X(int32_t, HAp_row) X(int32_t, HAp_col) X(int32_t, HAp_term_ptr) X(int32_t, HAp_term_a) X(int32_t, HAp_term_b)       \
	X(int32_t, HApf_row) X(int32_t, HApf_col) X(int32_t, HApf_term_ptr) X(int32_t, HApf_term_a) X(int32_t, HApf_term_f)  \
	X(int32_t, rowApf_ptr) X(int32_t, colApf_ptr) X(int32_t, colApf_idx)                                                 \
	X(int32_t, S_row) X(int32_t, S_col) X(int32_t, S_HAp) X(int32_t, S_term_ptr) X(int32_t, S_term_rowblk) X(int32_t, S_term_colblk) \
	X(int32_t, blkA_col) X(int32_t, blkF_col) X(int32_t, ldl_row) X(int32_t, ldl_tile_ij) X(int32_t, ldl_src) X(int32_t, ldl_upd) X(int32_t, ldl_orow) X(int32_t, ldl_phase)

// (type, name) of every workspace array.
#define SRBA_WS_ARRAYS(X)                                                                                               \
	X(double, edge_cur) X(double, edge_bak) X(double, lm_cur) X(double, lm_bak) X(double, slot) X(double, slot_bak)      \
	X(double, JA) X(double, JF) X(double, resid0) X(double, resid1)                                                      \
	X(double, HAp) X(double, HApf) X(double, Hf) X(double, Hf_inv) X(double, W) X(double, Sblk)                          \
	X(double, grad) X(double, delta) X(double, dense) X(double, grad0)                                                                  \
	X(int32_t, listed_idx) X(int32_t, refresh_idx) X(double, ldl_tiles)                                                                                             \
	X(uint8_t, valid) X(uint8_t, blk_ok) X(uint8_t, lm_inv) X(uint8_t, slot_listed)

enum {
#define X(T, name) SRBA_I_##name,
	SRBA_IMG_ARRAYS(X)
#undef X
	SRBA_I_COUNT
};
enum {
#define X(T, name) SRBA_W_##name,
	SRBA_WS_ARRAYS(X)
#undef X
	SRBA_W_COUNT
};

// How the dense reduced system is solved for this problem (decided by the packer from the kernel's tile capacity).
enum { SRBA_CHOL_REG = 0 /* on-chip block-sparse LDL^t */, SRBA_CHOL_BLOCKED = 1 /* dense blocked Cholesky in the L2-resident workspace */ };

struct DevProblem {
	srba_model_t model;
	int32_t n_edges, n_unk_edges, n_lms, n_unk_lms, n_pairs, n_path, n_obs;
	int32_t n_blkA, n_blkF, n_HAp, n_HAp_terms, n_HApf, n_HApf_terms, n_S, n_S_terms;
	int32_t n_total_obs_rows;
	int32_t n_scalars;   // p*n_unk_edges + 3*n_unk_lms
	int32_t n_dense;     // p*n_unk_edges
	int32_t ld_dense;    // leading dimension of the dense reduced system (n_dense + 1 augmented column, padded even)
	int32_t trace_cap;
	int32_t cost_rank;   // position in the batch's cost-descending order (scheduling only)
	int32_t want_dense;  // copy the dense reduced Hessian of the last solve into the result blob
	int32_t chol_mode;   // SRBA_CHOL_*
	int32_t cluster;     // 1: solved by an 8-CTA cluster (lm_kernel_cluster.cu), every array in the global workspace
	int32_t wave;        // 1: solved by the grid-wide lock-step pipeline (lm_wave.cuh), every array in the global workspace
	int32_t ldl_n_dst;   // wave: destinations / contributions of the level schedule (capi.cpp: symbolic_levels): ldl_phase holds
	int32_t ldl_n_contrib;  //    {first row, first destination} per level, ldl_upd [dst | contribution pointer | (tile(k,a), tile(k,b))]
	int32_t ldl_tail_level; // wave: first level of the TAIL -- the last <= SRBA_WAVE_TAIL_ROWS tile rows, eliminated densely in one CTA's shared memory
	int32_t panel_cap;   // SRBA_CHOL_BLOCKED: columns of the staged 8-row panel that fit the scratch
	int32_t ldl_n_tiles; // SRBA_CHOL_REG: 3x3 tiles of the filled pattern incl. the right-hand-side column
	int32_t ldl_n_phases; // barrier-separated elimination phases (1 or 2 independent block rows each)
	int32_t ldl_wide;     // 1: update descriptors are 3 ints {dst, tile(k,i), tile(k,j)} (more than 32767 tiles)
	int32_t ldl_tiles_in_ws; // 1: the tiles do not fit on chip and live in the workspace array ldl_tiles (L2-resident)
	int32_t s_alias_hap; // 1: no Schur terms and S block b == H_Ap block b -> the reduced blocks ARE the H_Ap blocks
	uint32_t sm_scratch; // byte offset of the Cholesky scratch inside the CTA's dynamic shared memory
	uint32_t sm_bytes;   // dynamic shared memory this problem uses (scratch + arena)

	uint64_t img_off[SRBA_I_COUNT];
	uint32_t img_len[SRBA_I_COUNT];  // bytes
	uint32_t img_sm[SRBA_I_COUNT];   // offset of the on-chip copy or SRBA_NOT_IN_SMEM
	uint64_t img_bytes;

	uint64_t ws_off[SRBA_W_COUNT];
	uint32_t ws_sm[SRBA_W_COUNT];    // offset in the arena or SRBA_NOT_IN_SMEM (then ws_off is used)
	uint64_t ws_bytes;

	// ---- result blob
	uint64_t res_stats, res_edge_pose, res_lm_pos, res_slot_pose, res_obs_valid, res_lm_inv, res_Hf_diag, res_dense, res_trace, res_trace_delta, res_stage_cycles;
	uint64_t res_bytes;
};

// Per-model kernel configuration shared by the packer and the launcher.
struct SrbaKernelCfg {
	int threads;      // CTA size
	int tpt;          // register tiles per thread of the dense solver
	int smem_bytes;   // dynamic shared memory per CTA (Cholesky scratch + arena)
};
#ifdef __cplusplus
extern "C" {
#endif
// model: 0 = SE2+RelativePoses2D, 1 = SE3+Cartesian3D, 2 = SE3+Mono, 3 = SE3+Stereo
SrbaKernelCfg srba_kernel_cfg(int model);
// doubles of shared-memory scratch the register-resident solver needs; <0 when the filled pattern does not fit the tiles
int srba_chol_reg_scratch_doubles(int model, int n_dense, int n_tiles_on_chip);
#ifdef __cplusplus
}
#endif
