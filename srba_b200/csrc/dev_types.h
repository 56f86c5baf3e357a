// Packed device image of one flattened optimize_edges() problem (host/device shared POD).
// The host packer (capi.cpp) lays every array of srba_problem_t (include/srba_b200.h) out in ONE contiguous,
// 16-byte-aligned blob per problem, all problems of a batch back to back, so a batch is one H2D copy.
// Offsets are in BYTES: img_* relative to the problem's blob in the input image, ws_* relative to the problem's
// workspace, res_* relative to the problem's result blob.
#pragma once
#include <stdint.h>

#include "../../include/srba_b200.h"

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

	// ---- input image
	uint64_t img_edge_pose, img_unk_edge, img_lm_pos, img_unk_lm;
	uint64_t img_pair_path_ptr, img_path_edge, img_path_dir;
	uint64_t img_obs_z, img_obs_slot, img_obs_lm, img_obs_mult;
	uint64_t img_colA_ptr, img_blkA_obs, img_blkA_col, img_blkA_slotA, img_blkA_slotD, img_blkA_normal_dir;
	uint64_t img_colF_ptr, img_blkF_obs, img_blkF_col, img_blkF_slot;
	uint64_t img_row_ptr, img_row_blk;
	uint64_t img_HAp_row, img_HAp_col, img_HAp_term_ptr, img_HAp_term_a, img_HAp_term_b;
	uint64_t img_HApf_row, img_HApf_col, img_HApf_term_ptr, img_HApf_term_a, img_HApf_term_f;
	uint64_t img_rowApf_ptr, img_colApf_ptr, img_colApf_idx;
	uint64_t img_S_row, img_S_col, img_S_HAp, img_S_term_ptr, img_S_term_rowblk, img_S_term_colblk;
	uint64_t img_bytes;

	// ---- workspace (doubles unless noted)
	uint64_t ws_edge_cur, ws_edge_bak, ws_lm_cur, ws_lm_bak, ws_slot, ws_slot_bak;
	uint64_t ws_JA, ws_JF, ws_resid0, ws_resid1;
	uint64_t ws_HAp, ws_HApf, ws_Hf, ws_Hf_inv, ws_W;
	uint64_t ws_grad, ws_delta, ws_dense;
	uint64_t ws_valid, ws_blk_ok, ws_lm_inv, ws_slot_listed, ws_pair_refresh;  // bytes
	uint64_t ws_bytes;

	// ---- result blob
	uint64_t res_stats, res_edge_pose, res_lm_pos, res_slot_pose, res_obs_valid, res_lm_inv, res_Hf_diag, res_dense, res_trace, res_trace_delta, res_stage_cycles;
	uint64_t res_bytes;
};
