// srba_b200_shim.hpp -- the SEAM-LEVEL shim: flattens the sub-problem selected by RbaEngine<>::optimize_edges out of the reference's own
// pointer-graph state (TRBA_Problem_state, srba_types.h:548-785: lin_system.dh_dAp / dh_df column maps with their symbolic entries,
// spanning_tree.sym.all_edges, spanning_tree.num, all_observations, all_lms, k2k_edges) into the index-addressed arrays of srba_problem_t
// (include/srba_b200.h), and scatters a srba_result_t back.  It is the pointer-graph counterpart of srba_b200::RbaEngine::flatten()
// (srba_b200/csrc/host_engine.hpp) and produces the SAME arrays in the same orders (tests/test_shim.py compares them element by element).
//
// The algorithm is written once, against a small ACCESS policy `A` of static functions that spell out how to reach the state:
//   shim/srba/impl/optimize_edges.h  supplies the policy for the reference's types (MRPT poses, MatrixBlockSparseCols::getCol, ...);
//   tests/shim/seam_oracle.cpp       supplies it for the test oracle's restatement of TRBA_Problem_state (same containers), which is
//                                    how the flattening and the scatter are exercised here, where MRPT itself cannot be compiled.
// Policy (ST = state type):
//   A::colAp(st, k2k_edge_id)  -> const std::map<size_t, ENTRY>&   ENTRY::sym.{kf_d, edge_normal_dir}      (dh_dAp column, srba_types.h:204-240)
//   A::colF(st, feat_id)       -> const std::map<size_t, ENTRY>&                                           (dh_df column, :246-266)
//   A::n_k2k(st), A::edge_from/edge_to(st, e), A::edge_pose_get/set(st, e, double *flat)                    (k2k_edge_t, :82-91)
//   A::lm_known(st, id), A::lm_base(st, id), A::lm_pos(st, id) -> double*                                   (TLandmarkEntry / TRelativeLandmarkPos, :106-135)
//   A::n_obs(st), A::obs_kf / obs_feat / obs_first_of_unknown(st, g), A::obs_data(st, g, double *out)       (k2f_edge_t, :515-525)
//   A::all_edges_from(st, kf)  -> const std::map<TKeyFrameID, std::deque<EDGE*>>* (NULL when kf has no row) (spanning_tree.sym.all_edges, :600-607)
//   A::edge_id(EDGE*), A::set_num(st, src, dst, const double *flat), A::set_validity(st, g, char)           (spanning_tree.num :609-618, :722)
//   A::POSE_NVALS (3: x y phi | 12: R row-major + t), A::OBS_DIMS
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "../include/srba_b200.h"

namespace srba_b200_shim {

struct Flat {
	srba_problem_t p;
	// ids needed to scatter the result back
	std::vector<uint64_t> k2k_ids, lm_ids, obs_global, edge_ids_touched, lm_ids_touched;
	std::vector<std::pair<uint64_t, uint64_t> > pairs;  // (from > to) of every numeric spanning-tree pair
	uint64_t num_span_tree_numeric_updates, num_kf_optimized, num_lm_optimized;
	// storage behind srba_problem_t
	std::vector<double> edge_pose, lm_pos, obs_z;
	std::vector<int32_t> unk_edge, unk_lm, pair_path_ptr, path_edge, obs_slot, obs_lm, obs_mult, colA_ptr, blkA_obs, blkA_slotA, blkA_slotD, colF_ptr, blkF_obs, blkF_slot, row_ptr,
	    row_blk, HAp_row, HAp_col, HAp_term_ptr, HAp_term_a, HAp_term_b, HApf_row, HApf_col, HApf_term_ptr, HApf_term_a, HApf_term_f, rowApf_ptr, colApf_ptr, colApf_idx, S_row, S_col,
	    S_HAp, S_term_ptr, S_term_rowblk, S_term_colblk;
	std::vector<uint8_t> path_dir, blkA_normal_dir;
	// caller-owned result buffers sized for this problem
	std::vector<double> r_edge_pose, r_lm_pos, r_slot_pose, r_Hf_diag;
	std::vector<uint8_t> r_obs_valid, r_lm_inv;
	std::vector<double> r_trace;   // 6 per LM trial: lambda, solved_ok, rho, new total squared error, accepted, relinearised

	void bind() {
#define SRBA_SHIM_BIND(name) p.name = name.empty() ? NULL : &name[0];
		SRBA_SHIM_BIND(edge_pose) SRBA_SHIM_BIND(unk_edge) SRBA_SHIM_BIND(lm_pos) SRBA_SHIM_BIND(unk_lm) SRBA_SHIM_BIND(pair_path_ptr) SRBA_SHIM_BIND(path_edge)
		SRBA_SHIM_BIND(path_dir) SRBA_SHIM_BIND(obs_z) SRBA_SHIM_BIND(obs_slot) SRBA_SHIM_BIND(obs_lm) SRBA_SHIM_BIND(obs_mult) SRBA_SHIM_BIND(colA_ptr) SRBA_SHIM_BIND(blkA_obs)
		SRBA_SHIM_BIND(blkA_slotA) SRBA_SHIM_BIND(blkA_slotD) SRBA_SHIM_BIND(blkA_normal_dir) SRBA_SHIM_BIND(colF_ptr) SRBA_SHIM_BIND(blkF_obs) SRBA_SHIM_BIND(blkF_slot)
		SRBA_SHIM_BIND(row_ptr) SRBA_SHIM_BIND(row_blk) SRBA_SHIM_BIND(HAp_row) SRBA_SHIM_BIND(HAp_col) SRBA_SHIM_BIND(HAp_term_ptr) SRBA_SHIM_BIND(HAp_term_a)
		SRBA_SHIM_BIND(HAp_term_b) SRBA_SHIM_BIND(HApf_row) SRBA_SHIM_BIND(HApf_col) SRBA_SHIM_BIND(HApf_term_ptr) SRBA_SHIM_BIND(HApf_term_a) SRBA_SHIM_BIND(HApf_term_f)
		SRBA_SHIM_BIND(rowApf_ptr) SRBA_SHIM_BIND(colApf_ptr) SRBA_SHIM_BIND(colApf_idx) SRBA_SHIM_BIND(S_row) SRBA_SHIM_BIND(S_col) SRBA_SHIM_BIND(S_HAp) SRBA_SHIM_BIND(S_term_ptr)
		SRBA_SHIM_BIND(S_term_rowblk) SRBA_SHIM_BIND(S_term_colblk)
#undef SRBA_SHIM_BIND
	}
	/** srba_result_t over buffers owned by this object; trace_cap > 0: the per-trial trace (lambda, rho, errors, decisions) of up to that many
	  * LM trials comes back too -- what feedback_user_iteration is replayed from (impl/optimize_edges.h:417-418,595-596) */
	srba_result_t make_result_buffers(int pose_nvals, int trace_cap = 0) {
		r_edge_pose.assign((size_t)p.n_unk_edges * pose_nvals, 0.0);
		r_lm_pos.assign((size_t)p.n_unk_lms * 3, 0.0);
		r_slot_pose.assign((size_t)2 * p.n_pairs * pose_nvals, 0.0);
		r_Hf_diag.assign((size_t)p.n_unk_lms * 9, 0.0);
		r_obs_valid.assign(p.n_obs, 0);
		r_lm_inv.assign(p.n_unk_lms, 0);
		srba_result_t r;
		std::memset(&r, 0, sizeof(r));
		r.edge_pose = r_edge_pose.empty() ? NULL : &r_edge_pose[0];
		r.lm_pos = r_lm_pos.empty() ? NULL : &r_lm_pos[0];
		r.slot_pose = r_slot_pose.empty() ? NULL : &r_slot_pose[0];
		r.Hf_diag = r_Hf_diag.empty() ? NULL : &r_Hf_diag[0];
		r.obs_valid = r_obs_valid.empty() ? NULL : &r_obs_valid[0];
		r.lm_invertible = r_lm_inv.empty() ? NULL : &r_lm_inv[0];
		if (trace_cap > 0) {
			r_trace.assign((size_t)trace_cap * 6, 0.0);
			r.trace = &r_trace[0];
			r.trace_cap = trace_cap;
		}
		return r;
	}
};

/** Steps A-D, I, O of optimize_edges (impl/optimize_edges.h:71-246,303-312; impl/sparse_hessian_build_symbolic.h; impl/schur.h:25-159) as index
  * arrays.  Returns NULL when no observation depends on any of the given unknowns (optimize_edges.h:132-138). */
template <class A, class ST>
std::unique_ptr<Flat> flatten(const ST &st, const srba_model_t &model, const std::vector<size_t> &k2k_in, const std::vector<size_t> &lms_in,
                              const std::vector<size_t> &obs_indices_in = std::vector<size_t>()) {
	const int S = A::POSE_NVALS, O = A::OBS_DIMS, P = (S == 3 ? 3 : 6);
	std::unique_ptr<Flat> fp(new Flat());
	Flat &F = *fp;
	std::memset(&F.p, 0, sizeof(F.p));
	// A. filter unknowns with empty columns (:71-119)
	std::set<uint64_t> touched_KFs, touched_LMs;
	for (size_t i = 0; i < k2k_in.size(); i++) {
		const size_t e = k2k_in[i];
		if (e >= A::n_k2k(st)) throw std::runtime_error("optimize_edges: k2k edge index out of range");
		if (A::colAp(st, e).empty()) continue;
		F.k2k_ids.push_back(e);
		touched_KFs.insert(A::edge_from(st, e)); touched_KFs.insert(A::edge_to(st, e));
	}
	for (size_t i = 0; i < lms_in.size(); i++) {
		const size_t id = lms_in[i];
		if (A::lm_known(st, id)) throw std::runtime_error("optimize_edges: a landmark with known relative position cannot be an unknown");
		if (A::colF(st, id).empty()) continue;
		F.lm_ids.push_back(id);
		touched_LMs.insert(id);
	}
	const int n_p = (int)F.k2k_ids.size(), n_f = (int)F.lm_ids.size();
	F.num_kf_optimized = touched_KFs.size();
	F.num_lm_optimized = touched_LMs.size();
	if (P * n_p + 3 * n_f == 0) return std::unique_ptr<Flat>();
	// C. involved observations (:171-234): first-touch order; the dh_dAp walk does NOT de-duplicate (obs_mult counts the repeats)
	std::unordered_map<uint64_t, int> local_obs;
	auto obs_local = [&](uint64_t g) -> int {
		std::unordered_map<uint64_t, int>::iterator it = local_obs.find(g);
		if (it != local_obs.end()) return it->second;
		const int l = (int)F.obs_global.size();
		local_obs[g] = l;
		F.obs_global.push_back(g);
		F.obs_mult.push_back(0);
		return l;
	};
	const bool user_list = !obs_indices_in.empty();
	for (size_t i = 0; i < obs_indices_in.size(); i++) {
		if (obs_indices_in[i] >= A::n_obs(st)) throw std::runtime_error("optimize_edges: observation index out of range");
		F.obs_mult[obs_local(obs_indices_in[i])]++;
	}
	F.colA_ptr.push_back(0);
	for (int i = 0; i < n_p; i++) {
		const typename A::colAp_t &col = A::colAp(st, F.k2k_ids[i]);
		for (typename A::colAp_t::const_iterator it = col.begin(); it != col.end(); ++it) {
			if (user_list && !local_obs.count(it->first)) throw std::runtime_error("ASSERT_(it_obs!=obs_global_idx2residual_idx.end()) (compute_minus_gradient.h:55)");
			const int l = obs_local(it->first);
			if (!user_list) F.obs_mult[l]++;
			F.blkA_obs.push_back(l);
		}
		F.colA_ptr.push_back((int32_t)F.blkA_obs.size());
	}
	F.colF_ptr.push_back(0);
	for (int i = 0; i < n_f; i++) {
		const typename A::colF_t &col = A::colF(st, F.lm_ids[i]);
		for (typename A::colF_t::const_iterator it = col.begin(); it != col.end(); ++it) {
			const bool fresh = !local_obs.count(it->first);
			if (user_list && fresh) throw std::runtime_error("ASSERT_(it_obs!=obs_global_idx2residual_idx.end()) (compute_minus_gradient.h:79)");
			const int l = obs_local(it->first);
			if (fresh) F.obs_mult[l] = 1;
			F.blkF_obs.push_back(l);
		}
		F.colF_ptr.push_back((int32_t)F.blkF_obs.size());
	}
	const int n_obs = (int)F.obs_global.size(), nA = (int)F.blkA_obs.size(), nF = (int)F.blkF_obs.size();
	int n_rows_dup = 0;
	for (size_t i = 0; i < F.obs_mult.size(); i++) n_rows_dup += F.obs_mult[i];
	// D. needed spanning-tree roots (jacobians.h:1020-1076) + the symbolic info of every dh_dAp block
	std::set<uint64_t> roots;
	std::vector<uint64_t> blkA_kf_d(nA);
	F.blkA_normal_dir.resize(nA);
	for (int i = 0; i < n_p; i++) {
		const typename A::colAp_t &col = A::colAp(st, F.k2k_ids[i]);
		int b = F.colA_ptr[i];
		for (typename A::colAp_t::const_iterator it = col.begin(); it != col.end(); ++it, ++b) {
			const uint64_t g = it->first;
			blkA_kf_d[b] = it->second.sym.kf_d;
			F.blkA_normal_dir[b] = it->second.sym.edge_normal_dir ? 1 : 0;
			roots.insert(it->second.sym.kf_d); roots.insert(A::obs_kf(st, g)); roots.insert(A::lm_base(st, A::obs_feat(st, g)));
		}
	}
	for (int b = 0; b < nF; b++) {
		const uint64_t g = F.obs_global[F.blkF_obs[b]];
		roots.insert(A::lm_base(st, A::obs_feat(st, g))); roots.insert(A::obs_kf(st, g));
	}
	std::map<std::pair<uint64_t, uint64_t>, int> pair_idx;
	std::unordered_map<uint64_t, int> local_edge;
	auto edge_local = [&](uint64_t e) -> int {
		std::unordered_map<uint64_t, int>::iterator it = local_edge.find(e);
		if (it != local_edge.end()) return it->second;
		const int l = (int)F.edge_ids_touched.size();
		local_edge[e] = l;
		F.edge_ids_touched.push_back(e);
		return l;
	};
	for (int i = 0; i < n_p; i++) { edge_local(F.k2k_ids[i]); F.unk_edge.push_back(i); }
	F.pair_path_ptr.push_back(0);
	F.num_span_tree_numeric_updates = 0;
	for (std::set<uint64_t>::const_iterator itr = roots.begin(); itr != roots.end(); ++itr) {
		const uint64_t from = *itr;
		const typename A::edges_row_t *row = A::all_edges_from(st, from);   // spanning_tree.sym.all_edges[from]: entries [from][to], from > to
		if (!row) continue;
		F.num_span_tree_numeric_updates += row->size();
		for (typename A::edges_row_t::const_iterator kv = row->begin(); kv != row->end(); ++kv) {
			pair_idx[std::make_pair(from, (uint64_t)kv->first)] = (int)F.pairs.size();
			F.pairs.push_back(std::make_pair(from, (uint64_t)kv->first));
			uint64_t cur = from;
			for (size_t q = 0; q < kv->second.size(); q++) {   // update_numeric_only_all_from_node (spantree_update_numeric.h:44-66)
				const uint64_t e = A::edge_id(kv->second[q]);
				F.path_edge.push_back(edge_local(e));
				const bool fwd = (A::edge_to(st, e) == cur);
				F.path_dir.push_back(fwd ? 1 : 0);
				cur = fwd ? A::edge_from(st, e) : A::edge_to(st, e);
			}
			F.pair_path_ptr.push_back((int32_t)F.path_edge.size());
		}
	}
	auto slot_of = [&](uint64_t src, uint64_t dst) -> int {  // slot of num[src][dst]: 2 per pair, [from][to] then [to][from]
		std::map<std::pair<uint64_t, uint64_t>, int>::const_iterator it = pair_idx.find(std::make_pair(std::max(src, dst), std::min(src, dst)));
		if (it == pair_idx.end()) throw std::runtime_error("srba_b200 shim: a numeric spanning-tree entry needed by a Jacobian is not among the refreshed rows");
		return 2 * it->second + (src > dst ? 0 : 1);
	};
	std::unordered_map<uint64_t, int> local_lm;
	auto lm_local = [&](uint64_t id) -> int {
		std::unordered_map<uint64_t, int>::iterator it = local_lm.find(id);
		if (it != local_lm.end()) return it->second;
		const int l = (int)F.lm_ids_touched.size();
		local_lm[id] = l;
		F.lm_ids_touched.push_back(id);
		return l;
	};
	for (int i = 0; i < n_f; i++) F.unk_lm.push_back(lm_local(F.lm_ids[i]));
	F.obs_z.resize((size_t)n_obs * O);
	F.obs_slot.resize(n_obs); F.obs_lm.resize(n_obs);
	for (int l = 0; l < n_obs; l++) {
		const uint64_t g = F.obs_global[l];
		A::obs_data(st, g, &F.obs_z[(size_t)l * O]);
		const uint64_t feat = A::obs_feat(st, g), base = A::lm_base(st, feat), kf = A::obs_kf(st, g);
		F.obs_slot[l] = (base == kf) ? -1 : slot_of(kf, base);
		F.obs_lm[l] = lm_local(feat);
	}
	F.blkA_slotA.resize(nA); F.blkA_slotD.resize(nA);
	for (int b = 0; b < nA; b++) {
		const uint64_t g = F.obs_global[F.blkA_obs[b]], kf = A::obs_kf(st, g), base = A::lm_base(st, A::obs_feat(st, g));
		F.blkA_slotD[b] = slot_of(blkA_kf_d[b], base);                                 // sym.rel_pose_base_from_d1 (jacobians.h:1040-1052)
		F.blkA_slotA[b] = (blkA_kf_d[b] == kf) ? -1 : slot_of(kf, blkA_kf_d[b]);       // sym.rel_pose_d1_from_obs
	}
	F.blkF_slot.resize(nF);
	for (int b = 0; b < nF; b++) {
		const uint64_t g = F.obs_global[F.blkF_obs[b]];
		F.blkF_slot[b] = A::obs_first_of_unknown(st, g) ? -1 : slot_of(A::obs_kf(st, g), A::lm_base(st, A::obs_feat(st, g)));   // sym.rel_pose_base_from_obs
	}
	// rows in walk order (dh_dAp columns in unknown order, then the dh_df block): the validity protocol of jacobians.h:1083-1117
	std::vector<std::vector<int> > row_lists(n_obs);
	for (int b = 0; b < nA; b++) row_lists[F.blkA_obs[b]].push_back(b);
	for (int b = 0; b < nF; b++) row_lists[F.blkF_obs[b]].push_back(nA + b);
	F.row_ptr.push_back(0);
	for (int l = 0; l < n_obs; l++) {
		for (size_t q = 0; q < row_lists[l].size(); q++) F.row_blk.push_back(row_lists[l][q]);
		F.row_ptr.push_back((int32_t)F.row_blk.size());
	}
	// I. Hessian symbolic (sparse_hessian_build_symbolic.h:22-237): terms of every block in ascending GLOBAL observation index
	std::vector<int> obs_sorted(n_obs);
	for (int l = 0; l < n_obs; l++) obs_sorted[l] = l;
	std::sort(obs_sorted.begin(), obs_sorted.end(), [&](int a, int b) { return F.obs_global[a] < F.obs_global[b]; });
	std::vector<int> blkA_ucol(nA), blkF_fcol(nF);
	for (int i = 0; i < n_p; i++) for (int b = F.colA_ptr[i]; b < F.colA_ptr[i + 1]; b++) blkA_ucol[b] = i;
	for (int i = 0; i < n_f; i++) for (int b = F.colF_ptr[i]; b < F.colF_ptr[i + 1]; b++) blkF_fcol[b] = i;
	typedef std::map<std::pair<int, int>, std::vector<std::pair<int, int> > > term_map_t;
	term_map_t HAp_terms, HApf_terms;   // key (col,row) -> (blk_row, blk_col);  key (row,fcol) -> (blkA, blkF)
	for (size_t qq = 0; qq < obs_sorted.size(); qq++) {
		const std::vector<int> &rl = row_lists[obs_sorted[qq]];
		for (size_t x = 0; x < rl.size(); x++) {
			if (rl[x] >= nA) continue;
			for (size_t y = 0; y < rl.size(); y++) {
				if (rl[y] >= nA) { HApf_terms[std::make_pair(blkA_ucol[rl[x]], blkF_fcol[rl[y] - nA])].push_back(std::make_pair(rl[x], rl[y] - nA)); continue; }
				const int ci = blkA_ucol[rl[x]], cj = blkA_ucol[rl[y]];
				if (ci <= cj && (ci < cj || x == y)) HAp_terms[std::make_pair(cj, ci)].push_back(std::make_pair(rl[x], rl[y]));
			}
		}
	}
	std::map<std::pair<int, int>, int> HAp_index;  // (row,col) -> block id
	F.HAp_term_ptr.push_back(0);
	for (term_map_t::const_iterator kv = HAp_terms.begin(); kv != HAp_terms.end(); ++kv) {
		HAp_index[std::make_pair(kv->first.second, kv->first.first)] = (int)F.HAp_row.size();
		F.HAp_row.push_back(kv->first.second); F.HAp_col.push_back(kv->first.first);
		for (size_t t = 0; t < kv->second.size(); t++) { F.HAp_term_a.push_back(kv->second[t].first); F.HAp_term_b.push_back(kv->second[t].second); }
		F.HAp_term_ptr.push_back((int32_t)F.HAp_term_a.size());
	}
	F.HApf_term_ptr.push_back(0);
	F.rowApf_ptr.assign(n_p + 1, 0);
	std::vector<std::vector<int> > lm_cols(n_f);
	for (term_map_t::const_iterator kv = HApf_terms.begin(); kv != HApf_terms.end(); ++kv) {
		const int id = (int)F.HApf_row.size();
		F.HApf_row.push_back(kv->first.first); F.HApf_col.push_back(kv->first.second);
		F.rowApf_ptr[kv->first.first + 1]++;
		lm_cols[kv->first.second].push_back(id);
		for (size_t t = 0; t < kv->second.size(); t++) { F.HApf_term_a.push_back(kv->second[t].first); F.HApf_term_f.push_back(kv->second[t].second); }
		F.HApf_term_ptr.push_back((int32_t)F.HApf_term_a.size());
	}
	for (int i = 0; i < n_p; i++) F.rowApf_ptr[i + 1] += F.rowApf_ptr[i];
	F.colApf_ptr.push_back(0);
	for (int k = 0; k < n_f; k++) {
		for (size_t q = 0; q < lm_cols[k].size(); q++) F.colApf_idx.push_back(lm_cols[k][q]);
		F.colApf_ptr.push_back((int32_t)F.colApf_idx.size());
	}
	// O. Schur symbolic (schur.h:56-157): block (row j <= col i) gets one triplet per shared landmark, ascending landmark
	term_map_t S_terms;  // key (col,row) -> (HApf(row,k), HApf(col,k))
	for (term_map_t::const_iterator kv = HAp_terms.begin(); kv != HAp_terms.end(); ++kv) S_terms[kv->first];
	for (int k = 0; k < n_f; k++) {
		const std::vector<int> &E = lm_cols[k];
		for (size_t x = 0; x < E.size(); x++)
			for (size_t y = x; y < E.size(); y++) S_terms[std::make_pair(F.HApf_row[E[y]], F.HApf_row[E[x]])].push_back(std::make_pair(E[x], E[y]));
	}
	F.S_term_ptr.push_back(0);
	for (term_map_t::const_iterator kv = S_terms.begin(); kv != S_terms.end(); ++kv) {
		F.S_row.push_back(kv->first.second); F.S_col.push_back(kv->first.first);
		std::map<std::pair<int, int>, int>::const_iterator it = HAp_index.find(std::make_pair(kv->first.second, kv->first.first));
		F.S_HAp.push_back(it == HAp_index.end() ? -1 : it->second);
		for (size_t t = 0; t < kv->second.size(); t++) { F.S_term_rowblk.push_back(kv->second[t].first); F.S_term_colblk.push_back(kv->second[t].second); }
		F.S_term_ptr.push_back((int32_t)F.S_term_rowblk.size());
	}
	// values
	F.edge_pose.resize(F.edge_ids_touched.size() * S);
	for (size_t i = 0; i < F.edge_ids_touched.size(); i++) A::edge_pose_get(st, F.edge_ids_touched[i], &F.edge_pose[i * S]);
	F.lm_pos.resize(F.lm_ids_touched.size() * 3);
	for (size_t i = 0; i < F.lm_ids_touched.size(); i++) std::memcpy(&F.lm_pos[i * 3], A::lm_pos(st, F.lm_ids_touched[i]), 3 * sizeof(double));
	F.p.model = model;
	F.p.n_edges = (int32_t)F.edge_ids_touched.size(); F.p.n_unk_edges = n_p; F.p.n_lms = (int32_t)F.lm_ids_touched.size(); F.p.n_unk_lms = n_f;
	F.p.n_pairs = (int32_t)F.pairs.size(); F.p.n_path = (int32_t)F.path_edge.size(); F.p.n_obs = n_obs; F.p.n_blkA = nA; F.p.n_blkF = nF;
	F.p.n_HAp = (int32_t)F.HAp_row.size(); F.p.n_HAp_terms = (int32_t)F.HAp_term_a.size(); F.p.n_HApf = (int32_t)F.HApf_row.size();
	F.p.n_HApf_terms = (int32_t)F.HApf_term_a.size(); F.p.n_S = (int32_t)F.S_row.size(); F.p.n_S_terms = (int32_t)F.S_term_rowblk.size();
	F.p.n_total_obs_rows = n_rows_dup;
	F.bind();
	return fp;
}

/** Writes a result back into the state exactly as optimize_edges does (impl/optimize_edges.h:508-539 unknowns, 550-565 / 664-670 numeric
  * spanning tree incl. its stale-entry semantics, :269-276 validity). */
template <class A, class ST>
void scatter(ST &st, const Flat &F, const srba_result_t &r) {
	const int S = A::POSE_NVALS;
	for (size_t i = 0; i < F.k2k_ids.size(); i++) A::edge_pose_set(st, F.k2k_ids[i], r.edge_pose + i * S);
	for (size_t i = 0; i < F.lm_ids.size(); i++) std::memcpy(A::lm_pos(st, F.lm_ids[i]), r.lm_pos + i * 3, 3 * sizeof(double));
	for (size_t t = 0; t < F.pairs.size(); t++) {
		A::set_num(st, F.pairs[t].first, F.pairs[t].second, r.slot_pose + (2 * t) * S);
		A::set_num(st, F.pairs[t].second, F.pairs[t].first, r.slot_pose + (2 * t + 1) * S);
	}
	for (size_t l = 0; l < F.obs_global.size(); l++) A::set_validity(st, F.obs_global[l], (char)r.obs_valid[l]);
}

/** parameters.srba.feedback_user_iteration(iter, total_sq_err, mean_sqroot_error) replayed from the trace of a finished call: once with the
  * initial error (optimize_edges.h:417-418), then once per ACCEPTED step with the outer iteration counter it had (:595-596). */
template <class CALLBACK>
void replay_feedback(CALLBACK cb, const Flat &F, const srba_result_t &r) {
	if (!cb) return;
	const double n = (double)F.p.n_total_obs_rows;
	cb(0u, r.stats.total_sqr_error_init, std::sqrt(r.stats.total_sqr_error_init / n));
	unsigned iter = 0;
	for (int t = 0; t < r.trace_len; t++)
		if (r.trace[6 * t + 1] != 0 && r.trace[6 * t + 4] != 0) {
			cb(iter, r.trace[6 * t + 3], std::sqrt(r.trace[6 * t + 3] / n));
			iter++;
		}
}

}  // namespace srba_b200_shim
