// srba_b200::RbaEngine — host-side mirror of the reference's RbaEngine<> (include/srba/RbaEngine.h:66-816) for the
// hot path.  The CPU keeps the incremental relative-coordinate graph in INDEX-addressed arrays (no pointer graph):
//   edges / keyframes / landmarks / observations as flat vectors,
//   per-observation Jacobian rows (edge, direction, d+1) instead of std::map nodes with raw pointers,
//   per-column observation lists (ascending) instead of std::map columns,
// so that one optimize_edges() call flattens into the CSR arrays of include/srba_b200.h in O(N log N) — replacing the
// reference's quadratic per-call symbolic rebuilds (impl/sparse_hessian_build_symbolic.h:22-237, impl/schur.h:25-159)
// — and the whole numeric LM loop runs on the GPU (srba_b200_optimize).  Selection/ordering of unknowns, observations,
// spanning-tree paths and Hessian terms reproduces the reference bit-exactly.
#pragma once
#include <algorithm>
#include <array>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <map>
#include <memory>
#include <queue>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/srba_b200_engine.h"
#include "host_pose.hpp"

namespace srba_b200 {

typedef uint64_t TKeyFrameID;
typedef uint64_t TLandmarkID;
static const uint64_t INVALID_ID = static_cast<uint64_t>(-1);

#define SRBA_CHECK(cond)                                                                      \
	do {                                                                                      \
		if (!(cond)) {                                                                        \
			std::ostringstream s_;                                                            \
			s_ << "srba_b200: check failed: " #cond " (" << __FILE__ << ":" << __LINE__ << ")"; \
			throw std::runtime_error(s_.str());                                               \
		}                                                                                     \
	} while (0)

inline int obs_dims_of(int k) { return k == SRBA_OBS_MONO ? 2 : (k == SRBA_OBS_STEREO ? 4 : 3); }

// (block key, term) record of the symbolic Hessian / Schur construction, and its stable sort by (hi, lo): two counting-sort passes
struct TermRec { int32_t hi, lo, a, b; };
inline void sort_terms(std::vector<TermRec> &v, int n_hi, int n_lo) {
	if (v.size() < 2) return;
	std::vector<TermRec> tmp(v.size());
	std::vector<uint32_t> cnt((size_t)std::max(n_hi, n_lo) + 1);
	std::fill(cnt.begin(), cnt.begin() + n_lo + 1, 0u);
	for (const TermRec &r : v) cnt[r.lo + 1]++;
	for (int i = 0; i < n_lo; i++) cnt[i + 1] += cnt[i];
	for (const TermRec &r : v) tmp[cnt[r.lo]++] = r;
	std::fill(cnt.begin(), cnt.begin() + n_hi + 1, 0u);
	for (const TermRec &r : tmp) cnt[r.hi + 1]++;
	for (int i = 0; i < n_hi; i++) cnt[i + 1] += cnt[i];
	for (const TermRec &r : tmp) v[cnt[r.hi]++] = r;
}

// Per-thread id -> local-index tables of flatten(): entry i is valid when its stamp equals the stamp of the current call.
struct FlattenScratch {
	uint32_t stamp = 0;
	std::vector<uint32_t> obs_stamp, lm_stamp, edge_stamp, kf_stamp;
	std::vector<int32_t> obs_local, lm_local, edge_local, kf_first;
	static FlattenScratch &get() { static thread_local FlattenScratch s; return s; }
	void begin(size_t n_obs, size_t n_lm, size_t n_edge, size_t n_kf) {
		if (++stamp == 0) {   // wrapped: forget everything
			for (std::vector<uint32_t> *v : {&obs_stamp, &lm_stamp, &edge_stamp, &kf_stamp}) std::fill(v->begin(), v->end(), 0u);
			stamp = 1;
		}
		auto grow = [](std::vector<uint32_t> &st, std::vector<int32_t> &loc, size_t n) { if (st.size() < n) { st.resize(n + n / 2, 0u); loc.resize(n + n / 2, -1); } };
		grow(obs_stamp, obs_local, n_obs); grow(lm_stamp, lm_local, n_lm); grow(edge_stamp, edge_local, n_edge); grow(kf_stamp, kf_first, n_kf);
	}
};

// A flattened optimize_edges() problem + the ids needed to scatter the result back.
struct FlatProblem {
	srba_problem_t p;
	std::vector<double> edge_pose, lm_pos, obs_z;
	std::vector<int32_t> unk_edge, unk_lm, pair_path_ptr, path_edge, obs_slot, obs_lm, obs_mult, colA_ptr, blkA_obs, blkA_slotA, blkA_slotD, colF_ptr,
	    blkF_obs, blkF_slot, row_ptr, row_blk, HAp_row, HAp_col, HAp_term_ptr, HAp_term_a, HAp_term_b, HApf_row, HApf_col, HApf_term_ptr, HApf_term_a,
	    HApf_term_f, rowApf_ptr, colApf_ptr, colApf_idx, S_row, S_col, S_HAp, S_term_ptr, S_term_rowblk, S_term_colblk;
	std::vector<uint8_t> path_dir, blkA_normal_dir;
	// scatter-back ids
	std::vector<uint64_t> k2k_ids, lm_ids, obs_global, edge_ids_touched, lm_ids_touched;
	std::vector<std::pair<uint64_t, uint64_t> > pairs;  // (from>to)
	srba_opt_info_t info;  // integer counters already known at flatten time
	void bind() {
		p.edge_pose = edge_pose.data(); p.unk_edge = unk_edge.data(); p.lm_pos = lm_pos.data(); p.unk_lm = unk_lm.data();
		p.pair_path_ptr = pair_path_ptr.data(); p.path_edge = path_edge.data(); p.path_dir = path_dir.data();
		p.obs_z = obs_z.data(); p.obs_slot = obs_slot.data(); p.obs_lm = obs_lm.data(); p.obs_mult = obs_mult.data();
		p.colA_ptr = colA_ptr.data(); p.blkA_obs = blkA_obs.data(); p.blkA_slotA = blkA_slotA.data(); p.blkA_slotD = blkA_slotD.data();
		p.blkA_normal_dir = blkA_normal_dir.data(); p.colF_ptr = colF_ptr.data(); p.blkF_obs = blkF_obs.data(); p.blkF_slot = blkF_slot.data();
		p.row_ptr = row_ptr.data(); p.row_blk = row_blk.data();
		p.HAp_row = HAp_row.data(); p.HAp_col = HAp_col.data(); p.HAp_term_ptr = HAp_term_ptr.data(); p.HAp_term_a = HAp_term_a.data(); p.HAp_term_b = HAp_term_b.data();
		p.HApf_row = HApf_row.data(); p.HApf_col = HApf_col.data(); p.HApf_term_ptr = HApf_term_ptr.data(); p.HApf_term_a = HApf_term_a.data();
		p.HApf_term_f = HApf_term_f.data(); p.rowApf_ptr = rowApf_ptr.data(); p.colApf_ptr = colApf_ptr.data(); p.colApf_idx = colApf_idx.data();
		p.S_row = S_row.data(); p.S_col = S_col.data(); p.S_HAp = S_HAp.data(); p.S_term_ptr = S_term_ptr.data();
		p.S_term_rowblk = S_term_rowblk.data(); p.S_term_colblk = S_term_colblk.data();
	}
};

class EngineBase {
  public:
	virtual ~EngineBase() {}
	srba_engine_params_t parameters;
	bool record_problems = false;
	std::deque<std::unique_ptr<FlatProblem> > recorded;
	int trace_cap = 0;
	std::vector<double> last_trace, last_trace_delta;
	uint64_t last_trace_scalars = 0;
	double flatten_seconds = 0, gpu_seconds = 0;
	uint64_t total_trials = 0;
};

template <class POSE>
class RbaEngine : public EngineBase {
  public:
	static const int P = POSE::P, S = POSE::S;

	struct Edge { TKeyFrameID from, to; POSE inv_pose; };                    // k2k_edge_t (srba_types.h:82-91)
	struct KeyFrame { std::vector<uint32_t> adj_k2k, adj_k2f; };              // keyframe_info (srba_types.h:528-532)
	struct Landmark {                                                         // TLandmarkEntry + TRelativeLandmarkPos
		bool exists = false, known = true;
		TKeyFrameID base = 0;
		double pos[3] = {0, 0, 0};
		int32_t fcol = -1;  // dh_df column (appendCol order)
	};
	struct RowEntry { uint32_t edge; uint8_t normal_dir; TKeyFrameID kf_d; }; // one dh_dAp block (srba_types.h:204-240)
	struct Observation {                                                      // k2f_edge_t (srba_types.h:501-525)
		TKeyFrameID kf;
		TLandmarkID lm;
		double z[4];
		bool known_lm, first_obs_of_unknown, has_F;
		uint32_t row_begin, row_len;
	};
	struct STE { TKeyFrameID next; uint64_t distance; };                      // TSpanTreeEntry (srba_types.h:538-542)
	struct new_kf_observation_t { TLandmarkID feat_id; double obs_data[4]; bool is_fixed, is_unknown_with_init_val; double feat_rel_pos[3]; };
	struct TNewEdgeInfo { uint64_t id = INVALID_ID; bool has_approx_init_val = false; TKeyFrameID loopclosure_observer_kf = INVALID_ID, loopclosure_base_kf = INVALID_ID; };
	struct TOptimizeLocalAreaParams {
		bool optimize_k2k_edges = true, optimize_landmarks = true;
		TKeyFrameID max_visitable_kf_id = INVALID_ID;
		size_t dont_optimize_landmarks_seen_less_than_n_times = 2;
	};
	struct TNewKeyFrameInfo { TKeyFrameID kf_id; std::vector<TNewEdgeInfo> created_edge_ids; srba_opt_info_t optimize_results, optimize_results_stg1; };

	// ---- state (TRBA_Problem_state, srba_types.h:548-785)
	std::vector<KeyFrame> keyframes;
	std::vector<Edge> k2k_edges;
	std::vector<Landmark> all_lms;
	std::vector<Observation> all_observations;
	std::vector<RowEntry> rows;
	std::vector<std::vector<uint32_t> > colA, colF;  // dh_dAp / dh_df columns: ascending observation indices
	std::vector<uint8_t> validity;                   // all_observations_Jacob_validity
	std::vector<std::map<TKeyFrameID, STE> > next_edge;                       // spanning_tree.sym.next_edge
	std::vector<std::map<TKeyFrameID, std::vector<uint32_t> > > all_edges;    // spanning_tree.sym.all_edges (edge ids)
	std::vector<std::map<TKeyFrameID, POSE> > num;                            // spanning_tree.num (values only)
	std::set<uint64_t> last_timestep_touched_kfs;
	std::map<TLandmarkID, std::array<double, 9> > unknown_lms_inf_matrices;
	int O = 3;
	SE3 sensor_pose3, right_cam_R2L;

	explicit RbaEngine(const srba_engine_params_t &p) { set_params(p); }
	void set_params(const srba_engine_params_t &p) {  // checks first: a rejected parameter set leaves the engine untouched
		SRBA_CHECK(p.pose_kind == (P == 3 ? SRBA_POSE_SE2 : SRBA_POSE_SE3));
		if (P == 3) SRBA_CHECK(p.obs_kind == SRBA_OBS_RELPOSES2D && p.sensor_pose_kind == SRBA_SENSORPOSE_NONE);
		if (P == 6) SRBA_CHECK(p.obs_kind == SRBA_OBS_CARTESIAN3D || p.obs_kind == SRBA_OBS_MONO || p.obs_kind == SRBA_OBS_STEREO);
		SRBA_CHECK(p.max_tree_depth >= 1);
		parameters = p;
		O = obs_dims_of(p.obs_kind);
		sensor_pose3 = SE3::from_xyzypr(p.sensor_pose);
		right_cam_R2L = SE3::from_xyzquat(p.right_cam_pose).inverse();
	}
	template <class V> static typename V::value_type &grow(V &v, size_t i) { if (i >= v.size()) v.resize(i + 1); return v[i]; }

	// ------------------------------------------------------------ graph construction
	TKeyFrameID alloc_keyframe() { keyframes.push_back(KeyFrame()); return keyframes.size() - 1; }  // impl/alloc_keyframe.h:19-28
	size_t alloc_kf2kf_edge(TKeyFrameID a, TKeyFrameID b, const POSE &init) {                        // impl/alloc_kf2kf_edge.h:17-61
		SRBA_CHECK(a != b);
		Edge e; e.from = a; e.to = b; e.inv_pose = init;
		k2k_edges.push_back(e);
		const uint32_t id = k2k_edges.size() - 1;
		keyframes[a].adj_k2k.push_back(id);
		keyframes[b].adj_k2k.push_back(id);
		colA.push_back(std::vector<uint32_t>());
		return id;
	}
	size_t create_kf2kf_edge(TKeyFrameID new_kf_id, TKeyFrameID a, TKeyFrameID b, const POSE &init = POSE()) {  // impl/create_kf2kf_edge.h:15-37
		const size_t id = alloc_kf2kf_edge(a, b, init);
		update_symbolic_new_node(new_kf_id, a, b, parameters.max_tree_depth);
		return id;
	}
	TKeyFrameID other_end(TKeyFrameID one, uint32_t e) const { return k2k_edges[e].from == one ? k2k_edges[e].to : k2k_edges[e].from; }

	// find_path_bfs (impl/spantree_update_symbolic.h:227-300): BFS with adjacency-order tie-breaking, edges from cur to trg.
	bool find_path_bfs(TKeyFrameID cur_node, TKeyFrameID trg_node, std::vector<uint32_t> &out_edges) const {
		out_edges.clear();
		if (cur_node == trg_node) return true;
		struct Prec { TKeyFrameID prev; uint32_t edge; uint64_t dist; };
		std::unordered_map<TKeyFrameID, Prec> prec;
		std::queue<TKeyFrameID> pending;
		pending.push(cur_node);
		prec[cur_node] = Prec{0, 0, 0};
		while (!pending.empty()) {
			const TKeyFrameID next_kf = pending.front();
			pending.pop();
			const uint64_t cur_dist = prec[next_kf].dist;
			if (next_kf == trg_node) {
				out_edges.resize(cur_dist);
				uint64_t d = cur_dist;
				TKeyFrameID node = trg_node;
				while (node != cur_node) { const Prec &pr = prec[node]; out_edges[--d] = pr.edge; node = pr.prev; }
				return true;
			}
			for (uint32_t e : keyframes[next_kf].adj_k2k) {
				const TKeyFrameID nk = other_end(next_kf, e);
				if (!prec.count(nk)) { pending.push(nk); prec[nk] = Prec{next_kf, e, cur_dist + 1}; }
			}
		}
		return false;
	}

	// BFS half of create_complete_spanning_tree (impl/spantree_create_complete.h:35-83): visiting order = adjacency order,
	// a keyframe's tree parent is the first dequeued neighbour that reaches it.  Stops early once every keyframe of `wanted`
	// (if given) has been reached -- parents are final at first visit, so the tree below is the same.
	struct BfsNode { TKeyFrameID prev; uint32_t edge; uint64_t dist; };
	void complete_spanning_tree_bfs(TKeyFrameID root, uint64_t max_depth, std::map<TKeyFrameID, BfsNode> &prec, const std::set<TKeyFrameID> *wanted = nullptr) const {
		prec.clear();
		SRBA_CHECK(root < keyframes.size());
		std::queue<TKeyFrameID> pending;
		pending.push(root);
		prec[root] = BfsNode{root, 0, 0};
		size_t missing = 0;
		if (wanted) for (TKeyFrameID w : *wanted) missing += (w != root) ? 1 : 0;
		while (!pending.empty() && !(wanted && missing == 0)) {
			const TKeyFrameID next_kf = pending.front();
			pending.pop();
			const uint64_t cur_dist = prec[next_kf].dist;
			if (cur_dist >= max_depth) continue;
			for (uint32_t e : keyframes[next_kf].adj_k2k) {
				const TKeyFrameID nk = other_end(next_kf, e);
				if (!prec.count(nk)) {
					pending.push(nk);
					prec[nk] = BfsNode{next_kf, e, cur_dist + 1};
					if (wanted && wanted->count(nk)) missing--;
				}
			}
		}
	}
	// chain of (edge, dir) from root to target through the BFS tree; dir 1: the edge's `from` is the node being reached
	// (compose inv_pose), 0: its `to` is (compose -inv_pose)  (spantree_create_complete.h:117-122)
	void bfs_chain(const std::map<TKeyFrameID, BfsNode> &prec, TKeyFrameID root, TKeyFrameID target, std::vector<int32_t> &edges, std::vector<uint8_t> &dirs) const {
		std::vector<std::pair<int32_t, uint8_t> > rev;
		for (TKeyFrameID node = target; node != root;) {
			const BfsNode &b = prec.at(node);
			rev.push_back(std::make_pair((int32_t)b.edge, (uint8_t)(k2k_edges[b.edge].to == node ? 0 : 1)));
			node = b.prev;
		}
		for (size_t i = rev.size(); i-- > 0;) { edges.push_back(rev[i].first); dirs.push_back(rev[i].second); }
	}
	void fill_eval_model(srba_model_t &m) const {
		std::memset(&m, 0, sizeof(m));
		m.pose_kind = parameters.pose_kind; m.obs_kind = parameters.obs_kind; m.sensor_pose_kind = parameters.sensor_pose_kind; m.noise_kind = parameters.noise_kind;
		std::memcpy(m.sensor_pose, sensor_pose3.v, sizeof(m.sensor_pose));
		m.noise_std = parameters.noise_std;
		std::memcpy(m.cam_left, parameters.cam_left, sizeof(m.cam_left));
		std::memcpy(m.cam_right, parameters.cam_right, sizeof(m.cam_right));
		std::memcpy(m.right_cam_R2L, right_cam_R2L.v, sizeof(m.right_cam_R2L));
	}
	std::vector<double> all_edge_poses() const {
		std::vector<double> ep(k2k_edges.size() * S + 1);
		for (size_t e = 0; e < k2k_edges.size(); e++) std::memcpy(&ep[e * S], k2k_edges[e].inv_pose.v, S * sizeof(double));
		return ep;
	}

	// create_complete_spanning_tree (impl/spantree_create_complete.h:18-127): ids ascending (the reference returns a std::map),
	// poses of every reached keyframe as seen from root, composed on the device.
	void create_complete_spanning_tree(TKeyFrameID root, uint64_t max_depth, std::vector<TKeyFrameID> &ids, std::vector<double> &poses) const {
		std::map<TKeyFrameID, BfsNode> prec;
		complete_spanning_tree_bfs(root, max_depth, prec);
		ids.clear();
		std::vector<int32_t> ptr(1, 0), pe;
		std::vector<uint8_t> pd;
		for (const auto &kv : prec) {
			ids.push_back(kv.first);
			bfs_chain(prec, root, kv.first, pe, pd);
			ptr.push_back((int32_t)pe.size());
		}
		srba_eval_problem_t q;
		std::memset(&q, 0, sizeof(q));
		fill_eval_model(q.model);
		const std::vector<double> ep = all_edge_poses();
		q.n_edges = (int32_t)k2k_edges.size(); q.n_pairs = (int32_t)ids.size(); q.n_path = (int32_t)pe.size();
		q.edge_pose = ep.data(); q.pair_path_ptr = ptr.data(); q.path_edge = pe.data(); q.path_dir = pd.data();
		poses.assign(ids.size() * S, 0.0);
		double unused = 0;
		if (srba_b200_eval_overall_squared_error(&q, &unused, poses.data()) != SRBA_OK)
			throw std::runtime_error(std::string("srba_b200_eval_overall_squared_error failed: ") + srba_b200_last_error());
	}

	// eval_overall_squared_error (impl/eval_overall_error.h:15-137): integer work (pairs, BFS trees, chains) here, every pose
	// composition, observation error and the sum on the device.
	double eval_overall_squared_error() const {
		std::map<TKeyFrameID, std::set<TKeyFrameID> > ob_pairs;  // minimum id first (:29-39)
		for (const Observation &ob : all_observations) {
			const TKeyFrameID base = all_lms[ob.lm].base;
			if (base != ob.kf) ob_pairs[std::min(ob.kf, base)].insert(std::max(ob.kf, base));
		}
		std::map<std::pair<TKeyFrameID, TKeyFrameID>, int32_t> pair_idx;
		std::vector<int32_t> ptr(1, 0), pe;
		std::vector<uint8_t> pd;
		std::map<TKeyFrameID, BfsNode> prec;
		for (const auto &rp : ob_pairs) {
			complete_spanning_tree_bfs(rp.first, std::numeric_limits<uint64_t>::max(), prec, &rp.second);
			for (TKeyFrameID w : rp.second) {
				if (!prec.count(w)) continue;  // not connected: the reference's ASSERT_ fires when the observation is evaluated (:106-110)
				pair_idx[std::make_pair(rp.first, w)] = (int32_t)ptr.size() - 1;
				bfs_chain(prec, rp.first, w, pe, pd);
				ptr.push_back((int32_t)pe.size());
			}
		}
		const size_t n_obs = all_observations.size();
		std::vector<double> z(n_obs * O + 1), lm(all_lms.size() * 3 + 1);
		std::vector<int32_t> op(n_obs + 1), ol(n_obs + 1);
		std::vector<uint8_t> oi(n_obs + 1);
		for (size_t i = 0; i < n_obs; i++) {
			const Observation &ob = all_observations[i];
			const TKeyFrameID base = all_lms[ob.lm].base;
			for (int k = 0; k < O; k++) z[i * O + k] = ob.z[k];
			ol[i] = (int32_t)ob.lm;
			if (base == ob.kf) { op[i] = -1; oi[i] = 0; continue; }
			const auto it = pair_idx.find(std::make_pair(std::min(ob.kf, base), std::max(ob.kf, base)));
			SRBA_CHECK(it != pair_idx.end());  // ASSERT_(itRelPose != ...end())  (:109)
			op[i] = it->second;
			oi[i] = ob.kf > base ? 1 : 0;  // pose of base as seen from the observer: the tree pose when the observer is the root
		}
		for (size_t l = 0; l < all_lms.size(); l++) std::memcpy(&lm[l * 3], all_lms[l].pos, 3 * sizeof(double));
		srba_eval_problem_t q;
		std::memset(&q, 0, sizeof(q));
		fill_eval_model(q.model);
		const std::vector<double> ep = all_edge_poses();
		q.n_edges = (int32_t)k2k_edges.size(); q.n_pairs = (int32_t)ptr.size() - 1; q.n_path = (int32_t)pe.size(); q.n_obs = (int32_t)n_obs; q.n_lms = (int32_t)all_lms.size();
		q.edge_pose = ep.data(); q.pair_path_ptr = ptr.data(); q.path_edge = pe.data(); q.path_dir = pd.data();
		q.obs_z = z.data(); q.obs_pair = op.data(); q.obs_inverse = oi.data(); q.obs_lm = ol.data(); q.lm_pos = lm.data();
		double sqerr = 0;
		if (srba_b200_eval_overall_squared_error(&q, &sqerr, nullptr) != SRBA_OK)
			throw std::runtime_error(std::string("srba_b200_eval_overall_squared_error failed: ") + srba_b200_last_error());
		return sqerr;
	}

	// update_symbolic_new_node (impl/spantree_update_symbolic.h:19-211)
	void update_symbolic_new_node(TKeyFrameID new_node_id, TKeyFrameID e_first, TKeyFrameID e_second, uint64_t max_depth) {
		SRBA_CHECK(max_depth >= 1);
		std::set<std::pair<TKeyFrameID, TKeyFrameID> > modified;
		const TKeyFrameID ik = (e_first == new_node_id) ? e_second : e_first;
		grow(next_edge, std::max<size_t>(std::max(ik, new_node_id), keyframes.size() - 1));  // no reallocation below
		std::vector<std::pair<TKeyFrameID, uint64_t> > tk;  // snapshot of distances
		for (const auto &kv : next_edge[ik]) if (kv.second.distance < max_depth) tk.push_back(std::make_pair(kv.first, kv.second.distance));
		tk.push_back(std::make_pair(ik, uint64_t(0)));
		std::vector<TKeyFrameID> STn;  // ids only; distances are read live below
		for (const auto &kv : next_edge[new_node_id]) STn.push_back(kv.first);
		STn.push_back(new_node_id);
		for (size_t r_idx = 0; r_idx < STn.size(); r_idx++) {
			const TKeyFrameID r = STn[r_idx];
			const bool r_is_new = (r == new_node_id);
			const uint64_t dist_r2n = r_is_new ? 0 : next_edge[new_node_id].at(r).distance;  // live value of ST[n][r]
			for (size_t s_idx = 0; s_idx < tk.size(); s_idx++) {
				const TKeyFrameID s = tk[s_idx].first;
				if (r == s) continue;
				const uint64_t dist_s2ik = tk[s_idx].second;
				std::map<TKeyFrameID, STE> &st_r = grow(next_edge, r);
				std::map<TKeyFrameID, STE> &st_s = grow(next_edge, s);
				const uint64_t new_dist = dist_r2n + dist_s2ik + 1;
				auto it_s_inSTr = st_r.find(s);
				const bool found = it_s_inSTr != st_r.end();
				if ((found && new_dist < it_s_inSTr->second.distance) || (!found && new_dist <= max_depth)) {
					const TKeyFrameID next_r = r_is_new ? ik : st_r.at(new_node_id).next;  // towards new_node_id
					const TKeyFrameID next_s = (s == ik) ? new_node_id : st_s.at(ik).next;  // towards ik
					STE &a = st_r[s];
					a.distance = new_dist; a.next = next_r;
					SRBA_CHECK(r != a.next);
					STE &b = st_s[r];
					b.distance = new_dist; b.next = next_s;
					SRBA_CHECK(s != b.next);
					modified.insert(std::make_pair(s, r));
					modified.insert(std::make_pair(r, s));
				}
			}
		}
		for (const auto &pr : modified) {  // all_edges rebuild (:167-188); (a,b) and (b,a) are always inserted together
			if (pr.first < pr.second) continue;
			const TKeyFrameID from = std::max(pr.first, pr.second), to = std::min(pr.first, pr.second);
			std::vector<uint32_t> &path = grow(all_edges, from)[to];
			const bool ok = find_path_bfs(from, to, path);
			SRBA_CHECK(ok);
		}
	}

	// numeric chain of one stored path (impl/spantree_update_numeric.h:47-65) — host copy, only for update_numeric_all()
	POSE compose_path(TKeyFrameID from, const std::vector<uint32_t> &path) const {
		POSE accum;
		TKeyFrameID cur = from;
		for (uint32_t e : path) {
			if (k2k_edges[e].to == cur) { accum = POSE::compose(accum, k2k_edges[e].inv_pose); cur = k2k_edges[e].from; }
			else { accum = POSE::compose(accum, k2k_edges[e].inv_pose.inverse()); cur = k2k_edges[e].to; }
		}
		return accum;
	}
	void update_numeric_all() {  // TSpanningTree::update_numeric(false) — test/support path only, not used by the optimiser
		for (size_t from = 0; from < all_edges.size(); from++)
			for (const auto &kv : all_edges[from]) {
				const POSE a = compose_path(from, kv.second);
				grow(num, from)[kv.first] = a;
				grow(num, kv.first)[from] = a.inverse();
			}
	}
	const POSE *get_kf_relative_pose(TKeyFrameID kf_query, TKeyFrameID kf_reference) const {  // RbaEngine.h:385-396
		if (kf_reference >= num.size()) return nullptr;
		auto it = num[kf_reference].find(kf_query);
		return it != num[kf_reference].end() ? &it->second : nullptr;
	}

	// ------------------------------------------------------------ sensor inverse models (models/sensors.h:122-141,293-315,405-413,826-834)
	void inverse_sensor_model(double *out, const double *obs) const {
		switch (parameters.obs_kind) {
			case SRBA_OBS_MONO:
				out[0] = (obs[0] - parameters.cam_left[2]) / parameters.cam_left[0];
				out[1] = (obs[1] - parameters.cam_left[3]) / parameters.cam_left[1];
				out[2] = 1;
				break;
			case SRBA_OBS_STEREO: {
				const double fxl = parameters.cam_left[0], fyl = parameters.cam_left[1], cxl = parameters.cam_left[2], cyl = parameters.cam_left[3];
				const double disparity = std::max(0.001f, static_cast<float>(obs[0]) - static_cast<float>(obs[2]));
				const double baseline = parameters.right_cam_pose[0];
				SRBA_CHECK(baseline != 0);
				const double Z = fxl * baseline / disparity;
				out[0] = (obs[0] - cxl) * Z / fxl; out[1] = (obs[1] - cyl) * Z / fyl; out[2] = Z;
				break;
			}
			default: out[0] = obs[0]; out[1] = obs[1]; out[2] = obs[2];
		}
		if (parameters.sensor_pose_kind == SRBA_SENSORPOSE_SE3) sensor_pose3.compose_point(out, out);  // sensor2robot_point
	}
	// landmark_matcher<OBS>::find_relative_pose
	bool find_relative_pose(const std::vector<const double *> &new_kf_obs, const std::vector<const double *> &old_kf_obs, POSE &out) const {
		const size_t N = new_kf_obs.size();
		switch (parameters.obs_kind) {
			case SRBA_OBS_RELPOSES2D:
				for (size_t i = 0; i < N; i++) {
					const double *a = new_kf_obs[i], *b = old_kf_obs[i];
					if ((a[0] != 0 || a[1] != 0 || a[2] != 0) && (b[0] != 0 || b[1] != 0 || b[2] != 0)) continue;
					from_se3(to_se3(SE2::ominus(SE2(b[0], b[1], b[2]), SE2(a[0], a[1], a[2]))), out);
					return true;
				}
				return false;
			case SRBA_OBS_MONO: return false;
			case SRBA_OBS_CARTESIAN3D: {
				std::vector<double> a, b;
				for (size_t i = 0; i < N; i++) for (int k = 0; k < 3; k++) { a.push_back(old_kf_obs[i][k]); b.push_back(new_kf_obs[i][k]); }
				SE3 T;
				if (!absolute_orientation(a, b, T)) return false;
				from_se3(T, out);
				return true;
			}
			case SRBA_OBS_STEREO: {
				const double cx = parameters.cam_left[2], cy = parameters.cam_left[3], bl = parameters.right_cam_pose[0], fl = parameters.cam_left[0];
				std::vector<double> a, b;
				for (size_t i = 0; i < N; i++) {
					const double d_old = old_kf_obs[i][0] - old_kf_obs[i][2], d_new = new_kf_obs[i][0] - new_kf_obs[i][2];
					if (d_old <= .0 || d_new <= .0) continue;
					a.push_back((old_kf_obs[i][0] - cx) * bl / d_old); a.push_back((old_kf_obs[i][1] - cy) * bl / d_old); a.push_back(fl * bl / d_old);
					b.push_back((new_kf_obs[i][0] - cx) * bl / d_new); b.push_back((new_kf_obs[i][1] - cy) * bl / d_new); b.push_back(fl * bl / d_new);
				}
				SE3 T;
				if (!absolute_orientation(a, b, T)) return false;
				from_se3(T, out);
				return true;
			}
		}
		return false;
	}

	// ------------------------------------------------------------ add_observation (impl/add-observations.h:17-264)
	size_t add_observation(TKeyFrameID observing_kf_id, TLandmarkID feat_id, const double *obs_data, const double *fixed_rel_pos, const double *unk_init) {
		SRBA_CHECK(!(fixed_rel_pos && unk_init));
		const bool is_1st_time_seen = (feat_id >= all_lms.size() || !all_lms[feat_id].exists);
		const bool is_fixed = fixed_rel_pos != nullptr || (!is_1st_time_seen && all_lms[feat_id].known);
		const size_t new_obs_idx = all_observations.size();
		if (is_1st_time_seen) {
			Landmark &L = grow(all_lms, feat_id);
			L.exists = true; L.known = is_fixed; L.base = observing_kf_id;
			if (is_fixed) std::memcpy(L.pos, fixed_rel_pos, sizeof(L.pos));
			else {
				if (unk_init) std::memcpy(L.pos, unk_init, sizeof(L.pos));
				else inverse_sensor_model(L.pos, obs_data);
				L.fcol = colF.size();
				colF.push_back(std::vector<uint32_t>());
			}
		}
		const Landmark &L = all_lms[feat_id];
		const TKeyFrameID base_id = L.base;
		Observation ob;
		ob.kf = observing_kf_id; ob.lm = feat_id;
		for (int i = 0; i < 4; i++) ob.z[i] = i < O ? obs_data[i] : 0.0;
		ob.first_obs_of_unknown = is_1st_time_seen && !is_fixed;
		ob.known_lm = is_fixed;
		ob.row_begin = rows.size(); ob.row_len = 0; ob.has_F = false;
		keyframes[observing_kf_id].adj_k2f.push_back(new_obs_idx);
		validity.push_back(1);
		bool ignore_obs = false;
		if (base_id != observing_kf_id) {
			const TKeyFrameID from = std::max(observing_kf_id, base_id), to = std::min(observing_kf_id, base_id);
			const bool all_edges_inverse = (from != observing_kf_id);
			if (from >= all_edges.size()) throw std::runtime_error("No ST.all_edges found for observing_id / base_id");
			auto it = all_edges[from].find(to);
			if (it != all_edges[from].end()) {
				const std::vector<uint32_t> &path = it->second;
				SRBA_CHECK(!path.empty());
				TKeyFrameID curKF = observing_kf_id;
				for (size_t j = 0; j < path.size(); j++) {
					const size_t i = all_edges_inverse ? (path.size() - j - 1) : j;
					SRBA_CHECK(curKF != to);
					const uint32_t e = path[i];
					const bool normal_dir = (k2k_edges[e].to == curKF);
					colA[e].push_back(new_obs_idx);
					rows.push_back(RowEntry{e, (uint8_t)normal_dir, curKF});
					ob.row_len++;
					// placeholders of the numeric spanning tree are created here by operator[] in the reference (:196-204)
					grow(num, curKF)[base_id];
					if (curKF != observing_kf_id) grow(num, observing_kf_id)[curKF];
					curKF = normal_dir ? k2k_edges[e].from : k2k_edges[e].to;
				}
			} else ignore_obs = true;
		}
		if (!is_fixed && !ignore_obs) {
			colF[L.fcol].push_back(new_obs_idx);
			ob.has_F = true;
			if (!ob.first_obs_of_unknown) grow(num, observing_kf_id)[base_id];
		}
		all_observations.push_back(ob);
		return new_obs_idx;
	}

	// ------------------------------------------------------------ edge-creation policy (ecps/local_areas_fixed_size.h:52-213)
	typedef std::multimap<size_t, TKeyFrameID, std::greater<size_t> > base_sorted_lst_t;
	TKeyFrameID center_of(TKeyFrameID kf) const { return parameters.submap_size * (kf / parameters.submap_size); }
	void ecp_eval(TKeyFrameID new_kf_id, const std::vector<new_kf_observation_t> &obs, std::vector<TNewEdgeInfo> &out) {
		SRBA_CHECK(new_kf_id >= 1);
		const TKeyFrameID cur_center = center_of(new_kf_id);
		const uint64_t min_dist_for_loop_closure = parameters.max_tree_depth + 1;
		std::map<TKeyFrameID, size_t> obs_for_each_base;  // make_ordered_list_base_kfs.h:16-43
		for (const auto &o : obs) {
			if (o.feat_id >= all_lms.size() || !all_lms[o.feat_id].exists) continue;
			obs_for_each_base[all_lms[o.feat_id].base]++;
		}
		base_sorted_lst_t sorted_bases;
		for (const auto &kv : obs_for_each_base) sorted_bases.insert(std::make_pair(kv.second, kv.first));
		std::map<TKeyFrameID, size_t> obs_for_each_area;
		std::map<TKeyFrameID, bool> base_is_center;
		std::map<TKeyFrameID, std::map<TKeyFrameID, size_t> > grouped;
		for (const auto &it : sorted_bases) {
			const TKeyFrameID c = center_of(it.second);
			obs_for_each_area[c] += it.first;
			grouped[c][it.second] += it.first;
			if (!base_is_center.count(c)) base_is_center[c] = true;
			if (it.second != c) base_is_center[c] = false;
		}
		base_sorted_lst_t areas_sorted;
		for (const auto &kv : obs_for_each_area) areas_sorted.insert(std::make_pair(kv.second, kv.first));
		std::map<TKeyFrameID, base_sorted_lst_t> grouped_sorted;
		for (const auto &kv : grouped)
			for (const auto &kv2 : kv.second) grouped_sorted[kv.first].insert(std::make_pair(kv2.second, kv2.first));
		if (cur_center != new_kf_id) {
			TNewEdgeInfo nei;
			nei.id = create_kf2kf_edge(new_kf_id, cur_center, new_kf_id);
			out.push_back(nei);
		}
		for (const auto &it : areas_sorted) {
			const size_t num_obs_this_base = it.first;
			const TKeyFrameID remote_center = it.second;
			if (cur_center == remote_center) continue;
			uint64_t found_distance = std::numeric_limits<uint64_t>::max();
			if (cur_center < next_edge.size()) {
				auto f = next_edge[cur_center].find(remote_center);
				if (f != next_edge[cur_center].end()) found_distance = f->second.distance;
			}
			uint64_t extra = 2;
			if (cur_center == new_kf_id) extra--;
			if (base_is_center[remote_center]) extra--;
			if (found_distance >= min_dist_for_loop_closure - extra && num_obs_this_base >= parameters.min_obs_to_loop_closure) {
				TNewEdgeInfo nei;
				nei.id = create_kf2kf_edge(cur_center, remote_center, cur_center);
				nei.loopclosure_observer_kf = new_kf_id;
				nei.loopclosure_base_kf = grouped_sorted[remote_center].begin()->second;
				out.push_back(nei);
			}
		}
		if (out.empty()) throw std::runtime_error("ECP: no suitable linking KF found: the node becomes isolated of the graph!");
	}

	// determine_kf2kf_edges_to_create (impl/determine_kf2kf_edges_to_create.h:17-266)
	void collect_matches(bool from_new_obs, TKeyFrameID last_kf, TKeyFrameID other_kf, const std::vector<new_kf_observation_t> &obs,
	                     std::vector<const double *> &new_kf_obs, std::vector<const double *> &old_kf_obs) const {
		new_kf_obs.clear(); old_kf_obs.clear();
		std::map<TLandmarkID, size_t> feats;
		const std::vector<uint32_t> *last_k2f = nullptr;
		if (!from_new_obs) {
			last_k2f = &keyframes[last_kf].adj_k2f;
			for (size_t k = 0; k < last_k2f->size(); k++) feats[all_observations[(*last_k2f)[k]].lm] = k;
		} else
			for (size_t k = 0; k < obs.size(); k++) feats[obs[k].feat_id] = k;
		for (uint32_t oi : keyframes[other_kf].adj_k2f) {
			auto it = feats.find(all_observations[oi].lm);
			if (it == feats.end()) continue;
			old_kf_obs.push_back(all_observations[oi].z);
			new_kf_obs.push_back(from_new_obs ? obs[it->second].obs_data : all_observations[(*last_k2f)[it->second]].z);
		}
	}
	void determine_kf2kf_edges_to_create(TKeyFrameID new_kf_id, const std::vector<new_kf_observation_t> &obs, std::vector<TNewEdgeInfo> &neis) {
		neis.clear();
		if (keyframes.size() == 1) return;
		ecp_eval(new_kf_id, obs, neis);
		for (auto &nei : neis) {
			if (nei.has_approx_init_val) continue;
			Edge &ed = k2k_edges[nei.id];
			const bool not_touching = (ed.to != new_kf_id) && (ed.from != new_kf_id);
			if (!not_touching && last_timestep_touched_kfs.count(ed.from)) {  // method #1
				const POSE *rel = get_kf_relative_pose(new_kf_id - 1, ed.from);
				if (rel) { ed.inv_pose = (ed.to == new_kf_id) ? rel->inverse() : *rel; nei.has_approx_init_val = true; }
			}
			if (nei.has_approx_init_val) continue;
			std::vector<const double *> new_kf_obs, old_kf_obs;  // method #2
			if (not_touching) collect_matches(false, ed.from, ed.to, obs, new_kf_obs, old_kf_obs);
			else collect_matches(true, new_kf_id, (ed.to == new_kf_id) ? ed.from : ed.to, obs, new_kf_obs, old_kf_obs);
			POSE pose_new_wrt_old;
			bool found_ok = find_relative_pose(new_kf_obs, old_kf_obs, pose_new_wrt_old);
			if (!found_ok && nei.loopclosure_observer_kf != INVALID_ID && nei.loopclosure_base_kf != INVALID_ID) {
				collect_matches(nei.loopclosure_observer_kf == new_kf_id, nei.loopclosure_observer_kf, nei.loopclosure_base_kf, obs, new_kf_obs, old_kf_obs);
				found_ok = find_relative_pose(new_kf_obs, old_kf_obs, pose_new_wrt_old);
			}
			if (!found_ok) continue;
			SE3 sp;
			if (parameters.sensor_pose_kind == SRBA_SENSORPOSE_SE3) sp = sensor_pose3;
			from_se3(SE3::compose(SE3::compose(sp, to_se3(pose_new_wrt_old)), sp.inverse()), pose_new_wrt_old);
			const bool edge_dir_to_newkf = (ed.to == new_kf_id);
			nei.has_approx_init_val = true;
			if (!not_touching) ed.inv_pose = edge_dir_to_newkf ? pose_new_wrt_old.inverse() : pose_new_wrt_old;
			else {
				const POSE ident;
				auto rel = [&](TKeyFrameID q, TKeyFrameID ref) -> const POSE * { return q == ref ? &ident : get_kf_relative_pose(q, ref); };
				const POSE *obs_wrt_to = rel(nei.loopclosure_observer_kf, ed.to), *base_wrt_to = rel(nei.loopclosure_base_kf, ed.to);
				const POSE *obs_wrt_from = rel(nei.loopclosure_observer_kf, ed.from), *base_wrt_from = rel(nei.loopclosure_base_kf, ed.from);
				const bool near_to = (obs_wrt_to || base_wrt_from) || !(obs_wrt_from || base_wrt_to);
				const POSE *obs_wrt_local = near_to ? (obs_wrt_to ? obs_wrt_to : &ident) : (obs_wrt_from ? obs_wrt_from : &ident);
				const POSE *base_wrt_remote = near_to ? (base_wrt_from ? base_wrt_from : &ident) : (base_wrt_to ? base_wrt_to : &ident);
				const POSE local_wrt_remote = POSE::compose(POSE::compose(*base_wrt_remote, pose_new_wrt_old), obs_wrt_local->inverse());
				ed.inv_pose = edge_dir_to_newkf ? local_wrt_remote.inverse() : local_wrt_remote;
			}
		}
		last_timestep_touched_kfs.clear();
		for (const auto &nei : neis) { last_timestep_touched_kfs.insert(k2k_edges[nei.id].from); last_timestep_touched_kfs.insert(k2k_edges[nei.id].to); }
	}

	// ------------------------------------------------------------ define_new_keyframe (impl/define_new_keyframe.h:16-116)
	void define_new_keyframe(const std::vector<new_kf_observation_t> &obs, TNewKeyFrameInfo &info, bool run_local_optimization = true) {
		info.kf_id = INVALID_ID;
		info.created_edge_ids.clear();
		std::memset(&info.optimize_results, 0, sizeof(srba_opt_info_t));
		const TKeyFrameID new_kf_id = alloc_keyframe();
		std::vector<TNewEdgeInfo> neis;
		determine_kf2kf_edges_to_create(new_kf_id, obs, neis);
		for (const auto &o : obs) add_observation(new_kf_id, o.feat_id, o.obs_data, o.is_fixed ? o.feat_rel_pos : nullptr, o.is_unknown_with_init_val ? o.feat_rel_pos : nullptr);
		if (run_local_optimization) {
			if (parameters.optimize_new_edges_alone && !neis.empty()) {
				const int old_kernel = parameters.use_robust_kernel;
				parameters.use_robust_kernel = parameters.use_robust_kernel_stage1;
				for (const auto &nei : neis) {
					if (nei.has_approx_init_val) continue;
					optimize_edges(std::vector<size_t>(1, nei.id), std::vector<size_t>(), info.optimize_results_stg1);
				}
				parameters.use_robust_kernel = old_kernel;
			}
			optimize_local_area(new_kf_id, parameters.max_optimize_depth, info.optimize_results, TOptimizeLocalAreaParams());
		}
		info.kf_id = new_kf_id;
		info.created_edge_ids.swap(neis);
	}

	// ------------------------------------------------------------ unknown selection: bfs_visitor (impl/bfs_visitor.h:21-177) with
	// VisitorOptimizeLocalArea (RbaEngine.h:543-618).  Integer work; order is the contract.
	void select_local_area(TKeyFrameID root_id, uint64_t win_size, const TOptimizeLocalAreaParams &prm, std::vector<size_t> &k2k_out, std::vector<size_t> &lm_out) const {
		const bool use_prebuilt_st = (win_size <= parameters.max_tree_depth);
		std::map<TLandmarkID, size_t> lm_times_seen;
		std::vector<uint8_t> k2k_visited(k2k_edges.size(), 0), k2f_visited(all_observations.size(), 0);
		auto visit_kf_edges = [&](TKeyFrameID kf, bool do_k2k) {
			for (uint32_t oi : keyframes[kf].adj_k2f) {
				if (k2f_visited[oi]) continue;
				k2f_visited[oi] = 1;
				if (prm.optimize_landmarks && !all_observations[oi].known_lm)
					if (++lm_times_seen[all_observations[oi].lm] == prm.dont_optimize_landmarks_seen_less_than_n_times) lm_out.push_back(all_observations[oi].lm);
			}
			(void)do_k2k;
		};
		if (!use_prebuilt_st) {
			std::set<TKeyFrameID> kf_visited;
			std::queue<TKeyFrameID> pending;
			std::map<TKeyFrameID, uint64_t> distances;
			pending.push(root_id); kf_visited.insert(root_id); distances[root_id] = 0;
			while (!pending.empty()) {
				const TKeyFrameID kf = pending.front();
				pending.pop();
				const uint64_t cur_dist = distances[kf];
				visit_kf_edges(kf, false);
				if (cur_dist >= win_size) continue;
				for (uint32_t e : keyframes[kf].adj_k2k) {
					const TKeyFrameID nk = other_end(kf, e);
					if (!kf_visited.count(nk)) {
						if (nk <= prm.max_visitable_kf_id) { pending.push(nk); distances[nk] = cur_dist + 1; }
						kf_visited.insert(nk);
					}
					if (!k2k_visited[e]) { k2k_visited[e] = 1; if (prm.optimize_k2k_edges) k2k_out.push_back(e); }
				}
			}
		} else {
			if (root_id >= next_edge.size()) return;
			std::vector<std::pair<TKeyFrameID, uint64_t> > KFs;
			KFs.push_back(std::make_pair(root_id, uint64_t(0)));
			for (const auto &kv : next_edge[root_id]) KFs.push_back(std::make_pair(kv.first, kv.second.distance));
			for (const auto &kd : KFs) {
				if (kd.second > win_size) continue;
				visit_kf_edges(kd.first, true);
				for (uint32_t e : keyframes[kd.first].adj_k2k)
					if (!k2k_visited[e]) { k2k_visited[e] = 1; if (prm.optimize_k2k_edges) k2k_out.push_back(e); }
			}
		}
	}
	// observation_indices_to_optimize (RbaEngine.h:235-241): empty = every observation of the selected unknowns
	void optimize_local_area(TKeyFrameID root_id, uint64_t win_size, srba_opt_info_t &out_info, const TOptimizeLocalAreaParams &prm,
	                         const std::vector<size_t> &observation_indices_to_optimize = std::vector<size_t>()) {
		std::vector<size_t> k2k, lms;
		select_local_area(root_id, win_size, prm, k2k, lms);
		if (!k2k.empty() || !lms.empty()) optimize_edges(k2k, lms, out_info, observation_indices_to_optimize);
	}

	// ------------------------------------------------------------ flatten: optimize_edges.h:71-312 + schur.h:25-159 as CSR
	std::unique_ptr<FlatProblem> flatten(const std::vector<size_t> &k2k_in, const std::vector<size_t> &lms_in,
	                                     const std::vector<size_t> &obs_indices_in = std::vector<size_t>()) {
		std::unique_ptr<FlatProblem> fp(new FlatProblem());
		FlatProblem &F = *fp;
		std::memset(&F.p, 0, sizeof(F.p));
		std::memset(&F.info, 0, sizeof(F.info));
		// A. filter unknowns with empty columns (:71-119)
		std::set<TKeyFrameID> touched_KFs;
		for (size_t e : k2k_in) {
			SRBA_CHECK(e < k2k_edges.size());
			if (colA[e].empty()) continue;
			F.k2k_ids.push_back(e);
			touched_KFs.insert(k2k_edges[e].from); touched_KFs.insert(k2k_edges[e].to);
		}
		std::set<TLandmarkID> touched_LMs;
		for (size_t id : lms_in) {
			SRBA_CHECK(id < all_lms.size() && all_lms[id].exists);
			SRBA_CHECK(!all_lms[id].known);
			if (colF[all_lms[id].fcol].empty()) continue;
			F.lm_ids.push_back(id);
			touched_LMs.insert(id);
		}
		const int n_p = F.k2k_ids.size(), n_f = F.lm_ids.size();
		F.info.num_kf_optimized = touched_KFs.size();
		F.info.num_lm_optimized = touched_LMs.size();
		if (P * n_p + 3 * n_f == 0) return nullptr;
		// id -> local index tables: stamped arrays kept per thread (no hashing, no per-call clearing; flatten() only READS the engine, so
		// several threads may flatten different areas of one engine at the same time -- srba_engine_flatten_local_areas)
		FlattenScratch &sc = FlattenScratch::get();
		sc.begin(all_observations.size(), all_lms.size(), k2k_edges.size(), keyframes.size());
		const uint32_t stamp = sc.stamp;

		// C. involved observations (:171-217): first-touch order, duplicates counted in obs_mult
		auto obs_known = [&](uint64_t g) -> bool { return sc.obs_stamp[g] == stamp; };
		auto obs_local = [&](uint64_t g) -> int {
			if (sc.obs_stamp[g] == stamp) return sc.obs_local[g];
			const int l = F.obs_global.size();
			sc.obs_stamp[g] = stamp; sc.obs_local[g] = l;
			F.obs_global.push_back(g);
			F.obs_mult.push_back(0);
			return l;
		};
		// in_observation_indices_to_optimize (:218-234): the residual rows are the caller's list (in its order, repeats counted);
		// compute_minus_gradient ASSERT_s that every block of an unknown column finds its observation in that list
		// (compute_minus_gradient.h:55-56,79-80), so the list must cover the columns; extra observations only add residual rows.
		const bool user_list = !obs_indices_in.empty();
		for (size_t g : obs_indices_in) {
			SRBA_CHECK(g < all_observations.size());
			F.obs_mult[obs_local(g)]++;
		}
		F.colA_ptr.push_back(0);
		for (int i = 0; i < n_p; i++) {
			for (uint32_t g : colA[F.k2k_ids[i]]) {
				if (user_list && !obs_known(g)) throw std::runtime_error("ASSERT_(it_obs!=obs_global_idx2residual_idx.end()): an observation of an unknown column is missing from observation_indices_to_optimize (compute_minus_gradient.h:55)");
				const int l = obs_local(g);
				if (!user_list) F.obs_mult[l]++;
				F.blkA_obs.push_back(l);
			}
			F.colA_ptr.push_back(F.blkA_obs.size());
		}
		F.colF_ptr.push_back(0);
		for (int i = 0; i < n_f; i++) {
			for (uint32_t g : colF[all_lms[F.lm_ids[i]].fcol]) {
				const bool fresh = !obs_known(g);
				if (user_list && fresh) throw std::runtime_error("ASSERT_(it_obs!=obs_global_idx2residual_idx.end()): an observation of an unknown column is missing from observation_indices_to_optimize (compute_minus_gradient.h:79)");
				const int l = obs_local(g);
				if (fresh) F.obs_mult[l] = 1;
				F.blkF_obs.push_back(l);
			}
			F.colF_ptr.push_back(F.blkF_obs.size());
		}
		const int n_obs = F.obs_global.size(), nA = F.blkA_obs.size(), nF = F.blkF_obs.size();
		int n_rows_dup = 0;
		for (int m : F.obs_mult) n_rows_dup += m;

		// D. needed spanning-tree roots (jacobians.h:1020-1076) and the pose-pair table of the touched rows
		std::vector<TKeyFrameID> roots;   // ascending, unique (the reference's std::set order)
		auto add_root = [&](TKeyFrameID kf) {
			if (sc.kf_stamp[kf] == stamp) return;
			sc.kf_stamp[kf] = stamp; sc.kf_first[kf] = -1;
			roots.push_back(kf);
		};
		std::vector<int> blkA_row(nA);  // index into rows[] of each dh_dAp block
		for (int i = 0; i < n_p; i++)
			for (int b = F.colA_ptr[i]; b < F.colA_ptr[i + 1]; b++) {
				const Observation &ob = all_observations[F.obs_global[F.blkA_obs[b]]];
				int r = -1;
				for (uint32_t k = 0; k < ob.row_len; k++) if (rows[ob.row_begin + k].edge == F.k2k_ids[i]) { r = ob.row_begin + k; break; }
				SRBA_CHECK(r >= 0);
				blkA_row[b] = r;
				add_root(rows[r].kf_d); add_root(ob.kf); add_root(all_lms[ob.lm].base);
			}
		for (int b = 0; b < nF; b++) {
			const Observation &ob = all_observations[F.obs_global[F.blkF_obs[b]]];
			add_root(all_lms[ob.lm].base); add_root(ob.kf);
		}
		std::sort(roots.begin(), roots.end());
		auto edge_local = [&](uint64_t e) -> int {
			if (sc.edge_stamp[e] == stamp) return sc.edge_local[e];
			const int l = F.edge_ids_touched.size();
			sc.edge_stamp[e] = stamp; sc.edge_local[e] = l;
			F.edge_ids_touched.push_back(e);
			return l;
		};
		for (int i = 0; i < n_p; i++) { edge_local(F.k2k_ids[i]); F.unk_edge.push_back(i); }
		F.pair_path_ptr.push_back(0);
		uint64_t span_updates = 0;
		for (TKeyFrameID from : roots) {
			if (from >= all_edges.size()) continue;
			span_updates += all_edges[from].size();
			sc.kf_first[from] = (int32_t)F.pairs.size();   // the pairs of one root are contiguous and ascending in `to`
			for (const auto &kv : all_edges[from]) {
				F.pairs.push_back(std::make_pair(from, kv.first));
				TKeyFrameID cur = from;
				for (uint32_t e : kv.second) {
					F.path_edge.push_back(edge_local(e));
					const bool fwd = (k2k_edges[e].to == cur);
					F.path_dir.push_back(fwd ? 1 : 0);
					cur = fwd ? k2k_edges[e].from : k2k_edges[e].to;
				}
				F.pair_path_ptr.push_back(F.path_edge.size());
			}
		}
		auto slot_of = [&](TKeyFrameID src, TKeyFrameID dst) -> int {  // slot of num[src][dst]
			const TKeyFrameID from = std::max(src, dst), to = std::min(src, dst);
			int idx = -1;
			if (sc.kf_stamp[from] == stamp && sc.kf_first[from] >= 0) {
				const auto first = F.pairs.begin() + sc.kf_first[from], last = first + all_edges[from].size();
				const auto it = std::lower_bound(first, last, std::make_pair(from, to));
				if (it != last && it->second == to) idx = (int)(it - F.pairs.begin());
			}
			// (reachable only through observation_indices_to_optimize: an EXTRA observation whose base pose the reference would read
			// from a stale, un-refreshed numeric spanning-tree entry -- not carried across the boundary)
			if (idx < 0) throw std::runtime_error("srba_b200: a listed observation needs a numeric spanning-tree entry outside the rows refreshed for the unknowns (unsupported)");
			return 2 * idx + (src > dst ? 0 : 1);
		};
		// landmarks touched
		auto lm_local = [&](uint64_t id) -> int {
			if (sc.lm_stamp[id] == stamp) return sc.lm_local[id];
			const int l = F.lm_ids_touched.size();
			sc.lm_stamp[id] = stamp; sc.lm_local[id] = l;
			F.lm_ids_touched.push_back(id);
			return l;
		};
		for (int i = 0; i < n_f; i++) F.unk_lm.push_back(lm_local(F.lm_ids[i]));
		// observations
		F.obs_z.resize((size_t)n_obs * O);
		F.obs_slot.resize(n_obs); F.obs_lm.resize(n_obs);
		for (int l = 0; l < n_obs; l++) {
			const Observation &ob = all_observations[F.obs_global[l]];
			for (int k = 0; k < O; k++) F.obs_z[(size_t)l * O + k] = ob.z[k];
			const TKeyFrameID base = all_lms[ob.lm].base;
			F.obs_slot[l] = (base == ob.kf) ? -1 : slot_of(ob.kf, base);
			F.obs_lm[l] = lm_local(ob.lm);
		}
		// blocks
		F.blkA_slotA.resize(nA); F.blkA_slotD.resize(nA); F.blkA_normal_dir.resize(nA);
		for (int b = 0; b < nA; b++) {
			const Observation &ob = all_observations[F.obs_global[F.blkA_obs[b]]];
			const RowEntry &re = rows[blkA_row[b]];
			F.blkA_normal_dir[b] = re.normal_dir;
			F.blkA_slotD[b] = slot_of(re.kf_d, all_lms[ob.lm].base);
			F.blkA_slotA[b] = (re.kf_d == ob.kf) ? -1 : slot_of(ob.kf, re.kf_d);
		}
		F.blkF_slot.resize(nF);
		for (int b = 0; b < nF; b++) {
			const Observation &ob = all_observations[F.obs_global[F.blkF_obs[b]]];
			F.blkF_slot[b] = ob.first_obs_of_unknown ? -1 : slot_of(ob.kf, all_lms[ob.lm].base);
		}
		// rows in walk order (dAp columns in unknown order, then the df block): a stable counting sort of the blocks by observation
		F.row_ptr.assign(n_obs + 1, 0);
		for (int b = 0; b < nA; b++) F.row_ptr[F.blkA_obs[b] + 1]++;
		for (int b = 0; b < nF; b++) F.row_ptr[F.blkF_obs[b] + 1]++;
		for (int l = 0; l < n_obs; l++) F.row_ptr[l + 1] += F.row_ptr[l];
		F.row_blk.resize(nA + nF);
		{
			std::vector<int32_t> fill(F.row_ptr.begin(), F.row_ptr.end() - 1);
			for (int b = 0; b < nA; b++) F.row_blk[fill[F.blkA_obs[b]]++] = b;
			for (int b = 0; b < nF; b++) F.row_blk[fill[F.blkF_obs[b]]++] = nA + b;
		}

		// I. Hessian symbolic (sparse_hessian_build_symbolic.h): terms of every block in ascending GLOBAL observation index.  The terms
		// are generated in that order as (block key, term) records and grouped by a stable two-pass counting sort on the key -- the
		// blocks come out in the order of the reference's std::map keys, the terms of a block in generation order.
		std::vector<int> obs_sorted(n_obs);
		for (int l = 0; l < n_obs; l++) obs_sorted[l] = l;
		std::sort(obs_sorted.begin(), obs_sorted.end(), [&](int a, int b) { return F.obs_global[a] < F.obs_global[b]; });
		std::vector<int> blkA_ucol(nA), blkF_fcol(nF);
		for (int i = 0; i < n_p; i++) for (int b = F.colA_ptr[i]; b < F.colA_ptr[i + 1]; b++) blkA_ucol[b] = i;
		for (int i = 0; i < n_f; i++) for (int b = F.colF_ptr[i]; b < F.colF_ptr[i + 1]; b++) blkF_fcol[b] = i;
		std::vector<TermRec> tA, tApf, tS;   // HAp: key (col,row) -> (blk_row, blk_col); HApf: key (row,fcol) -> (blkA, blkF)
		tA.reserve((size_t)nA * 2); tApf.reserve(nA);
		for (int l : obs_sorted) {
			const int r0 = F.row_ptr[l], r1 = F.row_ptr[l + 1];
			for (int x = r0; x < r1; x++) {
				const int bx = F.row_blk[x];
				if (bx >= nA) continue;
				const int ci = blkA_ucol[bx];
				for (int y = r0; y < r1; y++) {
					const int by = F.row_blk[y];
					if (by >= nA) { tApf.push_back(TermRec{ci, blkF_fcol[by - nA], bx, by - nA}); continue; }
					const int cj = blkA_ucol[by];
					if (ci <= cj && (ci < cj || x == y)) tA.push_back(TermRec{cj, ci, bx, by});
				}
			}
		}
		sort_terms(tA, n_p, n_p);
		F.HAp_term_ptr.push_back(0);
		for (size_t t = 0; t < tA.size(); t++) {
			if (t == 0 || tA[t].hi != tA[t - 1].hi || tA[t].lo != tA[t - 1].lo) {
				if (t) F.HAp_term_ptr.push_back(F.HAp_term_a.size());
				tS.push_back(TermRec{tA[t].hi, tA[t].lo, -1, (int32_t)F.HAp_row.size()});   // every H_Ap block is a block of the reduced system
				F.HAp_row.push_back(tA[t].lo); F.HAp_col.push_back(tA[t].hi);
			}
			F.HAp_term_a.push_back(tA[t].a); F.HAp_term_b.push_back(tA[t].b);
		}
		if (!tA.empty()) F.HAp_term_ptr.push_back(F.HAp_term_a.size());
		sort_terms(tApf, n_p, n_f);
		F.HApf_term_ptr.push_back(0);
		F.rowApf_ptr.assign(n_p + 1, 0);
		F.colApf_ptr.assign(n_f + 1, 0);
		for (size_t t = 0; t < tApf.size(); t++) {
			if (t == 0 || tApf[t].hi != tApf[t - 1].hi || tApf[t].lo != tApf[t - 1].lo) {
				if (t) F.HApf_term_ptr.push_back(F.HApf_term_a.size());
				F.HApf_row.push_back(tApf[t].hi); F.HApf_col.push_back(tApf[t].lo);
				F.rowApf_ptr[tApf[t].hi + 1]++;
				F.colApf_ptr[tApf[t].lo + 1]++;
			}
			F.HApf_term_a.push_back(tApf[t].a); F.HApf_term_f.push_back(tApf[t].b);
		}
		if (!tApf.empty()) F.HApf_term_ptr.push_back(F.HApf_term_a.size());
		for (int i = 0; i < n_p; i++) F.rowApf_ptr[i + 1] += F.rowApf_ptr[i];
		for (int k = 0; k < n_f; k++) F.colApf_ptr[k + 1] += F.colApf_ptr[k];
		F.colApf_idx.resize(F.HApf_row.size());   // blocks of each landmark column, ascending edge row
		{
			std::vector<int32_t> fill(F.colApf_ptr.begin(), F.colApf_ptr.end() - 1);
			for (size_t id = 0; id < F.HApf_col.size(); id++) F.colApf_idx[fill[F.HApf_col[id]]++] = (int32_t)id;
		}
		// O. Schur symbolic (schur.h:56-157): block (row j <= col i) gets one triplet per shared landmark, ascending k
		for (int k = 0; k < n_f; k++) {
			const int32_t *E = F.colApf_idx.data() + F.colApf_ptr[k];
			const int nE = F.colApf_ptr[k + 1] - F.colApf_ptr[k];
			for (int x = 0; x < nE; x++)
				for (int y = x; y < nE; y++) tS.push_back(TermRec{F.HApf_row[E[y]], F.HApf_row[E[x]], E[x], E[y]});   // key (col,row) -> (HApf(row,k), HApf(col,k))
		}
		sort_terms(tS, n_p, n_p);
		F.S_term_ptr.push_back(0);
		for (size_t t = 0; t < tS.size(); t++) {
			if (t == 0 || tS[t].hi != tS[t - 1].hi || tS[t].lo != tS[t - 1].lo) {
				if (t) F.S_term_ptr.push_back(F.S_term_rowblk.size());
				F.S_row.push_back(tS[t].lo); F.S_col.push_back(tS[t].hi); F.S_HAp.push_back(-1);
			}
			if (tS[t].a < 0) F.S_HAp.back() = tS[t].b;
			else { F.S_term_rowblk.push_back(tS[t].a); F.S_term_colblk.push_back(tS[t].b); }
		}
		if (!tS.empty()) F.S_term_ptr.push_back(F.S_term_rowblk.size());
		// values
		F.edge_pose.resize(F.edge_ids_touched.size() * S);
		for (size_t i = 0; i < F.edge_ids_touched.size(); i++) std::memcpy(&F.edge_pose[i * S], k2k_edges[F.edge_ids_touched[i]].inv_pose.v, S * sizeof(double));
		F.lm_pos.resize(F.lm_ids_touched.size() * 3);
		for (size_t i = 0; i < F.lm_ids_touched.size(); i++) std::memcpy(&F.lm_pos[i * 3], all_lms[F.lm_ids_touched[i]].pos, 3 * sizeof(double));
		// model + sizes
		srba_model_t &m = F.p.model;
		m.pose_kind = parameters.pose_kind; m.obs_kind = parameters.obs_kind; m.sensor_pose_kind = parameters.sensor_pose_kind; m.noise_kind = parameters.noise_kind;
		std::memcpy(m.sensor_pose, sensor_pose3.v, sizeof(m.sensor_pose));
		m.noise_std = parameters.noise_std;
		std::memcpy(m.noise_lambda, parameters.noise_lambda, sizeof(m.noise_lambda));
		std::memcpy(m.cam_left, parameters.cam_left, sizeof(m.cam_left));
		std::memcpy(m.cam_right, parameters.cam_right, sizeof(m.cam_right));
		std::memcpy(m.right_cam_R2L, right_cam_R2L.v, sizeof(m.right_cam_R2L));
		m.max_iters = (int32_t)std::min<uint64_t>(parameters.max_iters, 1u << 30);
		m.use_robust_kernel = parameters.use_robust_kernel;
		m.kernel_param = parameters.kernel_param;
		m.max_error_per_obs_to_stop = parameters.max_error_per_obs_to_stop;
		m.max_rho = parameters.max_rho; m.max_lambda = parameters.max_lambda;
		m.min_error_reduction_ratio_to_relinearize = parameters.min_error_reduction_ratio_to_relinearize;
		m.solver_kind = parameters.solver_kind;  // RBA_OPTIONS::solver_t (optimize_edges.h:54)
		m.reserved_ = 0;
		F.p.n_edges = F.edge_ids_touched.size(); F.p.n_unk_edges = n_p; F.p.n_lms = F.lm_ids_touched.size(); F.p.n_unk_lms = n_f;
		F.p.n_pairs = F.pairs.size(); F.p.n_path = F.path_edge.size(); F.p.n_obs = n_obs; F.p.n_blkA = nA; F.p.n_blkF = nF;
		F.p.n_HAp = F.HAp_row.size(); F.p.n_HAp_terms = F.HAp_term_a.size(); F.p.n_HApf = F.HApf_row.size(); F.p.n_HApf_terms = F.HApf_term_a.size();
		F.p.n_S = F.S_row.size(); F.p.n_S_terms = F.S_term_rowblk.size(); F.p.n_total_obs_rows = n_rows_dup;
		// unk_edge holds LOCAL edge indices: the unknown edges were registered first, so ucol i -> local edge i
		F.info.num_observations = n_rows_dup;
		F.info.num_jacobians = nA + (parameters.obs_kind == SRBA_OBS_RELPOSES2D ? 0 : nF);
		F.info.num_kf2kf_edges_optimized = n_p; F.info.num_kf2lm_edges_optimized = n_f;
		F.info.num_total_scalar_optimized = P * n_p + 3 * n_f;
		F.info.num_span_tree_numeric_updates = span_updates;
		F.bind();
		return fp;
	}

	// scatter a result back (optimize_edges.h:508-539,550-565,722-751)
	void apply(const FlatProblem &F, const srba_result_t &r, srba_opt_info_t &out) {
		for (size_t i = 0; i < F.k2k_ids.size(); i++) std::memcpy(k2k_edges[F.k2k_ids[i]].inv_pose.v, r.edge_pose + i * S, S * sizeof(double));
		for (size_t i = 0; i < F.lm_ids.size(); i++) std::memcpy(all_lms[F.lm_ids[i]].pos, r.lm_pos + i * 3, 3 * sizeof(double));
		for (size_t t = 0; t < F.pairs.size(); t++) {
			POSE a, b;
			std::memcpy(a.v, r.slot_pose + (2 * t) * S, S * sizeof(double));
			std::memcpy(b.v, r.slot_pose + (2 * t + 1) * S, S * sizeof(double));
			grow(num, F.pairs[t].first)[F.pairs[t].second] = a;
			grow(num, F.pairs[t].second)[F.pairs[t].first] = b;
		}
		for (size_t l = 0; l < F.obs_global.size(); l++) validity[F.obs_global[l]] = r.obs_valid[l];
		unknown_lms_inf_matrices.clear();
		if (parameters.cov_recovery == 1)
			for (size_t i = 0; i < F.lm_ids.size(); i++)
				if (r.lm_invertible[i]) { std::array<double, 9> M; std::memcpy(M.data(), r.Hf_diag + i * 9, sizeof(double) * 9); unknown_lms_inf_matrices[F.lm_ids[i]] = M; }
		out = F.info;
		out.obs_rmse = r.stats.obs_rmse;
		out.total_sqr_error_init = r.stats.total_sqr_error_init;
		out.total_sqr_error_final = r.stats.total_sqr_error_final;
		out.num_lm_trials = r.stats.num_trials;
		out.num_iters = r.stats.num_iters;
	}

	// optimize_edges (impl/optimize_edges.h:44-793) == flatten -> srba_b200_optimize -> scatter
	void optimize_edges(const std::vector<size_t> &k2k_in, const std::vector<size_t> &lms_in, srba_opt_info_t &out_info,
	                    const std::vector<size_t> &in_observation_indices_to_optimize = std::vector<size_t>()) {
		std::memset(&out_info, 0, sizeof(out_info));
		const auto t0 = std::chrono::steady_clock::now();
		std::unique_ptr<FlatProblem> fp = flatten(k2k_in, lms_in, in_observation_indices_to_optimize);
		if (!fp) return;  // "Skipping optimization since no observation depends on any of the given variables" (:132-138)
		FlatProblem &F = *fp;
		const int S_ = S;
		std::vector<double> edge_out(F.k2k_ids.size() * S_), lm_out(F.lm_ids.size() * 3 + 1), slot_out(2 * F.pairs.size() * S_ + 1), Hf(F.lm_ids.size() * 9 + 1);
		std::vector<uint8_t> valid(F.obs_global.size()), inv(F.lm_ids.size() + 1);
		srba_result_t r;
		std::memset(&r, 0, sizeof(r));
		r.edge_pose = edge_out.data(); r.lm_pos = lm_out.data(); r.slot_pose = slot_out.data(); r.obs_valid = valid.data();
		r.lm_invertible = inv.data(); r.Hf_diag = Hf.data();
		const uint64_t nsc = F.info.num_total_scalar_optimized;
		if (trace_cap > 0) {
			last_trace.assign((size_t)trace_cap * 6, 0.0);
			last_trace_delta.assign((size_t)trace_cap * nsc, 0.0);
			r.trace_cap = trace_cap; r.trace = last_trace.data(); r.trace_delta = last_trace_delta.data();
			last_trace_scalars = nsc;
		}
		const auto t1 = std::chrono::steady_clock::now();
		const int rc = srba_b200_optimize(&F.p, &r);
		const auto t2 = std::chrono::steady_clock::now();
		if (rc != SRBA_OK) throw std::runtime_error(std::string("srba_b200_optimize failed: ") + srba_b200_last_error());
		if (trace_cap > 0) { last_trace.resize((size_t)r.trace_len * 6); last_trace_delta.resize((size_t)r.trace_len * nsc); }
		apply(F, r, out_info);
		total_trials += r.stats.num_trials;
		const auto t3 = std::chrono::steady_clock::now();
		flatten_seconds += std::chrono::duration<double>(t1 - t0).count() + std::chrono::duration<double>(t3 - t2).count();
		gpu_seconds += std::chrono::duration<double>(t2 - t1).count();
		if (record_problems) recorded.push_back(std::move(fp));
	}
};

}  // namespace srba_b200
