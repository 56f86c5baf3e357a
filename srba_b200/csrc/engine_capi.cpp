// C view of srba_b200::RbaEngine (include/srba_b200_engine.h).
#include "host_engine.hpp"

#include <atomic>
#include <mutex>
#include <thread>

using namespace srba_b200;

namespace {
thread_local std::string g_eng_err;

struct Handle {
	int pose_kind;
	RbaEngine<SE2> *e2 = nullptr;
	RbaEngine<SE3> *e3 = nullptr;
	EngineBase *base() { return e2 ? static_cast<EngineBase *>(e2) : static_cast<EngineBase *>(e3); }
};
struct FlatHandle { std::unique_ptr<FlatProblem> fp; };

#define DISPATCH(h, ...)                                                            \
	do {                                                                            \
		Handle *H_ = static_cast<Handle *>(h);                                      \
		if (H_->e2) { RbaEngine<SE2> &E = *H_->e2; __VA_ARGS__; }                   \
		else { RbaEngine<SE3> &E = *H_->e3; __VA_ARGS__; }                          \
	} while (0)
#define GUARD_BEGIN try {
#define GUARD_END(ret) } catch (const std::exception &e) { g_eng_err = e.what(); return ret; }

template <class E>
void fill_obs(const E &eng, std::vector<typename E::new_kf_observation_t> &lst, uint64_t n, const uint64_t *feat_id, const double *obs_data,
              const uint8_t *is_fixed, const uint8_t *is_unk, const double *frp) {
	const int O = eng.O;
	const bool pixels = (eng.parameters.obs_kind == SRBA_OBS_MONO || eng.parameters.obs_kind == SRBA_OBS_STEREO);
	lst.resize(n);
	for (uint64_t i = 0; i < n; i++) {
		auto &o = lst[i];
		o.feat_id = feat_id[i];
		for (int k = 0; k < 4; k++) o.obs_data[k] = 0;
		// pixel observations are stored as float in the reference (TPixelCoordf, observations_MonocularCamera.h:28,33-34)
		for (int k = 0; k < O; k++) o.obs_data[k] = pixels ? (double)(float)obs_data[i * O + k] : obs_data[i * O + k];
		o.is_fixed = is_fixed ? is_fixed[i] != 0 : false;
		o.is_unknown_with_init_val = is_unk ? is_unk[i] != 0 : false;
		for (int k = 0; k < 3; k++) o.feat_rel_pos[k] = frp ? frp[i * 3 + k] : 0.0;
	}
}
}  // namespace

extern "C" {

const char *srba_engine_last_error(void) { return g_eng_err.c_str(); }

void srba_engine_default_params(srba_engine_params_t *p) {
	std::memset(p, 0, sizeof(*p));
	p->pose_kind = SRBA_POSE_SE3; p->obs_kind = SRBA_OBS_CARTESIAN3D; p->sensor_pose_kind = SRBA_SENSORPOSE_NONE; p->noise_kind = SRBA_NOISE_IDENTITY;
	p->noise_std = 1.0;
	for (int i = 0; i < 4; i++) p->noise_lambda[i * 4 + i] = 1.0;
	p->right_cam_pose[3] = 1.0;
	p->max_tree_depth = 4; p->max_optimize_depth = 4;  // impl/rba_problem_common.h:35-56
	p->optimize_new_edges_alone = 1; p->use_robust_kernel = 0; p->use_robust_kernel_stage1 = 0; p->kernel_param = 3.; p->max_iters = 20;
	p->max_error_per_obs_to_stop = 1e-6; p->max_rho = 10.0; p->max_lambda = 1e20; p->min_error_reduction_ratio_to_relinearize = 0.01;
	p->cov_recovery = 1;
	p->submap_size = 15; p->min_obs_to_loop_closure = 4;  // ecps/local_areas_fixed_size.h:30-33
}

void *srba_engine_create(const srba_engine_params_t *p) {
	GUARD_BEGIN
	Handle *h = new Handle();
	h->pose_kind = p->pose_kind;
	if (p->pose_kind == SRBA_POSE_SE2) h->e2 = new RbaEngine<SE2>(*p);
	else h->e3 = new RbaEngine<SE3>(*p);
	return h;
	GUARD_END(nullptr)
}
void srba_engine_destroy(void *h) {
	Handle *H = static_cast<Handle *>(h);
	if (!H) return;
	delete H->e2; delete H->e3; delete H;
}
int srba_engine_set_params(void *h, const srba_engine_params_t *p) {  // validated BEFORE anything is assigned (host_engine.hpp set_params)
	GUARD_BEGIN
	if (!h || !p) throw std::runtime_error("srba_engine_set_params: null argument");
	DISPATCH(h, E.set_params(*p));
	return 0;
	GUARD_END(-1)
}

int srba_engine_define_new_keyframe(void *h, uint64_t n_obs, const uint64_t *feat_id, const double *obs_data, const uint8_t *is_fixed,
                                    const uint8_t *is_unk, const double *frp, int run_local_optimization, uint64_t *out_kf_id,
                                    srba_new_edge_t *out_edges, uint64_t max_edges, uint64_t *out_n_edges, srba_opt_info_t *out_opt, srba_opt_info_t *out_stg1) {
	GUARD_BEGIN
	DISPATCH(h, {
		typedef typename std::remove_reference<decltype(E)>::type EngT;
		std::vector<typename EngT::new_kf_observation_t> lst;
		fill_obs(E, lst, n_obs, feat_id, obs_data, is_fixed, is_unk, frp);
		typename EngT::TNewKeyFrameInfo info;
		std::memset(&info.optimize_results, 0, sizeof(srba_opt_info_t));
		std::memset(&info.optimize_results_stg1, 0, sizeof(srba_opt_info_t));
		E.define_new_keyframe(lst, info, run_local_optimization != 0);
		if (out_kf_id) *out_kf_id = info.kf_id;
		if (out_n_edges) *out_n_edges = info.created_edge_ids.size();
		for (size_t i = 0; i < info.created_edge_ids.size() && i < max_edges; i++) {
			out_edges[i].id = info.created_edge_ids[i].id;
			out_edges[i].has_approx_init_val = info.created_edge_ids[i].has_approx_init_val;
			out_edges[i].loopclosure_observer_kf = info.created_edge_ids[i].loopclosure_observer_kf;
			out_edges[i].loopclosure_base_kf = info.created_edge_ids[i].loopclosure_base_kf;
		}
		if (out_opt) *out_opt = info.optimize_results;
		if (out_stg1) *out_stg1 = info.optimize_results_stg1;
	});
	return 0;
	GUARD_END(-1)
}

int srba_engine_optimize_local_area(void *h, uint64_t root, uint64_t win_size, int opt_k2k, int opt_lms, uint64_t max_visitable, uint64_t min_seen, srba_opt_info_t *out) {
	GUARD_BEGIN
	DISPATCH(h, {
		typedef typename std::remove_reference<decltype(E)>::type EngT;
		typename EngT::TOptimizeLocalAreaParams prm;
		prm.optimize_k2k_edges = opt_k2k != 0; prm.optimize_landmarks = opt_lms != 0;
		prm.max_visitable_kf_id = max_visitable; prm.dont_optimize_landmarks_seen_less_than_n_times = min_seen;
		srba_opt_info_t info;
		std::memset(&info, 0, sizeof(info));
		E.optimize_local_area(root, win_size, info, prm);
		if (out) *out = info;
	});
	return 0;
	GUARD_END(-1)
}
int srba_engine_optimize_local_area_obs(void *h, uint64_t root, uint64_t win_size, uint64_t n_obs_idx, const uint64_t *obs_idx, srba_opt_info_t *out) {
	GUARD_BEGIN
	if (!h || (n_obs_idx && !obs_idx)) throw std::runtime_error("srba_engine_optimize_local_area_obs: null argument");
	DISPATCH(h, {
		typedef typename std::remove_reference<decltype(E)>::type EngT;
		std::vector<size_t> lst(obs_idx, obs_idx + n_obs_idx);
		srba_opt_info_t info;
		std::memset(&info, 0, sizeof(info));
		E.optimize_local_area(root, win_size, info, typename EngT::TOptimizeLocalAreaParams(), lst);
		if (out) *out = info;
	});
	return 0;
	GUARD_END(-1)
}
int srba_engine_optimize_edges(void *h, uint64_t n_k2k, const uint64_t *k2k, uint64_t n_lm, const uint64_t *lms, srba_opt_info_t *out) {
	GUARD_BEGIN
	DISPATCH(h, {
		std::vector<size_t> a(k2k, k2k + n_k2k), b(lms, lms + n_lm);
		srba_opt_info_t info;
		std::memset(&info, 0, sizeof(info));
		E.optimize_edges(a, b, info);
		if (out) *out = info;
	});
	return 0;
	GUARD_END(-1)
}
int srba_engine_select_local_area(void *h, uint64_t root, uint64_t win_size, uint64_t *k2k_ids, uint64_t k2k_cap, uint64_t *n_k2k, uint64_t *lm_ids, uint64_t lm_cap, uint64_t *n_lm) {
	GUARD_BEGIN
	DISPATCH(h, {
		typedef typename std::remove_reference<decltype(E)>::type EngT;
		std::vector<size_t> a, b;
		E.select_local_area(root, win_size, typename EngT::TOptimizeLocalAreaParams(), a, b);
		*n_k2k = a.size(); *n_lm = b.size();
		for (size_t i = 0; i < a.size() && i < k2k_cap; i++) k2k_ids[i] = a[i];
		for (size_t i = 0; i < b.size() && i < lm_cap; i++) lm_ids[i] = b[i];
	});
	return 0;
	GUARD_END(-1)
}

void *srba_engine_flatten(void *h, uint64_t n_k2k, const uint64_t *k2k, uint64_t n_lm, const uint64_t *lms) {
	GUARD_BEGIN
	std::unique_ptr<FlatProblem> fp;
	DISPATCH(h, {
		std::vector<size_t> a(k2k, k2k + n_k2k), b(lms, lms + n_lm);
		fp = E.flatten(a, b);
	});
	if (!fp) return nullptr;
	FlatHandle *fh = new FlatHandle();
	fh->fp = std::move(fp);
	return fh;
	GUARD_END(nullptr)
}
void *srba_engine_flatten_local_area(void *h, uint64_t root, uint64_t win_size) {
	GUARD_BEGIN
	std::unique_ptr<FlatProblem> fp;
	DISPATCH(h, {
		typedef typename std::remove_reference<decltype(E)>::type EngT;
		std::vector<size_t> a, b;
		E.select_local_area(root, win_size, typename EngT::TOptimizeLocalAreaParams(), a, b);
		if (!a.empty() || !b.empty()) fp = E.flatten(a, b);
	});
	if (!fp) return nullptr;
	FlatHandle *fh = new FlatHandle();
	fh->fp = std::move(fp);
	return fh;
	GUARD_END(nullptr)
}
void *srba_engine_flatten_local_area_obs(void *h, uint64_t root, uint64_t win_size, uint64_t n_obs_idx, const uint64_t *obs_idx) {
	GUARD_BEGIN
	std::unique_ptr<FlatProblem> fp;
	DISPATCH(h, {
		typedef typename std::remove_reference<decltype(E)>::type EngT;
		std::vector<size_t> a, b, lst(obs_idx, obs_idx + n_obs_idx);
		E.select_local_area(root, win_size, typename EngT::TOptimizeLocalAreaParams(), a, b);
		if (!a.empty() || !b.empty()) fp = E.flatten(a, b, lst);
	});
	if (!fp) return nullptr;
	FlatHandle *fh = new FlatHandle();
	fh->fp = std::move(fp);
	return fh;
	GUARD_END(nullptr)
}
// Snapshot semantics (the local areas of many keyframes of ONE map state as independent problems): flatten() and
// select_local_area() only read the engine, so the areas are flattened by several host threads at once.
int srba_engine_flatten_local_areas(void *h, uint64_t n, const uint64_t *roots, uint64_t win_size, int n_threads, void **out) {
	GUARD_BEGIN
	if (!h || (n && (!roots || !out))) throw std::runtime_error("srba_engine_flatten_local_areas: null argument");
	for (uint64_t i = 0; i < n; i++) out[i] = nullptr;
	unsigned nt = n_threads > 0 ? (unsigned)n_threads : std::max(1u, std::thread::hardware_concurrency());
	nt = (unsigned)std::min<uint64_t>(nt, std::max<uint64_t>(n, 1));
	std::atomic<uint64_t> next(0);
	std::mutex err_mutex;
	std::string err;
	auto work = [&]() {
		for (;;) {
			const uint64_t i = next.fetch_add(1);
			if (i >= n) return;
			try {
				std::unique_ptr<FlatProblem> fp;
				DISPATCH(h, {
					typedef typename std::remove_reference<decltype(E)>::type EngT;
					std::vector<size_t> a, b;
					E.select_local_area(roots[i], win_size, typename EngT::TOptimizeLocalAreaParams(), a, b);
					if (!a.empty() || !b.empty()) fp = E.flatten(a, b);
				});
				if (fp) { FlatHandle *fh = new FlatHandle(); fh->fp = std::move(fp); out[i] = fh; }
			} catch (const std::exception &e) {
				std::lock_guard<std::mutex> lk(err_mutex);
				if (err.empty()) err = e.what();
				next.store(n);   // stop handing out work
				return;
			}
		}
	};
	std::vector<std::thread> pool;
	for (unsigned t = 1; t < nt; t++) pool.emplace_back(work);
	work();
	for (std::thread &t : pool) t.join();
	if (!err.empty()) {
		for (uint64_t i = 0; i < n; i++) { delete static_cast<FlatHandle *>(out[i]); out[i] = nullptr; }
		throw std::runtime_error(err);
	}
	return 0;
	GUARD_END(-1)
}
const srba_problem_t *srba_flat_problem(const void *flat) { return flat ? &static_cast<const FlatHandle *>(flat)->fp->p : nullptr; }
int srba_flat_apply(void *h, const void *flat, const srba_result_t *res, srba_opt_info_t *out) {
	GUARD_BEGIN
	srba_opt_info_t info;
	std::memset(&info, 0, sizeof(info));
	if (!h || !flat || !res) throw std::runtime_error("srba_flat_apply: null argument");
	DISPATCH(h, E.apply(*static_cast<const FlatHandle *>(flat)->fp, *res, info));
	if (out) *out = info;
	return 0;
	GUARD_END(-1)
}
void srba_flat_ids(const void *flat, uint64_t *k2k_ids, uint64_t *lm_ids, uint64_t *obs_global_idx, uint64_t *pair_from_to) {
	const FlatProblem &F = *static_cast<const FlatHandle *>(flat)->fp;
	if (k2k_ids) std::copy(F.k2k_ids.begin(), F.k2k_ids.end(), k2k_ids);
	if (lm_ids) std::copy(F.lm_ids.begin(), F.lm_ids.end(), lm_ids);
	if (obs_global_idx) std::copy(F.obs_global.begin(), F.obs_global.end(), obs_global_idx);
	if (pair_from_to) for (size_t t = 0; t < F.pairs.size(); t++) { pair_from_to[2 * t] = F.pairs[t].first; pair_from_to[2 * t + 1] = F.pairs[t].second; }
}
uint64_t srba_flat_touched_edges(const void *flat, uint64_t *edge_ids, uint64_t cap) {
	const FlatProblem &F = *static_cast<const FlatHandle *>(flat)->fp;
	if (edge_ids) for (size_t i = 0; i < F.edge_ids_touched.size() && i < cap; i++) edge_ids[i] = F.edge_ids_touched[i];
	return F.edge_ids_touched.size();
}
void srba_flat_destroy(void *flat) { delete static_cast<FlatHandle *>(flat); }

void srba_engine_record_problems(void *h, int on) { static_cast<Handle *>(h)->base()->record_problems = on != 0; }
uint64_t srba_engine_num_recorded(void *h) { return static_cast<Handle *>(h)->base()->recorded.size(); }
void *srba_engine_take_recorded(void *h, uint64_t i) {
	EngineBase *b = static_cast<Handle *>(h)->base();
	if (i >= b->recorded.size() || !b->recorded[i]) return nullptr;
	FlatHandle *fh = new FlatHandle();
	fh->fp = std::move(b->recorded[i]);
	// pop semantics: once every slot has been taken the list is emptied, so it does not grow (and get rescanned) forever
	bool all_taken = true;
	for (const auto &r : b->recorded) if (r) { all_taken = false; break; }
	if (all_taken) b->recorded.clear();
	return fh;
}
void srba_engine_trace(void *h, int trace_cap) { static_cast<Handle *>(h)->base()->trace_cap = trace_cap; }
uint64_t srba_engine_last_trace(void *h, double *trace6, double *delta, uint64_t cap_trials, uint64_t *n_scalars) {
	EngineBase *b = static_cast<Handle *>(h)->base();
	const uint64_t n = std::min<uint64_t>(b->last_trace.size() / 6, cap_trials);
	if (n_scalars) *n_scalars = b->last_trace_scalars;
	if (trace6) std::copy(b->last_trace.begin(), b->last_trace.begin() + n * 6, trace6);
	if (delta) std::copy(b->last_trace_delta.begin(), b->last_trace_delta.begin() + n * b->last_trace_scalars, delta);
	return b->last_trace.size() / 6;
}

uint64_t srba_engine_num_keyframes(void *h) { uint64_t r = 0; DISPATCH(h, r = E.keyframes.size()); return r; }
uint64_t srba_engine_num_edges(void *h) { uint64_t r = 0; DISPATCH(h, r = E.k2k_edges.size()); return r; }
uint64_t srba_engine_num_observations(void *h) { uint64_t r = 0; DISPATCH(h, r = E.all_observations.size()); return r; }
int srba_engine_pose_nvals(void *h) { return static_cast<Handle *>(h)->e2 ? 3 : 12; }
void srba_engine_get_edge(void *h, uint64_t id, uint64_t *from, uint64_t *to, double *inv_pose) {
	DISPATCH(h, {
		if (id >= E.k2k_edges.size()) { if (from) *from = srba_b200::INVALID_ID; if (to) *to = srba_b200::INVALID_ID; break; }
		if (from) *from = E.k2k_edges[id].from;
		if (to) *to = E.k2k_edges[id].to;
		if (inv_pose) std::memcpy(inv_pose, E.k2k_edges[id].inv_pose.v, sizeof(E.k2k_edges[id].inv_pose.v));
	});
}
void srba_engine_set_edge_pose(void *h, uint64_t id, const double *inv_pose) {
	DISPATCH(h, if (inv_pose && id < E.k2k_edges.size()) std::memcpy(E.k2k_edges[id].inv_pose.v, inv_pose, sizeof(E.k2k_edges[id].inv_pose.v)));
}
uint64_t srba_engine_num_unknown_lms(void *h) {
	uint64_t r = 0;
	DISPATCH(h, for (const auto &L : E.all_lms) r += (L.exists && !L.known) ? 1 : 0);
	return r;
}
void srba_engine_get_unknown_lms(void *h, uint64_t *ids, uint64_t *base, double *pos) {
	DISPATCH(h, {
		size_t k = 0;
		for (size_t id = 0; id < E.all_lms.size(); id++) {
			const auto &L = E.all_lms[id];
			if (!L.exists || L.known) continue;
			if (ids) ids[k] = id;
			if (base) base[k] = L.base;
			if (pos) for (int j = 0; j < 3; j++) pos[3 * k + j] = L.pos[j];
			k++;
		}
	});
}
void srba_engine_set_lm_pos(void *h, uint64_t id, const double *pos) {
	DISPATCH(h, if (id < E.all_lms.size() && E.all_lms[id].exists) for (int j = 0; j < 3; j++) E.all_lms[id].pos[j] = pos[j]);
}
uint64_t srba_engine_get_next_edge(void *h, uint64_t kf, uint64_t *target, uint64_t *next, uint64_t *dist, uint64_t cap) {
	uint64_t n = 0;
	DISPATCH(h, {
		if (kf < E.next_edge.size()) {
			n = E.next_edge[kf].size();
			uint64_t k = 0;
			for (auto it = E.next_edge[kf].begin(); it != E.next_edge[kf].end() && k < cap; ++it, ++k) {
				if (target) target[k] = it->first;
				if (next) next[k] = it->second.next;
				if (dist) dist[k] = it->second.distance;
			}
		}
	});
	return n;
}
uint64_t srba_engine_get_jacob_validity(void *h, uint8_t *out, uint64_t cap) {
	uint64_t n = 0;
	DISPATCH(h, {
		n = E.validity.size();
		for (uint64_t i = 0; out && i < std::min<uint64_t>(n, cap); i++) out[i] = E.validity[i];
	});
	return n;
}
int srba_engine_eval_overall_squared_error(void *h, double *out_sqerr) {
	GUARD_BEGIN
	DISPATCH(h, *out_sqerr = E.eval_overall_squared_error());
	return 0;
	GUARD_END(-1)
}
int64_t srba_engine_complete_spanning_tree(void *h, uint64_t root, uint64_t max_depth, uint64_t *ids, double *poses, uint64_t cap) {
	GUARD_BEGIN
	int64_t n = 0;
	DISPATCH(h, {
		typedef typename std::remove_reference<decltype(E)>::type EngT;
		std::vector<TKeyFrameID> id_;
		std::vector<double> pose_;
		if (!ids && !poses) {  // size query: the integer BFS only
			std::map<TKeyFrameID, typename EngT::BfsNode> prec;
			E.complete_spanning_tree_bfs(root, max_depth, prec);
			n = (int64_t)prec.size();
		} else {
			E.create_complete_spanning_tree(root, max_depth, id_, pose_);
			n = (int64_t)id_.size();
			for (uint64_t k = 0; k < std::min<uint64_t>(cap, id_.size()); k++) {
				if (ids) ids[k] = id_[k];
				if (poses) std::memcpy(poses + k * EngT::S, &pose_[k * EngT::S], sizeof(double) * EngT::S);
			}
		}
	});
	return n;
	GUARD_END(-1)
}
uint64_t srba_engine_get_num(void *h, uint64_t kf, uint64_t *target, double *poses, uint64_t cap) {
	uint64_t n = 0;
	DISPATCH(h, {
		typedef typename std::remove_reference<decltype(E)>::type EngT;
		if (kf < E.num.size()) {
			n = E.num[kf].size();
			uint64_t k = 0;
			for (auto it = E.num[kf].begin(); it != E.num[kf].end() && k < cap; ++it, ++k) {
				if (target) target[k] = it->first;
				if (poses) std::memcpy(poses + k * EngT::S, it->second.v, sizeof(double) * EngT::S);
			}
		}
	});
	return n;
}
void srba_engine_set_num(void *h, uint64_t src, uint64_t dst, const double *pose) {
	DISPATCH(h, {
		typedef typename std::remove_reference<decltype(E)>::type EngT;
		if (pose && src < E.keyframes.size() && dst < E.keyframes.size()) std::memcpy(EngT::grow(E.num, src)[dst].v, pose, sizeof(double) * EngT::S);
	});
}
int64_t srba_engine_get_all_edges_path(void *h, uint64_t from, uint64_t to, uint64_t *edge_ids, uint64_t cap) {
	int64_t n = -1;
	DISPATCH(h, {
		if (from < E.all_edges.size()) {
			auto it = E.all_edges[from].find(to);
			if (it != E.all_edges[from].end()) {
				n = it->second.size();
				for (size_t k = 0; k < it->second.size() && k < cap; k++) edge_ids[k] = it->second[k];
			}
		}
	});
	return n;
}
uint64_t srba_engine_alloc_keyframe(void *h) { uint64_t r = 0; DISPATCH(h, r = E.alloc_keyframe()); return r; }
uint64_t srba_engine_create_kf2kf_edge(void *h, uint64_t new_kf_id, uint64_t from, uint64_t to, const double *init_inv_pose) {
	GUARD_BEGIN
	uint64_t r = 0;
	DISPATCH(h, {
		typedef typename std::remove_reference<decltype(E)>::type EngT;
		typename std::remove_reference<decltype(E.k2k_edges[0].inv_pose)>::type p;
		if (init_inv_pose) std::memcpy(p.v, init_inv_pose, sizeof(double) * EngT::S);
		r = E.create_kf2kf_edge(new_kf_id, from, to, p);
	});
	return r;
	GUARD_END(static_cast<uint64_t>(-1))
}
double srba_engine_flatten_seconds(void *h) { return static_cast<Handle *>(h)->base()->flatten_seconds; }
double srba_engine_gpu_seconds(void *h) { return static_cast<Handle *>(h)->base()->gpu_seconds; }
uint64_t srba_engine_total_trials(void *h) { return static_cast<Handle *>(h)->base()->total_trials; }

}  // extern "C"
