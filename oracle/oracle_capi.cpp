// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).  C API over srba_oracle.hpp for ctypes.
#include "oracle_capi.h"

#include "srba_oracle.hpp"

using namespace srba_oracle;

namespace {
thread_local std::string g_last_error;

struct Handle {
	int pose_kind;
	Engine<Pose2> *e2;
	Engine<Pose3> *e3;
};

#define DISPATCH(h, ...)                                   \
	do {                                                   \
		Handle *H_ = static_cast<Handle *>(h);             \
		if (H_->pose_kind == ORC_POSE_SE2) { Engine<Pose2> &E = *H_->e2; __VA_ARGS__; } \
		else { Engine<Pose3> &E = *H_->e3; __VA_ARGS__; }  \
	} while (0)

#define GUARD_BEGIN try {
#define GUARD_END(ret_err)                     \
	}                                          \
	catch (const std::exception &e) {          \
		g_last_error = e.what();               \
		return ret_err;                        \
	}

template <class E>
void fill_obs(const E &eng, typename E::new_kf_observations_t &lst, uint64_t n_obs, const uint64_t *feat_id, const double *obs_data,
              const uint8_t *is_fixed, const uint8_t *is_unk_init, const double *feat_rel_pos) {
	const int O = eng.O;
	const bool pixels = (eng.parameters.obs_kind == ORC_OBS_MONO || eng.parameters.obs_kind == ORC_OBS_STEREO);
	for (uint64_t i = 0; i < n_obs; i++) {
		typename E::new_kf_observation_t o;
		o.feat_id = feat_id[i];
		for (int k = 0; k < MAX_OBS; k++) o.obs_data[k] = 0;
		// TPixelCoordf stores pixels as float (observations_MonocularCamera.h:28,33-34)
		for (int k = 0; k < O; k++) o.obs_data[k] = pixels ? static_cast<double>(static_cast<float>(obs_data[i * O + k])) : obs_data[i * O + k];
		o.is_fixed = is_fixed ? is_fixed[i] != 0 : false;
		o.is_unknown_with_init_val = is_unk_init ? is_unk_init[i] != 0 : false;
		for (int k = 0; k < LM_DIMS; k++) o.feat_rel_pos[k] = feat_rel_pos ? feat_rel_pos[i * LM_DIMS + k] : 0.0;
		lst.push_back(o);
	}
}
}  // namespace

namespace {
template <class POSE>
int eval_dh_dAp_t(const orc_params_t *p, const double *A, const double *D, const double *edge_inv_pose, int normal_dir, const double *lm_pos, double *out_J) {
	Engine<POSE> E(*p);
	typename Engine<POSE>::pose_flag_t fa, fd;
	if (A) { fa.pose.fromFlat(A); fa.updated = true; }
	fd.pose.fromFlat(D); fd.updated = true;
	typename Engine<POSE>::TRelativeLandmarkPos rfp;
	rfp.id_frame_base = 0;
	for (int k = 0; k < 3; k++) rfp.pos[k] = lm_pos[k];
	E.k2k_edges.push_back(typename Engine<POSE>::k2k_edge_t());
	if (edge_inv_pose) E.k2k_edges[0].inv_pose.fromFlat(edge_inv_pose);
	char valid = 1;
	typename Engine<POSE>::EntryAp ent;
	ent.sym.rel_pose_d1_from_obs = A ? &fa : NULL;
	ent.sym.rel_pose_base_from_d1 = &fd;
	ent.sym.edge_normal_dir = normal_dir != 0;
	ent.sym.feat_rel_pos = &rfp;
	ent.sym.kf_d = 1; ent.sym.kf_base = 0; ent.sym.k2k_edge_id = 0; ent.sym.obs_idx = 0; ent.sym.is_valid = &valid;
	typename Engine<POSE>::k2f_edge_t obs;
	obs.kf_id = 2;
	E.compute_jacobian_dh_dp(ent, obs, NULL);
	std::copy(ent.num, ent.num + E.O * POSE::DIMS, out_J);
	return valid;
}
}  // namespace

extern "C" {

const char *orc_last_error(void) { return g_last_error.c_str(); }

void orc_default_params(orc_params_t *p) {
	std::memset(p, 0, sizeof(*p));
	p->pose_kind = ORC_POSE_SE3;
	p->obs_kind = ORC_OBS_CARTESIAN3D;
	p->sensor_pose_kind = ORC_SENSORPOSE_NONE;   // RBA_OPTIONS_DEFAULT (RbaEngine.h:39-45)
	p->noise_kind = ORC_NOISE_IDENTITY;
	p->noise_std = 1.0;
	for (int i = 0; i < 4; i++) p->noise_lambda[i * 4 + i] = 1.0;
	p->right_cam_pose[3] = 1.0;
	// impl/rba_problem_common.h:35-56
	p->max_tree_depth = 4;
	p->max_optimize_depth = 4;
	p->optimize_new_edges_alone = 1;
	p->use_robust_kernel = 0;
	p->use_robust_kernel_stage1 = 0;
	p->kernel_param = 3.;
	p->max_iters = 20;
	p->max_error_per_obs_to_stop = 1e-6;
	p->max_rho = 10.0;
	p->max_lambda = 1e20;
	p->min_error_reduction_ratio_to_relinearize = 0.01;
	p->cov_recovery = 1;  // crpLandmarksApprox
	// ecps/local_areas_fixed_size.h:30-33
	p->submap_size = 15;
	p->min_obs_to_loop_closure = 4;
	p->solver_kind = 0;  // solver_LM_schur_dense_cholesky (RbaEngine.h:44)
	p->reserved_ = 0;
}

void *orc_create(const orc_params_t *p) {
	GUARD_BEGIN
	Handle *h = new Handle();
	h->pose_kind = p->pose_kind;
	h->e2 = NULL; h->e3 = NULL;
	if (p->pose_kind == ORC_POSE_SE2) h->e2 = new Engine<Pose2>(*p);
	else h->e3 = new Engine<Pose3>(*p);
	return h;
	GUARD_END(NULL)
}
void orc_destroy(void *h) {
	Handle *H = static_cast<Handle *>(h);
	if (!H) return;
	delete H->e2;
	delete H->e3;
	delete H;
}
void orc_set_params(void *h, const orc_params_t *p) { try { DISPATCH(h, E.set_params(*p)); } catch (const std::exception &e) { g_last_error = e.what(); } }

int orc_define_new_keyframe(void *h, uint64_t n_obs, const uint64_t *feat_id, const double *obs_data, const uint8_t *is_fixed,
                            const uint8_t *is_unknown_with_init_val, const double *feat_rel_pos, int run_local_optimization,
                            uint64_t *out_kf_id, orc_new_edge_t *out_edges, uint64_t max_edges, uint64_t *out_n_edges, orc_opt_info_t *out_opt,
                            orc_opt_info_t *out_opt_stg1) {
	GUARD_BEGIN
	DISPATCH(h, {
		typedef typename std::remove_reference<decltype(E)>::type EngT;
		typename EngT::new_kf_observations_t lst;
		fill_obs(E, lst, n_obs, feat_id, obs_data, is_fixed, is_unknown_with_init_val, feat_rel_pos);
		typename EngT::TNewKeyFrameInfo info;
		std::memset(&info.optimize_results, 0, sizeof(orc_opt_info_t));
		std::memset(&info.optimize_results_stg1, 0, sizeof(orc_opt_info_t));
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
		if (out_opt_stg1) *out_opt_stg1 = info.optimize_results_stg1;
	});
	return 0;
	GUARD_END(-1)
}

int orc_optimize_local_area(void *h, uint64_t root, uint64_t win_size, int optimize_k2k_edges, int optimize_landmarks, uint64_t max_visitable_kf_id,
                            uint64_t dont_optimize_landmarks_seen_less_than_n_times, orc_opt_info_t *out) {
	GUARD_BEGIN
	DISPATCH(h, {
		typedef typename std::remove_reference<decltype(E)>::type EngT;
		typename EngT::TOptimizeLocalAreaParams prm;
		prm.optimize_k2k_edges = optimize_k2k_edges != 0;
		prm.optimize_landmarks = optimize_landmarks != 0;
		prm.max_visitable_kf_id = max_visitable_kf_id;
		prm.dont_optimize_landmarks_seen_less_than_n_times = dont_optimize_landmarks_seen_less_than_n_times;
		orc_opt_info_t info;
		std::memset(&info, 0, sizeof(info));
		E.optimize_local_area(root, static_cast<unsigned int>(std::min<uint64_t>(win_size, 0xffffffffu)), info, prm);
		if (out) *out = info;
	});
	return 0;
	GUARD_END(-1)
}

int orc_optimize_local_area_obs(void *h, uint64_t root, uint64_t win_size, uint64_t n_obs_idx, const uint64_t *obs_idx, orc_opt_info_t *out) {
	GUARD_BEGIN
	DISPATCH(h, {
		typedef typename std::remove_reference<decltype(E)>::type EngT;
		std::vector<size_t> lst(obs_idx, obs_idx + n_obs_idx);
		orc_opt_info_t info;
		std::memset(&info, 0, sizeof(info));
		E.optimize_local_area(root, static_cast<unsigned int>(std::min<uint64_t>(win_size, 0xffffffffu)), info, typename EngT::TOptimizeLocalAreaParams(), lst);
		if (out) *out = info;
	});
	return 0;
	GUARD_END(-1)
}
void orc_set_num(void *h, uint64_t src, uint64_t dst, const double *pose) {
	DISPATCH(h, E.spanning_tree.num[src][dst].pose.fromFlat(pose));
}

int orc_optimize_edges(void *h, uint64_t n_k2k, const uint64_t *k2k, uint64_t n_lm, const uint64_t *lms, orc_opt_info_t *out) {
	GUARD_BEGIN
	DISPATCH(h, {
		std::vector<size_t> a(k2k, k2k + n_k2k), b(lms, lms + n_lm);
		orc_opt_info_t info;
		std::memset(&info, 0, sizeof(info));
		E.optimize_edges(a, b, info);
		if (out) *out = info;
	});
	return 0;
	GUARD_END(-1)
}

uint64_t orc_alloc_keyframe(void *h) {
	uint64_t r = 0;
	DISPATCH(h, r = E.alloc_keyframe());
	return r;
}
uint64_t orc_create_kf2kf_edge(void *h, uint64_t new_kf_id, uint64_t from, uint64_t to, const double *init_inv_pose) {
	GUARD_BEGIN
	uint64_t r = 0;
	DISPATCH(h, {
		typedef typename std::remove_reference<decltype(E)>::type EngT;
		typename EngT::pose_t p;
		if (init_inv_pose) p.fromFlat(init_inv_pose);
		r = E.create_kf2kf_edge(new_kf_id, from, to, p);
	});
	return r;
	GUARD_END(static_cast<uint64_t>(-1))
}
void orc_update_numeric_all(void *h) { DISPATCH(h, E.update_numeric(false)); }
int orc_are_keyframes_connected(void *h, uint64_t a, uint64_t b) {
	GUARD_BEGIN
	int r = 0;
	DISPATCH(h, r = E.are_keyframes_connected(a, b) ? 1 : 0);
	return r;
	GUARD_END(-1)
}
uint64_t orc_complete_spanning_tree(void *h, uint64_t root, uint64_t max_depth, uint64_t *ids, double *poses, uint64_t cap) {
	GUARD_BEGIN
	uint64_t n = 0;
	DISPATCH(h, {
		typedef typename std::remove_reference<decltype(E)>::type EngT;
		std::map<TKeyFrameID, typename EngT::pose_t> st;
		E.create_complete_spanning_tree(root, st, max_depth);
		n = st.size();
		uint64_t k = 0;
		for (typename std::map<TKeyFrameID, typename EngT::pose_t>::const_iterator it = st.begin(); it != st.end() && k < cap; ++it, ++k) {
			if (ids) ids[k] = it->first;
			if (poses) it->second.toFlat(poses + k * EngT::pose_t::NVALS);
		}
	});
	return n;
	GUARD_END(0)
}
uint64_t orc_get_jacob_validity(void *h, uint8_t *out, uint64_t cap) {  /* all_observations_Jacob_validity (srba_types.h:722) */
	GUARD_BEGIN
	uint64_t n = 0;
	DISPATCH(h, {
		n = E.all_observations_Jacob_validity.size();
		for (uint64_t i = 0; out && i < std::min<uint64_t>(n, cap); i++) out[i] = E.all_observations_Jacob_validity[i] ? 1 : 0;
	});
	return n;
	GUARD_END(0)
}
double orc_eval_overall_squared_error(void *h) {
	GUARD_BEGIN
	double r = 0;
	DISPATCH(h, r = E.eval_overall_squared_error());
	return r;
	GUARD_END(-1.0)
}
uint64_t orc_last_dense_hessian(void *h, double *out, uint64_t cap, int *valid) {
	GUARD_BEGIN
	uint64_t n = 0;
	DISPATCH(h, {
		n = E.last_dense_hessian.size();
		if (out) std::copy(E.last_dense_hessian.begin(), E.last_dense_hessian.begin() + std::min<uint64_t>(n, cap), out);
		if (valid) *valid = E.last_hessian_valid ? 1 : 0;
	});
	return n;
	GUARD_END(0)
}
int64_t orc_find_path_bfs_len(void *h, uint64_t from, uint64_t to) {
	GUARD_BEGIN
	int64_t r = -1;
	DISPATCH(h, {
		std::vector<TKeyFrameID> path;
		if (E.find_path_bfs(from, to, &path, NULL)) r = static_cast<int64_t>(path.size());
	});
	return r;
	GUARD_END(-2)
}

uint64_t orc_num_keyframes(void *h) { uint64_t r = 0; DISPATCH(h, r = E.keyframes.size()); return r; }
uint64_t orc_num_edges(void *h) { uint64_t r = 0; DISPATCH(h, r = E.k2k_edges.size()); return r; }
uint64_t orc_num_observations(void *h) { uint64_t r = 0; DISPATCH(h, r = E.all_observations.size()); return r; }
int orc_pose_nvals(void *h) { return static_cast<Handle *>(h)->pose_kind == ORC_POSE_SE2 ? 3 : 12; }
void orc_get_edge(void *h, uint64_t id, uint64_t *from, uint64_t *to, double *inv_pose) {
	DISPATCH(h, {
		if (from) *from = E.k2k_edges[id].from;
		if (to) *to = E.k2k_edges[id].to;
		if (inv_pose) E.k2k_edges[id].inv_pose.toFlat(inv_pose);
	});
}
void orc_set_edge_pose(void *h, uint64_t id, const double *inv_pose) { DISPATCH(h, E.k2k_edges[id].inv_pose.fromFlat(inv_pose)); }
uint64_t orc_num_unknown_lms(void *h) { uint64_t r = 0; DISPATCH(h, r = E.unknown_lms.size()); return r; }
uint64_t orc_num_known_lms(void *h) { uint64_t r = 0; DISPATCH(h, r = E.known_lms.size()); return r; }
void orc_get_unknown_lms(void *h, uint64_t *ids, uint64_t *base, double *pos) {
	DISPATCH(h, {
		size_t k = 0;
		for (auto it = E.unknown_lms.begin(); it != E.unknown_lms.end(); ++it, ++k) {
			if (ids) ids[k] = it->first;
			if (base) base[k] = it->second.id_frame_base;
			if (pos) for (int j = 0; j < LM_DIMS; j++) pos[k * LM_DIMS + j] = it->second.pos[j];
		}
	});
}
void orc_get_known_lms(void *h, uint64_t *ids, uint64_t *base, double *pos) {
	DISPATCH(h, {
		size_t k = 0;
		for (auto it = E.known_lms.begin(); it != E.known_lms.end(); ++it, ++k) {
			if (ids) ids[k] = it->first;
			if (base) base[k] = it->second.id_frame_base;
			if (pos) for (int j = 0; j < LM_DIMS; j++) pos[k * LM_DIMS + j] = it->second.pos[j];
		}
	});
}
void orc_set_lm_pos(void *h, uint64_t id, const double *pos) {
	DISPATCH(h, {
		if (id < E.all_lms.size() && E.all_lms[id].rfp)
			for (int j = 0; j < LM_DIMS; j++) E.all_lms[id].rfp->pos[j] = pos[j];
	});
}
uint64_t orc_get_next_edge(void *h, uint64_t kf, uint64_t *target, uint64_t *next, uint64_t *dist, uint64_t cap) {
	uint64_t n = 0;
	DISPATCH(h, {
		auto *m = E.spanning_tree.sym.next_edge.find(kf);
		if (m) {
			n = m->size();
			uint64_t k = 0;
			for (auto it = m->begin(); it != m->end() && k < cap; ++it, ++k) {
				if (target) target[k] = it->first;
				if (next) next[k] = it->second.next;
				if (dist) dist[k] = it->second.distance;
			}
		}
	});
	return n;
}
uint64_t orc_get_num(void *h, uint64_t kf, uint64_t *target, double *poses, uint8_t *updated, uint64_t cap) {
	uint64_t n = 0;
	DISPATCH(h, {
		typedef typename std::remove_reference<decltype(E)>::type EngT;
		auto *m = E.spanning_tree.num.find(kf);
		if (m) {
			n = m->size();
			uint64_t k = 0;
			for (auto it = m->begin(); it != m->end() && k < cap; ++it, ++k) {
				if (target) target[k] = it->first;
				if (poses) it->second.pose.toFlat(poses + k * EngT::pose_t::NVALS);
				if (updated) updated[k] = it->second.updated ? 1 : 0;
			}
		}
	});
	return n;
}
int64_t orc_get_all_edges_path(void *h, uint64_t from, uint64_t to, uint64_t *edge_ids, uint64_t cap) {
	int64_t n = -1;
	DISPATCH(h, {
		auto *m = E.spanning_tree.sym.all_edges.find(from);
		if (m) {
			auto it = m->find(to);
			if (it != m->end()) {
				n = static_cast<int64_t>(it->second.size());
				for (size_t k = 0; k < it->second.size() && k < cap; k++) edge_ids[k] = it->second[k]->id;
			}
		}
	});
	return n;
}

void orc_trace_enable(void *h, int on) { DISPATCH(h, E.trace_on = on != 0); }
void orc_trace_clear(void *h) { DISPATCH(h, E.traces.clear()); }
uint64_t orc_trace_num_calls(void *h) { uint64_t r = 0; DISPATCH(h, r = E.traces.size()); return r; }
void orc_trace_call_sizes(void *h, uint64_t call, uint64_t *n_k2k, uint64_t *n_lm, uint64_t *n_obs, uint64_t *n_trials, uint64_t *n_scalars) {
	DISPATCH(h, {
		const OptTrace &t = E.traces[call];
		*n_k2k = t.k2k_ids.size(); *n_lm = t.lm_ids.size(); *n_obs = t.obs_idx.size();
		*n_trials = t.trial_scalars.size() / 6; *n_scalars = t.n_scalars;
	});
}
void orc_trace_call_data(void *h, uint64_t call, uint64_t *k2k_ids, uint64_t *lm_ids, uint64_t *obs_idx, double *trial_scalars, double *deltas,
                         double *init_grad, double *lambda0) {
	DISPATCH(h, {
		const OptTrace &t = E.traces[call];
		if (k2k_ids) std::copy(t.k2k_ids.begin(), t.k2k_ids.end(), k2k_ids);
		if (lm_ids) std::copy(t.lm_ids.begin(), t.lm_ids.end(), lm_ids);
		if (obs_idx) std::copy(t.obs_idx.begin(), t.obs_idx.end(), obs_idx);
		if (trial_scalars) std::copy(t.trial_scalars.begin(), t.trial_scalars.end(), trial_scalars);
		if (deltas) std::copy(t.deltas.begin(), t.deltas.end(), deltas);
		if (init_grad) std::copy(t.init_grad.begin(), t.init_grad.end(), init_grad);
		if (lambda0) *lambda0 = t.lambda0;
	});
}
double orc_total_opt_seconds(void *h) { double r = 0; DISPATCH(h, r = E.total_opt_seconds); return r; }
uint64_t orc_total_opt_trials(void *h) { uint64_t r = 0; DISPATCH(h, r = E.total_opt_trials); return r; }
void orc_reset_opt_timer(void *h) { DISPATCH(h, { E.total_opt_seconds = 0; E.total_opt_trials = 0; }); }

// ---- tests/schur_unittest.cpp:71-279 restated on the oracle's containers
int orc_schur_case(uint64_t n_p, uint64_t n_f, const uint8_t *visible, const double *JA, const double *JF, double lambda, double *out_HAp_dense,
                   double *out_grad_Ap, uint64_t *out_n_invertible) {
	GUARD_BEGIN
	const int p = 6, f = 3, o = 3;
	static char valid_true = 1;
	std::vector<JacobColView> dh_dAp(n_p), dh_df(n_f);
	size_t idx_obs = 0, ia = 0, jf = 0;
	for (size_t nKF = 0; nKF <= n_p; nKF++)
		for (size_t nLM = 0; nLM < n_f; nLM++) {
			if (!visible[nKF * n_f + nLM]) continue;
			if (nKF > 0) { JacobBlockRef r = {JA + (ia++) * o * p, &valid_true}; dh_dAp[nKF - 1].push_back(std::make_pair(idx_obs, r)); }
			JacobBlockRef r = {JF + (jf++) * o * f, &valid_true};
			dh_df[nLM].push_back(std::make_pair(idx_obs, r));
			idx_obs++;
		}
	for (size_t j = 0; j < n_f; j++) std::sort(dh_df[j].begin(), dh_df[j].end(), [](const std::pair<size_t, JacobBlockRef> &a, const std::pair<size_t, JacobBlockRef> &b) { return a.first < b.first; });
	std::vector<double> minus_grad(p * n_p + f * n_f, 1.0);
	SparseBlockHessian HAp, Hf, HApf;
	sparse_hessian_build_symbolic(HAp, Hf, HApf, dh_dAp, dh_df, p, f);
	NoiseModel noise;
	noise.kind = ORC_NOISE_IDENTITY; noise.o = o; noise.std_noise = 1.0;
	sparse_hessian_update_numeric(HAp, noise);
	sparse_hessian_update_numeric(Hf, noise);
	sparse_hessian_update_numeric(HApf, noise);
	SchurComplement schur(HAp, Hf, HApf, &minus_grad[0], &minus_grad[p * n_p], p, f);
	schur.numeric_build_reduced_system(lambda);
	const size_t n = p * n_p;
	std::fill(out_HAp_dense, out_HAp_dense + n * n, 0.0);
	for (size_t i = 0; i < n_p; i++)
		for (SparseBlockHessian::col_t::const_iterator it = HAp.getCol(i).begin(); it != HAp.getCol(i).end(); ++it)
			for (int a = 0; a < p; a++)
				for (int b = 0; b < p; b++) {
					out_HAp_dense[(p * it->first + a) * n + p * i + b] = it->second.num[a * p + b];
					out_HAp_dense[(p * i + b) * n + p * it->first + a] = it->second.num[a * p + b];
				}
	std::copy(minus_grad.begin(), minus_grad.begin() + n, out_grad_Ap);
	if (out_n_invertible) *out_n_invertible = schur.getNumFeaturesFullRank();
	return 0;
	GUARD_END(-1)
}


// The three solver_engine<> variants (impl/lev-marq_solvers.h:27-209, 213-405, 409-591) on the system of orc_schur_case
// (minus_grad = ones): one solve(lambda); returns delta_eps [6 n_p | 3 n_f] and the gradient vector as solve() leaves it.
int orc_solver_case(int solver_kind, uint64_t n_p, uint64_t n_f, const uint8_t *visible, const double *JA, const double *JF, double lambda,
                    double *out_delta, double *out_minus_grad_after, int *out_solved) {
	GUARD_BEGIN
	const int p = 6, f = 3, o = 3;
	static char valid_true = 1;
	std::vector<JacobColView> dh_dAp(n_p), dh_df(n_f);
	size_t idx_obs = 0, ia = 0, jf = 0;
	for (size_t nKF = 0; nKF <= n_p; nKF++)
		for (size_t nLM = 0; nLM < n_f; nLM++) {
			if (!visible[nKF * n_f + nLM]) continue;
			if (nKF > 0) { JacobBlockRef r = {JA + (ia++) * o * p, &valid_true}; dh_dAp[nKF - 1].push_back(std::make_pair(idx_obs, r)); }
			JacobBlockRef r = {JF + (jf++) * o * f, &valid_true};
			dh_df[nLM].push_back(std::make_pair(idx_obs, r));
			idx_obs++;
		}
	for (size_t j = 0; j < n_f; j++) std::sort(dh_df[j].begin(), dh_df[j].end(), [](const std::pair<size_t, JacobBlockRef> &a, const std::pair<size_t, JacobBlockRef> &b) { return a.first < b.first; });
	std::vector<double> minus_grad(p * n_p + f * n_f, 1.0);
	SparseBlockHessian HAp, Hf, HApf;
	sparse_hessian_build_symbolic(HAp, Hf, HApf, dh_dAp, dh_df, p, f);
	NoiseModel noise;
	noise.kind = ORC_NOISE_IDENTITY; noise.o = o; noise.std_noise = 1.0;
	sparse_hessian_update_numeric(HAp, noise);
	sparse_hessian_update_numeric(Hf, noise);
	sparse_hessian_update_numeric(HApf, noise);
	std::unique_ptr<SolverBase> sv;
	if (solver_kind == SOLVER_LM_SCHUR_DENSE_CHOLESKY) sv.reset(new SolverSchurDense(HAp, Hf, HApf, minus_grad, n_p, n_f, p, f));
	else if (solver_kind == SOLVER_LM_SCHUR_SPARSE_CHOLESKY) sv.reset(new SolverSchurSparse(HAp, Hf, HApf, minus_grad, n_p, n_f, p, f));
	else if (solver_kind == SOLVER_LM_NO_SCHUR_SPARSE_CHOLESKY) sv.reset(new SolverNoSchurSparse(HAp, Hf, HApf, minus_grad, n_p, n_f, p, f));
	else throw std::runtime_error("orc_solver_case: unknown solver_kind");
	const bool ok = sv->solve(lambda);
	*out_solved = ok ? 1 : 0;
	if (ok) std::copy(sv->delta_eps.begin(), sv->delta_eps.end(), out_delta);
	std::copy(minus_grad.begin(), minus_grad.end(), out_minus_grad_after);
	return 0;
	GUARD_END(-1)
}

int orc_eval_dh_dAp(const orc_params_t *p, const double *A, const double *D, const double *edge_inv_pose, int normal_dir, const double *lm_pos, double *out_J) {
	GUARD_BEGIN
	if (p->pose_kind == ORC_POSE_SE2) return eval_dh_dAp_t<Pose2>(p, A, D, edge_inv_pose, normal_dir, lm_pos, out_J);
	return eval_dh_dAp_t<Pose3>(p, A, D, edge_inv_pose, normal_dir, lm_pos, out_J);
	GUARD_END(-1)
}
int orc_eval_dh_df(const orc_params_t *p, const double *pose_base_wrt_obs, const double *lm_pos, double *out_J) {
	GUARD_BEGIN
	if (p->pose_kind != ORC_POSE_SE3) return -1;
	Engine<Pose3> E(*p);
	Engine<Pose3>::pose_flag_t fp;
	if (pose_base_wrt_obs) { fp.pose.fromFlat(pose_base_wrt_obs); fp.updated = true; }
	Engine<Pose3>::TRelativeLandmarkPos rfp;
	rfp.id_frame_base = 0;
	for (int k = 0; k < 3; k++) rfp.pos[k] = lm_pos[k];
	char valid = 1;
	Engine<Pose3>::EntryF ent;
	ent.sym.feat_rel_pos = &rfp;
	ent.sym.rel_pose_base_from_obs = pose_base_wrt_obs ? &fp : NULL;
	ent.sym.obs_idx = 0; ent.sym.is_valid = &valid;
	E.compute_jacobian_dh_df(ent, NULL);
	std::copy(ent.num, ent.num + E.O * 3, out_J);
	return valid;
	GUARD_END(-1)
}
int orc_predict(const orc_params_t *p, const double *pose_base_wrt_obs, const double *lm_pos, const double *z, double *out_err) {
	GUARD_BEGIN
	if (p->pose_kind == ORC_POSE_SE2) {
		Engine<Pose2> E(*p);
		Pose2 P; P.fromFlat(pose_base_wrt_obs);
		E.observe_error(out_err, z, P, lm_pos);
	} else {
		Engine<Pose3> E(*p);
		Pose3 P; P.fromFlat(pose_base_wrt_obs);
		E.observe_error(out_err, z, P, lm_pos);
	}
	return 0;
	GUARD_END(-1)
}
void orc_pose_compose(int pose_kind, const double *a, const double *b, double *out) {
	if (pose_kind == ORC_POSE_SE2) { Pose2 A, B, C; A.fromFlat(a); B.fromFlat(b); C.composeFrom(A, B); C.toFlat(out); }
	else { Pose3 A, B, C; A.fromFlat(a); B.fromFlat(b); C.composeFrom(A, B); C.toFlat(out); }
}
void orc_pose_inverse(int pose_kind, const double *a, double *out) {
	if (pose_kind == ORC_POSE_SE2) { Pose2 A; A.fromFlat(a); A.inverse(); A.toFlat(out); }
	else { Pose3 A; A.fromFlat(a); A.inverse(); A.toFlat(out); }
}
void orc_pose_exp(int pose_kind, const double *v, double *out) {
	if (pose_kind == ORC_POSE_SE2) Pose2::pseudo_exp(v).toFlat(out);
	else Pose3::pseudo_exp(v).toFlat(out);
}
void orc_pose_from_xyzypr(const double *v, double *out12) { Pose3(v[0], v[1], v[2], v[3], v[4], v[5]).toFlat(out12); }

}  // extern "C"
