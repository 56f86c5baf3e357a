// Seam-level shim (shim/srba_b200_shim.hpp) exercised on the reference's DATA MODEL as the test oracle restates it (oracle/srba_oracle.hpp:
// the same containers as TRBA_Problem_state -- lin_system.dh_dAp / dh_df column maps with symbolic entries, spanning_tree.sym.all_edges,
// spanning_tree.num, all_observations, all_lms, k2k_edges; MRPT itself cannot be compiled here).  TEST PROGRAM: the oracle is test
// infrastructure and is only linked here.
//   1. an oracle engine runs a sequential SLAM over the frames of the scenario file (CPU);
//   2. srba_b200_shim::flatten walks ITS pointer graph for optimize_local_area(root, win)'s unknowns;
//   3. the product's host engine (C ABI) is fed the same frames, gets the oracle's state injected and flattens the same local area:
//      every array of the two srba_problem_t must be IDENTICAL (sizes, indices bit-exact, values bit-exact);
//   4. with "run": srba_b200_optimize() on the shim's problem (GPU), srba_b200_shim::scatter back into the oracle's state, and a second,
//      untouched copy of the oracle engine runs its own CPU optimize_local_area: edges / landmarks must agree to 1e-6.
// Usage: seam_oracle <scenario file> [run]
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "../../oracle/srba_oracle.hpp"
#include "../../include/srba_b200_engine.h"
#include "../../shim/srba_b200_shim.hpp"

using namespace srba_oracle;

template <class POSE, int OBS>
struct OracleAccess {
	typedef Engine<POSE> ST;
	enum { POSE_NVALS = (POSE::DIMS == 3 ? 3 : 12), OBS_DIMS = OBS };
	typedef typename ST::colAp_t colAp_t;
	typedef typename ST::colF_t colF_t;
	typedef std::map<TKeyFrameID, typename ST::k2k_edge_vector_t> edges_row_t;
	static const colAp_t &colAp(const ST &st, size_t e) { return st.lin_system.dh_dAp[e]; }
	static const colF_t &colF(const ST &st, size_t id) { return st.lin_system.dh_df[st.lin_system.dh_df_remap[id]]; }
	static size_t n_k2k(const ST &st) { return st.k2k_edges.size(); }
	static uint64_t edge_from(const ST &st, size_t e) { return st.k2k_edges[e].from; }
	static uint64_t edge_to(const ST &st, size_t e) { return st.k2k_edges[e].to; }
	static void edge_pose_get(const ST &st, size_t e, double *o) { st.k2k_edges[e].inv_pose.toFlat(o); }
	static void edge_pose_set(ST &st, size_t e, const double *o) { st.k2k_edges[e].inv_pose.fromFlat(o); }
	static bool lm_known(const ST &st, size_t id) { return st.all_lms[id].has_known_pos; }
	static uint64_t lm_base(const ST &st, size_t id) { return st.all_lms[id].rfp->id_frame_base; }
	static double *lm_pos(const ST &st, size_t id) { return st.all_lms[id].rfp->pos; }
	static size_t n_obs(const ST &st) { return st.all_observations.size(); }
	static uint64_t obs_kf(const ST &st, size_t g) { return st.all_observations[g].kf_id; }
	static uint64_t obs_feat(const ST &st, size_t g) { return st.all_observations[g].feat_id; }
	static bool obs_first_of_unknown(const ST &st, size_t g) { return st.all_observations[g].is_first_obs_of_unknown; }
	static void obs_data(const ST &st, size_t g, double *o) { for (int k = 0; k < OBS; k++) o[k] = st.all_observations[g].obs_arr[k]; }
	static const edges_row_t *all_edges_from(const ST &st, uint64_t kf) { return st.spanning_tree.sym.all_edges.find(kf); }
	static uint64_t edge_id(const typename ST::k2k_edge_t *e) { return e->id; }
	static void set_num(ST &st, uint64_t src, uint64_t dst, const double *o) { st.spanning_tree.num[src][dst].pose.fromFlat(o); }
	static void set_validity(ST &st, size_t g, char v) { st.all_observations_Jacob_validity[g] = v; }
};

struct FeedbackCall { unsigned iter; double err, rmse; };
static std::vector<FeedbackCall> g_feedback;
static void record_feedback(unsigned int iter, const double total_sq_err, const double mean_sqroot_error) { g_feedback.push_back({iter, total_sq_err, mean_sqroot_error}); }

#define CHECK(c) do { if (!(c)) { std::printf("FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)
template <class T>
static bool same(const T *a, const T *b, size_t n) { return n == 0 || std::memcmp(a, b, n * sizeof(T)) == 0; }

template <class POSE, int OBS>
static int run(const orc_params_t &prm, std::istream &f, bool run_gpu) {
	typedef Engine<POSE> Eng;
	typedef OracleAccess<POSE, OBS> A;
	const int S = A::POSE_NVALS;
	size_t n_frames, win;
	f >> n_frames >> win;
	Eng eo(prm);
	srba_engine_params_t pp;
	static_assert(sizeof(pp) == sizeof(prm), "the two parameter structs mirror each other field by field");
	std::memcpy(&pp, &prm, sizeof(pp));
	void *ep = srba_engine_create(&pp);
	CHECK(ep);
	for (size_t k = 0; k < n_frames; k++) {
		size_t n;
		f >> n;
		typename Eng::new_kf_observations_t obs(n);
		std::vector<uint64_t> ids(n);
		std::vector<double> z(n * OBS), rel(n * 3, 0.0);
		std::vector<uint8_t> fixed(n), init(n, 0);
		for (size_t i = 0; i < n; i++) {
			int fx;
			f >> ids[i] >> fx;
			fixed[i] = (uint8_t)fx;
			std::memset(&obs[i], 0, sizeof(obs[i]));
			obs[i].feat_id = ids[i]; obs[i].is_fixed = fx != 0; obs[i].is_unknown_with_init_val = false;
			for (int c = 0; c < OBS; c++) { f >> z[i * OBS + c]; obs[i].obs_data[c] = z[i * OBS + c]; }
		}
		typename Eng::TNewKeyFrameInfo info;
		eo.define_new_keyframe(obs, info, true);   // sequential SLAM on the CPU: a realistic state
		uint64_t kf, ne;
		std::vector<srba_new_edge_t> edges(n + 4);
		srba_opt_info_t o1, o2;
		CHECK(srba_engine_define_new_keyframe(ep, n, ids.data(), z.data(), fixed.data(), init.data(), rel.data(), 0, &kf, edges.data(), edges.size(), &ne, &o1, &o2) == 0);
		CHECK(kf == info.kf_id && ne == info.created_edge_ids.size());
	}
	// the sequential run has converged: move every edge a little so that the optimisation below has real work to do
	std::vector<double> v(12);
	for (size_t e = 0; e < eo.k2k_edges.size(); e++) {
		eo.k2k_edges[e].inv_pose.toFlat(v.data());
		v[S == 3 ? 0 : 9] += 2e-3 * (double)((e % 7) + 1);
		v[S == 3 ? 1 : 10] -= 1e-3 * (double)((e % 5) + 1);
		eo.k2k_edges[e].inv_pose.fromFlat(v.data());
	}
	// inject the oracle's state into the product's host engine (edges, landmarks): same numeric inputs on both sides
	for (size_t e = 0; e < eo.k2k_edges.size(); e++) { eo.k2k_edges[e].inv_pose.toFlat(v.data()); srba_engine_set_edge_pose(ep, e, v.data()); }
	for (typename Eng::TRelativeLandmarkPosMap::const_iterator it = eo.unknown_lms.begin(); it != eo.unknown_lms.end(); ++it) srba_engine_set_lm_pos(ep, it->first, it->second.pos);
	const uint64_t root = n_frames - 1;
	std::vector<size_t> k2k, lms;
	eo.select_local_area(root, (unsigned)win, typename Eng::TOptimizeLocalAreaParams(), k2k, lms);
	void *flat = srba_engine_flatten_local_area(ep, root, win);
	CHECK(flat);
	const srba_problem_t &q = *srba_flat_problem(flat);
	std::unique_ptr<srba_b200_shim::Flat> F = srba_b200_shim::flatten<A>(eo, q.model, k2k, lms);
	CHECK(F.get());
	const srba_problem_t &p = F->p;
	// ---- every array identical
	CHECK(p.n_edges == q.n_edges && p.n_unk_edges == q.n_unk_edges && p.n_lms == q.n_lms && p.n_unk_lms == q.n_unk_lms && p.n_pairs == q.n_pairs && p.n_path == q.n_path);
	CHECK(p.n_obs == q.n_obs && p.n_blkA == q.n_blkA && p.n_blkF == q.n_blkF && p.n_HAp == q.n_HAp && p.n_HAp_terms == q.n_HAp_terms && p.n_HApf == q.n_HApf);
	CHECK(p.n_HApf_terms == q.n_HApf_terms && p.n_S == q.n_S && p.n_S_terms == q.n_S_terms && p.n_total_obs_rows == q.n_total_obs_rows);
	CHECK(same(p.edge_pose, q.edge_pose, (size_t)p.n_edges * S) && same(p.lm_pos, q.lm_pos, (size_t)p.n_lms * 3) && same(p.obs_z, q.obs_z, (size_t)p.n_obs * OBS));
	CHECK(same(p.unk_edge, q.unk_edge, p.n_unk_edges) && same(p.unk_lm, q.unk_lm, p.n_unk_lms));
	CHECK(same(p.pair_path_ptr, q.pair_path_ptr, p.n_pairs + 1) && same(p.path_edge, q.path_edge, p.n_path) && same(p.path_dir, q.path_dir, p.n_path));
	CHECK(same(p.obs_slot, q.obs_slot, p.n_obs) && same(p.obs_lm, q.obs_lm, p.n_obs) && same(p.obs_mult, q.obs_mult, p.n_obs));
	CHECK(same(p.colA_ptr, q.colA_ptr, p.n_unk_edges + 1) && same(p.blkA_obs, q.blkA_obs, p.n_blkA) && same(p.blkA_slotA, q.blkA_slotA, p.n_blkA));
	CHECK(same(p.blkA_slotD, q.blkA_slotD, p.n_blkA) && same(p.blkA_normal_dir, q.blkA_normal_dir, p.n_blkA));
	CHECK(same(p.colF_ptr, q.colF_ptr, p.n_unk_lms + 1) && same(p.blkF_obs, q.blkF_obs, p.n_blkF) && same(p.blkF_slot, q.blkF_slot, p.n_blkF));
	CHECK(same(p.row_ptr, q.row_ptr, p.n_obs + 1) && same(p.row_blk, q.row_blk, p.n_blkA + p.n_blkF));
	CHECK(same(p.HAp_row, q.HAp_row, p.n_HAp) && same(p.HAp_col, q.HAp_col, p.n_HAp) && same(p.HAp_term_ptr, q.HAp_term_ptr, p.n_HAp + 1));
	CHECK(same(p.HAp_term_a, q.HAp_term_a, p.n_HAp_terms) && same(p.HAp_term_b, q.HAp_term_b, p.n_HAp_terms));
	CHECK(same(p.HApf_row, q.HApf_row, p.n_HApf) && same(p.HApf_col, q.HApf_col, p.n_HApf) && same(p.HApf_term_ptr, q.HApf_term_ptr, p.n_HApf + 1));
	CHECK(same(p.HApf_term_a, q.HApf_term_a, p.n_HApf_terms) && same(p.HApf_term_f, q.HApf_term_f, p.n_HApf_terms));
	CHECK(same(p.rowApf_ptr, q.rowApf_ptr, p.n_unk_edges + 1) && same(p.colApf_ptr, q.colApf_ptr, p.n_unk_lms + 1) && same(p.colApf_idx, q.colApf_idx, p.n_HApf));
	CHECK(same(p.S_row, q.S_row, p.n_S) && same(p.S_col, q.S_col, p.n_S) && same(p.S_HAp, q.S_HAp, p.n_S) && same(p.S_term_ptr, q.S_term_ptr, p.n_S + 1));
	CHECK(same(p.S_term_rowblk, q.S_term_rowblk, p.n_S_terms) && same(p.S_term_colblk, q.S_term_colblk, p.n_S_terms));
	std::printf("FLATTEN identical: n_p %d n_f %d n_obs %d blkA %d blkF %d pairs %d HAp %d HApf %d S %d S_terms %d\n", p.n_unk_edges, p.n_unk_lms, p.n_obs, p.n_blkA, p.n_blkF,
	            p.n_pairs, p.n_HAp, p.n_HApf, p.n_S, p.n_S_terms);
	{  // feedback_user_iteration replay on a synthetic trace (no device): accepted trials only, iter = number of accepted steps before
		srba_result_t fake;
		std::memset(&fake, 0, sizeof(fake));
		double tr[5 * 6] = {1e-3, 1, 0.8, 90.0, 1, 0,   1e-3, 0, 0, 0, 0, 0,   4e-3, 1, -0.1, 95.0, 0, 0,   1.6e-2, 1, 0.5, 40.0, 1, 1,   5e-3, 1, 0.9, 10.0, 1, 0};
		fake.trace = tr; fake.trace_len = 5; fake.trace_cap = 5;
		fake.stats.total_sqr_error_init = 100.0;
		g_feedback.clear();
		srba_b200_shim::replay_feedback(&record_feedback, *F, fake);
		CHECK(g_feedback.size() == 4);
		CHECK(g_feedback[0].iter == 0 && g_feedback[0].err == 100.0 && g_feedback[1].iter == 0 && g_feedback[1].err == 90.0);
		CHECK(g_feedback[2].iter == 1 && g_feedback[2].err == 40.0 && g_feedback[3].iter == 2 && g_feedback[3].err == 10.0);
		CHECK(std::fabs(g_feedback[3].rmse - std::sqrt(10.0 / p.n_total_obs_rows)) < 1e-15);
		srba_b200_shim::replay_feedback((void (*)(unsigned int, const double, const double))NULL, *F, fake);   // no callback installed: nothing happens
		g_feedback.clear();
	}
	if (run_gpu) {
		srba_result_t res = F->make_result_buffers(S, 256);   // with the per-trial trace: feedback_user_iteration is replayed from it
		const int rc = srba_b200_optimize(&F->p, &res);
		if (rc != SRBA_OK) { std::printf("FAILED: srba_b200_optimize: %s\n", srba_b200_last_error()); return 1; }
		// snapshot of the unknowns before scattering, then the oracle's own CPU solution from the same state
		std::vector<double> gpu_edges = F->r_edge_pose, gpu_lms = F->r_lm_pos;
		orc_opt_info_t info;
		std::memset(&info, 0, sizeof(info));
		eo.optimize_local_area(root, (unsigned)win, info);     // CPU reference result, in place
		double worst = 0, scale = 0;
		for (size_t i = 0; i < F->k2k_ids.size(); i++) {
			eo.k2k_edges[F->k2k_ids[i]].inv_pose.toFlat(v.data());
			for (int c = 0; c < S; c++) { worst = std::fmax(worst, std::fabs(v[c] - gpu_edges[i * S + c])); scale = std::fmax(scale, std::fabs(v[c])); }
		}
		double worst_lm = 0, scale_lm = 1e-300;
		for (size_t i = 0; i < F->lm_ids.size(); i++)
			for (int c = 0; c < 3; c++) {
				const double o = eo.all_lms[F->lm_ids[i]].rfp->pos[c];
				worst_lm = std::fmax(worst_lm, std::fabs(o - gpu_lms[i * 3 + c])); scale_lm = std::fmax(scale_lm, std::fabs(o));
			}
		std::printf("GPU through the seam shim vs the oracle's CPU optimize_local_area: trials %d vs %llu, edges %.3e, landmarks %.3e (relative)\n", res.stats.num_trials,
		            (unsigned long long)info.num_lm_trials, worst / scale, worst_lm / scale_lm);
		CHECK(worst <= 1e-6 * scale && worst_lm <= 1e-6 * scale_lm);
		// feedback_user_iteration replay (optimize_edges.h:417-418,595-596): (0, initial error), then one call per accepted step
		g_feedback.clear();
		srba_b200_shim::replay_feedback(&record_feedback, *F, res);
		CHECK(!g_feedback.empty());
		std::printf("FEEDBACK calls %d, GPU iters %d (oracle %llu), first %.17g (stats %.17g, oracle %.17g), last %.17g (stats %.17g)\n", (int)g_feedback.size(), res.stats.num_iters,
		            (unsigned long long)info.num_iters, g_feedback[0].err, res.stats.total_sqr_error_init, info.total_sqr_error_init, g_feedback.back().err, res.stats.total_sqr_error_final);
		CHECK(g_feedback.size() >= 2 && g_feedback.size() <= 1 + (size_t)res.stats.num_iters);  // final `iter` also counts the closing pass (trial counts at the noise floor may differ from the oracle's)
		CHECK(g_feedback[0].iter == 0 && g_feedback[0].err == res.stats.total_sqr_error_init);
		CHECK(std::fabs(g_feedback[0].err - info.total_sqr_error_init) <= 1e-9 * info.total_sqr_error_init);
		for (size_t i = 1; i < g_feedback.size(); i++) {
			CHECK(g_feedback[i].iter == i - 1 && g_feedback[i].err < g_feedback[i - 1].err);
			CHECK(std::fabs(g_feedback[i].rmse - std::sqrt(g_feedback[i].err / p.n_total_obs_rows)) <= 1e-15 + 1e-12 * g_feedback[i].rmse);
		}
		CHECK(g_feedback.back().err == res.stats.total_sqr_error_final);
		std::printf("FEEDBACK ok: %d calls\n", (int)g_feedback.size());
		// scatter: the GPU result written back into the pointer graph (edges, landmarks, numeric spanning tree, validity)
		srba_b200_shim::scatter<A>(eo, *F, res);
		eo.k2k_edges[F->k2k_ids[0]].inv_pose.toFlat(v.data());
		CHECK(same(v.data(), &gpu_edges[0], S));
		const std::pair<uint64_t, uint64_t> pr = F->pairs[0];
		eo.spanning_tree.num[pr.first][pr.second].pose.toFlat(v.data());
		CHECK(same(v.data(), &F->r_slot_pose[0], S));
		std::printf("SCATTER ok\n");
	}
	srba_flat_destroy(flat);
	srba_engine_destroy(ep);
	std::printf("OK\n");
	return 0;
}

int main(int argc, char **argv) {
	if (argc < 2) { std::printf("usage: seam_oracle <scenario> [run]\n"); return 2; }
	std::ifstream f(argv[1]);
	const bool run_gpu = argc > 2 && !std::strcmp(argv[2], "run");
	std::string kind;
	f >> kind;
	orc_params_t p;
	orc_default_params(&p);
	try {
		if (kind == "se2") {
			p.pose_kind = ORC_POSE_SE2; p.obs_kind = ORC_OBS_RELPOSES2D; p.noise_kind = ORC_NOISE_CONST_MATRIX;
			for (int i = 0; i < 16; i++) p.noise_lambda[i] = 0;
			f >> p.noise_lambda[0] >> p.noise_lambda[4] >> p.noise_lambda[8];
			p.max_tree_depth = p.max_optimize_depth = 3; p.submap_size = 10; p.min_obs_to_loop_closure = 1; p.max_error_per_obs_to_stop = 1e-8;
			return run<Pose2, 3>(p, f, run_gpu);
		}
		if (kind == "cartesian3d") {
			p.pose_kind = ORC_POSE_SE3; p.obs_kind = ORC_OBS_CARTESIAN3D; p.noise_kind = ORC_NOISE_IDENTITY;
			f >> p.noise_std;
			p.max_tree_depth = p.max_optimize_depth = 3; p.submap_size = 10; p.min_obs_to_loop_closure = 4;
			return run<Pose3, 3>(p, f, run_gpu);
		}
	} catch (const std::exception &e) {
		std::printf("FAILED: exception: %s\n", e.what());
		return 1;
	}
	std::printf("unknown scenario kind\n");
	return 2;
}
