// C ABI of the B200 hot path (include/srba_b200.h): packs flattened optimize_edges() problems into one pinned host
// image, moves it to HBM with ONE copy, launches the persistent LM kernel (lm_kernel.cu) and brings one packed result
// image back.  No CPU fallback: without a CUDA device every entry point fails with SRBA_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "dev_types.h"

extern "C" int srba_launch_lm_batch(int model, const char *img, const uint64_t *img_off, char *ws, const uint64_t *ws_off, char *res,
                                    const uint64_t *res_off, const int32_t *order, int n_problems, int grid, int panel_cap, void *stream);

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

int validate(const srba_problem_t &p) {
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
	const int P = m.pose_kind == SRBA_POSE_SE2 ? 3 : 6;
	const int o = obs_dims(m.obs_kind);
	// ASSERT_ABOVEEQ_(OBS_DIMS*nObs, POSE_DIMS*nUnknowns_k2k+LM_DIMS*nUnknowns_k2f)  optimize_edges.h:355
	if ((int64_t)o * p.n_total_obs_rows < (int64_t)P * p.n_unk_edges + 3LL * p.n_unk_lms)
		return fail(SRBA_ERR_RANK_DEFICIENT, "OBS_DIMS*nObs < number of scalar unknowns (optimize_edges.h:355)");
	if (m.noise_kind == SRBA_NOISE_IDENTITY && !(m.noise_std > 0)) return fail(SRBA_ERR_BAD_ARGUMENT, "noise_std must be > 0");
	return SRBA_OK;
}

template <class T>
void put(char *base, uint64_t off, const T *src, size_t n) {
	if (n) std::memcpy(base + off, src, n * sizeof(T));
}

// Fills the header's layout fields; when `base` != NULL also copies the arrays.
void pack_problem(const srba_problem_t &p, DevProblem &h, char *base, int trace_cap, int want_dense) {
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
	Layout L;
	L.add(sizeof(DevProblem));
#define PUT(T, name, count)                               \
	h.img_##name = L.add((uint64_t)(count) * sizeof(T));  \
	if (base) put<T>(base, h.img_##name, p.name, (size_t)(count));
	PUT(double, edge_pose, (size_t)p.n_edges * S)
	PUT(int32_t, unk_edge, p.n_unk_edges)
	PUT(double, lm_pos, (size_t)p.n_lms * 3)
	PUT(int32_t, unk_lm, p.n_unk_lms)
	PUT(int32_t, pair_path_ptr, p.n_pairs + 1)
	PUT(int32_t, path_edge, p.n_path)
	PUT(uint8_t, path_dir, p.n_path)
	PUT(double, obs_z, (size_t)p.n_obs * o)
	PUT(int32_t, obs_slot, p.n_obs)
	PUT(int32_t, obs_lm, p.n_obs)
	PUT(int32_t, obs_mult, p.n_obs)
	PUT(int32_t, colA_ptr, p.n_unk_edges + 1)
	PUT(int32_t, blkA_obs, p.n_blkA)
	PUT(int32_t, blkA_slotA, p.n_blkA)
	PUT(int32_t, blkA_slotD, p.n_blkA)
	PUT(uint8_t, blkA_normal_dir, p.n_blkA)
	PUT(int32_t, colF_ptr, p.n_unk_lms + 1)
	PUT(int32_t, blkF_obs, p.n_blkF)
	PUT(int32_t, blkF_slot, p.n_blkF)
	PUT(int32_t, row_ptr, p.n_obs + 1)
	PUT(int32_t, row_blk, p.n_blkA + p.n_blkF)
	PUT(int32_t, HAp_row, p.n_HAp)
	PUT(int32_t, HAp_col, p.n_HAp)
	PUT(int32_t, HAp_term_ptr, p.n_HAp + 1)
	PUT(int32_t, HAp_term_a, p.n_HAp_terms)
	PUT(int32_t, HAp_term_b, p.n_HAp_terms)
	PUT(int32_t, HApf_row, p.n_HApf)
	PUT(int32_t, HApf_col, p.n_HApf)
	PUT(int32_t, HApf_term_ptr, p.n_HApf + 1)
	PUT(int32_t, HApf_term_a, p.n_HApf_terms)
	PUT(int32_t, HApf_term_f, p.n_HApf_terms)
	PUT(int32_t, rowApf_ptr, p.n_unk_edges + 1)
	PUT(int32_t, colApf_ptr, p.n_unk_lms + 1)
	PUT(int32_t, colApf_idx, p.n_HApf)
	PUT(int32_t, S_row, p.n_S)
	PUT(int32_t, S_col, p.n_S)
	PUT(int32_t, S_HAp, p.n_S)
	PUT(int32_t, S_term_ptr, p.n_S + 1)
	PUT(int32_t, S_term_rowblk, p.n_S_terms)
	PUT(int32_t, S_term_colblk, p.n_S_terms)
#undef PUT
	// derived: column of every block
	h.img_blkA_col = L.add((uint64_t)p.n_blkA * 4);
	h.img_blkF_col = L.add((uint64_t)p.n_blkF * 4);
	if (base) {
		int32_t *ca = reinterpret_cast<int32_t *>(base + h.img_blkA_col);
		for (int c = 0; c < p.n_unk_edges; c++)
			for (int b = p.colA_ptr[c]; b < p.colA_ptr[c + 1]; b++) ca[b] = c;
		int32_t *cf = reinterpret_cast<int32_t *>(base + h.img_blkF_col);
		for (int c = 0; c < p.n_unk_lms; c++)
			for (int b = p.colF_ptr[c]; b < p.colF_ptr[c + 1]; b++) cf[b] = c;
	}
	h.img_bytes = L.total();

	Layout W;
#define WS(name, bytes) h.ws_##name = W.add((uint64_t)(bytes));
	WS(edge_cur, (size_t)p.n_edges * S * 8)
	WS(edge_bak, (size_t)p.n_unk_edges * S * 8)
	WS(lm_cur, (size_t)p.n_lms * 3 * 8)
	WS(lm_bak, (size_t)p.n_unk_lms * 3 * 8)
	WS(slot, (size_t)2 * p.n_pairs * S * 8)
	WS(slot_bak, (size_t)2 * p.n_pairs * S * 8)
	WS(JA, (size_t)p.n_blkA * o * P * 8)
	WS(JF, (size_t)p.n_blkF * o * 3 * 8)
	WS(resid0, (size_t)p.n_obs * o * 8)
	WS(resid1, (size_t)p.n_obs * o * 8)
	WS(HAp, (size_t)p.n_HAp * P * P * 8)
	WS(HApf, (size_t)p.n_HApf * P * 3 * 8)
	WS(Hf, (size_t)p.n_unk_lms * 9 * 8)
	WS(Hf_inv, (size_t)p.n_unk_lms * 9 * 8)
	WS(W, (size_t)p.n_HApf * P * 3 * 8)
	WS(grad, (size_t)h.n_scalars * 8)
	WS(delta, (size_t)h.n_scalars * 8)
	WS(dense, (size_t)h.n_dense * h.ld_dense * 8)
	WS(valid, p.n_obs)
	WS(blk_ok, p.n_blkA + p.n_blkF)
	WS(lm_inv, p.n_unk_lms)
	WS(slot_listed, 2 * p.n_pairs)
	WS(pair_refresh, p.n_pairs)
#undef WS
	h.ws_bytes = W.total();

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
	h.res_stage_cycles = R.add(16 * 8);
	h.res_bytes = R.total();
	if (base) std::memcpy(base, &h, sizeof(h));
}

double problem_cost(const DevProblem &h) {
	const double n = h.n_dense;
	return n * n * n / 3.0 + 40.0 * (h.n_blkA + h.n_blkF) + 100.0 * h.n_S_terms + 1000.0;
}
}  // namespace

// Grow-only buffer pool reused by the synchronous entry points (srba_b200_optimize / _optimize_batch): the library
// keeps no problem state between calls, only this reusable workspace (SURVEY.md §8b "optional reusable device workspace").
struct Pool {
	char *h_img = nullptr, *h_res = nullptr, *d_img = nullptr, *d_ws = nullptr, *d_res = nullptr;
	uint64_t h_img_cap = 0, h_res_cap = 0, d_img_cap = 0, d_ws_cap = 0, d_res_cap = 0;
	cudaStream_t stream = nullptr;
};
static thread_local Pool g_pool;

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
	std::vector<uint64_t> img_off, ws_off, res_off;
	std::vector<int32_t> order;      // grouped by model, cost-descending inside each group
	int group_begin[5] = {0, 0, 0, 0, 0};
	int group_max_n[4] = {0, 0, 0, 0};
	uint64_t tab_img_off = 0, tab_ws_off = 0, tab_res_off = 0, tab_order = 0;
	uint64_t img_bytes = 0, ws_bytes = 0, res_bytes = 0;
	char *h_img = nullptr, *h_res = nullptr;
	char *d_img = nullptr, *d_ws = nullptr, *d_res = nullptr;
	int64_t launches = 0;
	int sm_count = 148;
	bool uploaded = false, downloaded = false;
};

extern "C" {

const char *srba_b200_last_error(void) { return g_err.c_str(); }
const char *srba_b200_version(void) { return "srba_b200 0.1 (sm_100a)"; }
int srba_b200_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

void srba_b200_batch_destroy(srba_batch_t *b) {
	if (!b) return;
	if (b->pooled) { delete b; return; }
	if (b->h_img) cudaFreeHost(b->h_img);
	if (b->h_res) cudaFreeHost(b->h_res);
	if (b->d_img) cudaFree(b->d_img);
	if (b->d_ws) cudaFree(b->d_ws);
	if (b->d_res) cudaFree(b->d_res);
	delete b;
}

static srba_batch_t *batch_create_impl(int32_t n, const srba_problem_t *const *in, int32_t trace_cap, Pool *pool);
srba_batch_t *srba_b200_batch_create(int32_t n, const srba_problem_t *const *in, int32_t trace_cap) { return batch_create_impl(n, in, trace_cap, nullptr); }
static srba_batch_t *batch_create_impl(int32_t n, const srba_problem_t *const *in, int32_t trace_cap, Pool *pool) {
	if (n <= 0 || !in) { fail(SRBA_ERR_BAD_ARGUMENT, "empty batch"); return nullptr; }
	if (srba_b200_device_count() <= 0) { fail(SRBA_ERR_NO_DEVICE, "no CUDA device: the srba_b200 hot path has no CPU fallback"); return nullptr; }
	for (int i = 0; i < n; i++)
		if (!in[i] || validate(*in[i]) != SRBA_OK) {
			if (!in[i]) fail(SRBA_ERR_BAD_ARGUMENT, "null problem");
			return nullptr;
		}
	srba_batch_t *b = new srba_batch();
	b->n = n;
	b->hdr.resize(n);
	b->img_off.resize(n); b->ws_off.resize(n); b->res_off.resize(n); b->order.resize(n);
	const int want_dense = trace_cap < 0 ? 1 : 0;  // negative trace_cap: |trace_cap| trials of trace AND the dense Hessian
	const int tcap = trace_cap < 0 ? -trace_cap : trace_cap;
	Layout I;
	b->tab_img_off = I.add((uint64_t)n * 8);
	b->tab_ws_off = I.add((uint64_t)n * 8);
	b->tab_res_off = I.add((uint64_t)n * 8);
	b->tab_order = I.add((uint64_t)n * 4);
	uint64_t ws = 0, rs = 0;
	for (int i = 0; i < n; i++) {
		pack_problem(*in[i], b->hdr[i], nullptr, tcap, want_dense);
		b->img_off[i] = I.add(b->hdr[i].img_bytes);
		b->ws_off[i] = ws; ws += b->hdr[i].ws_bytes;
		b->res_off[i] = rs; rs += b->hdr[i].res_bytes;
	}
	b->img_bytes = I.total(); b->ws_bytes = std::max<uint64_t>(ws, 16); b->res_bytes = std::max<uint64_t>(rs, 16);
	std::iota(b->order.begin(), b->order.end(), 0);
	auto model_of = [&](int i) {
		const srba_model_t &m = b->hdr[i].model;
		return m.pose_kind == SRBA_POSE_SE2 ? 0 : (m.obs_kind == SRBA_OBS_CARTESIAN3D ? 1 : (m.obs_kind == SRBA_OBS_MONO ? 2 : 3));
	};
	std::stable_sort(b->order.begin(), b->order.end(), [&](int a, int c) {
		const int ma = model_of(a), mc = model_of(c);
		if (ma != mc) return ma < mc;
		return problem_cost(b->hdr[a]) > problem_cost(b->hdr[c]);
	});
	{
		int cnt[4] = {0, 0, 0, 0};
		for (int i = 0; i < n; i++) { const int g = model_of(i); cnt[g]++; b->group_max_n[g] = std::max(b->group_max_n[g], b->hdr[i].n_dense); }
		for (int g = 0; g < 4; g++) b->group_begin[g + 1] = b->group_begin[g] + cnt[g];
	}
	cudaDeviceProp prop;
	int dev = 0;
	if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess) b->sm_count = prop.multiProcessorCount;
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
	for (int i = 0; i < n; i++) {
		b->hdr[i].cost_rank = 0;
		pack_problem(*in[i], b->hdr[i], b->h_img + b->img_off[i], tcap, want_dense);
	}
	return b;
}

int srba_b200_batch_upload(srba_batch_t *b, void *stream) {
	if (!b) return fail(SRBA_ERR_BAD_ARGUMENT, "null batch");
	CUDA_TRY(cudaMemcpyAsync(b->d_img, b->h_img, b->img_bytes, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
	b->uploaded = true;
	return SRBA_OK;
}

int srba_b200_batch_run(srba_batch_t *b, void *stream) {
	if (!b) return fail(SRBA_ERR_BAD_ARGUMENT, "null batch");
	if (!b->uploaded) return fail(SRBA_ERR_BAD_ARGUMENT, "batch_run before batch_upload");
	b->launches = 0;
	for (int g = 0; g < 4; g++) {
		const int cnt = b->group_begin[g + 1] - b->group_begin[g];
		if (!cnt) continue;
		// one CTA per problem, dispatched in cost-descending order (largest first); staged Cholesky panel sized for the
		// largest problem of the group, capped so that 2 CTAs/SM stay resident
		const int panel_cap = std::min(std::max((b->group_max_n[g] + 1 + 3) / 4 * 4, 24), 1024);
		const int rc = srba_launch_lm_batch(g, b->d_img, reinterpret_cast<const uint64_t *>(b->d_img + b->tab_img_off), b->d_ws,
		                                    reinterpret_cast<const uint64_t *>(b->d_img + b->tab_ws_off), b->d_res,
		                                    reinterpret_cast<const uint64_t *>(b->d_img + b->tab_res_off),
		                                    reinterpret_cast<const int32_t *>(b->d_img + b->tab_order) + b->group_begin[g], cnt, cnt, panel_cap, stream);
		if (rc != 0) return fail(SRBA_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString((cudaError_t)rc));
		b->launches++;
	}
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
	return out->stats.status;
}

int srba_b200_batch_stage_cycles(srba_batch_t *b, int32_t i, uint64_t out[16]) {
	if (!b || i < 0 || i >= b->n || !b->downloaded) return fail(SRBA_ERR_BAD_ARGUMENT, "bad stage_cycles arguments");
	std::memcpy(out, b->h_res + b->res_off[i] + b->hdr[i].res_stage_cycles, 16 * 8);
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

int srba_b200_optimize_batch(int32_t n, const srba_problem_t *const *in, srba_result_t *const *out) {
	if (!out) return fail(SRBA_ERR_BAD_ARGUMENT, "null out");
	int tcap = 0, want_dense = 0;
	for (int i = 0; i < n; i++) {
		if (!out[i]) return fail(SRBA_ERR_BAD_ARGUMENT, "null result");
		tcap = std::max(tcap, out[i]->trace ? out[i]->trace_cap : 0);
		if (out[i]->dense_hessian) want_dense = 1;
	}
	if (want_dense && tcap == 0) tcap = 1;
	srba_batch_t *b = batch_create_impl(n, in, want_dense ? -tcap : tcap, &g_pool);
	if (!b) return g_code != SRBA_OK ? g_code : SRBA_ERR_INTERNAL;
	int rc = SRBA_OK;
	if (!g_pool.stream && cudaStreamCreateWithFlags(&g_pool.stream, cudaStreamNonBlocking) != cudaSuccess) rc = fail(SRBA_ERR_CUDA, "cudaStreamCreate failed");
	cudaStream_t st = g_pool.stream;
	if (rc == SRBA_OK) rc = srba_b200_batch_upload(b, st);
	if (rc == SRBA_OK) rc = srba_b200_batch_run(b, st);
	if (rc == SRBA_OK) rc = srba_b200_batch_download(b, st);
	if (rc == SRBA_OK)
		for (int i = 0; i < n; i++) {
			const int r = srba_b200_batch_unpack(b, i, out[i]);
			if (r != SRBA_OK && rc == SRBA_OK) rc = r;
		}
	srba_b200_batch_destroy(b);
	return rc;
}

int srba_b200_optimize(const srba_problem_t *in, srba_result_t *out) { return srba_b200_optimize_batch(1, &in, &out); }

}  // extern "C"
