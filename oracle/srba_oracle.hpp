// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of MRPT/srba's RbaEngine<> for the local Levenberg-Marquardt hot path
//  plus the graph front-end that produces its symbolic inputs.
// It deliberately keeps the reference's pointer-rich std::map / std::deque layout, its
// iteration orders and its behavioural quirks, because index order must be
// bit-exact.  Every function cites the reference file:line it follows (paths relative to
// /root/reference/include/srba/).
//
// PARITY STATUS: the reference cannot be compiled here (needs MRPT >= 1.3 + Eigen 3, neither
// vendored nor installed), so this restatement is pinned against the reference's own test
// *properties* (tests/{schur,spantree,fixed-transformations,sensor-pose}_unittest.cpp,
// tests/submaps_edge_init_values.cpp and the examples/cpp tutorials' ground truth), not
// against reference outputs: value-level parity of J/H/S/delta is "parity unpinned".
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
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
#include <vector>

#include "oracle_capi.h"
#include "poses.hpp"

namespace srba_oracle {

typedef uint64_t TKeyFrameID;
typedef uint64_t TLandmarkID;
typedef size_t topo_dist_t;
static const TKeyFrameID SRBA_INVALID_KEYFRAMEID = static_cast<TKeyFrameID>(-1);
static const size_t SRBA_INVALID_INDEX = static_cast<size_t>(-1);
static const int LM_DIMS = 3;  // Euclidean3D (landmarks.h:27) and RelativePoses2D (landmarks.h:99)
static const int MAX_OBS = 4;
static const int MAX_P = 6;

#define ORC_ASSERT(cond)                                                                         \
	do {                                                                                         \
		if (!(cond)) {                                                                           \
			std::ostringstream s_;                                                               \
			s_ << "ASSERT failed: " #cond " at " << __FILE__ << ":" << __LINE__;                 \
			throw std::runtime_error(s_.str());                                                  \
		}                                                                                        \
	} while (0)

inline int obs_dims_of(int obs_kind) {
	switch (obs_kind) {
		case ORC_OBS_RELPOSES2D: return 3;   // observations_RelativePoses_2D.h:25
		case ORC_OBS_CARTESIAN3D: return 3;  // observations_Cartesian_3D.h:22
		case ORC_OBS_MONO: return 2;         // observations_MonocularCamera.h
		case ORC_OBS_STEREO: return 4;       // observations_StereoCamera.h:32
	}
	throw std::runtime_error("bad obs_kind");
}

// mrpt::utils::map_as_vector<K,V>: operator[] grows, find() succeeds for any index < size().
template <class V>
struct MapAsVector {
	std::deque<V> d;
	V &operator[](size_t i) {
		if (i >= d.size()) d.resize(i + 1);
		return d[i];
	}
	V *find(size_t i) { return i < d.size() ? &d[i] : NULL; }
	const V *find(size_t i) const { return i < d.size() ? &d[i] : NULL; }
	size_t size() const { return d.size(); }
	void clear() { d.clear(); }
};

// ---------------------------------------------------------------------------------------------
// Eigen::FullPivLU<3x3>::isInvertible()/inverse()  (impl/schur.h:200-206)
// complete pivoting; rank = #pivots with |pivot| > eps*n*|max pivot|.
// ---------------------------------------------------------------------------------------------
inline bool fullpiv_lu_inverse3(const double *Ain, double *inv) {
	const int n = 3;
	double lu[9];
	std::memcpy(lu, Ain, sizeof(lu));
	int rowsTransp[3], colsTransp[3];
	int nonzero_pivots = n;
	double maxpivot = 0;
	for (int k = 0; k < n; k++) {
		int rbig = k, cbig = k;
		double big = 0;
		for (int r = k; r < n; r++)
			for (int c = k; c < n; c++)
				if (std::abs(lu[3 * r + c]) > big) { big = std::abs(lu[3 * r + c]); rbig = r; cbig = c; }
		if (big == 0) {
			nonzero_pivots = k;
			for (int i = k; i < n; i++) { rowsTransp[i] = i; colsTransp[i] = i; }
			break;
		}
		if (big > maxpivot) maxpivot = big;
		rowsTransp[k] = rbig; colsTransp[k] = cbig;
		if (k != rbig) for (int c = 0; c < n; c++) std::swap(lu[3 * k + c], lu[3 * rbig + c]);
		if (k != cbig) for (int r = 0; r < n; r++) std::swap(lu[3 * r + k], lu[3 * r + cbig]);
		if (k < n - 1) {
			for (int r = k + 1; r < n; r++) lu[3 * r + k] /= lu[3 * k + k];
			for (int r = k + 1; r < n; r++)
				for (int c = k + 1; c < n; c++) lu[3 * r + c] -= lu[3 * r + k] * lu[3 * k + c];
		}
	}
	// rank()
	const double thr = std::abs(maxpivot) * (std::numeric_limits<double>::epsilon() * n);
	int rank = 0;
	for (int i = 0; i < nonzero_pivots; i++) rank += (std::abs(lu[3 * i + i]) > thr) ? 1 : 0;
	if (rank != n) return false;
	// permutations: P = rowsTransp applied in order; Q = colsTransp.  A = P^-1 L U Q^-1.
	int p[3] = {0, 1, 2}, q[3] = {0, 1, 2};
	// Row permutation vector: result row i of P*A comes from row p[i] of A.
	for (int k = 0; k < n; k++) std::swap(p[k], p[rowsTransp[k]]);
	for (int k = 0; k < n; k++) std::swap(q[k], q[colsTransp[k]]);
	// Solve A X = I  column by column: L U (Q^-1 X) = P I
	for (int col = 0; col < n; col++) {
		double b[3];
		for (int i = 0; i < n; i++) b[i] = (p[i] == col) ? 1.0 : 0.0;
		for (int i = 0; i < n; i++)
			for (int j = 0; j < i; j++) b[i] -= lu[3 * i + j] * b[j];
		for (int i = n - 1; i >= 0; i--) {
			for (int j = i + 1; j < n; j++) b[i] -= lu[3 * i + j] * b[j];
			b[i] /= lu[3 * i + i];
		}
		for (int i = 0; i < n; i++) inv[3 * q[i] + col] = b[i];
	}
	return true;
}

// ---------------------------------------------------------------------------------------------
// Eigen::LLT<MatrixXd,Upper> of the self-adjoint upper view (impl/lev-marq_solvers.h:527-534,549).
// A is n x n row-major, only the upper triangle is read; U overwrites it (A = U^T U).
// Returns false when a pivot is <= 0 (Eigen's llt_inplace: "if (x<=0) return k").
// ---------------------------------------------------------------------------------------------
// Blocked right-looking evaluation order (cache-friendly, vectorisable -- Eigen's own LLT is blocked too), arranged so that every
// element receives exactly the operations of the textbook recurrence  u_kj = (a_kj - sum_{r<k} u_rk u_rj) / u_kk  with r ASCENDING:
// bit-identical to the unblocked loop, ~10x faster on the 1 000-2 000-dim systems of loop-closure areas.  (Entries below the
// diagonal are scratch: never read.)
inline bool dense_llt_upper(std::vector<double> &A, size_t n) {
	const size_t B = 48;
	double *a = A.data();
	for (size_t kb = 0; kb < n; kb += B) {
		const size_t ke = std::min(kb + B, n);
		for (size_t k = kb; k < ke; k++) {  // rows of the panel: contributions of the panel's earlier rows, then the pivot
			double *rk = a + k * n;
			for (size_t r = kb; r < k; r++) {
				const double *rr = a + r * n;
				const double m = rr[k];
				for (size_t j = k; j < n; j++) rk[j] -= m * rr[j];
			}
			double x = rk[k];
			if (x <= 0) return false;
			x = std::sqrt(x);
			rk[k] = x;
			for (size_t j = k + 1; j < n; j++) rk[j] = rk[j] / x;
		}
		size_t i = ke;
		for (; i + 4 <= n; i += 4) {  // trailing rows, four at a time (each panel row is streamed once per four rows)
			double *r0 = a + i * n, *r1 = r0 + n, *r2 = r1 + n, *r3 = r2 + n;
			for (size_t r = kb; r < ke; r++) {
				const double *rr = a + r * n;
				const double m0 = rr[i], m1 = rr[i + 1], m2 = rr[i + 2], m3 = rr[i + 3];
				for (size_t j = i; j < n; j++) {
					const double u = rr[j];
					r0[j] -= m0 * u; r1[j] -= m1 * u; r2[j] -= m2 * u; r3[j] -= m3 * u;
				}
			}
		}
		for (; i < n; i++) {
			double *ri = a + i * n;
			for (size_t r = kb; r < ke; r++) {
				const double *rr = a + r * n;
				const double m = rr[i];
				for (size_t j = i; j < n; j++) ri[j] -= m * rr[j];
			}
		}
	}
	return true;
}
inline void dense_llt_solve(const std::vector<double> &U, size_t n, const double *b, double *x) {
	std::vector<double> y(n);
	for (size_t i = 0; i < n; i++) {  // U^T y = b
		double s = b[i];
		for (size_t r = 0; r < i; r++) s -= U[r * n + i] * y[r];
		y[i] = s / U[i * n + i];
	}
	for (size_t ii = n; ii-- > 0;) {  // U x = y
		double s = y[ii];
		for (size_t c = ii + 1; c < n; c++) s -= U[ii * n + c] * x[c];
		x[ii] = s / U[ii * n + ii];
	}
}

// Horn's closed-form absolute orientation, standing in for mrpt::tfest::se3_l2 (un-vendored MRPT),
// call sites: models/observations_Cartesian_3D.h:73-78, observations_StereoCamera.h:110-116.
// Finds P such that this_i ~= P (+) other_i.  Needs >= 3 pairs.
inline bool se3_l2_horn(const std::vector<double> &pts_this, const std::vector<double> &pts_other, Pose3 &out) {
	const size_t N = pts_this.size() / 3;
	if (N < 3) return false;
	double ct[3] = {0, 0, 0}, co[3] = {0, 0, 0};
	for (size_t i = 0; i < N; i++)
		for (int k = 0; k < 3; k++) { ct[k] += pts_this[3 * i + k]; co[k] += pts_other[3 * i + k]; }
	for (int k = 0; k < 3; k++) { ct[k] /= N; co[k] /= N; }
	double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // S = sum other' * this'^T
	for (size_t i = 0; i < N; i++)
		for (int r = 0; r < 3; r++)
			for (int c = 0; c < 3; c++) S[3 * r + c] += (pts_other[3 * i + r] - co[r]) * (pts_this[3 * i + c] - ct[c]);
	const double Sxx = S[0], Sxy = S[1], Sxz = S[2], Syx = S[3], Syy = S[4], Syz = S[5], Szx = S[6], Szy = S[7], Szz = S[8];
	double Nm[16] = {Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx,
	                 Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz,
	                 Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy,
	                 Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz};
	// Jacobi eigen-decomposition of the symmetric 4x4 matrix.
	double V[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
	for (int sweep = 0; sweep < 100; sweep++) {
		double off = 0;
		for (int r = 0; r < 4; r++)
			for (int c = r + 1; c < 4; c++) off += Nm[4 * r + c] * Nm[4 * r + c];
		if (off < 1e-300) break;
		for (int p = 0; p < 4; p++)
			for (int q = p + 1; q < 4; q++) {
				if (Nm[4 * p + q] == 0) continue;
				const double theta = (Nm[4 * q + q] - Nm[4 * p + p]) / (2 * Nm[4 * p + q]);
				const double t = (theta >= 0 ? 1.0 : -1.0) / (std::abs(theta) + std::sqrt(theta * theta + 1));
				const double c = 1 / std::sqrt(t * t + 1), s = t * c;
				for (int k = 0; k < 4; k++) {
					const double akp = Nm[4 * k + p], akq = Nm[4 * k + q];
					Nm[4 * k + p] = c * akp - s * akq; Nm[4 * k + q] = s * akp + c * akq;
				}
				for (int k = 0; k < 4; k++) {
					const double apk = Nm[4 * p + k], aqk = Nm[4 * q + k];
					Nm[4 * p + k] = c * apk - s * aqk; Nm[4 * q + k] = s * apk + c * aqk;
				}
				for (int k = 0; k < 4; k++) {
					const double vkp = V[4 * k + p], vkq = V[4 * k + q];
					V[4 * k + p] = c * vkp - s * vkq; V[4 * k + q] = s * vkp + c * vkq;
				}
			}
	}
	int best = 0;
	for (int k = 1; k < 4; k++) if (Nm[4 * k + k] > Nm[4 * best + best]) best = k;
	double q[4] = {V[0 + best], V[4 + best], V[8 + best], V[12 + best]};
	const double nq = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
	for (int k = 0; k < 4; k++) q[k] /= nq;
	if (q[0] < 0) for (int k = 0; k < 4; k++) q[k] = -q[k];
	out = pose3_from_quat(0, 0, 0, q[0], q[1], q[2], q[3]);
	for (int r = 0; r < 3; r++)
		out.t[r] = ct[r] - (out.R[3 * r] * co[0] + out.R[3 * r + 1] * co[1] + out.R[3 * r + 2] * co[2]);
	return true;
}

// ---------------------------------------------------------------------------------------------
// Block-sparse Hessian containers (srba_types.h:274-382; MRPT MatrixBlockSparseCols semantics)
// ---------------------------------------------------------------------------------------------
struct THessianSymbolicInfoEntry {  // srba_types.h:290-303
	const double *J1, *J2;
	const char *J1_valid, *J2_valid;
	size_t obs_idx;
};
struct HEntry {
	double num[MAX_P * MAX_P];
	std::vector<THessianSymbolicInfoEntry> lst_jacob_blocks;
	HEntry() { std::memset(num, 0, sizeof(num)); }
};
struct SparseBlockHessian {
	typedef std::map<size_t, HEntry> col_t;
	std::vector<col_t> cols;
	int nr, nc;  // block dims
	SparseBlockHessian() : nr(0), nc(0) {}
	void setColCount(size_t n) { cols.resize(n); }
	size_t getColCount() const { return cols.size(); }
	col_t &getCol(size_t i) { return cols[i]; }
	const col_t &getCol(size_t i) const { return cols[i]; }
	// MatrixBlockSparseCols::copyNumericalValuesFrom: copy the values of every block of `o`; blocks that exist
	// here but not in `o` are zeroed (un-vendored MRPT; the zeroing is the semantics impl/schur.h:188-190 needs
	// for the Schur fill-in blocks created at :142-149, which are absent from HAp_original — see DESIGN.md).
	void copyNumericalValuesFrom(const SparseBlockHessian &o) {
		nr = o.nr; nc = o.nc;
		if (cols.size() != o.cols.size()) cols.resize(o.cols.size());
		for (size_t i = 0; i < o.cols.size(); i++) {
			col_t &dst = cols[i];
			const col_t &src = o.cols[i];
			col_t::iterator itd = dst.begin();
			col_t::const_iterator its = src.begin();
			while (its != src.end()) {
				if (itd == dst.end() || itd->first > its->first) {
					itd = dst.insert(itd, std::make_pair(its->first, HEntry()));
					std::memcpy(itd->second.num, its->second.num, sizeof(its->second.num));
					++itd; ++its;
				} else if (itd->first < its->first) {
					std::memset(itd->second.num, 0, sizeof(itd->second.num));
					++itd;
				} else {
					std::memcpy(itd->second.num, its->second.num, sizeof(its->second.num));
					++itd; ++its;
				}
			}
			for (; itd != dst.end(); ++itd) std::memset(itd->second.num, 0, sizeof(itd->second.num));
		}
	}
};

// Observation-noise policy (srba_options_noise.h:25-133), run-time dispatched.
struct NoiseModel {
	int kind;        // ORC_NOISE_*
	int o;           // OBS_DIMS
	double std_noise;
	double lambda[16];
	// H += J1^t * Lambda * J2      (accum_JtJ :45-50 / :103-109)
	void accum_JtJ(double *H, const double *J1, int m1, const double *J2, int m2) const {
		if (kind == ORC_NOISE_IDENTITY) {
			for (int a = 0; a < m1; a++)
				for (int b = 0; b < m2; b++) {
					double s = 0;
					for (int k = 0; k < o; k++) s += J1[k * m1 + a] * J2[k * m2 + b];
					H[a * m2 + b] += s;
				}
		} else {
			double T[MAX_P * MAX_OBS];  // J1^t * Lambda  (m1 x o)
			for (int a = 0; a < m1; a++)
				for (int k = 0; k < o; k++) {
					double s = 0;
					for (int l = 0; l < o; l++) s += J1[l * m1 + a] * lambda[l * o + k];
					T[a * o + k] = s;
				}
			for (int a = 0; a < m1; a++)
				for (int b = 0; b < m2; b++) {
					double s = 0;
					for (int k = 0; k < o; k++) s += T[a * o + k] * J2[k * m2 + b];
					H[a * m2 + b] += s;
				}
		}
	}
	void scale_H(double *H, int n) const {  // :52-56 — 1/sigma (sic), not 1/sigma^2
		if (kind == ORC_NOISE_IDENTITY) { const double k = 1.0 / std_noise; for (int i = 0; i < n; i++) H[i] *= k; }
	}
	void accum_Jtr(double *g, const double *J, int m, const double *r) const {  // :60-65 / :120-126
		if (kind == ORC_NOISE_IDENTITY) {
			for (int a = 0; a < m; a++) {
				double s = 0;
				for (int k = 0; k < o; k++) s += J[k * m + a] * r[k];
				g[a] += s;
			}
		} else {
			double T[MAX_P * MAX_OBS];
			for (int a = 0; a < m; a++)
				for (int k = 0; k < o; k++) {
					double s = 0;
					for (int l = 0; l < o; l++) s += J[l * m + a] * lambda[l * o + k];
					T[a * o + k] = s;
				}
			for (int a = 0; a < m; a++) {
				double s = 0;
				for (int k = 0; k < o; k++) s += T[a * o + k] * r[k];
				g[a] += s;
			}
		}
	}
	void scale_Jtr(double *g, int n) const {  // :67-71
		if (kind == ORC_NOISE_IDENTITY) { const double k = 1.0 / std_noise; for (int i = 0; i < n; i++) g[i] *= k; }
	}
};

// impl/sparse_hessian_update_numeric.h:22-60
inline size_t sparse_hessian_update_numeric(SparseBlockHessian &H, const NoiseModel &noise) {
	size_t nInvalid = 0;
	const int m1 = H.nr, m2 = H.nc;
	for (size_t i = 0; i < H.getColCount(); i++) {
		SparseBlockHessian::col_t &col = H.getCol(i);
		for (SparseBlockHessian::col_t::iterator it = col.begin(); it != col.end(); ++it) {
			HEntry &entry = it->second;
			double Hij[MAX_P * MAX_P];
			std::memset(Hij, 0, sizeof(Hij));
			for (size_t k = 0; k < entry.lst_jacob_blocks.size(); k++) {
				const THessianSymbolicInfoEntry &sym_k = entry.lst_jacob_blocks[k];
				if (*sym_k.J1_valid && *sym_k.J2_valid) noise.accum_JtJ(Hij, sym_k.J1, m1, sym_k.J2, m2);
				else nInvalid++;
			}
			noise.scale_H(Hij, m1 * m2);
			std::memcpy(entry.num, Hij, sizeof(double) * m1 * m2);
		}
	}
	return nInvalid;
}

// Minimal view of one Jacobian column for the symbolic builder: ordered (obs_idx -> {num, valid}).
struct JacobBlockRef { const double *num; const char *is_valid; };
typedef std::vector<std::pair<size_t, JacobBlockRef> > JacobColView;  // ascending obs_idx (std::map order)

// impl/sparse_hessian_build_symbolic.h:22-237
inline void sparse_hessian_build_symbolic(SparseBlockHessian &HAp, SparseBlockHessian &Hf, SparseBlockHessian &HApf,
                                          const std::vector<JacobColView> &dh_dAp, const std::vector<JacobColView> &dh_df,
                                          int p, int f) {
	const size_t nUnknowns_k2k = dh_dAp.size(), nUnknowns_k2f = dh_df.size();
	HAp.nr = p; HAp.nc = p; Hf.nr = f; Hf.nc = f; HApf.nr = p; HApf.nc = f;
	struct Isect {
		static void run(const JacobColView &ci, const JacobColView &cj, std::vector<THessianSymbolicInfoEntry> &out) {
			size_t a = 0, b = 0;
			while (a < ci.size() && b < cj.size()) {
				if (ci[a].first < cj[b].first) ++a;
				else if (cj[b].first < ci[a].first) ++b;
				else {
					THessianSymbolicInfoEntry e = {ci[a].second.num, cj[b].second.num, ci[a].second.is_valid, cj[b].second.is_valid, ci[a].first};
					out.push_back(e);
					++a; ++b;
				}
			}
		}
	};
	// (1) HAp  (:38-105): upper triangle, block (i,j) i<=j stored at col j, row i.
	HAp.setColCount(nUnknowns_k2k);
	for (size_t i = 0; i < nUnknowns_k2k; i++) {
		const JacobColView &col_i = dh_dAp[i];
		ORC_ASSERT(!col_i.empty());
		{
			std::vector<THessianSymbolicInfoEntry> &lst = HAp.getCol(i)[i].lst_jacob_blocks;
			for (size_t k = 0; k < col_i.size(); k++) {
				THessianSymbolicInfoEntry e = {col_i[k].second.num, col_i[k].second.num, col_i[k].second.is_valid, col_i[k].second.is_valid, col_i[k].first};
				lst.push_back(e);
			}
		}
		for (size_t j = i + 1; j < nUnknowns_k2k; j++) {
			ORC_ASSERT(!dh_dAp[j].empty());
			std::vector<THessianSymbolicInfoEntry> sym;
			Isect::run(col_i, dh_dAp[j], sym);
			if (!sym.empty()) sym.swap(HAp.getCol(j)[i].lst_jacob_blocks);
		}
	}
	// (2) Hf  (:110-177)
	Hf.setColCount(nUnknowns_k2f);
	for (size_t i = 0; i < nUnknowns_k2f; i++) {
		const JacobColView &col_i = dh_df[i];
		ORC_ASSERT(!col_i.empty());
		{
			std::vector<THessianSymbolicInfoEntry> &lst = Hf.getCol(i)[i].lst_jacob_blocks;
			for (size_t k = 0; k < col_i.size(); k++) {
				THessianSymbolicInfoEntry e = {col_i[k].second.num, col_i[k].second.num, col_i[k].second.is_valid, col_i[k].second.is_valid, col_i[k].first};
				lst.push_back(e);
			}
		}
		for (size_t j = i + 1; j < nUnknowns_k2f; j++) {
			ORC_ASSERT(!dh_df[j].empty());
			std::vector<THessianSymbolicInfoEntry> sym;
			Isect::run(col_i, dh_df[j], sym);
			if (!sym.empty()) sym.swap(Hf.getCol(j)[i].lst_jacob_blocks);
		}
	}
	// (3) HApf (:182-235): stored ROW-compressed: getCol(i=edge)[j=landmark].
	HApf.setColCount(nUnknowns_k2k);
	for (size_t i = 0; i < nUnknowns_k2k; i++) {
		const JacobColView &col_i = dh_dAp[i];
		ORC_ASSERT(!col_i.empty());
		for (size_t j = 0; j < nUnknowns_k2f; j++) {
			ORC_ASSERT(!dh_df[j].empty());
			std::vector<THessianSymbolicInfoEntry> sym;
			Isect::run(col_i, dh_df[j], sym);
			if (!sym.empty()) sym.swap(HApf.getCol(i)[j].lst_jacob_blocks);
		}
	}
}

// ---------------------------------------------------------------------------------------------
// SchurComplement  (impl/schur.h:14-414)
// ---------------------------------------------------------------------------------------------
class SchurComplement {
  public:
	SchurComplement(SparseBlockHessian &_HAp, SparseBlockHessian &_Hf, SparseBlockHessian &_HApf, double *_minus_grad_Ap,
	                double *_minus_grad_f, int p_, int f_)
	    : HAp(_HAp), Hf(_Hf), HApf(_HApf), minus_grad_Ap(_minus_grad_Ap), minus_grad_f(_minus_grad_f),
	      nUnknowns_Ap(_HAp.getColCount()), nUnknowns_f(_Hf.getColCount()), nHf_invertible_blocks(0), p(p_), f(f_) {
		if (!nUnknowns_f || !nUnknowns_Ap) return;  // :34
		HAp_original.copyNumericalValuesFrom(HAp);   // :38
		m_Hf_blocks_info.resize(nUnknowns_f);        // :42-49
		for (size_t i = 0; i < nUnknowns_f; i++) {
			const SparseBlockHessian::col_t &col_i = Hf.getCol(i);
			ORC_ASSERT(!col_i.empty() && col_i.rbegin()->first == i);
			m_Hf_blocks_info[i].sym_Hf_diag_blocks = col_i.rbegin()->second.num;
		}
		m_sym_HAp_reduce.clear();
		m_sym_GradAp_reduce.resize(nUnknowns_Ap);
		for (size_t i = 0; i < nUnknowns_Ap; i++) {  // :56-157
			SparseBlockHessian::col_t &HAp_col_i = HAp.getCol(i);
			for (size_t j = 0; j <= i; j++) {
				THApSymbolicEntry sym_ij;
				TGradApSymbolicEntry *grad_entries = (i == j) ? &m_sym_GradAp_reduce[i] : NULL;
				SparseBlockHessian::col_t &row_i = HApf.getCol(i);
				SparseBlockHessian::col_t &row_j = HApf.getCol(j);
				SparseBlockHessian::col_t::const_iterator it_i = row_i.begin(), it_j = row_j.begin();
				while (it_i != row_i.end() && it_j != row_j.end()) {
					if (it_i->first < it_j->first) ++it_i;
					else if (it_j->first < it_i->first) ++it_j;
					else {
						const size_t idx_feat = it_j->first;
						ORC_ASSERT(idx_feat < nUnknowns_f);
						double *out_temporary_result = NULL;
						if (grad_entries) {
							grad_entries->lst_terms_to_subtract.resize(grad_entries->lst_terms_to_subtract.size() + 1);
							TGradApSymbolicEntry::TEntry &ent = grad_entries->lst_terms_to_subtract.back();
							ent.feat_idx = idx_feat;
							out_temporary_result = ent.Hpi_lk_times_inv_Hf_lk;
						}
						THApSymbolicEntry::TEntry e = {it_j->second.num, &m_Hf_blocks_info[idx_feat], it_i->second.num, out_temporary_result};
						sym_ij.lst_terms_to_add.push_back(e);
						++it_i; ++it_j;
					}
				}
				if (!sym_ij.lst_terms_to_add.empty()) {  // :123-154
					SparseBlockHessian::col_t::iterator itExistingRowEntry;
					if (((i == j) || (i != j && HAp_col_i.size() != 1)) && (itExistingRowEntry = HAp_col_i.find(j)) != HAp_col_i.end()) {
						sym_ij.HAp_ij = itExistingRowEntry->second.num;
					} else {
						ORC_ASSERT(i != j);
						sym_ij.HAp_ij = HAp_col_i[j].num;
						std::memset(sym_ij.HAp_ij, 0, sizeof(double) * MAX_P * MAX_P);
					}
					m_sym_HAp_reduce.push_back(sym_ij);
				}
			}
		}
	}
	void realize_HAp_changed() { HAp_original.copyNumericalValuesFrom(HAp); }  // :165-168
	size_t getNumFeatures() const { return nUnknowns_f; }
	size_t getNumFeaturesFullRank() const { return nHf_invertible_blocks; }

	void numeric_build_reduced_system(const double lambda) {  // :180-268
		if (!nUnknowns_f || !nUnknowns_Ap) return;
		HAp.copyNumericalValuesFrom(HAp_original);
		nHf_invertible_blocks = 0;
		for (size_t i = 0; i < nUnknowns_f; i++) {  // :193-208
			double Hfi[9];
			std::memcpy(Hfi, m_Hf_blocks_info[i].sym_Hf_diag_blocks, sizeof(Hfi));
			for (int k = 0; k < f; k++) Hfi[k * f + k] += lambda;
			double inv[9];
			if ((m_Hf_blocks_info[i].num_Hf_diag_blocks_invertible = fullpiv_lu_inverse3(Hfi, inv))) {
				nHf_invertible_blocks++;
				std::memcpy(m_Hf_blocks_info[i].num_Hf_diag_blocks_inverses, inv, sizeof(inv));
			}
		}
		double aux[MAX_P * LM_DIMS];
		for (std::deque<THApSymbolicEntry>::const_iterator it = m_sym_HAp_reduce.begin(); it != m_sym_HAp_reduce.end(); ++it) {  // :213-244
			double *HAp_ij = it->HAp_ij;
			for (size_t t = 0; t < it->lst_terms_to_add.size(); t++) {
				const THApSymbolicEntry::TEntry &entry = it->lst_terms_to_add[t];
				if (!entry.inv_Hf_lk->num_Hf_diag_blocks_invertible) continue;
				double *prod = entry.out_Hpi_lk_times_inv_Hf_lk ? entry.out_Hpi_lk_times_inv_Hf_lk : aux;
				const double *inv = entry.inv_Hf_lk->num_Hf_diag_blocks_inverses;
				for (int a = 0; a < p; a++)
					for (int b = 0; b < f; b++) {
						double s = 0;
						for (int k = 0; k < f; k++) s += entry.Hpi_lk[a * f + k] * inv[k * f + b];
						prod[a * f + b] = s;
					}
				for (int a = 0; a < p; a++)
					for (int b = 0; b < p; b++) {
						double s = 0;
						for (int k = 0; k < f; k++) s += prod[a * f + k] * entry.Hpj_lk[b * f + k];
						HAp_ij[a * p + b] -= s;
					}
			}
		}
		double *modified_grad_Ap = minus_grad_Ap;  // :248-265 — IN PLACE on the caller's gradient
		for (size_t i = 0; i < nUnknowns_Ap; i++) {
			for (std::deque<TGradApSymbolicEntry::TEntry>::const_iterator it = m_sym_GradAp_reduce[i].lst_terms_to_subtract.begin();
			     it != m_sym_GradAp_reduce[i].lst_terms_to_subtract.end(); ++it) {
				if (m_Hf_blocks_info[it->feat_idx].num_Hf_diag_blocks_invertible) {
					const double *grad_df = minus_grad_f + it->feat_idx * f;
					for (int a = 0; a < p; a++) {
						double s = 0;
						for (int k = 0; k < f; k++) s += it->Hpi_lk_times_inv_Hf_lk[a * f + k] * grad_df[k];
						modified_grad_Ap[a] -= s;
					}
				}
			}
			modified_grad_Ap += p;
		}
	}

	void numeric_solve_for_features(double *in_deltas_Ap, double *out_deltas_feats) {  // :271-311
		for (size_t idx_Ap = 0; idx_Ap < nUnknowns_Ap; idx_Ap++) {
			SparseBlockHessian::col_t &row_i = HApf.getCol(idx_Ap);
			const double *delta_idx_Ap = in_deltas_Ap + idx_Ap * p;
			for (SparseBlockHessian::col_t::const_iterator itCol = row_i.begin(); itCol != row_i.end(); ++itCol) {
				const size_t idx_feat = itCol->first;
				if (!was_ith_feature_invertible(idx_feat)) continue;
				double *grad_df = minus_grad_f + idx_feat * f;
				for (int b = 0; b < f; b++) {
					double s = 0;
					for (int a = 0; a < p; a++) s += itCol->second.num[a * f + b] * delta_idx_Ap[a];
					grad_df[b] -= s;
				}
			}
		}
		for (size_t idx_feat = 0; idx_feat < nUnknowns_f; idx_feat++) {
			if (!was_ith_feature_invertible(idx_feat)) continue;
			double *delta_feat = out_deltas_feats + idx_feat * f;
			const double *grad_df = minus_grad_f + idx_feat * f;
			const double *inv = m_Hf_blocks_info[idx_feat].num_Hf_diag_blocks_inverses;
			for (int a = 0; a < f; a++) {
				double s = 0;
				for (int k = 0; k < f; k++) s += inv[a * f + k] * grad_df[k];
				delta_feat[a] = s;
			}
		}
	}
	bool was_ith_feature_invertible(const size_t i) const { return m_Hf_blocks_info[i].num_Hf_diag_blocks_invertible; }

  private:
	SparseBlockHessian HAp_original;
	SparseBlockHessian &HAp, &Hf, &HApf;
	double *minus_grad_Ap, *minus_grad_f;
	const size_t nUnknowns_Ap, nUnknowns_f;
	size_t nHf_invertible_blocks;
	int p, f;
	struct TInfoPerHfBlock {
		const double *sym_Hf_diag_blocks;
		double num_Hf_diag_blocks_inverses[9];
		bool num_Hf_diag_blocks_invertible;
		TInfoPerHfBlock() : sym_Hf_diag_blocks(NULL), num_Hf_diag_blocks_invertible(false) {}
	};
	std::deque<TInfoPerHfBlock> m_Hf_blocks_info;
	struct THApSymbolicEntry {
		struct TEntry {
			const double *Hpi_lk;
			const TInfoPerHfBlock *inv_Hf_lk;
			const double *Hpj_lk;
			double *out_Hpi_lk_times_inv_Hf_lk;
		};
		double *HAp_ij;
		std::deque<TEntry> lst_terms_to_add;
	};
	std::deque<THApSymbolicEntry> m_sym_HAp_reduce;
	struct TGradApSymbolicEntry {
		struct TEntry {
			size_t feat_idx;
			double Hpi_lk_times_inv_Hf_lk[MAX_P * LM_DIMS];
		};
		std::deque<TEntry> lst_terms_to_subtract;
	};
	std::vector<TGradApSymbolicEntry> m_sym_GradAp_reduce;
};

// ---------------------------------------------------------------------------------------------
// mrpt::math::CSparseMatrix + CSparseMatrix::CholeskyDecomp as the reference uses them (impl/lev-marq_solvers.h:84-184,
// 296-362).  MRPT wraps T. Davis' CSparse (bundled inside MRPT 1.3.x, absent here): insert_submatrix() appends
// triplets, compressFromTriplet() = cs_compress (duplicates summed when the factorisation walks the column),
// CholeskyDecomp(A) = cs_schol(order 1) + cs_chol, update(A) = cs_chol again on the same symbolic analysis (throws
// CExceptionNotDefPos when cs_chol meets a pivot d <= 0), backsub = P^t L^-t L^-1 P b.  Restated here: the up-looking
// Cholesky of cs_chol (elimination tree, row reach, "d <= 0 -> not positive definite") on the UPPER triangle only, in
// the NATURAL order -- cs_schol's AMD permutation is not reproduced (parity unpinned; any symmetric permutation gives
// the same solution up to rounding and fails iff the matrix is not positive definite).
// ---------------------------------------------------------------------------------------------
struct CSparseMatrixLite {
	size_t n;
	std::vector<int> ti, tj;  // triplets (row, col)
	std::vector<double> tx;
	std::vector<int> Ap, Ai;  // compressed columns
	std::vector<double> Ax;
	CSparseMatrixLite() : n(0) {}
	void clear(size_t n_) { n = n_; ti.clear(); tj.clear(); tx.clear(); Ap.clear(); Ai.clear(); Ax.clear(); }
	void insert_submatrix(size_t row, size_t col, const double *M, int nr, int nc) {
		for (int r = 0; r < nr; r++)
			for (int c = 0; c < nc; c++) { ti.push_back((int)(row + r)); tj.push_back((int)(col + c)); tx.push_back(M[r * nc + c]); }
	}
	void compressFromTriplet() {  // cs_compress: counting sort by column, triplet order kept inside a column
		Ap.assign(n + 1, 0);
		for (size_t k = 0; k < tj.size(); k++) Ap[tj[k] + 1]++;
		for (size_t c = 0; c < n; c++) Ap[c + 1] += Ap[c];
		Ai.resize(ti.size()); Ax.resize(ti.size());
		std::vector<int> w(Ap.begin(), Ap.end() - 1);
		for (size_t k = 0; k < ti.size(); k++) { const int q = w[tj[k]]++; Ai[q] = ti[k]; Ax[q] = tx[k]; }
	}
	void getAsDenseUpper(std::vector<double> &D) const {
		D.assign(n * n, 0.0);
		for (size_t c = 0; c < n; c++)
			for (int q = Ap[c]; q < Ap[c + 1]; q++) D[(size_t)Ai[q] * n + c] += Ax[q];
	}
};
struct SparseCholeskyLite {
	size_t n;
	std::vector<int> parent;                                 // elimination tree (cs_etree)
	std::vector<std::vector<std::pair<int, double> > > Lcol; // L by columns, diagonal first, rows ascending
	bool has_symbolic;
	SparseCholeskyLite() : n(0), has_symbolic(false) {}
	void symbolic(const CSparseMatrixLite &A) {
		n = A.n;
		parent.assign(n, -1);
		std::vector<int> ancestor(n, -1);
		for (size_t k = 0; k < n; k++)
			for (int q = A.Ap[k]; q < A.Ap[k + 1]; q++)
				for (int i = A.Ai[q]; i != -1 && i < (int)k;) {  // walk from i to the root of its subtree, compressing the path
					const int inext = ancestor[i];
					ancestor[i] = (int)k;
					if (inext == -1) parent[i] = (int)k;
					i = inext;
				}
		has_symbolic = true;
	}
	// cs_chol: row k of L solves L(0:k-1,0:k-1) x = A(0:k-1,k); d = A(k,k) - x.x must be > 0.
	bool numeric(const CSparseMatrixLite &A) {
		if (!has_symbolic) symbolic(A);
		Lcol.assign(n, std::vector<std::pair<int, double> >());
		std::vector<double> x(n, 0.0);
		std::vector<int> mark(n, -1), reach;
		for (size_t k = 0; k < n; k++) {
			reach.clear();
			mark[k] = (int)k;
			double d = 0;
			for (int q = A.Ap[k]; q < A.Ap[k + 1]; q++) {
				const int i = A.Ai[q];
				if (i > (int)k) continue;  // upper triangle only
				if (i == (int)k) { d += A.Ax[q]; continue; }
				x[i] += A.Ax[q];  // duplicates summed
				for (int t = i; t != -1 && mark[t] != (int)k; t = parent[t]) { mark[t] = (int)k; reach.push_back(t); }
			}
			std::sort(reach.begin(), reach.end());  // ascending = a topological order of the elimination tree
			for (size_t r = 0; r < reach.size(); r++) {
				const int i = reach[r];
				const double lki = x[i] / Lcol[i][0].second;
				x[i] = 0;
				for (size_t q = 1; q < Lcol[i].size(); q++) x[Lcol[i][q].first] -= Lcol[i][q].second * lki;
				d -= lki * lki;
				Lcol[i].push_back(std::make_pair((int)k, lki));
			}
			if (d <= 0) return false;  // not positive definite
			Lcol[k].push_back(std::make_pair((int)k, std::sqrt(d)));
		}
		return true;
	}
	void backsub(const double *b, double *out, size_t len) const {  // L L^t out = b
		std::vector<double> y(b, b + len);
		for (size_t j = 0; j < len; j++) {
			y[j] /= Lcol[j][0].second;
			for (size_t q = 1; q < Lcol[j].size(); q++) y[Lcol[j][q].first] -= Lcol[j][q].second * y[j];
		}
		for (size_t jj = len; jj-- > 0;) {
			for (size_t q = 1; q < Lcol[jj].size(); q++) y[jj] -= Lcol[jj][q].second * y[Lcol[jj][q].first];
			y[jj] /= Lcol[jj][0].second;
		}
		for (size_t j = 0; j < len; j++) out[j] = y[j];
	}
};

// RBA_OPTIONS::solver_t (srba_options_solver.h:21-77) is a compile-time choice in the reference (optimize_edges.h:54);
// here a run-time tag with the same three values.
enum { SOLVER_LM_SCHUR_DENSE_CHOLESKY = 0, SOLVER_LM_SCHUR_SPARSE_CHOLESKY = 1, SOLVER_LM_NO_SCHUR_SPARSE_CHOLESKY = 2 };
struct SolverBase {
	std::vector<double> delta_eps;
	virtual ~SolverBase() {}
	virtual bool solve(const double lambda) = 0;
	virtual void realize_lambda_changed() = 0;
	virtual void realize_relinearized() = 0;
	virtual bool was_ith_feature_invertible(const size_t i) = 0;
	virtual void get_extra_results(std::vector<double> &denseH, bool &valid) = 0;  // hessian of the last solve, densified
};

// solver_engine<false,false>  (impl/lev-marq_solvers.h:27-209): LM without Schur, sparse Cholesky of the WHOLE system.
// minus_grad is left untouched (unlike the Schur variants) and every landmark counts as invertible (:196-200).
struct SolverNoSchurSparse : public SolverBase {
	const size_t nUnknowns_k2k, nUnknowns_k2f;
	const int p, f;
	SparseBlockHessian &HAp, &Hf, &HApf;
	std::vector<double> &minus_grad;
	const size_t nUnknowns_scalars, idx_start_f;
	CSparseMatrixLite sS;
	bool sS_is_valid;
	SparseCholeskyLite chol;
	SolverNoSchurSparse(SparseBlockHessian &HAp_, SparseBlockHessian &Hf_, SparseBlockHessian &HApf_, std::vector<double> &minus_grad_, size_t n_k2k,
	                    size_t n_k2f, int p_, int f_)
	    : nUnknowns_k2k(n_k2k), nUnknowns_k2f(n_k2f), p(p_), f(f_), HAp(HAp_), Hf(Hf_), HApf(HApf_), minus_grad(minus_grad_),
	      nUnknowns_scalars(p_ * n_k2k + f_ * n_k2f), idx_start_f(p_ * n_k2k), sS_is_valid(false) {}
	void insert_diag_or_plain(SparseBlockHessian &H, size_t ncols, size_t off, int d, double lambda) {
		double tmp[MAX_P * MAX_P];
		for (size_t i = 0; i < ncols; i++) {  // only the upper-half triangle is stored (:93-116,135-158)
			const SparseBlockHessian::col_t &col_i = H.getCol(i);
			for (SparseBlockHessian::col_t::const_iterator it = col_i.begin(); it != col_i.end(); ++it) {
				for (int k = 0; k < d * d; k++) tmp[k] = it->second.num[k];
				if (it->first == i)
					for (int k = 0; k < d; k++) tmp[k * d + k] += lambda;
				sS.insert_submatrix(off + d * it->first, off + d * i, tmp, d, d);
			}
		}
	}
	bool solve(const double lambda) {  // :81-184
		sS.clear(nUnknowns_scalars);
		sS_is_valid = false;
		insert_diag_or_plain(HAp, nUnknowns_k2k, 0, p, lambda);  // 1/3
		for (size_t i = 0; i < nUnknowns_k2k; i++) {              // 2/3: HApf is row-compressed (:119-131)
			const SparseBlockHessian::col_t &row_i = HApf.getCol(i);
			for (SparseBlockHessian::col_t::const_iterator it = row_i.begin(); it != row_i.end(); ++it)
				sS.insert_submatrix(p * i, idx_start_f + f * it->first, it->second.num, p, f);
		}
		insert_diag_or_plain(Hf, nUnknowns_k2f, idx_start_f, f, lambda);  // 3/3
		sS.compressFromTriplet();
		if (!chol.numeric(sS)) return false;  // CExceptionNotDefPos (:171-175)
		sS_is_valid = true;
		delta_eps.assign(nUnknowns_scalars, 0.0);
		chol.backsub(&minus_grad[0], &delta_eps[0], nUnknowns_scalars);  // :182
		return true;
	}
	void realize_relinearized() {}
	void realize_lambda_changed() {}
	bool was_ith_feature_invertible(const size_t) { return true; }
	void get_extra_results(std::vector<double> &denseH, bool &valid) { valid = sS_is_valid; if (sS.Ap.empty()) denseH.clear(); else sS.getAsDenseUpper(denseH); }
};

// solver_engine<true,true>  (impl/lev-marq_solvers.h:409-591): LM with Schur + dense Cholesky.
struct SolverSchurDense : public SolverBase {
	const size_t nUnknowns_k2k, nUnknowns_k2f;
	const int p, f;
	SparseBlockHessian &HAp, &Hf, &HApf;
	std::vector<double> &minus_grad;
	SchurComplement schur_compl;
	std::vector<double> denseH, denseChol;
	bool denseChol_is_uptodate, hessian_is_valid;

	SolverSchurDense(SparseBlockHessian &HAp_, SparseBlockHessian &Hf_, SparseBlockHessian &HApf_, std::vector<double> &minus_grad_,
	                 size_t n_k2k, size_t n_k2f, int p_, int f_)
	    : nUnknowns_k2k(n_k2k), nUnknowns_k2f(n_k2f), p(p_), f(f_), HAp(HAp_), Hf(Hf_), HApf(HApf_), minus_grad(minus_grad_),
	      schur_compl(HAp_, Hf_, HApf_, &minus_grad_[0], n_k2f != 0 ? &minus_grad_[p_ * n_k2k] : NULL, p_, f_),
	      denseChol_is_uptodate(false), hessian_is_valid(false) {}

	bool solve(const double lambda) {  // :474-568
		schur_compl.numeric_build_reduced_system(lambda);
		const size_t n = nUnknowns_k2k * p;
		if (!denseChol_is_uptodate) {
			denseH.assign(n * n, 0.0);
			hessian_is_valid = false;
			for (size_t i = 0; i < nUnknowns_k2k; i++) {  // :497-519
				const SparseBlockHessian::col_t &col_i = HAp.getCol(i);
				for (SparseBlockHessian::col_t::const_iterator it = col_i.begin(); it != col_i.end(); ++it) {
					const size_t r0 = p * it->first, c0 = p * i;
					for (int a = 0; a < p; a++)
						for (int b = 0; b < p; b++) denseH[(r0 + a) * n + c0 + b] = it->second.num[a * p + b] + ((it->first == i && a == b) ? lambda : 0.0);
				}
			}
			hessian_is_valid = true;
			denseChol = denseH;
			if (!dense_llt_upper(denseChol, n)) return false;  // :527-534
			denseChol_is_uptodate = true;
		}
		delta_eps.assign(nUnknowns_k2k * p + nUnknowns_k2f * f, 0.0);  // :544-549
		if (n) dense_llt_solve(denseChol, n, &minus_grad[0], &delta_eps[0]);
		if (nUnknowns_k2f != 0 && nUnknowns_k2k != 0)  // the reference dereferences an empty vector when n_p==0 (schur.h:34,288)
			schur_compl.numeric_solve_for_features(&delta_eps[0], &delta_eps[nUnknowns_k2k * p]);  // :556-561
		return true;
	}
	void realize_lambda_changed() { denseChol_is_uptodate = false; }   // :570-573
	void realize_relinearized() { schur_compl.realize_HAp_changed(); }  // :574-580
	bool was_ith_feature_invertible(const size_t i) { return schur_compl.was_ith_feature_invertible(i); }
	void get_extra_results(std::vector<double> &out, bool &valid) { out.swap(denseH); valid = hessian_is_valid; }  // :582-586
};

// solver_engine<true,false>  (impl/lev-marq_solvers.h:213-405): LM with Schur + sparse Cholesky of the reduced system.
struct SolverSchurSparse : public SolverBase {
	const size_t nUnknowns_k2k, nUnknowns_k2f;
	const int p, f;
	SparseBlockHessian &HAp, &Hf, &HApf;
	std::vector<double> &minus_grad;
	CSparseMatrixLite sS;
	bool sS_is_valid;
	SparseCholeskyLite chol;
	SchurComplement schur_compl;
	SolverSchurSparse(SparseBlockHessian &HAp_, SparseBlockHessian &Hf_, SparseBlockHessian &HApf_, std::vector<double> &minus_grad_, size_t n_k2k,
	                  size_t n_k2f, int p_, int f_)
	    : nUnknowns_k2k(n_k2k), nUnknowns_k2f(n_k2f), p(p_), f(f_), HAp(HAp_), Hf(Hf_), HApf(HApf_), minus_grad(minus_grad_), sS_is_valid(false),
	      schur_compl(HAp_, Hf_, HApf_, &minus_grad_[0], n_k2f != 0 ? &minus_grad_[p_ * n_k2k] : NULL, p_, f_) {}
	bool solve(const double lambda) {  // :276-380
		schur_compl.numeric_build_reduced_system(lambda);
		const size_t n = nUnknowns_k2k * p;
		sS.clear(n);
		sS_is_valid = false;
		double tmp[MAX_P * MAX_P];
		for (size_t i = 0; i < nUnknowns_k2k; i++) {  // :296-321
			const SparseBlockHessian::col_t &col_i = HAp.getCol(i);
			for (SparseBlockHessian::col_t::const_iterator it = col_i.begin(); it != col_i.end(); ++it) {
				for (int k = 0; k < p * p; k++) tmp[k] = it->second.num[k];
				if (it->first == i)
					for (int k = 0; k < p; k++) tmp[k * p + k] += lambda;
				sS.insert_submatrix(p * it->first, p * i, tmp, p, p);
			}
		}
		sS.compressFromTriplet();
		if (!chol.numeric(sS)) return false;  // :347-351
		sS_is_valid = true;
		delta_eps.assign(nUnknowns_k2k * p + nUnknowns_k2f * f, 0.0);  // :359 "Important: It's initialized to zeros!"
		if (n) chol.backsub(&minus_grad[0], &delta_eps[0], n);
		if (nUnknowns_k2f != 0 && nUnknowns_k2k != 0) schur_compl.numeric_solve_for_features(&delta_eps[0], &delta_eps[n]);  // :371-375
		return true;
	}
	void realize_relinearized() { schur_compl.realize_HAp_changed(); }
	void realize_lambda_changed() {}
	bool was_ith_feature_invertible(const size_t i) { return schur_compl.was_ith_feature_invertible(i); }
	void get_extra_results(std::vector<double> &denseH, bool &valid) { valid = sS_is_valid; if (sS.Ap.empty()) denseH.clear(); else sS.getAsDenseUpper(denseH); }
};

// ---------------------------------------------------------------------------------------------
// Trace of one optimize_edges call (parity debugging; not in the reference)
// ---------------------------------------------------------------------------------------------
struct OptTrace {
	std::vector<uint64_t> k2k_ids, lm_ids, obs_idx;
	std::vector<double> trial_scalars;  // 6 per trial
	std::vector<double> deltas;         // n_scalars per trial
	std::vector<double> init_grad;
	double lambda0;
	size_t n_scalars;
	OptTrace() : lambda0(0), n_scalars(0) {}
};

// ---------------------------------------------------------------------------------------------
// The engine:  RbaEngine<KF2KF_POSE, LM, OBS, RBA_OPTIONS>  (RbaEngine.h:66-816)
// POSE = Pose2 (kf2kf_poses::SE2) or Pose3 (kf2kf_poses::SE3); landmark/observation/sensor-pose/
// noise types are run-time tags in orc_params_t.
// ---------------------------------------------------------------------------------------------
template <class POSE>
struct Engine {
	typedef POSE pose_t;
	static const int P = POSE::DIMS;

	struct pose_flag_t {  // srba_types.h:57-68
		pose_t pose;
		mutable bool updated;
		pose_flag_t() : updated(false) {}
		void mark_outdated() const { updated = false; }
	};
	typedef std::map<TKeyFrameID, pose_flag_t> frameid2pose_map_t;
	struct k2k_edge_t {  // srba_types.h:82-91
		TKeyFrameID from, to;
		pose_t inv_pose;
		size_t id;
	};
	typedef std::deque<k2k_edge_t *> k2k_edge_vector_t;
	struct TRelativeLandmarkPos {  // srba_types.h:106-121
		TKeyFrameID id_frame_base;
		double pos[LM_DIMS];
	};
	typedef std::map<TLandmarkID, TRelativeLandmarkPos> TRelativeLandmarkPosMap;
	struct TLandmarkEntry {  // srba_types.h:127-135
		bool has_known_pos;
		TRelativeLandmarkPos *rfp;
		TLandmarkEntry() : has_known_pos(true), rfp(NULL) {}
		TLandmarkEntry(bool k, TRelativeLandmarkPos *r) : has_known_pos(k), rfp(r) {}
	};
	struct k2f_edge_t {  // srba_types.h:501-525
		TKeyFrameID kf_id;
		TLandmarkID feat_id;
		double obs_data[MAX_OBS];  // obs.obs.obs_data (after the float rounding of TPixelCoordf for pixel observations)
		double obs_arr[MAX_OBS];
		bool feat_has_known_rel_pos, is_first_obs_of_unknown;
		TRelativeLandmarkPos *feat_rel_pos;
	};
	struct keyframe_info {  // srba_types.h:528-532
		std::deque<k2k_edge_t *> adjacent_k2k_edges;
		std::deque<k2f_edge_t *> adjacent_k2f_edges;
	};
	struct TSpanTreeEntry { TKeyFrameID next; topo_dist_t distance; };  // srba_types.h:538-542
	struct TJacobianSymbolicInfo_dh_dAp {  // srba_types.h:204-240
		const pose_flag_t *rel_pose_d1_from_obs, *rel_pose_base_from_d1;
		bool edge_normal_dir;
		const TRelativeLandmarkPos *feat_rel_pos;
		TKeyFrameID kf_d, kf_base;
		size_t k2k_edge_id, obs_idx;
		char *is_valid;
	};
	struct TJacobianSymbolicInfo_dh_df {  // srba_types.h:246-266
		TRelativeLandmarkPos *feat_rel_pos;
		const pose_flag_t *rel_pose_base_from_obs;
		size_t obs_idx;
		char *is_valid;
	};
	struct EntryAp { double num[MAX_OBS * MAX_P]; TJacobianSymbolicInfo_dh_dAp sym; EntryAp() { std::memset(num, 0, sizeof(num)); } };
	struct EntryF { double num[MAX_OBS * LM_DIMS]; TJacobianSymbolicInfo_dh_df sym; EntryF() { std::memset(num, 0, sizeof(num)); } };
	typedef std::map<size_t, EntryAp> colAp_t;
	typedef std::map<size_t, EntryF> colF_t;
	struct new_kf_observation_t {  // srba_types.h:474-497
		TLandmarkID feat_id;
		double obs_data[MAX_OBS];
		bool is_fixed, is_unknown_with_init_val;
		double feat_rel_pos[LM_DIMS];
	};
	typedef std::deque<new_kf_observation_t> new_kf_observations_t;
	struct TNewEdgeInfo {  // srba_types.h:178-200
		size_t id;
		bool has_approx_init_val;
		TKeyFrameID loopclosure_observer_kf, loopclosure_base_kf;
		TNewEdgeInfo() : id(SRBA_INVALID_INDEX), has_approx_init_val(false), loopclosure_observer_kf(SRBA_INVALID_KEYFRAMEID), loopclosure_base_kf(SRBA_INVALID_KEYFRAMEID) {}
	};
	struct TOptimizeLocalAreaParams {  // RbaEngine.h:214-227
		bool optimize_k2k_edges, optimize_landmarks;
		TKeyFrameID max_visitable_kf_id;
		size_t dont_optimize_landmarks_seen_less_than_n_times;
		TOptimizeLocalAreaParams() : optimize_k2k_edges(true), optimize_landmarks(true), max_visitable_kf_id(static_cast<TKeyFrameID>(-1)), dont_optimize_landmarks_seen_less_than_n_times(2) {}
	};
	struct TNewKeyFrameInfo {  // RbaEngine.h:180-195
		TKeyFrameID kf_id;
		std::vector<TNewEdgeInfo> created_edge_ids;
		orc_opt_info_t optimize_results, optimize_results_stg1;
	};

	// ------------------------------------------------------------------ TRBA_Problem_state (srba_types.h:548-785)
	std::deque<keyframe_info> keyframes;
	std::deque<k2k_edge_t> k2k_edges;
	TRelativeLandmarkPosMap unknown_lms, known_lms;
	std::map<TLandmarkID, std::vector<double> > unknown_lms_inf_matrices;
	std::deque<TLandmarkEntry> all_lms;
	struct TSpanningTree {
		struct {
			MapAsVector<std::map<TKeyFrameID, TSpanTreeEntry> > next_edge;
			MapAsVector<std::map<TKeyFrameID, k2k_edge_vector_t> > all_edges;
		} sym;
		MapAsVector<frameid2pose_map_t> num;
	} spanning_tree;
	std::deque<k2f_edge_t> all_observations;
	struct {
		std::deque<colAp_t> dh_dAp;
		std::deque<colF_t> dh_df;
		std::vector<size_t> dh_df_remap;  // feat_id -> column (getColInverseRemappedIndices)
	} lin_system;
	std::deque<char> all_observations_Jacob_validity;
	std::set<size_t> last_timestep_touched_kfs;

	orc_params_t parameters;
	pose_t sensor_pose_rel;   // parameters.sensor_pose.relative_pose (as Pose3 when sensor_pose_kind==SE3)
	Pose3 sensor_pose3;
	Pose3 right_cam_R2L;      // -rightCameraPose, hoisted (models/sensors.h:193,259)
	NoiseModel noise;
	int O;                    // OBS_DIMS
	const pose_t aux_null_pose;

	// instrumentation (not in the reference)
	bool trace_on;
	std::deque<OptTrace> traces;
	double total_opt_seconds;
	uint64_t total_opt_trials;

	explicit Engine(const orc_params_t &p) : trace_on(false), total_opt_seconds(0), total_opt_trials(0) { set_params(p); }

	void set_params(const orc_params_t &p) {
		parameters = p;
		O = obs_dims_of(p.obs_kind);
		noise.kind = p.noise_kind; noise.o = O; noise.std_noise = p.noise_std;
		std::memcpy(noise.lambda, p.noise_lambda, sizeof(noise.lambda));
		sensor_pose3 = Pose3(p.sensor_pose[0], p.sensor_pose[1], p.sensor_pose[2], p.sensor_pose[3], p.sensor_pose[4], p.sensor_pose[5]);
		Pose3 rc = pose3_from_quat(p.right_cam_pose[0], p.right_cam_pose[1], p.right_cam_pose[2], p.right_cam_pose[3], p.right_cam_pose[4], p.right_cam_pose[5], p.right_cam_pose[6]);
		right_cam_R2L = -rc;
		if (P == 3) ORC_ASSERT(p.obs_kind == ORC_OBS_RELPOSES2D && p.sensor_pose_kind == ORC_SENSORPOSE_NONE);
		if (P == 6) ORC_ASSERT(p.obs_kind != ORC_OBS_RELPOSES2D);
	}
	bool is_relpose_family() const { return parameters.obs_kind == ORC_OBS_RELPOSES2D; }  // landmarks.h:100 jacob_relpose_landmark

	// ------------------------------------------------------------------ sensor models (models/sensors.h)
	// landmark_t::composePosePoint (landmarks.h:42-46 Euclidean3D; :108-112 RelativePoses2D = no-op)
	void composePosePoint(double *pt, const Pose3 &pose) const { pose.composePoint(pt[0], pt[1], pt[2], pt[0], pt[1], pt[2]); }
	void composePosePoint(double *, const Pose2 &) const {}
	// sensor_pose_on_robot_t::point_robot2sensor (srba_options_sensor_pose.h:58-63,116-121)
	void point_robot2sensor(double *pt) const {
		if (parameters.sensor_pose_kind == ORC_SENSORPOSE_SE3) sensor_pose3.inverseComposePoint(pt[0], pt[1], pt[2], pt[0], pt[1], pt[2]);
	}
	// sensor_model<>::eval_jacob_dh_dx  (sensors.h:85-110 Mono, :230-280 Stereo, :383-393 Cartesian3D, :804-814 RelPoses2D)
	bool eval_jacob_dh_dx(double *dh_dx /* O x 3 */, const double *xji_l) const {
		switch (parameters.obs_kind) {
			case ORC_OBS_RELPOSES2D:
			case ORC_OBS_CARTESIAN3D:
				for (int i = 0; i < 9; i++) dh_dx[i] = 0;
				dh_dx[0] = dh_dx[4] = dh_dx[8] = 1;
				return true;
			case ORC_OBS_MONO: {
				if (xji_l[2] <= 0) return false;
				const double pz_inv = 1.0 / xji_l[2], pz_inv2 = pz_inv * pz_inv;
				const double fx = parameters.cam_left[0], fy = parameters.cam_left[1];
				dh_dx[0] = fx * pz_inv; dh_dx[1] = 0; dh_dx[2] = -fx * xji_l[0] * pz_inv2;
				dh_dx[3] = 0; dh_dx[4] = fy * pz_inv; dh_dx[5] = -fy * xji_l[1] * pz_inv2;
				return true;
			}
			case ORC_OBS_STEREO: {
				if (xji_l[2] <= 0) return false;
				{
					const double pz_inv = 1.0 / xji_l[2], pz_inv2 = pz_inv * pz_inv;
					const double fx = parameters.cam_left[0], fy = parameters.cam_left[1];
					dh_dx[0] = fx * pz_inv; dh_dx[1] = 0; dh_dx[2] = -fx * xji_l[0] * pz_inv2;
					dh_dx[3] = 0; dh_dx[4] = fy * pz_inv; dh_dx[5] = -fy * xji_l[1] * pz_inv2;
				}
				double r[3];
				right_cam_R2L.composePoint(xji_l[0], xji_l[1], xji_l[2], r[0], r[1], r[2]);
				{
					const double pz_inv = 1.0 / r[2], pz_inv2 = pz_inv * pz_inv;
					const double fx = parameters.cam_right[0], fy = parameters.cam_right[1];
					dh_dx[6] = fx * pz_inv; dh_dx[7] = 0; dh_dx[8] = -fx * r[0] * pz_inv2;
					dh_dx[9] = 0; dh_dx[10] = fy * pz_inv; dh_dx[11] = -fy * r[1] * pz_inv2;
				}
				return true;
			}
		}
		return false;
	}
	// jacob_dh_dx_rotate (srba_options_sensor_pose.h:66-71,124-129): dh_dx = dh_dx * R_s^T
	void jacob_dh_dx_rotate(double *dh_dx) const {
		if (parameters.sensor_pose_kind != ORC_SENSORPOSE_SE3) return;
		double t[MAX_OBS * 3];
		for (int r = 0; r < O; r++)
			for (int c = 0; c < 3; c++)
				t[3 * r + c] = dh_dx[3 * r + 0] * sensor_pose3.R[3 * c + 0] + dh_dx[3 * r + 1] * sensor_pose3.R[3 * c + 1] + dh_dx[3 * r + 2] * sensor_pose3.R[3 * c + 2];
		std::memcpy(dh_dx, t, sizeof(double) * O * 3);
	}
	// pose_robot2sensor + sensor_model<>::observe_error (reprojection_residuals.h:54-63; sensors.h:49-66,174-211,348-364,769-785)
	void observe_error(double *out_err, const double *z, const Pose2 &base_pose_wrt_observer, const double *) const {
		const Pose2 h = Pose2(z[0], z[1], z[2]) - base_pose_wrt_observer;
		out_err[0] = h.x; out_err[1] = h.y; out_err[2] = h.phi;
	}
	void observe_error(double *out_err, const double *z, const Pose3 &base_pose_wrt_robot, const double *lm_pos) const {
		Pose3 base_pose_wrt_sensor = base_pose_wrt_robot;
		if (parameters.sensor_pose_kind == ORC_SENSORPOSE_SE3) base_pose_wrt_sensor = base_pose_wrt_robot - sensor_pose3;  // inverseComposeFrom
		double x, y, zc;
		base_pose_wrt_sensor.composePoint(lm_pos[0], lm_pos[1], lm_pos[2], x, y, zc);
		double pred[MAX_OBS];
		switch (parameters.obs_kind) {
			case ORC_OBS_CARTESIAN3D: pred[0] = x; pred[1] = y; pred[2] = zc; break;
			case ORC_OBS_MONO:
				ORC_ASSERT(zc != 0);
				pred[0] = parameters.cam_left[2] + parameters.cam_left[0] * x / zc;
				pred[1] = parameters.cam_left[3] + parameters.cam_left[1] * y / zc;
				break;
			case ORC_OBS_STEREO: {
				ORC_ASSERT(zc != 0);
				pred[0] = parameters.cam_left[2] + parameters.cam_left[0] * x / zc;
				pred[1] = parameters.cam_left[3] + parameters.cam_left[1] * y / zc;
				double rx, ry, rz;
				right_cam_R2L.composePoint(x, y, zc, rx, ry, rz);
				ORC_ASSERT(rz != 0);
				pred[2] = parameters.cam_right[2] + parameters.cam_right[0] * rx / rz;
				pred[3] = parameters.cam_right[3] + parameters.cam_right[1] * ry / rz;
				break;
			}
			default: ORC_ASSERT(false);
		}
		for (int i = 0; i < O; i++) out_err[i] = z[i] - pred[i];
	}
	// sensor_model<>::inverse_sensor_model (sensors.h:122-141,293-315,405-413,826-834) + sensor2robot_point
	void inverse_sensor_model(double *out_lm_pos, const double *obs) const {
		switch (parameters.obs_kind) {
			case ORC_OBS_RELPOSES2D:
			case ORC_OBS_CARTESIAN3D: out_lm_pos[0] = obs[0]; out_lm_pos[1] = obs[1]; out_lm_pos[2] = obs[2]; break;
			case ORC_OBS_MONO:
				out_lm_pos[0] = (obs[0] - parameters.cam_left[2]) / parameters.cam_left[0];
				out_lm_pos[1] = (obs[1] - parameters.cam_left[3]) / parameters.cam_left[1];
				out_lm_pos[2] = 1;
				break;
			case ORC_OBS_STEREO: {
				const double fxl = parameters.cam_left[0], fyl = parameters.cam_left[1], cxl = parameters.cam_left[2], cyl = parameters.cam_left[3];
				const double disparity = std::max(0.001f, static_cast<float>(obs[0]) - static_cast<float>(obs[2]));
				const double baseline = parameters.right_cam_pose[0];
				ORC_ASSERT(baseline != 0);
				const double Z = fxl * baseline / disparity;
				out_lm_pos[0] = (obs[0] - cxl) * Z / fxl;
				out_lm_pos[1] = (obs[1] - cyl) * Z / fyl;
				out_lm_pos[2] = Z;
				break;
			}
		}
		if (parameters.sensor_pose_kind == ORC_SENSORPOSE_SE3) composePosePoint(out_lm_pos, sensor_pose3);
	}

	// ------------------------------------------------------------------ landmark_matcher<OBS>::find_relative_pose
	bool find_relative_pose(const std::vector<const double *> &new_kf_obs, const std::vector<const double *> &old_kf_obs, pose_t &pose_new_kf_wrt_old_kf) const {
		ORC_ASSERT(new_kf_obs.size() == old_kf_obs.size());
		const size_t N = new_kf_obs.size();
		switch (parameters.obs_kind) {
			case ORC_OBS_RELPOSES2D: {  // observations_RelativePoses_2D.h:51-71
				for (size_t i = 0; i < N; i++) {
					const double *kf0 = new_kf_obs[i], *kf1 = old_kf_obs[i];
					if ((kf0[0] != 0 || kf0[1] != 0 || kf0[2] != 0) && (kf1[0] != 0 || kf1[1] != 0 || kf1[2] != 0)) continue;
					const Pose2 new_obs(kf0[0], kf0[1], kf0[2]), old_obs(kf1[0], kf1[1], kf1[2]);
					from_pose3(to_pose3(old_obs - new_obs), pose_new_kf_wrt_old_kf);
					return true;
				}
				return false;
			}
			case ORC_OBS_MONO: return false;  // observations_MonocularCamera.h:50-57
			case ORC_OBS_CARTESIAN3D: {       // observations_Cartesian_3D.h:47-81
				std::vector<double> a, b;
				for (size_t i = 0; i < N; i++) for (int k = 0; k < 3; k++) { a.push_back(old_kf_obs[i][k]); b.push_back(new_kf_obs[i][k]); }
				Pose3 found;
				if (!se3_l2_horn(a, b, found)) return false;
				from_pose3(found, pose_new_kf_wrt_old_kf);
				return true;
			}
			case ORC_OBS_STEREO: {  // observations_StereoCamera.h:59-120
				const double cx = parameters.cam_left[2], cy = parameters.cam_left[3], baseline = parameters.right_cam_pose[0], fl = parameters.cam_left[0];
				std::vector<double> a, b;
				for (size_t i = 0; i < N; i++) {
					const double disparity_old = old_kf_obs[i][0] - old_kf_obs[i][2];
					if (disparity_old <= .0) continue;
					const double disparity_new = new_kf_obs[i][0] - new_kf_obs[i][2];
					if (disparity_new <= .0) continue;
					a.push_back((old_kf_obs[i][0] - cx) * baseline / disparity_old); a.push_back((old_kf_obs[i][1] - cy) * baseline / disparity_old); a.push_back(fl * baseline / disparity_old);
					b.push_back((new_kf_obs[i][0] - cx) * baseline / disparity_new); b.push_back((new_kf_obs[i][1] - cy) * baseline / disparity_new); b.push_back(fl * baseline / disparity_new);
				}
				Pose3 found;
				if (!se3_l2_horn(a, b, found)) return false;
				from_pose3(found, pose_new_kf_wrt_old_kf);
				return true;
			}
		}
		return false;
	}

	// ------------------------------------------------------------------ impl/alloc_keyframe.h:19-28
	TKeyFrameID alloc_keyframe() {
		const TKeyFrameID new_kf_id = keyframes.size();
		keyframes.push_back(keyframe_info());
		return new_kf_id;
	}
	// impl/alloc_kf2kf_edge.h:17-61
	size_t alloc_kf2kf_edge(TKeyFrameID first, TKeyFrameID second, const pose_t &init_inv_pose_val) {
		k2k_edges.push_back(k2k_edge_t());
		k2k_edge_t &new_edge = k2k_edges.back();
		new_edge.from = first; new_edge.to = second;
		ORC_ASSERT(new_edge.from != new_edge.to);
		new_edge.inv_pose = init_inv_pose_val;
		new_edge.id = k2k_edges.size() - 1;
		keyframes[first].adjacent_k2k_edges.push_back(&new_edge);
		keyframes[second].adjacent_k2k_edges.push_back(&new_edge);
		lin_system.dh_dAp.push_back(colAp_t());  // appendCol
		return new_edge.id;
	}
	// impl/create_kf2kf_edge.h:15-37
	size_t create_kf2kf_edge(TKeyFrameID new_kf_id, TKeyFrameID first, TKeyFrameID second, const pose_t &init_inv_pose_val = pose_t()) {
		const size_t ed_id = alloc_kf2kf_edge(first, second, init_inv_pose_val);
		update_symbolic_new_node(new_kf_id, first, second, parameters.max_tree_depth);
		return ed_id;
	}
	bool are_keyframes_connected(TKeyFrameID id1, TKeyFrameID id2) const {  // impl/spantree_misc.h
		ORC_ASSERT(id1 < keyframes.size() && id2 < keyframes.size());
		const std::deque<k2k_edge_t *> &e = keyframes[id1].adjacent_k2k_edges;
		for (size_t i = 0; i < e.size(); i++)
			if ((e[i]->from == id1 && e[i]->to == id2) || (e[i]->from == id2 && e[i]->to == id1)) return true;
		return false;
	}
	static TKeyFrameID other_end(TKeyFrameID one, const k2k_edge_t &e) { return e.from == one ? e.to : e.from; }  // getTheOtherFromPair2

	// ------------------------------------------------------------------ impl/spantree_update_symbolic.h:227-300
	bool find_path_bfs(TKeyFrameID cur_node, TKeyFrameID trg_node, std::vector<TKeyFrameID> *out_path_IDs, k2k_edge_vector_t *out_path_edges) const {
		if (out_path_IDs) out_path_IDs->clear();
		if (out_path_edges) out_path_edges->clear();
		if (cur_node == trg_node) return true;
		struct TBFSEntry { TKeyFrameID prev; k2k_edge_t *prev_edge; topo_dist_t dist; TBFSEntry() : prev(0), prev_edge(NULL), dist(std::numeric_limits<topo_dist_t>::max()) {} };
		std::set<TKeyFrameID> visited;
		std::queue<TKeyFrameID> pending;
		std::map<TKeyFrameID, TBFSEntry> preceding;
		pending.push(cur_node);
		visited.insert(cur_node);
		preceding[cur_node].dist = 0;
		while (!pending.empty()) {
			const TKeyFrameID next_kf = pending.front();
			pending.pop();
			TBFSEntry &bfs_data_next = preceding[next_kf];
			const topo_dist_t cur_dist = bfs_data_next.dist;
			if (next_kf == trg_node) {
				topo_dist_t dist = bfs_data_next.dist;
				if (out_path_IDs) out_path_IDs->resize(dist);
				if (out_path_edges) out_path_edges->resize(dist);
				TKeyFrameID path_node = trg_node;
				while (path_node != cur_node) {
					ORC_ASSERT(dist > 0);
					// NOTE: the reference decrements `dist` once per requested output; callers ask for one.
					typename std::map<TKeyFrameID, TBFSEntry>::const_iterator it_prec = preceding.find(path_node);
					ORC_ASSERT(it_prec != preceding.end());
					--dist;
					if (out_path_IDs) (*out_path_IDs)[dist] = path_node;
					path_node = it_prec->second.prev;
					if (out_path_edges) (*out_path_edges)[dist] = it_prec->second.prev_edge;
				}
				return true;
			}
			const keyframe_info &kfi = keyframes[next_kf];
			for (size_t i = 0; i < kfi.adjacent_k2k_edges.size(); i++) {
				const k2k_edge_t *ed = kfi.adjacent_k2k_edges[i];
				const TKeyFrameID new_kf = other_end(next_kf, *ed);
				if (!visited.count(new_kf)) {
					pending.push(new_kf);
					visited.insert(new_kf);
					TBFSEntry &pe = preceding[new_kf];
					if (pe.dist > cur_dist + 1) { pe.dist = cur_dist + 1; pe.prev = next_kf; pe.prev_edge = const_cast<k2k_edge_t *>(ed); }
				}
			}
		}
		return false;
	}

	// ------------------------------------------------------------------ impl/spantree_update_symbolic.h:19-211
	void update_symbolic_new_node(TKeyFrameID new_node_id, TKeyFrameID e_first, TKeyFrameID e_second, topo_dist_t max_depth) {
		ORC_ASSERT(max_depth >= 1);
		typedef std::pair<TKeyFrameID, TKeyFrameID> TPairKeyFrameID;
		std::set<TPairKeyFrameID> kfs_with_modified_next_edge;
		{
			const TKeyFrameID ik = (e_first == new_node_id) ? e_second : e_first;  // getTheOtherFromPair
			const std::map<TKeyFrameID, TSpanTreeEntry> &st_ik = spanning_tree.sym.next_edge[ik];
			std::vector<std::pair<TKeyFrameID, topo_dist_t> > tk;  // distances are COPIES (snapshot)
			for (typename std::map<TKeyFrameID, TSpanTreeEntry>::const_iterator it = st_ik.begin(); it != st_ik.end(); ++it)
				if (it->second.distance < max_depth) tk.push_back(std::make_pair(it->first, it->second.distance));
			tk.push_back(std::make_pair(ik, topo_dist_t(0)));
			const std::map<TKeyFrameID, TSpanTreeEntry> &st_n = spanning_tree.sym.next_edge[new_node_id];
			std::vector<std::pair<TKeyFrameID, const TSpanTreeEntry *> > STn;  // entries are POINTERS (live)
			for (typename std::map<TKeyFrameID, TSpanTreeEntry>::const_iterator it = st_n.begin(); it != st_n.end(); ++it)
				STn.push_back(std::make_pair(it->first, &it->second));
			STn.push_back(std::pair<TKeyFrameID, const TSpanTreeEntry *>(new_node_id, static_cast<const TSpanTreeEntry *>(NULL)));
			for (size_t r_idx = 0; r_idx < STn.size(); r_idx++) {
				const TKeyFrameID r = STn[r_idx].first;
				const TSpanTreeEntry *ste_n2r = STn[r_idx].second;
				const topo_dist_t dist_r2n = ste_n2r ? ste_n2r->distance : 0;
				std::map<TKeyFrameID, TSpanTreeEntry> &st_r = spanning_tree.sym.next_edge[r];
				TSpanTreeEntry *ste_r2n = NULL;
				if (r != new_node_id) {
					typename std::map<TKeyFrameID, TSpanTreeEntry>::iterator it = st_r.find(new_node_id);
					ORC_ASSERT(it != st_r.end());
					ste_r2n = &it->second;
				}
				for (size_t s_idx = 0; s_idx < tk.size(); s_idx++) {
					const TKeyFrameID s = tk[s_idx].first;
					if (r == s) continue;
					const topo_dist_t dist_s2ik = tk[s_idx].second;
					std::map<TKeyFrameID, TSpanTreeEntry> &st_s = spanning_tree.sym.next_edge[s];
					TSpanTreeEntry *ste_s2ik = NULL;
					if (s != ik) {
						typename std::map<TKeyFrameID, TSpanTreeEntry>::iterator it_ik_inSTs = st_s.find(ik);
						ORC_ASSERT(it_ik_inSTs != st_s.end());
						ste_s2ik = &it_ik_inSTs->second;
					}
					const topo_dist_t new_dist = dist_r2n + dist_s2ik + 1;
					typename std::map<TKeyFrameID, TSpanTreeEntry>::iterator it_s_inSTr = st_r.find(s);
					if (it_s_inSTr != st_r.end()) {
						if (new_dist < it_s_inSTr->second.distance) {
							it_s_inSTr->second.distance = new_dist;
							it_s_inSTr->second.next = ste_r2n ? ste_r2n->next : ik;
							ORC_ASSERT(r != it_s_inSTr->second.next);
							TSpanTreeEntry &ste_r_inSTs = st_s[r];
							ste_r_inSTs.distance = new_dist;
							ste_r_inSTs.next = ste_s2ik ? ste_s2ik->next : new_node_id;
							ORC_ASSERT(s != ste_r_inSTs.next);
							kfs_with_modified_next_edge.insert(std::make_pair(s, r));
							kfs_with_modified_next_edge.insert(std::make_pair(r, s));
						}
					} else {
						if (new_dist <= max_depth) {
							TSpanTreeEntry &ste_s_inSTr = st_r[s];
							ste_s_inSTr.distance = new_dist;
							ste_s_inSTr.next = ste_r2n ? ste_r2n->next : ik;
							ORC_ASSERT(r != ste_s_inSTr.next);
							TSpanTreeEntry &ste_r_inSTs = st_s[r];
							ste_r_inSTs.distance = new_dist;
							ste_r_inSTs.next = ste_s2ik ? ste_s2ik->next : new_node_id;
							ORC_ASSERT(s != ste_r_inSTs.next);
							kfs_with_modified_next_edge.insert(std::make_pair(r, s));
							kfs_with_modified_next_edge.insert(std::make_pair(s, r));
						}
					}
				}
			}
		}
		// Update "all_edges" (:167-188)
		for (typename std::set<TPairKeyFrameID>::const_iterator it = kfs_with_modified_next_edge.begin(); it != kfs_with_modified_next_edge.end(); ++it) {
			const TKeyFrameID kf_id = it->first;
			const std::map<TKeyFrameID, TSpanTreeEntry> &Ds = spanning_tree.sym.next_edge[kf_id];
			typename std::map<TKeyFrameID, TSpanTreeEntry>::const_iterator it2 = Ds.find(it->second);
			ORC_ASSERT(it2 != Ds.end());
			const TKeyFrameID dst_kf_id = it2->first;
			const TKeyFrameID from = std::max(dst_kf_id, kf_id), to = std::min(dst_kf_id, kf_id);
			k2k_edge_vector_t &path = spanning_tree.sym.all_edges[from][to];
			path.clear();
			const bool path_found = find_path_bfs(from, to, NULL, &path);
			ORC_ASSERT(path_found);
		}
	}

	// ------------------------------------------------------------------ impl/spantree_update_numeric.h:19-81
	size_t update_numeric_only_all_from_node(TKeyFrameID id_from, const std::map<TKeyFrameID, k2k_edge_vector_t> &row, bool skip_marked_as_uptodate) {
		frameid2pose_map_t &frameid2pose_map = spanning_tree.num[id_from];
		for (typename std::map<TKeyFrameID, k2k_edge_vector_t>::const_iterator itE = row.begin(); itE != row.end(); ++itE) {
			const TKeyFrameID id_to = itE->first;
			pose_flag_t &i2j = frameid2pose_map[id_to];
			pose_flag_t &j2i = spanning_tree.num[id_to][id_from];
			if (skip_marked_as_uptodate && i2j.updated && j2i.updated) continue;
			const k2k_edge_vector_t &ev = itE->second;
			pose_t accum;
			TKeyFrameID curKF = id_from;
			for (size_t k = 0; k < ev.size(); k++) {
				if (ev[k]->to == curKF) { accum.composeFrom(accum, ev[k]->inv_pose); curKF = ev[k]->from; }
				else { accum.composeFrom(accum, -ev[k]->inv_pose); curKF = ev[k]->to; }
			}
			i2j.pose = accum; i2j.updated = true;
			j2i.pose = -accum; j2i.updated = true;
		}
		return row.size();
	}
	size_t update_numeric(bool skip_marked_as_uptodate = false) {  // :98-109
		size_t pose_count = 0;
		for (size_t i = 0; i < spanning_tree.sym.all_edges.size(); i++)
			pose_count += update_numeric_only_all_from_node(i, spanning_tree.sym.all_edges.d[i], skip_marked_as_uptodate);
		return pose_count;
	}
	size_t update_numeric(const std::set<TKeyFrameID> &kfs_to_update, bool skip_marked_as_uptodate) {  // :111-127
		size_t pose_count = 0;
		for (std::set<TKeyFrameID>::const_iterator it = kfs_to_update.begin(); it != kfs_to_update.end(); ++it) {
			const std::map<TKeyFrameID, k2k_edge_vector_t> *row = spanning_tree.sym.all_edges.find(*it);
			if (row) pose_count += update_numeric_only_all_from_node(*it, *row, skip_marked_as_uptodate);
		}
		return pose_count;
	}

	// ------------------------------------------------------------------ impl/spantree_create_complete.h:18-127
	void create_complete_spanning_tree(TKeyFrameID root_id, std::map<TKeyFrameID, pose_t> &span_tree, size_t max_depth) const {
		span_tree.clear();
		struct TBFSEntryEdges { TKeyFrameID prev; const k2k_edge_t *edge; topo_dist_t dist; TBFSEntryEdges() : prev(0), edge(NULL), dist(std::numeric_limits<topo_dist_t>::max()) {} };
		std::set<TKeyFrameID> visited;
		std::queue<TKeyFrameID> pending;
		std::map<TKeyFrameID, TBFSEntryEdges> preceding;
		pending.push(root_id);
		visited.insert(root_id);
		preceding[root_id].dist = 0;
		while (!pending.empty()) {
			const TKeyFrameID next_kf = pending.front();
			pending.pop();
			const topo_dist_t cur_dist = preceding[next_kf].dist;
			if (cur_dist >= max_depth) continue;
			const keyframe_info &kfi = keyframes[next_kf];
			for (size_t i = 0; i < kfi.adjacent_k2k_edges.size(); i++) {
				const k2k_edge_t *ed = kfi.adjacent_k2k_edges[i];
				const TKeyFrameID new_kf = other_end(next_kf, *ed);
				if (!visited.count(new_kf)) {
					pending.push(new_kf);
					visited.insert(new_kf);
					TBFSEntryEdges &pe = preceding[new_kf];
					if (pe.dist > cur_dist + 1) { pe.dist = cur_dist + 1; pe.prev = next_kf; pe.edge = ed; }
				}
			}
		}
		std::multimap<topo_dist_t, std::pair<TKeyFrameID, TBFSEntryEdges> > kfs_by_depth;
		for (typename std::map<TKeyFrameID, TBFSEntryEdges>::const_iterator it = preceding.begin(); it != preceding.end(); ++it)
			kfs_by_depth.insert(std::make_pair(it->second.dist, *it));
		for (typename std::multimap<topo_dist_t, std::pair<TKeyFrameID, TBFSEntryEdges> >::const_iterator it = kfs_by_depth.begin(); it != kfs_by_depth.end(); ++it) {
			const TKeyFrameID kf_id = it->second.first;
			const TBFSEntryEdges &bfs_data = it->second.second;
			if (it->first == 0) span_tree[kf_id] = pose_t();
			else {
				ORC_ASSERT(bfs_data.edge != NULL);
				const pose_t parent_pose = span_tree[bfs_data.prev];
				pose_t &my_pose = span_tree[kf_id];
				if (bfs_data.edge->to == kf_id) my_pose.composeFrom(parent_pose, -bfs_data.edge->inv_pose);
				else my_pose.composeFrom(parent_pose, bfs_data.edge->inv_pose);
			}
		}
	}

	// ------------------------------------------------------------------ impl/eval_overall_error.h:15-137
	// One complete (unbounded-depth) BFS spanning tree per root = min(observer, base) of every observation; the pose of
	// the larger id as seen from the root comes from that tree, its inverse serves the opposite direction (:72-76).
	// No robust kernel here (:129 plain squaredNorm).
	double eval_overall_squared_error() const {
		double sqerr = 0;
		std::map<TKeyFrameID, std::set<TKeyFrameID> > ob_pairs;  // minimum id first
		for (typename std::deque<k2f_edge_t>::const_iterator itO = all_observations.begin(); itO != all_observations.end(); ++itO) {
			const TKeyFrameID obs_id = itO->kf_id, base_id = itO->feat_rel_pos->id_frame_base;
			ob_pairs[std::min(obs_id, base_id)].insert(std::max(obs_id, base_id));
		}
		std::map<TKeyFrameID, std::map<TKeyFrameID, pose_t> > all_rel_poses;  // [SOURCE][TARGET] = TARGET as seen from SOURCE
		for (std::map<TKeyFrameID, std::set<TKeyFrameID> >::const_iterator it1 = ob_pairs.begin(); it1 != ob_pairs.end(); ++it1) {
			const TKeyFrameID root = it1->first;
			std::map<TKeyFrameID, pose_t> span_tree;
			create_complete_spanning_tree(root, span_tree, std::numeric_limits<size_t>::max());
			for (std::set<TKeyFrameID>::const_iterator w = it1->second.begin(); w != it1->second.end(); ++w) {
				typename std::map<TKeyFrameID, pose_t>::const_iterator have = span_tree.find(*w);
				if (have == span_tree.end()) continue;
				all_rel_poses[root][*w] = have->second;
				all_rel_poses[*w][root] = -have->second;
			}
		}
		for (typename std::deque<k2f_edge_t>::const_iterator itO = all_observations.begin(); itO != all_observations.end(); ++itO) {
			const TKeyFrameID obs_frame_id = itO->kf_id, base_id = itO->feat_rel_pos->id_frame_base;
			const pose_t *base_pose_wrt_observer = NULL;
			if (base_id == obs_frame_id) base_pose_wrt_observer = &aux_null_pose;
			else {
				typename std::map<TKeyFrameID, std::map<TKeyFrameID, pose_t> >::const_iterator m = all_rel_poses.find(obs_frame_id);
				ORC_ASSERT(m != all_rel_poses.end());
				typename std::map<TKeyFrameID, pose_t>::const_iterator itRelPose = m->second.find(base_id);
				ORC_ASSERT(itRelPose != m->second.end());
				base_pose_wrt_observer = &itRelPose->second;
			}
			double delta[MAX_OBS];
			observe_error(delta, itO->obs_arr, *base_pose_wrt_observer, itO->feat_rel_pos->pos);
			for (int k = 0; k < O; k++) sqerr += delta[k] * delta[k];
		}
		return sqerr;
	}

	// RbaEngine.h:385-396
	const pose_t *get_kf_relative_pose(TKeyFrameID kf_query, TKeyFrameID kf_reference) const {
		const frameid2pose_map_t *m = spanning_tree.num.find(kf_reference);
		if (!m) return NULL;
		typename frameid2pose_map_t::const_iterator it = m->find(kf_query);
		return it != m->end() ? &it->second.pose : NULL;
	}

	// ------------------------------------------------------------------ impl/add-observations.h:17-264
	size_t add_observation(TKeyFrameID observing_kf_id, TLandmarkID feat_id, const double *obs_data, const double *fixed_relative_position, const double *unknown_relative_position_init_val) {
		ORC_ASSERT(!(fixed_relative_position != NULL && unknown_relative_position_init_val != NULL));
		const bool is_1st_time_seen = (feat_id >= all_lms.size() || all_lms[feat_id].rfp == NULL);
		const bool is_fixed = (fixed_relative_position != NULL) || (!is_1st_time_seen && all_lms[feat_id].has_known_pos);
		const size_t new_obs_idx = all_observations.size();
		all_observations.push_back(k2f_edge_t());
		all_observations_Jacob_validity.push_back(1);
		char *const jacob_valid_bit = &all_observations_Jacob_validity.back();
		k2f_edge_t &new_k2f_edge = all_observations.back();
		if (is_1st_time_seen) {
			if (is_fixed) {
				TRelativeLandmarkPos new_rfp;
				new_rfp.id_frame_base = observing_kf_id;
				std::memcpy(new_rfp.pos, fixed_relative_position, sizeof(new_rfp.pos));
				typename TRelativeLandmarkPosMap::iterator it_new = known_lms.insert(known_lms.end(), std::make_pair(feat_id, new_rfp));
				if (feat_id >= all_lms.size()) all_lms.resize(feat_id + 1);
				all_lms[feat_id] = TLandmarkEntry(true, &it_new->second);
			} else {
				TRelativeLandmarkPos new_rfp;
				new_rfp.id_frame_base = observing_kf_id;
				if (unknown_relative_position_init_val) std::memcpy(new_rfp.pos, unknown_relative_position_init_val, sizeof(new_rfp.pos));
				else inverse_sensor_model(new_rfp.pos, obs_data);
				typename TRelativeLandmarkPosMap::iterator it_new = unknown_lms.insert(unknown_lms.end(), std::make_pair(feat_id, new_rfp));
				if (feat_id >= all_lms.size()) all_lms.resize(feat_id + 1);
				all_lms[feat_id] = TLandmarkEntry(false, &it_new->second);
				if (feat_id >= lin_system.dh_df_remap.size()) lin_system.dh_df_remap.resize(feat_id + 1, SRBA_INVALID_INDEX);
				lin_system.dh_df_remap[feat_id] = lin_system.dh_df.size();  // appendCol(feat_id)
				lin_system.dh_df.push_back(colF_t());
			}
		}
		TRelativeLandmarkPos *lm_rel_pos = all_lms[feat_id].rfp;
		const TKeyFrameID base_id = lm_rel_pos->id_frame_base;
		new_k2f_edge.kf_id = observing_kf_id;
		new_k2f_edge.feat_id = feat_id;
		for (int i = 0; i < O; i++) { new_k2f_edge.obs_data[i] = obs_data[i]; new_k2f_edge.obs_arr[i] = obs_data[i]; }
		new_k2f_edge.is_first_obs_of_unknown = is_1st_time_seen && !is_fixed;
		new_k2f_edge.feat_has_known_rel_pos = is_fixed;
		new_k2f_edge.feat_rel_pos = lm_rel_pos;
		keyframe_info &kf_info = keyframes[observing_kf_id];
		kf_info.adjacent_k2f_edges.push_back(&new_k2f_edge);

		bool graph_says_ignore_this_obs = false;
		if (base_id != new_k2f_edge.kf_id) {  // :135-215
			const TKeyFrameID from = std::max(observing_kf_id, base_id), to = std::min(observing_kf_id, base_id);
			const bool all_edges_inverse = (from != observing_kf_id);
			const std::map<TKeyFrameID, k2k_edge_vector_t> *it_map = spanning_tree.sym.all_edges.find(from);
			if (!it_map) throw std::runtime_error("No ST.all_edges found for observing_id / base_id");
			typename std::map<TKeyFrameID, k2k_edge_vector_t>::const_iterator it_obs_ed = it_map->find(to);
			if (it_obs_ed != it_map->end()) {
				const k2k_edge_vector_t &obs_edges = it_obs_ed->second;
				ORC_ASSERT(!obs_edges.empty());
				TKeyFrameID curKF = observing_kf_id;
				for (size_t j = 0; j < obs_edges.size(); j++) {
					const size_t i = all_edges_inverse ? (obs_edges.size() - j - 1) : j;
					ORC_ASSERT(curKF != to);
					const size_t edge_id = obs_edges[i]->id;
					const bool normal_dir = (obs_edges[i]->to == curKF);
					colAp_t &col = lin_system.dh_dAp[edge_id];
					typename colAp_t::iterator it_new_entry = col.insert(col.end(), std::make_pair(new_obs_idx, EntryAp()));
					EntryAp &entry = it_new_entry->second;
					entry.sym.obs_idx = new_obs_idx;
					entry.sym.is_valid = jacob_valid_bit;
					entry.sym.edge_normal_dir = normal_dir;
					entry.sym.kf_d = curKF;
					entry.sym.kf_base = base_id;
					entry.sym.feat_rel_pos = lm_rel_pos;
					entry.sym.k2k_edge_id = edge_id;
					entry.sym.rel_pose_base_from_d1 = &spanning_tree.num[curKF][base_id];
					entry.sym.rel_pose_d1_from_obs = (curKF == observing_kf_id) ? NULL : &spanning_tree.num[observing_kf_id][curKF];
					curKF = normal_dir ? obs_edges[i]->from : obs_edges[i]->to;
				}
			} else graph_says_ignore_this_obs = true;
		}
		if (!is_fixed && !graph_says_ignore_this_obs) {  // :220-257
			ORC_ASSERT(feat_id < lin_system.dh_df_remap.size() && lin_system.dh_df_remap[feat_id] != SRBA_INVALID_INDEX);
			const size_t col_idx = lin_system.dh_df_remap[feat_id];
			colF_t &col = lin_system.dh_df[col_idx];
			typename colF_t::iterator it_new_entry = col.insert(col.end(), std::make_pair(new_obs_idx, EntryF()));
			EntryF &entry = it_new_entry->second;
			entry.sym.obs_idx = new_obs_idx;
			entry.sym.is_valid = jacob_valid_bit;
			entry.sym.feat_rel_pos = lm_rel_pos;
			entry.sym.rel_pose_base_from_obs = new_k2f_edge.is_first_obs_of_unknown ? NULL : &spanning_tree.num[observing_kf_id][base_id];
		}
		return new_obs_idx;
	}

	// ------------------------------------------------------------------ impl/make_ordered_list_base_kfs.h:16-43
	typedef std::multimap<size_t, TKeyFrameID, std::greater<size_t> > base_sorted_lst_t;  // srba_types.h:25
	void make_ordered_list_base_kfs(const new_kf_observations_t &obs, base_sorted_lst_t &obs_for_each_base_sorted) const {
		std::map<TKeyFrameID, size_t> obs_for_each_base;
		for (typename new_kf_observations_t::const_iterator itObs = obs.begin(); itObs != obs.end(); ++itObs) {
			const TLandmarkID lm_id = itObs->feat_id;
			if (lm_id >= all_lms.size()) continue;
			const TLandmarkEntry &lme = all_lms[lm_id];
			if (!lme.rfp) continue;
			obs_for_each_base[lme.rfp->id_frame_base]++;
		}
		for (std::map<TKeyFrameID, size_t>::const_iterator it = obs_for_each_base.begin(); it != obs_for_each_base.end(); ++it)
			obs_for_each_base_sorted.insert(std::make_pair(it->second, it->first));
	}

	// ------------------------------------------------------------------ ecps/local_areas_fixed_size.h:52-213
	TKeyFrameID get_center_kf_for_kf(TKeyFrameID kf_id) const { return parameters.submap_size * (kf_id / parameters.submap_size); }
	void ecp_eval(TKeyFrameID new_kf_id, const new_kf_observations_t &obs, std::vector<TNewEdgeInfo> &new_k2k_edge_ids) {
		ORC_ASSERT(new_kf_id >= 1);
		const size_t MINIMUM_OBS_TO_LOOP_CLOSURE = parameters.min_obs_to_loop_closure;
		const TKeyFrameID current_center_kf_id = get_center_kf_for_kf(new_kf_id);
		const topo_dist_t min_dist_for_loop_closure = parameters.max_tree_depth + 1;
		base_sorted_lst_t obs_for_each_base_sorted;
		make_ordered_list_base_kfs(obs, obs_for_each_base_sorted);
		std::map<TKeyFrameID, size_t> obs_for_each_area;
		std::map<TKeyFrameID, bool> base_is_center_for_all_obs_in_area;
		std::map<TKeyFrameID, std::map<TKeyFrameID, size_t> > obs_for_base_KF_grouped_by_area;
		for (base_sorted_lst_t::const_iterator it = obs_for_each_base_sorted.begin(); it != obs_for_each_base_sorted.end(); ++it) {
			const size_t num_obs_this_base = it->first;
			const TKeyFrameID base_id = it->second;
			const TKeyFrameID this_localmap_center = get_center_kf_for_kf(base_id);
			obs_for_each_area[this_localmap_center] += num_obs_this_base;
			obs_for_base_KF_grouped_by_area[this_localmap_center][base_id] += num_obs_this_base;
			if (base_is_center_for_all_obs_in_area.find(this_localmap_center) == base_is_center_for_all_obs_in_area.end())
				base_is_center_for_all_obs_in_area[this_localmap_center] = true;
			if (base_id != this_localmap_center) base_is_center_for_all_obs_in_area[this_localmap_center] = false;
		}
		base_sorted_lst_t obs_for_each_area_sorted;
		for (std::map<TKeyFrameID, size_t>::const_iterator it = obs_for_each_area.begin(); it != obs_for_each_area.end(); ++it)
			obs_for_each_area_sorted.insert(std::make_pair(it->second, it->first));
		std::map<TKeyFrameID, base_sorted_lst_t> obs_for_base_KF_grouped_by_area_sorted;
		for (std::map<TKeyFrameID, std::map<TKeyFrameID, size_t> >::const_iterator it = obs_for_base_KF_grouped_by_area.begin(); it != obs_for_base_KF_grouped_by_area.end(); ++it) {
			base_sorted_lst_t &bsl = obs_for_base_KF_grouped_by_area_sorted[it->first];
			for (std::map<TKeyFrameID, size_t>::const_iterator it2 = it->second.begin(); it2 != it->second.end(); ++it2)
				bsl.insert(std::make_pair(it2->second, it2->first));
		}
		if (current_center_kf_id != new_kf_id) {  // :118-130
			TNewEdgeInfo nei;
			nei.id = create_kf2kf_edge(new_kf_id, current_center_kf_id, new_kf_id);
			nei.has_approx_init_val = false;
			new_k2k_edge_ids.push_back(nei);
		}
		for (base_sorted_lst_t::const_iterator it = obs_for_each_area_sorted.begin(); it != obs_for_each_area_sorted.end(); ++it) {  // :133-197
			const size_t num_obs_this_base = it->first;
			const TKeyFrameID remote_center_kf_id = it->second;
			const TKeyFrameID from_id = current_center_kf_id, to_id = remote_center_kf_id;
			if (from_id == to_id) continue;
			topo_dist_t found_distance = std::numeric_limits<topo_dist_t>::max();
			const std::map<TKeyFrameID, TSpanTreeEntry> *from_Ds = spanning_tree.sym.next_edge.find(from_id);
			if (from_Ds) {
				typename std::map<TKeyFrameID, TSpanTreeEntry>::const_iterator it_to_dist = from_Ds->find(to_id);
				if (it_to_dist != from_Ds->end()) found_distance = it_to_dist->second.distance;
			}
			topo_dist_t dist_extra_edges = 2;
			if (current_center_kf_id == new_kf_id) dist_extra_edges--;
			if (base_is_center_for_all_obs_in_area[remote_center_kf_id]) dist_extra_edges--;
			if (found_distance >= min_dist_for_loop_closure - dist_extra_edges) {
				if (num_obs_this_base >= MINIMUM_OBS_TO_LOOP_CLOSURE) {
					TNewEdgeInfo nei;
					nei.id = create_kf2kf_edge(from_id, to_id, from_id);
					nei.has_approx_init_val = false;
					nei.loopclosure_observer_kf = new_kf_id;
					const base_sorted_lst_t &bsl = obs_for_base_KF_grouped_by_area_sorted[remote_center_kf_id];
					ORC_ASSERT(!bsl.empty());
					nei.loopclosure_base_kf = bsl.begin()->second;
					new_k2k_edge_ids.push_back(nei);
				}
			}
		}
		if (new_k2k_edge_ids.empty()) throw std::runtime_error("ECP: no suitable linking KF found: the node becomes isolated of the graph!");  // :199
	}

	// ------------------------------------------------------------------ impl/determine_kf2kf_edges_to_create.h:17-266
	void determine_kf2kf_edges_to_create(TKeyFrameID new_kf_id, const new_kf_observations_t &obs, std::vector<TNewEdgeInfo> &new_k2k_edge_ids) {
		new_k2k_edge_ids.clear();
		if (keyframes.size() == 1) return;
		ecp_eval(new_kf_id, obs, new_k2k_edge_ids);
		const size_t nNEI = new_k2k_edge_ids.size();
		for (size_t i = 0; i < nNEI; i++) {
			TNewEdgeInfo &nei = new_k2k_edge_ids[i];
			if (nei.has_approx_init_val) continue;
			k2k_edge_t &nei_edge = k2k_edges[nei.id];
			const bool nei_edge_does_not_touch_cur_kf = (nei_edge.to != new_kf_id) && (nei_edge.from != new_kf_id);
			if (!nei_edge_does_not_touch_cur_kf && last_timestep_touched_kfs.count(nei_edge.from) != 0) {  // Method #1 :48-62
				const pose_t *rel_pose = get_kf_relative_pose(new_kf_id - 1, nei_edge.from);
				if (rel_pose) {
					const bool edge_dir_to_newkf = (nei_edge.to == new_kf_id);
					if (edge_dir_to_newkf) nei_edge.inv_pose = -(*rel_pose);
					else nei_edge.inv_pose = *rel_pose;
					nei.has_approx_init_val = true;
				}
			}
			if (!nei.has_approx_init_val) {  // Method #2 :66-254
				TKeyFrameID last_kf_id, other_kf_id;
				if (nei_edge_does_not_touch_cur_kf) { last_kf_id = nei_edge.from; other_kf_id = nei_edge.to; }
				else { last_kf_id = new_kf_id; other_kf_id = (nei_edge.to == new_kf_id) ? nei_edge.from : nei_edge.to; }
				std::vector<const double *> new_kf_obs, old_kf_obs;
				{
					const std::deque<k2f_edge_t *> &other_k2f_edges = keyframes[other_kf_id].adjacent_k2f_edges;
					std::map<TLandmarkID, size_t> newkf_obs_feats;
					const std::deque<k2f_edge_t *> *last_k2f_edges = NULL;
					if (nei_edge_does_not_touch_cur_kf) {
						last_k2f_edges = &keyframes[last_kf_id].adjacent_k2f_edges;
						for (size_t k = 0; k < last_k2f_edges->size(); k++) newkf_obs_feats[(*last_k2f_edges)[k]->feat_id] = k;
					} else {
						for (size_t k = 0; k < obs.size(); k++) newkf_obs_feats[obs[k].feat_id] = k;
					}
					for (size_t k = 0; k < other_k2f_edges.size(); k++) {
						const TLandmarkID lm_id = other_k2f_edges[k]->feat_id;
						std::map<TLandmarkID, size_t>::const_iterator it_id = newkf_obs_feats.find(lm_id);
						if (it_id == newkf_obs_feats.end()) continue;
						old_kf_obs.push_back(other_k2f_edges[k]->obs_data);
						if (nei_edge_does_not_touch_cur_kf) new_kf_obs.push_back((*last_k2f_edges)[it_id->second]->obs_data);
						else new_kf_obs.push_back(obs[it_id->second].obs_data);
					}
				}
				pose_t pose_new_kf_wrt_old_kf;
				bool found_ok = find_relative_pose(new_kf_obs, old_kf_obs, pose_new_kf_wrt_old_kf);
				if (!found_ok && nei.loopclosure_observer_kf != SRBA_INVALID_KEYFRAMEID && nei.loopclosure_base_kf != SRBA_INVALID_KEYFRAMEID) {  // 2nd attempt :125-175
					last_kf_id = nei.loopclosure_observer_kf;
					other_kf_id = nei.loopclosure_base_kf;
					const std::deque<k2f_edge_t *> &other2_k2f_edges = keyframes[other_kf_id].adjacent_k2f_edges;
					const bool last_kf_is_new_kf = (last_kf_id == new_kf_id);
					new_kf_obs.clear(); old_kf_obs.clear();
					std::map<TLandmarkID, size_t> newkf_obs_feats;
					const std::deque<k2f_edge_t *> *last_k2f_edges = NULL;
					if (!last_kf_is_new_kf) {
						last_k2f_edges = &keyframes[last_kf_id].adjacent_k2f_edges;
						for (size_t k = 0; k < last_k2f_edges->size(); k++) newkf_obs_feats[(*last_k2f_edges)[k]->feat_id] = k;
					} else {
						for (size_t k = 0; k < obs.size(); k++) newkf_obs_feats[obs[k].feat_id] = k;
					}
					for (size_t k = 0; k < other2_k2f_edges.size(); k++) {
						const TLandmarkID lm_id = other2_k2f_edges[k]->feat_id;
						std::map<TLandmarkID, size_t>::const_iterator it_id = newkf_obs_feats.find(lm_id);
						if (it_id == newkf_obs_feats.end()) continue;
						old_kf_obs.push_back(other2_k2f_edges[k]->obs_data);
						if (!last_kf_is_new_kf) new_kf_obs.push_back((*last_k2f_edges)[it_id->second]->obs_data);
						else new_kf_obs.push_back(obs[it_id->second].obs_data);
					}
					found_ok = find_relative_pose(new_kf_obs, old_kf_obs, pose_new_kf_wrt_old_kf);
				}
				if (found_ok) {  // :177-247
					Pose3 sensor_pose;  // robot2sensor(CPose3D(), sensor_pose, params)
					if (parameters.sensor_pose_kind == ORC_SENSORPOSE_SE3) sensor_pose = sensor_pose3;
					from_pose3((sensor_pose + to_pose3(pose_new_kf_wrt_old_kf)) + (-sensor_pose), pose_new_kf_wrt_old_kf);
					const bool edge_dir_to_newkf = (nei_edge.to == new_kf_id);
					nei.has_approx_init_val = true;
					if (!nei_edge_does_not_touch_cur_kf) {
						if (edge_dir_to_newkf) nei_edge.inv_pose = -pose_new_kf_wrt_old_kf;
						else nei_edge.inv_pose = pose_new_kf_wrt_old_kf;
					} else {
						ORC_ASSERT(nei.loopclosure_observer_kf != SRBA_INVALID_KEYFRAMEID && nei.loopclosure_base_kf != SRBA_INVALID_KEYFRAMEID);
						const pose_t default_identity_pose;
						const pose_t *pose_observer_wrt_to = (nei.loopclosure_observer_kf == nei_edge.to) ? &default_identity_pose : get_kf_relative_pose(nei.loopclosure_observer_kf, nei_edge.to);
						const pose_t *pose_base_wrt_to = (nei.loopclosure_base_kf == nei_edge.to) ? &default_identity_pose : get_kf_relative_pose(nei.loopclosure_base_kf, nei_edge.to);
						const pose_t *pose_observer_wrt_from = (nei.loopclosure_observer_kf == nei_edge.from) ? &default_identity_pose : get_kf_relative_pose(nei.loopclosure_observer_kf, nei_edge.from);
						const pose_t *pose_base_wrt_from = (nei.loopclosure_base_kf == nei_edge.from) ? &default_identity_pose : get_kf_relative_pose(nei.loopclosure_base_kf, nei_edge.from);
						const bool observer_is_near_to = (pose_observer_wrt_to || pose_base_wrt_from) || !(pose_observer_wrt_from || pose_base_wrt_to);
						const pose_t *pose_observer_wrt_local = observer_is_near_to ? (pose_observer_wrt_to ? pose_observer_wrt_to : &default_identity_pose)
						                                                            : (pose_observer_wrt_from ? pose_observer_wrt_from : &default_identity_pose);
						const pose_t *pose_base_wrt_remote = observer_is_near_to ? (pose_base_wrt_from ? pose_base_wrt_from : &default_identity_pose)
						                                                         : (pose_base_wrt_to ? pose_base_wrt_to : &default_identity_pose);
						const pose_t pose_local_wrt_remote = ((*pose_base_wrt_remote) + (pose_new_kf_wrt_old_kf)) + (-(*pose_observer_wrt_local));
#ifdef ORC_EXPERIMENT_FIX_LC_INIT
						if (observer_is_near_to) nei_edge.inv_pose = -pose_local_wrt_remote; else nei_edge.inv_pose = pose_local_wrt_remote;
#else
						if (edge_dir_to_newkf) nei_edge.inv_pose = -pose_local_wrt_remote;
						else nei_edge.inv_pose = pose_local_wrt_remote;
#endif
					}
				}
			}
		}
		last_timestep_touched_kfs.clear();  // :257-262
		for (size_t i = 0; i < nNEI; i++) {
			const k2k_edge_t &nei_edge = k2k_edges[new_k2k_edge_ids[i].id];
			last_timestep_touched_kfs.insert(nei_edge.from);
			last_timestep_touched_kfs.insert(nei_edge.to);
		}
	}

	// ------------------------------------------------------------------ impl/define_new_keyframe.h:16-116
	void define_new_keyframe(const new_kf_observations_t &obs, TNewKeyFrameInfo &out_new_kf_info, const bool run_local_optimization = true) {
		out_new_kf_info.kf_id = static_cast<TKeyFrameID>(-1);
		out_new_kf_info.created_edge_ids.clear();
		std::memset(&out_new_kf_info.optimize_results, 0, sizeof(orc_opt_info_t));
		const TKeyFrameID new_kf_id = alloc_keyframe();
		std::vector<TNewEdgeInfo> new_k2k_edge_ids;
		determine_kf2kf_edges_to_create(new_kf_id, obs, new_k2k_edge_ids);
		for (typename new_kf_observations_t::const_iterator it_obs = obs.begin(); it_obs != obs.end(); ++it_obs) {
			const double *fixed_rel_pos = it_obs->is_fixed ? it_obs->feat_rel_pos : NULL;
			const double *unk_rel_pos_initval = it_obs->is_unknown_with_init_val ? it_obs->feat_rel_pos : NULL;
			add_observation(new_kf_id, it_obs->feat_id, it_obs->obs_data, fixed_rel_pos, unk_rel_pos_initval);
		}
		if (run_local_optimization) {
			if (parameters.optimize_new_edges_alone) {
				if (!new_k2k_edge_ids.empty()) {
					const int old_kernel = parameters.use_robust_kernel;
					parameters.use_robust_kernel = parameters.use_robust_kernel_stage1;
					std::vector<size_t> k2f_edges_to_opt;
					std::vector<size_t> k2k_edges_to_opt(1);
					for (size_t i = 0; i < new_k2k_edge_ids.size(); i++) {
						if (new_k2k_edge_ids[i].has_approx_init_val) continue;
						k2k_edges_to_opt[0] = new_k2k_edge_ids[i].id;
						optimize_edges(k2k_edges_to_opt, k2f_edges_to_opt, out_new_kf_info.optimize_results_stg1);
					}
					parameters.use_robust_kernel = old_kernel;
				}
			}
			TOptimizeLocalAreaParams opt_params;
			optimize_local_area(new_kf_id, parameters.max_optimize_depth, out_new_kf_info.optimize_results, opt_params);
		}
		out_new_kf_info.kf_id = new_kf_id;
		out_new_kf_info.created_edge_ids.swap(new_k2k_edge_ids);
	}

	// ------------------------------------------------------------------ impl/optimize_local_area.h:15-58 + impl/bfs_visitor.h:21-177
	// with VisitorOptimizeLocalArea (RbaEngine.h:543-618) inlined.
	void select_local_area(TKeyFrameID root_id, unsigned int win_size, const TOptimizeLocalAreaParams &params, std::vector<size_t> &k2k_edges_to_optimize, std::vector<size_t> &lm_IDs_to_optimize) const {
		const bool use_prebuilt_st = (win_size <= parameters.max_tree_depth);
		std::map<TLandmarkID, size_t> lm_times_seen;
		std::set<TLandmarkID> lm_visited;
		std::set<const k2k_edge_t *> k2k_visited;
		std::set<const k2f_edge_t *> k2f_visited;
		struct V {
			static void k2f(const k2f_edge_t *edge, const TOptimizeLocalAreaParams &params, std::map<TLandmarkID, size_t> &lm_times_seen, std::vector<size_t> &out) {
				if (!edge->feat_has_known_rel_pos) {
					const TLandmarkID lm_ID = edge->feat_id;
					if (++lm_times_seen[lm_ID] == params.dont_optimize_landmarks_seen_less_than_n_times) out.push_back(lm_ID);
				}
			}
		};
		const topo_dist_t max_distance = win_size;
		if (!use_prebuilt_st) {  // bfs_visitor.h:35-104
			std::set<TKeyFrameID> kf_visited;
			std::queue<TKeyFrameID> pending;
			std::map<TKeyFrameID, topo_dist_t> distances;
			pending.push(root_id);
			kf_visited.insert(root_id);
			distances[root_id] = 0;
			while (!pending.empty()) {
				const TKeyFrameID next_kf = pending.front();
				pending.pop();
				const topo_dist_t cur_dist = distances[next_kf];
				const keyframe_info &kfi = keyframes[next_kf];
				for (size_t i = 0; i < kfi.adjacent_k2f_edges.size(); i++) {
					const k2f_edge_t *ed = kfi.adjacent_k2f_edges[i];
					lm_visited.insert(ed->feat_id);
					if (!k2f_visited.count(ed)) {
						if (params.optimize_landmarks) V::k2f(ed, params, lm_times_seen, lm_IDs_to_optimize);
						k2f_visited.insert(ed);
					}
				}
				if (cur_dist >= max_distance) continue;
				for (size_t i = 0; i < kfi.adjacent_k2k_edges.size(); i++) {
					const k2k_edge_t *ed = kfi.adjacent_k2k_edges[i];
					const TKeyFrameID new_kf = other_end(next_kf, *ed);
					if (!kf_visited.count(new_kf)) {
						if (new_kf <= params.max_visitable_kf_id) { pending.push(new_kf); distances[new_kf] = cur_dist + 1; }
						kf_visited.insert(new_kf);
					}
					if (!k2k_visited.count(ed)) {
						if (params.optimize_k2k_edges) k2k_edges_to_optimize.push_back(ed->id);
						k2k_visited.insert(ed);
					}
				}
			}
		} else {  // bfs_visitor.h:105-176
			const std::map<TKeyFrameID, TSpanTreeEntry> *root_ST = spanning_tree.sym.next_edge.find(root_id);
			if (!root_ST) return;
			std::vector<std::pair<TKeyFrameID, topo_dist_t> > KFs;
			KFs.push_back(std::make_pair(root_id, topo_dist_t(0)));
			for (typename std::map<TKeyFrameID, TSpanTreeEntry>::const_iterator it = root_ST->begin(); it != root_ST->end(); ++it)
				KFs.push_back(std::make_pair(it->first, it->second.distance));
			for (size_t i = 0; i < KFs.size(); i++) {
				const TKeyFrameID kf_id = KFs[i].first;
				const size_t cur_dist = KFs[i].second;
				if (cur_dist > max_distance) continue;
				const keyframe_info &kfi = keyframes[kf_id];
				for (size_t k = 0; k < kfi.adjacent_k2f_edges.size(); k++) {
					const k2f_edge_t *ed = kfi.adjacent_k2f_edges[k];
					lm_visited.insert(ed->feat_id);
					if (!k2f_visited.count(ed)) {
						if (params.optimize_landmarks) V::k2f(ed, params, lm_times_seen, lm_IDs_to_optimize);
						k2f_visited.insert(ed);
					}
				}
				for (size_t k = 0; k < kfi.adjacent_k2k_edges.size(); k++) {
					const k2k_edge_t *ed = kfi.adjacent_k2k_edges[k];
					if (!k2k_visited.count(ed)) {
						if (params.optimize_k2k_edges) k2k_edges_to_optimize.push_back(ed->id);
						k2k_visited.insert(ed);
					}
				}
			}
		}
	}
	void optimize_local_area(TKeyFrameID root_id, unsigned int win_size, orc_opt_info_t &out_info, const TOptimizeLocalAreaParams &params = TOptimizeLocalAreaParams(),
	                         const std::vector<size_t> &observation_indices_to_optimize = std::vector<size_t>()) {  // RbaEngine.h:235-241
		std::vector<size_t> k2k_edges_to_optimize, lm_IDs_to_optimize;
		select_local_area(root_id, win_size, params, k2k_edges_to_optimize, lm_IDs_to_optimize);
		if (!k2k_edges_to_optimize.empty() || !lm_IDs_to_optimize.empty())
			optimize_edges(k2k_edges_to_optimize, lm_IDs_to_optimize, out_info, observation_indices_to_optimize);  // optimize_local_area.h:54
	}

	// ------------------------------------------------------------------ impl/jacobians.h
	static void add_edge_ij_to_list_needed_roots(std::set<TKeyFrameID> &lst, TKeyFrameID i, TKeyFrameID j) { lst.insert(i); lst.insert(j); }
	// :1020-1076
	void prepare_Jacobians_required_tree_roots(std::set<TKeyFrameID> &lst, const std::vector<colAp_t *> &lst_JacobCols_dAp, const std::vector<colF_t *> &lst_JacobCols_df) const {
		lst.clear();
		for (size_t i = 0; i < lst_JacobCols_dAp.size(); i++)
			for (typename colAp_t::const_iterator it = lst_JacobCols_dAp[i]->begin(); it != lst_JacobCols_dAp[i]->end(); ++it) {
				const TKeyFrameID obs_id = all_observations[it->first].kf_id;
				add_edge_ij_to_list_needed_roots(lst, it->second.sym.kf_d, obs_id);
				add_edge_ij_to_list_needed_roots(lst, it->second.sym.kf_base, it->second.sym.kf_d);
			}
		for (size_t i = 0; i < lst_JacobCols_df.size(); i++)
			for (typename colF_t::const_iterator it = lst_JacobCols_df[i]->begin(); it != lst_JacobCols_df[i]->end(); ++it) {
				const k2f_edge_t &k2f = all_observations[it->first];
				ORC_ASSERT(k2f.feat_rel_pos);
				add_edge_ij_to_list_needed_roots(lst, k2f.feat_rel_pos->id_frame_base, k2f.kf_id);
			}
	}

	// compute_jacobian_dAepsDx_deps<jacob_point_landmark,3,6>  (:363-494)
	void dAepsDx_deps(double *jacob, const double *dh_dx, bool is_inverse_edge_jacobian, const double *xji_i, const Pose3 *A, const Pose3 &Dp, const Pose3 &p_d_d1) const {
		Pose3 D = Dp;
		double H_ROTA[MAX_OBS * 3];
		if (!is_inverse_edge_jacobian) {
			if (A) mat_mul(H_ROTA, dh_dx, O, 3, A->R, 3);
			else std::memcpy(H_ROTA, dh_dx, sizeof(double) * O * 3);
		} else {
			D.composeFrom(p_d_d1, Dp);
			double Rt[9];
			for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Rt[3 * r + c] = p_d_d1.R[3 * c + r];
			if (A) { double T[MAX_OBS * 3]; mat_mul(T, dh_dx, O, 3, A->R, 3); mat_mul(H_ROTA, T, O, 3, Rt, 3); }
			else mat_mul(H_ROTA, dh_dx, O, 3, Rt, 3);
		}
		double v[3];
		for (int r = 0; r < 3; r++) v[r] = -D.t[r] - xji_i[0] * D.R[3 * r + 0] - xji_i[1] * D.R[3 * r + 1] - xji_i[2] * D.R[3 * r + 2];
		const double aux[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
		double second[MAX_OBS * 3];
		mat_mul(second, H_ROTA, O, 3, aux, 3);
		for (int r = 0; r < O; r++)
			for (int c = 0; c < 3; c++) { jacob[r * 6 + c] = H_ROTA[3 * r + c]; jacob[r * 6 + 3 + c] = second[3 * r + c]; }
		if (is_inverse_edge_jacobian) for (int i = 0; i < O * 6; i++) jacob[i] = -jacob[i];
	}
	// compute_jacobian_dAepsDx_deps<jacob_relpose_landmark,3,3>  (:644-744)
	void dAepsDx_deps(double *jacob, const double *dh_dx, bool is_inverse_edge_jacobian, const double *, const Pose2 *A, const Pose2 &Dp, const Pose2 &p_d_d1) const {
		double Xd, Yd, PHIa;
		Pose2 base_wrt_obs;
		if (!is_inverse_edge_jacobian) {
			Xd = Dp.x; Yd = Dp.y; PHIa = 0;
			if (A) { PHIa = A->phi; base_wrt_obs.composeFrom(*A, Dp); }
			else base_wrt_obs = Dp;
		} else {
			Pose2 pose_base_wrt_d1_prime;
			pose_base_wrt_d1_prime.composeFrom(p_d_d1, Dp);
			const Pose2 p_d_d1_inv = -p_d_d1;
			const Pose2 A_prime = A ? (*A + p_d_d1_inv) : p_d_d1_inv;
			Xd = pose_base_wrt_d1_prime.x; Yd = pose_base_wrt_d1_prime.y;
			PHIa = A_prime.phi;
			base_wrt_obs.composeFrom(A_prime, pose_base_wrt_d1_prime);
		}
		const double PHIad = base_wrt_obs.phi;
		const double ccos_ad = std::cos(PHIad), ssin_ad = std::sin(PHIad);
		const double ccos_a = std::cos(PHIa), ssin_a = std::sin(PHIa);
		const double J0[9] = {ccos_ad, ssin_ad, 0, -ssin_ad, ccos_ad, 0, 0, 0, 1};
		const double J1[9] = {1, 0, -Xd * ssin_a - Yd * ccos_a, 0, 1, Xd * ccos_a - Yd * ssin_a, 0, 0, 1};
		const double J2[9] = {ccos_a, -ssin_a, 0, ssin_a, ccos_a, 0, 0, 0, 1};
		double T0[9], T1[9];
		mat_mul(T0, dh_dx, 3, 3, J0, 3);
		mat_mul(T1, T0, 3, 3, J1, 3);
		mat_mul(jacob, T1, 3, 3, J2, 3);
		if (is_inverse_edge_jacobian) for (int i = 0; i < 9; i++) jacob[i] = -jacob[i];
	}
	static void mat_mul(double *C, const double *A, int n, int k, const double *B, int m) {
		for (int r = 0; r < n; r++)
			for (int c = 0; c < m; c++) {
				double s = 0;
				for (int t = 0; t < k; t++) s += A[r * k + t] * B[t * m + c];
				C[r * m + c] = s;
			}
	}

	// compute_jacobian_dh_dp (:207-354)
	void compute_jacobian_dh_dp(EntryAp &jacob, const k2f_edge_t &observation, std::vector<const pose_flag_t *> *out_list_of_required_num_poses) const {
		ORC_ASSERT(observation.kf_id != jacob.sym.kf_base);
		if (!*jacob.sym.is_valid) return;
		const pose_flag_t *pose_d1_wrt_obs = jacob.sym.rel_pose_d1_from_obs;
		const pose_flag_t &pose_base_wrt_d1 = *jacob.sym.rel_pose_base_from_d1;
		const bool is_inverse_edge_jacobian = !jacob.sym.edge_normal_dir;
		if (out_list_of_required_num_poses) {
			if (jacob.sym.rel_pose_d1_from_obs) out_list_of_required_num_poses->push_back(jacob.sym.rel_pose_d1_from_obs);
			out_list_of_required_num_poses->push_back(jacob.sym.rel_pose_base_from_d1);
		}
		ORC_ASSERT(pose_base_wrt_d1.updated);
		if (pose_d1_wrt_obs) ORC_ASSERT(pose_d1_wrt_obs->updated);
		pose_t pose_i_wrt_l;
		if (pose_d1_wrt_obs != NULL) pose_i_wrt_l.composeFrom(pose_d1_wrt_obs->pose, pose_base_wrt_d1.pose);
		else pose_i_wrt_l = pose_base_wrt_d1.pose;
		const double *xji_i = jacob.sym.feat_rel_pos->pos;
		double xji_l[3] = {xji_i[0], xji_i[1], xji_i[2]};
		composePosePoint(xji_l, pose_i_wrt_l);
		double dh_dx[MAX_OBS * 3];
		if (!is_relpose_family()) point_robot2sensor(xji_l);
		if (!eval_jacob_dh_dx(dh_dx, xji_l)) {
			*jacob.sym.is_valid = 0;
			std::memset(jacob.num, 0, sizeof(jacob.num));
			return;
		}
		jacob_dh_dx_rotate(dh_dx);
		dAepsDx_deps(jacob.num, dh_dx, is_inverse_edge_jacobian, xji_i, pose_d1_wrt_obs ? &pose_d1_wrt_obs->pose : NULL, pose_base_wrt_d1.pose, k2k_edges[jacob.sym.k2k_edge_id].inv_pose);
	}
	static const double *rotation_of(const Pose3 &p) { return p.R; }
	static const double *rotation_of(const Pose2 &) { static const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; return I3; }
	// compute_jacobian_dh_df (:888-1013)
	void compute_jacobian_dh_df(EntryF &jacob, std::vector<const pose_flag_t *> *out_list_of_required_num_poses) const {
		if (!*jacob.sym.is_valid) return;
		const pose_flag_t *rel_pose_base_from_obs = jacob.sym.rel_pose_base_from_obs;
		if (out_list_of_required_num_poses && rel_pose_base_from_obs) out_list_of_required_num_poses->push_back(rel_pose_base_from_obs);
		if (rel_pose_base_from_obs) ORC_ASSERT(rel_pose_base_from_obs->updated);
		const double *xji_i = jacob.sym.feat_rel_pos->pos;
		double xji_l[3] = {xji_i[0], xji_i[1], xji_i[2]};
		if (rel_pose_base_from_obs != NULL) composePosePoint(xji_l, rel_pose_base_from_obs->pose);
		double dh_dx[MAX_OBS * 3];
		point_robot2sensor(xji_l);
		if (!eval_jacob_dh_dx(dh_dx, xji_l)) {
			*jacob.sym.is_valid = 0;
			std::memset(jacob.num, 0, sizeof(jacob.num));
			return;
		}
		jacob_dh_dx_rotate(dh_dx);
		if (rel_pose_base_from_obs != NULL) mat_mul(jacob.num, dh_dx, O, 3, rotation_of(rel_pose_base_from_obs->pose), 3);
		else std::memcpy(jacob.num, dh_dx, sizeof(double) * O * 3);
	}
	// recompute_all_Jacobians (:1083-1117)
	size_t recompute_all_Jacobians(std::vector<colAp_t *> &lst_JacobCols_dAp, std::vector<colF_t *> &lst_JacobCols_df, std::vector<const pose_flag_t *> *out_list_of_required_num_poses = NULL) {
		size_t nJacobs = 0;
		if (out_list_of_required_num_poses) out_list_of_required_num_poses->clear();
		for (size_t i = 0; i < lst_JacobCols_dAp.size(); i++)
			for (typename colAp_t::iterator it = lst_JacobCols_dAp[i]->begin(); it != lst_JacobCols_dAp[i]->end(); ++it) {
				compute_jacobian_dh_dp(it->second, all_observations[it->first], out_list_of_required_num_poses);
				nJacobs++;
			}
		if (!is_relpose_family())  // :72-84: never evaluated for jacob_relpose_landmark
			for (size_t i = 0; i < lst_JacobCols_df.size(); i++)
				for (typename colF_t::iterator it = lst_JacobCols_df[i]->begin(); it != lst_JacobCols_df[i]->end(); ++it) {
					compute_jacobian_dh_df(it->second, out_list_of_required_num_poses);
					nJacobs++;
				}
		return nJacobs;
	}

	// ------------------------------------------------------------------ impl/reprojection_residuals.h:16-81
	struct TObsUsed { size_t obs_idx; k2f_edge_t *k2f; };
	static double huber_kernel(double delta, const double kernel_param) {  // RbaEngine.h:810-813
		return std::abs(2 * (kernel_param * kernel_param) * (std::sqrt(1 + (delta / kernel_param) * (delta / kernel_param)) - 1));
	}
	double reprojection_residuals(std::vector<double> &residuals, const std::vector<TObsUsed> &observations) const {
		const size_t nObs = observations.size();
		if (residuals.size() != nObs * O) residuals.resize(nObs * O);
		double total_sqr_err = 0;
		for (size_t i = 0; i < nObs; i++) {
			const TKeyFrameID obs_frame_id = observations[i].k2f->kf_id;
			const TRelativeLandmarkPos *feat_rel_pos = observations[i].k2f->feat_rel_pos;
			const TKeyFrameID base_id = feat_rel_pos->id_frame_base;
			pose_t const *base_pose_wrt_observer = NULL;
			if (base_id == obs_frame_id) base_pose_wrt_observer = &aux_null_pose;
			else {
				const frameid2pose_map_t *m = spanning_tree.num.find(obs_frame_id);
				ORC_ASSERT(m != NULL);
				typename frameid2pose_map_t::const_iterator itRelPose = m->find(base_id);
				ORC_ASSERT(itRelPose != m->end());
				base_pose_wrt_observer = &itRelPose->second.pose;
			}
			double *delta = &residuals[i * O];
			observe_error(delta, observations[i].k2f->obs_arr, *base_pose_wrt_observer, feat_rel_pos->pos);
			double sum_2 = 0;
			for (int k = 0; k < O; k++) sum_2 += delta[k] * delta[k];
			if (parameters.use_robust_kernel) {
				const double nrm = std::max(1e-11, std::sqrt(sum_2));
				const double w = std::sqrt(huber_kernel(nrm, parameters.kernel_param)) / nrm;
				for (int k = 0; k < O; k++) delta[k] *= w;
				total_sqr_err += (w * w) * sum_2;
			} else total_sqr_err += sum_2;
		}
		return total_sqr_err;
	}

	// ------------------------------------------------------------------ impl/compute_minus_gradient.h:20-91
	void compute_minus_gradient(std::vector<double> &minus_grad, const std::vector<colAp_t *> &sparse_jacobs_Ap, const std::vector<colF_t *> &sparse_jacobs_f,
	                            const std::vector<double> &residuals, const std::map<size_t, size_t> &obs_global_idx2residual_idx) const {
		const size_t nUnknowns_k2k = sparse_jacobs_Ap.size(), nUnknowns_k2f = sparse_jacobs_f.size();
		const size_t idx_start_f = P * nUnknowns_k2k;
		const size_t nUnknowns_scalars = P * nUnknowns_k2k + LM_DIMS * nUnknowns_k2f;
		if (minus_grad.size() != nUnknowns_scalars) minus_grad.resize(nUnknowns_scalars);
		for (size_t i = 0; i < nUnknowns_k2k; i++) {
			double accum_g_i[MAX_P] = {0, 0, 0, 0, 0, 0};
			for (typename colAp_t::const_iterator itJ = sparse_jacobs_Ap[i]->begin(); itJ != sparse_jacobs_Ap[i]->end(); ++itJ) {
				std::map<size_t, size_t>::const_iterator it_obs = obs_global_idx2residual_idx.find(itJ->first);
				ORC_ASSERT(it_obs != obs_global_idx2residual_idx.end());
				noise.accum_Jtr(accum_g_i, itJ->second.num, P, &residuals[it_obs->second * O]);
			}
			noise.scale_Jtr(accum_g_i, P);
			for (int k = 0; k < P; k++) minus_grad[i * P + k] = accum_g_i[k];
		}
		for (size_t i = 0; i < nUnknowns_k2f; i++) {
			double accum_g_i[LM_DIMS] = {0, 0, 0};
			for (typename colF_t::const_iterator itJ = sparse_jacobs_f[i]->begin(); itJ != sparse_jacobs_f[i]->end(); ++itJ) {
				std::map<size_t, size_t>::const_iterator it_obs = obs_global_idx2residual_idx.find(itJ->first);
				ORC_ASSERT(it_obs != obs_global_idx2residual_idx.end());
				noise.accum_Jtr(accum_g_i, itJ->second.num, LM_DIMS, &residuals[it_obs->second * O]);
			}
			noise.scale_Jtr(accum_g_i, LM_DIMS);
			for (int k = 0; k < LM_DIMS; k++) minus_grad[idx_start_f + i * LM_DIMS + k] = accum_g_i[k];
		}
	}

	// ------------------------------------------------------------------ impl/optimize_edges.h:44-793
	void optimize_edges(const std::vector<size_t> &run_k2k_edges_in, const std::vector<size_t> &run_feat_ids_in, orc_opt_info_t &out_info,
	                    const std::vector<size_t> &in_observation_indices_to_optimize = std::vector<size_t>()) {
		const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();  // m_profiler.enter("opt") :56
		struct Leave {
			Engine *e; std::chrono::steady_clock::time_point t0;
			~Leave() { e->total_opt_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
		} leave = {this, t0};
		(void)leave;
		std::memset(&out_info, 0, sizeof(out_info));  // out_info.clear() :58
		const size_t POSE_DIMS = P;
		const size_t OBS_DIMS = O;

		// A. filter unknowns (:71-119)
		std::vector<size_t> run_k2k_edges, run_feat_ids;
		std::set<TKeyFrameID> touched_KFs;
		std::set<TLandmarkID> touched_LMs;
		for (size_t i = 0; i < run_k2k_edges_in.size(); i++) {
			ORC_ASSERT(run_k2k_edges_in[i] < lin_system.dh_dAp.size());
			const colAp_t &col_i = lin_system.dh_dAp[run_k2k_edges_in[i]];
			if (!col_i.empty()) {
				run_k2k_edges.push_back(run_k2k_edges_in[i]);
				touched_KFs.insert(k2k_edges[run_k2k_edges_in[i]].from);
				touched_KFs.insert(k2k_edges[run_k2k_edges_in[i]].to);
			}
		}
		for (size_t i = 0; i < run_feat_ids_in.size(); i++) {
			const TLandmarkID feat_id = run_feat_ids_in[i];
			ORC_ASSERT(feat_id < all_lms.size());
			const TLandmarkEntry &lm_e = all_lms[feat_id];
			ORC_ASSERT(lm_e.rfp != NULL);
			ORC_ASSERT(!lm_e.has_known_pos);
			ORC_ASSERT(feat_id < lin_system.dh_df_remap.size() && lin_system.dh_df_remap[feat_id] != SRBA_INVALID_INDEX);
			const colF_t &col_i = lin_system.dh_df[lin_system.dh_df_remap[feat_id]];
			if (!col_i.empty()) { run_feat_ids.push_back(feat_id); touched_LMs.insert(feat_id); }
		}
		// B. (:125-163)
		const size_t nUnknowns_k2k = run_k2k_edges.size(), nUnknowns_k2f = run_feat_ids.size();
		out_info.num_kf_optimized = touched_KFs.size();
		out_info.num_lm_optimized = touched_LMs.size();
		const size_t idx_start_f = POSE_DIMS * nUnknowns_k2k;
		const size_t nUnknowns_scalars = POSE_DIMS * nUnknowns_k2k + LM_DIMS * nUnknowns_k2f;
		if (!nUnknowns_scalars) return;  // :132-138
		std::vector<colAp_t *> dh_dAp(nUnknowns_k2k);
		std::vector<k2k_edge_t *> k2k_edge_unknowns(nUnknowns_k2k);
		for (size_t i = 0; i < nUnknowns_k2k; i++) { dh_dAp[i] = &lin_system.dh_dAp[run_k2k_edges[i]]; k2k_edge_unknowns[i] = &k2k_edges[run_k2k_edges[i]]; }
		std::vector<colF_t *> dh_df(nUnknowns_k2f);
		std::vector<TRelativeLandmarkPos *> k2f_edge_unknowns(nUnknowns_k2f);
		for (size_t i = 0; i < nUnknowns_k2f; i++) {
			dh_df[i] = &lin_system.dh_df[lin_system.dh_df_remap[run_feat_ids[i]]];
			k2f_edge_unknowns[i] = all_lms[run_feat_ids[i]].rfp;
		}
		// C. involved observations (:171-234).  NOTE the dAp loop does NOT de-duplicate.
		std::map<size_t, size_t> obs_global_idx2residual_idx;
		std::vector<TObsUsed> involved_obs;
		if (!in_observation_indices_to_optimize.empty()) {  // :218-234: use only those observations, in the caller's order
			for (size_t i = 0; i < in_observation_indices_to_optimize.size(); i++) {
				const size_t global_obs_idx = in_observation_indices_to_optimize[i];
				ORC_ASSERT(global_obs_idx < all_observations.size());
				obs_global_idx2residual_idx[global_obs_idx] = involved_obs.size();
				TObsUsed o = {global_obs_idx, &all_observations[global_obs_idx]};
				involved_obs.push_back(o);
			}
		} else
		for (size_t i = 0; i < nUnknowns_k2k; i++)
			for (typename colAp_t::iterator it = dh_dAp[i]->begin(); it != dh_dAp[i]->end(); ++it) {
				const size_t global_obs_idx = it->first;
				obs_global_idx2residual_idx[global_obs_idx] = involved_obs.size();
				TObsUsed o = {global_obs_idx, &all_observations[global_obs_idx]};
				involved_obs.push_back(o);
			}
		if (in_observation_indices_to_optimize.empty())
		for (size_t i = 0; i < nUnknowns_k2f; i++)
			for (typename colF_t::iterator it = dh_df[i]->begin(); it != dh_df[i]->end(); ++it) {
				const size_t global_obs_idx = it->first;
				if (obs_global_idx2residual_idx.find(global_obs_idx) == obs_global_idx2residual_idx.end()) {
					obs_global_idx2residual_idx[global_obs_idx] = involved_obs.size();
					TObsUsed o = {global_obs_idx, &all_observations[global_obs_idx]};
					involved_obs.push_back(o);
				}
			}
		const size_t nObs = involved_obs.size();
		// D/E (:245-256)
		std::set<TKeyFrameID> kfs_num_spantrees_to_update;
		prepare_Jacobians_required_tree_roots(kfs_num_spantrees_to_update, dh_dAp, dh_df);
		const size_t count_span_tree_num_update = update_numeric(kfs_num_spantrees_to_update, false);
		// F/G/H (:262-283)
		std::vector<const pose_flag_t *> list_of_required_num_poses;
		for (size_t i = 0; i < nObs; i++) all_observations_Jacob_validity[involved_obs[i].obs_idx] = 1;
		const size_t count_jacobians = recompute_all_Jacobians(dh_dAp, dh_df, &list_of_required_num_poses);
		for (size_t i = 0; i < list_of_required_num_poses.size(); i++) list_of_required_num_poses[i]->mark_outdated();
		// I/J (:303-331)
		SparseBlockHessian HAp, Hf, HApf;
		{
			std::vector<JacobColView> vAp(nUnknowns_k2k), vF(nUnknowns_k2f);
			for (size_t i = 0; i < nUnknowns_k2k; i++)
				for (typename colAp_t::const_iterator it = dh_dAp[i]->begin(); it != dh_dAp[i]->end(); ++it) {
					JacobBlockRef r = {it->second.num, it->second.sym.is_valid};
					vAp[i].push_back(std::make_pair(it->first, r));
				}
			for (size_t i = 0; i < nUnknowns_k2f; i++)
				for (typename colF_t::const_iterator it = dh_df[i]->begin(); it != dh_df[i]->end(); ++it) {
					JacobBlockRef r = {it->second.num, it->second.sym.is_valid};
					vF[i].push_back(std::make_pair(it->first, r));
				}
			sparse_hessian_build_symbolic(HAp, Hf, HApf, vAp, vF, P, LM_DIMS);
		}
		sparse_hessian_update_numeric(HAp, noise);
		sparse_hessian_update_numeric(Hf, noise);
		sparse_hessian_update_numeric(HApf, noise);
		// K (:355)
		if (OBS_DIMS * nObs < POSE_DIMS * nUnknowns_k2k + LM_DIMS * nUnknowns_k2f) throw std::runtime_error("optimize_edges: ASSERT_ABOVEEQ_(OBS_DIMS*nObs,nUnknowns) failed (rank-deficient problem)");
		// L (:361-390)
		double nu = 2;
		const double max_gradient_to_stop = 1e-15;
		double lambda = -1;
		{
			double Hess_diag_max = 0;
			for (size_t i = 0; i < nUnknowns_k2k; i++) {
				const double *H = HAp.getCol(i)[i].num;
				double m = H[0];
				for (int k = 1; k < P; k++) m = std::max(m, H[k * P + k]);
				Hess_diag_max = std::max(Hess_diag_max, m);
			}
			for (size_t i = 0; i < nUnknowns_k2f; i++) {
				const double *H = Hf.getCol(i)[i].num;
				double m = H[0];
				for (int k = 1; k < LM_DIMS; k++) m = std::max(m, H[k * LM_DIMS + k]);
				Hess_diag_max = std::max(Hess_diag_max, m);
			}
			const double tau = 1e-3;
			lambda = tau * Hess_diag_max;
		}
		// M (:395-412)
		std::vector<double> residuals(nObs * O);
		double total_proj_error = reprojection_residuals(residuals, involved_obs);
		double RMSE = std::sqrt(total_proj_error / nObs);
		out_info.num_observations = nObs;
		out_info.num_jacobians = count_jacobians;
		out_info.num_kf2kf_edges_optimized = run_k2k_edges.size();
		out_info.num_kf2lm_edges_optimized = run_feat_ids.size();
		out_info.num_total_scalar_optimized = nUnknowns_scalars;
		out_info.num_span_tree_numeric_updates = count_span_tree_num_update;
		out_info.total_sqr_error_init = total_proj_error;
		// N (:422-426)
		std::vector<double> minus_grad;
		compute_minus_gradient(minus_grad, dh_dAp, dh_df, residuals, obs_global_idx2residual_idx);

		OptTrace *tr = NULL;
		if (trace_on) {
			traces.push_back(OptTrace());
			tr = &traces.back();
			tr->k2k_ids.assign(run_k2k_edges.begin(), run_k2k_edges.end());
			tr->lm_ids.assign(run_feat_ids.begin(), run_feat_ids.end());
			for (size_t i = 0; i < nObs; i++) tr->obs_idx.push_back(involved_obs[i].obs_idx);
			tr->init_grad = minus_grad;
			tr->lambda0 = lambda;
			tr->n_scalars = nUnknowns_scalars;
		}
		// O (:431-436)
		std::unique_ptr<SolverBase> my_solver_ptr;  // my_solver_t = solver_engine<USE_SCHUR,DENSE_CHOLESKY> (:54)
		switch (parameters.solver_kind) {
			case SOLVER_LM_SCHUR_DENSE_CHOLESKY: my_solver_ptr.reset(new SolverSchurDense(HAp, Hf, HApf, minus_grad, nUnknowns_k2k, nUnknowns_k2f, P, LM_DIMS)); break;
			case SOLVER_LM_SCHUR_SPARSE_CHOLESKY: my_solver_ptr.reset(new SolverSchurSparse(HAp, Hf, HApf, minus_grad, nUnknowns_k2k, nUnknowns_k2f, P, LM_DIMS)); break;
			case SOLVER_LM_NO_SCHUR_SPARSE_CHOLESKY: my_solver_ptr.reset(new SolverNoSchurSparse(HAp, Hf, HApf, minus_grad, nUnknowns_k2k, nUnknowns_k2f, P, LM_DIMS)); break;
			default: throw std::runtime_error("optimize_edges: unknown solver_kind");
		}
		SolverBase &my_solver = *my_solver_ptr;
		const double MAX_LAMBDA = parameters.max_lambda;
		std::vector<k2k_edge_t> old_k2k_edge_unknowns;
		std::vector<pose_t> old_span_tree;
		std::vector<TRelativeLandmarkPos> old_k2f_edge_unknowns;
		// P (:452-696)
		size_t iter;
		bool stop = false;
		for (iter = 0; iter < parameters.max_iters && !stop; iter++) {
			double rho = 0;
			if (lambda >= MAX_LAMBDA) stop = true;
			if (RMSE < parameters.max_error_per_obs_to_stop) stop = true;
			while (rho <= 0 && !stop) {
				out_info.num_lm_trials++;
				total_opt_trials++;
				if (!my_solver.solve(lambda)) {
					if (tr) { const double s[6] = {lambda, 0, 0, 0, 0, 0}; tr->trial_scalars.insert(tr->trial_scalars.end(), s, s + 6); tr->deltas.insert(tr->deltas.end(), nUnknowns_scalars, 0.0); }
					lambda *= nu;
					nu *= 2.;
					stop = (lambda > MAX_LAMBDA);
					continue;
				}
				old_k2k_edge_unknowns.resize(nUnknowns_k2k);
				for (size_t i = 0; i < nUnknowns_k2k; i++) old_k2k_edge_unknowns[i] = *k2k_edge_unknowns[i];
				old_k2f_edge_unknowns.resize(nUnknowns_k2f);
				for (size_t i = 0; i < nUnknowns_k2f; i++) old_k2f_edge_unknowns[i] = *k2f_edge_unknowns[i];
				for (size_t i = 0; i < nUnknowns_k2k; i++) {  // :508-527
					const pose_t &old_pose = k2k_edge_unknowns[i]->inv_pose;
					const pose_t incrPose = pose_t::pseudo_exp(&my_solver.delta_eps[POSE_DIMS * i]);
					pose_t new_pose;
					new_pose.composeFrom(incrPose, old_pose);
					k2k_edge_unknowns[i]->inv_pose = new_pose;
				}
				for (size_t i = 0; i < nUnknowns_k2f; i++) {  // :534-539
					const double *delta_feat = &my_solver.delta_eps[idx_start_f + LM_DIMS * i];
					for (int k = 0; k < LM_DIMS; k++) k2f_edge_unknowns[i]->pos[k] += delta_feat[k];
				}
				{  // :550-565
					const size_t nReqNumPoses = list_of_required_num_poses.size();
					if (old_span_tree.size() != nReqNumPoses) old_span_tree.resize(nReqNumPoses);
					for (size_t i = 0; i < nReqNumPoses; i++) old_span_tree[i] = list_of_required_num_poses[i]->pose;
				}
				for (size_t i = 0; i < list_of_required_num_poses.size(); i++) list_of_required_num_poses[i]->mark_outdated();
				update_numeric(kfs_num_spantrees_to_update, true);
				std::vector<double> new_residuals;  // :570-579
				double new_total_proj_error = reprojection_residuals(new_residuals, involved_obs);
				const double new_RMSE = std::sqrt(new_total_proj_error / nObs);
				const double error_reduction_ratio = total_proj_error > 0 ? (total_proj_error - new_total_proj_error) / total_proj_error : 0;
				double denom = 0;  // :585
				for (size_t i = 0; i < nUnknowns_scalars; i++) denom += my_solver.delta_eps[i] * (lambda * my_solver.delta_eps[i] + minus_grad[i]);
				rho = (total_proj_error - new_total_proj_error) / denom;
				if (tr) {
					const double s[6] = {lambda, 1, rho, new_total_proj_error, rho > 0 ? 1.0 : 0.0, 0};
					tr->trial_scalars.insert(tr->trial_scalars.end(), s, s + 6);
					tr->deltas.insert(tr->deltas.end(), my_solver.delta_eps.begin(), my_solver.delta_eps.end());
				}
				if (rho > 0) {
					const bool do_relinearize = (error_reduction_ratio < 0 || error_reduction_ratio > parameters.min_error_reduction_ratio_to_relinearize);
					if (tr) tr->trial_scalars[tr->trial_scalars.size() - 1] = do_relinearize ? 1.0 : 0.0;
					residuals.swap(new_residuals);
					total_proj_error = new_total_proj_error;
					RMSE = new_RMSE;
					if (do_relinearize) {  // :606-629
						for (size_t i = 0; i < nObs; i++) all_observations_Jacob_validity[involved_obs[i].obs_idx] = 1;
						recompute_all_Jacobians(dh_dAp, dh_df);
						sparse_hessian_update_numeric(HAp, noise);
						sparse_hessian_update_numeric(Hf, noise);
						sparse_hessian_update_numeric(HApf, noise);
						my_solver.realize_relinearized();
					}
					compute_minus_gradient(minus_grad, dh_dAp, dh_df, residuals, obs_global_idx2residual_idx);  // :633
					double norm_inf_min_grad = 0;
					for (size_t i = 0; i < minus_grad.size(); i++) norm_inf_min_grad = std::max(norm_inf_min_grad, std::abs(minus_grad[i]));
					if (norm_inf_min_grad <= max_gradient_to_stop) stop = true;
					if (RMSE < parameters.max_error_per_obs_to_stop) stop = true;
					if (rho > parameters.max_rho) stop = true;
					lambda *= 1.0 / 3.0;
					nu = 2.0;
					my_solver.realize_lambda_changed();
				} else {  // :658-690
					for (size_t i = 0; i < list_of_required_num_poses.size(); i++) const_cast<pose_flag_t *>(list_of_required_num_poses[i])->pose = old_span_tree[i];
					for (size_t i = 0; i < nUnknowns_k2k; i++) *k2k_edge_unknowns[i] = old_k2k_edge_unknowns[i];
					for (size_t i = 0; i < nUnknowns_k2f; i++) *k2f_edge_unknowns[i] = old_k2f_edge_unknowns[i];
					lambda *= nu;
					nu *= 2.0;
					stop = (lambda > MAX_LAMBDA);
					my_solver.realize_lambda_changed();
				}
			}
		}
		// Q (:722-751)
		out_info.total_sqr_error_final = total_proj_error;
		out_info.num_iters = iter;
		unknown_lms_inf_matrices.clear();
		if (parameters.cov_recovery == 1 /* crpLandmarksApprox */ && nUnknowns_k2k != 0) {
			for (size_t i = 0; i < nUnknowns_k2f; i++) {
				if (!my_solver.was_ith_feature_invertible(i)) continue;
				const SparseBlockHessian::col_t &col_i = Hf.getCol(i);
				ORC_ASSERT(col_i.rbegin()->first == i);
				unknown_lms_inf_matrices[run_feat_ids[i]].assign(col_i.rbegin()->second.num, col_i.rbegin()->second.num + LM_DIMS * LM_DIMS);
			}
		}
		my_solver.get_extra_results(last_dense_hessian, last_hessian_valid);  // :769
		out_info.obs_rmse = RMSE;
	}
	std::vector<double> last_dense_hessian;
	bool last_hessian_valid;
};

}  // namespace srba_oracle
