// Host-side pose algebra used by the graph front-end of srba_b200::RbaEngine (edge bootstrap values and the
// landmark matchers); same conventions as MRPT's CPose2D / CPose3D that the reference calls
// (impl/determine_kf2kf_edges_to_create.h:50-62,198-243; models/observations_*.h landmark_matcher<>).
// Flat layouts match include/srba_b200.h: SE2 = x,y,phi ; SE3 = R row-major (9) + t (3).
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace srba_b200 {

inline double wrap_to_pi(double a) {
	a += M_PI;
	const bool neg = a < 0;
	a = std::fmod(a, 2.0 * M_PI);
	if (neg) a += 2.0 * M_PI;
	return a - M_PI;
}

struct SE2 {
	static const int P = 3, S = 3, KIND = 0;
	double v[3];
	SE2() { v[0] = v[1] = v[2] = 0; }
	SE2(double x, double y, double phi) { v[0] = x; v[1] = y; v[2] = phi; }
	static SE2 compose(const SE2 &a, const SE2 &b) {
		const double c = std::cos(a.v[2]), s = std::sin(a.v[2]);
		return SE2(a.v[0] + b.v[0] * c - b.v[1] * s, a.v[1] + b.v[0] * s + b.v[1] * c, wrap_to_pi(a.v[2] + b.v[2]));
	}
	SE2 inverse() const {
		const double c = std::cos(v[2]), s = std::sin(v[2]);
		return SE2(-v[0] * c - v[1] * s, v[0] * s - v[1] * c, -v[2]);
	}
	// a (-) b  == CPose2D::inverseComposeFrom
	static SE2 ominus(const SE2 &a, const SE2 &b) {
		const double c = std::cos(b.v[2]), s = std::sin(b.v[2]);
		return SE2((a.v[0] - b.v[0]) * c + (a.v[1] - b.v[1]) * s, -(a.v[0] - b.v[0]) * s + (a.v[1] - b.v[1]) * c, wrap_to_pi(a.v[2] - b.v[2]));
	}
};

struct SE3 {
	static const int P = 6, S = 12, KIND = 1;
	double v[12];
	SE3() { std::memset(v, 0, sizeof(v)); v[0] = v[4] = v[8] = 1; }
	static SE3 from_xyzypr(const double *q) {
		SE3 r;
		const double cy = std::cos(q[3]), sy = std::sin(q[3]), cp = std::cos(q[4]), sp = std::sin(q[4]), cr = std::cos(q[5]), sr = std::sin(q[5]);
		const double R[9] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr, sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr, -sp, cp * sr, cp * cr};
		std::memcpy(r.v, R, sizeof(R));
		r.v[9] = q[0]; r.v[10] = q[1]; r.v[11] = q[2];
		return r;
	}
	static SE3 from_xyzquat(const double *q) {
		SE3 r;
		const double qr = q[3], qx = q[4], qy = q[5], qz = q[6];
		const double R[9] = {qr * qr + qx * qx - qy * qy - qz * qz, 2 * (qx * qy - qr * qz), 2 * (qz * qx + qr * qy),
		                     2 * (qx * qy + qr * qz), qr * qr - qx * qx + qy * qy - qz * qz, 2 * (qy * qz - qr * qx),
		                     2 * (qz * qx - qr * qy), 2 * (qy * qz + qr * qx), qr * qr - qx * qx - qy * qy + qz * qz};
		std::memcpy(r.v, R, sizeof(R));
		r.v[9] = q[0]; r.v[10] = q[1]; r.v[11] = q[2];
		return r;
	}
	static SE3 compose(const SE3 &a, const SE3 &b) {
		SE3 o;
		for (int r = 0; r < 3; r++) o.v[9 + r] = a.v[9 + r] + a.v[3 * r] * b.v[9] + a.v[3 * r + 1] * b.v[10] + a.v[3 * r + 2] * b.v[11];
		for (int r = 0; r < 3; r++)
			for (int c = 0; c < 3; c++) o.v[3 * r + c] = a.v[3 * r] * b.v[c] + a.v[3 * r + 1] * b.v[3 + c] + a.v[3 * r + 2] * b.v[6 + c];
		return o;
	}
	SE3 inverse() const {
		SE3 o;
		for (int r = 0; r < 3; r++)
			for (int c = 0; c < 3; c++) o.v[3 * r + c] = v[3 * c + r];
		for (int r = 0; r < 3; r++) o.v[9 + r] = -(o.v[3 * r] * v[9] + o.v[3 * r + 1] * v[10] + o.v[3 * r + 2] * v[11]);
		return o;
	}
	static SE3 ominus(const SE3 &a, const SE3 &b) { return compose(b.inverse(), a); }
	void compose_point(const double *l, double *g) const {
		const double x = v[0] * l[0] + v[1] * l[1] + v[2] * l[2] + v[9], y = v[3] * l[0] + v[4] * l[1] + v[5] * l[2] + v[10],
		             z = v[6] * l[0] + v[7] * l[1] + v[8] * l[2] + v[11];
		g[0] = x; g[1] = y; g[2] = z;
	}
};

inline SE3 to_se3(const SE3 &p) { return p; }
inline SE3 to_se3(const SE2 &p) { const double q[6] = {p.v[0], p.v[1], 0, p.v[2], 0, 0}; return SE3::from_xyzypr(q); }
inline void from_se3(const SE3 &p, SE3 &o) { o = p; }
inline void from_se3(const SE3 &p, SE2 &o) { o = SE2(p.v[9], p.v[10], std::atan2(p.v[3], p.v[0])); }

// Closed-form absolute orientation (Horn 1987, unit quaternions): finds T with  a_i ~= T (+) b_i.  Stands in for the
// un-vendored mrpt::tfest::se3_l2 used by landmark_matcher<Cartesian_3D|StereoCamera> (observations_Cartesian_3D.h:73-78).
inline bool absolute_orientation(const std::vector<double> &a, const std::vector<double> &b, SE3 &out) {
	const size_t N = a.size() / 3;
	if (N < 3) return false;
	double ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
	for (size_t i = 0; i < N; i++)
		for (int k = 0; k < 3; k++) { ca[k] += a[3 * i + k]; cb[k] += b[3 * i + k]; }
	for (int k = 0; k < 3; k++) { ca[k] /= N; cb[k] /= N; }
	double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
	for (size_t i = 0; i < N; i++)
		for (int r = 0; r < 3; r++)
			for (int c = 0; c < 3; c++) M[3 * r + c] += (b[3 * i + r] - cb[r]) * (a[3 * i + c] - ca[c]);
	double K[16] = {M[0] + M[4] + M[8], M[5] - M[7], M[6] - M[2], M[1] - M[3],
	                M[5] - M[7], M[0] - M[4] - M[8], M[1] + M[3], M[6] + M[2],
	                M[6] - M[2], M[1] + M[3], -M[0] + M[4] - M[8], M[5] + M[7],
	                M[1] - M[3], M[6] + M[2], M[5] + M[7], -M[0] - M[4] + M[8]};
	double V[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
	for (int sweep = 0; sweep < 100; sweep++) {  // cyclic Jacobi on the symmetric 4x4
		double off = 0;
		for (int r = 0; r < 4; r++)
			for (int c = r + 1; c < 4; c++) off += K[4 * r + c] * K[4 * r + c];
		if (off < 1e-300) break;
		for (int p = 0; p < 4; p++)
			for (int q = p + 1; q < 4; q++) {
				if (K[4 * p + q] == 0) continue;
				const double th = (K[4 * q + q] - K[4 * p + p]) / (2 * K[4 * p + q]);
				const double t = (th >= 0 ? 1.0 : -1.0) / (std::abs(th) + std::sqrt(th * th + 1));
				const double c = 1 / std::sqrt(t * t + 1), s = t * c;
				for (int k = 0; k < 4; k++) { const double x = K[4 * k + p], y = K[4 * k + q]; K[4 * k + p] = c * x - s * y; K[4 * k + q] = s * x + c * y; }
				for (int k = 0; k < 4; k++) { const double x = K[4 * p + k], y = K[4 * q + k]; K[4 * p + k] = c * x - s * y; K[4 * q + k] = s * x + c * y; }
				for (int k = 0; k < 4; k++) { const double x = V[4 * k + p], y = V[4 * k + q]; V[4 * k + p] = c * x - s * y; V[4 * k + q] = s * x + c * y; }
			}
	}
	int best = 0;
	for (int k = 1; k < 4; k++) if (K[4 * k + k] > K[4 * best + best]) best = k;
	double q[7] = {0, 0, 0, V[best], V[4 + best], V[8 + best], V[12 + best]};
	const double nq = std::sqrt(q[3] * q[3] + q[4] * q[4] + q[5] * q[5] + q[6] * q[6]);
	for (int k = 3; k < 7; k++) q[k] /= nq;
	if (q[3] < 0) for (int k = 3; k < 7; k++) q[k] = -q[k];
	out = SE3::from_xyzquat(q);
	for (int r = 0; r < 3; r++) out.v[9 + r] = ca[r] - (out.v[3 * r] * cb[0] + out.v[3 * r + 1] * cb[1] + out.v[3 * r + 2] * cb[2]);
	return true;
}

}  // namespace srba_b200
