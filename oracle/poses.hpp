// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// MRPT-lite pose arithmetic: a CPU restatement of the (un-vendored) MRPT 1.3.x
// pose classes that the reference's hot path calls.  MRPT is NOT present under
// /root/reference (README.md:13, srba-config.cmake.in:32), so these follow MRPT's
// published semantics and are anchored on the reference's
// own call sites:
//   CPose2D/CPose3D::composeFrom, unary '-', binary '-'  : impl/spantree_update_numeric.h:51,59,72
//   composePoint / inverseComposePoint                   : models/sensors.h:58,183,358; srba_options_sensor_pose.h:112-120
//   SE_traits<2|3>::pseudo_exp                           : impl/optimize_edges.h:517
//   CPose3DQuat (stereo right camera)                    : models/sensors.h:193,259
// Parity status: "parity unpinned" at the value level (no MRPT binary to compare with);
// pinned at the property level by tests/test_oracle_spantree.py (reference
// tests/spantree_unittest.cpp:208,216) and the analytic-vs-finite-difference Jacobian tests.
#pragma once
#include <cmath>
#include <cstring>

namespace srba_oracle {

// mrpt::math::wrapTo2Pi / wrapToPi (mrpt/math/wrap2pi.h)
inline double wrapTo2Pi(double a) {
	const bool was_neg = a < 0;
	a = std::fmod(a, 2.0 * M_PI);
	if (was_neg) a += 2.0 * M_PI;
	return a;
}
inline double wrapToPi(double a) { return wrapTo2Pi(a + M_PI) - M_PI; }

// ---------------------------------------------------------------- SE(2)
struct Pose2 {
	static const int DIMS = 3;       // kf2kf_poses::SE2::REL_POSE_DIMS (models/kf2kf_poses.h:26-31)
	static const int NVALS = 3;      // doubles in the flat representation
	double x, y, phi;
	Pose2() : x(0), y(0), phi(0) {}
	Pose2(double x_, double y_, double phi_) : x(x_), y(y_), phi(phi_) {}

	// this = A (+) B.  Safe when &A==this or &B==this.
	void composeFrom(const Pose2 &A, const Pose2 &B) {
		const double ccos = std::cos(A.phi), ssin = std::sin(A.phi);
		const double nx = A.x + B.x * ccos - B.y * ssin;
		const double ny = A.y + B.x * ssin + B.y * ccos;
		const double nphi = wrapToPi(A.phi + B.phi);
		x = nx; y = ny; phi = nphi;
	}
	void inverse() {
		const double ccos = std::cos(phi), ssin = std::sin(phi);
		const double ox = x, oy = y;
		x = -ox * ccos - oy * ssin;
		y = ox * ssin - oy * ccos;
		phi = -phi;
	}
	Pose2 operator-() const { Pose2 r = *this; r.inverse(); return r; }
	Pose2 operator+(const Pose2 &B) const { Pose2 r; r.composeFrom(*this, B); return r; }
	// A - B = (-B) (+) A   (CPose2D::inverseComposeFrom)
	Pose2 operator-(const Pose2 &B) const {
		Pose2 r;
		const double ccos = std::cos(B.phi), ssin = std::sin(B.phi);
		r.x = (x - B.x) * ccos + (y - B.y) * ssin;
		r.y = -(x - B.x) * ssin + (y - B.y) * ccos;
		r.phi = wrapToPi(phi - B.phi);
		return r;
	}
	void toFlat(double *o) const { o[0] = x; o[1] = y; o[2] = phi; }
	void fromFlat(const double *o) { x = o[0]; y = o[1]; phi = o[2]; }
	// SE_traits<2>::pseudo_exp : plain copy of (x,y,phi)
	static Pose2 pseudo_exp(const double *v) { return Pose2(v[0], v[1], v[2]); }
};

// ---------------------------------------------------------------- SE(3)
struct Pose3 {
	static const int DIMS = 6;       // kf2kf_poses::SE3::REL_POSE_DIMS (models/kf2kf_poses.h:33-38)
	static const int NVALS = 12;     // R row-major (9) + t (3)
	double R[9];
	double t[3];
	Pose3() { setIdentity(); }
	void setIdentity() {
		R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
		t[0] = t[1] = t[2] = 0;
	}
	// CPose3D(x,y,z,yaw,pitch,roll): R = Rz(yaw) Ry(pitch) Rx(roll)
	Pose3(double x, double y, double z, double yaw, double pitch, double roll) {
		t[0] = x; t[1] = y; t[2] = z;
		const double cy = std::cos(yaw), sy = std::sin(yaw);
		const double cp = std::cos(pitch), sp = std::sin(pitch);
		const double cr = std::cos(roll), sr = std::sin(roll);
		R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
		R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
		R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
	}
	void composeFrom(const Pose3 &A, const Pose3 &B) {
		double nt[3], nR[9];
		for (int r = 0; r < 3; r++)
			nt[r] = A.t[r] + A.R[3 * r + 0] * B.t[0] + A.R[3 * r + 1] * B.t[1] + A.R[3 * r + 2] * B.t[2];
		for (int r = 0; r < 3; r++)
			for (int c = 0; c < 3; c++)
				nR[3 * r + c] = A.R[3 * r + 0] * B.R[c] + A.R[3 * r + 1] * B.R[3 + c] + A.R[3 * r + 2] * B.R[6 + c];
		std::memcpy(t, nt, sizeof(nt));
		std::memcpy(R, nR, sizeof(nR));
	}
	void inverse() {
		double nR[9], nt[3];
		for (int r = 0; r < 3; r++)
			for (int c = 0; c < 3; c++) nR[3 * r + c] = R[3 * c + r];
		for (int r = 0; r < 3; r++) nt[r] = -(nR[3 * r + 0] * t[0] + nR[3 * r + 1] * t[1] + nR[3 * r + 2] * t[2]);
		std::memcpy(t, nt, sizeof(nt));
		std::memcpy(R, nR, sizeof(nR));
	}
	Pose3 operator-() const { Pose3 r = *this; r.inverse(); return r; }
	Pose3 operator+(const Pose3 &B) const { Pose3 r; r.composeFrom(*this, B); return r; }
	Pose3 operator-(const Pose3 &B) const { Pose3 bi = -B; Pose3 r; r.composeFrom(bi, *this); return r; }
	void composePoint(double lx, double ly, double lz, double &gx, double &gy, double &gz) const {
		const double ox = R[0] * lx + R[1] * ly + R[2] * lz + t[0];
		const double oy = R[3] * lx + R[4] * ly + R[5] * lz + t[1];
		const double oz = R[6] * lx + R[7] * ly + R[8] * lz + t[2];
		gx = ox; gy = oy; gz = oz;
	}
	void inverseComposePoint(double gx, double gy, double gz, double &lx, double &ly, double &lz) const {
		const double dx = gx - t[0], dy = gy - t[1], dz = gz - t[2];
		const double ox = R[0] * dx + R[3] * dy + R[6] * dz;
		const double oy = R[1] * dx + R[4] * dy + R[7] * dz;
		const double oz = R[2] * dx + R[5] * dy + R[8] * dz;
		lx = ox; ly = oy; lz = oz;
	}
	void toFlat(double *o) const { std::memcpy(o, R, 9 * sizeof(double)); std::memcpy(o + 9, t, 3 * sizeof(double)); }
	void fromFlat(const double *o) { std::memcpy(R, o, 9 * sizeof(double)); std::memcpy(t, o + 9, 3 * sizeof(double)); }

	// SE_traits<3>::pseudo_exp == CPose3D::exp(v, pseudo_exponential=true):
	// translation copied verbatim, rotation = exp([w]x) (Rodrigues with MRPT's small-angle series).
	static Pose3 pseudo_exp(const double *v) {
		Pose3 P;
		P.t[0] = v[0]; P.t[1] = v[1]; P.t[2] = v[2];
		const double w0 = v[3], w1 = v[4], w2 = v[5];
		const double theta_sq = w0 * w0 + w1 * w1 + w2 * w2;
		const double theta = std::sqrt(theta_sq);
		double A, B;
		if (theta_sq < 1e-8) { A = 1.0 - theta_sq / 6.0; B = 0.5; }
		else if (theta_sq < 1e-6) { B = 0.5 - 0.25 * (1.0 / 6.0) * theta_sq; A = 1.0 - theta_sq * (1.0 / 6.0) * (1.0 - (1.0 / 20.0) * theta_sq); }
		else { const double inv_theta = 1.0 / theta; A = std::sin(theta) * inv_theta; B = (1 - std::cos(theta)) * (inv_theta * inv_theta); }
		const double wx2 = w0 * w0, wy2 = w1 * w1, wz2 = w2 * w2;
		P.R[0] = 1.0 - B * (wy2 + wz2);
		P.R[4] = 1.0 - B * (wx2 + wz2);
		P.R[8] = 1.0 - B * (wx2 + wy2);
		{ const double a = A * w2, b = B * (w0 * w1); P.R[1] = b - a; P.R[3] = b + a; }
		{ const double a = A * w1, b = B * (w0 * w2); P.R[2] = b + a; P.R[6] = b - a; }
		{ const double a = A * w0, b = B * (w1 * w2); P.R[5] = b - a; P.R[7] = b + a; }
		return P;
	}
};

// CPose3DQuat(x,y,z,q(r,x,y,z)) -> Pose3 (used for TStereoCamera::rightCameraPose and test ground truth).
inline Pose3 pose3_from_quat(double x, double y, double z, double qr, double qx, double qy, double qz) {
	Pose3 P;
	P.t[0] = x; P.t[1] = y; P.t[2] = z;
	P.R[0] = qr * qr + qx * qx - qy * qy - qz * qz; P.R[1] = 2 * (qx * qy - qr * qz);               P.R[2] = 2 * (qz * qx + qr * qy);
	P.R[3] = 2 * (qx * qy + qr * qz);               P.R[4] = qr * qr - qx * qx + qy * qy - qz * qz; P.R[5] = 2 * (qy * qz - qr * qx);
	P.R[6] = 2 * (qz * qx - qr * qy);               P.R[7] = 2 * (qy * qz + qr * qx);               P.R[8] = qr * qr - qx * qx - qy * qy + qz * qz;
	return P;
}

// Conversions used by determine_kf2kf_edges_to_create.h:198 "pose_t( CPose3D )".
inline Pose3 to_pose3(const Pose3 &p) { return p; }
inline Pose3 to_pose3(const Pose2 &p) { return Pose3(p.x, p.y, 0, p.phi, 0, 0); }
inline void from_pose3(const Pose3 &p, Pose3 &o) { o = p; }
inline void from_pose3(const Pose3 &p, Pose2 &o) { o = Pose2(p.t[0], p.t[1], std::atan2(p.R[3], p.R[0])); }

}  // namespace srba_oracle
