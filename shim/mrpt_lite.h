// mrpt_lite.h -- the handful of MRPT types that appear in the PUBLIC signatures of srba::RbaEngine<> (poses, camera calibration),
// for building the srba_b200 facade (shim/srba.h) where MRPT itself is not installed.  Same names, same accessors, same conventions
// as MRPT 1.3 (yaw-pitch-roll about Z-Y-X; CPose3D + CPose3D composes, unary minus inverts).  With real MRPT available define
// SRBA_B200_USE_REAL_MRPT and include <mrpt/poses.h> / <mrpt/utils/TCamera.h> before shim/srba.h instead of this file.
#pragma once
#include <cmath>
#include <cstdio>
#include <ostream>
#include <sstream>
#include <string>

#ifndef DEG2RAD
#define DEG2RAD(x) ((x) * M_PI / 180.0)
#endif
#ifndef RAD2DEG
#define RAD2DEG(x) ((x) * 180.0 / M_PI)
#endif

namespace mrpt {
namespace poses {

class CPose2D {
	double m_x, m_y, m_phi;

public:
	CPose2D(double x = 0, double y = 0, double phi = 0) : m_x(x), m_y(y), m_phi(phi) {}
	double x() const { return m_x; }
	double y() const { return m_y; }
	double phi() const { return m_phi; }
	void x(double v) { m_x = v; }
	void y(double v) { m_y = v; }
	void phi(double v) { m_phi = v; }
	CPose2D operator+(const CPose2D &b) const {  // this (+) b
		const double c = std::cos(m_phi), s = std::sin(m_phi);
		double ph = m_phi + b.m_phi;
		ph = std::fmod(ph + M_PI, 2 * M_PI);
		if (ph < 0) ph += 2 * M_PI;
		return CPose2D(m_x + c * b.m_x - s * b.m_y, m_y + s * b.m_x + c * b.m_y, ph - M_PI);
	}
	CPose2D operator-() const {
		const double c = std::cos(m_phi), s = std::sin(m_phi);
		return CPose2D(-m_x * c - m_y * s, m_x * s - m_y * c, -m_phi);
	}
	CPose2D operator-(const CPose2D &b) const { return (-b) + *this; }  // this (-) b
	std::string asString() const {
		char buf[128];
		std::snprintf(buf, sizeof(buf), "[%f %f %f]", m_x, m_y, RAD2DEG(m_phi));
		return buf;
	}
};
inline std::ostream &operator<<(std::ostream &o, const CPose2D &p) { return o << "(" << p.x() << "," << p.y() << "," << RAD2DEG(p.phi()) << "deg)"; }

class CPose3D {
	double m_R[9], m_t[3];  // rotation row-major, translation

public:
	CPose3D() { setFromValues(0, 0, 0, 0, 0, 0); }
	CPose3D(double x, double y, double z, double yaw = 0, double pitch = 0, double roll = 0) { setFromValues(x, y, z, yaw, pitch, roll); }
	void setFromValues(double x, double y, double z, double yaw, double pitch, double roll) {
		const double cy = std::cos(yaw), sy = std::sin(yaw), cp = std::cos(pitch), sp = std::sin(pitch), cr = std::cos(roll), sr = std::sin(roll);
		m_R[0] = cy * cp; m_R[1] = cy * sp * sr - sy * cr; m_R[2] = cy * sp * cr + sy * sr;
		m_R[3] = sy * cp; m_R[4] = sy * sp * sr + cy * cr; m_R[5] = sy * sp * cr - cy * sr;
		m_R[6] = -sp;     m_R[7] = cp * sr;                m_R[8] = cp * cr;
		m_t[0] = x; m_t[1] = y; m_t[2] = z;
	}
	static CPose3D fromRotationAndTranslation(const double R_row_major[9], const double t[3]) {
		CPose3D p;
		for (int i = 0; i < 9; i++) p.m_R[i] = R_row_major[i];
		for (int i = 0; i < 3; i++) p.m_t[i] = t[i];
		return p;
	}
	static CPose3D fromQuaternion(double x, double y, double z, double qr, double qx, double qy, double qz) {
		const double n = std::sqrt(qr * qr + qx * qx + qy * qy + qz * qz);
		qr /= n; qx /= n; qy /= n; qz /= n;
		const double R[9] = {qr * qr + qx * qx - qy * qy - qz * qz, 2 * (qx * qy - qr * qz), 2 * (qz * qx + qr * qy),
		                     2 * (qx * qy + qr * qz), qr * qr - qx * qx + qy * qy - qz * qz, 2 * (qy * qz - qr * qx),
		                     2 * (qz * qx - qr * qy), 2 * (qy * qz + qr * qx), qr * qr - qx * qx - qy * qy + qz * qz};
		const double t[3] = {x, y, z};
		return fromRotationAndTranslation(R, t);
	}
	double x() const { return m_t[0]; }
	double y() const { return m_t[1]; }
	double z() const { return m_t[2]; }
	const double *rotation() const { return m_R; }     // row-major 3x3 (MRPT: getRotationMatrix())
	const double *translation() const { return m_t; }
	void getYawPitchRoll(double &yaw, double &pitch, double &roll) const {
		pitch = std::atan2(-m_R[6], std::hypot(m_R[0], m_R[3]));
		if (std::fabs(std::fabs(pitch) - M_PI / 2) < 1e-12) { roll = 0; yaw = std::atan2(-m_R[1], m_R[4]); }
		else { roll = std::atan2(m_R[7], m_R[8]); yaw = std::atan2(m_R[3], m_R[0]); }
	}
	double yaw() const { double a, b, c; getYawPitchRoll(a, b, c); return a; }
	double pitch() const { double a, b, c; getYawPitchRoll(a, b, c); return b; }
	double roll() const { double a, b, c; getYawPitchRoll(a, b, c); return c; }
	CPose3D operator+(const CPose3D &b) const {
		CPose3D o;
		for (int i = 0; i < 3; i++) {
			for (int j = 0; j < 3; j++) o.m_R[3 * i + j] = m_R[3 * i] * b.m_R[j] + m_R[3 * i + 1] * b.m_R[3 + j] + m_R[3 * i + 2] * b.m_R[6 + j];
			o.m_t[i] = m_R[3 * i] * b.m_t[0] + m_R[3 * i + 1] * b.m_t[1] + m_R[3 * i + 2] * b.m_t[2] + m_t[i];
		}
		return o;
	}
	CPose3D operator-() const {
		CPose3D o;
		for (int i = 0; i < 3; i++)
			for (int j = 0; j < 3; j++) o.m_R[3 * i + j] = m_R[3 * j + i];
		for (int i = 0; i < 3; i++) o.m_t[i] = -(o.m_R[3 * i] * m_t[0] + o.m_R[3 * i + 1] * m_t[1] + o.m_R[3 * i + 2] * m_t[2]);
		return o;
	}
	CPose3D operator-(const CPose3D &b) const { return (-b) + *this; }
};
inline std::ostream &operator<<(std::ostream &o, const CPose3D &p) {
	double y, pi, r;
	p.getYawPitchRoll(y, pi, r);
	return o << "(x,y,z,yaw,pitch,roll)=(" << p.x() << "," << p.y() << "," << p.z() << "," << RAD2DEG(y) << "deg," << RAD2DEG(pi) << "deg," << RAD2DEG(r) << "deg)";
}

// x y z + quaternion, only what the stereo calibration needs (TStereoCamera::rightCameraPose)
class CPose3DQuat {
public:
	double v[7];
	CPose3DQuat() : v{0, 0, 0, 1, 0, 0, 0} {}
	CPose3DQuat(double x, double y, double z, double qr, double qx, double qy, double qz) : v{x, y, z, qr, qx, qy, qz} {}
	void fromString(const std::string &s) {  // "[x y z qr qx qy qz]"
		std::string t = s;
		for (char &c : t)
			if (c == '[' || c == ']' || c == ',') c = ' ';
		std::istringstream is(t);
		for (double &d : v) is >> d;
	}
};
}  // namespace poses

namespace utils {
class TCamera {
public:
	unsigned int ncols = 640, nrows = 480;
	double m_fx = 1, m_fy = 1, m_cx = 0, m_cy = 0;
	double fx() const { return m_fx; }
	double fy() const { return m_fy; }
	double cx() const { return m_cx; }
	double cy() const { return m_cy; }
	void fx(double v) { m_fx = v; }
	void fy(double v) { m_fy = v; }
	void cx(double v) { m_cx = v; }
	void cy(double v) { m_cy = v; }
	void setIntrinsicParamsFromValues(double fx, double fy, double cx, double cy) { m_fx = fx; m_fy = fy; m_cx = cx; m_cy = cy; }
};
class TStereoCamera {
public:
	TCamera leftCamera, rightCamera;
	mrpt::poses::CPose3DQuat rightCameraPose;
};
}  // namespace utils
}  // namespace mrpt
