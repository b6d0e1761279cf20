// host_math.hpp — the handful of vector/matrix operations the reference takes from GLM, restated with
// GLM 0.9.9's scalar (non-SIMD) operation order so that init-time tables are bit-identical
// (GLM itself is an unpinned external dependency of the reference).
#pragma once
#include <cmath>

namespace ssbh {

struct F3 { float x, y, z; };
inline F3 f3(float x, float y, float z) { return F3{ x, y, z }; }
inline F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline F3 operator*(F3 a, F3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline F3 operator/(F3 a, F3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline F3 operator*(F3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
inline float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline F3 cross(F3 x, F3 y) { return f3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
inline F3 normalize(F3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }

// column-major 3x3: m[c*3+r]
struct M3 { float m[9]; float& at(int c, int r) { return m[c * 3 + r]; } float at(int c, int r) const { return m[c * 3 + r]; } };
inline M3 m3_cols(F3 c0, F3 c1, F3 c2) { return M3{ { c0.x, c0.y, c0.z, c1.x, c1.y, c1.z, c2.x, c2.y, c2.z } }; }
inline M3 transpose(M3 const& a) { M3 r; for (int c = 0; c < 3; ++c) for (int rr = 0; rr < 3; ++rr) r.at(c, rr) = a.at(rr, c); return r; }
inline M3 inverse(M3 const& m) {
	float OneOverDeterminant = 1.0f / (+m.at(0, 0) * (m.at(1, 1) * m.at(2, 2) - m.at(2, 1) * m.at(1, 2))
	                                   - m.at(1, 0) * (m.at(0, 1) * m.at(2, 2) - m.at(2, 1) * m.at(0, 2))
	                                   + m.at(2, 0) * (m.at(0, 1) * m.at(1, 2) - m.at(1, 1) * m.at(0, 2)));
	M3 I;
	I.at(0, 0) = +(m.at(1, 1) * m.at(2, 2) - m.at(2, 1) * m.at(1, 2)) * OneOverDeterminant;
	I.at(1, 0) = -(m.at(1, 0) * m.at(2, 2) - m.at(2, 0) * m.at(1, 2)) * OneOverDeterminant;
	I.at(2, 0) = +(m.at(1, 0) * m.at(2, 1) - m.at(2, 0) * m.at(1, 1)) * OneOverDeterminant;
	I.at(0, 1) = -(m.at(0, 1) * m.at(2, 2) - m.at(2, 1) * m.at(0, 2)) * OneOverDeterminant;
	I.at(1, 1) = +(m.at(0, 0) * m.at(2, 2) - m.at(2, 0) * m.at(0, 2)) * OneOverDeterminant;
	I.at(2, 1) = -(m.at(0, 0) * m.at(2, 1) - m.at(2, 0) * m.at(0, 1)) * OneOverDeterminant;
	I.at(0, 2) = +(m.at(0, 1) * m.at(1, 2) - m.at(1, 1) * m.at(0, 2)) * OneOverDeterminant;
	I.at(1, 2) = -(m.at(0, 0) * m.at(1, 2) - m.at(1, 0) * m.at(0, 2)) * OneOverDeterminant;
	I.at(2, 2) = +(m.at(0, 0) * m.at(1, 1) - m.at(1, 0) * m.at(0, 1)) * OneOverDeterminant;
	return I;
}
inline F3 operator*(M3 const& m, F3 v) {
	return f3(m.at(0, 0) * v.x + m.at(1, 0) * v.y + m.at(2, 0) * v.z,
	          m.at(0, 1) * v.x + m.at(1, 1) * v.y + m.at(2, 1) * v.z,
	          m.at(0, 2) * v.x + m.at(1, 2) * v.y + m.at(2, 2) * v.z);
}

// column-major 4x4: m[c*4+r]
template <typename T> struct M4 { T m[16]; T& at(int c, int r) { return m[c * 4 + r]; } T at(int c, int r) const { return m[c * 4 + r]; } };
template <typename T> inline M4<T> m4_diag(T s) { M4<T> r; for (int i = 0; i < 16; ++i) r.m[i] = T(0); for (int i = 0; i < 4; ++i) r.at(i, i) = s; return r; }
inline M4<float> perspectiveFov(float fov, float width, float height, float zNear, float zFar) {  // RH, depth -1..1
	float const rad = fov;
	float const h = std::cos(0.5f * rad) / std::sin(0.5f * rad);
	float const w = h * height / width;
	M4<float> R = m4_diag(0.0f);
	R.at(0, 0) = w; R.at(1, 1) = h;
	R.at(2, 2) = -(zFar + zNear) / (zFar - zNear);
	R.at(2, 3) = -1.0f;
	R.at(3, 2) = -(2.0f * zFar * zNear) / (zFar - zNear);
	return R;
}
inline M4<float> lookAt(F3 eye, F3 center, F3 up) {  // RH
	F3 const f = normalize(center - eye);
	F3 const s = normalize(cross(f, up));
	F3 const u = cross(s, f);
	M4<float> R = m4_diag(1.0f);
	R.at(0, 0) = s.x; R.at(1, 0) = s.y; R.at(2, 0) = s.z;
	R.at(0, 1) = u.x; R.at(1, 1) = u.y; R.at(2, 1) = u.z;
	R.at(0, 2) = -f.x; R.at(1, 2) = -f.y; R.at(2, 2) = -f.z;
	R.at(3, 0) = -dot(s, eye); R.at(3, 1) = -dot(u, eye); R.at(3, 2) = dot(f, eye);
	return R;
}
inline M4<double> widen(M4<float> const& a) { M4<double> r; for (int i = 0; i < 16; ++i) r.m[i] = (double)a.m[i]; return r; }
inline M4<double> mul(M4<double> const& a, M4<double> const& b) {
	M4<double> R;
	for (int c = 0; c < 4; ++c)
		for (int r = 0; r < 4; ++r)
			R.at(c, r) = a.at(0, r) * b.at(c, 0) + a.at(1, r) * b.at(c, 1) + a.at(2, r) * b.at(c, 2) + a.at(3, r) * b.at(c, 3);
	return R;
}
inline M4<double> inverse(M4<double> const& m) {
	typedef double T;
	T Coef00 = m.at(2, 2) * m.at(3, 3) - m.at(3, 2) * m.at(2, 3);
	T Coef02 = m.at(1, 2) * m.at(3, 3) - m.at(3, 2) * m.at(1, 3);
	T Coef03 = m.at(1, 2) * m.at(2, 3) - m.at(2, 2) * m.at(1, 3);
	T Coef04 = m.at(2, 1) * m.at(3, 3) - m.at(3, 1) * m.at(2, 3);
	T Coef06 = m.at(1, 1) * m.at(3, 3) - m.at(3, 1) * m.at(1, 3);
	T Coef07 = m.at(1, 1) * m.at(2, 3) - m.at(2, 1) * m.at(1, 3);
	T Coef08 = m.at(2, 1) * m.at(3, 2) - m.at(3, 1) * m.at(2, 2);
	T Coef10 = m.at(1, 1) * m.at(3, 2) - m.at(3, 1) * m.at(1, 2);
	T Coef11 = m.at(1, 1) * m.at(2, 2) - m.at(2, 1) * m.at(1, 2);
	T Coef12 = m.at(2, 0) * m.at(3, 3) - m.at(3, 0) * m.at(2, 3);
	T Coef14 = m.at(1, 0) * m.at(3, 3) - m.at(3, 0) * m.at(1, 3);
	T Coef15 = m.at(1, 0) * m.at(2, 3) - m.at(2, 0) * m.at(1, 3);
	T Coef16 = m.at(2, 0) * m.at(3, 2) - m.at(3, 0) * m.at(2, 2);
	T Coef18 = m.at(1, 0) * m.at(3, 2) - m.at(3, 0) * m.at(1, 2);
	T Coef19 = m.at(1, 0) * m.at(2, 2) - m.at(2, 0) * m.at(1, 2);
	T Coef20 = m.at(2, 0) * m.at(3, 1) - m.at(3, 0) * m.at(2, 1);
	T Coef22 = m.at(1, 0) * m.at(3, 1) - m.at(3, 0) * m.at(1, 1);
	T Coef23 = m.at(1, 0) * m.at(2, 1) - m.at(2, 0) * m.at(1, 1);
	T Fac0[4] = { Coef00, Coef00, Coef02, Coef03 }, Fac1[4] = { Coef04, Coef04, Coef06, Coef07 };
	T Fac2[4] = { Coef08, Coef08, Coef10, Coef11 }, Fac3[4] = { Coef12, Coef12, Coef14, Coef15 };
	T Fac4[4] = { Coef16, Coef16, Coef18, Coef19 }, Fac5[4] = { Coef20, Coef20, Coef22, Coef23 };
	T Vec0[4] = { m.at(1, 0), m.at(0, 0), m.at(0, 0), m.at(0, 0) }, Vec1[4] = { m.at(1, 1), m.at(0, 1), m.at(0, 1), m.at(0, 1) };
	T Vec2[4] = { m.at(1, 2), m.at(0, 2), m.at(0, 2), m.at(0, 2) }, Vec3[4] = { m.at(1, 3), m.at(0, 3), m.at(0, 3), m.at(0, 3) };
	T SignA[4] = { +1, -1, +1, -1 }, SignB[4] = { -1, +1, -1, +1 };
	M4<double> Inv;
	for (int i = 0; i < 4; ++i) {
		T Inv0 = Vec1[i] * Fac0[i] - Vec2[i] * Fac1[i] + Vec3[i] * Fac2[i];
		T Inv1 = Vec0[i] * Fac0[i] - Vec2[i] * Fac3[i] + Vec3[i] * Fac4[i];
		T Inv2 = Vec0[i] * Fac1[i] - Vec1[i] * Fac3[i] + Vec3[i] * Fac5[i];
		T Inv3 = Vec0[i] * Fac2[i] - Vec1[i] * Fac4[i] + Vec2[i] * Fac5[i];
		Inv.at(0, i) = Inv0 * SignA[i]; Inv.at(1, i) = Inv1 * SignB[i];
		Inv.at(2, i) = Inv2 * SignA[i]; Inv.at(3, i) = Inv3 * SignB[i];
	}
	T Dot0[4] = { m.at(0, 0) * Inv.at(0, 0), m.at(0, 1) * Inv.at(1, 0), m.at(0, 2) * Inv.at(2, 0), m.at(0, 3) * Inv.at(3, 0) };
	T Dot1 = (Dot0[0] + Dot0[1]) + (Dot0[2] + Dot0[3]);
	T OneOverDeterminant = 1.0 / Dot1;
	for (int i = 0; i < 16; ++i) Inv.m[i] = Inv.m[i] * OneOverDeterminant;
	return Inv;
}

}  // namespace ssbh
