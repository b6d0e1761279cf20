// TEST INFRASTRUCTURE ONLY — not part of the product.
//
// Minimal GLM-compatible shim used ONLY to compile the unmodified reference
// (/root/reference/src) into oracle/_ref/.  GLM itself is an external, unpinned,
// un-vendored dependency of the reference (cmake/FindGLM.cmake:3-24, README.md:16-19)
// and is not installed in this image.  This file restates the *published scalar
// (non-SIMD) semantics of GLM 0.9.9* for exactly the surface the reference touches
// (list: SURVEY.md §8c).  Operation order is part of the contract:
//   dot(vec3)  = (x*x' + y*y') + z*z'          dot(vec4) = (x*x'+y*y') + (z*z'+w*w')
//   cross(x,y) = (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
//   normalize  = v * inversesqrt(dot(v,v)),  inversesqrt(x) = 1/sqrt(x)
//   mat3*vec3  = m[0][r]*v.x + m[1][r]*v.y + m[2][r]*v.z   (left to right)
//   mat4*vec4  = (m[0]*v.x + m[1]*v.y) + (m[2]*v.z + m[3]*v.w)
// Everything is component-wise, evaluated left-to-right, nothing fused
// (build with -ffp-contract=off).
#pragma once

#include <cmath>
#include <cstddef>
#include <type_traits>

namespace glm {

enum qualifier { packed_highp, defaultp = packed_highp };
typedef std::size_t length_t;  // GLM_FORCE_SIZE_T_LENGTH (stdafx.hpp:26)

template <length_t L, typename T, qualifier Q = defaultp> struct vec;

// ---------------------------------------------------------------- vec2
template <typename T, qualifier Q> struct vec<2, T, Q> {
	typedef T value_type;
	union { T x, r, s; };
	union { T y, g, t; };

	vec() : x(0), y(0) {}
	vec(vec const&) = default;
	vec& operator=(vec const&) = default;
	template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>>
	explicit vec(U sc) : x(static_cast<T>(sc)), y(static_cast<T>(sc)) {}
	template <typename A, typename B,
		typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
	vec(A a, B b) : x(static_cast<T>(a)), y(static_cast<T>(b)) {}
	template <typename U, qualifier P> vec(vec<2, U, P> const& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
	template <typename U, qualifier P> explicit vec(vec<3, U, P> const& v);
	template <typename U, qualifier P> explicit vec(vec<4, U, P> const& v);

	static constexpr length_t length() { return 2; }
	T& operator[](length_t i) { return i == 0 ? x : y; }
	T const& operator[](length_t i) const { return i == 0 ? x : y; }

	template <typename U> vec& operator+=(vec<2, U, Q> const& v) { x += static_cast<T>(v.x); y += static_cast<T>(v.y); return *this; }
	template <typename U> vec& operator-=(vec<2, U, Q> const& v) { x -= static_cast<T>(v.x); y -= static_cast<T>(v.y); return *this; }
	template <typename U> vec& operator*=(vec<2, U, Q> const& v) { x *= static_cast<T>(v.x); y *= static_cast<T>(v.y); return *this; }
	template <typename U> vec& operator/=(vec<2, U, Q> const& v) { x /= static_cast<T>(v.x); y /= static_cast<T>(v.y); return *this; }
	template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>> vec& operator+=(U sc) { T c = static_cast<T>(sc); x += c; y += c; return *this; }
	template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>> vec& operator-=(U sc) { T c = static_cast<T>(sc); x -= c; y -= c; return *this; }
	template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>> vec& operator*=(U sc) { T c = static_cast<T>(sc); x *= c; y *= c; return *this; }
	template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>> vec& operator/=(U sc) { T c = static_cast<T>(sc); x /= c; y /= c; return *this; }
};

// ---------------------------------------------------------------- vec3
template <typename T, qualifier Q> struct vec<3, T, Q> {
	typedef T value_type;
	union { T x, r, s; };
	union { T y, g, t; };
	union { T z, b, p; };

	vec() : x(0), y(0), z(0) {}
	vec(vec const&) = default;
	vec& operator=(vec const&) = default;
	template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>>
	explicit vec(U sc) : x(static_cast<T>(sc)), y(static_cast<T>(sc)), z(static_cast<T>(sc)) {}
	template <typename A, typename B, typename C,
		typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value && std::is_arithmetic<C>::value>>
	vec(A a, B b_, C c) : x(static_cast<T>(a)), y(static_cast<T>(b_)), z(static_cast<T>(c)) {}
	template <typename U, qualifier P> vec(vec<3, U, P> const& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}
	template <typename U, qualifier P> explicit vec(vec<4, U, P> const& v);
	template <typename U, qualifier P, typename C, typename = std::enable_if_t<std::is_arithmetic<C>::value>>
	vec(vec<2, U, P> const& v, C c) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(c)) {}

	static constexpr length_t length() { return 3; }
	T& operator[](length_t i) { return i == 0 ? x : (i == 1 ? y : z); }
	T const& operator[](length_t i) const { return i == 0 ? x : (i == 1 ? y : z); }

	template <typename U> vec& operator+=(vec<3, U, Q> const& v) { x += static_cast<T>(v.x); y += static_cast<T>(v.y); z += static_cast<T>(v.z); return *this; }
	template <typename U> vec& operator-=(vec<3, U, Q> const& v) { x -= static_cast<T>(v.x); y -= static_cast<T>(v.y); z -= static_cast<T>(v.z); return *this; }
	template <typename U> vec& operator*=(vec<3, U, Q> const& v) { x *= static_cast<T>(v.x); y *= static_cast<T>(v.y); z *= static_cast<T>(v.z); return *this; }
	template <typename U> vec& operator/=(vec<3, U, Q> const& v) { x /= static_cast<T>(v.x); y /= static_cast<T>(v.y); z /= static_cast<T>(v.z); return *this; }
	template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>> vec& operator+=(U sc) { T c = static_cast<T>(sc); x += c; y += c; z += c; return *this; }
	template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>> vec& operator-=(U sc) { T c = static_cast<T>(sc); x -= c; y -= c; z -= c; return *this; }
	template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>> vec& operator*=(U sc) { T c = static_cast<T>(sc); x *= c; y *= c; z *= c; return *this; }
	template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>> vec& operator/=(U sc) { T c = static_cast<T>(sc); x /= c; y /= c; z /= c; return *this; }
};

// ---------------------------------------------------------------- vec4
template <typename T, qualifier Q> struct vec<4, T, Q> {
	typedef T value_type;
	union { T x, r, s; };
	union { T y, g, t; };
	union { T z, b, p; };
	union { T w, a, q; };

	vec() : x(0), y(0), z(0), w(0) {}
	vec(vec const&) = default;
	vec& operator=(vec const&) = default;
	template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>>
	explicit vec(U sc) : x(static_cast<T>(sc)), y(static_cast<T>(sc)), z(static_cast<T>(sc)), w(static_cast<T>(sc)) {}
	template <typename A, typename B, typename C, typename D,
		typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value && std::is_arithmetic<C>::value && std::is_arithmetic<D>::value>>
	vec(A a_, B b_, C c, D d) : x(static_cast<T>(a_)), y(static_cast<T>(b_)), z(static_cast<T>(c)), w(static_cast<T>(d)) {}
	template <typename U, qualifier P> vec(vec<4, U, P> const& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)), w(static_cast<T>(v.w)) {}
	template <typename U, qualifier P, typename D, typename = std::enable_if_t<std::is_arithmetic<D>::value>>
	vec(vec<3, U, P> const& v, D d) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)), w(static_cast<T>(d)) {}
	template <typename U, qualifier P, typename C, typename D,
		typename = std::enable_if_t<std::is_arithmetic<C>::value && std::is_arithmetic<D>::value>>
	vec(vec<2, U, P> const& v, C c, D d) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(c)), w(static_cast<T>(d)) {}

	static constexpr length_t length() { return 4; }
	T& operator[](length_t i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
	T const& operator[](length_t i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }

	template <typename U> vec& operator+=(vec<4, U, Q> const& v) { x += static_cast<T>(v.x); y += static_cast<T>(v.y); z += static_cast<T>(v.z); w += static_cast<T>(v.w); return *this; }
	template <typename U> vec& operator-=(vec<4, U, Q> const& v) { x -= static_cast<T>(v.x); y -= static_cast<T>(v.y); z -= static_cast<T>(v.z); w -= static_cast<T>(v.w); return *this; }
	template <typename U> vec& operator*=(vec<4, U, Q> const& v) { x *= static_cast<T>(v.x); y *= static_cast<T>(v.y); z *= static_cast<T>(v.z); w *= static_cast<T>(v.w); return *this; }
	template <typename U> vec& operator/=(vec<4, U, Q> const& v) { x /= static_cast<T>(v.x); y /= static_cast<T>(v.y); z /= static_cast<T>(v.z); w /= static_cast<T>(v.w); return *this; }
	template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>> vec& operator+=(U sc) { T c = static_cast<T>(sc); x += c; y += c; z += c; w += c; return *this; }
	template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>> vec& operator-=(U sc) { T c = static_cast<T>(sc); x -= c; y -= c; z -= c; w -= c; return *this; }
	template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>> vec& operator*=(U sc) { T c = static_cast<T>(sc); x *= c; y *= c; z *= c; w *= c; return *this; }
	template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value>> vec& operator/=(U sc) { T c = static_cast<T>(sc); x /= c; y /= c; z /= c; w /= c; return *this; }
};

// truncating conversions
template <typename T, qualifier Q> template <typename U, qualifier P>
vec<2, T, Q>::vec(vec<3, U, P> const& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
template <typename T, qualifier Q> template <typename U, qualifier P>
vec<2, T, Q>::vec(vec<4, U, P> const& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
template <typename T, qualifier Q> template <typename U, qualifier P>
vec<3, T, Q>::vec(vec<4, U, P> const& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}

// ---------------------------------------------------------------- vec operators (binary)
#define SSB_GLM_VEC_BINOP(OP)                                                                                   \
	template <typename T, qualifier Q> inline vec<2, T, Q> operator OP(vec<2, T, Q> const& a, vec<2, T, Q> const& b) { return vec<2, T, Q>(a.x OP b.x, a.y OP b.y); } \
	template <typename T, qualifier Q> inline vec<3, T, Q> operator OP(vec<3, T, Q> const& a, vec<3, T, Q> const& b) { return vec<3, T, Q>(a.x OP b.x, a.y OP b.y, a.z OP b.z); } \
	template <typename T, qualifier Q> inline vec<4, T, Q> operator OP(vec<4, T, Q> const& a, vec<4, T, Q> const& b) { return vec<4, T, Q>(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); } \
	template <typename T, qualifier Q> inline vec<2, T, Q> operator OP(vec<2, T, Q> const& a, T b) { return vec<2, T, Q>(a.x OP b, a.y OP b); } \
	template <typename T, qualifier Q> inline vec<3, T, Q> operator OP(vec<3, T, Q> const& a, T b) { return vec<3, T, Q>(a.x OP b, a.y OP b, a.z OP b); } \
	template <typename T, qualifier Q> inline vec<4, T, Q> operator OP(vec<4, T, Q> const& a, T b) { return vec<4, T, Q>(a.x OP b, a.y OP b, a.z OP b, a.w OP b); } \
	template <typename T, qualifier Q> inline vec<2, T, Q> operator OP(T a, vec<2, T, Q> const& b) { return vec<2, T, Q>(a OP b.x, a OP b.y); } \
	template <typename T, qualifier Q> inline vec<3, T, Q> operator OP(T a, vec<3, T, Q> const& b) { return vec<3, T, Q>(a OP b.x, a OP b.y, a OP b.z); } \
	template <typename T, qualifier Q> inline vec<4, T, Q> operator OP(T a, vec<4, T, Q> const& b) { return vec<4, T, Q>(a OP b.x, a OP b.y, a OP b.z, a OP b.w); }
SSB_GLM_VEC_BINOP(+)
SSB_GLM_VEC_BINOP(-)
SSB_GLM_VEC_BINOP(*)
SSB_GLM_VEC_BINOP(/)
#undef SSB_GLM_VEC_BINOP

template <typename T, qualifier Q> inline vec<2, T, Q> operator-(vec<2, T, Q> const& v) { return vec<2, T, Q>(-v.x, -v.y); }
template <typename T, qualifier Q> inline vec<3, T, Q> operator-(vec<3, T, Q> const& v) { return vec<3, T, Q>(-v.x, -v.y, -v.z); }
template <typename T, qualifier Q> inline vec<4, T, Q> operator-(vec<4, T, Q> const& v) { return vec<4, T, Q>(-v.x, -v.y, -v.z, -v.w); }

typedef vec<2, float, defaultp> vec2;
typedef vec<3, float, defaultp> vec3;
typedef vec<4, float, defaultp> vec4;
typedef vec<2, double, defaultp> dvec2;
typedef vec<3, double, defaultp> dvec3;
typedef vec<4, double, defaultp> dvec4;

// ---------------------------------------------------------------- scalar / vector functions
template <typename T> inline T min(T a, T b) { return (b < a) ? b : a; }
template <typename T> inline T max(T a, T b) { return (a < b) ? b : a; }
template <typename T, typename = std::enable_if_t<std::is_arithmetic<T>::value>>
inline T clamp(T x, T lo, T hi) { return glm::min(glm::max(x, lo), hi); }
template <length_t L, typename T, qualifier Q>
inline vec<L, T, Q> clamp(vec<L, T, Q> const& v, vec<L, T, Q> const& lo, vec<L, T, Q> const& hi) {
	vec<L, T, Q> out;
	for (length_t i = 0; i < L; ++i) out[i] = glm::min(glm::max(v[i], lo[i]), hi[i]);
	return out;
}
template <length_t L, typename T, qualifier Q> inline vec<L, T, Q> abs(vec<L, T, Q> const& v) {
	vec<L, T, Q> out;
	for (length_t i = 0; i < L; ++i) out[i] = std::abs(v[i]);
	return out;
}
template <length_t L, typename T, qualifier Q> inline vec<L, T, Q> round(vec<L, T, Q> const& v) {
	vec<L, T, Q> out;
	for (length_t i = 0; i < L; ++i) out[i] = std::round(v[i]);
	return out;
}

template <typename T, qualifier Q> inline T dot(vec<2, T, Q> const& a, vec<2, T, Q> const& b) {
	vec<2, T, Q> t(a * b);
	return t.x + t.y;
}
template <typename T, qualifier Q> inline T dot(vec<3, T, Q> const& a, vec<3, T, Q> const& b) {
	vec<3, T, Q> t(a * b);
	return t.x + t.y + t.z;
}
template <typename T, qualifier Q> inline T dot(vec<4, T, Q> const& a, vec<4, T, Q> const& b) {
	vec<4, T, Q> t(a * b);
	return (t.x + t.y) + (t.z + t.w);
}
template <typename T, qualifier Q> inline vec<3, T, Q> cross(vec<3, T, Q> const& x, vec<3, T, Q> const& y) {
	return vec<3, T, Q>(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
template <typename T, typename = std::enable_if_t<std::is_floating_point<T>::value>>
inline T inversesqrt(T x) { return static_cast<T>(1) / std::sqrt(x); }
template <length_t L, typename T, qualifier Q> inline T length(vec<L, T, Q> const& v) { return std::sqrt(dot(v, v)); }
template <length_t L, typename T, qualifier Q> inline vec<L, T, Q> normalize(vec<L, T, Q> const& v) {
	return v * inversesqrt(dot(v, v));
}
template <typename T, typename = std::enable_if_t<std::is_floating_point<T>::value>>
inline T radians(T deg) { return deg * static_cast<T>(0.01745329251994329576923690768489); }
template <typename T, typename = std::enable_if_t<std::is_floating_point<T>::value>>
inline T degrees(T rad) { return rad * static_cast<T>(57.295779513082320876798154814105); }

// ---------------------------------------------------------------- mat3x3 (column-major)
template <length_t C, length_t R, typename T, qualifier Q = defaultp> struct mat;

template <typename T, qualifier Q> struct mat<3, 3, T, Q> {
	typedef vec<3, T, Q> col_type;
	col_type value[3];

	mat() {}
	mat(col_type const& c0, col_type const& c1, col_type const& c2) { value[0] = c0; value[1] = c1; value[2] = c2; }
	mat(T x0, T y0, T z0, T x1, T y1, T z1, T x2, T y2, T z2) {
		value[0] = col_type(x0, y0, z0); value[1] = col_type(x1, y1, z1); value[2] = col_type(x2, y2, z2);
	}
	col_type& operator[](length_t i) { return value[i]; }
	col_type const& operator[](length_t i) const { return value[i]; }
};
typedef mat<3, 3, float, defaultp> mat3x3;
typedef mat<3, 3, float, defaultp> mat3;

template <typename T, qualifier Q> inline mat<3, 3, T, Q> transpose(mat<3, 3, T, Q> const& m) {
	mat<3, 3, T, Q> r;
	for (length_t c = 0; c < 3; ++c)
		for (length_t rr = 0; rr < 3; ++rr) r[c][rr] = m[rr][c];
	return r;
}
template <typename T, qualifier Q> inline mat<3, 3, T, Q> inverse(mat<3, 3, T, Q> const& m) {
	T OneOverDeterminant = static_cast<T>(1) / (
		+ m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2])
		- m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2])
		+ m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]));
	mat<3, 3, T, Q> Inverse;
	Inverse[0][0] = + (m[1][1] * m[2][2] - m[2][1] * m[1][2]) * OneOverDeterminant;
	Inverse[1][0] = - (m[1][0] * m[2][2] - m[2][0] * m[1][2]) * OneOverDeterminant;
	Inverse[2][0] = + (m[1][0] * m[2][1] - m[2][0] * m[1][1]) * OneOverDeterminant;
	Inverse[0][1] = - (m[0][1] * m[2][2] - m[2][1] * m[0][2]) * OneOverDeterminant;
	Inverse[1][1] = + (m[0][0] * m[2][2] - m[2][0] * m[0][2]) * OneOverDeterminant;
	Inverse[2][1] = - (m[0][0] * m[2][1] - m[2][0] * m[0][1]) * OneOverDeterminant;
	Inverse[0][2] = + (m[0][1] * m[1][2] - m[1][1] * m[0][2]) * OneOverDeterminant;
	Inverse[1][2] = - (m[0][0] * m[1][2] - m[1][0] * m[0][2]) * OneOverDeterminant;
	Inverse[2][2] = + (m[0][0] * m[1][1] - m[1][0] * m[0][1]) * OneOverDeterminant;
	return Inverse;
}
template <typename T, qualifier Q> inline vec<3, T, Q> operator*(mat<3, 3, T, Q> const& m, vec<3, T, Q> const& v) {
	return vec<3, T, Q>(
		m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z,
		m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
		m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
}
template <typename T, qualifier Q> inline mat<3, 3, T, Q> operator*(mat<3, 3, T, Q> const& m, T s) {
	return mat<3, 3, T, Q>(m[0] * s, m[1] * s, m[2] * s);
}

// ---------------------------------------------------------------- mat4x4 (column-major)
template <typename T, qualifier Q> struct mat<4, 4, T, Q> {
	typedef vec<4, T, Q> col_type;
	col_type value[4];

	mat() {}
	explicit mat(T s) {
		value[0] = col_type(s, 0, 0, 0); value[1] = col_type(0, s, 0, 0);
		value[2] = col_type(0, 0, s, 0); value[3] = col_type(0, 0, 0, s);
	}
	mat(col_type const& c0, col_type const& c1, col_type const& c2, col_type const& c3) {
		value[0] = c0; value[1] = c1; value[2] = c2; value[3] = c3;
	}
	template <typename U, qualifier P> mat(mat<4, 4, U, P> const& m) {
		for (length_t i = 0; i < 4; ++i) value[i] = col_type(m[i]);
	}
	col_type& operator[](length_t i) { return value[i]; }
	col_type const& operator[](length_t i) const { return value[i]; }
};
typedef mat<4, 4, float, defaultp> mat4x4;
typedef mat<4, 4, float, defaultp> mat4;
typedef mat<4, 4, double, defaultp> dmat4x4;
typedef mat<4, 4, double, defaultp> dmat4;

template <typename T, qualifier Q> inline vec<4, T, Q> operator*(mat<4, 4, T, Q> const& m, vec<4, T, Q> const& v) {
	typedef vec<4, T, Q> col;
	col const Mov0(v[0]); col const Mov1(v[1]);
	col const Mul0 = m[0] * Mov0; col const Mul1 = m[1] * Mov1;
	col const Add0 = Mul0 + Mul1;
	col const Mov2(v[2]); col const Mov3(v[3]);
	col const Mul2 = m[2] * Mov2; col const Mul3 = m[3] * Mov3;
	col const Add1 = Mul2 + Mul3;
	col const Add2 = Add0 + Add1;
	return Add2;
}
template <typename T, qualifier Q> inline mat<4, 4, T, Q> operator*(mat<4, 4, T, Q> const& m1, mat<4, 4, T, Q> const& m2) {
	typedef vec<4, T, Q> col;
	col const SrcA0 = m1[0], SrcA1 = m1[1], SrcA2 = m1[2], SrcA3 = m1[3];
	col const SrcB0 = m2[0], SrcB1 = m2[1], SrcB2 = m2[2], SrcB3 = m2[3];
	mat<4, 4, T, Q> Result;
	Result[0] = SrcA0 * SrcB0[0] + SrcA1 * SrcB0[1] + SrcA2 * SrcB0[2] + SrcA3 * SrcB0[3];
	Result[1] = SrcA0 * SrcB1[0] + SrcA1 * SrcB1[1] + SrcA2 * SrcB1[2] + SrcA3 * SrcB1[3];
	Result[2] = SrcA0 * SrcB2[0] + SrcA1 * SrcB2[1] + SrcA2 * SrcB2[2] + SrcA3 * SrcB2[3];
	Result[3] = SrcA0 * SrcB3[0] + SrcA1 * SrcB3[1] + SrcA2 * SrcB3[2] + SrcA3 * SrcB3[3];
	return Result;
}
template <typename T, qualifier Q> inline mat<4, 4, T, Q> operator*(mat<4, 4, T, Q> const& m, T s) {
	return mat<4, 4, T, Q>(m[0] * s, m[1] * s, m[2] * s, m[3] * s);
}
template <typename T, qualifier Q> inline mat<4, 4, T, Q> inverse(mat<4, 4, T, Q> const& m) {
	typedef vec<4, T, Q> v4;
	T Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
	T Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
	T Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
	T Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
	T Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
	T Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
	T Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
	T Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
	T Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
	T Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
	T Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
	T Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
	T Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
	T Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
	T Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
	T Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
	T Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
	T Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
	v4 Fac0(Coef00, Coef00, Coef02, Coef03);
	v4 Fac1(Coef04, Coef04, Coef06, Coef07);
	v4 Fac2(Coef08, Coef08, Coef10, Coef11);
	v4 Fac3(Coef12, Coef12, Coef14, Coef15);
	v4 Fac4(Coef16, Coef16, Coef18, Coef19);
	v4 Fac5(Coef20, Coef20, Coef22, Coef23);
	v4 Vec0(m[1][0], m[0][0], m[0][0], m[0][0]);
	v4 Vec1(m[1][1], m[0][1], m[0][1], m[0][1]);
	v4 Vec2(m[1][2], m[0][2], m[0][2], m[0][2]);
	v4 Vec3(m[1][3], m[0][3], m[0][3], m[0][3]);
	v4 Inv0(Vec1 * Fac0 - Vec2 * Fac1 + Vec3 * Fac2);
	v4 Inv1(Vec0 * Fac0 - Vec2 * Fac3 + Vec3 * Fac4);
	v4 Inv2(Vec0 * Fac1 - Vec1 * Fac3 + Vec3 * Fac5);
	v4 Inv3(Vec0 * Fac2 - Vec1 * Fac4 + Vec2 * Fac5);
	v4 SignA(+1, -1, +1, -1);
	v4 SignB(-1, +1, -1, +1);
	mat<4, 4, T, Q> Inverse(Inv0 * SignA, Inv1 * SignB, Inv2 * SignA, Inv3 * SignB);
	v4 Row0(Inverse[0][0], Inverse[1][0], Inverse[2][0], Inverse[3][0]);
	v4 Dot0(m[0] * Row0);
	T Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
	T OneOverDeterminant = static_cast<T>(1) / Dot1;
	return Inverse * OneOverDeterminant;
}

// ---------------------------------------------------------------- gtc/matrix_transform (RH, depth -1..1)
template <typename T> inline mat<4, 4, T, defaultp> perspectiveFov(T fov, T width, T height, T zNear, T zFar) {
	T const rad = fov;
	T const h = std::cos(static_cast<T>(0.5) * rad) / std::sin(static_cast<T>(0.5) * rad);
	T const w = h * height / width;
	mat<4, 4, T, defaultp> Result(static_cast<T>(0));
	Result[0][0] = w;
	Result[1][1] = h;
	Result[2][2] = -(zFar + zNear) / (zFar - zNear);
	Result[2][3] = -static_cast<T>(1);
	Result[3][2] = -(static_cast<T>(2) * zFar * zNear) / (zFar - zNear);
	return Result;
}
template <typename T, qualifier Q>
inline mat<4, 4, T, Q> lookAt(vec<3, T, Q> const& eye, vec<3, T, Q> const& center, vec<3, T, Q> const& up) {
	vec<3, T, Q> const f(normalize(center - eye));
	vec<3, T, Q> const s(normalize(cross(f, up)));
	vec<3, T, Q> const u(cross(s, f));
	mat<4, 4, T, Q> Result(static_cast<T>(1));
	Result[0][0] = s.x; Result[1][0] = s.y; Result[2][0] = s.z;
	Result[0][1] = u.x; Result[1][1] = u.y; Result[2][1] = u.z;
	Result[0][2] = -f.x; Result[1][2] = -f.y; Result[2][2] = -f.z;
	Result[3][0] = -dot(s, eye);
	Result[3][1] = -dot(u, eye);
	Result[3][2] = dot(f, eye);
	return Result;
}

}  // namespace glm
