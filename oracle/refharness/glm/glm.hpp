// Minimal GLM stand-in used ONLY to compile the reference headers for the oracle
// (oracle/_ref/libsmref.so).  TEST INFRASTRUCTURE - never linked into the product.
//
// GLM itself is not vendored by the reference (Makefile:8-9,40 only adds an include
// path) and is not installed in this image, so the arithmetic contract of the calls
// the hot path makes is restated here.  Semantics follow GLM 0.9.9.x (the version
// shipped by the distribution the reference README names, README.md:21):
//   dot        : component products summed left to right
//   length     : sqrt(dot(v,v))
//   normalize  : v * inversesqrt(dot(v,v)),  inversesqrt(x) = 1/sqrt(x)
//   cross      : (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y)
//   mix(x,y,a) : T( U(x)*(U(1)-a) + U(y)*a )   evaluated in the weight's type U
//   round      : std::round (half away from zero);  floor: std::floor; fract: x-floor(x)
//   vec/scalar : per-component division (no reciprocal multiply)
//   conversions between vector types: per-component static_cast (truncation)
// Call sites that depend on this: layermap.h:344-375,382-383,430-431; water.h:14,45,
// 56-66,252; wind.h:17,60,76-87,99,107; particle.h:26.
// No reference test pins results at this boundary ("parity unpinned" at the GLM
// level): this header IS the definition the CUDA code mirrors.
#pragma once
#include <cmath>
#include <cstddef>

namespace glm {

typedef unsigned int uint;

template <typename T> struct tvec2 {
  T x, y;
  tvec2() : x(0), y(0) {}
  template <typename A> explicit tvec2(A s) : x(static_cast<T>(s)), y(static_cast<T>(s)) {}
  template <typename A, typename B> tvec2(A a, B b) : x(static_cast<T>(a)), y(static_cast<T>(b)) {}
  template <typename U> tvec2(const tvec2<U>& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
  tvec2& operator+=(const tvec2& o) { x += o.x; y += o.y; return *this; }
  tvec2& operator-=(const tvec2& o) { x -= o.x; y -= o.y; return *this; }
};

template <typename T> struct tvec3 {
  T x, y, z;
  tvec3() : x(0), y(0), z(0) {}
  template <typename A> explicit tvec3(A s) : x(static_cast<T>(s)), y(static_cast<T>(s)), z(static_cast<T>(s)) {}
  template <typename A, typename B, typename C>
  tvec3(A a, B b, C c) : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)) {}
  template <typename U> tvec3(const tvec3<U>& v)
      : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
  tvec3& operator+=(const tvec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
  tvec3& operator-=(const tvec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
};

template <typename T> struct tvec4 {
  T x, y, z, w;
  tvec4() : x(0), y(0), z(0), w(0) {}
  template <typename A> explicit tvec4(A s)
      : x(static_cast<T>(s)), y(static_cast<T>(s)), z(static_cast<T>(s)), w(static_cast<T>(s)) {}
  template <typename A, typename B, typename C, typename D>
  tvec4(A a, B b, C c, D d)
      : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)), w(static_cast<T>(d)) {}
  template <typename U> tvec4(const tvec4<U>& v)
      : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)), w(static_cast<T>(v.w)) {}
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
};

typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;
typedef tvec2<int> ivec2;
typedef tvec3<int> ivec3;
typedef tvec2<bool> bvec2;

// ---- component-wise arithmetic -------------------------------------------------
#define SMREF_GLM_BINOP(OP)                                                                          \
  template <typename T> tvec2<T> operator OP(const tvec2<T>& a, const tvec2<T>& b) {                 \
    return tvec2<T>(a.x OP b.x, a.y OP b.y); }                                                       \
  template <typename T> tvec2<T> operator OP(const tvec2<T>& a, T s) {                               \
    return tvec2<T>(a.x OP s, a.y OP s); }                                                           \
  template <typename T> tvec2<T> operator OP(T s, const tvec2<T>& a) {                               \
    return tvec2<T>(s OP a.x, s OP a.y); }                                                           \
  template <typename T> tvec3<T> operator OP(const tvec3<T>& a, const tvec3<T>& b) {                 \
    return tvec3<T>(a.x OP b.x, a.y OP b.y, a.z OP b.z); }                                           \
  template <typename T> tvec3<T> operator OP(const tvec3<T>& a, T s) {                               \
    return tvec3<T>(a.x OP s, a.y OP s, a.z OP s); }                                                 \
  template <typename T> tvec3<T> operator OP(T s, const tvec3<T>& a) {                               \
    return tvec3<T>(s OP a.x, s OP a.y, s OP a.z); }                                                 \
  template <typename T> tvec4<T> operator OP(const tvec4<T>& a, const tvec4<T>& b) {                 \
    return tvec4<T>(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); }                               \
  template <typename T> tvec4<T> operator OP(const tvec4<T>& a, T s) {                               \
    return tvec4<T>(a.x OP s, a.y OP s, a.z OP s, a.w OP s); }                                       \
  template <typename T> tvec4<T> operator OP(T s, const tvec4<T>& a) {                               \
    return tvec4<T>(s OP a.x, s OP a.y, s OP a.z, s OP a.w); }
SMREF_GLM_BINOP(+)
SMREF_GLM_BINOP(-)
SMREF_GLM_BINOP(*)
SMREF_GLM_BINOP(/)
#undef SMREF_GLM_BINOP

template <typename T> tvec2<T> operator-(const tvec2<T>& a) { return tvec2<T>(-a.x, -a.y); }
template <typename T> tvec3<T> operator-(const tvec3<T>& a) { return tvec3<T>(-a.x, -a.y, -a.z); }

// ---- geometric ---------------------------------------------------------------------
template <typename T> T dot(const tvec2<T>& a, const tvec2<T>& b) {
  tvec2<T> t(a * b);
  return t.x + t.y;
}
template <typename T> T dot(const tvec3<T>& a, const tvec3<T>& b) {
  tvec3<T> t(a * b);
  return t.x + t.y + t.z;
}
template <typename T> T length(const tvec2<T>& v) { return std::sqrt(dot(v, v)); }
template <typename T> T length(const tvec3<T>& v) { return std::sqrt(dot(v, v)); }
template <typename T> T inversesqrt(T x) { return static_cast<T>(1) / std::sqrt(x); }
template <typename T> tvec2<T> normalize(const tvec2<T>& v) { return v * inversesqrt(dot(v, v)); }
template <typename T> tvec3<T> normalize(const tvec3<T>& v) { return v * inversesqrt(dot(v, v)); }
template <typename T> tvec3<T> cross(const tvec3<T>& x, const tvec3<T>& y) {
  return tvec3<T>(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}

// ---- common ------------------------------------------------------------------------
template <typename T> tvec2<T> round(const tvec2<T>& v) { return tvec2<T>(std::round(v.x), std::round(v.y)); }
template <typename T> tvec2<T> floor(const tvec2<T>& v) { return tvec2<T>(std::floor(v.x), std::floor(v.y)); }
template <typename T> tvec2<T> fract(const tvec2<T>& v) { return v - floor(v); }

template <typename T, typename U> tvec2<T> mix(const tvec2<T>& x, const tvec2<T>& y, U a) {
  return tvec2<T>(tvec2<U>(x) * (static_cast<U>(1) - a) + tvec2<U>(y) * a);
}
template <typename T, typename U> tvec3<T> mix(const tvec3<T>& x, const tvec3<T>& y, U a) {
  return tvec3<T>(tvec3<U>(x) * (static_cast<U>(1) - a) + tvec3<U>(y) * a);
}
template <typename T, typename U> tvec4<T> mix(const tvec4<T>& x, const tvec4<T>& y, U a) {
  return tvec4<T>(tvec4<U>(x) * (static_cast<U>(1) - a) + tvec4<U>(y) * a);
}

// ---- relational --------------------------------------------------------------------
template <typename T> bvec2 greaterThanEqual(const tvec2<T>& a, const tvec2<T>& b) {
  return bvec2(a.x >= b.x, a.y >= b.y);
}
template <typename T> bvec2 lessThan(const tvec2<T>& a, const tvec2<T>& b) {
  return bvec2(a.x < b.x, a.y < b.y);
}
inline bool all(const bvec2& v) { return v.x && v.y; }

}  // namespace glm
