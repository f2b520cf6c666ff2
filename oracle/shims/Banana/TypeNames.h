// Headless stand-in for the un-vendored Banana/TypeNames.h.  TEST INFRASTRUCTURE:
// only used to compile the reference's own Source/SceneManager.cpp, unmodified, into oracle/_ref/.
#pragma once
#include <vector>
#include <cmath>
#include <memory>
namespace Banana
{
template<class T> struct Vec3
{
    T v[3];
    Vec3() : v{ T(0), T(0), T(0) } {}
    explicit Vec3(T s) : v{ s, s, s } {}
    template<class A, class B, class C> Vec3(A a, B b, C c) : v{ static_cast<T>(a), static_cast<T>(b), static_cast<T>(c) } {}
    T&       operator [](int i) { return v[i]; }
    const T& operator [](int i) const { return v[i]; }
    Vec3& operator +=(const Vec3& o) { v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2]; return *this; }
};
template<class T> inline Vec3<T> operator +(const Vec3<T>& a, const Vec3<T>& b) { return Vec3<T>(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
template<class T> inline Vec3<T> operator -(const Vec3<T>& a, const Vec3<T>& b) { return Vec3<T>(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
template<class T> inline Vec3<T> operator *(T s, const Vec3<T>& a) { return Vec3<T>(s * a[0], s * a[1], s * a[2]); }
template<class T> inline Vec3<T> operator *(const Vec3<T>& a, T s) { return Vec3<T>(a[0] * s, a[1] * s, a[2] * s); }
template<class T> using Vec_Vec3 = std::vector<Vec3<T> >;
}
namespace glm
{
// glm::length(vec3) = sqrt(dot(v, v)), dot = (x*x + y*y) + z*z
inline float length(const Banana::Vec3<float>& a) { return std::sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]); }
}
