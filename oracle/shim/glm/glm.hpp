// TEST INFRASTRUCTURE ONLY (oracle tier A). Stand-in for the un-vendored glm dependency of the
// reference (triangle.h:2, scene.h:7, ...), so that the reference .cpp files compile UNMODIFIED.
// Only what the reference uses is provided; every formula restates glm 0.9.9's generic (scalar)
// definitions: dot = (x+y)+z of the products, normalize = v * (1/sqrt(dot)), min(x,y) = (y<x)?y:x,
// max(x,y) = (x<y)?y:x, vec/scalar = per-component IEEE divide.  "parity unpinned": the reference
// pins no glm version and ships no tests.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <string>

namespace glm
{
struct vec2
{
    float x, y;
    vec2() : x(0.f), y(0.f) {}
    template <typename A>
    explicit vec2(A s) : x(static_cast<float>(s)), y(static_cast<float>(s)) {}
    template <typename A, typename B>
    vec2(A a, B b) : x(static_cast<float>(a)), y(static_cast<float>(b)) {}
};

struct vec3
{
    float x, y, z;
    vec3() : x(0.f), y(0.f), z(0.f) {}
    template <typename A>
    explicit vec3(A s) : x(static_cast<float>(s)), y(static_cast<float>(s)), z(static_cast<float>(s)) {}
    template <typename A, typename B, typename C>
    vec3(A a, B b, C c) : x(static_cast<float>(a)), y(static_cast<float>(b)), z(static_cast<float>(c)) {}
    vec3 &operator+=(const vec3 &o)
    {
        x += o.x, y += o.y, z += o.z;
        return *this;
    }
    float &operator[](int i) { return (&x)[i]; }
    const float &operator[](int i) const { return (&x)[i]; }
};

inline vec3 operator+(const vec3 &a, const vec3 &b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3 &a, const vec3 &b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3 &a, const vec3 &b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3 &a, const vec3 &b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(const vec3 &a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3 &a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3 &a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator-(const vec3 &a) { return vec3(-a.x, -a.y, -a.z); }
inline bool operator==(const vec3 &a, const vec3 &b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(const vec3 &a, const vec3 &b) { return !(a == b); }

template <typename T>
inline T min(T x, T y) { return (y < x) ? y : x; }
template <typename T>
inline T max(T x, T y) { return (x < y) ? y : x; }
inline vec3 min(const vec3 &a, const vec3 &b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline vec3 max(const vec3 &a, const vec3 &b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
template <typename T>
inline T clamp(T x, T lo, T hi) { return min(max(x, lo), hi); }
template <typename T>
inline T radians(T degrees) { return degrees * static_cast<T>(0.01745329251994329576923690768489); }

inline float dot(const vec3 &a, const vec3 &b)
{
    vec3 tmp(a * b);
    return tmp.x + tmp.y + tmp.z;
}
inline vec3 cross(const vec3 &x, const vec3 &y)
{
    return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
inline float length(const vec3 &v) { return std::sqrt(dot(v, v)); }
inline float length2(const vec3 &v) { return dot(v, v); }
inline vec3 normalize(const vec3 &v) { return v * (1.0f / std::sqrt(dot(v, v))); }
inline vec3 reflect(const vec3 &I, const vec3 &N) { return I - N * dot(N, I) * 2.0f; }
inline vec3 refract(const vec3 &I, const vec3 &N, float eta)
{
    float const d = dot(N, I);
    float const k = 1.0f - eta * eta * (1.0f - d * d);
    return (k >= 0.0f) ? (eta * I - (eta * d + std::sqrt(k)) * N) : vec3(0);
}
} // namespace glm
