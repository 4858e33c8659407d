// Device float3 helpers in the operation order the reference inherits from glm's scalar path
// (SURVEY App. A.1).  The library is compiled with --fmad=false, so `a*b + c` below is two IEEE roundings,
// exactly like the reference's x86-64 build without FMA; fused operations are only ever written as explicit
// __fmaf_rn in code that does not need to match reference arithmetic (conservative culling tests).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace trt
{
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ float3 cross3(float3 a, float3 b)
{
    return f3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
__device__ __forceinline__ float3 normalize3(float3 v) { return v * (1.0f / sqrtf(dot3(v, v))); }
__device__ __forceinline__ float length3(float3 v) { return sqrtf(dot3(v, v)); }
// glm::min / glm::max: NaN in the second argument returns the first
__device__ __forceinline__ float gmin(float x, float y) { return (y < x) ? y : x; }
__device__ __forceinline__ float gmax(float x, float y) { return (x < y) ? y : x; }
} // namespace trt
