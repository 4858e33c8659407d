// HitRecord::pn and texture coordinates need the reference's barycentrics: the least-squares solution of
// [v0 v1 v2; 1 1 1] b = [P; 1] in double (triangle.cpp:12-29, Eigen colPivHouseholderQr().solve()).
//
// Solved in closed form instead of by a QR factorisation.  With s = b0 + b1 + b2, u = s - 1, e1 = v1 - v0, e2 = v2 - v0,
// q = P - v0 the problem reads  min |v0 u + e1 b1 + e2 b2 - q|^2 + u^2.  For a fixed u the best (b1, b2) is the in-plane
// solution (b1, b2)(q) - u (b1, b2)(v0)  (2x2 Gram system of e1, e2), which leaves  |(v0.n^) u - (q.n^)|^2 + u^2  along the
// unit normal n^, so  u = (v0.n)(q.n) / ((v0.n)^2 + n.n)  with n = e1 x e2.  Every piece is conditioned by the triangle's
// SHAPE only, where a factorisation of [v0 v1 v2; 1 1 1] is conditioned by |v| / (triangle size): against exact rational
// arithmetic on 0.01-sized triangles at distance 1000 this is accurate to 2.5e-16, a double QR / SVD solve to 9e-10
// (oracle-side check: tests/test_oracle.py).  About 70 multiply-adds and two divisions with short dependency chains,
// where the Householder version (3 square roots, 7 divisions, one long chain) was a quarter of k_shade's time.
#pragma once
#include "device_math.cuh"

namespace trt
{
__device__ __forceinline__ double dotd(double ax, double ay, double az, double bx, double by, double bz)
{
    return fma(ax, bx, fma(ay, by, az * bz));
}

__device__ inline void baryLeastSquares(const float *v9, float3 P, float &bx, float &by, float &bz)
{
    const double v0x = (double)v9[0], v0y = (double)v9[1], v0z = (double)v9[2];
    // differences of floats: exact in double unless the exponents are more than 2^29 apart
    const double e1x = (double)v9[3] - v0x, e1y = (double)v9[4] - v0y, e1z = (double)v9[5] - v0z;
    const double e2x = (double)v9[6] - v0x, e2y = (double)v9[7] - v0y, e2z = (double)v9[8] - v0z;
    const double qx = (double)P.x - v0x, qy = (double)P.y - v0y, qz = (double)P.z - v0z;
    const double g11 = dotd(e1x, e1y, e1z, e1x, e1y, e1z), g12 = dotd(e1x, e1y, e1z, e2x, e2y, e2z),
                 g22 = dotd(e2x, e2y, e2z, e2x, e2y, e2z);
    const double inv = 1.0 / fma(g11, g22, -(g12 * g12));
    const double a1 = dotd(e1x, e1y, e1z, qx, qy, qz), a2 = dotd(e2x, e2y, e2z, qx, qy, qz);
    const double c1 = dotd(e1x, e1y, e1z, v0x, v0y, v0z), c2 = dotd(e2x, e2y, e2z, v0x, v0y, v0z);
    const double p1 = fma(a1, g22, -(a2 * g12)) * inv, p2 = fma(a2, g11, -(a1 * g12)) * inv; // in-plane coordinates of q
    const double w1 = fma(c1, g22, -(c2 * g12)) * inv, w2 = fma(c2, g11, -(c1 * g12)) * inv; // ... and of v0
    const double nx = fma(e1y, e2z, -(e1z * e2y)), ny = fma(e1z, e2x, -(e1x * e2z)), nz = fma(e1x, e2y, -(e1y * e2x));
    const double vn = dotd(v0x, v0y, v0z, nx, ny, nz), qn = dotd(qx, qy, qz, nx, ny, nz), nn = dotd(nx, ny, nz, nx, ny, nz);
    const double u = vn * qn / fma(vn, vn, nn);
    const double b1 = fma(-u, w1, p1), b2 = fma(-u, w2, p2);
    bx = (float)(((1.0 + u) - b1) - b2), by = (float)b1, bz = (float)b2;
}

// bvh.cpp:224: pn = normalize(vn0*b.x + vn1*b.y + vn2*b.z)
__device__ __forceinline__ float3 shadingNormal(const float *vn9, float bx, float by, float bz)
{
    const float3 n0 = f3(vn9[0], vn9[1], vn9[2]), n1 = f3(vn9[3], vn9[4], vn9[5]), n2 = f3(vn9[6], vn9[7], vn9[8]);
    return normalize3((n0 * bx + n1 * by) + n2 * bz);
}
} // namespace trt
