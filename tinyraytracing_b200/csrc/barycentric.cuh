// HitRecord::pn and texture coordinates need the reference's barycentrics: the least-squares solution of
// [v0 v1 v2; 1 1 1] b = [P; 1] in double (triangle.cpp:12-29, Eigen colPivHouseholderQr().solve()).
// Same column-pivoted Householder algorithm here, in registers, un-fused double arithmetic.
#pragma once
#include "device_math.cuh"

namespace trt
{
__device__ inline void baryLeastSquares(const float *v9, float3 P, float &bx, float &by, float &bz)
{
    double A[4][3] = {{(double)v9[0], (double)v9[3], (double)v9[6]},
                      {(double)v9[1], (double)v9[4], (double)v9[7]},
                      {(double)v9[2], (double)v9[5], (double)v9[8]},
                      {1.0, 1.0, 1.0}};
    double b[4] = {(double)P.x, (double)P.y, (double)P.z, 1.0};
    int perm[3] = {0, 1, 2};
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        int best = k;
        double bn = -1.0;
#pragma unroll
        for (int j = k; j < 3; ++j)
        {
            double s = 0.0;
#pragma unroll
            for (int i = k; i < 4; ++i)
                s += A[i][j] * A[i][j];
            if (s > bn)
                bn = s, best = j;
        }
        // column swap k <-> best with compile-time indices only (keeps A and perm in registers)
#pragma unroll
        for (int j = k + 1; j < 3; ++j)
        {
            const bool sw = (best == j);
#pragma unroll
            for (int i = 0; i < 4; ++i)
            {
                const double a = A[i][k], b = A[i][j];
                A[i][k] = sw ? b : a;
                A[i][j] = sw ? a : b;
            }
            const int pa = perm[k], pb = perm[j];
            perm[k] = sw ? pb : pa;
            perm[j] = sw ? pa : pb;
        }
        const double norm = sqrt(bn);
        if (norm == 0.0)
            continue;
        const double alpha = (A[k][k] > 0.0) ? -norm : norm;
        double w[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int i = k; i < 4; ++i)
            w[i] = A[i][k];
        w[k] -= alpha;
        double wtw = 0.0;
#pragma unroll
        for (int i = k; i < 4; ++i)
            wtw += w[i] * w[i];
        if (wtw == 0.0)
            continue;
        const double beta = 2.0 / wtw;
#pragma unroll
        for (int j = k; j < 3; ++j)
        {
            double s = 0.0;
#pragma unroll
            for (int i = k; i < 4; ++i)
                s += w[i] * A[i][j];
            s *= beta;
#pragma unroll
            for (int i = k; i < 4; ++i)
                A[i][j] -= s * w[i];
        }
        double s = 0.0;
#pragma unroll
        for (int i = k; i < 4; ++i)
            s += w[i] * b[i];
        s *= beta;
#pragma unroll
        for (int i = k; i < 4; ++i)
            b[i] -= s * w[i];
    }
    double y[3];
#pragma unroll
    for (int k = 2; k >= 0; --k)
    {
        double s = b[k];
#pragma unroll
        for (int j = k + 1; j < 3; ++j)
            s -= A[k][j] * y[j];
        y[k] = (A[k][k] != 0.0) ? s / A[k][k] : 0.0;
    }
    double r[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        // perm[k] is data dependent: select without dynamic register indexing
        r[0] = (perm[k] == 0) ? y[k] : r[0];
        r[1] = (perm[k] == 1) ? y[k] : r[1];
        r[2] = (perm[k] == 2) ? y[k] : r[2];
    }
    bx = (float)r[0], by = (float)r[1], bz = (float)r[2];
}

// bvh.cpp:224: pn = normalize(vn0*b.x + vn1*b.y + vn2*b.z)
__device__ __forceinline__ float3 shadingNormal(const float *vn9, float bx, float by, float bz)
{
    const float3 n0 = f3(vn9[0], vn9[1], vn9[2]), n1 = f3(vn9[3], vn9[4], vn9[5]), n2 = f3(vn9[6], vn9[7], vn9[8]);
    return normalize3((n0 * bx + n1 * by) + n2 * bz);
}
} // namespace trt
