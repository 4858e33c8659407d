// Closest-hit traversal: device code shared by the fixed-batch kernels (trace.cu) and the wavefront path
// tracer (wavefront.cu).  Replaces traverseBVH / interactAABB / interactBVHNode / interactTriangle
// (bvh.cpp:146-245).  All arithmetic that decides a hit is un-fused IEEE float in the reference's order.
//
// Exactness argument: the reference walks EVERY child whose padded box passes interactAABB
// and keeps the nearest hit with the tie rule of bvh.cpp:168-172,219.  That rule is a total order on
// (t, key) with a per-triangle key computed at flatten time, so any visiting order gives the
// reference's winner as long as every leaf whose own box passes is scanned unless it provably cannot win:
//   * ancestor boxes contain descendant boxes exactly (min/max of the same floats) and the slab test is
//     monotone in the box, so "leaf box passes" implies "all ancestors pass" — testing fewer or other
//     enclosing boxes cannot add or lose leaves (rays with an exactly-zero direction component, where
//     inf*0 = NaN breaks monotonicity, are routed to the exhaustive walk);
//   * a subtree is skipped only when its entry distance exceeds the best t found so far.
#pragma once
#include "accel.h"
#include "device_math.cuh"

namespace trt
{
struct Hit
{
    float t;      // TRT_INF on miss
    int32_t id;   // post-build triangle index, -1 on miss
    uint32_t key; // tie key of the current winner
};

#define TRT_REF_STACK TRT_REF_STACK_LIMIT

// interactAABB (bvh.cpp:231-245): returns whether the reference would descend (result > 0); t0 = entry distance.
__device__ __forceinline__ bool boxPass(float3 S, float3 inv, float ax, float ay, float az, float bx, float by, float bz,
                                        float &t0)
{
    const float inx = (bx - S.x) * inv.x, iny = (by - S.y) * inv.y, inz = (bz - S.z) * inv.z;
    const float outx = (ax - S.x) * inv.x, outy = (ay - S.y) * inv.y, outz = (az - S.z) * inv.z;
    // tmax = glm::max(in, out), tmin = glm::min(in, out) — first argument `in`
    const float t1 = gmin(gmax(inx, outx), gmin(gmax(iny, outy), gmax(inz, outz)));
    t0 = gmax(gmin(inx, outx), gmax(gmin(iny, outy), gmin(inz, outz)));
    return (t1 >= t0) && (((t0 > 0.0f) ? t0 : t1) > 0.0f);
}

// interactTriangle (bvh.cpp:177-209) with one extra early-out: a candidate farther than the current best can
// never be accepted, so its inside test is skipped.
__device__ __forceinline__ bool triangleHit(const TriGeom &g, float3 S, float3 d, float best_t, float &t_out)
{
    const float3 N = f3(g.p1nx.w, g.p2ny.w, g.p3nz.w);
    const float dn = dot3(N, d);
    if (fabsf(dn) < 0.00001f)
        return false;
    const float3 p1 = f3(g.p1nx.x, g.p1nx.y, g.p1nx.z);
    const float t = dot3(p1 - S, N) / dn;
    if (t < 0.0005f)
        return false;
    if (t > best_t)
        return false;
    const float3 p2 = f3(g.p2ny.x, g.p2ny.y, g.p2ny.z), p3 = f3(g.p3nz.x, g.p3nz.y, g.p3nz.z);
    const float3 P = S + d * t;
    const float dir1 = dot3(cross3(p2 - p1, P - p1), N);
    const float dir2 = dot3(cross3(p3 - p2, P - p2), N);
    const float dir3 = dot3(cross3(p1 - p3, P - p3), N);
    t_out = t;
    return (dir1 > 0.f && dir2 > 0.f && dir3 > 0.f) || (dir1 < 0.f && dir2 < 0.f && dir3 < 0.f);
}

// interactBVHNode (bvh.cpp:211-229) over one reference leaf, merged into the running best by (t, key).
__device__ __forceinline__ void scanLeaf(const SceneView &sv, int first, int num, float3 S, float3 d, Hit &hit)
{
    for (int i = first; i < first + num; ++i)
    {
        const float4 *gp = reinterpret_cast<const float4 *>(sv.tri_geom + i);
        TriGeom g;
        g.p1nx = __ldg(gp), g.p2ny = __ldg(gp + 1), g.p3nz = __ldg(gp + 2);
        float t;
        if (triangleHit(g, S, d, hit.t, t))
        {
            if (t < hit.t)
            {
                hit.t = t, hit.id = i, hit.key = 0xFFFFFFFFu; // key fetched lazily, only if a tie shows up
            }
            else if (t == hit.t)
            {
                if (hit.key == 0xFFFFFFFFu)
                    hit.key = (hit.id < 0) ? TRT_MISS_KEY : __ldg(sv.tri_key + hit.id);
                const uint32_t k = __ldg(sv.tri_key + i);
                if (k > hit.key)
                    hit.id = i, hit.key = k;
            }
        }
    }
}

__device__ __forceinline__ float3 rcpDir(float3 d)
{
    // (float)(1.0 / d) of bvh.cpp:233: a correctly rounded single division (double rounding is innocuous)
    return f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
}

// Walk of the reference binary topology.  EXHAUSTIVE: the reference's own visiting rule (both children whose
// box passes, no pruning).  Otherwise near-child-first with entry-distance pruning.
template <bool EXHAUSTIVE>
__device__ __forceinline__ void traceRefTopology(const SceneView &sv, float3 S, float3 d, Hit &hit)
{
    hit.t = TRT_INF, hit.id = -1, hit.key = 0xFFFFFFFFu;
    if (sv.root_link == 0x7fffffff)
        return;
    const float3 inv = rcpDir(d);
    int32_t stack_link[TRT_REF_STACK];
    float stack_t[TRT_REF_STACK];
    int sp = 0;
    int32_t cur = sv.root_link;
    for (;;)
    {
        if (cur < 0)
        {
            const int leaf = ~cur;
            scanLeaf(sv, leaf >> 3, (leaf & 7) + 1, S, d, hit);
        }
        else
        {
            const float4 *np = reinterpret_cast<const float4 *>(sv.ref_nodes + cur);
            const float4 a = __ldg(np), b = __ldg(np + 1), c = __ldg(np + 2);
            const int4 lk = __ldg(reinterpret_cast<const int4 *>(np + 3));
            float tl, tr;
            bool hl = boxPass(S, inv, a.x, a.y, a.z, a.w, b.x, b.y, tl);
            bool hr = boxPass(S, inv, b.z, b.w, c.x, c.y, c.z, c.w, tr);
            if (!EXHAUSTIVE)
            {
                hl = hl && !(tl > hit.t);
                hr = hr && !(tr > hit.t);
            }
            if (hl && hr)
            {
                const bool leftFirst = EXHAUSTIVE || !(tr < tl);
                stack_link[sp] = leftFirst ? lk.y : lk.x;
                stack_t[sp] = leftFirst ? tr : tl;
                ++sp;
                cur = leftFirst ? lk.x : lk.y;
                continue;
            }
            if (hl || hr)
            {
                cur = hl ? lk.x : lk.y;
                continue;
            }
        }
        // pop
        for (;;)
        {
            if (sp == 0)
                return;
            --sp;
            cur = stack_link[sp];
            if (EXHAUSTIVE || !(stack_t[sp] > hit.t))
                break;
        }
    }
}

// ---- fast layout -------------------------------------------------------------------------------------------
// Rays for which the slab test is not monotone in the box (a direction component that is exactly +-0 gives
// inf*0 = NaN; non-finite origins / directions) take the reference's own exhaustive walk.
__device__ __forceinline__ bool needsStrictWalk(float3 S, float3 d)
{
    // 1/d overflows to +-inf for zero AND for denormal components: test the reciprocal itself
    const float3 inv = rcpDir(d);
    const bool inf_rcp = !(fabsf(inv.x) <= 3.4028235e38f) || !(fabsf(inv.y) <= 3.4028235e38f) || !(fabsf(inv.z) <= 3.4028235e38f);
    const float sum = ((S.x + S.y) + S.z) + ((d.x + d.y) + d.z); // inf or NaN anywhere -> not finite
    return inf_rcp || !(fabsf(sum) < 3.0e38f);
}

struct TraceCounters
{
    uint32_t nodes, boxes, leaves, tris;
};

// Slab test of one child of a wide node: the reference's arithmetic ((B - S) * inv, un-fused).  For rays
// admitted here every operand is finite and inv is finite and non-zero, so no NaN can appear and IEEE
// fminf / fmaxf (one instruction) equal glm's compare-select min / max.
__device__ __forceinline__ bool childPass(float3 S, float3 inv, float ax, float ay, float az, float bx, float by, float bz,
                                          float &t0)
{
    const float inx = (bx - S.x) * inv.x, iny = (by - S.y) * inv.y, inz = (bz - S.z) * inv.z;
    const float outx = (ax - S.x) * inv.x, outy = (ay - S.y) * inv.y, outz = (az - S.z) * inv.z;
    const float t1 = fminf(fmaxf(inx, outx), fminf(fmaxf(iny, outy), fmaxf(inz, outz)));
    t0 = fmaxf(fminf(inx, outx), fmaxf(fminf(iny, outy), fminf(inz, outz)));
    return (t1 >= t0) && (((t0 > 0.0f) ? t0 : t1) > 0.0f);
}

// 4-wide walk over the reference leaves: while-while (inner nodes until a leaf is reached, then the leaf scan),
// nearest child first, entry-distance pruning at push and at pop.
template <bool STATS>
__device__ __forceinline__ void traceWide(const SceneView &sv, float3 S, float3 d, Hit &hit, TraceCounters *cnt)
{
    hit.t = TRT_INF, hit.id = -1, hit.key = 0xFFFFFFFFu;
    if (sv.wide_root == TRT_LINK_EMPTY)
        return;
    const float3 inv = rcpDir(d);
    int32_t stack_link[TRT_WIDE_STACK];
    float stack_t[TRT_WIDE_STACK];
    stack_link[0] = TRT_LINK_EXIT;
    stack_t[0] = -1.f;
    int sp = 1;
    int32_t cur = sv.wide_root;
    for (;;)
    {
        while (cur >= 0)
        {
            const float4 *np = reinterpret_cast<const float4 *>(sv.wide_nodes + cur);
            const float4 lox = __ldg(np), loy = __ldg(np + 1), loz = __ldg(np + 2);
            const float4 hix = __ldg(np + 3), hiy = __ldg(np + 4), hiz = __ldg(np + 5);
            const int4 lk = __ldg(reinterpret_cast<const int4 *>(np + 6));
            float t0, t1, t2, t3;
            bool h0 = childPass(S, inv, lox.x, loy.x, loz.x, hix.x, hiy.x, hiz.x, t0);
            bool h1 = childPass(S, inv, lox.y, loy.y, loz.y, hix.y, hiy.y, hiz.y, t1);
            bool h2 = (lk.z != TRT_LINK_EMPTY) && childPass(S, inv, lox.z, loy.z, loz.z, hix.z, hiy.z, hiz.z, t2);
            bool h3 = (lk.w != TRT_LINK_EMPTY) && childPass(S, inv, lox.w, loy.w, loz.w, hix.w, hiy.w, hiz.w, t3);
            if (STATS)
                cnt->nodes++, cnt->boxes += 2 + (lk.z != TRT_LINK_EMPTY) + (lk.w != TRT_LINK_EMPTY);
            h0 = h0 && !(t0 > hit.t);
            h1 = h1 && !(t1 > hit.t);
            h2 = h2 && !(t2 > hit.t);
            h3 = h3 && !(t3 > hit.t);
            // nearest hit child continues, the others are pushed (far ones first would need a full sort; the
            // pop-time distance check prunes them anyway)
            float tn = 3.0e38f;
            int32_t next = TRT_LINK_EMPTY;
            if (h0)
                tn = t0, next = lk.x;
            if (h1 && t1 < tn)
                tn = t1, next = lk.y;
            if (h2 && t2 < tn)
                tn = t2, next = lk.z;
            if (h3 && t3 < tn)
                tn = t3, next = lk.w;
            if (next == TRT_LINK_EMPTY)
            {
                // nothing hit: pop
                do
                {
                    --sp;
                    cur = stack_link[sp];
                } while (stack_t[sp] > hit.t);
                continue;
            }
            if (h0 && lk.x != next)
                stack_link[sp] = lk.x, stack_t[sp] = t0, ++sp;
            if (h1 && lk.y != next)
                stack_link[sp] = lk.y, stack_t[sp] = t1, ++sp;
            if (h2 && lk.z != next)
                stack_link[sp] = lk.z, stack_t[sp] = t2, ++sp;
            if (h3 && lk.w != next)
                stack_link[sp] = lk.w, stack_t[sp] = t3, ++sp;
            cur = next;
        }
        if (cur == TRT_LINK_EXIT)
            return;
        const int leaf = ~cur;
        if (STATS)
            cnt->leaves++, cnt->tris += (leaf & 7) + 1;
        scanLeaf(sv, leaf >> 3, (leaf & 7) + 1, S, d, hit);
        do
        {
            --sp;
            cur = stack_link[sp];
        } while (stack_t[sp] > hit.t);
    }
}

// Closest hit with the default (fast) layout; falls back to the reference topology for strict rays and for
// scenes without a wide layout.
__device__ __forceinline__ void traceClosest(const SceneView &sv, float3 S, float3 d, Hit &hit)
{
    if (!sv.use_wide)
        traceRefTopology<false>(sv, S, d, hit);
    else if (needsStrictWalk(S, d))
        traceRefTopology<true>(sv, S, d, hit);
    else
        traceWide<false>(sv, S, d, hit, nullptr);
}
} // namespace trt
