// Closest-hit traversal: device code shared by the fixed-batch kernels (trace.cu) and the wavefront path
// tracer (wavefront.cu).  Replaces traverseBVH / interactAABB / interactBVHNode / interactTriangle
// (bvh.cpp:146-245).  All arithmetic that decides a hit is un-fused IEEE float in the reference's order.
//
// Exactness argument: the reference walks EVERY child whose padded box passes interactAABB and keeps
// the nearest hit with the tie rule of bvh.cpp:168-172,219.  That rule is a total order on (t, key) with a
// per-triangle key computed at flatten time, so any visiting order gives the reference's winner as long
// as every triangle the reference tests is tested unless it provably cannot win, and no triangle it does NOT test is
// ever reported:
//   * a triangle is tested by the reference iff its leaf's box and all ancestor boxes pass; ancestor boxes contain
//     descendant boxes exactly and the slab test is monotone in the box, so "leaf box passes" implies the rest
//     (rays with a direction component whose reciprocal overflows, where inf*0 = NaN breaks monotonicity, and rays
//     with non-finite or very distant origins are routed to the exhaustive walk of the reference topology);
//   * the fast layout is its own tree over triangles: its boxes only cull (they contain every point at which
//     interactTriangle can accept the triangle), and a candidate counts only if its REFERENCE leaf's box passes;
//   * a subtree is skipped only when its entry distance exceeds the best t found so far.
#pragma once
#include "accel.h"
#include "device_math.cuh"

namespace trt
{
// 256-bit global load (LDG.E.256, sm_100): p must be 32-byte aligned
__device__ __forceinline__ void ldg256(const void *p, float4 &a, float4 &b)
{
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

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

// plane part: dn = dot(N,d), t = dot(p1-S,N)/dn with the rejections of bvh.cpp:185,189 and the "cannot win" cut
__device__ __forceinline__ bool trianglePlane(float4 q0, float4 q1, float3 S, float3 d, float best_t, float &t_out)
{
    const float3 N = f3(q0.x, q0.y, q0.z);
    const float dn = dot3(N, d);
    if (fabsf(dn) < 0.00001f)
        return false;
    const float3 p1 = f3(q0.w, q1.x, q1.y);
    const float a = dot3(p1 - S, N);
    // opposite signs (or a == +-0 against a negative dn, ...) give t <= -0 < 0.0005: rejected by bvh.cpp:189
    // whatever the quotient is, so the IEEE division is skipped (a NaN operand is a miss on either path).
    // The same holds for |a| < 4e-9: |dn| >= 0.00001 here, so |t| < 0.00041 < 0.0005 whatever the quotient.  This is not
    // a rare case: a ray that leaves an axis-aligned surface starts EXACTLY on the plane of every coplanar triangle about
    // half the time (a == +-0), and FCHK sends a zero dividend to the division's ~110-instruction slow path — ncu on
    // the staircase render walk: taken by 6 of the 13 lanes that divide, in two thirds of the leaf scans.
    if (((__float_as_int(a) ^ __float_as_int(dn)) < 0) || fabsf(a) < 4.0e-9f)
        return false;
    const float t = a / dn;
    if (t < 0.0005f)
        return false;
    if (t > best_t)
        return false;
    t_out = t;
    return true;
}

// inside part (bvh.cpp:191-200): three edge cross products against N, all > 0 or all < 0
__device__ __forceinline__ bool triangleInside(float4 q0, float4 q1, float4 q2, float3 S, float3 d, float t)
{
    const float3 N = f3(q0.x, q0.y, q0.z);
    const float3 p1 = f3(q0.w, q1.x, q1.y), p2 = f3(q1.z, q1.w, q2.x), p3 = f3(q2.y, q2.z, q2.w);
    const float3 P = S + d * t;
    const float dir1 = dot3(cross3(p2 - p1, P - p1), N);
    const float dir2 = dot3(cross3(p3 - p2, P - p2), N);
    const float dir3 = dot3(cross3(p1 - p3, P - p3), N);
    return (dir1 > 0.f && dir2 > 0.f && dir3 > 0.f) || (dir1 < 0.f && dir2 < 0.f && dir3 < 0.f);
}

// interactTriangle (bvh.cpp:177-209) with one extra early-out: a candidate farther than the current best can
// never be accepted, so its inside test is skipped.
__device__ __forceinline__ bool triangleHit(const TriGeom &g, float3 S, float3 d, float best_t, float &t_out)
{
    float t;
    if (!trianglePlane(g.q0, g.q1, S, d, best_t, t))
        return false;
    t_out = t;
    return triangleInside(g.q0, g.q1, g.q2, S, d, t);
}

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

// Culling test of one of the FAST layout's own child boxes (walkNodeStep / traceWide).  These boxes never decide a
// hit — a candidate counts only if its reference leaf's box passes childPass above — they only have to contain, with
// the 256-ulp pad they carry, every point at which a triangle below them can be accepted.  So the test need not be the
// reference's arithmetic, only conservative within that pad:
//   * (b - S) * inv is evaluated as fma(b, inv, -(S * inv)): one instruction per plane instead of two.  The result
//     differs from the exact value by at most 2^-24 |S inv| (the rounding of the precomputed product), i.e. it is the
//     exact slab distance of an origin moved by half an ulp of |S| <= 4 x scene scale — 1/64 of the pad;
//   * the far distance is clamped to the best t so far and the near distance to 0, which folds the reference's
//     "result > 0" rule and the pruning test into one comparison: pass <=> max(t0, 0) <= min(t1, best).  (t1 == 0 with
//     t0 <= 0 passes here and not in the reference: a superset, harmless for a culling test.)
// inv is finite here (cullInv clamps it), so no inf - inf can appear; a NaN would be dropped by fminf / fmaxf, which
// only ever widens the test.  31 -> 18 instructions per child box: the slab tests were 28 % of the walker's issued
// instructions (profiles/r01_closest_staircase_default.txt).
__device__ __forceinline__ bool childCull(float3 inv, float3 nsi, float ax, float ay, float az, float bx, float by, float bz,
                                          float best_t, float &t0c)
{
    const float inx = __fmaf_rn(bx, inv.x, nsi.x), iny = __fmaf_rn(by, inv.y, nsi.y), inz = __fmaf_rn(bz, inv.z, nsi.z);
    const float outx = __fmaf_rn(ax, inv.x, nsi.x), outy = __fmaf_rn(ay, inv.y, nsi.y), outz = __fmaf_rn(az, inv.z, nsi.z);
    const float t1 = fminf(fminf(fmaxf(inx, outx), fmaxf(iny, outy)), fminf(fmaxf(inz, outz), best_t));
    t0c = fmaxf(fmaxf(fminf(inx, outx), fminf(iny, outy)), fmaxf(fminf(inz, outz), 0.0f));
    return t0c <= t1;
}

// interactBVHNode (bvh.cpp:211-229) over one reference leaf, merged into the running best by (t, key).
// One fused loop per lane.  Two alternatives were measured on staircase (4 Mi config-2 rays) and rejected:
// splitting into a plane pass + an inside pass per lane (1.70 vs 2.13 Grays/s: the reloads and the recomputed
// division cost more than the divergence they remove), and pooling the warp's candidates in shared memory
// (walkPersistent<POOLED = true>, 1.93 Grays/s, latency-bound).
// FAST = false: triangles in post-build order (reference-topology walks; hit.id is the post-build index).
// FAST = true: the fast layout's own leaf order; a candidate counts only if the REFERENCE leaf that holds the
// triangle passes the reference's slab test for this ray (the reference never tests a triangle otherwise), and
// hit.id is the fast index until the walk ends (toPostBuildId).
// The boxes the reference tests on its way from the root to reference leaf `leaf` (interactAABB on every node of the
// path, bvh.cpp:156-166), with the reference's own arithmetic including its NaN behaviour: the leaf is scanned by the
// reference iff all of them pass.  Used instead of the leaf-box-only gate for rays with a direction component whose
// reciprocal overflows, where "leaf box passes" no longer implies "its ancestors pass".
// (out of line and with plain pointer arguments: the rare path must not weigh on the register allocation of the walks)
static __device__ __noinline__ bool refPathPasses(const int32_t *ref_leaf_parent, const RefNode *ref_nodes, int leaf, float3 S,
                                                  float3 d)
{
    const float3 inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z); // the reference's own reciprocal (bvh.cpp:233), +-inf included
    int32_t p = __ldg(ref_leaf_parent + leaf);
    while (p >= 0)
    {
        const float4 *np = reinterpret_cast<const float4 *>(ref_nodes + (p >> 1));
        const float4 a = __ldg(np), b = __ldg(np + 1), c = __ldg(np + 2);
        float t0;
        const bool ok = (p & 1) ? boxPass(S, inv, b.z, b.w, c.x, c.y, c.z, c.w, t0) : boxPass(S, inv, a.x, a.y, a.z, a.w, b.x, b.y, t0);
        if (!ok)
            return false;
        p = __ldg(reinterpret_cast<const int32_t *>(np + 3) + 2);
    }
    return true;
}

template <bool FAST>
__device__ __forceinline__ void scanLeaf(const SceneView &sv, int first, int num, float3 S, float3 d, float3 inv, Hit &hit,
                                         bool pathGate = false)
{
    const TriGeom *geom = FAST ? sv.fast_geom : sv.tri_geom;
    const uint32_t *keys = FAST ? sv.fast_key : sv.tri_key;
    for (int i = first; i < first + num; ++i)
    {
        // 48-byte records, three 128-bit loads (padding to 64 B for two 256-bit loads was measured: no gain)
        const float4 *gp = reinterpret_cast<const float4 *>(geom + i);
        TriGeom g;
        g.q0 = __ldg(gp), g.q1 = __ldg(gp + 1), g.q2 = __ldg(gp + 2);
        float t;
        if (!trianglePlane(g.q0, g.q1, S, d, hit.t, t))
            continue;
        // the reference leaf of a candidate is fetched before the inside test, not after it: the gate below is two
        // dependent loads (leaf index, then its box), and the first one's latency now hides behind the inside test
        int32_t ref_leaf = 0;
        if (FAST)
            ref_leaf = __ldg(sv.fast_leaf + i);
        if (triangleInside(g.q0, g.q1, g.q2, S, d, t))
        {
            if (FAST && pathGate)
            {
                if (!refPathPasses(sv.ref_leaf_parent, sv.ref_nodes, ref_leaf, S, d))
                    continue;
            }
            else if (FAST && sv.check_leaf_box)
            {
                const float4 *bp = sv.ref_leaf_box + 2 * ref_leaf;
                const float4 lo = __ldg(bp), hi = __ldg(bp + 1);
                float t0;
                if (!childPass(S, inv, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, t0))
                    continue;
            }
            if (t < hit.t)
            {
                hit.t = t, hit.id = i, hit.key = 0xFFFFFFFFu; // key fetched lazily, only if a tie shows up
            }
            else if (t == hit.t)
            {
                if (hit.key == 0xFFFFFFFFu)
                    hit.key = (hit.id < 0) ? TRT_MISS_KEY : __ldg(keys + hit.id);
                const uint32_t k = __ldg(keys + i);
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
            scanLeaf<false>(sv, leaf >> 3, (leaf & 7) + 1, S, d, inv, hit);
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
// 0: ordinary ray.  2: non-finite origin / direction, or an origin so far out that the rounding of S + d*t exceeds what
// the fast layout's box pad was sized for: the reference's own exhaustive walk.  1: finite ray with a direction component
// that is zero or denormal (1/d overflows, e.g. the centre column of an axis-aligned camera): the fast layout is still
// walked — with inv = +-inf a child box is passed iff the ray's coordinate lies strictly inside its slab, entry distances
// are +-inf and prune correctly, and a NaN (coordinate exactly on a box plane) makes fminf / fmaxf drop the box, which is
// safe because an acceptable hit point lies strictly inside the padded boxes of its triangle — but a candidate is
// accepted only if EVERY box on the reference's path to its leaf passes the reference's test (refPathPasses).  An
// exhaustive walk for these rays cost 1.5 ms PER RAY on a 100 k-triangle scene and set the time of the whole launch.
__device__ __forceinline__ int rayClass(const SceneView &sv, float3 S, float3 d)
{
    // 1/d overflows to +-inf for zero AND for denormal components: test the reciprocal itself
    // (and for components so small that b * inv could overflow in the culling test's fused form, see cullInv)
    const float3 inv = rcpDir(d);
    const float lim = sv.inv_cull_limit;
    const bool inf_rcp = !(fabsf(inv.x) <= lim) || !(fabsf(inv.y) <= lim) || !(fabsf(inv.z) <= lim);
    const float sum = ((S.x + S.y) + S.z) + ((d.x + d.y) + d.z); // inf or NaN anywhere -> not finite
    const bool far = fmaxf(fabsf(S.x), fmaxf(fabsf(S.y), fabsf(S.z))) > sv.strict_origin_limit;
    if (far || !(fabsf(sum) < 3.0e38f))
        return 2;
    return inf_rcp ? 1 : 0;
}
__device__ __forceinline__ bool needsStrictWalk(const SceneView &sv, float3 S, float3 d) { return rayClass(sv, S, d) != 0; }

// Reciprocal direction for the culling tests of a class-1 ray: a component beyond the limit (the ray does not move
// along that axis within any t <= INF: |d| < 1e-30 x scene scale) is clamped to +-limit, so that fma(b, inv, -S inv)
// stays finite and keeps the sign of b - S — the slab is passed iff the origin lies inside it, which is what
// inv = +-inf meant in the un-fused form.  Class-0 rays are below the limit: their culling reciprocal IS 1/d.
__device__ __forceinline__ float3 cullInv(const SceneView &sv, float3 inv)
{
    const float lim = sv.inv_cull_limit;
    return f3(copysignf(fminf(fabsf(inv.x), lim), inv.x), copysignf(fminf(fabsf(inv.y), lim), inv.y),
              copysignf(fminf(fabsf(inv.z), lim), inv.z));
}

// A 128-byte wide node in four 256-bit loads (LDG.E.256, sm_100) instead of seven 128-bit ones: with every lane on
// a different node the L1 data pipe pays one wavefront per lane per load instruction, and ncu shows that pipe at 73 %
// of peak on staircase (profiles/r01_closest_staircase_default.txt).
struct NodeRegs
{
    float4 lox, loy, loz, hix, hiy, hiz;
    int4 lk;
};
__device__ __forceinline__ NodeRegs loadWideNode(const WideNode *n)
{
    NodeRegs r;
    float4 l, pad;
    const char *p = reinterpret_cast<const char *>(n);
    ldg256(p, r.lox, r.loy);
    ldg256(p + 32, r.loz, r.hix);
    ldg256(p + 64, r.hiy, r.hiz);
    ldg256(p + 96, l, pad);
    r.lk = make_int4(__float_as_int(l.x), __float_as_int(l.y), __float_as_int(l.z), __float_as_int(l.w));
    return r;
}

// Traversal stack entry: (entry distance, link) packed in 64 bits, so that a push / pop is ONE local-memory access.
// With every lane at its own stack depth each access costs a wavefront per lane in the L1 data pipe — the pipe that
// bounds this kernel on staircase — so halving the stack instructions matters as much as the node loads.
typedef unsigned long long StackEntry;
__device__ __forceinline__ StackEntry packEntry(float t, int32_t link)
{
    return ((StackEntry)__float_as_uint(t) << 32) | (StackEntry)(uint32_t)link;
}
__device__ __forceinline__ float entryT(StackEntry e) { return __uint_as_float((uint32_t)(e >> 32)); }
__device__ __forceinline__ int32_t entryLink(StackEntry e) { return (int32_t)(uint32_t)e; }

struct TraceCounters
{
    uint32_t nodes, boxes, leaves, tris;
};

// One visit of a 4-wide node: culling tests of the (up to) four children, the passing ones sorted by entry distance
// (5-comparator network; misses sort to the end).  On return nh = number of passing children, (k0,l0) the nearest,
// (k1,l1) .. (k3,l3) the others near-to-far.  Shared by the plain and the persistent walk.
struct NodeOrder
{
    float k0, k1, k2, k3;
    int32_t l0, l1, l2, l3;
    int nh;
};
__device__ __forceinline__ NodeOrder visitWideNode(const WideNode *node, float3 inv, float3 nsi, float best_t)
{
    const NodeRegs nr = loadWideNode(node);
    const float4 lox = nr.lox, loy = nr.loy, loz = nr.loz, hix = nr.hix, hiy = nr.hiy, hiz = nr.hiz;
    const int4 lk = nr.lk;
    float t0, t1, t2, t3;
    const bool h0 = childCull(inv, nsi, lox.x, loy.x, loz.x, hix.x, hiy.x, hiz.x, best_t, t0);
    const bool h1 = childCull(inv, nsi, lox.y, loy.y, loz.y, hix.y, hiy.y, hiz.y, best_t, t1);
    const bool h2 = (lk.z != TRT_LINK_EMPTY) && childCull(inv, nsi, lox.z, loy.z, loz.z, hix.z, hiy.z, hiz.z, best_t, t2);
    const bool h3 = (lk.w != TRT_LINK_EMPTY) && childCull(inv, nsi, lox.w, loy.w, loz.w, hix.w, hiy.w, hiz.w, best_t, t3);
    NodeOrder o;
    o.k0 = h0 ? t0 : 3.0e38f, o.k1 = h1 ? t1 : 3.0e38f, o.k2 = h2 ? t2 : 3.0e38f, o.k3 = h3 ? t3 : 3.0e38f;
    o.l0 = lk.x, o.l1 = lk.y, o.l2 = lk.z, o.l3 = lk.w;
#define TRT_CSWAP(ka, la, kb, lb)                                                                                   \
    {                                                                                                               \
        const bool sw = kb < ka;                                                                                    \
        const float tk = sw ? kb : ka;                                                                              \
        const int32_t tl = sw ? lb : la;                                                                            \
        kb = sw ? ka : kb, lb = sw ? la : lb, ka = tk, la = tl;                                                     \
    }
    TRT_CSWAP(o.k0, o.l0, o.k1, o.l1)
    TRT_CSWAP(o.k2, o.l2, o.k3, o.l3)
    TRT_CSWAP(o.k0, o.l0, o.k2, o.l2)
    TRT_CSWAP(o.k1, o.l1, o.k3, o.l3)
    TRT_CSWAP(o.k1, o.l1, o.k2, o.l2)
#undef TRT_CSWAP
    o.nh = (int)h0 + (int)h1 + (int)h2 + (int)h3;
    return o;
}

// 4-wide walk of the fast layout: while-while (inner nodes until a leaf is reached, then the leaf scan),
// nearest child first, entry-distance pruning at push and at pop.  cls = rayClass (0 or 1).
template <bool STATS>
__device__ __forceinline__ void traceWide(const SceneView &sv, float3 S, float3 d, Hit &hit, TraceCounters *cnt, int cls = 0)
{
    hit.t = TRT_INF, hit.id = -1, hit.key = 0xFFFFFFFFu;
    if (sv.wide_root == TRT_LINK_EMPTY)
        return;
    const float3 inv = rcpDir(d);
    const float3 cinv = (cls == 1) ? cullInv(sv, inv) : inv;
    const float3 nsi = f3(-(S.x * cinv.x), -(S.y * cinv.y), -(S.z * cinv.z));
    StackEntry stack[TRT_WIDE_STACK];
    stack[0] = packEntry(-1.f, TRT_LINK_EXIT);
    int sp = 1;
    int32_t cur = sv.wide_root;
    for (;;)
    {
        while (cur >= 0)
        {
            const NodeOrder o = visitWideNode(sv.wide_nodes + cur, cinv, nsi, hit.t);
            if (STATS)
            {
                const int4 lk = __ldg(&sv.wide_nodes[cur].link);
                cnt->nodes++, cnt->boxes += 2 + (lk.z != TRT_LINK_EMPTY) + (lk.w != TRT_LINK_EMPTY);
            }
            if (o.nh == 0)
            {
                StackEntry e;
                do
                {
                    e = stack[--sp];
                } while (entryT(e) > hit.t);
                cur = entryLink(e);
                continue;
            }
            if (o.nh > 3)
                stack[sp++] = packEntry(o.k3, o.l3);
            if (o.nh > 2)
                stack[sp++] = packEntry(o.k2, o.l2);
            if (o.nh > 1)
                stack[sp++] = packEntry(o.k1, o.l1);
            cur = o.l0;
        }
        if (cur == TRT_LINK_EXIT)
        {
            hit.id = (hit.id >= 0) ? __ldg(sv.fast_orig + hit.id) : -1; // fast index -> post-build index
            return;
        }
        const int leaf = ~cur;
        if (STATS)
            cnt->leaves++, cnt->tris += (leaf & 7) + 1;
        scanLeaf<true>(sv, leaf >> 3, (leaf & 7) + 1, S, d, inv, hit, cls == 1);
        StackEntry e;
        do
        {
            e = stack[--sp];
        } while (entryT(e) > hit.t);
        cur = entryLink(e);
    }
}

// ---- warp-persistent walker ---------------------------------------------------------------------------------
// Thread-per-ray over the 4-wide layout, organised for SIMT efficiency on incoherent rays (ncu on the plain
// while-while kernel: 7.7 of 32 lanes active, half of the loss being lanes whose ray had finished):
//   * lanes fetch a new ray from a global counter as soon as enough lanes of the warp are idle (no tail);
//   * the inner-node loop and the leaf loop are warp-uniform: a lane that reaches a leaf postpones it and keeps
//     walking inner nodes until no lane of the warp still lacks a leaf, then the warp scans leaves together.
//     (Putting the next phase to a vote each step — the one with more lanes ready wins — was measured in round 2:
//     two more ballots per step cost more than the stragglers it removes, staircase 6.44 -> 5.96 Grays/s.)
// Results are identical to traceWide: only the order in which a ray's own nodes are visited changes, and the
// (t, key) order decides the winner (see the header comment).
// ncu on the render walk (profiles/r02_walk_staircase.txt): 52 % of the stall samples are waits on loads, and the three
// pop loops alone — a local-memory load, a compare, a branch, at 2-4 lanes while the rest of the warp waits — hold
// 20 % of all samples.  Two remedies were built and measured in round 2, and both lost: mirroring the
// top stack entry in registers so that a pop only STARTS the load of the entry below (staircase 6.38 -> 6.18 Grays/s,
// shadow walk 58.2 -> 61.2 ms: two more live registers at the 56-register cap), and handing a nearest child that is a leaf
// straight to the postponed-leaf slot instead of pushing and popping it back (6.38 -> 6.22: the six selects that shift
// the sorted children run for every lane, the saved store + load only for a fifth of the visits).  Folding the two pop
// sites of a node step (no child passed / a leaf was reached and is postponed) into one block executed once per step:
// no change (staircase render 82.6 -> 82.4 ms, veach-mis closest hit 14.1 -> 13.8 Grays/s).
struct WalkState
{
    float3 S, d, inv, nsi; // inv: 1/d (the culling reciprocal of a class-1 ray, see cullInv); nsi = -(S * inv)
    Hit hit;
    int32_t cur, leaf; // cur: inner node / leaf link / TRT_LINK_EXIT; leaf: postponed leaf link or TRT_LINK_EMPTY
    int sp;
};

#define TRT_WALK_PUSH(st, stack, e) (stack[st.sp++] = (e))
#define TRT_WALK_POP(st, stack)                                                                                      \
    do                                                                                                              \
    {                                                                                                               \
        StackEntry e__;                                                                                             \
        do                                                                                                          \
        {                                                                                                           \
            e__ = stack[--st.sp];                                                                                   \
        } while (entryT(e__) > st.hit.t);                                                                           \
        st.cur = entryLink(e__);                                                                                    \
    } while (0)
#define TRT_WALK_RESET(st, stack)                                                                                    \
    do                                                                                                              \
    {                                                                                                               \
        stack[0] = packEntry(-1.f, TRT_LINK_EXIT);                                                                  \
        st.sp = 1;                                                                                                  \
    } while (0)

// One inner-node step of lane state `st` (st.cur >= 0 on entry).
__device__ __forceinline__ void walkNodeStep(const SceneView &sv, WalkState &st, StackEntry *stack)
{
    const NodeOrder o = visitWideNode(sv.wide_nodes + st.cur, st.inv, st.nsi, st.hit.t);
    if (o.nh == 0)
    {
        TRT_WALK_POP(st, stack);
        return;
    }
    if (o.nh > 3)
        TRT_WALK_PUSH(st, stack, packEntry(o.k3, o.l3));
    if (o.nh > 2)
        TRT_WALK_PUSH(st, stack, packEntry(o.k2, o.l2));
    if (o.nh > 1)
        TRT_WALK_PUSH(st, stack, packEntry(o.k1, o.l1));
    st.cur = o.l0;
}

#define TRT_WALK_WARPS 4 // warps per CTA of every kernel that calls walkPersistent (128 threads)
#ifndef TRT_WALK_REFILL
#define TRT_WALK_REFILL 8 // refill once this many lanes of the warp are idle
#endif

// Inside tests of n (<= 32) pooled candidates, one per lane; the ray comes from the owner lane's registers.
__device__ __forceinline__ void walkPoolInside(const SceneView &sv, const WalkState &st, const int32_t *pool_tri,
                                               const float *pool_t, const int32_t *pool_owner, int n,
                                               unsigned long long *best, int lane)
{
    constexpr unsigned FULL = 0xffffffffu;
    const bool valid = lane < n;
    const int tri = valid ? pool_tri[lane] : 0;
    const float t = valid ? pool_t[lane] : 0.f;
    const int owner = valid ? pool_owner[lane] : lane;
    const float3 S = f3(__shfl_sync(FULL, st.S.x, owner), __shfl_sync(FULL, st.S.y, owner), __shfl_sync(FULL, st.S.z, owner));
    const float3 d = f3(__shfl_sync(FULL, st.d.x, owner), __shfl_sync(FULL, st.d.y, owner), __shfl_sync(FULL, st.d.z, owner));
    if (valid)
    {
        const float4 *gp = reinterpret_cast<const float4 *>(sv.fast_geom + tri);
        bool ok = triangleInside(__ldg(gp), __ldg(gp + 1), __ldg(gp + 2), S, d, t);
        if (ok && sv.check_leaf_box)
        {
            const float4 *bp = sv.ref_leaf_box + 2 * __ldg(sv.fast_leaf + tri);
            const float4 lo = __ldg(bp), hi = __ldg(bp + 1);
            float t0;
            ok = childPass(S, rcpDir(d), lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, t0);
        }
        if (ok)
            atomicMin(best + owner, ((unsigned long long)__float_as_uint(t) << 32) | __ldg(sv.fast_rank + tri));
    }
}

// RAYS: struct with  unsigned locate(unsigned i)  (ray index -> the 32-bit token the walker keeps while the ray is in
// flight, below 2^31; below 2^29 apart from RAYS's own flag bit 30 when kHasBound),  void load(token, float3 &S,
// float3 &d, float &tmax)  (tmax: initial search bound, TRT_INF for none),  bool bounded(token),  void store(token,
// const Hit &)  (hit.id = post-build triangle index)  and  void storeFast(sv, token, const Hit &)  (hit.id = index in
// the fast layout's own order).  A bounded ray that finds nothing within its bound is walked again without one, so the
// result is the unbounded search's in every case (kHasBound = false compiles the second attempt out).
// `counter` must be zero at launch; n = number of rays.  Every thread of the block must call this.
// POOLED selects the leaf phase: false = every lane scans its own leaf (scanLeaf); true = the warp pools the
// candidates that pass the plane test and deals their inside tests out one per lane (see below; measured in
// profiles/r01_closest_staircase_leaflayout_pooled.txt: 8 % fewer instructions at 15 instead of 9.5 lanes, but the shorter
// dependent chains expose L1 latency — 69 % instead of 83 % issue-active — and it is 9 % slower, so it is off
// by default and kept selectable (TRT_TRACE_POOLED)).
template <bool POOLED, typename RAYS>
__device__ __forceinline__ void walkPersistent(const SceneView &sv, RAYS &rays, unsigned int n, unsigned int *counter)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int kRefillIdle = TRT_WALK_REFILL;
    constexpr unsigned kPathGateBit = 0x80000000u, kRetryBit = 0x20000000u;
    constexpr unsigned kOwnBits = kPathGateBit | (RAYS::kHasBound ? kRetryBit : 0u);
    const int lane = threadIdx.x & 31;
    // per warp: pool of (triangle, t, owner lane) candidates awaiting their inside test, and each lane's best
    // (t bits << 32 | rank) so far
    constexpr int kPoolWarps = POOLED ? TRT_WALK_WARPS : 1, kPoolSlots = POOLED ? 64 : 1, kBestSlots = POOLED ? 32 : 1;
    __shared__ int32_t s_pool_tri[kPoolWarps][kPoolSlots];
    __shared__ float s_pool_t[kPoolWarps][kPoolSlots];
    __shared__ int32_t s_pool_owner[kPoolWarps][kPoolSlots];
    __shared__ unsigned long long s_best[kPoolWarps][kBestSlots];
    const int warp = POOLED ? (threadIdx.x >> 5) : 0;
    int32_t *pool_tri = s_pool_tri[warp], *pool_owner = s_pool_owner[warp];
    float *pool_t = s_pool_t[warp];
    unsigned long long *best = s_best[warp];
    StackEntry stack[TRT_WIDE_STACK];
    WalkState st;
    st.cur = TRT_LINK_EXIT, st.leaf = TRT_LINK_EMPTY, st.sp = 1;
    unsigned int ray = 0xffffffffu; // idle
    bool drained = false;           // the pool is empty
    for (;;)
    {
        // ---- refill
        const unsigned idle = __ballot_sync(FULL, ray == 0xffffffffu);
        if (idle == FULL && drained)
            return;
        if (!drained && (__popc(idle) >= kRefillIdle))
        {
            unsigned int base = 0;
            const int leader = __ffs(idle) - 1;
            if (lane == leader)
                base = atomicAdd(counter, (unsigned int)__popc(idle));
            base = __shfl_sync(FULL, base, leader);
            if (ray == 0xffffffffu)
            {
                const unsigned int mine = base + (unsigned int)__popc(idle & ((1u << lane) - 1u));
                if (mine < n)
                {
                    ray = rays.locate(mine);
                    rays.load(ray, st.S, st.d, st.hit.t);
                    st.hit.id = -1, st.hit.key = 0xFFFFFFFFu;
                    const int cls = sv.use_wide ? rayClass(sv, st.S, st.d) : 3;
                    if (cls >= 2 || (POOLED && cls == 1)) // (the pooled inside tests have only the leaf-box gate)
                    {
                        // rare: the reference's own walk (it sets its own, unbounded, start state), finished on the spot
                        if (cls != 3)
                        {
                            atomicAdd(sv.strict_counter, 1ull);
                            traceRefTopology<true>(sv, st.S, st.d, st.hit);
                        }
                        else
                            traceRefTopology<false>(sv, st.S, st.d, st.hit);
                        rays.store(ray, st.hit);
                        ray = 0xffffffffu;
                    }
                    else
                    {
                        if (POOLED)
                            best[lane] = ((unsigned long long)__float_as_uint(TRT_INF) << 32) | sv.miss_rank;
                        st.inv = rcpDir(st.d);
                        if (cls == 1)
                        {
                            ray |= kPathGateBit; // tokens stay below 2^31: the top bit marks a class-1 ray (rayClass)
                            st.inv = cullInv(sv, st.inv);
                        }
                        st.nsi = f3(-(st.S.x * st.inv.x), -(st.S.y * st.inv.y), -(st.S.z * st.inv.z));
                        TRT_WALK_RESET(st, stack);
                        st.cur = sv.wide_root;
                        st.leaf = TRT_LINK_EMPTY;
                        if (st.cur == TRT_LINK_EMPTY) // empty scene
                            st.cur = TRT_LINK_EXIT;
                        else if (st.cur < 0) // the whole scene is one reference leaf: scanned without a box test
                            st.leaf = st.cur, st.cur = TRT_LINK_EXIT;
                    }
                }
            }
            if (base + (unsigned int)__popc(idle) >= n)
                drained = true;
        }
        const bool live = ray != 0xffffffffu;
        // ---- inner nodes, until no live lane still lacks a leaf
        for (;;)
        {
            const bool want = live && st.cur >= 0;
            if (!__any_sync(FULL, want && st.leaf == TRT_LINK_EMPTY))
                break;
            if (want)
            {
                walkNodeStep(sv, st, stack);
                if (st.cur < 0 && st.cur != TRT_LINK_EXIT && st.leaf == TRT_LINK_EMPTY)
                {
                    st.leaf = st.cur; // postpone the leaf, keep walking
                    TRT_WALK_POP(st, stack);
                }
            }
        }
        // ---- leaves: every lane runs the plane part for the triangles of ITS leaf; the candidates of the whole warp
        // are pooled in shared memory and their inside tests (two thirds of a triangle test) are dealt out 32 at a
        // time, one per lane, whoever owns the ray — ncu on the per-lane scan: 65 % of issued instructions were
        // triangle tests at 7-10 of 32 lanes.  Results merge through a 64-bit atomicMin on (t bits, rank).
        while (__any_sync(FULL, live && st.leaf != TRT_LINK_EMPTY))
        {
            const bool has = live && st.leaf != TRT_LINK_EMPTY;
            if (!POOLED)
            {
                if (has)
                {
                    const int lf = ~st.leaf;
                    const int32_t before = st.hit.id;
                    scanLeaf<true>(sv, lf >> 3, (lf & 7) + 1, st.S, st.d, st.inv, st.hit, (ray & kPathGateBit) != 0);
                    if (RAYS::kHasBound && st.hit.id != before && rays.canStop(sv, ray & ~kOwnBits, st.hit, st.inv, st.nsi))
                        st.cur = TRT_LINK_EXIT, st.leaf = TRT_LINK_EMPTY; // the answer is known (see WalkRays::canStop)
                    else if (st.cur < 0 && st.cur != TRT_LINK_EXIT)
                    {
                        st.leaf = st.cur; // a second leaf was reached while the first was postponed
                        TRT_WALK_POP(st, stack);
                    }
                    else
                        st.leaf = TRT_LINK_EMPTY;
                }
                continue;
            }
            const int leaf = has ? ~st.leaf : 0;
            const int first = leaf >> 3, num = has ? (leaf & 7) + 1 : 0;
            const int maxnum = __reduce_max_sync(FULL, num);
            int cnt = 0; // candidates waiting in the pool (warp-uniform)
            for (int k = 0; k < maxnum; ++k)
            {
                bool cand = false;
                float t = 0.f;
                if (k < num)
                {
                    const float4 *gp = reinterpret_cast<const float4 *>(sv.fast_geom + first + k);
                    cand = trianglePlane(__ldg(gp), __ldg(gp + 1), st.S, st.d, st.hit.t, t);
                }
                const unsigned m = __ballot_sync(FULL, cand);
                if (m == 0)
                    continue;
                if (cand)
                {
                    const int pos = cnt + __popc(m & ((1u << lane) - 1u));
                    pool_tri[pos] = first + k, pool_t[pos] = t, pool_owner[pos] = lane;
                }
                cnt += __popc(m);
                if (cnt >= 32)
                {
                    __syncwarp();
                    walkPoolInside(sv, st, pool_tri + (cnt - 32), pool_t + (cnt - 32), pool_owner + (cnt - 32), 32, best, lane);
                    cnt -= 32;
                    __syncwarp();
                }
            }
            if (cnt > 0)
            {
                __syncwarp();
                walkPoolInside(sv, st, pool_tri, pool_t, pool_owner, cnt, best, lane);
            }
            __syncwarp();
            st.hit.t = __uint_as_float((unsigned int)(best[lane] >> 32)); // tighter pruning bound for what follows
            if (has)
            {
                if (st.cur < 0 && st.cur != TRT_LINK_EXIT)
                {
                    st.leaf = st.cur; // a second leaf was reached while the first was postponed
                    TRT_WALK_POP(st, stack);
                }
                else
                    st.leaf = TRT_LINK_EMPTY;
            }
        }
        // ---- finished rays
        if (live && st.cur == TRT_LINK_EXIT && st.leaf == TRT_LINK_EMPTY)
        {
            if (RAYS::kHasBound && !POOLED && st.hit.id < 0 && !(ray & kRetryBit) && rays.bounded(ray & ~kOwnBits))
            {
                // nothing within the bound: the same ray once more, unbounded
                ray |= kRetryBit;
                st.hit.t = TRT_INF, st.hit.key = 0xFFFFFFFFu;
                TRT_WALK_RESET(st, stack);
                st.cur = sv.wide_root;
                if (st.cur < 0) // (wide_root is never EMPTY here: the ray was admitted to the fast layout)
                    st.leaf = st.cur, st.cur = TRT_LINK_EXIT;
            }
            else
            {
                if (POOLED)
                {
                    const unsigned long long b = best[lane];
                    st.hit.t = __uint_as_float((unsigned int)(b >> 32));
                    st.hit.id = __ldg(sv.rank_tri + (unsigned int)b);
                    rays.store(ray & ~kOwnBits, st.hit);
                }
                else
                    rays.storeFast(sv, ray & ~kOwnBits, st.hit);
                ray = 0xffffffffu;
            }
        }
    }
}

// Closest hit with the default (fast) layout; falls back to the reference topology for strict rays and for
// scenes without a wide layout.
__device__ __forceinline__ void traceClosest(const SceneView &sv, float3 S, float3 d, Hit &hit)
{
    const int cls = sv.use_wide ? rayClass(sv, S, d) : 3;
    if (cls == 3)
        traceRefTopology<false>(sv, S, d, hit);
    else if (cls == 2)
    {
        atomicAdd(sv.strict_counter, 1ull);
        traceRefTopology<true>(sv, S, d, hit);
    }
    else
        traceWide<false>(sv, S, d, hit, nullptr, cls);
}
} // namespace trt
