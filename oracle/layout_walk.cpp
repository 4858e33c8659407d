/* layout_walk.cpp — TEST INFRASTRUCTURE ONLY (part of liboracle.so, never linked or loaded by the product).
 *
 * Walks the product's fast layout (trt_layout_view from include/trt.h, built on the host by trt_layout_build) on the CPU
 * with the traversal RULES of DESIGN.md §3, restated here independently of the CUDA code:
 *   - 4-wide nodes, children in entry-distance order, a subtree skipped when its entry distance exceeds the best t;
 *   - a candidate triangle must pass the reference's plane + three-edge test (bvh.cpp:177-209, un-fused float, glm's
 *     operation order) AND its reference leaf's box must pass the reference's slab test (bvh.cpp:231-245); for rays
 *     with a zero / denormal direction component every box on the reference's root-to-leaf path is tested instead;
 *   - ties at equal t are decided by the SURVEY A.4 key.
 * Purpose: (1) `-m "not gpu"` tests compare its ids / distances with the oracle's exhaustive reference walk, which
 * checks the layout's DATA (boxes, leaf tags, keys, parent links) without a GPU; (2) work counters per ray class for
 * tuning the builder offline.  Compiled with -ffp-contract=off like the rest of the oracle. */
#include "../include/trt.h"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <omp.h>

namespace
{
struct V3
{
    float x, y, z;
};
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; } /* glm: products, then left-to-right adds */
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float gmin(float x, float y) { return (y < x) ? y : x; } /* glm::min / max: NaN in y returns x */
inline float gmax(float x, float y) { return (x < y) ? y : x; }

/* interactAABB with the reference's semantics (result > 0 <=> descend); t0 = entry distance */
inline bool refBoxPass(V3 S, V3 inv, const float *lo, const float *hi, float &t0)
{
    const float inx = (hi[0] - S.x) * inv.x, iny = (hi[1] - S.y) * inv.y, inz = (hi[2] - S.z) * inv.z;
    const float outx = (lo[0] - S.x) * inv.x, outy = (lo[1] - S.y) * inv.y, outz = (lo[2] - S.z) * inv.z;
    const float t1 = gmin(gmax(inx, outx), gmin(gmax(iny, outy), gmax(inz, outz)));
    t0 = gmax(gmin(inx, outx), gmax(gmin(iny, outy), gmin(inz, outz)));
    return (t1 >= t0) && (((t0 > 0.0f) ? t0 : t1) > 0.0f);
}
/* the culling test of the layout's own boxes: IEEE fmin / fmax (a NaN operand is dropped) */
inline bool cullBoxPass(V3 S, V3 inv, const float *lo, const float *hi, float &t0)
{
    const float inx = (hi[0] - S.x) * inv.x, iny = (hi[1] - S.y) * inv.y, inz = (hi[2] - S.z) * inv.z;
    const float outx = (lo[0] - S.x) * inv.x, outy = (lo[1] - S.y) * inv.y, outz = (lo[2] - S.z) * inv.z;
    const float t1 = std::fmin(std::fmax(inx, outx), std::fmin(std::fmax(iny, outy), std::fmax(inz, outz)));
    t0 = std::fmax(std::fmin(inx, outx), std::fmax(std::fmin(iny, outy), std::fmin(inz, outz)));
    return (t1 >= t0) && (((t0 > 0.0f) ? t0 : t1) > 0.0f);
}

struct Best
{
    float t;
    int32_t id; /* fast index */
};

inline uint32_t keyOf(const trt_layout_view &L, int32_t fast_id) { return fast_id < 0 ? L.miss_key : L.fast_key[fast_id]; }

/* every box on the reference's path from the root to `leaf` */
bool refPathPasses(const trt_layout_view &L, int leaf, V3 S, V3 inv)
{
    int32_t p = L.ref_leaf_parent[leaf];
    while (p >= 0)
    {
        const float *n = L.ref_nodes + (size_t)(p >> 1) * 16;
        float t0;
        const float *lo = (p & 1) ? n + 6 : n, *hi = lo + 3;
        if (!refBoxPass(S, inv, lo, hi, t0))
            return false;
        int32_t parent;
        std::memcpy(&parent, n + 14, 4);
        p = parent;
    }
    return true;
}

void scanLeaf(const trt_layout_view &L, int first, int num, V3 S, V3 d, V3 inv, bool pathGate, Best &best, uint64_t *cnt)
{
    for (int i = first; i < first + num; ++i)
    {
        cnt[3]++;
        const float *g = L.fast_geom + (size_t)i * 12;
        const V3 N = {g[0], g[1], g[2]}, p1 = {g[3], g[4], g[5]}, p2 = {g[6], g[7], g[8]}, p3 = {g[9], g[10], g[11]};
        const float dn = dot(N, d);
        if (std::fabs(dn) < 0.00001f)
            continue;
        const float t = dot(sub(p1, S), N) / dn;
        if (t < 0.0005f || t > best.t)
            continue;
        const V3 P = {S.x + d.x * t, S.y + d.y * t, S.z + d.z * t};
        const float dir1 = dot(cross(sub(p2, p1), sub(P, p1)), N);
        const float dir2 = dot(cross(sub(p3, p2), sub(P, p2)), N);
        const float dir3 = dot(cross(sub(p1, p3), sub(P, p3)), N);
        if (!((dir1 > 0.f && dir2 > 0.f && dir3 > 0.f) || (dir1 < 0.f && dir2 < 0.f && dir3 < 0.f)))
            continue;
        const int leaf = L.fast_leaf[i];
        if (pathGate)
        {
            if (!refPathPasses(L, leaf, S, inv))
                continue;
        }
        else if (L.check_leaf_box)
        {
            float t0;
            const float *b = L.ref_leaf_box + (size_t)leaf * 8;
            if (!cullBoxPass(S, inv, b, b + 4, t0)) /* finite, non-NaN operands here: equals the reference's test */
                continue;
        }
        if (t < best.t || (t == best.t && L.fast_key[i] > keyOf(L, best.id)))
            best.t = t, best.id = i;
    }
}
} // namespace

extern "C"
{
/* id: post-build triangle index, -1 miss, -2 = ray class the fast layout does not serve (non-finite, or origin
 * beyond strict_origin_limit: the product walks the reference topology exhaustively for those); t: distance,
 * TRT_INF on miss.  counters[4] += {wide nodes visited, child boxes tested, leaves scanned, triangles tested}. */
void orc_walk_layout(const trt_layout_view *Lp, const float *rays6, int64_t n, int32_t *id_out, float *t_out,
                     uint64_t *counters, int32_t threads)
{
    const trt_layout_view &L = *Lp;
    if (threads > 0)
        omp_set_num_threads(threads);
    uint64_t total[4] = {0, 0, 0, 0};
#pragma omp parallel
    {
        uint64_t cnt[4] = {0, 0, 0, 0};
#pragma omp for schedule(dynamic, 1024)
        for (int64_t r = 0; r < n; ++r)
        {
            const float *q = rays6 + r * 6;
            const V3 S = {q[0], q[1], q[2]}, d = {q[3], q[4], q[5]};
            const V3 inv = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
            const float sum = ((S.x + S.y) + S.z) + ((d.x + d.y) + d.z);
            const bool far = std::fmax(std::fabs(S.x), std::fmax(std::fabs(S.y), std::fabs(S.z))) > L.strict_origin_limit;
            if (far || !(std::fabs(sum) < 3.0e38f) || L.wide_root == 0x7fffffff)
            {
                id_out[r] = -2, t_out[r] = TRT_INF;
                continue;
            }
            const bool pathGate = !(std::fabs(inv.x) <= 3.4028235e38f) || !(std::fabs(inv.y) <= 3.4028235e38f) ||
                                  !(std::fabs(inv.z) <= 3.4028235e38f);
            Best best = {TRT_INF, -1};
            struct Entry
            {
                float t;
                int32_t link;
            } stack[128];
            int sp = 0;
            int32_t cur = L.wide_root;
            for (;;)
            {
                if (cur >= 0)
                {
                    cnt[0]++;
                    const float *nd = L.wide_nodes + (size_t)cur * 32;
                    int32_t link[4];
                    std::memcpy(link, nd + 24, 16);
                    Entry hit[4];
                    int nh = 0;
                    for (int k = 0; k < 4; ++k)
                    {
                        if (link[k] == 0x7fffffff)
                            continue;
                        cnt[1]++;
                        const float lo[3] = {nd[k], nd[4 + k], nd[8 + k]}, hi[3] = {nd[12 + k], nd[16 + k], nd[20 + k]};
                        float t0;
                        if (cullBoxPass(S, inv, lo, hi, t0) && !(t0 > best.t))
                            hit[nh++] = {t0, link[k]};
                    }
                    for (int a = 1; a < nh; ++a) /* insertion sort by entry distance, stable */
                        for (int b = a; b > 0 && hit[b].t < hit[b - 1].t; --b)
                        {
                            const Entry e = hit[b];
                            hit[b] = hit[b - 1], hit[b - 1] = e;
                        }
                    for (int k = nh - 1; k >= 1; --k)
                        stack[sp++] = hit[k];
                    if (nh > 0)
                    {
                        cur = hit[0].link;
                        continue;
                    }
                }
                else
                {
                    const int leaf = ~cur;
                    cnt[2]++;
                    scanLeaf(L, leaf >> 3, (leaf & 7) + 1, S, d, inv, pathGate, best, cnt);
                }
                bool found = false;
                while (sp > 0)
                {
                    const Entry e = stack[--sp];
                    if (!(e.t > best.t))
                    {
                        cur = e.link, found = true;
                        break;
                    }
                }
                if (!found)
                    break;
            }
            id_out[r] = best.id >= 0 ? L.fast_orig[best.id] : -1;
            t_out[r] = best.t;
        }
#pragma omp critical
        for (int k = 0; k < 4; ++k)
            total[k] += cnt[k];
    }
    if (counters)
        for (int k = 0; k < 4; ++k)
            counters[k] += total[k];
}
}
