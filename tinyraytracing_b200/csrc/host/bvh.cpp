// buildBVH on the host: the reference's topology and in-place triangle reorder (bvh.cpp:16-144), computed on
// lean 16-byte sort records instead of 168-byte Triangles with std::string members.
//
// Why the result is identical to the reference built with the same standard library: every step of the
// reference touches `triangles[l..r]` only through (a) std::sort with a comparator on one centroid
// component and (b) reads of vertex extents.  std::sort's element moves depend only on comparison outcomes
// and positions, so sorting records {centroid, source index} with the same comparators in the same sequence
// (x, y, z, then the winning axis again: the sorts are unstable, each starts from the previous one's order)
// yields the same permutation; the fat triangles are permuted once at the end.
#include "tinyrt.h"

#include <algorithm>
#include <deque>
#include <future>
#include <mutex>

namespace trt
{
namespace
{
struct Rec
{
    float c[3];
    int src;
};
struct Ext
{
    float lo[3], hi[3];
};
struct Arena
{
    // stable addresses; one segment per build task so that concurrent subtree builds never share a container
    std::mutex mutex;
    std::vector<std::unique_ptr<std::deque<BVHNode>>> segments;
    std::deque<BVHNode> *newSegment()
    {
        std::lock_guard<std::mutex> g(mutex);
        segments.emplace_back(new std::deque<BVHNode>());
        return segments.back().get();
    }
};
std::mutex g_arenaMutex;
std::unordered_map<const BVHNode *, std::unique_ptr<Arena>> g_arenas;

// Subtrees over disjoint ranges are independent once the parent's final sort has run (every later step of the
// reference only touches its own range), so the upper levels fork one task per left child.  The sorts themselves
// stay std::sort on one thread each: their permutation is part of the contract.
constexpr int kForkDepth = 6;     // at most 2^6 concurrent subtree builds
constexpr int kForkMinRange = 1 << 14;

struct Builder
{
    std::vector<Rec> &rec;
    const std::vector<Ext> &ext; // per source triangle: min / max over its three vertices
    int leaf_num;
    Arena &arena;
    std::deque<BVHNode> *segment;
    std::vector<float> pre, suf; // prefix / suffix extents, 6 floats per slot, reused across nodes

    BVHNode *build(int l, int r, int depth = 0)
    {
        if (l > r)
            return nullptr;
        segment->emplace_back();
        BVHNode *node = &segment->back();
        float AA[3] = {1145141919.f, 1145141919.f, 1145141919.f}, BB[3] = {-1145141919.f, -1145141919.f, -1145141919.f};
        for (int i = l; i <= r; ++i)
        {
            const Ext &e = ext[rec[i].src];
            for (int a = 0; a < 3; ++a)
            {
                AA[a] = fmin2(AA[a], e.lo[a] - 0.001f);
                BB[a] = fmax2(BB[a], e.hi[a] + 0.001f);
            }
        }
        node->AA = vec3(AA[0], AA[1], AA[2]);
        node->BB = vec3(BB[0], BB[1], BB[2]);
        const int n = r - l + 1;
        if (n <= leaf_num)
        {
            node->num = n;
            node->index = l;
            return node;
        }
        float Cost = INF;
        int Axis = 0, Split = (l + r) / 2;
        pre.resize((size_t)n * 6);
        suf.resize((size_t)n * 6);
        for (int axis = 0; axis < 3; ++axis)
        {
            sortAxis(l, r, axis);
            // running extents start from +-INF (114514), not from the first triangle (bvh.cpp:62-63,79-80)
            for (int i = 0; i < n; ++i)
            {
                const Ext &e = ext[rec[l + i].src];
                for (int a = 0; a < 3; ++a)
                {
                    const float pmax = (i == 0) ? -INF : pre[(size_t)(i - 1) * 6 + 3 + a];
                    const float pmin = (i == 0) ? INF : pre[(size_t)(i - 1) * 6 + a];
                    pre[(size_t)i * 6 + 3 + a] = fmax2(pmax, e.hi[a]);
                    pre[(size_t)i * 6 + a] = fmin2(pmin, e.lo[a]);
                }
            }
            for (int i = n - 1; i >= 0; --i)
            {
                const Ext &e = ext[rec[l + i].src];
                for (int a = 0; a < 3; ++a)
                {
                    const float smax = (i == n - 1) ? -INF : suf[(size_t)(i + 1) * 6 + 3 + a];
                    const float smin = (i == n - 1) ? INF : suf[(size_t)(i + 1) * 6 + a];
                    suf[(size_t)i * 6 + 3 + a] = fmax2(smax, e.hi[a]);
                    suf[(size_t)i * 6 + a] = fmin2(smin, e.lo[a]);
                }
            }
            float cost = INF;
            int split = l;
            for (int i = 0; i < n - 1; ++i)
            {
                const float *L = &pre[(size_t)i * 6], *R = &suf[(size_t)(i + 1) * 6];
                float lx = L[3] - L[0], ly = L[4] - L[1], lz = L[5] - L[2];
                // `2.0 * (...)` is a double product rounded to float (bvh.cpp:106,114)
                float la = (float)(2.0 * (double)((lx * ly) + (lx * lz) + (ly * lz)));
                float lc = la * (float)(i + 1);
                float rx = R[3] - R[0], ry = R[4] - R[1], rz = R[5] - R[2];
                float ra = (float)(2.0 * (double)((rx * ry) + (rx * rz) + (ry * rz)));
                float rc = ra * (float)(n - 1 - i);
                float total = lc + rc;
                if (total < cost)
                {
                    cost = total;
                    split = l + i;
                }
            }
            if (cost < Cost)
            {
                Cost = cost;
                Axis = axis;
                Split = split;
            }
        }
        sortAxis(l, r, Axis);
        if (depth < kForkDepth && n >= kForkMinRange)
        {
            auto left = std::async(std::launch::async, [this, l, Split, depth] {
                Builder sub{rec, ext, leaf_num, arena, arena.newSegment(), {}, {}};
                return sub.build(l, Split, depth + 1);
            });
            node->right = build(Split + 1, r, depth + 1);
            node->left = left.get();
        }
        else
        {
            node->left = build(l, Split, depth + 1);
            node->right = build(Split + 1, r, depth + 1);
        }
        return node;
    }

    void sortAxis(int l, int r, int axis)
    {
        auto b = rec.begin() + l, e = rec.begin() + r + 1;
        if (axis == 0)
            std::sort(b, e, [](const Rec &p, const Rec &q) { return p.c[0] < q.c[0]; });
        else if (axis == 1)
            std::sort(b, e, [](const Rec &p, const Rec &q) { return p.c[1] < q.c[1]; });
        else
            std::sort(b, e, [](const Rec &p, const Rec &q) { return p.c[2] < q.c[2]; });
    }
};
} // namespace

BVHNode *buildBVH(std::vector<Triangle> &triangles, int l, int r, int leaf_num)
{
    if (l > r || l < 0 || r >= (int)triangles.size())
        return nullptr;
    std::vector<Rec> rec(triangles.size());
    std::vector<Ext> ext(triangles.size());
    for (size_t i = 0; i < triangles.size(); ++i)
    {
        const Triangle &t = triangles[i];
        rec[i] = {{t.center.x, t.center.y, t.center.z}, (int)i};
        const float vx[3] = {t.v[0].x, t.v[1].x, t.v[2].x}, vy[3] = {t.v[0].y, t.v[1].y, t.v[2].y},
                    vz[3] = {t.v[0].z, t.v[1].z, t.v[2].z};
        const float *c[3] = {vx, vy, vz};
        for (int a = 0; a < 3; ++a)
        {
            ext[i].lo[a] = fmin2(c[a][0], fmin2(c[a][1], c[a][2]));
            ext[i].hi[a] = fmax2(c[a][0], fmax2(c[a][1], c[a][2]));
        }
    }
    std::unique_ptr<Arena> arena(new Arena());
    Builder b{rec, ext, leaf_num, *arena, arena->newSegment(), {}, {}};
    BVHNode *root = b.build(l, r);
    // apply the permutation to the fat triangles once
    std::vector<Triangle> sorted;
    sorted.reserve((size_t)(r - l + 1));
    for (int i = l; i <= r; ++i)
        sorted.push_back(std::move(triangles[rec[i].src]));
    for (int i = l; i <= r; ++i)
        triangles[i] = std::move(sorted[i - l]);
    std::lock_guard<std::mutex> g(g_arenaMutex);
    g_arenas[root] = std::move(arena);
    return root;
}

void freeBVH(BVHNode *root)
{
    std::lock_guard<std::mutex> g(g_arenaMutex);
    g_arenas.erase(root);
}

int nodeCountBVH(const BVHNode *n) { return n ? 1 + nodeCountBVH(n->left) + nodeCountBVH(n->right) : 0; }
} // namespace trt
