// Host-side construction of the GPU layouts from the reference topology handed over in trt_scene_desc.
#include "accel.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <future>
#include <memory>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <queue>

namespace trt
{
std::string buildAccel(const trt_scene_desc &desc, AccelBuild &out)
{
    const int n = desc.n_tris, nn = desc.n_nodes;
    out = AccelBuild();
    out.tri_geom.resize(n);
    out.tri_key.assign(n, 0);
    for (int i = 0; i < n; ++i)
    {
        const float *v = desc.v + (size_t)i * 9, *N = desc.normal + (size_t)i * 3;
        out.tri_geom[i].q0 = make_float4(N[0], N[1], N[2], v[0]);
        out.tri_geom[i].q1 = make_float4(v[1], v[2], v[3], v[4]);
        out.tri_geom[i].q2 = make_float4(v[5], v[6], v[7], v[8]);
        if (desc.mtl[i] < 0 || desc.mtl[i] >= desc.n_materials)
            return "triangle material index out of range";
    }
    if (n == 0 || nn == 0)
        return (n == 0 && nn == 0) ? "" : "triangles without nodes (or nodes without triangles)";

    // pass 1: classify nodes, number leaves left to right (pre-order with left before right = ascending index)
    std::vector<int32_t> inner_index(nn, -1), leaf_ord(nn, -1);
    int n_inner = 0, n_leaves = 0;
    for (int i = 0; i < nn; ++i)
    {
        const int32_t *lk = desc.node_link + (size_t)i * 4;
        if (lk[3] > 0)
        {
            if (lk[3] > 8)
                return "leaf with more than 8 triangles (the GPU layout encodes num-1 in 3 bits; main.cpp:76 uses 8)";
            if (lk[2] < 0 || lk[2] + lk[3] > n)
                return "leaf triangle range out of bounds";
            leaf_ord[i] = n_leaves++;
        }
        else
        {
            if (lk[0] <= i || lk[0] >= nn || lk[1] <= i || lk[1] >= nn)
                return "inner node needs two children placed after it (pre-order)";
            inner_index[i] = n_inner++;
        }
    }
    if (n_leaves >= (1 << 27))
        return "too many leaves for the 32-bit tie key";
    out.n_leaves = n_leaves;

    // tie key: (emissive, emissive ? -leafOrdinal : +leafOrdinal, emissive ? +index : -index)
    for (int i = 0; i < nn; ++i)
    {
        const int32_t *lk = desc.node_link + (size_t)i * 4;
        if (lk[3] <= 0)
            continue;
        for (int k = 0; k < lk[3]; ++k)
        {
            const int tri = lk[2] + k;
            const bool em = desc.materials[desc.mtl[tri]].is_emissive != 0;
            out.tri_key[tri] = em ? (0x80000000u | ((uint32_t)(n_leaves - 1 - leaf_ord[i]) << 3) | (uint32_t)k)
                                  : (((uint32_t)leaf_ord[i] << 3) | (uint32_t)(7 - k));
        }
    }

    // ranks: position in descending key order with the miss key inserted
    {
        std::vector<int32_t> order(n);
        for (int i = 0; i < n; ++i)
            order[i] = i;
        std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return out.tri_key[a] > out.tri_key[b]; });
        out.tri_rank.assign(n, 0);
        out.rank_tri.assign((size_t)n + 1, -1);
        uint32_t r = 0;
        bool missPlaced = false;
        for (int i = 0; i < n; ++i)
        {
            if (!missPlaced && out.tri_key[order[i]] < TRT_MISS_KEY)
            {
                out.miss_rank = r++;
                missPlaced = true;
            }
            out.tri_rank[order[i]] = r;
            out.rank_tri[r] = order[i];
            ++r;
        }
        if (!missPlaced)
            out.miss_rank = r;
    }

    auto link = [&](int node) -> int32_t {
        const int32_t *lk = desc.node_link + (size_t)node * 4;
        return (lk[3] > 0) ? ~((lk[2] << 3) | (lk[3] - 1)) : inner_index[node];
    };
    if ((long long)n >= (1ll << 28))
        return "too many triangles for the leaf link encoding";
    out.ref_nodes.resize(n_inner);
    for (int i = 0; i < nn; ++i)
    {
        if (inner_index[i] < 0)
            continue;
        const int32_t *lk = desc.node_link + (size_t)i * 4;
        const float *L = desc.node_box + (size_t)lk[0] * 6, *R = desc.node_box + (size_t)lk[1] * 6;
        RefNode &o = out.ref_nodes[inner_index[i]];
        o.a = make_float4(L[0], L[1], L[2], L[3]);
        o.b = make_float4(L[4], L[5], R[0], R[1]);
        o.c = make_float4(R[2], R[3], R[4], R[5]);
        o.d = make_int4(link(lk[0]), link(lk[1]), -1, 0);
    }
    out.root_link = link(0);
    // parent links (the path of boxes the reference tests on its way to a leaf, walked upwards by refPathPasses)
    out.ref_leaf_parent.assign(n_leaves, -1);
    for (int i = 0; i < nn; ++i)
    {
        if (inner_index[i] < 0)
            continue;
        const int32_t *lk = desc.node_link + (size_t)i * 4;
        for (int side = 0; side < 2; ++side)
        {
            const int c = lk[side];
            const int32_t token = (inner_index[i] << 1) | side;
            if (inner_index[c] >= 0)
                out.ref_nodes[inner_index[c]].d.z = token;
            else
                out.ref_leaf_parent[leaf_ord[c]] = token;
        }
    }

    // depth of the reference tree (iterative: staircase reaches 59, degenerate inputs may go deeper)
    std::vector<int> depth(nn, 0);
    int maxd = 0;
    for (int i = 0; i < nn; ++i)
    {
        const int32_t *lk = desc.node_link + (size_t)i * 4;
        if (lk[3] > 0)
            continue;
        depth[lk[0]] = depth[lk[1]] = depth[i] + 1;
        maxd = depth[i] + 1 > maxd ? depth[i] + 1 : maxd;
    }
    out.ref_depth = maxd;
    return "";
}
} // namespace trt

// ------------------------------------------------------------------------------------------------------------
// Fast layout: binned-SAH binary tree collapsed to 4 children per node, built over
//   "triangles" (default)  the scene's triangles, own leaves of <= 2 triangles, boxes padded generously; a hit is
//                          accepted only if the triangle's REFERENCE leaf box also passes the reference's slab test
//   "leaves"               the reference's leaves as atomic scan units with their own boxes bit for bit
// (TRT_WIDE_SOURCE selects; "off" builds nothing).  Why not keep the reference's leaves: its SAH is capped at
// INF = 114514 (bvh.cpp:49-51,125-134), so in any scene with coordinates beyond a few tens of units EVERY split
// is a median split on x and the leaves are thin x-slabs spanning the whole object (Cornell shell + 100k-triangle
// sphere: median leaf extent 192 units, 60 leaves / 367 triangle tests per ray).
namespace trt
{
namespace
{
struct Prim
{
    float lo[3], hi[3], c[3];
    int32_t payload; // triangle index (triangle mode) or reference leaf ordinal (leaf mode)
    float w;
};
struct BNode
{
    float lo[3], hi[3];
    int32_t left = -1, right = -1; // children; for a leaf: right == -2 and [first, first + count) of `order`
    int32_t first = 0, count = 0;
};

inline float halfArea(const float *lo, const float *hi)
{
    const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
    return x * y + y * z + z * x;
}
inline void grow(float *lo, float *hi, const float *plo, const float *phi)
{
    for (int a = 0; a < 3; ++a)
    {
        lo[a] = plo[a] < lo[a] ? plo[a] : lo[a];
        hi[a] = phi[a] > hi[a] ? phi[a] : hi[a];
    }
}

struct WideBuilder
{
    std::vector<Prim> &prims;
    int maxLeaf;
    std::vector<int32_t> order;
    std::vector<BNode> bn;         // pre-sized to 2 * prims: slots are handed out by `next`, so that the upper
    std::atomic<int32_t> next{0};  // levels can build disjoint ranges on separate threads

    int32_t build(int l, int r, int depth = 0) // [l, r)
    {
        const int32_t me = next.fetch_add(1);
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int i = l; i < r; ++i)
        {
            const Prim &p = prims[order[i]];
            grow(lo, hi, p.lo, p.hi);
            grow(clo, chi, p.c, p.c);
        }
        for (int a = 0; a < 3; ++a)
            bn[me].lo[a] = lo[a], bn[me].hi[a] = hi[a];
        if (r - l <= maxLeaf)
        {
            bn[me].right = -2, bn[me].first = l, bn[me].count = r - l;
            return me;
        }
        constexpr int NB = 32;
        int bestAxis = -1, bestBin = -1;
        float bestCost = INFINITY;
        for (int a = 0; a < 3; ++a)
        {
            const float ext = chi[a] - clo[a];
            if (!(ext > 0.f))
                continue;
            float blo[NB][3], bhi[NB][3], bw[NB];
            for (int b = 0; b < NB; ++b)
            {
                bw[b] = 0.f;
                for (int k = 0; k < 3; ++k)
                    blo[b][k] = INFINITY, bhi[b][k] = -INFINITY;
            }
            const float scale = NB / ext;
            for (int i = l; i < r; ++i)
            {
                const Prim &p = prims[order[i]];
                int b = (int)((p.c[a] - clo[a]) * scale);
                b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                grow(blo[b], bhi[b], p.lo, p.hi);
                bw[b] += p.w;
            }
            float rlo[NB][3], rhi[NB][3], rw[NB];
            float alo[3] = {INFINITY, INFINITY, INFINITY}, ahi[3] = {-INFINITY, -INFINITY, -INFINITY}, aw = 0.f;
            for (int b = NB - 1; b > 0; --b)
            {
                grow(alo, ahi, blo[b], bhi[b]);
                aw += bw[b];
                for (int k = 0; k < 3; ++k)
                    rlo[b][k] = alo[k], rhi[b][k] = ahi[k];
                rw[b] = aw;
            }
            float llo[3] = {INFINITY, INFINITY, INFINITY}, lhi[3] = {-INFINITY, -INFINITY, -INFINITY}, lw = 0.f;
            for (int b = 0; b < NB - 1; ++b)
            {
                grow(llo, lhi, blo[b], bhi[b]);
                lw += bw[b];
                if (lw == 0.f || rw[b + 1] == 0.f)
                    continue;
                const float cost = halfArea(llo, lhi) * lw + halfArea(rlo[b + 1], rhi[b + 1]) * rw[b + 1];
                if (cost < bestCost)
                    bestCost = cost, bestAxis = a, bestBin = b;
            }
        }
        int mid;
        if (bestAxis < 0)
            mid = (l + r) / 2; // all centroids coincide
        else
        {
            const float ext = chi[bestAxis] - clo[bestAxis], scale = NB / ext, c0 = clo[bestAxis];
            const int a = bestAxis, bb = bestBin;
            auto it = std::partition(order.begin() + l, order.begin() + r, [&](int32_t pi) {
                int b = (int)((prims[pi].c[a] - c0) * scale);
                b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
                return b <= bb;
            });
            mid = (int)(it - order.begin());
            if (mid == l || mid == r)
                mid = (l + r) / 2;
        }
        int32_t L, R;
        if (depth < 6 && r - l >= (1 << 16))
        {
            auto left = std::async(std::launch::async, [this, l, mid, depth] { return build(l, mid, depth + 1); });
            R = build(mid, r, depth + 1);
            L = left.get();
        }
        else
        {
            L = build(l, mid, depth + 1);
            R = build(mid, r, depth + 1);
        }
        bn[me].left = L, bn[me].right = R;
        return me;
    }
};

// Insertion-based optimisation of the binary tree (after Bittner, Hapala, Havran: "Fast insertion-based optimization
// of bounding volume hierarchies", 2013): an inner node whose box is large for what it holds is taken out, and its two
// subtrees are re-inserted where they increase the summed surface area of the inner nodes least (branch-and-bound search
// from the root).  Leaves — the sets of triangles — are never changed and every box stays the exact union of the padded
// triangle boxes below it, so the layout's invariants (checkLayout) hold by construction; only the expected number of
// box tests per ray (the SAH sum) drops.  Deterministic: no randomness, fixed visiting order.
struct Reinserter
{
    std::vector<BNode> &bn;
    int32_t n_nodes;
    std::vector<int32_t> parent;
    size_t max_per_pass = 200000;
    std::vector<std::pair<float, int32_t>> heap_; // findTarget's queue, kept to avoid an allocation per search

    explicit Reinserter(std::vector<BNode> &b, int32_t n) : bn(b), n_nodes(n), parent(n, -1)
    {
        for (int32_t i = 0; i < n; ++i)
            if (bn[i].right != -2)
                parent[bn[i].left] = i, parent[bn[i].right] = i;
    }
    bool isLeaf(int32_t i) const { return bn[i].right == -2; }
    float area(int32_t i) const { return halfArea(bn[i].lo, bn[i].hi); }
    float unionArea(int32_t a, int32_t b) const
    {
        float lo[3], hi[3];
        for (int k = 0; k < 3; ++k)
            lo[k] = std::fmin(bn[a].lo[k], bn[b].lo[k]), hi[k] = std::fmax(bn[a].hi[k], bn[b].hi[k]);
        return halfArea(lo, hi);
    }
    // `start` is recomputed unconditionally (a re-used free node carries a stale box); above it the walk stops at the
    // first node whose box does not change: every ancestor was the exact union of its children before
    void refitUp(const int32_t start)
    {
        for (int32_t i = start; i >= 0; i = parent[i])
        {
            const BNode &l = bn[bn[i].left], &r = bn[bn[i].right];
            bool changed = false;
            for (int k = 0; k < 3; ++k)
            {
                const float lo = std::fmin(l.lo[k], r.lo[k]), hi = std::fmax(l.hi[k], r.hi[k]);
                changed = changed || lo != bn[i].lo[k] || hi != bn[i].hi[k];
                bn[i].lo[k] = lo, bn[i].hi[k] = hi;
            }
            if (!changed && i != start)
                break;
        }
    }
    double innerAreaSum() const
    {
        double sum = 0;
        std::vector<int32_t> st{0};
        while (!st.empty())
        {
            const int32_t i = st.back();
            st.pop_back();
            if (isLeaf(i))
                continue;
            sum += area(i);
            st.push_back(bn[i].left), st.push_back(bn[i].right);
        }
        return sum;
    }
    int depth() const
    {
        int best = 0;
        std::vector<std::pair<int32_t, int>> st{{0, 1}};
        while (!st.empty())
        {
            const auto it = st.back();
            st.pop_back();
            best = std::max(best, it.second);
            if (!isLeaf(it.first))
                st.push_back({bn[it.first].left, it.second + 1}), st.push_back({bn[it.first].right, it.second + 1});
        }
        return best;
    }
    // best place for subtree x: the node t that x should become the sibling of (never the root)
    int32_t findTarget(int32_t x)
    {
        const float ax = area(x);
        float best = INFINITY;
        int32_t target = -1;
        typedef std::pair<float, int32_t> QE; // (induced cost of the ancestors, node), smallest first
        std::vector<QE> &heap = heap_;
        heap.clear();
        auto push = [&](QE e) {
            heap.push_back(e);
            std::push_heap(heap.begin(), heap.end(), std::greater<QE>());
        };
        const float rootInduced = unionArea(0, x) - area(0);
        push({rootInduced, bn[0].left}), push({rootInduced, bn[0].right});
        while (!heap.empty())
        {
            std::pop_heap(heap.begin(), heap.end(), std::greater<QE>());
            const QE e = heap.back();
            heap.pop_back();
            if (e.first + ax >= best)
                break; // every remaining candidate costs at least its induced part plus area(x)
            const float direct = unionArea(e.second, x);
            if (e.first + direct < best)
                best = e.first + direct, target = e.second;
            const float below = e.first + direct - area(e.second);
            if (!isLeaf(e.second) && below + ax < best)
                push({below, bn[e.second].left}), push({below, bn[e.second].right});
        }
        return target;
    }
    void replaceChild(int32_t p, int32_t from, int32_t to)
    {
        (bn[p].left == from ? bn[p].left : bn[p].right) = to;
        parent[to] = p;
    }
    // one pass over the inner nodes in decreasing order of the paper's combined inefficiency measure
    void pass(double fraction)
    {
        std::vector<std::pair<float, int32_t>> cand;
        for (int32_t i = 1; i < n_nodes; ++i)
        {
            if (isLeaf(i) || parent[i] <= 0) // needs a parent and a grandparent
                continue;
            const float a = area(i), al = area(bn[i].left), ar = area(bn[i].right);
            const float msum = a / (0.5f * (al + ar) + 1e-30f), mmin = a / (std::fmin(al, ar) + 1e-30f);
            cand.push_back({a * msum * mmin, i});
        }
        std::sort(cand.begin(), cand.end(), [](const std::pair<float, int32_t> &x, const std::pair<float, int32_t> &y) {
            return x.first > y.first || (x.first == y.first && x.second < y.second);
        });
        // large scenes: the nodes that matter are the big ones near the top; a cap keeps the build time bounded
        const size_t take = std::min((size_t)(fraction * (double)cand.size()), (size_t)max_per_pass);
        for (size_t c = 0; c < take; ++c)
        {
            const int32_t n = cand[c].second;
            if (isLeaf(n) || parent[n] <= 0)
                continue; // restructured earlier in this pass
            const int32_t p = parent[n], g = parent[p];
            const int32_t sib = (bn[p].left == n) ? bn[p].right : bn[p].left;
            int32_t sub[2] = {bn[n].left, bn[n].right};
            if (area(sub[0]) < area(sub[1]))
                std::swap(sub[0], sub[1]);
            replaceChild(g, p, sib); // p and n are now free nodes
            refitUp(g);
            const int32_t freeNode[2] = {n, p};
            for (int k = 0; k < 2; ++k)
            {
                const int32_t x = sub[k], t = findTarget(x), q = freeNode[k];
                const int32_t pt = parent[t];
                replaceChild(pt, t, q);
                bn[q].left = t, bn[q].right = x;
                parent[t] = q, parent[x] = q;
                refitUp(q);
            }
        }
    }
};

// Summed surface area of all child boxes of the 4-wide tree that the greedy collapse of buildWide makes of `bn` (the
// expected number of box tests per random ray, up to the root's area), and that tree's depth.
double collapseCost(const std::vector<BNode> &bn, int *depth_out)
{
    if (bn[0].right == -2)
    {
        *depth_out = 0;
        return 0;
    }
    double sum = 0;
    int maxDepth = 1;
    std::vector<std::pair<int32_t, int>> todo{{0, 1}};
    while (!todo.empty())
    {
        const auto it = todo.back();
        todo.pop_back();
        maxDepth = std::max(maxDepth, it.second);
        int32_t kids[4] = {bn[it.first].left, bn[it.first].right, -1, -1};
        int nk = 2;
        while (nk < 4)
        {
            int pick = -1;
            float best = -1.f;
            for (int k = 0; k < nk; ++k)
                if (bn[kids[k]].right != -2)
                {
                    const float a = halfArea(bn[kids[k]].lo, bn[kids[k]].hi);
                    if (a > best)
                        best = a, pick = k;
                }
            if (pick < 0)
                break;
            const int32_t c = kids[pick];
            kids[pick] = bn[c].left;
            kids[nk++] = bn[c].right;
        }
        for (int k = 0; k < nk; ++k)
        {
            sum += halfArea(bn[kids[k]].lo, bn[kids[k]].hi);
            if (bn[kids[k]].right != -2)
                todo.push_back({kids[k], it.second + 1});
        }
    }
    *depth_out = maxDepth;
    return sum;
}

// Optimal 4-wide collapse by dynamic programming (after Ylitie, Karras, Laine 2017, "Efficient incoherent ray traversal
// on GPUs through compressed wide BVHs", §3.1), for the measure collapseCost uses: the summed surface area of all child
// boxes.  S[x] = cost of binary node x occupying one child slot (its own box + the best 4 slots below it, if inner);
// G[x][k] = cheapest cover of x's subtree with at most k slots; F[x][k] = the same when x itself may not be a slot.
// EXPERIMENTAL (TRT_COLLAPSE=optimal): the host-side figure improves by 0.2-7 % over the greedy collapse; unmeasured
// on the GPU so far, hence not the default.
struct OptimalCollapse
{
    const std::vector<BNode> &bn;
    std::vector<float> S;
    std::vector<std::array<float, 5>> F, G;

    explicit OptimalCollapse(const std::vector<BNode> &b, int32_t used) : bn(b), S(used, 0.f), F(used), G(used)
    {
        std::vector<int32_t> order, st{0};
        while (!st.empty())
        {
            const int32_t i = st.back();
            st.pop_back();
            order.push_back(i);
            if (bn[i].right != -2)
                st.push_back(bn[i].left), st.push_back(bn[i].right);
        }
        for (size_t oi = order.size(); oi-- > 0;)
        {
            const int32_t x = order[oi];
            const float a = halfArea(bn[x].lo, bn[x].hi);
            if (bn[x].right == -2)
            {
                S[x] = a;
                for (int k = 1; k <= 4; ++k)
                    G[x][k] = a, F[x][k] = INFINITY;
                continue;
            }
            const int32_t l = bn[x].left, r = bn[x].right;
            F[x][1] = INFINITY;
            for (int k = 2; k <= 4; ++k)
            {
                float best = INFINITY;
                for (int i = 1; i < k; ++i)
                    best = std::fmin(best, G[l][i] + G[r][k - i]);
                F[x][k] = best;
            }
            S[x] = a + F[x][4];
            G[x][1] = S[x];
            for (int k = 2; k <= 4; ++k)
                G[x][k] = std::fmin(S[x], F[x][k]);
        }
    }
    // the binary nodes that become the slots covering y's subtree with at most k slots
    void cover(int32_t y, int k, int32_t *slots, int &n) const
    {
        if (k == 1 || bn[y].right == -2 || !(F[y][k] < S[y]))
        {
            slots[n++] = y;
            return;
        }
        split(y, k, slots, n);
    }
    // y itself is not a slot: distribute k slots over its two children
    void split(int32_t y, int k, int32_t *slots, int &n) const
    {
        const int32_t l = bn[y].left, r = bn[y].right;
        int bi = 1;
        for (int i = 1; i < k; ++i)
            if (G[l][i] + G[r][k - i] < G[l][bi] + G[r][k - bi])
                bi = i;
        cover(l, bi, slots, n);
        cover(r, k - bi, slots, n);
    }
};
} // namespace

std::string buildWide(const trt_scene_desc &desc, AccelBuild &out)
{
    out.wide_nodes.clear();
    out.wide_root = TRT_LINK_EMPTY;
    out.wide_depth = 0;
    out.fast_geom.clear(), out.fast_key.clear(), out.fast_rank.clear(), out.fast_orig.clear(), out.fast_leaf.clear();
    out.ref_leaf_box.clear();
    out.light_box.clear();
    const int nn = desc.n_nodes, n = desc.n_tris;
    // TRT_BUILD_TIMING: host-clock phases of this function on stderr (diagnostics)
    const bool timing = getenv("TRT_BUILD_TIMING") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto phase = [&](const char *what) {
        if (!timing)
            return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "buildWide: %-28s %8.3f s\n", what, std::chrono::duration<double>(now - t_prev).count());
        t_prev = now;
    };
    const char *env = getenv("TRT_WIDE_SOURCE");
    const std::string mode = env ? env : "triangles";
    if (mode == "off")
        return "wide layout disabled by TRT_WIDE_SOURCE=off";
    const bool leafMode = (mode == "leaves");

    // reference leaves: box, triangle range, and the leaf each triangle belongs to
    struct RefLeaf
    {
        int32_t first, num;
    };
    std::vector<RefLeaf> leaves;
    std::vector<int32_t> leafOfTri(n, -1);
    float scale = 0.f;
    for (int i = 0; i < nn; ++i)
    {
        const int32_t *lk = desc.node_link + (size_t)i * 4;
        if (lk[3] <= 0)
            continue;
        const float *b = desc.node_box + (size_t)i * 6;
        out.ref_leaf_box.push_back(make_float4(b[0], b[1], b[2], 0.f));
        out.ref_leaf_box.push_back(make_float4(b[3], b[4], b[5], 0.f));
        for (int k = 0; k < lk[3]; ++k)
            leafOfTri[lk[2] + k] = (int32_t)leaves.size();
        leaves.push_back({lk[2], lk[3]});
        for (int a = 0; a < 6; ++a)
            if (std::isfinite(b[a]))
                scale = std::fmax(scale, std::fabs(b[a]));
    }
    if (leaves.empty())
        return "";
    // a reference tree that is ONE leaf is scanned without any box test (bvh.cpp:151-154): no per-hit box check then
    out.root_is_reference_leaf = (leaves.size() == 1);

    // Own boxes are padded by 256 ulp(scene scale), scale = largest |coordinate| of the reference leaf boxes.  A
    // reported hit point S + d*t lies within a few ulp(|S| + t) of its (non-sliver) triangle, and rays whose origin
    // is farther than 4 * scale from the coordinate origin take the exhaustive walk, so the ray passes the padded box of
    // every triangle it can hit with a margin of well over ten times the rounding of the slab test.
    // A pad tied to the scale rather than a fixed length matters for finely tessellated geometry: staircase has
    // 25 920 millimetre-sized triangles, which a fixed 4e-3 pad would blow up eightfold.
    const float pad = 256.f * (std::nextafter(scale, INFINITY) - scale) + 1e-30f;
    out.scene_scale = scale;
    std::vector<Prim> prims;
    if (leafMode)
    {
        for (size_t l = 0; l < leaves.size(); ++l)
        {
            Prim p;
            const float4 lo = out.ref_leaf_box[2 * l], hi = out.ref_leaf_box[2 * l + 1];
            p.lo[0] = lo.x, p.lo[1] = lo.y, p.lo[2] = lo.z, p.hi[0] = hi.x, p.hi[1] = hi.y, p.hi[2] = hi.z;
            for (int a = 0; a < 3; ++a)
                p.c[a] = 0.5f * (p.lo[a] + p.hi[a]);
            p.payload = (int32_t)l;
            p.w = 1.0f + (float)leaves[l].num;
            prims.push_back(p);
        }
    }
    else
    {
        for (int t = 0; t < n; ++t)
        {
            const float *v = desc.v + (size_t)t * 9, *N = desc.normal + (size_t)t * 3;
            bool ok = leafOfTri[t] >= 0 && std::isfinite(N[0]) && std::isfinite(N[1]) && std::isfinite(N[2]);
            for (int k = 0; k < 9; ++k)
                ok = ok && std::isfinite(v[k]);
            if (!ok)
                continue; // a NaN normal / non-finite vertex can never pass interactTriangle's comparisons
            Prim p;
            // Needle / sliver triangles: the three edge tests of interactTriangle bound the hit point only up to
            // (rounding noise) / sin(smallest angle): an accepted point may lie beyond the triangle along its long
            // axis by about ext = 1e-6 * emax / (height / emax)  (1e-6 = twice the ~8-operation float noise of an edge
            // test relative to |edge| |P - p|).  The triangle's box is padded by max(pad, 4 ext); a needle so thin that
            // 4 ext exceeds a quarter of its length keeps the box the reference itself culls it with — its reference
            // leaf's — and so behaves exactly as in the leaf-atomic layout.
            const double e1[3] = {(double)v[3] - v[0], (double)v[4] - v[1], (double)v[5] - v[2]};
            const double e2[3] = {(double)v[6] - v[0], (double)v[7] - v[1], (double)v[8] - v[2]};
            const double e3[3] = {(double)v[6] - v[3], (double)v[7] - v[4], (double)v[8] - v[5]};
            const double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
            const double area2 = std::sqrt(cx * cx + cy * cy + cz * cz);
            double emax2 = 0;
            for (const double *e : {e1, e2, e3})
                emax2 = std::fmax(emax2, e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
            const double emax = std::sqrt(emax2), ext = 1e-6 * emax2 * emax / area2;
            const double pad_t = std::fmax((double)pad, 4.0 * ext);
            if (4.0 * ext < 0.25 * emax) // also false for NaN / zero-area
            {
                for (int a = 0; a < 3; ++a)
                {
                    p.lo[a] = std::fmin(v[a], std::fmin(v[3 + a], v[6 + a])) - (float)pad_t;
                    p.hi[a] = std::fmax(v[a], std::fmax(v[3 + a], v[6 + a])) + (float)pad_t;
                }
                if (pad_t > (double)pad)
                    ++out.n_sliver;
            }
            else
            {
                const float4 lo = out.ref_leaf_box[2 * leafOfTri[t]], hi = out.ref_leaf_box[2 * leafOfTri[t] + 1];
                p.lo[0] = lo.x, p.lo[1] = lo.y, p.lo[2] = lo.z, p.hi[0] = hi.x, p.hi[1] = hi.y, p.hi[2] = hi.z;
                ++out.n_needle;
            }
            for (int a = 0; a < 3; ++a)
                p.c[a] = 0.5f * (p.lo[a] + p.hi[a]);
            p.payload = t;
            p.w = 1.0f;
            prims.push_back(p);
        }
        if (prims.empty())
            return "";
    }

    // per light: union of the boxes that hold its material's triangles (SceneView::light_box)
    {
        const float big = 3.0e38f;
        out.light_box.assign((size_t)2 * std::max(0, desc.n_lights), make_float4(0.f, 0.f, 0.f, 0.f));
        for (int l = 0; l < desc.n_lights; ++l)
            out.light_box[2 * l] = make_float4(big, big, big, 0.f), out.light_box[2 * l + 1] = make_float4(-big, -big, -big, 0.f);
        for (const Prim &p : prims)
        {
            const int t0 = leafMode ? leaves[p.payload].first : p.payload;
            const int cnt = leafMode ? leaves[p.payload].num : 1;
            for (int t = t0; t < t0 + cnt; ++t)
                for (int l = 0; l < desc.n_lights; ++l)
                    if (desc.lights[l].material == desc.mtl[t])
                    {
                        float4 &lo = out.light_box[2 * l], &hi = out.light_box[2 * l + 1];
                        lo.x = std::fmin(lo.x, p.lo[0]), lo.y = std::fmin(lo.y, p.lo[1]), lo.z = std::fmin(lo.z, p.lo[2]);
                        hi.x = std::fmax(hi.x, p.hi[0]), hi.y = std::fmax(hi.y, p.hi[1]), hi.z = std::fmax(hi.z, p.hi[2]);
                    }
        }
    }

    const char *lenv = getenv("TRT_FAST_LEAF"); // triangles per leaf of the fast layout (1..8), default 2: measured best on every scene from 26 to 10 M triangles
    const int maxLeafTris = lenv ? std::max(1, std::min(8, atoi(lenv))) : 2;
    phase("scan units (boxes, pads)");
    WideBuilder wb{prims, leafMode ? 1 : maxLeafTris, {}, {}};
    wb.order.resize(prims.size());
    for (size_t i = 0; i < prims.size(); ++i)
        wb.order[i] = (int32_t)i;
    wb.bn.resize(prims.size() * 2);
    wb.build(0, (int)prims.size());
    phase("binned-SAH binary build");
    {
        // insertion-based optimisation (Reinserter above): up to two passes (one beyond 500 k triangles, where a pass
        // costs seconds), keeping whichever tree — the binned-SAH one included — collapses to the cheapest 4-wide tree
        // that still fits the traversal stack.  TRT_REINSERT=0 switches it off, =N sets the number of passes.
        const char *renv = getenv("TRT_REINSERT");
        const int maxPasses = renv ? atoi(renv) : (prims.size() > 500000 ? 1 : 2);
        const int32_t used = wb.next.load();
        if (maxPasses > 0 && used >= 7)
        {
            int depth = 0;
            double bestCost = collapseCost(wb.bn, &depth);
            std::vector<BNode> best(wb.bn.begin(), wb.bn.begin() + used);
            Reinserter ri(wb.bn, used);
            for (int pass = 0; pass < maxPasses; ++pass)
            {
                ri.pass(pass == 0 ? 1.0 : 0.5);
                const double cost = collapseCost(wb.bn, &depth);
                if (cost < bestCost && 3 * depth + 2 <= TRT_WIDE_STACK)
                {
                    bestCost = cost;
                    std::copy(wb.bn.begin(), wb.bn.begin() + used, best.begin());
                }
            }
            std::copy(best.begin(), best.end(), wb.bn.begin());
        }
    }
    const std::vector<BNode> &bn = wb.bn;
    phase("reinsertion passes");

    // triangles of a binary leaf, appended to the fast arrays in leaf order; returns the leaf link
    auto emitLeaf = [&](const BNode &c) -> int32_t {
        const int32_t first = (int32_t)out.fast_orig.size();
        for (int i = c.first; i < c.first + c.count; ++i)
        {
            const Prim &p = prims[wb.order[i]];
            const int t0 = leafMode ? leaves[p.payload].first : p.payload;
            const int cnt = leafMode ? leaves[p.payload].num : 1;
            for (int t = t0; t < t0 + cnt; ++t)
            {
                out.fast_orig.push_back(t);
                out.fast_geom.push_back(out.tri_geom[t]);
                out.fast_key.push_back(out.tri_key[t]);
                out.fast_rank.push_back(out.tri_rank[t]);
                out.fast_leaf.push_back(leafOfTri[t]);
            }
        }
        const int32_t count = (int32_t)out.fast_orig.size() - first;
        return ~((first << 3) | (count - 1)); // count <= 8 in either mode
    };

    if (bn[0].right == -2)
    {
        // a single scan unit: no node, no box of this layout is ever tested
        out.wide_root = emitLeaf(bn[0]);
        return "";
    }

    // collapse: a wide node adopts grandchildren, largest surface area first, until it has 4 children
    struct Item
    {
        int32_t bnode, wide, depth;
    };
    std::vector<Item> todo;
    std::unique_ptr<OptimalCollapse> optimal;
    {
        const char *cenv = getenv("TRT_COLLAPSE");
        if (cenv && std::string(cenv) == "optimal" && prims.size() <= 4000000)
            optimal.reset(new OptimalCollapse(bn, wb.next.load()));
    }
    out.wide_nodes.emplace_back();
    todo.push_back({0, 0, 1});
    int maxDepth = 1;
    double sah = 0;
    const double rootArea = halfArea(bn[0].lo, bn[0].hi);
    while (!todo.empty())
    {
        const Item it = todo.back();
        todo.pop_back();
        maxDepth = it.depth > maxDepth ? it.depth : maxDepth;
        int32_t kids[4] = {bn[it.bnode].left, bn[it.bnode].right, -1, -1};
        int nk = 2;
        if (optimal)
        {
            nk = 0;
            optimal->split(it.bnode, 4, kids, nk);
        }
        while (!optimal && nk < 4)
        {
            int pick = -1;
            float best = -1.f;
            for (int k = 0; k < nk; ++k)
                if (bn[kids[k]].right != -2)
                {
                    const float a = halfArea(bn[kids[k]].lo, bn[kids[k]].hi);
                    if (a > best)
                        best = a, pick = k;
                }
            if (pick < 0)
                break;
            const int32_t c = kids[pick];
            kids[pick] = bn[c].left;
            kids[nk++] = bn[c].right;
        }
        float lo[3][4], hi[3][4];
        int32_t link[4];
        for (int k = 0; k < 4; ++k)
        {
            if (k >= nk)
            {
                for (int a = 0; a < 3; ++a)
                    lo[a][k] = hi[a][k] = NAN;
                link[k] = TRT_LINK_EMPTY;
                continue;
            }
            const BNode &c = bn[kids[k]];
            for (int a = 0; a < 3; ++a)
                lo[a][k] = c.lo[a], hi[a][k] = c.hi[a];
            sah += halfArea(c.lo, c.hi) / rootArea;
            if (c.right == -2)
                link[k] = emitLeaf(c);
            else
            {
                link[k] = (int32_t)out.wide_nodes.size();
                out.wide_nodes.emplace_back();
                todo.push_back({kids[k], link[k], it.depth + 1});
            }
        }
        WideNode &w = out.wide_nodes[it.wide];
        w.lox = make_float4(lo[0][0], lo[0][1], lo[0][2], lo[0][3]);
        w.loy = make_float4(lo[1][0], lo[1][1], lo[1][2], lo[1][3]);
        w.loz = make_float4(lo[2][0], lo[2][1], lo[2][2], lo[2][3]);
        w.hix = make_float4(hi[0][0], hi[0][1], hi[0][2], hi[0][3]);
        w.hiy = make_float4(hi[1][0], hi[1][1], hi[1][2], hi[1][3]);
        w.hiz = make_float4(hi[2][0], hi[2][1], hi[2][2], hi[2][3]);
        w.link = make_int4(link[0], link[1], link[2], link[3]);
        w.pad = make_int4(0, 0, 0, 0);
    }
    phase("collapse + emission");
    out.wide_root = 0;
    out.wide_depth = maxDepth;
    out.sah_wide = sah;
    if (3 * maxDepth + 2 > TRT_WIDE_STACK)
    {
        // deeper than the per-thread stack: keep the reference-topology kernel for this scene
        out.wide_nodes.clear();
        out.wide_root = TRT_LINK_EMPTY;
        return "wide layout too deep for its stack";
    }
    return "";
}

// ------------------------------------------------------------------------------------------------------------
std::string checkLayout(const trt_scene_desc &desc, const AccelBuild &ab, trt_layout_report &rep)
{
    const int n = desc.n_tris, nn = desc.n_nodes;
    rep = trt_layout_report();
    rep.n_tris = n;
    rep.n_fast_tris = (int32_t)ab.fast_orig.size();
    rep.wide_nodes = (int32_t)ab.wide_nodes.size();
    rep.wide_depth = ab.wide_depth;
    rep.ref_leaves = ab.n_leaves;
    rep.ref_depth = ab.ref_depth;
    rep.slivers = ab.n_sliver, rep.needles = ab.n_needle;
    rep.sah_wide = ab.sah_wide;
    rep.use_wide = (ab.wide_root != TRT_LINK_EMPTY) ? 1 : 0;
    std::string first;
    auto bad = [&](const std::string &what) {
        if (rep.violations++ == 0)
            first = what;
    };
    const size_t nf = ab.fast_orig.size();
    if (ab.fast_geom.size() != nf || ab.fast_key.size() != nf || ab.fast_leaf.size() != nf || ab.fast_rank.size() != nf)
        bad("fast arrays differ in length");
    if (rep.violations)
        return first;

    // reference leaf of every triangle, leaf boxes in node order
    std::vector<int32_t> leafOfTri(n, -1);
    std::vector<const float *> leafBox;
    float scale = 0.f;
    for (int i = 0; i < nn; ++i)
    {
        const int32_t *lk = desc.node_link + (size_t)i * 4;
        if (lk[3] <= 0)
            continue;
        for (int k = 0; k < lk[3]; ++k)
            leafOfTri[lk[2] + k] = (int32_t)leafBox.size();
        leafBox.push_back(desc.node_box + (size_t)i * 6);
        for (int a = 0; a < 6; ++a)
            if (std::isfinite(leafBox.back()[a]))
                scale = std::fmax(scale, std::fabs(leafBox.back()[a]));
    }
    if (!rep.use_wide)
        return first; // the scene keeps the reference-topology kernels: there is no fast layout to check
    if (ab.ref_leaf_box.size() != 2 * leafBox.size())
        bad("ref_leaf_box does not hold one box per reference leaf");
    else
        for (size_t l = 0; l < leafBox.size(); ++l)
        {
            const float4 lo = ab.ref_leaf_box[2 * l], hi = ab.ref_leaf_box[2 * l + 1];
            const float *b = leafBox[l];
            if (std::memcmp(&lo, b, 12) != 0 || std::memcmp(&hi, b + 3, 12) != 0)
                bad("ref_leaf_box differs from the reference's leaf box");
        }

    // every triangle at most once, the missing ones can never be hit, per-triangle data carried over bit for bit
    std::vector<int32_t> seen(n, 0);
    for (size_t i = 0; i < nf; ++i)
    {
        const int32_t t = ab.fast_orig[i];
        if (t < 0 || t >= n)
        {
            bad("fast_orig out of range");
            continue;
        }
        if (seen[t]++)
            bad("triangle appears twice in the fast layout");
        if (ab.fast_leaf[i] != leafOfTri[t])
            bad("fast_leaf is not the triangle's reference leaf");
        if (std::memcmp(&ab.fast_geom[i], &ab.tri_geom[t], sizeof(TriGeom)) != 0 || ab.fast_key[i] != ab.tri_key[t] ||
            ab.fast_rank[i] != ab.tri_rank[t])
            bad("fast_geom / fast_key / fast_rank differ from the post-build arrays");
    }
    if (rep.use_wide)
        for (int t = 0; t < n; ++t)
        {
            if (seen[t])
                continue;
            ++rep.n_dropped;
            const float *v = desc.v + (size_t)t * 9, *N = desc.normal + (size_t)t * 3;
            bool finite = std::isfinite(N[0]) && std::isfinite(N[1]) && std::isfinite(N[2]);
            for (int k = 0; k < 9; ++k)
                finite = finite && std::isfinite(v[k]);
            if (finite && leafOfTri[t] >= 0)
                bad("a triangle that can be hit is missing from the fast layout");
        }
    if (!rep.use_wide || rep.violations)
        return first;

    // walk the tree: leaves tile [0, nf), every node is reached once, boxes nest and hold their triangles with the pad
    const float pad = 256.f * (std::nextafter(scale, INFINITY) - scale);

    // light boxes (the early stop of occluded light samples rests on them): every triangle of the fast layout that
    // carries a light's material lies inside that light's box, with the pad unless it is a needle under its reference
    // leaf's box — the same containment its path of boxes has to offer
    if (ab.light_box.size() != (size_t)2 * std::max(0, desc.n_lights))
        bad("light_box does not hold one box per light");
    else
        for (size_t i = 0; i < nf; ++i)
        {
            const int32_t t = ab.fast_orig[i];
            const float *v = desc.v + (size_t)t * 9;
            const float *rb = leafBox[ab.fast_leaf[i]];
            for (int l = 0; l < desc.n_lights; ++l)
            {
                if (desc.lights[l].material != desc.mtl[t])
                    continue;
                const float4 lo4 = ab.light_box[2 * l], hi4 = ab.light_box[2 * l + 1];
                const float lo[3] = {lo4.x, lo4.y, lo4.z}, hi[3] = {hi4.x, hi4.y, hi4.z};
                bool padded = true, refbox = true;
                for (int a = 0; a < 3; ++a)
                {
                    const float vlo = std::fmin(v[a], std::fmin(v[3 + a], v[6 + a])), vhi = std::fmax(v[a], std::fmax(v[3 + a], v[6 + a]));
                    padded = padded && lo[a] <= vlo - pad && hi[a] >= vhi + pad;
                    refbox = refbox && lo[a] <= rb[a] && hi[a] >= rb[3 + a];
                }
                if (!(padded || refbox))
                    bad("a light's box does not contain one of its material's triangles with the pad");
            }
        }
    if (rep.violations)
        return first;
    std::vector<int32_t> covered(nf, 0), visited(ab.wide_nodes.size(), 0);
    auto leafInside = [&](int32_t link, const float *lo, const float *hi) {
        const int first = (~link) >> 3, count = ((~link) & 7) + 1;
        if (first < 0 || (size_t)first + count > nf)
        {
            bad("leaf link out of range");
            return;
        }
        for (int i = first; i < first + count; ++i)
        {
            covered[i]++;
            const float *v = desc.v + (size_t)ab.fast_orig[i] * 9;
            const float *rb = leafBox[ab.fast_leaf[i]];
            bool padded = true, plain = true, refbox = true;
            for (int a = 0; a < 3; ++a)
            {
                const float vlo = std::fmin(v[a], std::fmin(v[3 + a], v[6 + a])), vhi = std::fmax(v[a], std::fmax(v[3 + a], v[6 + a]));
                padded = padded && lo[a] <= vlo - pad && hi[a] >= vhi + pad;
                plain = plain && lo[a] <= vlo && hi[a] >= vhi;
                refbox = refbox && lo[a] <= rb[a] && hi[a] >= rb[3 + a];
            }
            // a needle keeps its reference leaf's box (which holds its vertices); everything else gets the pad
            if (!plain || !(padded || refbox))
                bad("a box on the path to a triangle does not contain it with the pad");
        }
    };
    if (ab.wide_root < 0)
    {
        const float lo[3] = {-INFINITY, -INFINITY, -INFINITY}, hi[3] = {INFINITY, INFINITY, INFINITY};
        leafInside(ab.wide_root, lo, hi); // single scan unit: no box of this layout is tested
    }
    else
    {
        struct Item
        {
            int32_t node, depth;
            float lo[3], hi[3];
        };
        std::vector<Item> todo;
        todo.push_back({0, 1, {-INFINITY, -INFINITY, -INFINITY}, {INFINITY, INFINITY, INFINITY}});
        int maxDepth = 0;
        // the root's box = union of its children's boxes (the measure buildWide normalises sah_wide with)
        double rootHalfArea = 1.0;
        {
            const WideNode &w = ab.wide_nodes[0];
            const float *b6[6] = {&w.lox.x, &w.loy.x, &w.loz.x, &w.hix.x, &w.hiy.x, &w.hiz.x};
            float rlo[3] = {INFINITY, INFINITY, INFINITY}, rhi[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int k = 0; k < 4; ++k)
                if ((&w.link.x)[k] != TRT_LINK_EMPTY)
                    for (int a = 0; a < 3; ++a)
                        rlo[a] = std::fmin(rlo[a], b6[a][k]), rhi[a] = std::fmax(rhi[a], b6[3 + a][k]);
            rootHalfArea = (double)halfArea(rlo, rhi);
        }
        while (!todo.empty())
        {
            const Item it = todo.back();
            todo.pop_back();
            if (it.node < 0 || (size_t)it.node >= ab.wide_nodes.size())
            {
                bad("inner link out of range");
                continue;
            }
            if (visited[it.node]++)
            {
                bad("wide node reached twice");
                continue;
            }
            maxDepth = it.depth > maxDepth ? it.depth : maxDepth;
            const WideNode &w = ab.wide_nodes[it.node];
            const float *lox = &w.lox.x, *loy = &w.loy.x, *loz = &w.loz.x, *hix = &w.hix.x, *hiy = &w.hiy.x, *hiz = &w.hiz.x;
            const int32_t *lk = &w.link.x;
            int nk = 0;
            for (int k = 0; k < 4; ++k)
            {
                if (lk[k] == TRT_LINK_EMPTY)
                {
                    if (k < 2)
                        bad("a wide node needs at least two children in slots 0 and 1");
                    continue;
                }
                ++nk;
                const float lo[3] = {lox[k], loy[k], loz[k]}, hi[3] = {hix[k], hiy[k], hiz[k]};
                for (int a = 0; a < 3; ++a)
                    if (!(lo[a] <= hi[a]) || !(lo[a] >= it.lo[a]) || !(hi[a] <= it.hi[a]))
                        bad("child box not finite, inverted, or not inside its parent's box");
                const double rel = (double)halfArea(lo, hi) / rootHalfArea;
                (lk[k] < 0 ? rep.sah_leaf : rep.sah_inner) += rel;
                if (lk[k] < 0)
                    leafInside(lk[k], lo, hi);
                else
                {
                    Item c{lk[k], it.depth + 1, {lo[0], lo[1], lo[2]}, {hi[0], hi[1], hi[2]}};
                    todo.push_back(c);
                }
            }
            if (nk < 2)
                bad("wide node with fewer than two children");
        }
        for (size_t i = 0; i < visited.size(); ++i)
            if (!visited[i])
                bad("unreachable wide node");
        if (maxDepth != ab.wide_depth)
            bad("wide_depth differs from the tree's depth");
        if (3 * maxDepth + 2 > TRT_WIDE_STACK)
            bad("tree deeper than the traversal stack allows");
    }
    for (size_t i = 0; i < nf; ++i)
        if (covered[i] != 1)
            bad("a fast-layout triangle is not in exactly one leaf");
    return first;
}
} // namespace trt
