// Host-side construction of the GPU layouts from the reference topology handed over in trt_scene_desc.
#include "accel.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <functional>

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
        o.d = make_int4(link(lk[0]), link(lk[1]), 0, 0);
    }
    out.root_link = link(0);

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
// Fast layout: binned-SAH binary tree over the reference's leaves, collapsed to 4 children per node.
// The reference tree's upper levels are median-x splits (its SAH is capped at INF = 114514, bvh.cpp:49-51,
// 125-134), which costs ~140 box tests per ray on staircase; the leaves themselves stay the scan units, so
// the set of triangles tested together — and with it the reference's result — is unchanged (traverse.cuh).
namespace trt
{
namespace
{
struct Prim
{
    float lo[3], hi[3], c[3];
    int32_t link;
    float w; // scan cost of the leaf: one box test + num triangle tests
};
struct BNode
{
    float lo[3], hi[3];
    int32_t left = -1, right = -1; // binary children, or prim index in `left` when right == -2
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
    std::vector<int32_t> order;
    std::vector<BNode> bn;

    int32_t build(int l, int r) // [l, r)
    {
        const int32_t me = (int32_t)bn.size();
        bn.emplace_back();
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
        if (r - l == 1)
        {
            bn[me].left = order[l];
            bn[me].right = -2;
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
        const int32_t L = build(l, mid);
        const int32_t R = build(mid, r);
        bn[me].left = L, bn[me].right = R;
        return me;
    }
};
} // namespace

std::string buildWide(const trt_scene_desc &desc, AccelBuild &out)
{
    out.wide_nodes.clear();
    out.wide_root = TRT_LINK_EMPTY;
    out.wide_depth = 0;
    const int nn = desc.n_nodes;
    std::vector<Prim> prims;
    for (int i = 0; i < nn; ++i)
    {
        const int32_t *lk = desc.node_link + (size_t)i * 4;
        if (lk[3] <= 0)
            continue;
        const float *b = desc.node_box + (size_t)i * 6;
        Prim p;
        for (int a = 0; a < 3; ++a)
            p.lo[a] = b[a], p.hi[a] = b[3 + a], p.c[a] = 0.5f * (b[a] + b[3 + a]);
        p.link = ~((lk[2] << 3) | (lk[3] - 1));
        p.w = 1.0f + (float)lk[3];
        prims.push_back(p);
    }
    if (prims.empty())
        return "";
    if (prims.size() == 1)
    {
        out.wide_root = prims[0].link; // the reference scans a root leaf without any box test (bvh.cpp:151-154)
        return "";
    }
    WideBuilder wb{prims, {}, {}};
    // TRT_WIDE_SOURCE: "ref" collapses the reference binary tree as it is (experiments); "off" builds no wide
    // layout at all, which is what happens to scenes too deep for its stack (lets tests cover that path)
    const char *mode = getenv("TRT_WIDE_SOURCE");
    if (mode && std::string(mode) == "off")
        return "wide layout disabled by TRT_WIDE_SOURCE=off";
    if (mode && std::string(mode) == "ref")
    {
        // binary nodes = the reference's, pre-order; leaf prims in the same left-to-right order as `prims`
        std::vector<int32_t> leafPrim(nn, -1);
        int32_t lp = 0;
        for (int i = 0; i < nn; ++i)
            if (desc.node_link[(size_t)i * 4 + 3] > 0)
                leafPrim[i] = lp++;
        wb.bn.resize(nn);
        for (int i = 0; i < nn; ++i)
        {
            const int32_t *lk = desc.node_link + (size_t)i * 4;
            const float *b = desc.node_box + (size_t)i * 6;
            for (int a = 0; a < 3; ++a)
                wb.bn[i].lo[a] = b[a], wb.bn[i].hi[a] = b[3 + a];
            if (lk[3] > 0)
                wb.bn[i].left = leafPrim[i], wb.bn[i].right = -2;
            else
                wb.bn[i].left = lk[0], wb.bn[i].right = lk[1];
        }
    }
    else
    {
        wb.order.resize(prims.size());
        for (size_t i = 0; i < prims.size(); ++i)
            wb.order[i] = (int32_t)i;
        wb.bn.reserve(prims.size() * 2);
        wb.build(0, (int)prims.size());
    }
    const std::vector<BNode> &bn = wb.bn;

    // collapse: a wide node adopts grandchildren, largest surface area first, until it has 4 children
    struct Item
    {
        int32_t bnode, wide, depth;
    };
    std::vector<Item> todo;
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
                link[k] = prims[c.left].link;
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
} // namespace trt
