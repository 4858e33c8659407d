// Host-side construction of the GPU layouts from the reference topology handed over in trt_scene_desc.
#include "accel.h"

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
        out.tri_geom[i].p1nx = make_float4(v[0], v[1], v[2], N[0]);
        out.tri_geom[i].p2ny = make_float4(v[3], v[4], v[5], N[1]);
        out.tri_geom[i].p3nz = make_float4(v[6], v[7], v[8], N[2]);
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
