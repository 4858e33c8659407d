// GPU acceleration layouts and the device-resident scene (internal header shared by the .cu files).
#pragma once
#include "../../include/trt.h"
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

namespace trt
{
// Tie key of the "miss" state: above every non-emissive key, below every emissive key (bit 31), so that a hit
// at exactly t == INF is accepted only for an emissive triangle, as bvh.cpp:219 does.
#define TRT_MISS_KEY 0x7FFFFFFFu

// per-thread traversal stack entries for the reference topology; scenes whose reference tree is deeper are
// rejected at trt_scene_create (staircase: depth 59)
#define TRT_REF_STACK_LIMIT 64

// ---- reference topology, two child boxes per inner node (64 B) --------------------------------------------
// a = (L.AA.x L.AA.y L.AA.z L.BB.x)  b = (L.BB.y L.BB.z R.AA.x R.AA.y)  c = (R.AA.z R.BB.x R.BB.y R.BB.z)
// d = (left link, right link, parent, -);  link >= 0: inner node index;  link < 0: leaf, ~link = first<<3 | (num-1);
// parent = (parent's inner index << 1) | (1 if this node is its right child), -1 for the root
struct __align__(16) RefNode
{
    float4 a, b, c;
    int4 d;
};

// ---- fast layout: 4-wide BVH whose primitives are the reference's LEAVES (bvh.cpp:43-48), 128 B per node ----
// Child k's box is (lox[k] loy[k] loz[k]) - (hix[k] hiy[k] hiz[k]).  A leaf child carries the reference leaf's
// own padded box bit for bit (so its slab test IS the reference's interactAABB on that leaf); an inner child's
// box is the exact float union of the leaf boxes below it.  link >= 0: wide node index; link < 0: reference
// leaf, ~link = first<<3 | (num-1); TRT_LINK_EMPTY: unused slot.
#define TRT_LINK_EMPTY 0x7fffffff
#define TRT_LINK_EXIT ((int32_t)0x80000000)
#define TRT_WIDE_STACK 96
struct __align__(16) WideNode
{
    float4 lox, loy, loz, hix, hiy, hiz;
    int4 link;
    int4 pad;
};

// Triangle for the intersection test (48 B): q0 = (N.xyz p1.x)  q1 = (p1.y p1.z p2.x p2.y)  q2 = (p2.z p3.xyz).
// The plane part of the test (dot(N,d), t) needs q0 and q1 only; the inside test reads q2 as well.
struct __align__(16) TriGeom
{
    float4 q0, q1, q2;
};

// Shading attributes of a triangle (fetched once per path vertex, not during traversal)
struct __align__(16) TriShade
{
    float vn[9];
    float vt[6];
    int32_t mtl;
};

struct DeviceMaterial
{
    float3 Kd, Ks, Tr, radiance;
    float Ns, Ni;
    int32_t is_emissive, texture;
    double area;
    // lobe-pick probabilities of nextRay (pathTracing.cpp:191-192): |Kd| / (|Kd| + |Ks|) and |Ks| / (|Kd| + |Ks|) with the
    // MTL's Kd, lengths in float, quotients in double — a function of the material alone, computed once on the host
    double kd, ks;
};

struct DeviceLight
{
    int32_t material, first_tri, n_tris;
    float pdf; // (float)(double(1) / total area) of pathTracing.cpp:62, the same for every sample of the light
};

struct DeviceTexture
{
    int32_t rows, cols;
    const uint8_t *bgr;
};

struct DeviceCamera
{
    float3 eye, llc, horizontal, vertical;
    int32_t width, height;
};

// Everything the kernels read, by value in kernel parameters.
struct SceneView
{
    // reference-topology layout
    const RefNode *ref_nodes;
    int32_t root_link; // link of the root (leaf link when the whole scene is one leaf); 0x7fffffff = empty scene
    // fast layout (wide_root: node 0, or a leaf link when the scene is a single leaf, or TRT_LINK_EMPTY)
    const WideNode *wide_nodes;
    int32_t wide_root;
    int32_t use_wide; // 0: the scene has no wide layout (too deep for its stack): reference-topology kernels only
    // the fast layout's triangles in ITS leaf order: test data, tie key / rank, post-build index, and the reference
    // leaf each triangle belongs to (a hit counts only if that leaf's box passes the reference's slab test)
    const TriGeom *fast_geom;
    const uint32_t *fast_key, *fast_rank;
    const int32_t *fast_orig, *fast_leaf;
    const int32_t *fast_mtl; // material of each fast-layout triangle (the only thing a light-sample ray needs of its hit)
    // 2 per light: the union of the fast layout's (padded) boxes of every triangle that carries the light's material.
    // No triangle of that material can be hit before a ray enters this box — what lets an occluded light sample stop at
    // the first occluder it finds in front of it (WalkRays::canStop)
    const float4 *light_box;
    const float4 *ref_leaf_box; // 2 per reference leaf: (AA.xyz, -) (BB.xyz, -)
    const int32_t *ref_leaf_parent; // per reference leaf: (parent's inner index << 1) | right-child bit, -1 if the leaf is the root
    int32_t check_leaf_box;     // 0 only when the whole scene is ONE reference leaf (scanned without a box test)
    float strict_origin_limit;  // rays starting farther than this from the coordinate origin take the strict walk
    unsigned long long *strict_counter; // rays that took the exhaustive reference walk (trt_stats.rays_strict)
    float inv_cull_limit;       // |1/d| beyond this (incl. inf) makes a class-1 ray: culling reciprocal clamped, path gate
    int32_t n_tris;
    const TriGeom *tri_geom; // post-build order
    const uint32_t *tri_key; // tie key, higher wins at equal t
    // the same order as ranks (0 = highest key; the miss state has rank `miss_rank`, between the emissive and the
    // non-emissive triangles): (t bits << 32 | rank) is then a single 64-bit key whose minimum is the winner
    const uint32_t *tri_rank;
    const int32_t *rank_tri; // rank -> triangle index (-1 for miss_rank)
    uint32_t miss_rank;
    const float *tri_v;      // n*9 original vertices (barycentric solve)
    const TriShade *tri_shade;
    const DeviceMaterial *materials;
    const DeviceLight *lights;
    const float *light_v, *light_vn;
    const double *light_cum_area;
    const DeviceTexture *textures;
    int32_t n_lights, n_materials;
    double first_light_area; // quirk A.5-1: the static distribution's range (pathTracing.cpp:38)
    DeviceCamera cam;
};

struct AccelBuild
{
    std::vector<RefNode> ref_nodes;
    int32_t root_link = 0x7fffffff;
    std::vector<TriGeom> tri_geom;
    std::vector<uint32_t> tri_key, tri_rank;
    std::vector<int32_t> rank_tri;
    uint32_t miss_rank = 0;
    int32_t n_leaves = 0, ref_depth = 0;
    std::vector<WideNode> wide_nodes;
    std::vector<TriGeom> fast_geom;
    std::vector<uint32_t> fast_key, fast_rank;
    std::vector<int32_t> fast_orig, fast_leaf;
    std::vector<float4> ref_leaf_box;
    std::vector<float4> light_box; // 2 per light (lo, hi), see SceneView::light_box
    std::vector<int32_t> ref_leaf_parent;
    bool root_is_reference_leaf = false;
    float scene_scale = 0.f;
    int32_t n_sliver = 0; // triangles whose box got the larger sliver pad
    int32_t n_needle = 0; // triangles kept under their reference leaf's box
    int32_t wide_root = TRT_LINK_EMPTY, wide_depth = 0;
    double sah_ref = 0, sah_wide = 0; // expected box tests per random ray (surface-area heuristic), for the log
};

// Builds the layouts from the reference topology in `desc`. Returns "" or an error text.
std::string buildAccel(const trt_scene_desc &desc, AccelBuild &out);
// Host-side verification of the fast layout's structural invariants (the ones the exactness argument of DESIGN.md §3
// rests on); fills `report`, returns "" or the first violation.
std::string checkLayout(const trt_scene_desc &desc, const AccelBuild &ab, trt_layout_report &report);
// Builds the 4-wide fast layout over the reference leaves into `out` (after buildAccel). A non-empty return
// means the layout is unavailable for this scene (the reference-topology kernels are used instead).
std::string buildWide(const trt_scene_desc &desc, AccelBuild &out);

// ---- layout cache (layout_cache.cu): the whole AccelBuild in one file, keyed by a hash of what it was built from ----
uint64_t layoutKey(const trt_scene_desc &desc);
std::string saveLayout(const AccelBuild &ab, bool use_wide, uint64_t key, const char *path); // "" or why not
// true: `path` held the layouts for this key, `ab` / `use_wide` are filled; false: `why` says what was wrong with the file
bool loadLayout(const char *path, uint64_t key, AccelBuild &ab, bool &use_wide, std::string &why);
} // namespace trt
