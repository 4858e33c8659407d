// TEST INFRASTRUCTURE ONLY (oracle tier A bridge) — never linked into, or called by, the product path.
//
// Links the UNMODIFIED reference objects (bvh.cpp, triangle.cpp, camera.cpp, scene.cpp, material.cpp,
// pathTracing.cpp compiled from /root/reference/RayTracingOnCPU against oracle/shim) and exposes them
// through a small C ABI for the ctypes test harness and bench.py's cpu_baseline / reference arm:
//   * the reference's own loaders + buildBVH (main.cpp:66-76)          -> post-build triangle order
//   * the reference's own traverseBVH (bvh.cpp:146-175) over a ray batch -> closest-hit goldens
//   * the reference's own render-loop body (main.cpp:88-108, shade())   -> statistical image reference
// Nothing here restates an algorithm: every geometric / shading decision is taken by reference code.
#include "scene.h"
#include "bvh.h"
#include "pathtracing.h"

#include <cstdint>
#include <cstring>
#include <map>
#include <omp.h>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

namespace
{
struct RefScene
{
    Scene scene;
    BVHNode *root = nullptr;
    std::vector<std::string> mtl_names;              // index -> name (first-use order over post-build tris)
    std::unordered_map<std::string, int> mtl_index;  // name -> index
    std::unordered_map<std::string, int> canon;      // triangle content bytes -> first post-build index
    std::vector<int> canon_of;                       // post-build index -> canonical index
};

std::string tri_key(const Triangle &t)
{
    std::string k;
    k.append(reinterpret_cast<const char *>(t.v), sizeof(t.v));
    k.append(reinterpret_cast<const char *>(t.vn), sizeof(t.vn));
    k.append(reinterpret_cast<const char *>(t.vt), sizeof(t.vt));
    k.append(t.mtl_name);
    return k;
}

void node_stats(const BVHNode *n, int depth, int &nodes, int &leaves, int &maxdepth)
{
    if (!n)
        return;
    nodes++;
    if (depth > maxdepth)
        maxdepth = depth;
    if (n->num > 0)
    {
        leaves++;
        return;
    }
    node_stats(n->left, depth + 1, nodes, leaves, maxdepth);
    node_stats(n->right, depth + 1, nodes, leaves, maxdepth);
}

void flatten(const BVHNode *n, std::vector<float> &boxes, std::vector<int32_t> &links)
{
    // pre-order; links = {left child slot or -1, right child slot or -1, index, num}
    size_t me = links.size() / 4;
    links.insert(links.end(), {-1, -1, n->index, n->num});
    boxes.insert(boxes.end(), {n->AA.x, n->AA.y, n->AA.z, n->BB.x, n->BB.y, n->BB.z});
    if (n->num > 0)
        return;
    if (n->left)
    {
        links[me * 4 + 0] = (int32_t)(links.size() / 4);
        flatten(n->left, boxes, links);
    }
    if (n->right)
    {
        links[me * 4 + 1] = (int32_t)(links.size() / 4);
        flatten(n->right, boxes, links);
    }
}
} // namespace

extern "C"
{
// Loads with the reference's own loaders in the reference's order (main.cpp:66-69) and builds the BVH
// (main.cpp:76).  Returns NULL never: the reference loaders exit() on failure (scene.cpp:7-11,61-65).
void *ref_scene_load(const char *xml, const char *obj, const char *mtl, const char *basedir, int build)
{
    RefScene *h = new RefScene();
    h->scene.readxml(xml);
    h->scene.readobj(obj);
    h->scene.readmtl(mtl, basedir);
    if (build)
        h->root = buildBVH(h->scene.triangles, 0, (int)h->scene.triangles.size() - 1, 8);
    const auto &tris = h->scene.triangles;
    h->canon_of.resize(tris.size());
    for (size_t i = 0; i < tris.size(); i++)
    {
        if (!h->mtl_index.count(tris[i].mtl_name))
        {
            h->mtl_index[tris[i].mtl_name] = (int)h->mtl_names.size();
            h->mtl_names.push_back(tris[i].mtl_name);
        }
        auto it = h->canon.emplace(tri_key(tris[i]), (int)i).first;
        h->canon_of[i] = it->second;
    }
    return h;
}

int ref_num_triangles(void *hp) { return (int)static_cast<RefScene *>(hp)->scene.triangles.size(); }
int ref_num_materials(void *hp) { return (int)static_cast<RefScene *>(hp)->mtl_names.size(); }
const char *ref_material_name(void *hp, int i) { return static_cast<RefScene *>(hp)->mtl_names[i].c_str(); }
int ref_image_size(void *hp, int *w, int *h)
{
    RefScene *s = static_cast<RefScene *>(hp);
    *w = s->scene.img_width;
    *h = s->scene.img_height;
    return 0;
}

// Triangles in their current (post-build if built) order. Any pointer may be NULL.
void ref_get_triangles(void *hp, float *v9, float *vn9, float *vt6, float *normal3, float *center3, double *area,
                       int32_t *emissive, int32_t *mtl, int32_t *canon)
{
    RefScene *h = static_cast<RefScene *>(hp);
    const auto &tris = h->scene.triangles;
    for (size_t i = 0; i < tris.size(); i++)
    {
        const Triangle &t = tris[i];
        if (v9)
            memcpy(v9 + 9 * i, t.v, 36);
        if (vn9)
            memcpy(vn9 + 9 * i, t.vn, 36);
        if (vt6)
            memcpy(vt6 + 6 * i, t.vt, 24);
        if (normal3)
            memcpy(normal3 + 3 * i, &t.normal, 12);
        if (center3)
            memcpy(center3 + 3 * i, &t.center, 12);
        if (area)
            area[i] = t.area;
        if (emissive)
            emissive[i] = t.is_emissive ? 1 : 0;
        if (mtl)
            mtl[i] = h->mtl_index[t.mtl_name];
        if (canon)
            canon[i] = h->canon_of[i];
    }
}

// Material record of material i (index into ref_material_name order).
// out16 = Kd(3) Ks(3) Tr(3) Ns Ni radiance(3) is_emissive has_texture ; area_out = Material::area
void ref_get_material(void *hp, int i, float *out16, double *area_out, int32_t *n_light_tris)
{
    RefScene *h = static_cast<RefScene *>(hp);
    Material &m = h->scene.materials[h->mtl_names[i]];
    float o[16] = {m.Kd.x, m.Kd.y, m.Kd.z, m.Ks.x, m.Ks.y, m.Ks.z, m.Tr.x, m.Tr.y,
                   m.Tr.z, m.Ns, m.Ni, m.radiance.x, m.radiance.y, m.radiance.z, m.is_emissive ? 1.f : 0.f,
                   m.map_Kd != "" ? 1.f : 0.f};
    memcpy(out16, o, sizeof o);
    *area_out = m.area;
    *n_light_tris = (int32_t)m.triangles.size();
}

// camera vectors as computed by the reference's setCamera (camera.cpp:3-17): eye, llc, horizontal, vertical
void ref_get_camera(void *hp, float *out12)
{
    Camera &c = static_cast<RefScene *>(hp)->scene.camera;
    float o[12] = {c.eye.x, c.eye.y, c.eye.z, c.lower_left_corner.x, c.lower_left_corner.y, c.lower_left_corner.z,
                   c.horizontal.x, c.horizontal.y, c.horizontal.z, c.vertical.x, c.vertical.y, c.vertical.z};
    memcpy(out12, o, sizeof o);
}

// BVH shape: node count, leaf count, max depth (root = depth 0).
void ref_bvh_stats(void *hp, int32_t *nodes, int32_t *leaves, int32_t *maxdepth)
{
    int n = 0, l = 0, d = 0;
    node_stats(static_cast<RefScene *>(hp)->root, 0, n, l, d);
    *nodes = n, *leaves = l, *maxdepth = d;
}

// Pre-order dump of the reference tree: boxes[6*n], links[4*n] = {left, right, index, num}. Returns n.
int ref_bvh_flatten(void *hp, float *boxes, int32_t *links, int cap)
{
    RefScene *h = static_cast<RefScene *>(hp);
    std::vector<float> b;
    std::vector<int32_t> l;
    if (h->root)
        flatten(h->root, b, l);
    int n = (int)(l.size() / 4);
    if (boxes && links && n <= cap)
    {
        memcpy(boxes, b.data(), b.size() * 4);
        memcpy(links, l.data(), l.size() * 4);
    }
    return n;
}

// The reference's own traverseBVH on every ray (org xyz, dir xyz). Outputs per ray: distance (INF=114514 on
// miss), canonical post-build triangle index (-1 on miss), optional shading normal pn and hit point.
// threads<=0 -> omp default. Returns the number of hits.
long ref_trace(void *hp, const float *rays6, long n, float *t_out, int32_t *id_out, float *pn3, float *hitp3,
               int threads)
{
    RefScene *h = static_cast<RefScene *>(hp);
    long hits = 0;
    if (threads > 0)
        omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : hits)
    for (long i = 0; i < n; i++)
    {
        Ray ray(vec3(rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2]),
                vec3(rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5]));
        HitRecord rec = traverseBVH(ray, h->scene.triangles, h->root);
        if (t_out)
            t_out[i] = rec.distance;
        int id = -1;
        if (rec.is_hit)
        {
            hits++;
            auto it = h->canon.find(tri_key(rec.triangle));
            id = (it == h->canon.end()) ? -2 : it->second;
        }
        if (id_out)
            id_out[i] = id;
        if (pn3)
            memcpy(pn3 + 3 * i, &rec.pn, 12);
        if (hitp3)
            memcpy(hitp3 + 3 * i, &rec.hitpoint, 12);
    }
    return hits;
}

// The reference's render loop body (main.cpp:79-113) driven from here so that the LINEAR double image is
// available (the reference binary only writes the gamma-quantised PNG).  Pixel mapping, getRay,
// traverseBVH, shade and the /SAMPLE accumulate are the reference's; the jitter engine is a local
// default_random_engine like main.cpp:57-58.  shade()'s internal static engines stay racy/time-seeded,
// exactly as in the reference, so the result is only statistically reproducible.
void ref_render(void *hp, int spp, double *image, int threads, unsigned seed)
{
    RefScene *h = static_cast<RefScene *>(hp);
    Scene &scene = h->scene;
    const int W = scene.img_width, H = scene.img_height;
    memset(image, 0, sizeof(double) * W * H * 3);
    if (threads > 0)
        omp_set_num_threads(threads);
#pragma omp parallel
    {
        std::default_random_engine e(seed + 7919u * omp_get_thread_num());
        std::uniform_real_distribution<double> u1(0, 1);
        std::vector<double> local((size_t)W * H * 3, 0.0);
#pragma omp for schedule(dynamic, 1)
        for (int k = 0; k < spp; k++)
        {
            double *p = local.data();
            for (int i = 0; i < H; i++)
                for (int j = 0; j < W; j++)
                {
                    double x = double(j) / double(W - 1.0);
                    double y = double(H - i) / double(H - 1.0);
                    x += (u1(e) - 0.5f) / double(W);
                    y += (u1(e) - 0.5f) / double(H);
                    Ray ray = scene.camera.getRay(x, y);
                    HitRecord rec = traverseBVH(ray, scene.triangles, h->root);
                    vec3 color = vec3(0);
                    if (rec.is_hit)
                        color = shade(rec, -ray.direction, scene, h->root) / (float)spp;
                    *p++ += color.x;
                    *p++ += color.y;
                    *p++ += color.z;
                }
        }
#pragma omp critical
        for (size_t q = 0; q < local.size(); q++)
            image[q] += local[q];
    }
}

int ref_num_threads() { return omp_get_max_threads(); }
}
