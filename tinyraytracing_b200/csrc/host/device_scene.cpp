// Bridge from the host classes to the C ABI (include/trt.h): POD conversion after buildBVH, and the GPU-backed
// replacements for the reference's traverseBVH (bvh.cpp:146-175) and sample loop (main.cpp:79-113).
#include "tinyrt.h"

#include <algorithm>
#include <cstring>
#include <functional>

namespace trt
{
std::unique_ptr<SceneArrays> makeSceneArrays(Scene &scene, const BVHNode *root)
{
    std::unique_ptr<SceneArrays> a(new SceneArrays());
    const size_t n = scene.triangles.size();

    // material table: lights first in XML order, then the remaining names in first-use order over the
    // post-build triangles, then whatever else the MTL defined (sorted, so the table is deterministic)
    std::unordered_map<std::string, int> mindex;
    auto intern = [&](const std::string &name) {
        auto it = mindex.find(name);
        if (it != mindex.end())
            return it->second;
        int id = (int)a->material_names.size();
        mindex[name] = id;
        a->material_names.push_back(name);
        return id;
    };
    for (auto &l : scene.lights)
        intern(l.mtl_name);
    for (auto &t : scene.triangles)
        intern(t.mtl_name);
    {
        std::vector<std::string> rest;
        for (auto &kv : scene.materials)
            if (!mindex.count(kv.first))
                rest.push_back(kv.first);
        std::sort(rest.begin(), rest.end());
        for (auto &s : rest)
            intern(s);
    }

    a->v.resize(n * 9), a->vn.resize(n * 9), a->vt.resize(n * 6), a->normal.resize(n * 3), a->mtl.resize(n), a->face.resize(n);
    for (size_t i = 0; i < n; ++i)
    {
        const Triangle &t = scene.triangles[i];
        for (int k = 0; k < 3; ++k)
        {
            a->v[i * 9 + k * 3 + 0] = t.v[k].x, a->v[i * 9 + k * 3 + 1] = t.v[k].y, a->v[i * 9 + k * 3 + 2] = t.v[k].z;
            a->vn[i * 9 + k * 3 + 0] = t.vn[k].x, a->vn[i * 9 + k * 3 + 1] = t.vn[k].y, a->vn[i * 9 + k * 3 + 2] = t.vn[k].z;
            a->vt[i * 6 + k * 2 + 0] = t.vt[k].x, a->vt[i * 6 + k * 2 + 1] = t.vt[k].y;
        }
        a->normal[i * 3 + 0] = t.normal.x, a->normal[i * 3 + 1] = t.normal.y, a->normal[i * 3 + 2] = t.normal.z;
        a->mtl[i] = mindex[t.mtl_name];
        a->face[i] = t.face;
    }

    // pre-order flattening of the pointer tree
    std::function<int(const BVHNode *)> flat = [&](const BVHNode *nd) -> int {
        const int me = (int)(a->node_link.size() / 4);
        a->node_link.insert(a->node_link.end(), {-1, -1, nd->index, nd->num});
        a->node_box.insert(a->node_box.end(), {nd->AA.x, nd->AA.y, nd->AA.z, nd->BB.x, nd->BB.y, nd->BB.z});
        if (nd->num == 0)
        {
            if (nd->left)
            {
                int c = flat(nd->left);
                a->node_link[(size_t)me * 4 + 0] = c;
            }
            if (nd->right)
            {
                int c = flat(nd->right);
                a->node_link[(size_t)me * 4 + 1] = c;
            }
        }
        return me;
    };
    if (root)
        flat(root);

    for (auto &name : a->material_names)
    {
        const Material &m = scene.materials[name];
        trt_material o;
        std::memset(&o, 0, sizeof o);
        o.Kd[0] = m.Kd.x, o.Kd[1] = m.Kd.y, o.Kd[2] = m.Kd.z;
        o.Ks[0] = m.Ks.x, o.Ks[1] = m.Ks.y, o.Ks[2] = m.Ks.z;
        o.Tr[0] = m.Tr.x, o.Tr[1] = m.Tr.y, o.Tr[2] = m.Tr.z;
        o.Ns = m.Ns, o.Ni = m.Ni;
        o.radiance[0] = m.radiance.x, o.radiance[1] = m.radiance.y, o.radiance[2] = m.radiance.z;
        o.is_emissive = m.is_emissive ? 1 : 0;
        o.texture = -1;
        o.area = m.area;
        if (m.map_Kd != "")
        {
            // shade() indexes m.img whenever map_Kd != "" (pathTracing.cpp:17-25); an unreadable texture is
            // undefined behaviour there — here it becomes a 1x1 black texture
            o.texture = (int)a->textures.size();
            Image im = m.img;
            if (im.empty())
            {
                im.rows = im.cols = 1;
                im.data = std::make_shared<std::vector<unsigned char>>(3, 0);
            }
            a->texture_images.push_back(im);
            a->textures.push_back(trt_texture{im.rows, im.cols, im.data->data()});
        }
        a->materials.push_back(o);
    }

    for (auto &l : scene.lights)
    {
        const Material &m = scene.materials[l.mtl_name];
        trt_light o;
        o.material = mindex[l.mtl_name];
        o.first_tri = (int)a->light_cum_area.size();
        o.n_tris = (int)m.triangles.size();
        o._pad = 0;
        for (const Triangle &t : m.triangles)
        {
            for (int k = 0; k < 3; ++k)
            {
                a->light_v.insert(a->light_v.end(), {t.v[k].x, t.v[k].y, t.v[k].z});
                a->light_vn.insert(a->light_vn.end(), {t.vn[k].x, t.vn[k].y, t.vn[k].z});
            }
            a->light_cum_area.push_back(t.area);
        }
        a->lights.push_back(o);
    }

    trt_scene_desc &d = a->desc;
    std::memset(&d, 0, sizeof d);
    d.n_tris = (int)n;
    d.v = a->v.data(), d.vn = a->vn.data(), d.vt = a->vt.data(), d.normal = a->normal.data(), d.mtl = a->mtl.data();
    d.n_nodes = (int)(a->node_link.size() / 4);
    d.node_box = a->node_box.data(), d.node_link = a->node_link.data();
    d.n_materials = (int)a->materials.size(), d.materials = a->materials.data();
    d.n_lights = (int)a->lights.size(), d.lights = a->lights.data();
    d.n_light_tris = (int)a->light_cum_area.size();
    d.light_v = a->light_v.data(), d.light_vn = a->light_vn.data(), d.light_cum_area = a->light_cum_area.data();
    d.n_textures = (int)a->textures.size(), d.textures = a->textures.data();
    const Camera &c = scene.camera;
    const vec3 cv[4] = {c.eye, c.lower_left_corner, c.horizontal, c.vertical};
    float *dst[4] = {d.eye, d.lower_left_corner, d.horizontal, d.vertical};
    for (int k = 0; k < 4; ++k)
        dst[k][0] = cv[k].x, dst[k][1] = cv[k].y, dst[k][2] = cv[k].z;
    d.width = scene.img_width, d.height = scene.img_height;
    return a;
}

DeviceScene::DeviceScene(Scene &scene, BVHNode *root, int device, const char *layout_cache) : scene_(&scene)
{
    auto arrays = makeSceneArrays(scene, root);
    const int rc = layout_cache ? trt_scene_create_cached(&arrays->desc, device, layout_cache, nullptr, &h_)
                                : trt_scene_create(&arrays->desc, device, &h_);
    if (rc != TRT_OK)
        throw std::runtime_error(std::string("trt_scene_create: ") + trt_last_error());
}

DeviceScene::DeviceScene(const DeviceScene &src, int device, bool) : scene_(src.scene_)
{
    if (trt_scene_replicate(src.h_, device, &h_) != TRT_OK)
        throw std::runtime_error(std::string("trt_scene_replicate: ") + trt_last_error());
}

DeviceScene::~DeviceScene() { trt_scene_destroy(h_); }

std::vector<HitRecord> traverseBVH(const std::vector<Ray> &rays, DeviceScene &dev)
{
    const size_t n = rays.size();
    std::vector<float> r6(n * 6), t(n), hp(n * 3), pn(n * 3);
    std::vector<int32_t> id(n);
    for (size_t i = 0; i < n; ++i)
    {
        const Ray &r = rays[i];
        float *p = &r6[i * 6];
        p[0] = r.startpoint.x, p[1] = r.startpoint.y, p[2] = r.startpoint.z;
        p[3] = r.direction.x, p[4] = r.direction.y, p[5] = r.direction.z;
    }
    if (trt_trace_closest(dev.handle(), r6.data(), n, id.data(), t.data(), 0) != TRT_OK ||
        trt_hit_attributes(dev.handle(), r6.data(), id.data(), t.data(), n, hp.data(), pn.data()) != TRT_OK)
        throw std::runtime_error(std::string("traverseBVH: ") + trt_last_error());
    std::vector<HitRecord> out(n);
    for (size_t i = 0; i < n; ++i)
    {
        if (id[i] < 0)
            continue;
        HitRecord &h = out[i];
        h.is_hit = true;
        h.distance = t[i];
        h.hitpoint = vec3(hp[i * 3], hp[i * 3 + 1], hp[i * 3 + 2]);
        h.direction = rays[i].direction;
        h.pn = vec3(pn[i * 3], pn[i * 3 + 1], pn[i * 3 + 2]);
        h.triangle = dev.scene().triangles[id[i]];
        h.triangle_index = id[i];
        h.startpoint = rays[i].startpoint;
    }
    return out;
}

HitRecord traverseBVH(Ray ray, DeviceScene &dev) { return traverseBVH(std::vector<Ray>{ray}, dev)[0]; }

void renderImage(DeviceScene &dev, int spp, double *image, uint64_t seed, int max_depth)
{
    trt_render_params p;
    std::memset(&p, 0, sizeof p);
    p.spp = spp, p.sample_begin = 0, p.sample_end = spp, p.max_depth = max_depth, p.seed = seed;
    if (trt_render(dev.handle(), &p, image) != TRT_OK)
        throw std::runtime_error(std::string("trt_render: ") + trt_last_error());
}

void renderImage(const std::vector<DeviceScene *> &devs, int spp, double *image, uint64_t seed, int max_depth, uint32_t flags)
{
    std::vector<trt_scene *> handles;
    for (DeviceScene *d : devs)
        handles.push_back(d->handle());
    trt_render_params p;
    std::memset(&p, 0, sizeof p);
    p.spp = spp, p.sample_begin = 0, p.sample_end = spp, p.max_depth = max_depth, p.seed = seed, p.flags = flags;
    if (trt_render_multi(handles.data(), (int32_t)handles.size(), &p, image, nullptr) != TRT_OK)
        throw std::runtime_error(std::string("trt_render_multi: ") + trt_last_error());
}

int renderImageCheckpointed(DeviceScene &dev, int spp, double *image, const std::string &checkpoint, int every, uint64_t seed,
                            int max_depth)
{
    trt_scene *h = dev.handle();
    double *acc = trt_accum_create(h);
    if (!acc)
        throw std::runtime_error(std::string("trt_accum_create: ") + trt_last_error());
    int done = 0, rendered = 0;
    {
        int32_t c_done = 0, c_spp = 0, c_depth = 0;
        uint64_t c_seed = 0;
        // a missing / foreign / damaged file simply means "start from sample 0" (and is overwritten below)
        if (trt_accum_load(h, checkpoint.c_str(), acc, &c_done, &c_spp, &c_seed, &c_depth) == TRT_OK && c_spp == spp &&
            c_seed == seed && c_depth == max_depth)
            done = c_done;
        else
        {
            trt_accum_destroy(h, acc);
            if (!(acc = trt_accum_create(h)))
                throw std::runtime_error(std::string("trt_accum_create: ") + trt_last_error());
        }
    }
    every = std::max(1, every);
    trt_render_params p;
    std::memset(&p, 0, sizeof p);
    p.spp = spp, p.max_depth = max_depth, p.seed = seed;
    int rc = TRT_OK;
    while (done < spp && rc == TRT_OK)
    {
        p.sample_begin = done, p.sample_end = std::min(spp, done + every);
        if ((rc = trt_render_accumulate(h, &p, acc, nullptr)) != TRT_OK)
            break;
        rendered += p.sample_end - done;
        done = p.sample_end;
        rc = trt_accum_save(h, acc, done, spp, seed, max_depth, checkpoint.c_str());
    }
    if (rc == TRT_OK)
        rc = trt_resolve(h, acc, spp, image, nullptr, nullptr);
    const std::string err = rc == TRT_OK ? "" : trt_last_error();
    trt_accum_destroy(h, acc);
    if (rc != TRT_OK)
        throw std::runtime_error("renderImageCheckpointed: " + err);
    return rendered;
}

std::vector<vec3> shade(const std::vector<HitRecord> &records, DeviceScene &dev, uint64_t seed, int sample, int max_depth)
{
    const size_t n = records.size();
    std::vector<float> r6(n * 6), t(n), rad(n * 3);
    std::vector<int32_t> id(n);
    for (size_t i = 0; i < n; ++i)
    {
        const HitRecord &h = records[i];
        float *p = &r6[i * 6];
        p[0] = h.startpoint.x, p[1] = h.startpoint.y, p[2] = h.startpoint.z;
        p[3] = h.direction.x, p[4] = h.direction.y, p[5] = h.direction.z;
        id[i] = h.is_hit ? h.triangle_index : -1;
        t[i] = h.distance;
    }
    trt_shade_params sp;
    std::memset(&sp, 0, sizeof sp);
    sp.seed = seed, sp.sample = sample, sp.max_depth = max_depth;
    if (trt_shade(dev.handle(), r6.data(), id.data(), t.data(), n, &sp, rad.data()) != TRT_OK)
        throw std::runtime_error(std::string("shade: ") + trt_last_error());
    std::vector<vec3> out(n);
    for (size_t i = 0; i < n; ++i)
        out[i] = vec3(rad[i * 3], rad[i * 3 + 1], rad[i * 3 + 2]);
    return out;
}

vec3 shade(HitRecord &res, vec3 dir, DeviceScene &dev, uint64_t seed, int sample, int max_depth)
{
    HitRecord r = res;
    r.direction = vec3(-dir.x, -dir.y, -dir.z); // the reference passes wi = -ray.direction (main.cpp:101)
    return shade(std::vector<HitRecord>{r}, dev, seed, sample, max_depth)[0];
}

// triangle.cpp:12-29: least-squares solution of [v0 v1 v2; 1 1 1] b = [p; 1] in double, by the reference's own route
// (column-pivoted Householder QR).  The device computes the same solution in closed form (csrc/barycentric.cuh).
vec3 Triangle::findBaryCor(vec3 p)
{
    double A[4][3] = {{v[0].x, v[1].x, v[2].x}, {v[0].y, v[1].y, v[2].y}, {v[0].z, v[1].z, v[2].z}, {1, 1, 1}};
    double b[4] = {p.x, p.y, p.z, 1};
    int perm[3] = {0, 1, 2};
    for (int k = 0; k < 3; ++k)
    {
        int best = k;
        double bn = -1;
        for (int j = k; j < 3; ++j)
        {
            double s = 0;
            for (int i = k; i < 4; ++i)
                s += A[i][j] * A[i][j];
            if (s > bn)
                bn = s, best = j;
        }
        if (best != k)
        {
            for (int i = 0; i < 4; ++i)
                std::swap(A[i][k], A[i][best]);
            std::swap(perm[k], perm[best]);
        }
        double norm = std::sqrt(bn);
        if (norm == 0)
            continue;
        double alpha = (A[k][k] > 0) ? -norm : norm;
        double w[4] = {0, 0, 0, 0};
        for (int i = k; i < 4; ++i)
            w[i] = A[i][k];
        w[k] -= alpha;
        double wtw = 0;
        for (int i = k; i < 4; ++i)
            wtw += w[i] * w[i];
        if (wtw == 0)
            continue;
        const double beta = 2 / wtw;
        for (int j = k; j < 3; ++j)
        {
            double s = 0;
            for (int i = k; i < 4; ++i)
                s += w[i] * A[i][j];
            s *= beta;
            for (int i = k; i < 4; ++i)
                A[i][j] -= s * w[i];
        }
        double s = 0;
        for (int i = k; i < 4; ++i)
            s += w[i] * b[i];
        s *= beta;
        for (int i = k; i < 4; ++i)
            b[i] -= s * w[i];
    }
    double y[3];
    for (int k = 2; k >= 0; --k)
    {
        double s = b[k];
        for (int j = k + 1; j < 3; ++j)
            s -= A[k][j] * y[j];
        y[k] = (A[k][k] != 0) ? s / A[k][k] : 0;
    }
    double r[3];
    for (int k = 0; k < 3; ++k)
        r[perm[k]] = y[k];
    return vec3((float)r[0], (float)r[1], (float)r[2]);
}
} // namespace trt
