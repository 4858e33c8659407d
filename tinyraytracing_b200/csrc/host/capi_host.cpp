// C view of the host side (include/trt_host.h) for non-C++ bindings.
#include "tinyrt.h"
#include "png_writer.h"
#include "../../../include/trt_host.h"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

namespace trt
{
void setLastError(const std::string &s); // capi.cu
}

struct trt_host_scene
{
    trt::Scene scene;
    trt::BVHNode *root = nullptr;
    std::unique_ptr<trt::SceneArrays> arrays;
    double build_s = 0;
    ~trt_host_scene() { trt::freeBVH(root); }
};

namespace
{
int finish(std::unique_ptr<trt_host_scene> &s, int leaf_num, trt_host_scene **out)
{
    auto t0 = std::chrono::steady_clock::now();
    s->root = trt::buildBVH(s->scene.triangles, 0, (int)s->scene.triangles.size() - 1, leaf_num);
    s->build_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    s->arrays = trt::makeSceneArrays(s->scene, s->root);
    *out = s.release();
    return TRT_OK;
}
} // namespace

// ---- binary scene cache ------------------------------------------------------------------------------------------
// Layout: 8-byte magic, u32 version, u32 sizeof(trt_material), then a sequence of length-prefixed sections (u64 byte
// count + raw little-endian data) in a fixed order, then the FNV-1a 64 hash of everything before it.
namespace
{
const char kCacheMagic[8] = {'T', 'R', 'T', 'S', 'C', 'N', '0', '1'};
const uint32_t kCacheVersion = 1;

// FNV-1a over 64-bit words (bytes for the tail) of one chunk; the file hash chains the chunk hashes in file order
uint64_t chunkHash(const void *p, size_t n)
{
    const unsigned char *b = static_cast<const unsigned char *>(p);
    uint64_t h = 1469598103934665603ull;
    size_t i = 0;
    for (; i + 8 <= n; i += 8)
    {
        uint64_t w;
        std::memcpy(&w, b + i, 8);
        h = (h ^ w) * 1099511628211ull;
    }
    for (; i < n; ++i)
        h = (h ^ b[i]) * 1099511628211ull;
    return h ^ (uint64_t)n;
}

struct CacheWriter
{
    FILE *fp;
    uint64_t hash = 1469598103934665603ull;
    bool ok = true;
    void raw(const void *p, size_t n)
    {
        hash = (hash ^ chunkHash(p, n)) * 1099511628211ull;
        ok = ok && (n == 0 || std::fwrite(p, 1, n, fp) == n);
    }
    void section(const void *p, size_t n)
    {
        const uint64_t len = n;
        raw(&len, 8);
        raw(p, n);
    }
    template <typename T>
    void vec(const std::vector<T> &v) { section(v.data(), v.size() * sizeof(T)); }
};

struct CacheReader // hashes what it reads, chunk by chunk, exactly as CacheWriter::raw did
{
    const unsigned char *p, *end;
    uint64_t hash = 1469598103934665603ull;
    bool take(void *dst, size_t n)
    {
        if ((size_t)(end - p) < n)
            return false;
        std::memcpy(dst, p, n);
        hash = (hash ^ chunkHash(p, n)) * 1099511628211ull;
        p += n;
        return true;
    }
    template <typename T>
    bool vec(std::vector<T> &v)
    {
        uint64_t len = 0;
        if (!take(&len, 8) || len % sizeof(T) != 0 || (uint64_t)(end - p) < len)
            return false;
        v.resize(len / sizeof(T));
        if (len)
            std::memcpy(v.data(), p, len);
        hash = (hash ^ chunkHash(p, len)) * 1099511628211ull;
        p += len;
        return true;
    }
};
} // namespace

extern "C"
{
int trt_host_scene_load(const char *xml, const char *obj, const char *mtl, const char *basedir, int leaf_num,
                        trt_host_scene **out)
{
    if (!xml || !obj || !mtl || !basedir || !out || leaf_num < 1)
    {
        trt::setLastError("trt_host_scene_load: bad argument");
        return TRT_ERR_INVALID;
    }
    try
    {
        std::unique_ptr<trt_host_scene> s(new trt_host_scene());
        s->scene.readxml(xml); // the order cannot be changed (main.cpp:66-69): readobj needs is_emissive
        s->scene.readobj(obj);
        s->scene.readmtl(mtl, basedir);
        return finish(s, leaf_num, out);
    }
    catch (const std::exception &e)
    {
        trt::setLastError(e.what());
        return TRT_ERR_INVALID;
    }
}

int trt_host_scene_from_arrays(int32_t n, const float *v9, const float *vn9, const float *vt6, const int32_t *mtl,
                               int32_t n_materials, const trt_material *materials, int32_t n_lights,
                               const int32_t *light_materials, const float *light_radiance3, const float *eye,
                               const float *lookat, const float *up, double fovy, int32_t width, int32_t height,
                               int leaf_num, trt_host_scene **out)
{
    if (n < 0 || !v9 || !mtl || n_materials < 1 || !materials || !eye || !lookat || !up || width < 2 || height < 2 ||
        leaf_num < 1 || !out)
    {
        trt::setLastError("trt_host_scene_from_arrays: bad argument");
        return TRT_ERR_INVALID;
    }
    try
    {
        using namespace trt;
        std::unique_ptr<trt_host_scene> s(new trt_host_scene());
        Scene &sc = s->scene;
        sc.img_width = width, sc.img_height = height;
        sc.camera.aspect_ratio = (double)width / (double)height;
        sc.camera.fovy = fovy;
        sc.camera.eye = vec3(eye[0], eye[1], eye[2]);
        sc.camera.lookat = vec3(lookat[0], lookat[1], lookat[2]);
        sc.camera.up = vec3(up[0], up[1], up[2]);
        sc.camera.setCamera();
        auto mname = [](int i) { return "m" + std::to_string(i); };
        for (int i = 0; i < n_materials; ++i)
        {
            Material &m = sc.materials[mname(i)];
            const trt_material &o = materials[i];
            m.Kd = vec3(o.Kd[0], o.Kd[1], o.Kd[2]), m.Ks = vec3(o.Ks[0], o.Ks[1], o.Ks[2]);
            m.Tr = vec3(o.Tr[0], o.Tr[1], o.Tr[2]), m.Ns = o.Ns, m.Ni = o.Ni;
        }
        for (int l = 0; l < n_lights; ++l)
        {
            if (light_materials[l] < 0 || light_materials[l] >= n_materials)
                throw LoadError("light material out of range");
            vec3 r(light_radiance3[l * 3], light_radiance3[l * 3 + 1], light_radiance3[l * 3 + 2]);
            sc.lights.push_back(Light(mname(light_materials[l]), r));
            sc.materials[mname(light_materials[l])].is_emissive = true;
            sc.materials[mname(light_materials[l])].radiance = r;
        }
        sc.triangles.reserve(n);
        for (int i = 0; i < n; ++i)
        {
            if (mtl[i] < 0 || mtl[i] >= n_materials)
                throw LoadError("triangle material out of range");
            Triangle t;
            for (int k = 0; k < 3; ++k)
            {
                t.v[k] = vec3(v9[i * 9 + k * 3], v9[i * 9 + k * 3 + 1], v9[i * 9 + k * 3 + 2]);
                if (vn9)
                    t.vn[k] = vec3(vn9[i * 9 + k * 3], vn9[i * 9 + k * 3 + 1], vn9[i * 9 + k * 3 + 2]);
                if (vt6)
                    t.vt[k] = vec2(vt6[i * 6 + k * 2], vt6[i * 6 + k * 2 + 1]);
            }
            t.normal = normalize(cross(t.v[1] - t.v[0], t.v[2] - t.v[0]));
            t.center = (t.v[0] + t.v[1] + t.v[2]) / vec3(3.0f);
            t.mtl_name = mname(mtl[i]);
            t.face = i;
            Material &m = sc.materials[t.mtl_name];
            if (m.is_emissive)
            {
                t.is_emissive = true;
                m.area += t.calAera();
                t.area = m.area;
                m.triangles.push_back(t);
            }
            sc.triangles.push_back(std::move(t));
        }
        return finish(s, leaf_num, out);
    }
    catch (const std::exception &e)
    {
        trt::setLastError(e.what());
        return TRT_ERR_INVALID;
    }
}

int trt_host_scene_save(trt_host_scene *s, const char *path)
{
    if (!s || !path)
    {
        trt::setLastError("trt_host_scene_save: null argument");
        return TRT_ERR_INVALID;
    }
    FILE *fp = std::fopen(path, "wb");
    if (!fp)
    {
        trt::setLastError(std::string("trt_host_scene_save: cannot open ") + path);
        return TRT_ERR_INVALID;
    }
    const trt::SceneArrays &a = *s->arrays;
    CacheWriter w{fp};
    const uint32_t head[2] = {kCacheVersion, (uint32_t)sizeof(trt_material)};
    w.raw(kCacheMagic, 8);
    w.raw(head, 8);
    w.vec(a.v), w.vec(a.vn), w.vec(a.vt), w.vec(a.normal), w.vec(a.mtl), w.vec(a.face);
    w.vec(a.node_box), w.vec(a.node_link);
    w.vec(a.materials), w.vec(a.lights);
    w.vec(a.light_v), w.vec(a.light_vn), w.vec(a.light_cum_area);
    std::vector<int32_t> meta = {(int32_t)a.material_names.size(), (int32_t)a.texture_images.size(), a.desc.width, a.desc.height};
    w.vec(meta);
    for (const std::string &n : a.material_names)
        w.section(n.data(), n.size());
    for (const trt::Image &im : a.texture_images)
    {
        const int32_t rc[2] = {im.rows, im.cols};
        w.section(rc, 8);
        w.section(im.data->data(), im.data->size());
    }
    float cam[12];
    std::memcpy(cam, a.desc.eye, 12), std::memcpy(cam + 3, a.desc.lower_left_corner, 12);
    std::memcpy(cam + 6, a.desc.horizontal, 12), std::memcpy(cam + 9, a.desc.vertical, 12);
    w.section(cam, sizeof cam);
    const uint64_t h = w.hash;
    const bool ok = w.ok && std::fwrite(&h, 1, 8, fp) == 8;
    if (std::fclose(fp) != 0 || !ok)
    {
        trt::setLastError(std::string("trt_host_scene_save: write failed: ") + path);
        return TRT_ERR_INVALID;
    }
    return TRT_OK;
}

int trt_host_scene_load_cache(const char *path, trt_host_scene **out)
{
    if (!path || !out)
    {
        trt::setLastError("trt_host_scene_load_cache: null argument");
        return TRT_ERR_INVALID;
    }
    auto fail = [&](const std::string &why) {
        trt::setLastError("trt_host_scene_load_cache: " + why + ": " + path);
        return TRT_ERR_INVALID;
    };
    FILE *fp = std::fopen(path, "rb");
    if (!fp)
        return fail("cannot open");
    std::vector<unsigned char> buf;
    std::fseek(fp, 0, SEEK_END);
    const long size = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    if (size < 24)
    {
        std::fclose(fp);
        return fail("not a scene cache");
    }
    buf.resize((size_t)size);
    const bool rd = std::fread(buf.data(), 1, buf.size(), fp) == buf.size();
    std::fclose(fp);
    if (!rd)
        return fail("read failed");
    if (std::memcmp(buf.data(), kCacheMagic, 8) != 0)
        return fail("not a scene cache");
    uint64_t stored = 0;
    std::memcpy(&stored, buf.data() + buf.size() - 8, 8);
    CacheReader r{buf.data(), buf.data() + buf.size() - 8};
    char magic[8];
    uint32_t head[2] = {0, 0};
    if (!r.take(magic, 8) || !r.take(head, 8) || head[0] != kCacheVersion || head[1] != sizeof(trt_material))
        return fail("unsupported cache version");
    try
    {
        std::unique_ptr<trt_host_scene> s(new trt_host_scene());
        s->arrays.reset(new trt::SceneArrays());
        trt::SceneArrays &a = *s->arrays;
        std::vector<int32_t> meta;
        bool ok = r.vec(a.v) && r.vec(a.vn) && r.vec(a.vt) && r.vec(a.normal) && r.vec(a.mtl) && r.vec(a.face) &&
                  r.vec(a.node_box) && r.vec(a.node_link) && r.vec(a.materials) && r.vec(a.lights) && r.vec(a.light_v) &&
                  r.vec(a.light_vn) && r.vec(a.light_cum_area) && r.vec(meta) && meta.size() == 4 && meta[0] >= 0 && meta[1] >= 0;
        for (int i = 0; ok && i < meta[0]; ++i)
        {
            std::vector<char> name;
            ok = r.vec(name);
            a.material_names.emplace_back(name.begin(), name.end());
        }
        for (int i = 0; ok && i < meta[1]; ++i)
        {
            std::vector<int32_t> rc;
            trt::Image im;
            im.data = std::make_shared<std::vector<unsigned char>>();
            ok = r.vec(rc) && rc.size() == 2 && r.vec(*im.data) && rc[0] >= 1 && rc[1] >= 1 &&
                 im.data->size() == (size_t)rc[0] * rc[1] * 3;
            if (ok)
            {
                im.rows = rc[0], im.cols = rc[1];
                a.texture_images.push_back(im);
                a.textures.push_back(trt_texture{im.rows, im.cols, im.data->data()});
            }
        }
        std::vector<float> cam;
        ok = ok && r.vec(cam) && cam.size() == 12 && r.p == r.end;
        if (!ok || r.hash != stored) // nothing read so far is used before this check
            return fail("checksum mismatch or malformed file (truncated, corrupted or foreign)");
        // consistency of the sections with each other (a valid checksum only says the file is what was written)
        const size_t n = a.mtl.size();
        ok = ok && a.v.size() == n * 9 && a.vn.size() == n * 9 && a.vt.size() == n * 6 && a.normal.size() == n * 3 &&
             a.face.size() == n && a.node_link.size() % 4 == 0 && a.node_box.size() == a.node_link.size() / 4 * 6 &&
             a.light_v.size() == a.light_cum_area.size() * 9 && a.light_vn.size() == a.light_v.size() &&
             a.material_names.size() == a.materials.size();
        if (!ok)
            return fail("malformed cache");
        trt_scene_desc &d = a.desc;
        std::memset(&d, 0, sizeof d);
        d.n_tris = (int)n;
        d.v = a.v.data(), d.vn = a.vn.data(), d.vt = a.vt.data(), d.normal = a.normal.data(), d.mtl = a.mtl.data();
        d.n_nodes = (int)(a.node_link.size() / 4);
        d.node_box = a.node_box.data(), d.node_link = a.node_link.data();
        d.n_materials = (int)a.materials.size(), d.materials = a.materials.data();
        d.n_lights = (int)a.lights.size(), d.lights = a.lights.data();
        d.n_light_tris = (int)a.light_cum_area.size();
        d.light_v = a.light_v.data(), d.light_vn = a.light_vn.data(), d.light_cum_area = a.light_cum_area.data();
        d.n_textures = (int)a.textures.size(), d.textures = a.textures.data();
        std::memcpy(d.eye, cam.data(), 12), std::memcpy(d.lower_left_corner, cam.data() + 3, 12);
        std::memcpy(d.horizontal, cam.data() + 6, 12), std::memcpy(d.vertical, cam.data() + 9, 12);
        d.width = meta[2], d.height = meta[3];
        *out = s.release();
        return TRT_OK;
    }
    catch (const std::exception &e)
    {
        return fail(e.what());
    }
}

const trt_scene_desc *trt_host_scene_desc(trt_host_scene *s) { return s ? &s->arrays->desc : nullptr; }
const int32_t *trt_host_scene_faces(trt_host_scene *s) { return s ? s->arrays->face.data() : nullptr; }
const char *trt_host_scene_material_name(trt_host_scene *s, int i)
{
    if (!s || i < 0 || i >= (int)s->arrays->material_names.size())
        return nullptr;
    return s->arrays->material_names[i].c_str();
}
double trt_host_scene_build_seconds(trt_host_scene *s) { return s ? s->build_s : 0.0; }
void trt_host_scene_free(trt_host_scene *s) { delete s; }

int trt_decode_jpeg(const char *path, int32_t *rows, int32_t *cols, uint8_t *bgr_out, size_t capacity)
{
    if (!path || !rows || !cols)
    {
        trt::setLastError("trt_decode_jpeg: null argument");
        return TRT_ERR_INVALID;
    }
    trt::Image img;
    std::string why;
    if (!trt::decodeJpegFile(path, img, why))
    {
        trt::setLastError("trt_decode_jpeg: " + why);
        return TRT_ERR_INVALID;
    }
    *rows = img.rows, *cols = img.cols;
    const size_t need = (size_t)img.rows * img.cols * 3;
    if (bgr_out)
    {
        if (capacity < need)
        {
            trt::setLastError("trt_decode_jpeg: output buffer too small");
            return TRT_ERR_LIMIT;
        }
        std::memcpy(bgr_out, img.data->data(), need);
    }
    return TRT_OK;
}

int trt_write_pfm(const char *path, int32_t w, int32_t h, const double *rgb_linear)
{
    if (!path || !rgb_linear || w < 1 || h < 1)
    {
        trt::setLastError("trt_write_pfm: bad argument");
        return TRT_ERR_INVALID;
    }
    FILE *fp = std::fopen(path, "wb");
    if (!fp)
    {
        trt::setLastError(std::string("cannot open ") + path);
        return TRT_ERR_INVALID;
    }
    bool ok = std::fprintf(fp, "PF\n%d %d\n-1.0\n", w, h) > 0; // negative scale = little-endian
    std::vector<float> row((size_t)w * 3);
    for (int y = h - 1; ok && y >= 0; --y) // PFM stores the bottom row first
    {
        const double *src = rgb_linear + (size_t)y * w * 3;
        for (size_t i = 0; i < row.size(); ++i)
            row[i] = (float)src[i];
        ok = std::fwrite(row.data(), sizeof(float), row.size(), fp) == row.size();
    }
    ok = (std::fclose(fp) == 0) && ok;
    if (!ok)
        trt::setLastError(std::string("trt_write_pfm: write failed: ") + path);
    return ok ? TRT_OK : TRT_ERR_INVALID;
}

int trt_write_png(const char *path, int32_t w, int32_t h, const uint8_t *rgb, int alpha)
{
    FILE *fp = std::fopen(path, "wb");
    if (!fp)
    {
        trt::setLastError(std::string("cannot open ") + path);
        return TRT_ERR_INVALID;
    }
    bool ok = trt::writePNG(fp, (unsigned)w, (unsigned)h, rgb, alpha);
    std::fclose(fp);
    return ok ? TRT_OK : TRT_ERR_INVALID;
}
}
