// C view of the host side (include/trt_host.h) for non-C++ bindings.
#include "tinyrt.h"
#include "png_writer.h"
#include "../../../include/trt_host.h"

#include <chrono>
#include <cstring>

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
