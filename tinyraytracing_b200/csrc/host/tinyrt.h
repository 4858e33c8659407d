// Host-side mirror of the reference's interfaces (same class / member / function names and meaning) for the
// part of TinyRayTracing that stays on the CPU: Scene / Camera / Material / Light / Triangle / Ray /
// HitRecord / BVHNode, the XML / MTL / OBJ loaders and buildBVH.  The hot path (traverseBVH, shade, the
// sample loop) is NOT implemented here: those entry points forward to the sm_100a kernels through the C ABI
// of include/trt.h and fail when no GPU is present.
//   reference: ray.h:5-21, triangle.h:9-26, camera.h:6-20, material.h:11-33, light.h:9-18, scene.h:19-36,
//              bvh.h:5-32, pathtracing.h:11-17
#pragma once
#include "vecmath.h"
#include "../../../include/trt.h"

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace trt
{
const int DIFFUSE = 0, SPECULAR = 1, TRANSMISSION = 2, INVALID = 3; // ray.h:5-8
const float INF = 114514.0f;                                        // bvh.h:5
const float PI = 3.1415926f;                                        // pathtracing.h:11
const float P_RR = 0.8f;                                            // pathtracing.h:12

class Ray // ray.h:10-21
{
public:
    Ray() {}
    Ray(vec3 s, vec3 d) : startpoint(s), direction(d) {}
    Ray(vec3 s, vec3 d, int r) : startpoint(s), direction(d), ray_type(r) {}
    vec3 startpoint, direction;
    int ray_type = INVALID;
};

class Triangle // triangle.h:9-26
{
public:
    double calAera();            // law-of-cosines area in double (triangle.cpp:3-10)
    vec3 findBaryCor(vec3 hitp); // least-squares barycentrics in double (triangle.cpp:12-29)
    vec3 v[3], vn[3];
    vec2 vt[3];
    vec3 normal, center;
    double area = 0.0; // running sum of its light's area (scene.cpp:201-203)
    std::string mtl_name;
    bool is_emissive = false;
    int face = -1; // ordinal of the `f` statement in the OBJ (side table: post-build index -> OBJ face)
};

class Camera // camera.h:6-20
{
public:
    void setCamera();
    Ray getRay(float u, float v);
    void Print();
    double fovy = 90;
    vec3 eye = vec3(278.0f, 273.0f, -800.0f), lookat = vec3(278.0f, 273.0f, -799.0f), up = vec3(0.f, 1.f, 0.f);
    double aspect_ratio = 1.0;
    vec3 lower_left_corner, horizontal, vertical;
};

// cv::Mat stand-in: decoded 8-bit BGR rows x cols, shared on copy like cv::Mat's header copy
struct Image
{
    int rows = 0, cols = 0;
    std::shared_ptr<std::vector<unsigned char>> data;
    bool empty() const { return !data || data->empty(); }
};

// Baseline JPEG -> 8-bit BGR (rows x cols x 3), bit-identical to OpenCV's cv::imread on such files (jpeg_decoder.cpp).
bool decodeJpeg(const unsigned char *bytes, size_t n, Image &out, std::string &error);
bool decodeJpegFile(const std::string &path, Image &out, std::string &error);

class Material // material.h:11-33
{
public:
    void readinMap(); // decodes map_Kd (baseline JPEG) itself; other formats through a "<map_Kd>.bgr" side-car
    vec3 Kd, Ks, Tr;
    float Ns = 1, Ni = 1;
    std::string map_Kd;
    bool is_emissive = false;
    vec3 radiance;
    double area = 0.0;
    std::vector<Triangle> triangles; // the light's triangles, OBJ order, cumulative area
    Image img;
    int map_height = 0, map_width = 0;
};

class Light // light.h:9-18
{
public:
    Light() {}
    Light(std::string m, vec3 r) : mtl_name(m), radiance(r) {}
    std::string mtl_name;
    vec3 radiance;
};

struct LoadError : std::runtime_error
{
    using std::runtime_error::runtime_error;
};

class Scene // scene.h:19-36
{
public:
    // The reference prints and exit()s on failure (scene.cpp:7-11,61-65,119-123); these throw LoadError.
    void readxml(std::string xml_path);
    void readmtl(std::string mtl_path, std::string base_dir);
    void readobj(std::string obj_path);

    int img_width = 0, img_height = 0;
    std::vector<Triangle> triangles;
    std::vector<Light> lights;
    std::unordered_map<std::string, Material> materials;
    Camera camera;
};

struct BVHNode // bvh.h:16-22
{
    BVHNode *left = nullptr, *right = nullptr;
    int index = 0, num = 0;
    vec3 AA, BB;
};

struct HitRecord // bvh.h:7-15
{
    bool is_hit = false;
    float distance = INF;
    vec3 hitpoint, direction, pn;
    Triangle triangle;
    int triangle_index = -1; // extra: position in scene.triangles (post-build)
    vec3 startpoint;         // extra: origin of the ray that produced the record (shade() re-derives hitpoint from it)
};

// bvh.cpp:16-144 — same topology and the same in-place reorder of `triangles` as the reference
// (sort-sweep "SAH" with the INF cap, median-x fallback, leaf <= leaf_num), computed on lean sort records.
// Nodes are owned by an arena that lives as long as the returned root's BVH (see freeBVH).
BVHNode *buildBVH(std::vector<Triangle> &triangles, int l, int r, int leaf_num);
void freeBVH(BVHNode *root);
int nodeCountBVH(const BVHNode *root);

// ---- GPU-backed hot path -------------------------------------------------------------------------
// Owns the device copy of a built scene.  Created once after buildBVH (main.cpp:76).
class DeviceScene
{
public:
    // layout_cache: a file for the acceleration layouts (trt_scene_create_cached: read when it holds this scene's, else
    // built and written there); nullptr = build every time
    DeviceScene(Scene &scene, BVHNode *root, int device = 0, const char *layout_cache = nullptr);
    // replica of `src` on another GPU (trt_scene_replicate: device-to-device copy, the layouts are not rebuilt)
    DeviceScene(const DeviceScene &src, int device, bool replicate);
    ~DeviceScene();
    DeviceScene(const DeviceScene &) = delete;
    DeviceScene &operator=(const DeviceScene &) = delete;
    trt_scene *handle() const { return h_; }
    Scene &scene() const { return *scene_; }

private:
    Scene *scene_;
    trt_scene *h_ = nullptr;
};

// POD view of (scene, root) for trt_scene_create; the arrays live inside the returned object.
struct SceneArrays
{
    std::vector<float> v, vn, vt, normal, node_box, light_v, light_vn;
    std::vector<int32_t> mtl, node_link, face;
    std::vector<double> light_cum_area;
    std::vector<trt_material> materials;
    std::vector<std::string> material_names;
    std::vector<trt_light> lights;
    std::vector<trt_texture> textures;
    std::vector<Image> texture_images;
    trt_scene_desc desc;
};
std::unique_ptr<SceneArrays> makeSceneArrays(Scene &scene, const BVHNode *root);

// traverseBVH (bvh.cpp:146-175) for a batch of rays, on the GPU. The single-ray form is the drop-in
// signature of bvh.h:32 and costs a full launch per call: use the batch form in loops.
std::vector<HitRecord> traverseBVH(const std::vector<Ray> &rays, DeviceScene &dev);
HitRecord traverseBVH(Ray ray, DeviceScene &dev);

// ---- PathTracing (pathtracing.h:14-17) ------------------------------------------------------------------------------
// shade() (pathTracing.cpp:3-102) on the GPU.  The batch form runs ONE wavefront from the given records (trt_shade):
// next-event estimation over the lights, Russian roulette and the whole bounce chain, radiance per record, wi = the
// negated ray direction as at main.cpp:101.  Random numbers: Philox stream (seed; record index, sample).  The
// single-record form is the drop-in signature of pathtracing.h:17 and costs a wavefront per call: use the batch form in
// loops.  Records must come from traverseBVH above (they carry triangle_index and startpoint).
// RR / Sample / nextRay (pathtracing.h:14-16) have no callers in the reference other than shade() and nextRay()
// themselves; they live inside the device shade kernel (csrc/wavefront.cu: k_shade, sampleLobe) and are not exported.
std::vector<vec3> shade(const std::vector<HitRecord> &records, DeviceScene &dev, uint64_t seed = 0, int sample = 0,
                        int max_depth = 0);
vec3 shade(HitRecord &res, vec3 dir, DeviceScene &dev, uint64_t seed = 0, int sample = 0, int max_depth = 0);

// The sample loop of main.cpp:79-113 (getRay + traverseBVH + shade + accumulate) on the GPU:
// fills image[H*W*3] (double, divided by spp) exactly as the reference's `image` buffer.
void renderImage(DeviceScene &dev, int spp, double *image, uint64_t seed = 0, int max_depth = 0);
// The same loop fanned out over several GPUs of one box (the role of main.cpp:79-81's OpenMP loop over samples):
// devs are replicas of one scene on different devices; trt_render_multi shards the samples, sums the per-GPU buffers
// with one reduce (NCCL, or the library's peer-memory kernel with TRT_RENDER_PEER_REDUCE in flags) and resolves.
void renderImage(const std::vector<DeviceScene *> &devs, int spp, double *image, uint64_t seed = 0, int max_depth = 0,
                 uint32_t flags = 0);
// Checkpointed form: renders `every` samples at a time, writing the accumulation buffer + progress to `checkpoint`
// after each step (trt_accum_save); if the file exists and matches (frame size, spp, seed, max_depth) the render
// resumes from it — bit-identical to an uninterrupted render.  Returns the number of samples rendered by THIS call.
int renderImageCheckpointed(DeviceScene &dev, int spp, double *image, const std::string &checkpoint, int every,
                            uint64_t seed = 0, int max_depth = 0);
} // namespace trt
