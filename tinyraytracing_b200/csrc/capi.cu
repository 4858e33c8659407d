// The C ABI of include/trt.h: scene upload, blocking / async closest-hit entry points, render entry points.
#include "scene_impl.h"

#include <algorithm>
#include <memory>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <stdexcept>

namespace trt
{
static thread_local std::string g_lastError;
void setLastError(const std::string &s) { g_lastError = s; }

namespace
{
// dst: a pointer member of s->view (recorded, so that trt_scene_replicate can re-point the copy), or anything else
template <typename T>
int upload(trt_scene *s, const T *src, size_t count, const T **dst, int texture_index = -1)
{
    *dst = nullptr;
    if (count == 0)
        return TRT_OK;
    void *p = nullptr;
    TRT_CUDA(cudaMalloc(&p, count * sizeof(T)));
    const char *view0 = reinterpret_cast<const char *>(&s->view), *d = reinterpret_cast<const char *>(dst);
    const size_t off = (d >= view0 && d < view0 + sizeof(SceneView)) ? (size_t)(d - view0) : trt_scene::kNotInView;
    s->allocations.push_back({p, count * sizeof(T), off, texture_index});
    TRT_CUDA(cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice));
    *dst = static_cast<const T *>(p);
    return TRT_OK;
}

int initStreams(trt_scene *s)
{
    cudaDeviceProp prop;
    TRT_CUDA(cudaGetDeviceProperties(&prop, s->device));
    s->sm_count = prop.multiProcessorCount;
    TRT_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    TRT_CUDA(cudaStreamCreateWithFlags(&s->copy_in, cudaStreamNonBlocking));
    TRT_CUDA(cudaStreamCreateWithFlags(&s->copy_out, cudaStreamNonBlocking));
    TRT_CUDA(cudaEventCreate(&s->ev[0]));
    TRT_CUDA(cudaEventCreate(&s->ev[1]));
    return TRT_OK;
}

int newStrictCounter(trt_scene *s)
{
    void *p = nullptr;
    TRT_CUDA(cudaMalloc(&p, sizeof(unsigned long long)));
    s->allocations.push_back({p, sizeof(unsigned long long), offsetof(SceneView, strict_counter), -1});
    TRT_CUDA(cudaMemset(p, 0, sizeof(unsigned long long)));
    s->view.strict_counter = static_cast<unsigned long long *>(p);
    return TRT_OK;
}

int fail(int code, const std::string &msg)
{
    setLastError(msg);
    return code;
}

// No C++ exception may unwind through the C ABI (the layout builders allocate large vectors and run std::async tasks):
// every extern "C" body that can throw runs inside guarded().
template <typename F>
int guarded(const char *what, F &&body)
{
    try
    {
        return body();
    }
    catch (const std::bad_alloc &)
    {
        return fail(TRT_ERR_LIMIT, std::string(what) + ": out of host memory");
    }
    catch (const std::length_error &e)
    {
        return fail(TRT_ERR_LIMIT, std::string(what) + ": " + e.what());
    }
    catch (const std::exception &e)
    {
        return fail(TRT_ERR_INVALID, std::string(what) + ": " + e.what());
    }
    catch (...)
    {
        return fail(TRT_ERR_INVALID, std::string(what) + ": unknown exception");
    }
}

// counts and the arrays they size: negative counts, or a NULL array where the count is positive, are rejected up front
std::string validateDesc(const trt_scene_desc &d)
{
    if (d.n_tris < 0 || d.n_nodes < 0 || d.n_materials < 1 || d.n_lights < 0 || d.n_light_tris < 0 || d.n_textures < 0)
        return "negative count (or no material)";
    if (d.width < 2 || d.height < 2)
        return "frame smaller than 2 x 2 (main.cpp:88-89 divides by W-1 and H-1)";
    if ((long long)d.width * d.height > 0x3fffffffll)
        return "frame larger than 2^30 pixels";
    if (d.n_tris > 0 && (!d.v || !d.normal || !d.mtl))
        return "triangle arrays missing";
    if (d.n_nodes > 0 && (!d.node_box || !d.node_link))
        return "node arrays missing";
    if (!d.materials)
        return "materials missing";
    if (d.n_lights > 0 && !d.lights)
        return "lights missing";
    if (d.n_light_tris > 0 && (!d.light_v || !d.light_vn || !d.light_cum_area))
        return "light triangle arrays missing";
    if (d.n_textures > 0 && !d.textures)
        return "textures missing";
    for (int i = 0; i < d.n_tris; ++i)
        if (d.mtl[i] < 0 || d.mtl[i] >= d.n_materials)
            return "triangle material index out of range";
    return "";
}

bool isSm100(int device)
{
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess)
        return false;
    return p.major == 10; // the library ships sm_100a SASS only
}

// Rays that start farther than this many scene scales from the coordinate origin take the exhaustive reference walk.
// The fast layout's boxes carry a pad of 256 ulp(scale) >= 128 x 2^-23 scale; the float rounding the pad has to absorb
// (hit point S + d t, slab distances) is about 3 x 2^-23 (|S| + |hit - S|) <= 3 x 2^-23 x (8 + 9) scale = 51 x 2^-23
// scale: a margin of 2.5 at the limit (round 1 stopped at 4 x scale; a camera a few object radii away from a small
// centred object then sent every primary ray down the slow path).
constexpr float kStrictOriginFactor = 8.0f;

// staging for the blocking host-pointer entry point
constexpr size_t kChunkRays = 1u << 19; // small chunks keep the un-overlapped head (first H2D) and tail (last kernel + D2H) short

int ensureStaging(trt_scene *s)
{
    if (s->chunk_rays)
        return TRT_OK;
    for (int b = 0; b < 2; ++b)
    {
        TRT_CUDA(cudaMallocHost(&s->stage_in[b], kChunkRays * 6 * sizeof(float)));
        TRT_CUDA(cudaMallocHost(&s->stage_out[b], kChunkRays * 8));
        TRT_CUDA(cudaMalloc((void **)&s->d_rays[b], kChunkRays * 6 * sizeof(float)));
        TRT_CUDA(cudaMalloc((void **)&s->d_id[b], kChunkRays * sizeof(int32_t)));
        TRT_CUDA(cudaMalloc((void **)&s->d_t[b], kChunkRays * sizeof(float)));
        TRT_CUDA(cudaEventCreateWithFlags(&s->ev_in[b], cudaEventDisableTiming));
        TRT_CUDA(cudaEventCreateWithFlags(&s->ev_k[b], cudaEventDisableTiming));
        TRT_CUDA(cudaEventCreateWithFlags(&s->ev_out[b], cudaEventDisableTiming));
    }
    s->chunk_rays = kChunkRays;
    return TRT_OK;
}

bool isPinnedOrManaged(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
} // namespace
} // namespace trt

trt_scene::~trt_scene()
{
    if (device >= 0)
        cudaSetDevice(device);
    cudaDeviceSynchronize();
    trt::destroyWavefront(this);
    for (const Alloc &a : allocations)
        cudaFree(a.p);
    cudaFree(d_counter);
    cudaFree(d_frame_image), cudaFree(d_frame_accum), cudaFree(d_frame_rgb8);
    for (int b = 0; b < 2; ++b)
    {
        if (stage_in[b])
            cudaFreeHost(stage_in[b]);
        if (stage_out[b])
            cudaFreeHost(stage_out[b]);
        cudaFree(d_rays[b]), cudaFree(d_id[b]), cudaFree(d_t[b]);
        for (cudaEvent_t e : {ev_in[b], ev_k[b], ev_out[b], ev[b]})
            if (e)
                cudaEventDestroy(e);
    }
    for (cudaStream_t st : {stream, copy_in, copy_out})
        if (st)
            cudaStreamDestroy(st);
}

using namespace trt;

extern "C"
{
int trt_version(void) { return TRT_VERSION; }
const char *trt_last_error(void) { return g_lastError.c_str(); }

int trt_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n; ++i)
        ok += isSm100(i) ? 1 : 0;
    return ok;
}

void *trt_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess)
    {
        setLastError("cudaMallocHost failed");
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void trt_host_free(void *p)
{
    if (p)
        cudaFreeHost(p);
}

static int sceneCreate(const trt_scene_desc *desc, int device, const char *layout_cache, int32_t *from_cache, trt_scene **out);

int trt_scene_create(const trt_scene_desc *desc, int device, trt_scene **out)
{
    if (!desc || !out)
        return fail(TRT_ERR_INVALID, "trt_scene_create: null argument");
    *out = nullptr;
    return guarded("trt_scene_create", [&] { return sceneCreate(desc, device, nullptr, nullptr, out); });
}

int trt_scene_create_cached(const trt_scene_desc *desc, int device, const char *layout_cache_path, int32_t *from_cache,
                            trt_scene **out)
{
    if (!desc || !out || !layout_cache_path)
        return fail(TRT_ERR_INVALID, "trt_scene_create_cached: null argument");
    *out = nullptr;
    if (from_cache)
        *from_cache = 0;
    return guarded("trt_scene_create_cached", [&] { return sceneCreate(desc, device, layout_cache_path, from_cache, out); });
}

// The layouts of a (validated) description: from the cache file when it holds them, built (and written there) otherwise.
// use_wide = 0: the scene keeps the reference-topology kernels (single leaf / too deep / empty).  "" or an error text.
static std::string makeLayouts(const trt_scene_desc &desc, const char *layout_cache, AccelBuild &ab, bool &use_wide,
                               bool &from_cache)
{
    from_cache = false;
    uint64_t key = 0;
    if (layout_cache)
    {
        key = layoutKey(desc);
        std::string why;
        if (loadLayout(layout_cache, key, ab, use_wide, why))
        {
            from_cache = true;
            return "";
        }
    }
    const std::string err = buildAccel(desc, ab);
    if (!err.empty())
        return err;
    use_wide = buildWide(desc, ab).empty(); // an error text here only means "this scene keeps the reference-topology kernels"
    if (layout_cache)
        saveLayout(ab, use_wide, key, layout_cache); // best effort: a cache that cannot be written costs the next build, nothing else
    return "";
}

static int sceneCreate(const trt_scene_desc *desc, int device, const char *layout_cache, int32_t *from_cache_out, trt_scene **out)
{
    const std::string bad = validateDesc(*desc);
    if (!bad.empty())
        return fail(TRT_ERR_INVALID, "trt_scene_create: " + bad);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        cudaGetLastError();
        return fail(TRT_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
    }
    if (device < 0 || device >= ndev || !isSm100(device))
        return fail(TRT_ERR_NO_DEVICE, "device is not an sm_100 (B200) GPU: this library ships sm_100a code only");

    AccelBuild ab;
    bool use_wide = false, from_cache = false;
    std::string err = makeLayouts(*desc, layout_cache, ab, use_wide, from_cache);
    if (!err.empty())
        return fail(TRT_ERR_INVALID, "trt_scene_create: " + err);
    if (from_cache_out)
        *from_cache_out = from_cache ? 1 : 0;
    if (ab.ref_depth >= TRT_REF_STACK_LIMIT)
        return fail(TRT_ERR_LIMIT, "reference tree deeper than the traversal stack (" +
                                       std::to_string(ab.ref_depth) + " >= " + std::to_string(TRT_REF_STACK_LIMIT) + ")");

    std::unique_ptr<trt_scene> s(new trt_scene());
    s->device = device;
    TRT_CUDA(cudaSetDevice(device));
    int rc;
    if ((rc = initStreams(s.get())))
        return rc;

    SceneView &v = s->view;
    if ((rc = upload(s.get(), ab.ref_nodes.data(), ab.ref_nodes.size(), &v.ref_nodes)))
        return rc;
    if ((rc = upload(s.get(), ab.tri_geom.data(), ab.tri_geom.size(), &v.tri_geom)))
        return rc;
    if ((rc = upload(s.get(), ab.tri_key.data(), ab.tri_key.size(), &v.tri_key)))
        return rc;
    if ((rc = upload(s.get(), desc->v, (size_t)desc->n_tris * 9, &v.tri_v)))
        return rc;
    if (ab.rank_tri.empty())
        ab.rank_tri.assign(1, -1); // empty scene: rank 0 = miss
    if ((rc = upload(s.get(), ab.tri_rank.data(), ab.tri_rank.size(), &v.tri_rank)))
        return rc;
    if ((rc = upload(s.get(), ab.rank_tri.data(), ab.rank_tri.size(), &v.rank_tri)))
        return rc;
    v.miss_rank = ab.miss_rank;
    v.root_link = ab.root_link;
    v.n_tris = desc->n_tris;
    v.use_wide = use_wide ? 1 : 0;
    v.wide_root = ab.wide_root;
    if ((rc = upload(s.get(), ab.wide_nodes.data(), ab.wide_nodes.size(), &v.wide_nodes)) ||
        (rc = upload(s.get(), ab.fast_geom.data(), ab.fast_geom.size(), &v.fast_geom)) ||
        (rc = upload(s.get(), ab.fast_key.data(), ab.fast_key.size(), &v.fast_key)) ||
        (rc = upload(s.get(), ab.fast_rank.data(), ab.fast_rank.size(), &v.fast_rank)) ||
        (rc = upload(s.get(), ab.fast_orig.data(), ab.fast_orig.size(), &v.fast_orig)) ||
        (rc = upload(s.get(), ab.fast_leaf.data(), ab.fast_leaf.size(), &v.fast_leaf)) ||
        (rc = upload(s.get(), ab.ref_leaf_box.data(), ab.ref_leaf_box.size(), &v.ref_leaf_box)) ||
        (rc = upload(s.get(), ab.light_box.data(), ab.light_box.size(), &v.light_box)) ||
        (rc = upload(s.get(), ab.ref_leaf_parent.data(), ab.ref_leaf_parent.size(), &v.ref_leaf_parent)))
        return rc;
    v.check_leaf_box = ab.root_is_reference_leaf ? 0 : 1;
    {
        const char *e = getenv("TRT_SHADOW_STOP");
        s->shadow_stop = v.use_wide && (e ? atoi(e) != 0 : ab.fast_orig.size() >= 8192);
        e = getenv("TRT_CLOSEST_PLAIN");
        s->closest_plain = v.use_wide && (e ? atoi(e) != 0 : ab.wide_nodes.size() <= 64);
    }
    v.strict_origin_limit = kStrictOriginFactor * ab.scene_scale;
    if ((rc = newStrictCounter(s.get())))
        return rc;
    // b * inv and S * inv of the fused culling test stay below 4e30 for every box plane b and admitted origin S
    v.inv_cull_limit = 1.0e30f / std::max(ab.scene_scale, 1.0f);

    std::vector<TriShade> shade(desc->n_tris);
    for (int i = 0; i < desc->n_tris; ++i)
    {
        if (desc->vn)
            std::memcpy(shade[i].vn, desc->vn + (size_t)i * 9, 36);
        else
            std::memset(shade[i].vn, 0, 36);
        if (desc->vt)
            std::memcpy(shade[i].vt, desc->vt + (size_t)i * 6, 24);
        else
            std::memset(shade[i].vt, 0, 24);
        shade[i].mtl = desc->mtl[i];
    }
    if ((rc = upload(s.get(), shade.data(), shade.size(), &v.tri_shade)))
        return rc;
    {
        std::vector<int32_t> fast_mtl(ab.fast_orig.size());
        for (size_t i = 0; i < fast_mtl.size(); ++i)
            fast_mtl[i] = desc->mtl[ab.fast_orig[i]];
        if ((rc = upload(s.get(), fast_mtl.data(), fast_mtl.size(), &v.fast_mtl)))
            return rc;
    }

    std::vector<DeviceTexture> tex(desc->n_textures);
    for (int i = 0; i < desc->n_textures; ++i)
    {
        const trt_texture &t = desc->textures[i];
        if (t.rows < 1 || t.cols < 1 || !t.bgr)
            return fail(TRT_ERR_INVALID, "trt_scene_create: empty texture");
        tex[i].rows = t.rows, tex[i].cols = t.cols;
        if ((rc = upload(s.get(), t.bgr, (size_t)t.rows * t.cols * 3, &tex[i].bgr, i)))
            return rc;
    }
    if ((rc = upload(s.get(), tex.data(), tex.size(), &v.textures)))
        return rc;
    s->host_textures = tex;

    std::vector<DeviceMaterial> mats(desc->n_materials);
    for (int i = 0; i < desc->n_materials; ++i)
    {
        const trt_material &m = desc->materials[i];
        if (m.texture >= desc->n_textures || m.texture < -1)
            return fail(TRT_ERR_INVALID, "trt_scene_create: material texture index out of range");
        mats[i].Kd = make_float3(m.Kd[0], m.Kd[1], m.Kd[2]);
        mats[i].Ks = make_float3(m.Ks[0], m.Ks[1], m.Ks[2]);
        mats[i].Tr = make_float3(m.Tr[0], m.Tr[1], m.Tr[2]);
        mats[i].radiance = make_float3(m.radiance[0], m.radiance[1], m.radiance[2]);
        mats[i].Ns = m.Ns, mats[i].Ni = m.Ni;
        mats[i].is_emissive = m.is_emissive, mats[i].texture = m.texture;
        mats[i].area = m.area;
        {
            // glm::length in float (products first, then left-to-right adds; this file is compiled un-fused), as the
            // device computed it per vertex until round 2
            const float kd2 = (m.Kd[0] * m.Kd[0] + m.Kd[1] * m.Kd[1]) + m.Kd[2] * m.Kd[2];
            const float ks2 = (m.Ks[0] * m.Ks[0] + m.Ks[1] * m.Ks[1]) + m.Ks[2] * m.Ks[2];
            const double Kd_len = (double)std::sqrt(kd2), Ks_len = (double)std::sqrt(ks2);
            mats[i].kd = Kd_len / (Kd_len + Ks_len), mats[i].ks = Ks_len / (Kd_len + Ks_len);
        }
    }
    if ((rc = upload(s.get(), mats.data(), mats.size(), &v.materials)))
        return rc;
    v.n_materials = desc->n_materials;

    std::vector<DeviceLight> lights(desc->n_lights);
    for (int i = 0; i < desc->n_lights; ++i)
    {
        const trt_light &l = desc->lights[i];
        if (l.material < 0 || l.material >= desc->n_materials || l.first_tri < 0 || l.n_tris < 0 ||
            l.first_tri + l.n_tris > desc->n_light_tris)
            return fail(TRT_ERR_INVALID, "trt_scene_create: light out of range");
        if (l.n_tris >= (1 << 23))
            return fail(TRT_ERR_LIMIT, "trt_scene_create: a light of 2^23 triangles or more");
        lights[i] = DeviceLight{l.material, l.first_tri, l.n_tris, (float)(double(1) / desc->materials[l.material].area)};
    }
    if ((rc = upload(s.get(), lights.data(), lights.size(), &v.lights)))
        return rc;
    if ((rc = upload(s.get(), desc->light_v, (size_t)desc->n_light_tris * 9, &v.light_v)))
        return rc;
    if ((rc = upload(s.get(), desc->light_vn, (size_t)desc->n_light_tris * 9, &v.light_vn)))
        return rc;
    if ((rc = upload(s.get(), desc->light_cum_area, (size_t)desc->n_light_tris, &v.light_cum_area)))
        return rc;
    v.n_lights = desc->n_lights;
    v.first_light_area = desc->n_lights > 0 ? desc->materials[desc->lights[0].material].area : 0.0;

    v.cam.eye = make_float3(desc->eye[0], desc->eye[1], desc->eye[2]);
    v.cam.llc = make_float3(desc->lower_left_corner[0], desc->lower_left_corner[1], desc->lower_left_corner[2]);
    v.cam.horizontal = make_float3(desc->horizontal[0], desc->horizontal[1], desc->horizontal[2]);
    v.cam.vertical = make_float3(desc->vertical[0], desc->vertical[1], desc->vertical[2]);
    v.cam.width = desc->width, v.cam.height = desc->height;
    s->width = desc->width, s->height = desc->height;

    s->stats.accel_nodes = (int32_t)(v.use_wide ? ab.wide_nodes.size() : ab.ref_nodes.size());
    s->stats.accel_leaves = ab.n_leaves;
    s->stats.ref_depth = ab.ref_depth;
    s->stats.device = device;
    s->stats.accel_slivers = ab.n_sliver;
    s->stats.accel_needles = ab.n_needle;
    *out = s.release();
    return TRT_OK;
}

void trt_scene_destroy(trt_scene *s)
{
    delete s; // ~trt_scene releases every device / pinned allocation, stream and event
}

static int sceneReplicate(const trt_scene *src, int device, trt_scene **out)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev || !isSm100(device))
    {
        cudaGetLastError();
        return fail(TRT_ERR_NO_DEVICE, "trt_scene_replicate: device is not an sm_100 (B200) GPU");
    }
    std::unique_ptr<trt_scene> s(new trt_scene());
    s->device = device;
    TRT_CUDA(cudaSetDevice(device));
    int rc;
    if ((rc = initStreams(s.get())))
        return rc;
    s->view = src->view;
    s->width = src->width, s->height = src->height;
    s->host_textures = src->host_textures;
    s->shadow_stop = src->shadow_stop;
    s->closest_plain = src->closest_plain;
    s->stats = trt_stats{};
    s->stats.accel_nodes = src->stats.accel_nodes, s->stats.accel_leaves = src->stats.accel_leaves;
    s->stats.ref_depth = src->stats.ref_depth, s->stats.accel_slivers = src->stats.accel_slivers;
    s->stats.accel_needles = src->stats.accel_needles, s->stats.device = device;
    for (const trt_scene::Alloc &a : src->allocations)
    {
        if (a.view_offset == offsetof(SceneView, strict_counter) || a.view_offset == offsetof(SceneView, textures))
            continue; // rebuilt below
        void *p = nullptr;
        TRT_CUDA(cudaMalloc(&p, a.bytes));
        s->allocations.push_back({p, a.bytes, a.view_offset, a.texture_index});
        // device to device (over NVLink when the two are peers, staged by the driver otherwise)
        TRT_CUDA(cudaMemcpyPeer(p, device, a.p, src->device, a.bytes));
        if (a.view_offset != trt_scene::kNotInView)
            std::memcpy(reinterpret_cast<char *>(&s->view) + a.view_offset, &p, sizeof p);
        else if (a.texture_index >= 0 && a.texture_index < (int)s->host_textures.size())
            s->host_textures[a.texture_index].bgr = static_cast<const uint8_t *>(p);
        else
            return fail(TRT_ERR_INVALID, "trt_scene_replicate: allocation without an owner");
    }
    s->view.textures = nullptr;
    if ((rc = upload(s.get(), s->host_textures.data(), s->host_textures.size(), &s->view.textures)))
        return rc;
    if ((rc = newStrictCounter(s.get())))
        return rc;
    TRT_CUDA(cudaDeviceSynchronize());
    *out = s.release();
    return TRT_OK;
}

int trt_scene_replicate(const trt_scene *src, int device, trt_scene **out)
{
    if (!src || !out)
        return fail(TRT_ERR_INVALID, "trt_scene_replicate: null argument");
    *out = nullptr;
    return guarded("trt_scene_replicate", [&] { return sceneReplicate(src, device, out); });
}

int trt_trace_closest_async(trt_scene *s, const float *d_rays6, size_t n, int32_t *d_id, float *d_t, uint32_t flags,
                            void *stream)
{
    if (!s || (!d_rays6 && n))
        return fail(TRT_ERR_INVALID, "trt_trace_closest_async: null argument");
    TRT_CUDA(cudaSetDevice(s->device));
    return launchClosest(s, d_rays6, n, d_id, d_t, flags, static_cast<cudaStream_t>(stream));
}

static int traceHostPipelined(trt_scene *s, const float *rays6, size_t n, int32_t *tri_id, float *t, uint32_t flags)
{
    int rc = ensureStaging(s);
    if (rc)
        return rc;
    const bool in_pinned = isPinnedOrManaged(rays6);
    const bool id_pinned = tri_id && isPinnedOrManaged(tri_id), t_pinned = t && isPinnedOrManaged(t);
    const size_t C = s->chunk_rays;
    const size_t nchunks = (n + C - 1) / C;
    // pending D2H of chunk c-2 must be drained (and unstaged) before buffer b is reused
    auto drain = [&](size_t c) -> int {
        const int b = (int)(c & 1);
        const size_t off = c * C, m = std::min(C, n - off);
        TRT_CUDA(cudaEventSynchronize(s->ev_out[b]));
        if (tri_id && !id_pinned)
            std::memcpy(tri_id + off, s->stage_out[b], m * 4);
        if (t && !t_pinned)
            std::memcpy(t + off, (char *)s->stage_out[b] + C * 4, m * 4);
        return TRT_OK;
    };
    TRT_CUDA(cudaEventRecord(s->ev[0], s->stream));
    for (size_t c = 0; c < nchunks; ++c)
    {
        const int b = (int)(c & 1);
        const size_t off = c * C, m = std::min(C, n - off);
        if (c >= 2 && (rc = drain(c - 2)))
            return rc;
        const float *src = rays6 + off * 6;
        if (!in_pinned)
        {
            std::memcpy(s->stage_in[b], src, m * 24);
            src = static_cast<const float *>(s->stage_in[b]);
        }
        // the kernel that last read d_rays[b] (chunk c-2) must be done before it is overwritten
        TRT_CUDA(cudaStreamWaitEvent(s->copy_in, s->ev_k[b], 0));
        TRT_CUDA(cudaMemcpyAsync(s->d_rays[b], src, m * 24, cudaMemcpyHostToDevice, s->copy_in));
        TRT_CUDA(cudaEventRecord(s->ev_in[b], s->copy_in));
        TRT_CUDA(cudaStreamWaitEvent(s->stream, s->ev_in[b], 0));
        TRT_CUDA(cudaStreamWaitEvent(s->stream, s->ev_out[b], 0)); // outputs of chunk c-2 copied out
        if ((rc = launchClosest(s, s->d_rays[b], m, s->d_id[b], s->d_t[b], flags, s->stream)))
            return rc;
        TRT_CUDA(cudaEventRecord(s->ev_k[b], s->stream));
        TRT_CUDA(cudaStreamWaitEvent(s->copy_out, s->ev_k[b], 0));
        if (tri_id)
            TRT_CUDA(cudaMemcpyAsync(id_pinned ? (void *)(tri_id + off) : s->stage_out[b], s->d_id[b], m * 4,
                                     cudaMemcpyDeviceToHost, s->copy_out));
        if (t)
            TRT_CUDA(cudaMemcpyAsync(t_pinned ? (void *)(t + off) : (void *)((char *)s->stage_out[b] + C * 4), s->d_t[b],
                                     m * 4, cudaMemcpyDeviceToHost, s->copy_out));
        TRT_CUDA(cudaEventRecord(s->ev_out[b], s->copy_out));
    }
    TRT_CUDA(cudaEventRecord(s->ev[1], s->stream));
    for (size_t c = (nchunks >= 2 ? nchunks - 2 : 0); c < nchunks; ++c)
        if ((rc = drain(c)))
            return rc;
    TRT_CUDA(cudaStreamSynchronize(s->stream));
    float ms = 0;
    TRT_CUDA(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
    s->stats.last_trace_ms = ms;
    return TRT_OK;
}

int trt_trace_closest(trt_scene *s, const float *rays6, size_t n, int32_t *tri_id, float *t, uint32_t flags)
{
    if (!s || (!rays6 && n))
        return fail(TRT_ERR_INVALID, "trt_trace_closest: null argument");
    TRT_CUDA(cudaSetDevice(s->device));
    if (flags & TRT_TRACE_DEVICE_PTRS)
    {
        TRT_CUDA(cudaEventRecord(s->ev[0], s->stream));
        int rc = launchClosest(s, rays6, n, tri_id, t, flags, s->stream);
        if (rc)
            return rc;
        TRT_CUDA(cudaEventRecord(s->ev[1], s->stream));
        TRT_CUDA(cudaStreamSynchronize(s->stream));
        float ms = 0;
        TRT_CUDA(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
        s->stats.last_trace_ms = ms;
        return TRT_OK;
    }
    // host pointers: chunked, double-buffered pipeline  H2D (copy_in) -> kernel (stream) -> D2H (copy_out)
    const int rc = traceHostPipelined(s, rays6, n, tri_id, t, flags);
    if (rc != TRT_OK)
    {
        // copies into the caller's buffers may still be in flight: nothing of this call may touch them after it returns
        cudaStreamSynchronize(s->copy_in), cudaStreamSynchronize(s->stream), cudaStreamSynchronize(s->copy_out);
        cudaGetLastError();
    }
    return rc;
}

int trt_hit_attributes(trt_scene *s, const float *rays6, const int32_t *tri_id, const float *t, size_t n,
                       float *hitpoint3, float *pn3)
{
    if (!s || ((!rays6 || !tri_id || !t) && n))
        return fail(TRT_ERR_INVALID, "trt_hit_attributes: null argument");
    if (n == 0)
        return TRT_OK;
    TRT_CUDA(cudaSetDevice(s->device));
    float *d_r = nullptr, *d_t = nullptr, *d_h = nullptr, *d_p = nullptr;
    int32_t *d_i = nullptr;
    int rc = TRT_OK;
    auto cleanup = [&] { cudaFree(d_r), cudaFree(d_t), cudaFree(d_h), cudaFree(d_p), cudaFree(d_i); };
#define TRT_TRY(call)                                                                                               \
    do                                                                                                              \
    {                                                                                                               \
        cudaError_t e__ = (call);                                                                                   \
        if (e__ != cudaSuccess)                                                                                     \
        {                                                                                                           \
            setLastError(std::string(#call) + ": " + cudaGetErrorString(e__));                                      \
            cleanup();                                                                                              \
            return TRT_ERR_CUDA;                                                                                    \
        }                                                                                                           \
    } while (0)
    TRT_TRY(cudaMalloc((void **)&d_r, n * 24));
    TRT_TRY(cudaMalloc((void **)&d_t, n * 4));
    TRT_TRY(cudaMalloc((void **)&d_i, n * 4));
    TRT_TRY(cudaMalloc((void **)&d_h, n * 12));
    TRT_TRY(cudaMalloc((void **)&d_p, n * 12));
    TRT_TRY(cudaMemcpyAsync(d_r, rays6, n * 24, cudaMemcpyHostToDevice, s->stream));
    TRT_TRY(cudaMemcpyAsync(d_t, t, n * 4, cudaMemcpyHostToDevice, s->stream));
    TRT_TRY(cudaMemcpyAsync(d_i, tri_id, n * 4, cudaMemcpyHostToDevice, s->stream));
    rc = launchHitAttributes(s, d_r, d_i, d_t, n, d_h, d_p, s->stream);
    if (rc == TRT_OK)
    {
        if (hitpoint3)
            TRT_TRY(cudaMemcpyAsync(hitpoint3, d_h, n * 12, cudaMemcpyDeviceToHost, s->stream));
        if (pn3)
            TRT_TRY(cudaMemcpyAsync(pn3, d_p, n * 12, cudaMemcpyDeviceToHost, s->stream));
        TRT_TRY(cudaStreamSynchronize(s->stream));
    }
    cleanup();
    return rc;
#undef TRT_TRY
}

int trt_trace_counters(trt_scene *s, const float *rays6, size_t n, uint64_t out4[4])
{
    if (!s || !out4 || (!rays6 && n))
        return fail(TRT_ERR_INVALID, "trt_trace_counters: null argument");
    out4[0] = out4[1] = out4[2] = out4[3] = 0;
    if (n == 0)
        return TRT_OK;
    TRT_CUDA(cudaSetDevice(s->device));
    float *d_r = nullptr;
    unsigned long long *d_o = nullptr;
    TRT_CUDA(cudaMalloc((void **)&d_r, n * 24));
    if (cudaMalloc((void **)&d_o, 32) != cudaSuccess)
    {
        cudaFree(d_r);
        return fail(TRT_ERR_CUDA, "cudaMalloc failed");
    }
    cudaError_t e = cudaMemcpyAsync(d_r, rays6, n * 24, cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess)
        e = cudaMemsetAsync(d_o, 0, 32, s->stream);
    int rc = (e == cudaSuccess) ? launchClosestCounters(s, d_r, n, d_o, s->stream) : TRT_OK;
    if (e == cudaSuccess && rc == TRT_OK)
        e = cudaMemcpyAsync(out4, d_o, 32, cudaMemcpyDeviceToHost, s->stream);
    const cudaError_t es = cudaStreamSynchronize(s->stream);
    e = (e == cudaSuccess) ? es : e;
    cudaFree(d_r), cudaFree(d_o);
    if (rc == TRT_OK && e != cudaSuccess)
        return fail(TRT_ERR_CUDA, std::string("trt_trace_counters: ") + cudaGetErrorString(e));
    return rc;
}

int trt_render_accumulate(trt_scene *s, const trt_render_params *p, double *d_accum, void *stream)
{
    if (!s || !p || !d_accum)
        return fail(TRT_ERR_INVALID, "trt_render_accumulate: null argument");
    if (p->spp < 1 || p->sample_begin < 0 || p->sample_end < p->sample_begin || p->sample_end > p->spp || p->max_depth < 0)
        return fail(TRT_ERR_INVALID, "trt_render_accumulate: bad sample range / spp / max_depth");
    TRT_CUDA(cudaSetDevice(s->device));
    return renderAccumulate(s, *p, d_accum, static_cast<cudaStream_t>(stream));
}

int trt_resolve(trt_scene *s, const double *d_accum, int32_t spp, double *image_rgb, uint8_t *rgb8, void *stream_)
{
    if (!s || !d_accum || spp < 1)
        return fail(TRT_ERR_INVALID, "trt_resolve: bad argument");
    TRT_CUDA(cudaSetDevice(s->device));
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const size_t n = (size_t)s->width * s->height * 3;
    if (!s->d_frame_image)
        TRT_CUDA(cudaMalloc((void **)&s->d_frame_image, n * sizeof(double)));
    if (!s->d_frame_rgb8)
        TRT_CUDA(cudaMalloc((void **)&s->d_frame_rgb8, n));
    int rc = resolveImage(s, d_accum, spp, s->d_frame_image, s->d_frame_rgb8, stream);
    if (rc != TRT_OK)
        return rc;
    // pinned destinations (trt_host_alloc) are written by the copy engine directly; pageable ones are staged by the driver
    if (image_rgb)
        TRT_CUDA(cudaMemcpyAsync(image_rgb, s->d_frame_image, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (rgb8)
        TRT_CUDA(cudaMemcpyAsync(rgb8, s->d_frame_rgb8, n, cudaMemcpyDeviceToHost, stream));
    TRT_CUDA(cudaStreamSynchronize(stream));
    return TRT_OK;
}

int trt_render(trt_scene *s, const trt_render_params *p, double *image_rgb)
{
    if (!s || !p || !image_rgb)
        return fail(TRT_ERR_INVALID, "trt_render: null argument");
    TRT_CUDA(cudaSetDevice(s->device));
    const size_t n = (size_t)s->width * s->height * 3;
    if (!s->d_frame_accum)
        TRT_CUDA(cudaMalloc((void **)&s->d_frame_accum, n * sizeof(double)));
    TRT_CUDA(cudaMemsetAsync(s->d_frame_accum, 0, n * sizeof(double), s->stream));
    int rc = trt_render_accumulate(s, p, s->d_frame_accum, s->stream);
    if (rc == TRT_OK)
        rc = trt_resolve(s, s->d_frame_accum, p->spp, image_rgb, nullptr, s->stream);
    return rc;
}

int trt_layout_check(const trt_scene_desc *desc, trt_layout_report *report)
{
    if (!desc || !report)
        return fail(TRT_ERR_INVALID, "trt_layout_check: null argument");
    return guarded("trt_layout_check", [&]() -> int {
    const std::string bad = validateDesc(*desc);
    if (!bad.empty())
        return fail(TRT_ERR_INVALID, "trt_layout_check: " + bad);
    AccelBuild ab;
    std::string err = buildAccel(*desc, ab);
    if (!err.empty())
        return fail(TRT_ERR_INVALID, "trt_layout_check: " + err);
    buildWide(*desc, ab); // an error text here only means "this scene keeps the reference-topology kernels"
    err = checkLayout(*desc, ab, *report);
    if (!err.empty())
        return fail(TRT_ERR_INVALID, "trt_layout_check: " + err);
    return TRT_OK;
    });
}

struct trt_layout
{
    trt::AccelBuild ab;
};

static int layoutBuild(const trt_scene_desc *desc, const char *layout_cache, int32_t *from_cache_out, trt_layout **out,
                       trt_layout_view *view);

int trt_layout_build(const trt_scene_desc *desc, trt_layout **out, trt_layout_view *view)
{
    return layoutBuild(desc, nullptr, nullptr, out, view);
}

int trt_layout_build_cached(const trt_scene_desc *desc, const char *layout_cache_path, int32_t *from_cache, trt_layout **out,
                            trt_layout_view *view)
{
    if (!layout_cache_path)
        return fail(TRT_ERR_INVALID, "trt_layout_build_cached: null argument");
    if (from_cache)
        *from_cache = 0;
    return layoutBuild(desc, layout_cache_path, from_cache, out, view);
}

static int layoutBuild(const trt_scene_desc *desc, const char *layout_cache, int32_t *from_cache_out, trt_layout **out,
                       trt_layout_view *view)
{
    if (!desc || !out || !view)
        return fail(TRT_ERR_INVALID, "trt_layout_build: null argument");
    return guarded("trt_layout_build", [&]() -> int {
    const std::string bad = validateDesc(*desc);
    if (!bad.empty())
        return fail(TRT_ERR_INVALID, "trt_layout_build: " + bad);
    std::unique_ptr<trt_layout> l(new trt_layout());
    AccelBuild &ab = l->ab;
    bool use_wide = false, from_cache = false;
    const std::string err = makeLayouts(*desc, layout_cache, ab, use_wide, from_cache);
    if (!err.empty())
        return fail(TRT_ERR_INVALID, "trt_layout_build: " + err);
    if (from_cache_out)
        *from_cache_out = from_cache ? 1 : 0;
    static_assert(sizeof(WideNode) == 128 && sizeof(TriGeom) == 48 && sizeof(RefNode) == 64, "layout records");
    std::memset(view, 0, sizeof *view);
    view->n_wide_nodes = (int32_t)ab.wide_nodes.size();
    view->wide_root = use_wide ? ab.wide_root : TRT_LINK_EMPTY;
    view->n_fast_tris = (int32_t)ab.fast_orig.size();
    view->n_ref_leaves = (int32_t)ab.ref_leaf_parent.size();
    view->n_ref_inner = (int32_t)ab.ref_nodes.size();
    view->check_leaf_box = ab.root_is_reference_leaf ? 0 : 1;
    view->strict_origin_limit = kStrictOriginFactor * ab.scene_scale;
    view->miss_key = TRT_MISS_KEY;
    view->wide_nodes = reinterpret_cast<const float *>(ab.wide_nodes.data());
    view->fast_geom = reinterpret_cast<const float *>(ab.fast_geom.data());
    view->fast_key = ab.fast_key.data();
    view->fast_orig = ab.fast_orig.data();
    view->fast_leaf = ab.fast_leaf.data();
    view->ref_leaf_box = reinterpret_cast<const float *>(ab.ref_leaf_box.data());
    view->ref_leaf_parent = ab.ref_leaf_parent.data();
    view->ref_nodes = reinterpret_cast<const float *>(ab.ref_nodes.data());
    *out = l.release();
    return TRT_OK;
    });
}

void trt_layout_free(trt_layout *l) { delete l; }

int trt_get_stats(trt_scene *s, trt_stats *out)
{
    if (!s || !out)
        return fail(TRT_ERR_INVALID, "trt_get_stats: null argument");
    *out = s->stats;
    if (s->view.strict_counter)
    {
        TRT_CUDA(cudaSetDevice(s->device));
        unsigned long long c = 0;
        TRT_CUDA(cudaMemcpy(&c, s->view.strict_counter, sizeof c, cudaMemcpyDeviceToHost)); // waits for the work in flight
        out->rays_strict = c;
    }
    return TRT_OK;
}

int trt_reset_stats(trt_scene *s)
{
    if (!s)
        return fail(TRT_ERR_INVALID, "trt_reset_stats: null argument");
    s->stats.rays_closest = s->stats.rays_shadow = s->stats.paths = s->stats.kernel_launches = 0;
    if (s->view.strict_counter)
    {
        TRT_CUDA(cudaSetDevice(s->device));
        TRT_CUDA(cudaMemset(s->view.strict_counter, 0, sizeof(unsigned long long)));
    }
    return TRT_OK;
}
}
