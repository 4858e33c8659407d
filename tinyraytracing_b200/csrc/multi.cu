// Entry points that sit on top of the wavefront renderer: multi-GPU render inside the library (trt_render_multi),
// accumulation checkpoints (trt_accum_*), and the batch form of PathTracing::shade (trt_shade).
//
// trt_render_multi replaces the reference's own fan-out — `omp parallel for` over the samples, main.cpp:79-81 — by
// one host thread per GPU, each rendering its share of the samples of every pixel into its GPU's float64
// accumulation buffer.  The buffers are then summed onto the first GPU in one of two ways:
//   * ncclReduce(sum, float64, W*H*3) on communicators from ncclCommInitAll (north_star: "one NCCL reduce over
//     NVLink").  libnccl.so.2 is loaded with dlopen on first use, so the library has no link-time NCCL dependency and
//     shares the copy a host program (torch) may already have loaded;
//   * TRT_RENDER_PEER_REDUCE: k_peer_reduce_resolve, one kernel on the first GPU that reads every peer's buffer
//     through NVLink peer mappings, adds them in rank order and resolves (divide by spp, gamma, 8-bit pack) in the same
//     pass — the frame crosses NVLink once and is never written back as a sum.
#include "scene_impl.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <memory>
#include <mutex>
#include <nccl.h> // types and enums only: every NCCL function is reached through dlsym
#include <thread>
#include <vector>

namespace trt
{
namespace
{
int fail(int code, const std::string &msg)
{
    setLastError(msg);
    return code;
}

// ---------------------------------------------------------------------------------------------- NCCL through dlopen
struct Nccl
{
    void *lib = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclReduce) Reduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string error;
    std::map<std::vector<int>, std::vector<ncclComm_t>> comms; // one clique per device list, kept for the process
};

Nccl &nccl()
{
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *name : {"libnccl.so.2", "libnccl.so"})
            if ((n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL)))
                break;
        if (!n.lib)
        {
            n.error = std::string("cannot load libnccl.so.2: ") + dlerror();
            return;
        }
        auto sym = [&](const char *name) {
            void *p = dlsym(n.lib, name);
            if (!p && n.error.empty())
                n.error = std::string("libnccl.so.2 lacks ") + name;
            return p;
        };
        n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(sym("ncclCommInitAll"));
        n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
        n.Reduce = reinterpret_cast<decltype(n.Reduce)>(sym("ncclReduce"));
        n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
        n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return n;
}
std::mutex g_ncclMutex;

// ------------------------------------------------------------------------- fused peer reduce + resolve (rank order)
constexpr int kMaxPeers = 16;
struct PeerBuffers
{
    const double *p[kMaxPeers];
    int n;
};

__global__ void __launch_bounds__(256) k_peer_reduce_resolve(PeerBuffers src, size_t n, int spp, double *accum0, double *image,
                                                             uint8_t *rgb8)
{
    // two doubles per thread: 128-bit loads from the peers' HBM over NVLink
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i >= n)
        return;
    double a, b = 0.0;
    if (i + 1 < n)
    {
        double2 v = *reinterpret_cast<const double2 *>(src.p[0] + i);
        for (int k = 1; k < src.n; ++k)
        {
            const double2 w = *reinterpret_cast<const double2 *>(src.p[k] + i);
            v.x += w.x, v.y += w.y; // rank order: the sum does not depend on which GPU finished first
        }
        a = v.x, b = v.y;
    }
    else
    {
        a = src.p[0][i];
        for (int k = 1; k < src.n; ++k)
            a += src.p[k][i];
    }
    const double s[2] = {a, b};
    for (int k = 0; k < 2 && i + k < n; ++k)
    {
        if (accum0)
            accum0[i + k] = s[k]; // the summed buffer stays available on the first GPU (checkpoints)
        const double v = s[k] / (double)spp; // main.cpp:101
        if (image)
            image[i + k] = v;
        if (rgb8)
        {
            // main.cpp:34: (unsigned char)clamp(pow(v, 1.0f / 2.2f) * 255, 0.0, 255.0)
            double g = pow(v, (double)(1.0f / 2.2f)) * 255;
            g = (g < 0.0) ? 0.0 : g;
            g = (255.0 < g) ? 255.0 : g;
            rgb8[i + k] = (uint8_t)g;
        }
    }
}

int ensureFrameBuffers(trt_scene *s)
{
    const size_t n = (size_t)s->width * s->height * 3;
    if (!s->d_frame_accum)
        TRT_CUDA(cudaMalloc((void **)&s->d_frame_accum, n * sizeof(double)));
    if (!s->d_frame_image)
        TRT_CUDA(cudaMalloc((void **)&s->d_frame_image, n * sizeof(double)));
    if (!s->d_frame_rgb8)
        TRT_CUDA(cudaMalloc((void **)&s->d_frame_rgb8, n));
    return TRT_OK;
}

int renderMulti(trt_scene *const *scenes, int n, const trt_render_params &p, double *image_rgb, uint8_t *rgb8)
{
    trt_scene *s0 = scenes[0];
    const size_t count = (size_t)s0->width * s0->height * 3;
    const bool peer = (p.flags & TRT_RENDER_PEER_REDUCE) != 0 || n == 1;
    std::vector<int> devs(n);
    for (int i = 0; i < n; ++i)
    {
        if (!scenes[i])
            return fail(TRT_ERR_INVALID, "trt_render_multi: null scene");
        if (scenes[i]->width != s0->width || scenes[i]->height != s0->height || scenes[i]->view.n_tris != s0->view.n_tris ||
            scenes[i]->view.n_lights != s0->view.n_lights)
            return fail(TRT_ERR_INVALID, "trt_render_multi: the scenes are not replicas of one description");
        for (int j = 0; j < i; ++j)
            if (scenes[j] == scenes[i])
                return fail(TRT_ERR_INVALID, "trt_render_multi: the same scene handle twice");
        devs[i] = scenes[i]->device;
    }
    if (n > kMaxPeers)
        return fail(TRT_ERR_LIMIT, "trt_render_multi: more than 16 scenes");

    // NCCL clique for this device list (created once per process and list)
    std::vector<ncclComm_t> *comms = nullptr;
    if (!peer)
    {
        std::vector<int> sorted = devs;
        std::sort(sorted.begin(), sorted.end());
        if (std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end())
            return fail(TRT_ERR_INVALID, "trt_render_multi: two scenes on one device need TRT_RENDER_PEER_REDUCE (NCCL "
                                         "refuses duplicate devices)");
        std::lock_guard<std::mutex> lock(g_ncclMutex);
        Nccl &nc = nccl();
        if (!nc.error.empty())
            return fail(TRT_ERR_NCCL, "trt_render_multi: " + nc.error);
        auto it = nc.comms.find(devs);
        if (it == nc.comms.end())
        {
            std::vector<ncclComm_t> c(n);
            const ncclResult_t r = nc.CommInitAll(c.data(), n, devs.data());
            if (r != ncclSuccess)
                return fail(TRT_ERR_NCCL, std::string("ncclCommInitAll: ") + nc.GetErrorString(r));
            it = nc.comms.emplace(devs, std::move(c)).first;
        }
        comms = &it->second;
    }
    else
    {
        // peer mappings: the first GPU reads the others' buffers
        TRT_CUDA(cudaSetDevice(s0->device));
        for (int i = 1; i < n; ++i)
        {
            if (devs[i] == s0->device)
                continue;
            int can = 0;
            TRT_CUDA(cudaDeviceCanAccessPeer(&can, s0->device, devs[i]));
            if (!can)
                return fail(TRT_ERR_CUDA, "trt_render_multi: device " + std::to_string(s0->device) + " cannot map device " +
                                              std::to_string(devs[i]) + "'s memory (no NVLink / P2P path)");
            const cudaError_t e = cudaDeviceEnablePeerAccess(devs[i], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(TRT_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            cudaGetLastError();
        }
    }

    // every GPU: zero its buffer, render its share of the samples (one host thread each: the depth loop synchronises
    // with its own device), then record "my buffer is complete"
    const int total = p.sample_end - p.sample_begin;
    std::vector<int> rcs(n, TRT_OK);
    std::vector<std::string> errs(n);
    std::vector<cudaEvent_t> done(n, nullptr);
    int rc = TRT_OK;
    for (int i = 0; i < n && rc == TRT_OK; ++i)
    {
        if (cudaSetDevice(devs[i]) != cudaSuccess || ensureFrameBuffers(scenes[i]) != TRT_OK ||
            cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess)
            rc = fail(TRT_ERR_CUDA, "trt_render_multi: per-device set-up failed on device " + std::to_string(devs[i]));
    }
    auto cleanup = [&] {
        for (int i = 0; i < n; ++i)
            if (done[i])
            {
                cudaSetDevice(devs[i]);
                cudaEventDestroy(done[i]);
            }
    };
    if (rc != TRT_OK)
    {
        cleanup();
        return rc;
    }
    TRT_CUDA(cudaSetDevice(s0->device));
    TRT_CUDA(cudaEventRecord(s0->ev[0], s0->stream));
    const bool timing = getenv("TRT_MULTI_TIMING") != nullptr; // host-clock phases on stderr (diagnostics)
    const auto t_start = std::chrono::steady_clock::now();
    auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
    std::vector<double> t_done(n, 0.0);
    auto work = [&](int i) {
        trt_scene *s = scenes[i];
        trt_render_params q = p;
        // contiguous shares whose sizes differ by at most one and tile [sample_begin, sample_end)
        const int base = total / n, rem = total % n;
        q.sample_begin = p.sample_begin + i * base + std::min(i, rem);
        q.sample_end = q.sample_begin + base + (i < rem ? 1 : 0);
        q.flags &= ~TRT_RENDER_PEER_REDUCE;
        cudaError_t e = cudaSetDevice(s->device);
        if (e == cudaSuccess)
            e = cudaMemsetAsync(s->d_frame_accum, 0, count * sizeof(double), s->stream);
        if (e != cudaSuccess)
        {
            rcs[i] = TRT_ERR_CUDA, errs[i] = cudaGetErrorString(e);
            return;
        }
        rcs[i] = renderAccumulate(s, q, s->d_frame_accum, s->stream);
        if (rcs[i] != TRT_OK)
            errs[i] = trt_last_error(); // thread-local: carried back to the caller's thread below
        else if ((e = cudaEventRecord(done[i], s->stream)) != cudaSuccess)
            rcs[i] = TRT_ERR_CUDA, errs[i] = cudaGetErrorString(e);
        t_done[i] = since();
    };
    {
        std::vector<std::thread> threads;
        for (int i = 1; i < n; ++i)
            threads.emplace_back(work, i);
        work(0);
        for (auto &t : threads)
            t.join();
    }
    for (int i = 0; i < n; ++i)
        if (rcs[i] != TRT_OK)
        {
            cleanup();
            return fail(rcs[i], "trt_render_multi: device " + std::to_string(devs[i]) + ": " + errs[i]);
        }

    TRT_CUDA(cudaSetDevice(s0->device));
    if (!peer)
    {
        Nccl &nc = nccl();
        std::lock_guard<std::mutex> lock(g_ncclMutex);
        ncclResult_t r = nc.GroupStart();
        for (int i = 0; i < n && r == ncclSuccess; ++i)
        {
            cudaSetDevice(devs[i]);
            r = nc.Reduce(scenes[i]->d_frame_accum, scenes[i]->d_frame_accum, count, ncclDouble, ncclSum, 0, (*comms)[i],
                          scenes[i]->stream);
        }
        const ncclResult_t r2 = nc.GroupEnd();
        if (r != ncclSuccess || r2 != ncclSuccess)
        {
            cleanup();
            return fail(TRT_ERR_NCCL, std::string("ncclReduce: ") + nc.GetErrorString(r != ncclSuccess ? r : r2));
        }
        TRT_CUDA(cudaSetDevice(s0->device));
        if ((rc = resolveImage(s0, s0->d_frame_accum, p.spp, s0->d_frame_image, s0->d_frame_rgb8, s0->stream)))
        {
            cleanup();
            return rc;
        }
    }
    else
    {
        PeerBuffers src{};
        src.n = n;
        for (int i = 0; i < n; ++i)
        {
            src.p[i] = scenes[i]->d_frame_accum;
            if (i > 0)
                TRT_CUDA(cudaStreamWaitEvent(s0->stream, done[i], 0));
        }
        const size_t threads = (count + 1) / 2;
        k_peer_reduce_resolve<<<(unsigned)((threads + 255) / 256), 256, 0, s0->stream>>>(src, count, p.spp, s0->d_frame_accum,
                                                                                         s0->d_frame_image, s0->d_frame_rgb8);
        TRT_CUDA(cudaGetLastError());
        s0->stats.kernel_launches++;
    }
    TRT_CUDA(cudaEventRecord(s0->ev[1], s0->stream));
    const double t_reduce_enqueued = since();
    if (image_rgb)
        TRT_CUDA(cudaMemcpyAsync(image_rgb, s0->d_frame_image, count * sizeof(double), cudaMemcpyDeviceToHost, s0->stream));
    if (rgb8)
        TRT_CUDA(cudaMemcpyAsync(rgb8, s0->d_frame_rgb8, count, cudaMemcpyDeviceToHost, s0->stream));
    TRT_CUDA(cudaStreamSynchronize(s0->stream));
    // the other GPUs' streams are idle by now (NCCL: their part of the reduce; peer: their buffers were read)
    for (int i = 1; i < n; ++i)
    {
        cudaSetDevice(devs[i]);
        cudaStreamSynchronize(scenes[i]->stream);
    }
    TRT_CUDA(cudaSetDevice(s0->device));
    float ms = 0;
    TRT_CUDA(cudaEventElapsedTime(&ms, s0->ev[0], s0->ev[1]));
    s0->stats.last_render_ms = ms;
    if (timing)
    {
        std::fprintf(stderr, "trt_render_multi[%s, %d GPUs]: renders done at", peer ? "peer" : "nccl", n);
        for (int i = 0; i < n; ++i)
            std::fprintf(stderr, " %.2f", t_done[i]);
        std::fprintf(stderr, " ms; reduce + resolve enqueued at %.2f; all done at %.2f; device time %.2f ms\n", t_reduce_enqueued,
                     since(), ms);
    }
    cleanup();
    return TRT_OK;
}

// ------------------------------------------------------------------------------------------- checkpoint file format
struct CheckpointHeader
{
    char magic[8]; // "TRTACCUM"
    uint32_t version, width, height, channels;
    int32_t samples_done, spp, max_depth, _pad;
    uint64_t seed, payload_bytes, checksum; // FNV-1a 64 of the payload
};
uint64_t fnv1a(const void *data, size_t n, uint64_t h = 1469598103934665603ull)
{
    const unsigned char *p = static_cast<const unsigned char *>(data);
    for (size_t i = 0; i < n; ++i)
        h = (h ^ p[i]) * 1099511628211ull;
    return h;
}
} // namespace
} // namespace trt

using namespace trt;

extern "C"
{
int trt_render_multi(trt_scene *const *scenes, int32_t n, const trt_render_params *p, double *image_rgb, uint8_t *rgb8)
{
    if (!scenes || n < 1 || !p || !scenes[0])
        return fail(TRT_ERR_INVALID, "trt_render_multi: null argument / no scene");
    if (p->spp < 1 || p->sample_begin < 0 || p->sample_end < p->sample_begin || p->sample_end > p->spp || p->max_depth < 0)
        return fail(TRT_ERR_INVALID, "trt_render_multi: bad sample range / spp / max_depth");
    try
    {
        return renderMulti(scenes, n, *p, image_rgb, rgb8);
    }
    catch (const std::exception &e)
    {
        return fail(TRT_ERR_INVALID, std::string("trt_render_multi: ") + e.what());
    }
}

int trt_trace_closest_multi(trt_scene *const *scenes, int32_t k, const float *rays6, size_t n, int32_t *tri_id, float *t,
                            uint32_t flags)
{
    if (!scenes || k < 1 || (!rays6 && n))
        return fail(TRT_ERR_INVALID, "trt_trace_closest_multi: null argument / no scene");
    if (flags & TRT_TRACE_DEVICE_PTRS)
        return fail(TRT_ERR_INVALID, "trt_trace_closest_multi: host pointers only");
    for (int i = 0; i < k; ++i)
    {
        if (!scenes[i])
            return fail(TRT_ERR_INVALID, "trt_trace_closest_multi: null scene");
        for (int j = 0; j < i; ++j)
            if (scenes[j] == scenes[i])
                return fail(TRT_ERR_INVALID, "trt_trace_closest_multi: the same scene handle twice");
    }
    std::vector<int> rcs(k, TRT_OK);
    std::vector<std::string> errs(k);
    auto work = [&](int i) {
        const size_t lo = n * (size_t)i / (size_t)k, hi = n * (size_t)(i + 1) / (size_t)k; // contiguous shards that tile [0, n)
        if (hi == lo)
            return;
        rcs[i] = trt_trace_closest(scenes[i], rays6 + lo * 6, hi - lo, tri_id ? tri_id + lo : nullptr, t ? t + lo : nullptr, flags);
        if (rcs[i] != TRT_OK)
            errs[i] = trt_last_error(); // thread-local: carried back to the caller's thread below
    };
    try
    {
        std::vector<std::thread> threads;
        for (int i = 1; i < k; ++i)
            threads.emplace_back(work, i);
        work(0);
        for (auto &th : threads)
            th.join();
    }
    catch (const std::exception &e)
    {
        return fail(TRT_ERR_LIMIT, std::string("trt_trace_closest_multi: ") + e.what());
    }
    for (int i = 0; i < k; ++i)
        if (rcs[i] != TRT_OK)
            return fail(rcs[i], "trt_trace_closest_multi: scene " + std::to_string(i) + ": " + errs[i]);
    return TRT_OK;
}

double *trt_accum_create(trt_scene *s)
{
    if (!s)
    {
        setLastError("trt_accum_create: null scene");
        return nullptr;
    }
    const size_t bytes = (size_t)s->width * s->height * 3 * sizeof(double);
    double *p = nullptr;
    if (cudaSetDevice(s->device) != cudaSuccess || cudaMalloc((void **)&p, bytes) != cudaSuccess ||
        cudaMemset(p, 0, bytes) != cudaSuccess)
    {
        setLastError(std::string("trt_accum_create: ") + cudaGetErrorString(cudaGetLastError()));
        cudaFree(p);
        return nullptr;
    }
    return p;
}

void trt_accum_destroy(trt_scene *s, double *d_accum)
{
    if (s && d_accum)
    {
        cudaSetDevice(s->device);
        cudaFree(d_accum);
    }
}

int trt_accum_save(trt_scene *s, const double *d_accum, int32_t samples_done, int32_t spp, uint64_t seed, int32_t max_depth,
                   const char *path)
{
    if (!s || !d_accum || !path || samples_done < 0 || spp < 1 || samples_done > spp)
        return fail(TRT_ERR_INVALID, "trt_accum_save: bad argument");
    TRT_CUDA(cudaSetDevice(s->device));
    const size_t n = (size_t)s->width * s->height * 3;
    std::vector<double> host;
    try
    {
        host.resize(n);
    }
    catch (const std::exception &)
    {
        return fail(TRT_ERR_LIMIT, "trt_accum_save: out of host memory");
    }
    TRT_CUDA(cudaMemcpy(host.data(), d_accum, n * sizeof(double), cudaMemcpyDeviceToHost)); // waits for work in flight
    CheckpointHeader h;
    std::memset(&h, 0, sizeof h);
    std::memcpy(h.magic, "TRTACCUM", 8);
    h.version = 1, h.width = (uint32_t)s->width, h.height = (uint32_t)s->height, h.channels = 3;
    h.samples_done = samples_done, h.spp = spp, h.max_depth = max_depth, h.seed = seed;
    h.payload_bytes = n * sizeof(double);
    h.checksum = fnv1a(host.data(), n * sizeof(double));
    // written beside the target and renamed over it: a crash mid-write never leaves a half checkpoint under `path`
    const std::string tmp = std::string(path) + ".tmp";
    FILE *f = std::fopen(tmp.c_str(), "wb");
    if (!f)
        return fail(TRT_ERR_IO, "trt_accum_save: cannot open " + tmp);
    const bool ok = std::fwrite(&h, sizeof h, 1, f) == 1 && std::fwrite(host.data(), sizeof(double), n, f) == n;
    if (std::fclose(f) != 0 || !ok || std::rename(tmp.c_str(), path) != 0)
    {
        std::remove(tmp.c_str());
        return fail(TRT_ERR_IO, std::string("trt_accum_save: cannot write ") + path);
    }
    return TRT_OK;
}

int trt_accum_load(trt_scene *s, const char *path, double *d_accum, int32_t *samples_done, int32_t *spp, uint64_t *seed,
                   int32_t *max_depth)
{
    if (!s || !d_accum || !path)
        return fail(TRT_ERR_INVALID, "trt_accum_load: bad argument");
    FILE *f = std::fopen(path, "rb");
    if (!f)
        return fail(TRT_ERR_IO, std::string("trt_accum_load: cannot open ") + path);
    CheckpointHeader h;
    const size_t n = (size_t)s->width * s->height * 3;
    std::vector<double> host;
    std::string err;
    if (std::fread(&h, sizeof h, 1, f) != 1 || std::memcmp(h.magic, "TRTACCUM", 8) != 0)
        err = "not a checkpoint file";
    else if (h.version != 1)
        err = "unknown checkpoint version";
    else if (h.width != (uint32_t)s->width || h.height != (uint32_t)s->height || h.channels != 3 ||
             h.payload_bytes != n * sizeof(double))
        err = "checkpoint was made for another frame size";
    else if (h.samples_done < 0 || h.spp < 1 || h.samples_done > h.spp)
        err = "corrupt header";
    else
    {
        try
        {
            host.resize(n);
        }
        catch (const std::exception &)
        {
            err = "out of host memory";
        }
        char extra;
        if (err.empty() && (std::fread(host.data(), sizeof(double), n, f) != n || std::fread(&extra, 1, 1, f) != 0))
            err = "truncated or oversized file";
        else if (err.empty() && fnv1a(host.data(), n * sizeof(double)) != h.checksum)
            err = "checksum mismatch";
    }
    std::fclose(f);
    if (!err.empty())
        return fail(TRT_ERR_IO, std::string("trt_accum_load: ") + path + ": " + err);
    TRT_CUDA(cudaSetDevice(s->device));
    TRT_CUDA(cudaMemcpy(d_accum, host.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    if (samples_done)
        *samples_done = h.samples_done;
    if (spp)
        *spp = h.spp;
    if (seed)
        *seed = h.seed;
    if (max_depth)
        *max_depth = h.max_depth;
    return TRT_OK;
}

int trt_shade(trt_scene *s, const float *rays6, const int32_t *tri_id, const float *t, size_t n, const trt_shade_params *p,
              float *radiance3)
{
    if (!s || !p || ((!rays6 || !tri_id || !t || !radiance3) && n))
        return fail(TRT_ERR_INVALID, "trt_shade: null argument");
    if (p->max_depth < 0 || p->sample < 0)
        return fail(TRT_ERR_INVALID, "trt_shade: negative sample / max_depth");
    if (n == 0)
        return TRT_OK;
    for (size_t i = 0; i < n; ++i)
        if (tri_id[i] >= s->view.n_tris)
            return fail(TRT_ERR_INVALID, "trt_shade: tri_id[" + std::to_string(i) + "] is not a triangle of this scene");
    TRT_CUDA(cudaSetDevice(s->device));
    float *d_r = nullptr, *d_t = nullptr, *d_o = nullptr;
    int32_t *d_i = nullptr;
    auto cleanup = [&] { cudaFree(d_r), cudaFree(d_t), cudaFree(d_o), cudaFree(d_i); };
    cudaError_t e = cudaMalloc((void **)&d_r, n * 24);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&d_t, n * 4);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&d_i, n * 4);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&d_o, n * 12);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_r, rays6, n * 24, cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_t, t, n * 4, cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_i, tri_id, n * 4, cudaMemcpyHostToDevice, s->stream);
    int rc = TRT_OK;
    if (e == cudaSuccess)
        rc = shadeBatch(s, d_r, d_i, d_t, n, *p, d_o, s->stream);
    if (e == cudaSuccess && rc == TRT_OK)
        e = cudaMemcpyAsync(radiance3, d_o, n * 12, cudaMemcpyDeviceToHost, s->stream);
    const cudaError_t es = cudaStreamSynchronize(s->stream);
    cleanup();
    if (rc != TRT_OK)
        return rc;
    if (e != cudaSuccess || es != cudaSuccess)
        return fail(TRT_ERR_CUDA, std::string("trt_shade: ") + cudaGetErrorString(e != cudaSuccess ? e : es));
    return TRT_OK;
}
}
