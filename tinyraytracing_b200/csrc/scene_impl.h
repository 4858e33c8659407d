// trt_scene: the device-resident scene behind the opaque handle of include/trt.h (internal).
#pragma once
#include "accel.h"

#include <string>
#include <vector>

namespace trt
{
void setLastError(const std::string &s);

#define TRT_CUDA(call)                                                                                              \
    do                                                                                                              \
    {                                                                                                               \
        cudaError_t e__ = (call);                                                                                   \
        if (e__ != cudaSuccess)                                                                                     \
        {                                                                                                           \
            trt::setLastError(std::string(#call) + ": " + cudaGetErrorString(e__));                                 \
            return TRT_ERR_CUDA;                                                                                    \
        }                                                                                                           \
    } while (0)

struct Wavefront; // wavefront.cu
}

struct trt_scene
{
    trt_scene() = default;
    trt_scene(const trt_scene &) = delete;
    trt_scene &operator=(const trt_scene &) = delete;
    ~trt_scene(); // capi.cu: frees everything below, so every early return of trt_scene_create cleans up
    int device = -1;
    int sm_count = 148;
    trt::SceneView view{};
    // every cudaMalloc owned by the scene, with what trt_scene_replicate needs to rebuild it on another device:
    // view_offset = byte offset of the SceneView pointer member that refers to it (kNotInView: a texture's pixels,
    // referred to by texture_index of host_textures / the device texture table instead)
    struct Alloc
    {
        void *p;
        size_t bytes, view_offset;
        int texture_index;
    };
    static constexpr size_t kNotInView = ~(size_t)0;
    std::vector<Alloc> allocations;
    std::vector<trt::DeviceTexture> host_textures; // the device texture table as uploaded (device pointers inside)
    cudaStream_t stream = nullptr;   // library-owned stream for the blocking entry points
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    // pinned staging for pageable host buffers (two chunks in flight)
    void *stage_in[2] = {nullptr, nullptr}, *stage_out[2] = {nullptr, nullptr};
    float *d_rays[2] = {nullptr, nullptr};
    int32_t *d_id[2] = {nullptr, nullptr};
    float *d_t[2] = {nullptr, nullptr};
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    size_t chunk_rays = 0;
    unsigned int *d_counter = nullptr; // ray-pool cursors of the persistent kernels: a ring, one slot per launch in flight
    unsigned int counter_slot = 0;
    bool closest_plain = false; // fixed batches use the plain thread-per-ray kernel (tiny scenes: every ray is short, trace.cu)
    bool shadow_stop = false; // the wavefront walks with the early stop of occluded light samples (wavefront.cu: WalkRays)
    int persistent_blocks_per_sm = 1, pooled_blocks_per_sm = 1;
    trt::Wavefront *wf = nullptr;
    // frame buffers of trt_resolve / trt_render, allocated on first use and kept: a cudaMalloc / cudaFree pair per
    // call synchronises the device and costs more than the resolve kernel itself on small frames
    double *d_frame_image = nullptr, *d_frame_accum = nullptr;
    uint8_t *d_frame_rgb8 = nullptr;
    trt_stats stats{};
    int width = 0, height = 0;
};

namespace trt
{
// trace.cu
int launchClosest(trt_scene *s, const float *d_rays6, size_t n, int32_t *d_id, float *d_t, uint32_t flags,
                  cudaStream_t stream);
int launchClosestCounters(trt_scene *s, const float *d_rays6, size_t n, unsigned long long *d_out4, cudaStream_t stream);
int launchHitAttributes(trt_scene *s, const float *d_rays6, const int32_t *d_id, const float *d_t, size_t n,
                        float *d_hitp3, float *d_pn3, cudaStream_t stream);
// wavefront.cu
int renderAccumulate(trt_scene *s, const trt_render_params &p, double *d_accum, cudaStream_t stream);
int resolveImage(trt_scene *s, const double *d_accum, int spp, double *d_image, uint8_t *d_rgb8, cudaStream_t stream);
int shadeBatch(trt_scene *s, const float *d_rays6, const int32_t *d_id, const float *d_t, size_t n, const trt_shade_params &p,
               float *d_radiance3, cudaStream_t stream);
void destroyWavefront(trt_scene *s);
} // namespace trt
