// Fixed-batch closest-hit kernels (BASELINE config 2): one thread per ray over the device-resident scene.
#include "scene_impl.h"
#include "traverse.cuh"
#include "barycentric.cuh"

#include <algorithm>

namespace trt
{
namespace
{
constexpr int kTraceBlock = 128;

__device__ __forceinline__ void loadRay(const float *rays6, size_t i, float3 &S, float3 &d)
{
    // 24-byte records: three aligned 8-byte loads per ray
    const float2 *p = reinterpret_cast<const float2 *>(rays6 + i * 6);
    const float2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    S = f3(a.x, a.y, b.x);
    d = f3(b.y, c.x, c.y);
}

template <int MODE> // 0: ordered + pruned reference topology, 1: exhaustive reference walk, 2: fast layout
__global__ void __launch_bounds__(kTraceBlock) k_closest(SceneView sv, const float *__restrict__ rays6, size_t n,
                                                         int32_t *__restrict__ out_id, float *__restrict__ out_t)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    float3 S, d;
    loadRay(rays6, i, S, d);
    Hit hit;
    if (MODE == 1)
        traceRefTopology<true>(sv, S, d, hit);
    else if (MODE == 0)
        traceRefTopology<false>(sv, S, d, hit);
    else
        traceClosest(sv, S, d, hit);
    if (out_id)
        out_id[i] = hit.id;
    if (out_t)
        out_t[i] = hit.t;
}

struct BatchRays
{
    const float *rays6;
    int32_t *out_id;
    float *out_t;
    static constexpr bool kHasBound = false; // every ray searches the whole of [0.0005, INF]
    __device__ __forceinline__ unsigned int locate(unsigned int i) const { return i; }
    __device__ __forceinline__ void load(unsigned int i, float3 &S, float3 &d, float &tmax) const
    {
        loadRay(rays6, i, S, d);
        tmax = TRT_INF;
    }
    __device__ __forceinline__ bool bounded(unsigned int) const { return false; }
    __device__ __forceinline__ bool canStop(const SceneView &, unsigned int, const Hit &, float3, float3) const { return false; }
    __device__ __forceinline__ void store(unsigned int i, const Hit &h) const
    {
        if (out_id)
            out_id[i] = h.id;
        if (out_t)
            out_t[i] = h.t;
    }
    __device__ __forceinline__ void storeFast(const SceneView &sv, unsigned int i, Hit h) const
    {
        h.id = (h.id >= 0) ? __ldg(sv.fast_orig + h.id) : -1; // fast index -> post-build index
        store(i, h);
    }
};

// 9 CTAs per SM (56 registers): the walker is latency-bound, see k_walk in wavefront.cu
#ifndef TRT_WALK_CTAS
#define TRT_WALK_CTAS 9
#endif
template <bool POOLED>
__global__ void __launch_bounds__(kTraceBlock, TRT_WALK_CTAS) k_closest_persistent(SceneView sv, const float *__restrict__ rays6, unsigned int n,
                                                                    int32_t *__restrict__ out_id, float *__restrict__ out_t,
                                                                    unsigned int *counter)
{
    BatchRays r{rays6, out_id, out_t};
    walkPersistent<POOLED>(sv, r, n, counter);
}

// Work counters of the fast layout's walk (design evaluation / reporting): sums over the batch.
__global__ void __launch_bounds__(kTraceBlock) k_closest_counters(SceneView sv, const float *__restrict__ rays6, size_t n,
                                                                  unsigned long long *__restrict__ out4)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    TraceCounters c{0, 0, 0, 0};
    if (i < n)
    {
        float3 S, d;
        loadRay(rays6, i, S, d);
        Hit hit;
        if (!needsStrictWalk(sv, S, d))
            traceWide<true>(sv, S, d, hit, &c);
    }
    uint32_t v[4] = {c.nodes, c.boxes, c.leaves, c.tris};
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        uint32_t x = v[k];
        for (int o = 16; o > 0; o >>= 1)
            x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x)
            atomicAdd(out4 + k, (unsigned long long)x);
    }
}

__global__ void __launch_bounds__(128) k_hit_attributes(SceneView sv, const float *__restrict__ rays6,
                                                        const int32_t *__restrict__ id, const float *__restrict__ t,
                                                        size_t n, float *__restrict__ hitp3, float *__restrict__ pn3)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    float3 P = f3(0.f, 0.f, 0.f), pn = f3(0.f, 0.f, 0.f); // HitRecord defaults (bvh.h:11-13)
    const int tri = id[i];
    if (tri >= 0)
    {
        float3 S, d;
        loadRay(rays6, i, S, d);
        P = S + d * t[i]; // bvh.cpp:191
        float bx, by, bz;
        baryLeastSquares(sv.tri_v + (size_t)tri * 9, P, bx, by, bz);
        pn = shadingNormal(sv.tri_shade[tri].vn, bx, by, bz);
    }
    if (hitp3)
        hitp3[i * 3] = P.x, hitp3[i * 3 + 1] = P.y, hitp3[i * 3 + 2] = P.z;
    if (pn3)
        pn3[i * 3] = pn.x, pn3[i * 3 + 1] = pn.y, pn3[i * 3 + 2] = pn.z;
}
} // namespace

int launchClosest(trt_scene *s, const float *d_rays6, size_t n, int32_t *d_id, float *d_t, uint32_t flags,
                  cudaStream_t stream)
{
    if (n == 0)
        return TRT_OK;
    const unsigned grid = (unsigned)((n + kTraceBlock - 1) / kTraceBlock);
    if (flags & TRT_TRACE_EXHAUSTIVE)
        k_closest<1><<<grid, kTraceBlock, 0, stream>>>(s->view, d_rays6, n, d_id, d_t);
    else if (flags & TRT_TRACE_REFTOPO)
        k_closest<0><<<grid, kTraceBlock, 0, stream>>>(s->view, d_rays6, n, d_id, d_t);
    // Which walker serves a fixed batch by default: on a scene of a few nodes (the 26-triangle Cornell shell: 7) every ray
    // is two or three node visits long, nothing is left for the persistent walker's refill machinery to balance, and the
    // plain thread-per-ray kernel is faster (back, 16 Mi rays: 16.6 against 16.0 Grays/s; 1 Mi rays: 14.5 against 12.9).
    // From veach-mis (600 nodes) up the persistent walker wins or ties (15.2 against 13.9; staircase 6.6 = 6.7), and in
    // the wavefront it wins everywhere (staircase render 90 against 106 ms).  TRT_TRACE_PERSISTENT forces it.
    else if ((flags & TRT_TRACE_PLAIN) || (s->closest_plain && !(flags & (TRT_TRACE_POOLED | TRT_TRACE_PERSISTENT))) ||
             n > 0x7fffffffull) // the walker's ray tokens are 31-bit (top bit: class-1 mark)
        k_closest<2><<<grid, kTraceBlock, 0, stream>>>(s->view, d_rays6, n, d_id, d_t);
    else
    {
        constexpr unsigned kCursorRing = 64; // launches of one scene that may be in flight on different streams at once
        if (!s->d_counter)
        {
            TRT_CUDA(cudaMalloc((void **)&s->d_counter, kCursorRing * sizeof(unsigned int)));
            TRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->persistent_blocks_per_sm,
                                                                   k_closest_persistent<false>, kTraceBlock, 0));
            TRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->pooled_blocks_per_sm, k_closest_persistent<true>,
                                                                   kTraceBlock, 0));
        }
        // every launch gets its own cursor: two traces enqueued on different streams overlap on the device and would
        // otherwise advance / reset each other's
        unsigned int *cursor = s->d_counter + (s->counter_slot++ % kCursorRing);
        TRT_CUDA(cudaMemsetAsync(cursor, 0, 4, stream));
        const bool pooled = (flags & TRT_TRACE_POOLED) != 0;
        const int bps = pooled ? s->pooled_blocks_per_sm : s->persistent_blocks_per_sm;
        const unsigned pgrid = (unsigned)std::min<size_t>((size_t)s->sm_count * bps, grid);
        if (pooled)
            k_closest_persistent<true><<<pgrid, kTraceBlock, 0, stream>>>(s->view, d_rays6, (unsigned int)n, d_id, d_t, cursor);
        else
            k_closest_persistent<false><<<pgrid, kTraceBlock, 0, stream>>>(s->view, d_rays6, (unsigned int)n, d_id, d_t, cursor);
    }
    TRT_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    s->stats.rays_closest += n;
    return TRT_OK;
}

int launchClosestCounters(trt_scene *s, const float *d_rays6, size_t n, unsigned long long *d_out4, cudaStream_t stream)
{
    if (n == 0 || !s->view.use_wide)
        return TRT_OK;
    k_closest_counters<<<(unsigned)((n + kTraceBlock - 1) / kTraceBlock), kTraceBlock, 0, stream>>>(s->view, d_rays6, n, d_out4);
    TRT_CUDA(cudaGetLastError());
    return TRT_OK;
}

int launchHitAttributes(trt_scene *s, const float *d_rays6, const int32_t *d_id, const float *d_t, size_t n,
                        float *d_hitp3, float *d_pn3, cudaStream_t stream)
{
    if (n == 0)
        return TRT_OK;
    k_hit_attributes<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(s->view, d_rays6, d_id, d_t, n, d_hitp3, d_pn3);
    TRT_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return TRT_OK;
}
} // namespace trt
