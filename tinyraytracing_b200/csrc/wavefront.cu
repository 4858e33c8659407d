// Wavefront path tracer: the loop body main.cpp:88-108 + shade / nextRay / Sample / RR (pathTracing.cpp:3-209)
// as a sequence of kernels over a batch of paths held in HBM (SoA), one iteration per path depth:
//
//   k_raygen      (K1)  main.cpp:88-95 + Camera::getRay           -> ray[slot], queue = all slots
//   k_trace       (K2)  traverseBVH for the queue                  -> hit[slot]
//   k_shade       (K3)  shade(): emissive return, Kd/texture, per-light NEE sample (emits "shadow" rays),
//                       RR, nextRay/Sample -> next ray, throughput weight, next queue (warp-aggregated append)
//   k_shadow      (K4)  pathTracing.cpp:51-58: closest hit of the light sample ray; the sample counts only if
//                       the CLOSEST hit's material is the light's material (not an any-hit test)
//   k_accumulate  (K5)  L[slot] += throughput * sum(visible light samples, XML order); throughput *= weight
//   k_deposit           after the batch: accum[pixel] += sum over the batch's samples of L (double, fixed order)
//   k_resolve     (K6)  accum / spp (main.cpp:101) and the gamma-2.2 8-bit pack of imshow (main.cpp:30-38)
//
// A path's slot = (sample_in_batch * W*H + pixel): pixel / sample never need storing, each pixel-sample is
// owned by exactly one slot, so there are no floating-point atomics and the image is bit-reproducible for a
// given (seed, spp) whatever the batch size or GPU count (per-sample sums are added in double, sample order).
//
// Random numbers: Philox4x32-10, key = seed, counter = (pixel, sample, depth, slot >> 1) — twin of the
// oracle's generator (oracle/oracle.cpp), see DESIGN.md §5 for the slot table.
#include "scene_impl.h"
#include "traverse.cuh"
#include "barycentric.cuh"

#include <algorithm>
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace trt
{
namespace
{
constexpr int kBlock = 128;
// k_shade is bound by registers (112 unconstrained) and insensitive to its instruction count (ncu + A/B):
// 64-thread CTAs with at least 9 resident per SM cap it at 96 registers (48 bytes of spills) = 5 warps per scheduler
// instead of 4 — the register file is partitioned per scheduler, so 112 registers give 4 warps whatever the CTA size.
// Measured against 128 threads / 112 registers (k_shade's own device time, TRT_RENDER_PROFILE): back 2.73 -> 2.66 ms,
// veach-mis 22.4 -> 21.4 ms, staircase 32.2 -> 30.9 ms.
constexpr int kShadeBlock = 64, kShadeMinBlocks = 9;
constexpr int kMaxLights = 32;
constexpr float kPI = 3.1415926f; // pathtracing.h:11
constexpr float kPRR = 0.8f;      // pathtracing.h:12
enum RayType
{
    DIFFUSE = 0,
    SPECULAR = 1,
    TRANSMISSION = 2,
    INVALID = 3,
    CAMERA = 4
};
enum Slot
{
    S_JITTER_X = 0,
    S_JITTER_Y = 1,
    S_RR = 2,
    S_FRESNEL = 3,
    S_LOBE = 4,
    S_PHI = 5,
    S_THETA = 6,
    S_LIGHT0 = 8
};

struct WfBuffers
{
    float4 *ray_o, *ray_d; // ray_d.w = ray type (int bits) the path arrived with
    int32_t *hit_id;
    float *hit_t;
    float4 *thr, *L, *weight;
    uint32_t *nee_mask;
    int32_t *queue[2];
    float4 *sh_o, *sh_d; // sh_o.w = contribution index (slot*n_lights + light), sh_d.w = light material
    float4 *sh_contrib;  // [slot*n_lights + light]
    // [0],[1]: path queue sizes (ping-pong); [2],[3]: ray-pool cursors of the persistent trace / shadow kernels;
    // [kShadowCount + l]: shadow rays queued for light l (segment l of sh_o / sh_d starts at l * capacity)
    int32_t *counters;
    int32_t capacity; // paths per batch = length of one per-light shadow segment
};
constexpr int kPoolTrace = 2, kPoolShadow = 3, kShadowCount = 8, kNumCounters = 8 + 32;

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r)
    {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0, c1 = lo1, c2 = n2, c3 = lo0;
        k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
    }
    out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

struct Rng
{
    uint32_t k0, k1, pixel, sample, depth;
    // both uniforms of one block: u[0] from words (0,1), u[1] from words (2,3)
    __device__ __forceinline__ void block(uint32_t blk, double u[2]) const
    {
        uint32_t o[4];
        philox4x32_10(pixel, sample, depth, blk, k0, k1, o);
        u[0] = (double)(((uint64_t)o[0] << 21) | (uint64_t)(o[1] >> 11)) * (1.0 / 9007199254740992.0);
        u[1] = (double)(((uint64_t)o[2] << 21) | (uint64_t)(o[3] >> 11)) * (1.0 / 9007199254740992.0);
    }
};

__device__ __forceinline__ float3 xyz(float4 v) { return f3(v.x, v.y, v.z); }

// IEEE x / c for a finite c > 0, bit for bit, without the division's slow path on a zero dividend: FCHK sends
// 0 / c to a ~100-instruction subroutine (ncu, profiles/r01_render_shade_staircase.txt: 11 % of k_shade's warp
// instructions were that subroutine, fed by Ks = 0 and by colours with a zero channel).  +-0 / c = +-0 = x.
__device__ __forceinline__ float divByPositive(float x, float c)
{
    const bool zero = (x == 0.f);
    const float q = (zero ? 1.0f : x) / c;
    return zero ? x : q;
}
__device__ __forceinline__ float3 divByPositive(float3 v, float c)
{
    return f3(divByPositive(v.x, c), divByPositive(v.y, c), divByPositive(v.z, c));
}
__device__ __forceinline__ float4 xyzw(float3 v, float w) { return make_float4(v.x, v.y, v.z, w); }

// ------------------------------------------------------------------------------------------------ K1
__global__ void __launch_bounds__(kBlock) k_raygen(SceneView sv, WfBuffers wf, int n_paths, int sample0, uint64_t seed)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_paths)
        return;
    const int W = sv.cam.width, H = sv.cam.height, npix = W * H;
    const int pix = slot % npix, k = sample0 + slot / npix;
    const int i = pix / W, j = pix % W;
    Rng rng{(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)pix, (uint32_t)k, 0u};
    double u[2];
    rng.block(0, u);
    // main.cpp:88-93
    double x = double(j) / double(W - 1.0);
    double y = double(H - i) / double(H - 1.0);
    x += (u[0] - 0.5f) / double(W);
    y += (u[1] - 0.5f) / double(H);
    const float sx = (float)x, sy = (float)y;
    // camera.cpp:19-28
    const float3 d = normalize3(sv.cam.llc + sx * sv.cam.horizontal + sy * sv.cam.vertical - sv.cam.eye);
    wf.ray_o[slot] = xyzw(sv.cam.eye, 0.f);
    wf.ray_d[slot] = xyzw(d, __int_as_float(CAMERA));
    wf.thr[slot] = make_float4(1.f, 1.f, 1.f, 0.f);
    wf.L[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    wf.queue[0][slot] = slot;
    if (slot == 0)
    {
        wf.counters[0] = n_paths;
        wf.counters[1] = 0;
        wf.counters[kPoolTrace] = 0;
    }
}

// ------------------------------------------------------------------------------------------------ K2 / K4
struct QueueRays // closest-hit rays of the live paths; token = path slot
{
    WfBuffers wf;
    int qsel;
    __device__ __forceinline__ unsigned int locate(unsigned int i) const { return (unsigned int)wf.queue[qsel][i]; }
    __device__ __forceinline__ void load(unsigned int slot, float3 &S, float3 &d) const
    {
        S = xyz(wf.ray_o[slot]), d = xyz(wf.ray_d[slot]);
    }
    __device__ __forceinline__ void store(unsigned int slot, const Hit &h) const
    {
        wf.hit_id[slot] = h.id;
        wf.hit_t[slot] = h.t;
    }
    __device__ __forceinline__ void storeFast(const SceneView &sv, unsigned int slot, Hit h) const
    {
        h.id = (h.id >= 0) ? __ldg(sv.fast_orig + h.id) : -1; // fast index -> post-build index
        store(slot, h);
    }
};

// light-sample rays, one queue segment per light (neighbouring lanes: same light, nearby pixels); token = entry index
// in sh_o / sh_d (light * capacity + position, below 2^31 by the batch-size check of renderAccumulate)
struct ShadowRays
{
    WfBuffers wf;
    const TriShade *tri_shade;
    int n_lights;
    __device__ __forceinline__ unsigned int locate(unsigned int i) const
    {
        int l = 0;
        for (; l < n_lights - 1; ++l)
        {
            const unsigned int c = (unsigned int)wf.counters[kShadowCount + l];
            if (i < c)
                break;
            i -= c;
        }
        return (unsigned int)l * (unsigned int)wf.capacity + i;
    }
    __device__ __forceinline__ void load(unsigned int e, float3 &S, float3 &d) const
    {
        S = xyz(wf.sh_o[e]), d = xyz(wf.sh_d[e]);
    }
    // pathTracing.cpp:54-58: visible iff the closest hit's material is the light's material
    __device__ __forceinline__ void finish(unsigned int e, bool hit, int mtl) const
    {
        const bool visible = hit && mtl == __float_as_int(wf.sh_d[e].w);
        if (!visible)
            wf.sh_contrib[__float_as_int(wf.sh_o[e].w)] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ void store(unsigned int e, const Hit &h) const
    {
        finish(e, h.id >= 0, h.id >= 0 ? tri_shade[h.id].mtl : -1);
    }
    __device__ __forceinline__ void storeFast(const SceneView &sv, unsigned int e, const Hit &h) const
    {
        finish(e, h.id >= 0, h.id >= 0 ? __ldg(sv.fast_mtl + h.id) : -1); // one load from a dense 4-byte table
    }
};

// MODE 0: warp-persistent walker over the fast layout; 1: reference topology; 2: fast layout, plain thread per ray
template <int MODE, typename RAYS>
__device__ __forceinline__ void traceGridStride(const SceneView &sv, RAYS &rays, unsigned int n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        const unsigned int token = rays.locate((unsigned int)i);
        float3 S, d;
        rays.load(token, S, d);
        Hit hit;
        if (MODE == 1)
            traceRefTopology<false>(sv, S, d, hit);
        else
            traceClosest(sv, S, d, hit);
        rays.store(token, hit);
    }
}

template <int MODE>
__global__ void __launch_bounds__(kBlock) k_trace(SceneView sv, WfBuffers wf, int qsel)
{
    QueueRays r{wf, qsel};
    const unsigned int n = (unsigned int)wf.counters[qsel];
    if (MODE != 0)
        traceGridStride<MODE>(sv, r, n);
    else
        walkPersistent<false>(sv, r, n, reinterpret_cast<unsigned int *>(wf.counters + kPoolTrace));
}

// 9 resident CTAs per SM = 56 registers (24 bytes of spills) instead of 64 / 8 CTAs: the walker is latency-bound (ncu:
// 1.4 eligible warps per cycle), one more CTA of warps pays (staircase 72.1 -> 70.8 ms of k_shadow, veach-mis 32.6 -> 32.2);
// 10 CTAs = 48 registers spill 134 bytes and lose (83.6 ms).  k_trace needs 56 registers as it is.
template <int MODE>
__global__ void __launch_bounds__(kBlock, 9) k_shadow(SceneView sv, WfBuffers wf)
{
    ShadowRays r{wf, sv.tri_shade, sv.n_lights};
    unsigned int n = 0;
    for (int l = 0; l < sv.n_lights; ++l)
        n += (unsigned int)wf.counters[kShadowCount + l];
    if (MODE != 0)
        traceGridStride<MODE>(sv, r, n);
    else
        walkPersistent<false>(sv, r, n, reinterpret_cast<unsigned int *>(wf.counters + kPoolShadow));
}

// ------------------------------------------------------------------------------------------------ K3
// pathTracing.cpp:111-145
__device__ __forceinline__ float3 sampleLobe(float3 direction, int ray_type, double Ns, double u_phi, double u_theta)
{
    const double phi = u_phi * 2 * (double)kPI;
    double theta;
    if (ray_type == DIFFUSE)
        theta = asin(sqrt(u_theta));
    else
        theta = acos(pow(u_theta, (double)1 / (Ns + 1)));
    double st, ct, sp, cp;
    sincos(theta, &st, &ct);
    sincos(phi, &sp, &cp);
    const float3 sample = f3((float)(st * cp), (float)ct, (float)(st * sp));
    float3 front;
    if (fabsf(direction.x) > fabsf(direction.y))
        front = normalize3(f3(direction.z, 0.f, -direction.x));
    else
        front = normalize3(f3(0.f, -direction.z, direction.y));
    const float3 right = cross3(direction, front);
    return normalize3(((right * sample.x) + (direction * sample.y)) + (front * sample.z));
}

__device__ __forceinline__ float3 reflect3(float3 I, float3 N) { return I - N * dot3(N, I) * 2.0f; }

__device__ __forceinline__ void appendQueue(int32_t *counter, int32_t *queue, bool pred, int value)
{
    // warp-aggregated append: one atomic per warp
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0)
        return;
    const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader)
        base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred)
        queue[base + __popc(mask & ((1u << lane) - 1u))] = value;
}

__global__ void __launch_bounds__(kShadeBlock, kShadeMinBlocks) k_shade(SceneView sv, WfBuffers wf, int qsel, int depth, int max_depth,
                                                  int sample0, uint64_t seed)
{
  const int count = wf.counters[qsel];
  // warp-uniform grid-stride loop: the queue appends below are warp-collective
  for (int base = blockIdx.x * blockDim.x; base < count; base += gridDim.x * blockDim.x)
  {
    const int i = base + threadIdx.x;
    const bool active = i < count;
    bool survives = false;
    int slot = 0;
    uint32_t mask = 0;
    float4 weight = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active)
    {
        slot = wf.queue[qsel][i];
        const int tri = wf.hit_id[slot];
        const float4 rd4 = wf.ray_d[slot];
        const int via = __float_as_int(rd4.w);
        if (tri >= 0)
        {
            const TriShade ts = sv.tri_shade[tri];
            // material fields are fetched where they are used (L1-resident table) instead of holding the whole record
            // in registers across the light loop: k_shade's occupancy is register-bound
            const DeviceMaterial *mp = sv.materials + ts.mtl;
            if (mp->is_emissive)
            {
                // :9-12 returns the radiance; DIFFUSE / SPECULAR arrivals drop it (:87-94), camera and
                // TRANSMISSION arrivals keep it (main.cpp:101, :95-96)
                if (via == CAMERA || via == TRANSMISSION)
                {
                    const float4 T = wf.thr[slot];
                    float4 L = wf.L[slot];
                    const float3 rad = mp->radiance;
                    L.x += T.x * rad.x, L.y += T.y * rad.y, L.z += T.z * rad.z;
                    wf.L[slot] = L;
                }
            }
            else
            {
                const int npix = sv.cam.width * sv.cam.height;
                Rng rng{(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)(slot % npix), (uint32_t)(sample0 + slot / npix),
                        (uint32_t)depth};
                const float3 S = xyz(wf.ray_o[slot]), d = xyz(rd4);
                const float3 wi = -d;
                const float3 P = S + d * wf.hit_t[slot]; // bvh.cpp:191
                float bx, by, bz;
                baryLeastSquares(sv.tri_v + (size_t)tri * 9, P, bx, by, bz);
                const float3 pn = shadingNormal(ts.vn, bx, by, bz); // bvh.cpp:223-224
                float3 Kd = mp->Kd;
                const int m_texture = mp->texture;
                const float3 m_Ks = mp->Ks;
                const float m_Ns = mp->Ns;
                if (m_texture >= 0) // :17-26
                {
                    const double col = (ts.vt[0] * bx + ts.vt[2] * by) + ts.vt[4] * bz;
                    const double row = (ts.vt[1] * bx + ts.vt[3] * by) + ts.vt[5] * bz;
                    const double irow = row - floor(row), icol = col - floor(col);
                    const DeviceTexture tx = sv.textures[m_texture];
                    const int r = (int)(irow * tx.rows), c = (int)(icol * tx.cols);
                    const uint8_t *px = tx.bgr + ((size_t)r * tx.cols + c) * 3;
                    Kd = f3((float)((double)px[2] / 255), (float)((double)px[1] / 255), (float)((double)px[0] / 255));
                }

                // ---- direct illumination :33-75
                for (int li = 0; li < sv.n_lights; ++li)
                {
                    const DeviceLight lt = sv.lights[li];
                    double u0[2];
                    rng.block(4 + 2 * li, u0);
                    const double rnd = u0[0] * sv.first_light_area; // quirk A.5-1 (:38)
                    // first light triangle whose cumulative area exceeds rnd: the reference walks the list
                    // linearly (:40-42).  rnd is drawn from [0, area of the FIRST light), so the answer is usually the
                    // first triangle or none at all; otherwise gallop, then bisect (cumulative areas are non-decreasing)
                    const double *cum = sv.light_cum_area + lt.first_tri;
                    const int nt = lt.n_tris;
                    if (nt == 0)
                        continue;
                    int lo = 0;
                    if (!(rnd < __ldg(cum)))
                    {
                        if (!(rnd < __ldg(cum + nt - 1)))
                            continue; // no triangle selected: this light is not sampled at this vertex
                        int a = 0, b = 1; // !(rnd < cum[a]); looking for the first b with rnd < cum[b]
                        while (b < nt - 1 && !(rnd < __ldg(cum + b)))
                            a = b, b = min(2 * b + 1, nt - 1);
                        while (b - a > 1)
                        {
                            const int mid = (a + b) >> 1;
                            if (rnd < __ldg(cum + mid))
                                b = mid;
                            else
                                a = mid;
                        }
                        lo = b;
                    }
                    double u1[2];
                    rng.block(5 + 2 * li, u1);
                    const double rnd1 = u0[1], rnd2 = u1[0], rnd3 = u1[1];
                    const double rs = (rnd1 + rnd2) + rnd3;
                    const float p1 = (float)(rnd1 / rs), p2 = (float)(rnd2 / rs), p3 = (float)(rnd3 / rs);
                    const float *lv = sv.light_v + (size_t)(lt.first_tri + lo) * 9;
                    const float *ln = sv.light_vn + (size_t)(lt.first_tri + lo) * 9;
                    const float3 light_p = (f3(lv[0], lv[1], lv[2]) * p1 + f3(lv[3], lv[4], lv[5]) * p2) + f3(lv[6], lv[7], lv[8]) * p3;
                    const float3 light_n =
                        normalize3((f3(ln[0], ln[1], ln[2]) * p1 + f3(ln[3], ln[4], ln[5]) * p2) + f3(ln[6], ln[7], ln[8]) * p3);
                    const float3 wo = normalize3(light_p - P);
                    const float wo_pn = dot3(wo, pn);
                    if (!(wo_pn > 0.f)) // :60 — the sample cannot contribute: the shadow ray is not traced
                        continue;
                    const DeviceMaterial lm = sv.materials[lt.material];
                    const float pdf_light = (float)(double(1) / lm.area);
                    const float cos_theta_p = fabsf(dot3(wo, light_n));
                    const float cos_theta = fabsf(wo_pn / length3(pn));
                    const float3 diff = light_p - P;
                    const float len2 = dot3(diff, diff);
                    const float3 intensity = (((lm.radiance * cos_theta_p) * cos_theta) / len2) / pdf_light;
                    const float3 h = normalize3((wi + wo) * 0.5f);
                    const double cos_alpha = fmax((double)dot3(pn, h), 0.0);
                    // Ks == 0 (every diffuse material): Ks * (Ns+2) * pow(...) is +0 for any finite power, and the
                    // power is finite for cos_alpha in [0,1] and Ns >= 0 — the double-precision pow is skipped
                    const bool no_spec = (m_Ks.x == 0.f) && (m_Ks.y == 0.f) && (m_Ks.z == 0.f) && (m_Ns >= 0.f);
                    // the specular term of a diffuse material is (+0 * (Ns + 2)) * 0 / 2pi = +0 in every channel
                    float3 spec_term = f3(0.f, 0.f, 0.f);
                    if (!no_spec)
                        spec_term = divByPositive((m_Ks * (m_Ns + 2.0f)) * (float)pow(cos_alpha, (double)m_Ns), 2.0f * kPI);
                    const float3 brdf = divByPositive(Kd, kPI) + spec_term;
                    const float3 contrib = intensity * brdf;
                    const int cidx = slot * sv.n_lights + li;
                    wf.sh_contrib[cidx] = xyzw(contrib, 0.f);
                    {
                        // one queue segment per light, appended by the lanes that are converged here with a single
                        // atomic: neighbouring entries are neighbouring pixels aiming at the same light
                        cg::coalesced_group g = cg::coalesced_threads();
                        int at = 0;
                        if (g.thread_rank() == 0)
                            at = atomicAdd(&wf.counters[kShadowCount + li], (int)g.size());
                        at = g.shfl(at, 0);
                        const size_t e = (size_t)li * wf.capacity + at + g.thread_rank();
                        wf.sh_o[e] = xyzw(P, __int_as_float(cidx));
                        wf.sh_d[e] = xyzw(wo, __int_as_float(lt.material));
                    }
                    mask |= 1u << li;
                }

                // ---- indirect :78-99
                const bool may_bounce = (max_depth == 0) || (depth + 1 < max_depth);
                double ur[2];
                rng.block(1, ur);
                if (may_bounce && ur[0] < (double)kPRR) // RR(): :104-109
                {
                    // nextRay(): :147-209
                    const float3 ray_direction = d; // -wi
                    int type = INVALID;
                    float3 ndir = f3(0.f, 0.f, 0.f);
                    bool decided = false;
                    const float m_Ni = mp->Ni;
                    if (m_Ni > 1.f)
                    {
                        double n1, n2;
                        const double cos_in = (double)dot3(ray_direction, pn);
                        float3 normal;
                        if (cos_in > 0)
                            normal = -pn, n1 = (double)m_Ni, n2 = 1.0;
                        else
                            normal = pn, n1 = 1.0, n2 = (double)m_Ni;
                        const double q = (n1 - n2) / (n1 + n2);
                        const double rf0 = q * q;
                        const double a = (double)1.0f - fabs(cos_in);
                        const double fresnel = rf0 + ((double)1.0f - rf0) * (((a * a) * (a * a)) * a);
                        if (fresnel < ur[1])
                        {
                            const float eta = (float)(n1 / n2);
                            const float dv = dot3(normal, ray_direction);
                            const float k = 1.0f - eta * eta * (1.0f - dv * dv);
                            if (k >= 0.0f)
                                ndir = eta * ray_direction - (eta * dv + sqrtf(k)) * normal;
                            if (ndir.x != 0.f || ndir.y != 0.f || ndir.z != 0.f)
                                type = TRANSMISSION;
                            else
                                ndir = reflect3(ray_direction, normal), type = SPECULAR;
                            decided = true;
                        }
                    }
                    if (!decided)
                    {
                        const double Kd_len = (double)length3(mp->Kd), Ks_len = (double)length3(m_Ks);
                        const double kd = Kd_len / (Kd_len + Ks_len), ks = Ks_len / (Kd_len + Ks_len);
                        double ul[2], ut[2];
                        rng.block(2, ul);
                        const double p = ul[0];
                        if (p < kd)
                        {
                            rng.block(3, ut);
                            ndir = sampleLobe(pn, DIFFUSE, (double)m_Ns, ul[1], ut[0]);
                            type = DIFFUSE;
                        }
                        else if (m_Ns > 1.f && p < kd + ks)
                        {
                            rng.block(3, ut);
                            ndir = sampleLobe(reflect3(ray_direction, pn), SPECULAR, (double)m_Ns, ul[1], ut[0]);
                            type = SPECULAR;
                        }
                    }
                    if (type != INVALID) // the reference traces the INVALID ray too and discards it (:81-82)
                    {
                        const float3 w = (type == TRANSMISSION) ? mp->Tr : Kd; // SPECULAR also weights by Kd (:92-93)
                        weight = xyzw(divByPositive(w, kPRR), 0.f);
                        wf.ray_o[slot] = xyzw(P, 0.f);
                        wf.ray_d[slot] = xyzw(ndir, __int_as_float(type));
                        survives = true;
                    }
                }
            }
        }
        wf.nee_mask[slot] = mask;
        wf.weight[slot] = weight;
    }
    appendQueue(&wf.counters[qsel ^ 1], wf.queue[qsel ^ 1], survives, slot);
  }
}

// ------------------------------------------------------------------------------------------------ K5
__global__ void __launch_bounds__(kBlock) k_accumulate(SceneView sv, WfBuffers wf, int qsel)
{
  const int count = wf.counters[qsel];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
  {
    const int slot = wf.queue[qsel][i];
    uint32_t mask = wf.nee_mask[slot];
    float4 T = wf.thr[slot];
    if (mask)
    {
        float3 Ldir = f3(0.f, 0.f, 0.f);
        while (mask)
        {
            const int li = __ffs(mask) - 1;
            mask &= mask - 1;
            Ldir = Ldir + xyz(wf.sh_contrib[slot * sv.n_lights + li]); // L_dir += ..., lights in XML order (:70)
        }
        float4 L = wf.L[slot];
        L.x += T.x * Ldir.x, L.y += T.y * Ldir.y, L.z += T.z * Ldir.z;
        wf.L[slot] = L;
    }
    const float4 w = wf.weight[slot];
    T.x *= w.x, T.y *= w.y, T.z *= w.z;
    wf.thr[slot] = T;
  }
}

__global__ void k_reset_counters(WfBuffers wf, int qnext)
{
    // before k_shade of an iteration: the queue it fills, the shadow segments and the ray-pool cursors start at 0
    if (threadIdx.x == 0)
        wf.counters[qnext] = 0, wf.counters[kPoolTrace] = 0, wf.counters[kPoolShadow] = 0;
    if (threadIdx.x < 32)
        wf.counters[kShadowCount + threadIdx.x] = 0;
}

__global__ void __launch_bounds__(256) k_deposit(WfBuffers wf, double *accum, int npix, int samples_in_batch)
{
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix)
        return;
    double r = 0, g = 0, b = 0;
    for (int s = 0; s < samples_in_batch; ++s)
    {
        const float4 L = wf.L[(size_t)s * npix + pix];
        r += (double)L.x, g += (double)L.y, b += (double)L.z;
    }
    accum[(size_t)pix * 3 + 0] += r;
    accum[(size_t)pix * 3 + 1] += g;
    accum[(size_t)pix * 3 + 2] += b;
}

__global__ void __launch_bounds__(256) k_resolve(const double *accum, size_t n, int spp, double *image, uint8_t *rgb8)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const double v = accum[i] / (double)spp;
    if (image)
        image[i] = v;
    if (rgb8)
    {
        // main.cpp:34: (unsigned char)clamp(pow(v, 1.0f / 2.2f) * 255, 0.0, 255.0)
        double g = pow(v, (double)(1.0f / 2.2f)) * 255;
        g = (g < 0.0) ? 0.0 : g;
        g = (255.0 < g) ? 255.0 : g;
        rgb8[i] = (uint8_t)g;
    }
}
} // namespace

struct Wavefront
{
    WfBuffers buf{};
    int capacity = 0; // paths
    int n_lights = 0;
    std::vector<void *> allocs;
    static constexpr int kRing = 8;
    int32_t *h_ring = nullptr; // pinned: kRing snapshots of the device counters
    cudaEvent_t ring_ev[kRing] = {};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int blocks_trace = 1, blocks_shadow = 1, blocks_shade = 1;
    std::vector<cudaEvent_t> prof_ev; // TRT_RENDER_PROFILE: five timestamps per iteration, grown on demand
};

void destroyWavefront(trt_scene *s)
{
    if (!s->wf)
        return;
    for (void *p : s->wf->allocs)
        cudaFree(p);
    if (s->wf->h_ring)
        cudaFreeHost(s->wf->h_ring);
    for (cudaEvent_t e : s->wf->ring_ev)
        if (e)
            cudaEventDestroy(e);
    if (s->wf->ev0)
        cudaEventDestroy(s->wf->ev0), cudaEventDestroy(s->wf->ev1);
    for (cudaEvent_t e : s->wf->prof_ev)
        cudaEventDestroy(e);
    delete s->wf;
    s->wf = nullptr;
}

static int ensureWavefront(trt_scene *s, int paths)
{
    if (s->wf && s->wf->capacity >= paths)
        return TRT_OK;
    destroyWavefront(s);
    Wavefront *w = new Wavefront();
    s->wf = w;
    w->capacity = paths;
    w->n_lights = std::max(1, s->view.n_lights);
    const size_t N = (size_t)paths, NL = N * w->n_lights;
    auto alloc = [&](void **p, size_t bytes) -> int {
        TRT_CUDA(cudaMalloc(p, bytes));
        w->allocs.push_back(*p);
        return TRT_OK;
    };
    int rc;
    WfBuffers &b = w->buf;
    b.capacity = paths;
    if ((rc = alloc((void **)&b.ray_o, N * 16)) || (rc = alloc((void **)&b.ray_d, N * 16)) ||
        (rc = alloc((void **)&b.hit_id, N * 4)) || (rc = alloc((void **)&b.hit_t, N * 4)) ||
        (rc = alloc((void **)&b.thr, N * 16)) || (rc = alloc((void **)&b.L, N * 16)) ||
        (rc = alloc((void **)&b.weight, N * 16)) || (rc = alloc((void **)&b.nee_mask, N * 4)) ||
        (rc = alloc((void **)&b.queue[0], N * 4)) || (rc = alloc((void **)&b.queue[1], N * 4)) ||
        (rc = alloc((void **)&b.sh_o, NL * 16)) || (rc = alloc((void **)&b.sh_d, NL * 16)) ||
        (rc = alloc((void **)&b.sh_contrib, NL * 16)) ||
        (rc = alloc((void **)&b.counters, kNumCounters * 4)))
        return rc;
    TRT_CUDA(cudaMemset(b.counters, 0, kNumCounters * 4));
    TRT_CUDA(cudaMallocHost((void **)&w->h_ring, Wavefront::kRing * kNumCounters * 4));
    for (cudaEvent_t &e : w->ring_ev)
        TRT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    TRT_CUDA(cudaEventCreate(&w->ev0));
    TRT_CUDA(cudaEventCreate(&w->ev1));
    TRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w->blocks_trace, k_trace<0>, kBlock, 0));
    TRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w->blocks_shadow, k_shadow<0>, kBlock, 0));
    TRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w->blocks_shade, k_shade, kShadeBlock, 0));
    return TRT_OK;
}

int renderAccumulate(trt_scene *s, const trt_render_params &p, double *d_accum, cudaStream_t stream)
{
    const int W = s->width, H = s->height;
    const long long npix = (long long)W * H;
    if (s->view.n_lights > kMaxLights)
    {
        setLastError("more than 32 lights");
        return TRT_ERR_LIMIT;
    }
    // Paths in flight per batch.  Every depth iteration costs five launches and at least one wave of rays (~0.2 ms on
    // staircase) whatever the queue length, and path counts decay by 0.8 per depth, so large batches amortise the
    // long tail.  Measured: veach-mis 1280x720 256 spp 506.7 / 499.6 / 493.6 ms and staircase 1920x1080 128 spp
    // 2127 / 2102 / 2094 ms at 32 / 64 / 128 Mi paths.  Default 128 Mi paths (52 GB of path state with 6 lights:
    // HBM is 180 GB and the scene itself is megabytes), bounded by a third of the free memory.
    long long target = p.batch_paths > 0 ? p.batch_paths : (128ll << 20);
    if (p.batch_paths <= 0)
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
        {
            const long long per_path = 100 + 48ll * std::max(1, s->view.n_lights);
            const long long have = (long long)(s->wf ? (size_t)s->wf->capacity * per_path : 0);
            target = std::min(target, std::max(1ll << 20, ((long long)free_b / 3 + have) / per_path));
        }
        // slot * n_lights + light must fit 32-bit indices (checked below for explicit batch sizes)
        target = std::min(target, 0x7fffffffll / std::max(1, s->view.n_lights));
    }
    const int total_samples = p.sample_end - p.sample_begin;
    if (total_samples <= 0)
        return TRT_OK;
    int spb = (int)std::max(1ll, std::min((long long)total_samples, target / npix));
    if (npix * spb > 0x7fffffffll / std::max(1, s->view.n_lights))
    {
        setLastError("batch too large for 32-bit slot indices");
        return TRT_ERR_LIMIT;
    }
    int rc = ensureWavefront(s, (int)(npix * spb));
    if (rc)
        return rc;
    Wavefront *w = s->wf;
    const WfBuffers &b = w->buf;
    const int mode = (p.flags & TRT_RENDER_REFTOPO) ? 1 : ((p.flags & TRT_RENDER_PLAIN) ? 2 : 0);
    const int nl = s->view.n_lights;
    // The host never waits for an iteration it has just launched: every kernel reads its queue length from device
    // memory, and the host looks at the counters of iteration it - kLag to learn when the batch has died out.
    constexpr int kLag = 2;
    const unsigned grid_trace = (unsigned)(s->sm_count * std::max(1, w->blocks_trace));
    const unsigned grid_shadow = (unsigned)(s->sm_count * std::max(1, w->blocks_shadow));
    const unsigned grid_shade = (unsigned)(s->sm_count * 8);
    // k_shade: exactly the resident CTAs (grid-stride loop inside): a second, partly filled wave of CTAs would cost a
    // whole extra pass of ~40 us warp iterations
    const unsigned grid_kshade = (unsigned)(s->sm_count * std::max(1, w->blocks_shade));
    const bool profile = (p.flags & TRT_RENDER_PROFILE) != 0;
    size_t prof_used = 0;
    double prof_ms[4] = {0, 0, 0, 0};
    auto stamp = [&]() -> int { // a timestamp on the stream between two launches
        if (prof_used == w->prof_ev.size())
        {
            cudaEvent_t e;
            TRT_CUDA(cudaEventCreate(&e));
            w->prof_ev.push_back(e);
        }
        TRT_CUDA(cudaEventRecord(w->prof_ev[prof_used++], stream));
        return TRT_OK;
    };
    TRT_CUDA(cudaEventRecord(w->ev0, stream));
    for (int s0 = p.sample_begin; s0 < p.sample_end; s0 += spb)
    {
        const int ns = std::min(spb, p.sample_end - s0);
        const int n_paths = (int)(npix * ns);
        k_raygen<<<(unsigned)((n_paths + kBlock - 1) / kBlock), kBlock, 0, stream>>>(s->view, b, n_paths, s0, p.seed);
        s->stats.kernel_launches++;
        s->stats.paths += (uint64_t)n_paths;
        s->stats.rays_closest += (uint64_t)n_paths; // iteration 0 traces every path's camera ray
        int q = 0, consumed = 0;
        bool dead = false;
        auto consume = [&](int it) -> int { // counters as they stood after k_shade of iteration `it`
            TRT_CUDA(cudaEventSynchronize(w->ring_ev[it % Wavefront::kRing]));
            const int32_t *c = w->h_ring + (size_t)(it % Wavefront::kRing) * kNumCounters;
            const int next_live = c[(it & 1) ^ 1];
            uint64_t shadow = 0;
            for (int l = 0; l < nl; ++l)
                shadow += (uint64_t)c[kShadowCount + l];
            s->stats.rays_shadow += shadow;
            s->stats.rays_closest += (uint64_t)next_live; // traced by iteration it + 1
            if (next_live == 0)
                dead = true;
            return TRT_OK;
        };
        int depth = 0;
        for (; !dead; ++depth)
        {
            const unsigned grid_plain = (unsigned)(s->sm_count * 16);
            if (profile && (rc = stamp()))
                return rc;
            if (mode == 1)
                k_trace<1><<<grid_plain, kBlock, 0, stream>>>(s->view, b, q);
            else if (mode == 2)
                k_trace<2><<<grid_plain, kBlock, 0, stream>>>(s->view, b, q);
            else
                k_trace<0><<<grid_trace, kBlock, 0, stream>>>(s->view, b, q);
            if (profile && (rc = stamp()))
                return rc;
            k_reset_counters<<<1, 32, 0, stream>>>(b, q ^ 1);
            k_shade<<<grid_kshade, kShadeBlock, 0, stream>>>(s->view, b, q, depth, p.max_depth, s0, p.seed);
            TRT_CUDA(cudaMemcpyAsync(w->h_ring + (size_t)(depth % Wavefront::kRing) * kNumCounters, b.counters,
                                     kNumCounters * 4, cudaMemcpyDeviceToHost, stream));
            TRT_CUDA(cudaEventRecord(w->ring_ev[depth % Wavefront::kRing], stream));
            if (profile && (rc = stamp()))
                return rc;
            if (mode == 1)
                k_shadow<1><<<grid_plain, kBlock, 0, stream>>>(s->view, b);
            else if (mode == 2)
                k_shadow<2><<<grid_plain, kBlock, 0, stream>>>(s->view, b);
            else
                k_shadow<0><<<grid_shadow, kBlock, 0, stream>>>(s->view, b);
            if (profile && (rc = stamp()))
                return rc;
            k_accumulate<<<grid_shade, kBlock, 0, stream>>>(s->view, b, q);
            if (profile && (rc = stamp()))
                return rc;
            s->stats.kernel_launches += 5;
            q ^= 1;
            if (depth >= kLag && (rc = consume(consumed++)))
                return rc;
        }
        // iterations launched after the batch died are no-ops on empty queues; drain their snapshots
        for (; consumed < depth; ++consumed)
            if ((rc = consume(consumed)))
                return rc;
        if (profile)
        {
            // five timestamps per iteration: trace | reset + shade + counter snapshot | shadow | accumulate
            TRT_CUDA(cudaStreamSynchronize(stream));
            for (size_t i = 0; i + 4 < prof_used; i += 5)
                for (int k = 0; k < 4; ++k)
                {
                    float ms = 0;
                    TRT_CUDA(cudaEventElapsedTime(&ms, w->prof_ev[i + k], w->prof_ev[i + k + 1]));
                    prof_ms[k] += ms;
                }
            prof_used = 0;
        }
        k_deposit<<<(unsigned)((npix + 255) / 256), 256, 0, stream>>>(b, d_accum, (int)npix, ns);
        s->stats.kernel_launches++;
        TRT_CUDA(cudaGetLastError());
    }
    TRT_CUDA(cudaEventRecord(w->ev1, stream));
    TRT_CUDA(cudaStreamSynchronize(stream));
    float ms = 0;
    TRT_CUDA(cudaEventElapsedTime(&ms, w->ev0, w->ev1));
    s->stats.last_render_ms = ms;
    s->stats.ms_trace = prof_ms[0], s->stats.ms_shade = prof_ms[1], s->stats.ms_shadow = prof_ms[2];
    s->stats.ms_accumulate = prof_ms[3];
    return TRT_OK;
}

int resolveImage(trt_scene *s, const double *d_accum, int spp, double *d_image, uint8_t *d_rgb8, cudaStream_t stream)
{
    const size_t n = (size_t)s->width * s->height * 3;
    k_resolve<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_accum, n, spp, d_image, d_rgb8);
    TRT_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return TRT_OK;
}
} // namespace trt
