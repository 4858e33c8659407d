// Wavefront path tracer: the loop body main.cpp:88-108 + shade / nextRay / Sample / RR (pathTracing.cpp:3-209)
// as a sequence of kernels over a batch of paths held in HBM (SoA), one iteration per path depth:
//
//   k_raygen      (K1)  main.cpp:88-95 + Camera::getRay           -> ray[slot], queue = all slots
//   k_walk        (K2)  ONE persistent walker launch per depth over two kinds of rays: traverseBVH for the queue's
//                       paths -> hit[slot], and pathTracing.cpp:51-58 for the light-sample ("shadow") rays the
//                       previous depth's k_shade emitted: closest hit, the sample counts only if the CLOSEST hit's
//                       material is the light's material (not an any-hit test)
//   k_shade       (K3)  first settles the path's previous vertex — L[slot] += throughput * sum(visible light samples,
//                       XML order); throughput *= weight — then shade(): emissive return, Kd/texture, per-light NEE
//                       sample (emits shadow rays), RR, nextRay/Sample -> next ray, weight, next queue
//   k_finish            the tail of a batch: once few paths are alive, ONE launch runs each of them to its end (the
//                       same shade / walk device functions, one thread per path) instead of two launches per depth
//   k_deposit           after the batch: settles the last vertex of every path the same way, then
//                       accum[pixel] += sum over the batch's samples of L (double, fixed order)
//   k_resolve     (K4)  accum / spp (main.cpp:101) and the gamma-2.2 8-bit pack of imshow (main.cpp:30-38)
//
// Two launches per depth (round 1 had five: trace, counter reset, shade, shadow, accumulate).  A 512x512 16-spp job is
// 42 depths of ever fewer paths, each launch at least one wave and a launch gap whatever the queue length, so the
// launches per depth set its time.
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
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace trt
{
namespace
{
constexpr int kBlock = 128;
#ifndef TRT_TAIL_PATHS_DEFAULT
#define TRT_TAIL_PATHS_DEFAULT 65536
#endif
// k_shade is a long dependent chain per warp (ncu, profiles/r02_shade_staircase.txt: issue-active 44 %, 0.75 eligible warps
// per cycle): its throughput is resident warps / that latency, so registers are traded for warps.  64-thread CTAs with a
// minimum of 14 resident per SM cap it at 72 registers (88 bytes of spills) = 7 warps per scheduler.  k_shade's own
// device time on back / veach-mis / staircase (TRT_RENDER_PROFILE; 512x512x16, 1280x720x32, 1280x720x16):
//   min. CTAs  9 (86 registers, 5 warps / scheduler, round 1's setting)   2.48 / 18.39 / 27.26 ms
//             12 (80, no spills, 6)                                        2.43 / 16.54 / 25.30
//             14 (72, 88 B of spills, 7)                                   2.42 / 15.75 / 24.62   <- shipped
//             16 (64, 194 B, 8)                                            2.52 / 15.38 / 24.45
//             18 (56, 200 B, 9)                                            2.60 / 15.10 / 24.21
//             21 (40, 530 B, 12)                                           3.13 / 15.80 / 25.67
#ifndef TRT_SHADE_MINBLOCKS
#define TRT_SHADE_MINBLOCKS 14
#endif
constexpr int kShadeBlock = 64, kShadeMinBlocks = TRT_SHADE_MINBLOCKS;
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
    float4 *sh_o, *sh_d; // sh_o.w = contribution index (slot*n_lights + light), sh_d.w = search bound (just beyond the light point)
    float4 *sh_contrib;  // [slot*n_lights + light]
    // [0],[1]: path queue sizes (ping-pong); [kPool]: ray-pool cursor of the persistent walker; [kDone]: CTAs of the
    // running k_walk that have finished; [kShadowCount + l]: shadow rays queued for light l (segment l of sh_o / sh_d
    // starts at l * capacity)
    int32_t *counters;
    int32_t capacity; // paths per batch = length of one per-light shadow segment
};
constexpr int kPool = 2, kDone = 3, kShadowCount = 8, kNumCounters = 8 + 32;
// walker tokens: bits 31 and 29 are the walker's own (class-1 mark, "second, unbounded attempt"), bit 30 tells a
// shadow-queue entry from a path slot
constexpr unsigned kShadowBit = 0x40000000u;
constexpr unsigned kTokenLimit = 0x20000000u; // slots and shadow-queue entries stay below 2^29
#ifndef TRT_SHADOW_BOUND
#define TRT_SHADOW_BOUND 1
#endif

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

// ((hi << 21) | (lo >> 11)) * 2^-53, the 53-bit uniform of DESIGN.md §5, without the 64-bit integer-to-double conversion
// (a multi-instruction sequence on the GPU): the upper 21 and the lower 32 bits of that integer convert exactly one by
// one, their weighted sum is below 2^53 and therefore exact too, and so is the scaling — the same double, bit for bit.
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo)
{
    const uint32_t top = hi >> 11, low = (hi << 21) | (lo >> 11);
    return fma((double)top, 4294967296.0, (double)low) * (1.0 / 9007199254740992.0);
}

struct Rng
{
    uint32_t k0, k1, pixel, sample, depth;
    // both uniforms of one block: u[0] from words (0,1), u[1] from words (2,3)
    __device__ __forceinline__ void block(uint32_t blk, double u[2]) const
    {
        uint32_t o[4];
        philox4x32_10(pixel, sample, depth, blk, k0, k1, o);
        u[0] = u53(o[0], o[1]);
        u[1] = u53(o[2], o[3]);
    }
};

__device__ __forceinline__ float3 xyz(float4 v) { return f3(v.x, v.y, v.z); }

// Programmatic dependent launch (sm_90+): the two kernels of a depth are launched with the programmatic-stream-
// serialization attribute, so the next kernel's CTAs are set up and scheduled while the previous one drains, and wait
// HERE, before they touch anything the previous kernel wrote, until it has completed and its writes are visible.  The
// first statement of every kernel of the wavefront: a kernel launched the ordinary way falls through both.
__device__ __forceinline__ void pdlEntry()
{
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// x / c for a finite c > 0 without the division's slow path.  FCHK sends a zero, denormal or tiny dividend to a
// ~100-instruction subroutine that runs with one or two lanes of the warp (ncu, profiles/r02_shade_staircase.txt: 9 % of
// k_shade's warp instructions at 1.7 lanes, fed by Ks = 0, by colours with a zero channel and by the specular term
// Ks (Ns+2) cos^Ns of a narrow lobe, which is denormal for ~1.5 % of the evaluations at Ns = 1000).
//   x == +-0:      +-0 / c = +-0 = x, exactly;
//   |x| < 1e-30:   x * (1/c) — within one ulp of the IEEE quotient of a quantity that is 1e-30 of any radiance the frame
//                  can show (the one place where this kernel's radiance arithmetic is not the reference's operation,
//                  DESIGN.md §6);
//   otherwise the IEEE division, bit for bit.
__device__ __forceinline__ float divByPositive(float x, float c)
{
    const bool tiny = fabsf(x) < 1.0e-30f;
    float dividend = tiny ? 1.0f : x;
    // The select must stay in front of the division: with a constant c the compiler folds 1 / c, divides x itself on
    // every lane and picks afterwards — and FCHK then still sends the zero and tiny x to the slow path (ncu on the
    // first round-2 build: 45 000 of 55 000 specular evaluations per launch, three calls each, at 2 lanes).
    asm volatile("" : "+f"(dividend));
    const float q = dividend / c;
    return tiny ? x * q : q; // (tiny: q = 1/c; x = +-0 gives +-0)
}

// cos^Ns of the specular term.  A whole exponent from 1 to 4096 (every material of the cg22 scenes: 1 .. 1000) is
// evaluated by squaring and multiplying in double — at most 12 + 12 products, relative error below Ns 2^-53 = 1e-13 —
// instead of the library's pow, a ~180-instruction subroutine that ncu shows at 2 lanes of the warp and at 11 % of
// k_shade's warp instructions on staircase.  The value is rounded to float right after: the two agree except where the
// exact power lies within 1e-13 of a float rounding boundary (the oracle's glibc pow and CUDA's pow differ the same
// way).  Any other exponent: pow.
__device__ __forceinline__ double powLobe(double x, float Ns)
{
    const int n = (int)Ns;
    if ((float)n == Ns && n >= 1 && n <= 4096)
    {
        double r = (n & 1) ? x : 1.0, b = x;
        for (int k = n >> 1; k; k >>= 1)
        {
            b *= b;
            if (k & 1)
                r *= b;
        }
        return r;
    }
    return pow(x, (double)Ns);
}
__device__ __forceinline__ float3 divByPositive(float3 v, float c)
{
    return f3(divByPositive(v.x, c), divByPositive(v.y, c), divByPositive(v.z, c));
}
__device__ __forceinline__ float4 xyzw(float3 v, float w) { return make_float4(v.x, v.y, v.z, w); }

// ------------------------------------------------------------------------------------------------ K1
__global__ void __launch_bounds__(kBlock) k_raygen(SceneView sv, WfBuffers wf, int n_paths, int sample0, uint64_t seed)
{
    pdlEntry();
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x < kNumCounters) // queue 0 = every slot; empty next queue, shadow segments, cursors
        wf.counters[threadIdx.x] = (threadIdx.x == 0) ? n_paths : 0;
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
    wf.nee_mask[slot] = 0u; // no vertex to settle yet
    wf.queue[0][slot] = slot;
}

// ------------------------------------------------------------------------------------------------ K2
// The walker's ray pool of one depth: first the closest-hit rays of the live paths (token = path slot), then the
// light-sample rays, one queue segment per light (neighbouring lanes: same light, nearby pixels; token = kShadowBit |
// entry index in sh_o / sh_d = light * capacity + position; both below 2^30 by the batch-size check of renderAccumulate).
// STOP: compile the early stop of occluded light samples in (canStop).  It costs the walker registers (44 instead of 24
// bytes of spills at the 56-register cap) and a box test per improved hit, and pays only where light samples are
// often occluded: staircase shadow walk 58.3 -> 54.4 ms, but veach-mis (four plates under open lights) 28.2 -> 29.0 ms
// and the Cornell shell 1.66 -> 1.71 ms.  Two instantiations of k_walk, chosen per scene at trt_scene_create
// (trt_scene::shadow_stop: scenes of 8192 triangles and more, TRT_SHADOW_STOP=0|1 overrides).
template <bool STOP>
struct WalkRays
{
    // A light-sample ray starts its search with the distance just beyond its light point as the bound (sh_d.w): the
    // closest hit of the whole ray is what pathTracing.cpp:51-58 asks for, and whenever anything is hit up to the light
    // point — normally the light's own triangle — that IS the closest hit, found without walking or stacking what lies
    // behind the light.  Only if nothing at all is hit within the bound (a sample on a triangle edge that the inside
    // test rejects) does the walker start the ray again without one.
    static constexpr bool kHasBound = TRT_SHADOW_BOUND != 0;
    WfBuffers wf;
    const TriShade *tri_shade;
    const DeviceLight *lights;
    int n_lights, qsel;
    unsigned int n_closest;
    __device__ __forceinline__ unsigned int locate(unsigned int i) const
    {
        if (i < n_closest)
            return (unsigned int)wf.queue[qsel][i];
        i -= n_closest;
        int l = 0;
        for (; l < n_lights - 1; ++l)
        {
            const unsigned int c = (unsigned int)wf.counters[kShadowCount + l];
            if (i < c)
                break;
            i -= c;
        }
        return kShadowBit | ((unsigned int)l * (unsigned int)wf.capacity + i);
    }
    __device__ __forceinline__ void load(unsigned int tok, float3 &S, float3 &d, float &tmax) const
    {
        const bool sh = (tok & kShadowBit) != 0;
        const unsigned int e = tok & ~kShadowBit;
        const float4 o = sh ? wf.sh_o[e] : wf.ray_o[e], dd = sh ? wf.sh_d[e] : wf.ray_d[e];
        S = xyz(o), d = xyz(dd);
        tmax = (kHasBound && sh) ? dd.w : TRT_INF;
    }
    __device__ __forceinline__ bool bounded(unsigned int tok) const { return kHasBound && (tok & kShadowBit) != 0; }
    // After a leaf scan that changed the best hit of a light-sample ray: may the walk stop here?  The sample counts only
    // if the CLOSEST hit carries the light's material (pathTracing.cpp:54-58).  Every triangle of that material lies
    // inside sv.light_box[light] (the union of the fast layout's padded boxes of those triangles), so none of them can be
    // hit before the ray enters that box: once a triangle of ANOTHER material has been hit in front of the box — or the
    // ray misses the box, or the box lies behind it — the closest hit cannot be the light's whatever else the ray would
    // still find, and the sample is known to be lost.  (hit.id is the fast-layout index here; inv / nsi as in childCull.)
    __device__ __forceinline__ bool canStop(const SceneView &sv, unsigned int tok, const Hit &h, float3 inv, float3 nsi) const
    {
        if (!STOP || !(tok & kShadowBit) || h.id < 0 || !sv.light_box)
            return false;
        const unsigned int li = (tok & ~kShadowBit) / (unsigned int)wf.capacity;
        if (__ldg(sv.fast_mtl + h.id) == lights[li].material)
            return false; // the light itself so far: something may still lie in front of it
        const float4 lo = __ldg(sv.light_box + 2 * li), hi = __ldg(sv.light_box + 2 * li + 1);
        const float inx = __fmaf_rn(hi.x, inv.x, nsi.x), iny = __fmaf_rn(hi.y, inv.y, nsi.y), inz = __fmaf_rn(hi.z, inv.z, nsi.z);
        const float outx = __fmaf_rn(lo.x, inv.x, nsi.x), outy = __fmaf_rn(lo.y, inv.y, nsi.y), outz = __fmaf_rn(lo.z, inv.z, nsi.z);
        const float t1 = fminf(fmaxf(inx, outx), fminf(fmaxf(iny, outy), fmaxf(inz, outz)));
        const float t0 = fmaxf(fminf(inx, outx), fmaxf(fminf(iny, outy), fminf(inz, outz)));
        return (t0 > t1) || (t1 < 0.0f) || (h.t < t0);
    }
    // pathTracing.cpp:54-58: visible iff the closest hit's material is the light's material (entry e lies in the queue
    // segment of its light)
    __device__ __forceinline__ void finishShadow(unsigned int e, bool hit, int mtl) const
    {
        const bool visible = hit && mtl == lights[e / (unsigned int)wf.capacity].material;
        if (!visible)
            wf.sh_contrib[__float_as_int(wf.sh_o[e].w)] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ void store(unsigned int tok, const Hit &h) const
    {
        if (tok & kShadowBit)
            finishShadow(tok & ~kShadowBit, h.id >= 0, h.id >= 0 ? tri_shade[h.id].mtl : -1);
        else
            wf.hit_id[tok] = h.id, wf.hit_t[tok] = h.t;
    }
    __device__ __forceinline__ void storeFast(const SceneView &sv, unsigned int tok, const Hit &h) const
    {
        if (tok & kShadowBit) // a light sample needs only the hit's material: one load from a dense 4-byte table
            finishShadow(tok & ~kShadowBit, h.id >= 0, h.id >= 0 ? __ldg(sv.fast_mtl + h.id) : -1);
        else
            wf.hit_id[tok] = (h.id >= 0) ? __ldg(sv.fast_orig + h.id) : -1, wf.hit_t[tok] = h.t; // fast -> post-build index
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
        float tmax; // these walks ignore the bound: the unbounded search gives the same closest hit
        rays.load(token, S, d, tmax);
        Hit hit;
        if (MODE == 1)
            traceRefTopology<false>(sv, S, d, hit);
        else
            traceClosest(sv, S, d, hit);
        rays.store(token, hit);
    }
}

// what: bit 0 = the queue's closest-hit rays, bit 1 = the shadow rays, bit 2 = this is the depth's last walk: the CTA
// that finishes last zeroes the counters k_shade is about to fill (next queue size, shadow segments).  The pool cursor
// is zeroed by every launch's last CTA.  (TRT_RENDER_PROFILE walks the two kinds in two launches to time them apart.)
// 9 resident CTAs per SM = 56 registers instead of 64 / 8 CTAs: the walker is latency-bound (ncu: 1.4 eligible warps
// per cycle), one more CTA of warps pays; 10 CTAs = 48 registers spill 134 bytes and lose.
#ifndef TRT_WALK_CTAS
#define TRT_WALK_CTAS 9
#endif
// snap (what & 4 only, may be null): a slot of page-locked HOST memory mapped into the device.  The last CTA copies the
// counters there before it resets them — the queue length and shadow-ray counts the previous depth's k_shade left —
// then the sequence number `seq`: the host learns how the batch is decaying by polling that word, with no copy and no
// event between the kernels of the stream (round 1's per-depth cudaMemcpyAsync + event put a copy-engine round trip
// on the critical path of every depth).
template <int MODE, bool STOP>
__global__ void __launch_bounds__(kBlock, TRT_WALK_CTAS) k_walk(SceneView sv, WfBuffers wf, int qsel, int what, int32_t *snap,
                                                                int32_t seq)
{
    pdlEntry();
    unsigned int n_sh = 0;
    if (what & 2)
        for (int l = 0; l < sv.n_lights; ++l)
            n_sh += (unsigned int)wf.counters[kShadowCount + l];
    WalkRays<STOP> r{wf, sv.tri_shade, sv.lights, sv.n_lights, qsel, (what & 1) ? (unsigned int)wf.counters[qsel] : 0u};
    const unsigned int n = r.n_closest + n_sh;
    if (MODE != 0)
        traceGridStride<MODE>(sv, r, n);
    else
        walkPersistent<false>(sv, r, n, reinterpret_cast<unsigned int *>(wf.counters + kPool));
    // every CTA has read the counters before it got here; the last one to arrive resets them for what follows
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0)
    {
        __threadfence();
        s_last = atomicAdd(reinterpret_cast<unsigned int *>(wf.counters + kDone), 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last)
    {
        if ((what & 4) && snap)
        {
            if (threadIdx.x < kNumCounters)
            {
                snap[threadIdx.x] = wf.counters[threadIdx.x];
                __threadfence_system();
            }
            __syncthreads(); // (s_last is uniform over the CTA)
            if (threadIdx.x == 0)
                *reinterpret_cast<volatile int32_t *>(snap + kNumCounters) = seq;
        }
        if (threadIdx.x == 0)
            wf.counters[kPool] = 0, wf.counters[kDone] = 0;
        if (what & 4)
        {
            if (threadIdx.x == 0)
                wf.counters[qsel ^ 1] = 0;
            if (threadIdx.x < 32)
                wf.counters[kShadowCount + threadIdx.x] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------------ K3
// pathTracing.cpp:111-145.  The reference draws theta = asin(sqrt(u)) (diffuse) or acos(u^(1/(Ns+1))) (specular) and
// then takes sin(theta) and cos(theta); here sin(asin(x)) is x, cos(acos(y)) is y and the other one is the square root
// of one minus the square — the same values to within the 1-2 ulp by which any two double libms differ (the oracle's
// glibc and CUDA's already do), without two of the kernel's four double-precision transcendental calls per bounce.
__device__ __forceinline__ float3 sampleLobe(float3 direction, int ray_type, double Ns, double u_phi, double u_theta)
{
    const double phi = u_phi * 2 * (double)kPI;
    double st, ct;
    if (ray_type == DIFFUSE)
    {
        st = sqrt(u_theta);       // sin(asin(sqrt(u)))
        ct = sqrt(1.0 - u_theta); // cos(asin(sqrt(u)))
    }
    else
    {
        ct = pow(u_theta, (double)1 / (Ns + 1)); // cos(acos(.))
        st = sqrt(fmax(1.0 - ct * ct, 0.0));     // sin(acos(.))
    }
    double sp, cp;
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

// Settles the vertex a path was shaded at last: the light samples whose shadow ray found the light are added in XML
// order (L_dir += ..., pathTracing.cpp:70) and weighted by the throughput up to that vertex.  Round 1 did this in a
// kernel of its own after the shadow kernel; now the path's NEXT k_shade (or k_deposit, for a path that ended) does it,
// with the same operations in the same order per path.
__device__ __forceinline__ void settleVertex(const WfBuffers &wf, int n_lights, int slot, uint32_t mask, float4 T, float4 &L)
{
    float3 Ldir = f3(0.f, 0.f, 0.f);
    while (mask)
    {
        const int li = __ffs(mask) - 1;
        mask &= mask - 1;
        Ldir = Ldir + xyz(wf.sh_contrib[slot * n_lights + li]);
    }
    L.x += T.x * Ldir.x, L.y += T.y * Ldir.y, L.z += T.z * Ldir.z;
}

// K3.  (Round 2 also built a three-phase form of this kernel — vertex records in shared memory, the light picks of
// a CTA compacted into a task list, the picked (vertex, light) pairs dealt out to full warps — to lift the light loop
// from 18 to ~30 lanes per instruction.  Bit-identical frames, and 5 % SLOWER (k_shade 24.1 -> 25.3 ms staircase, 15.5
// -> 16.2 veach-mis): the kernel's time is load and dependency latency — ncu: 27 % of stall samples on loads, among them
// the round trip of the shadow-queue atomic of every light iteration, 26 % on fixed-latency dependencies, 14 % on
// instruction fetch — not the lanes its instructions serve.  DESIGN.md §10.)
// One path vertex: settles the previous vertex, then shade() / nextRay() / RR() of pathTracing.cpp for the hit the walk
// left in hit_id / hit_t[slot].  Returns whether the path goes on (its next ray is then in ray_o / ray_d[slot]).
// TAIL = false: light-sample rays go to the per-light queues for the next k_walk; TAIL = true (k_finish): they are walked
// here, one after the other, and their outcome is written straight into sh_contrib.
template <bool TAIL>
__device__ __forceinline__ bool shadeVertex(const SceneView &sv, const WfBuffers &wf, int slot, int depth, int max_depth, int sample0,
                                            uint64_t seed, int npix, int pixel0, unsigned int &tail_shadow)
{
    bool survives = false;
    uint32_t mask = 0;
    float4 weight = make_float4(0.f, 0.f, 0.f, 0.f);
    const int tri = wf.hit_id[slot];
    const float4 rd4 = wf.ray_d[slot];
    const int via = __float_as_int(rd4.w);
    if (depth > 0)
    {
        // the previous vertex: its light samples (their shadow rays were walked together with this path's ray),
        // then throughput *= weight of the bounce that led here (:84-98)
        float4 T = wf.thr[slot];
        const uint32_t pm = wf.nee_mask[slot];
        if (pm)
        {
            float4 L = wf.L[slot];
            settleVertex(wf, sv.n_lights, slot, pm, T, L);
            wf.L[slot] = L;
        }
        const float4 w = wf.weight[slot];
        T.x *= w.x, T.y *= w.y, T.z *= w.z;
        wf.thr[slot] = T;
    }
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
            Rng rng{(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)(pixel0 + slot % npix), (uint32_t)(sample0 + slot / npix),
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
                // frac() of a tiny negative coordinate is exactly 1.0: the reference then reads past the image
                // (:24, undefined); the oracle and this kernel clamp to the last texel
                const int r = min((int)(irow * tx.rows), tx.rows - 1), c = min((int)(icol * tx.cols), tx.cols - 1);
                const uint8_t *px = tx.bgr + ((size_t)r * tx.cols + c) * 3;
                Kd = f3((float)((double)px[2] / 255), (float)((double)px[1] / 255), (float)((double)px[0] / 255));
            }

            // the same for every light: the diffuse BRDF term Kd / PI (:67) and |pn| (:63)
            const float3 Kd_pi = divByPositive(Kd, kPI);
            const float pn_len = length3(pn);

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
                const float3 radiance = sv.materials[lt.material].radiance;
                const float pdf_light = lt.pdf; // (float)(double(1) / area) of :62, computed once per light on the host
                const float cos_theta_p = fabsf(dot3(wo, light_n));
                const float cos_theta = fabsf(wo_pn / pn_len);
                const float3 diff = light_p - P;
                const float len2 = dot3(diff, diff);
                const float3 intensity = (((radiance * cos_theta_p) * cos_theta) / len2) / pdf_light;
                const float3 h = normalize3((wi + wo) * 0.5f);
                const double cos_alpha = fmax((double)dot3(pn, h), 0.0);
                // Ks == 0 (every diffuse material): Ks * (Ns+2) * pow(...) is +0 for any finite power, and the
                // power is finite for cos_alpha in [0,1] and Ns >= 0 — the double-precision pow is skipped
                const bool no_spec = (m_Ks.x == 0.f) && (m_Ks.y == 0.f) && (m_Ks.z == 0.f) && (m_Ns >= 0.f);
                // the specular term of a diffuse material is (+0 * (Ns + 2)) * 0 / 2pi = +0 in every channel
                float3 spec_term = f3(0.f, 0.f, 0.f);
                if (!no_spec)
                {
                    // cos^Ns below 2^-110 (a narrow lobe seen from outside it: most evaluations at Ns = 250..1000) is
                    // taken as 0 without running the double-precision pow — a 300-instruction subroutine that ncu
                    // shows at 4 lanes; what is dropped is below 1e-30 of the diffuse term it would be added to.
                    // (cos_alpha = 0: log2 = -inf; Ns = 0: the product is NaN or 0 and the pow runs.)
                    float pw = 0.f;
                    if (!(m_Ns * __log2f((float)cos_alpha) < -110.f))
                        pw = (float)powLobe(cos_alpha, m_Ns);
                    spec_term = divByPositive((m_Ks * (m_Ns + 2.0f)) * pw, 2.0f * kPI);
                }
                const float3 brdf = Kd_pi + spec_term;
                const float3 contrib = intensity * brdf;
                const int cidx = slot * sv.n_lights + li;
                if (TAIL)
                {
                    // the light sample's ray walked on the spot (pathTracing.cpp:51-58: the CLOSEST hit decides; the
                    // bounded search + retry of the queued form returns the same hit, see WalkRays)
                    Hit sh;
                    traceClosest(sv, P, wo, sh);
                    ++tail_shadow;
                    const bool visible = sh.id >= 0 && sv.tri_shade[sh.id].mtl == lt.material;
                    wf.sh_contrib[cidx] = visible ? xyzw(contrib, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                else
                {
                    wf.sh_contrib[cidx] = xyzw(contrib, 0.f);
                    // one queue segment per light, appended by the lanes that are converged here with a single
                    // atomic: neighbouring entries are neighbouring pixels aiming at the same light
                    cg::coalesced_group g = cg::coalesced_threads();
                    int at = 0;
                    if (g.thread_rank() == 0)
                        at = atomicAdd(&wf.counters[kShadowCount + li], (int)g.size());
                    at = g.shfl(at, 0);
                    const size_t e = (size_t)li * wf.capacity + at + g.thread_rank();
                    wf.sh_o[e] = xyzw(P, __int_as_float(cidx));
                    // the light point lies at |diff| along wo (a unit vector up to rounding): search up to 1.001 x that
                    wf.sh_d[e] = xyzw(wo, sqrtf(len2) * 1.001f);
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
                    const double kd = mp->kd, ks = mp->ks; // |Kd| / (|Kd| + |Ks|), |Ks| / (...), :191-192 (host, once)
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
    return survives;
}

// K3: one thread per entry of the path queue.
__global__ void __launch_bounds__(kShadeBlock, kShadeMinBlocks) k_shade(SceneView sv, WfBuffers wf, int qsel, int depth, int max_depth,
                                                  int sample0, uint64_t seed, int npix, int pixel0)
{
    pdlEntry();
    const int count = wf.counters[qsel];
    unsigned int unused = 0;
    // warp-uniform grid-stride loop: the queue append below is warp-collective
    for (int base = blockIdx.x * blockDim.x; base < count; base += gridDim.x * blockDim.x)
    {
        const int i = base + threadIdx.x;
        bool survives = false;
        int slot = 0;
        if (i < count)
        {
            slot = wf.queue[qsel][i];
            survives = shadeVertex<false>(sv, wf, slot, depth, max_depth, sample0, seed, npix, pixel0, unused);
        }
        appendQueue(&wf.counters[qsel ^ 1], wf.queue[qsel ^ 1], survives, slot);
    }
}

// The tail of a batch in ONE launch: once few paths are left, every depth costs two launches of almost no work — a
// 512x512 16-spp job is 42 depths, 30 of them with fewer paths than the GPU has threads — so from there on each remaining
// path is run to its end by one thread: shadeVertex<true> (light samples walked on the spot), then the walk of the next
// ray, until the path dies.  The same device functions on the same per-path state in the same order as the per-depth
// kernels: the frame is bit-identical whatever the switch point (TRT_TAIL_PATHS, tests/test_gpu_render.py).
// Lanes of a warp stay in step (all shade, then all walk); what is lost is lanes of paths that have ended.
// The last CTA publishes the snapshot that tells the host the batch is over, with the rays traced here in two spare
// counters (trt_stats).
constexpr int kTailClosest = 4, kTailShadow = 5;
#ifndef TRT_FINISH_MINBLOCKS
#define TRT_FINISH_MINBLOCKS 4 // 158 registers, no spills (8: 113 registers, within 0.5 %)
#endif
__global__ void __launch_bounds__(kShadeBlock, TRT_FINISH_MINBLOCKS) k_finish(SceneView sv, WfBuffers wf, int qsel, int depth0, int max_depth, int sample0,
                                                        uint64_t seed, int npix, int pixel0, int32_t *snap, int32_t seq)
{
    pdlEntry();
    const int count = wf.counters[qsel];
    unsigned int n_closest = 0, n_shadow = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        const int slot = wf.queue[qsel][i];
        for (int depth = depth0; shadeVertex<true>(sv, wf, slot, depth, max_depth, sample0, seed, npix, pixel0, n_shadow); ++depth)
        {
            Hit h;
            traceClosest(sv, xyz(wf.ray_o[slot]), xyz(wf.ray_d[slot]), h);
            wf.hit_id[slot] = h.id, wf.hit_t[slot] = h.t;
            ++n_closest;
        }
    }
    __syncwarp();
    n_closest = __reduce_add_sync(0xffffffffu, n_closest), n_shadow = __reduce_add_sync(0xffffffffu, n_shadow);
    if ((threadIdx.x & 31) == 0)
    {
        if (n_closest)
            atomicAdd(&wf.counters[kTailClosest], (int)n_closest);
        if (n_shadow)
            atomicAdd(&wf.counters[kTailShadow], (int)n_shadow);
    }
    __syncthreads();
    __shared__ bool s_last;
    if (threadIdx.x == 0)
    {
        __threadfence();
        s_last = atomicAdd(reinterpret_cast<unsigned int *>(wf.counters + kDone), 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last && snap)
    {
        // the batch is over: both queues read as empty, no light samples pending
        if (threadIdx.x < kNumCounters)
        {
            const bool kept = threadIdx.x == kTailClosest || threadIdx.x == kTailShadow;
            snap[threadIdx.x] = kept ? __ldcg(wf.counters + threadIdx.x) : 0;
            __threadfence_system();
        }
        __syncthreads();
        if (threadIdx.x == 0)
            *reinterpret_cast<volatile int32_t *>(snap + kNumCounters) = seq;
    }
    if (s_last && threadIdx.x == 0) // leave the counters as the last walk of a batch leaves them
        wf.counters[kDone] = 0, wf.counters[qsel] = 0, wf.counters[kTailClosest] = 0, wf.counters[kTailShadow] = 0;
}

// accum[pixel] += L of the batch's samples, ONE AT A TIME in sample order: the running double sum then goes through the
// same sequence of additions however the samples were split into batches (round 1 summed a batch first and added the
// partial sum, which changed last bits with the batch size on large frames), and a render resumed from a checkpoint
// continues the very sequence an uninterrupted one follows.
__global__ void __launch_bounds__(256) k_deposit(WfBuffers wf, double *accum, int npix, int samples_in_batch, int n_lights)
{
    pdlEntry();
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= npix)
        return;
    double r = accum[(size_t)pix * 3 + 0], g = accum[(size_t)pix * 3 + 1], b = accum[(size_t)pix * 3 + 2];
    for (int s = 0; s < samples_in_batch; ++s)
    {
        const int slot = s * npix + pix;
        float4 L = wf.L[slot];
        const uint32_t pm = wf.nee_mask[slot]; // a path that ended after a vertex with light samples: settle it
        if (pm)
            settleVertex(wf, n_lights, slot, pm, wf.thr[slot], L);
        r += (double)L.x, g += (double)L.y, b += (double)L.z;
    }
    accum[(size_t)pix * 3 + 0] = r;
    accum[(size_t)pix * 3 + 1] = g;
    accum[(size_t)pix * 3 + 2] = b;
}

// ---- trt_shade: a wavefront that starts from caller-supplied hits instead of camera rays (pathtracing.h:14) ----
__global__ void __launch_bounds__(kBlock) k_inject(WfBuffers wf, const float *__restrict__ rays6, const int32_t *__restrict__ id,
                                                   const float *__restrict__ t, int n)
{
    pdlEntry();
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x < kNumCounters)
        wf.counters[threadIdx.x] = (threadIdx.x == 0) ? n : 0;
    if (slot >= n)
        return;
    const float *r = rays6 + (size_t)slot * 6;
    wf.ray_o[slot] = make_float4(r[0], r[1], r[2], 0.f);
    // arrives like a camera ray: an emissive hit returns its radiance (pathTracing.cpp:9-12, main.cpp:101)
    wf.ray_d[slot] = make_float4(r[3], r[4], r[5], __int_as_float(CAMERA));
    wf.hit_id[slot] = id[slot];
    wf.hit_t[slot] = t[slot];
    wf.thr[slot] = make_float4(1.f, 1.f, 1.f, 0.f);
    wf.L[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    wf.nee_mask[slot] = 0u;
    wf.queue[0][slot] = slot;
}

__global__ void __launch_bounds__(kBlock) k_collect(WfBuffers wf, float *__restrict__ radiance3, int n, int n_lights)
{
    pdlEntry();
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n)
        return;
    float4 L = wf.L[slot];
    const uint32_t pm = wf.nee_mask[slot];
    if (pm)
        settleVertex(wf, n_lights, slot, pm, wf.thr[slot], L);
    radiance3[(size_t)slot * 3] = L.x, radiance3[(size_t)slot * 3 + 1] = L.y, radiance3[(size_t)slot * 3 + 2] = L.z;
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

// Launch with the programmatic-stream-serialization attribute (see pdlEntry).  TRT_PDL=0 launches the ordinary way.
static bool usePdl()
{
    static const bool on = [] {
        const char *e = getenv("TRT_PDL");
        return !e || atoi(e) != 0;
    }();
    return on;
}
template <typename... KArgs, typename... Args>
static void launchPdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t stream, bool pdl, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(block), cfg.dynamicSmemBytes = 0, cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...); // errors surface at the cudaGetLastError that follows
}

// Path state of the batch in flight.
struct Wavefront
{
    WfBuffers buf{};
    int capacity = 0; // paths
    int n_lights = 0;
    std::vector<void *> allocs;
    static constexpr int kRing = 8, kSlot = kNumCounters + 8; // a slot: the counters, then the sequence word
    int32_t *h_ring = nullptr; // page-locked and mapped: kRing snapshots of the device counters, written by k_walk itself
    int32_t *d_ring = nullptr; // the same memory as the device sees it
    int32_t seq = 0;           // sequence number of the last snapshot asked for
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int blocks_walk = 1, blocks_shade = 1;
    std::vector<cudaEvent_t> prof_ev; // TRT_RENDER_PROFILE: four timestamps per iteration, grown on demand
    ~Wavefront()
    {
        for (void *p : allocs)
            cudaFree(p);
        if (h_ring)
            cudaFreeHost(h_ring);
        for (cudaEvent_t e : {ev0, ev1})
            if (e)
                cudaEventDestroy(e);
        for (cudaEvent_t e : prof_ev)
            cudaEventDestroy(e);
    }
};

void destroyWavefront(trt_scene *s)
{
    delete s->wf;
    s->wf = nullptr;
}

// Path state for `paths` paths.  Built aside and published only when every allocation has succeeded, so that a
// failed attempt (an explicit batch_paths beyond the free memory) leaves the scene without a wavefront and a retry
// with a smaller batch starts clean.
static int ensureWavefront(trt_scene *s, int paths)
{
    if (s->wf && s->wf->capacity >= paths)
        return TRT_OK;
    destroyWavefront(s);
    std::unique_ptr<Wavefront> w(new Wavefront());
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
    {
        cudaGetLastError(); // the failed cudaMalloc must not poison the next call
        return rc;
    }
    TRT_CUDA(cudaMemset(b.counters, 0, kNumCounters * 4));
    TRT_CUDA(cudaHostAlloc((void **)&w->h_ring, Wavefront::kRing * Wavefront::kSlot * 4, cudaHostAllocMapped));
    std::memset(w->h_ring, 0, Wavefront::kRing * Wavefront::kSlot * 4);
    TRT_CUDA(cudaHostGetDevicePointer((void **)&w->d_ring, w->h_ring, 0));
    TRT_CUDA(cudaEventCreate(&w->ev0));
    TRT_CUDA(cudaEventCreate(&w->ev1));
    if (s->shadow_stop)
        TRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w->blocks_walk, k_walk<0, true>, kBlock, 0));
    else
        TRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w->blocks_walk, k_walk<0, false>, kBlock, 0));
    TRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&w->blocks_shade, k_shade, kShadeBlock, 0));
    w->capacity = paths;
    s->wf = w.release();
    return TRT_OK;
}

// Paths in flight per batch.  Path counts decay by 0.8 per depth and every depth costs two launches of at least one
// wave whatever the queue length, so a batch should be large against that tail — but the tail is short now (round 1
// paid five launches per depth and defaulted to 128 Mi paths = 52 GB with six lights for 2.6 % over 32 Mi; with two
// launches per depth, keeping a second batch in flight on a second stream to cover the tails was measured at 0.5 % on
// veach-mis 256 spp and 0.1 % on staircase 1080p 128 spp, and taken out again).  Default: 32 Mi paths (13 GB with six
// lights), bounded by a third of the free memory; batch_paths asks for more.
static long long batchTarget(trt_scene *s, int batch_paths, long long &max_paths)
{
    const int nl1 = std::max(1, s->view.n_lights);
    long long target = batch_paths > 0 ? batch_paths : (32ll << 20);
    // walker tokens carry two flag bits, and slot * n_lights + light indexes the light-sample contributions
    max_paths = (long long)(kTokenLimit - 1) / nl1;
    if (batch_paths <= 0)
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
        {
            const long long per_path = 100 + 48ll * nl1;
            const long long have = (long long)(s->wf ? (size_t)s->wf->capacity * per_path : 0);
            target = std::min(target, std::max(1ll << 20, ((long long)free_b / 3 + have) / per_path));
        }
        target = std::min(target, max_paths);
    }
    return target;
}

// The depth loop over the n_paths paths that k_raygen / k_inject have just set up in queue 0 of a lane (slot = index),
// as a state machine that never blocks: step() launches the next depth when the host may run that far ahead and takes
// in the counter snapshots that have arrived (runToEnd drives it; a host that has other work can interleave it).
// Stream keys of slot: pixel = pixel0 + slot % npix, sample = sample0 + slot / npix.  first_walk = false: the hits of
// depth 0 are already in hit_id / hit_t (trt_shade).  prof_ms (TRT_RENDER_PROFILE) += {closest-hit walk, shadow walk, shade}.
// Switch point of k_finish: live paths at or below which the rest of a batch is run by one launch (0: never).
// Measured on the 512x512 16-spp Cornell job: 4.38 ms without, 4.06 / 3.96 / 3.97 / 4.04 ms at 16 Ki / 64 Ki / 128 Ki /
// 256 Ki paths (profiles/r02_ab_tail.txt); read at every batch so that tests can move it (TRT_TAIL_PATHS).
static long long tailPaths()
{
    const char *e = getenv("TRT_TAIL_PATHS");
    return e ? std::max(0ll, atoll(e)) : (long long)TRT_TAIL_PATHS_DEFAULT;
}

class DepthLoop
{
  public:
    // The host never waits for an iteration it has just launched: every kernel reads its queue length from device
    // memory, and the host looks at the counters of iteration it - kLag to learn when the batch has died out.  (At
    // least one walk is always launched after the k_shade that emptied the queue: it serves the light samples of the
    // last vertices and publishes that depth's counters.)
    static constexpr int kLag = 2;

    void begin(trt_scene *scene, Wavefront *wave, cudaStream_t st, int paths, int npix_, int pixel0_, int sample0_, uint64_t seed_,
               int max_depth_, uint32_t flags, bool first_walk_, double *prof_ms_)
    {
        s = scene, w = wave, stream = st, n_paths = paths, npix = npix_, pixel0 = pixel0_, sample0 = sample0_, seed = seed_;
        max_depth = max_depth_, first_walk = first_walk_, prof_ms = prof_ms_;
        mode = (flags & TRT_RENDER_REFTOPO) ? 1 : ((flags & TRT_RENDER_PLAIN) ? 2 : 0);
        profile = (flags & TRT_RENDER_PROFILE) != 0 && prof_ms;
        pdl = usePdl() && !profile; // (the profile's event records sit between the kernels)
        // (a profile lists every depth's kernels; TRT_RENDER_REFTOPO / _PLAIN ask for one particular walk to the end of
        // every path — k_finish walks the fast layout thread per ray — so that they stay independent checks of it)
        tail_paths = (profile || mode != 0) ? 0 : tailPaths();
        nl = s->view.n_lights;
        seq0 = w->seq;
        q = 0, depth = 0, consumed = 0, dead = false, final_walk = false, prof_used = 0, live_bound = n_paths;
        active = true;
    }
    bool finished() const { return active && final_walk && consumed == depth; }

    // > 0: made progress; 0: nothing to do right now; < 0: a trt_status
    int step()
    {
        int progress = 0, rc;
        while (awaited() && arrived(consumed)) // every snapshot that has come in, oldest first
        {
            consume(consumed++);
            progress = 1;
        }
        if (final_walk)
            ; // k_finish (or the last walk) is on its way: only its snapshot is still to come
        else if (!dead)
        {
            if (depth - consumed <= kLag) // the host runs at most kLag depths ahead of what it has seen
            {
                if ((rc = launchDepth()))
                    return rc;
                progress = 1;
            }
        }
        else
        {
            // one more walk: it serves whatever light samples the last k_shade emitted and publishes its counters
            walk(q, 1 | 2 | 4, 1, depth - 1);
            w->seq = seq0 + depth;
            final_walk = true;
            TRT_CUDA(cudaGetLastError());
            progress = 1;
        }
        return progress;
    }

    // the snapshot never came although the stream has drained (or failed): report instead of spinning for ever
    int checkStalled()
    {
        const cudaError_t e = cudaStreamQuery(stream);
        if (e != cudaErrorNotReady && awaited() && !arrived(consumed))
        {
            setLastError(std::string("wavefront: counter snapshot never arrived: ") +
                         (e == cudaSuccess ? "stream idle" : cudaGetErrorString(e)));
            return TRT_ERR_CUDA;
        }
        return TRT_OK;
    }

    // TRT_RENDER_PROFILE: four timestamps per iteration: closest-hit walk | shadow walk | shade
    int finishProfile()
    {
        if (!profile)
            return TRT_OK;
        TRT_CUDA(cudaStreamSynchronize(stream));
        const bool per_depth = getenv("TRT_RENDER_PROFILE_DEPTHS") != nullptr; // diagnostics: one stderr line per depth
        for (size_t i = 0; i + 3 < prof_used; i += 4)
        {
            float d[3] = {0, 0, 0};
            for (int k = 0; k < 3; ++k)
            {
                TRT_CUDA(cudaEventElapsedTime(&d[k], w->prof_ev[i + k], w->prof_ev[i + k + 1]));
                prof_ms[k] += d[k];
            }
            if (per_depth)
                std::fprintf(stderr, "depth %2zu: closest walk %.4f  shadow walk %.4f  shade %.4f ms\n", i / 4, d[0], d[1], d[2]);
        }
        return TRT_OK;
    }

  private:
    trt_scene *s = nullptr;
    Wavefront *w = nullptr;
    cudaStream_t stream = nullptr;
    int n_paths = 0, npix = 1, pixel0 = 0, sample0 = 0, max_depth = 0, mode = 0, nl = 0;
    uint64_t seed = 0;
    bool first_walk = true, profile = false, pdl = false, active = false;
    double *prof_ms = nullptr;
    int32_t seq0 = 0;
    int q = 0, depth = 0, consumed = 0;
    bool dead = false, final_walk = false;
    long long live_bound = 0; // no queue from here on is longer
    long long tail_paths = 0; // at or below this many live paths the rest of the batch is one k_finish launch
    size_t prof_used = 0;

    size_t slotOf(int it) const { return (size_t)(it % Wavefront::kRing) * Wavefront::kSlot; }
    // is the snapshot of depth `consumed` on its way?  A depth's counters are published by the walk of the NEXT depth, or
    // by the final walk
    bool awaited() const { return consumed < depth && (consumed < depth - 1 || final_walk); }
    bool arrived(int it) const
    {
        const volatile int32_t *c = w->h_ring + slotOf(it);
        return c[kNumCounters] == seq0 + 1 + it;
    }
    void consume(int it) // counters as they stood after k_shade of iteration `it`
    {
        const volatile int32_t *c = w->h_ring + slotOf(it);
        const int next_live = c[(it & 1) ^ 1];
        uint64_t shadow = 0;
        for (int l = 0; l < nl; ++l)
            shadow += (uint64_t)c[kShadowCount + l];
        s->stats.rays_shadow += shadow + (uint64_t)c[kTailShadow]; // (the spare counters are zero except in k_finish's snapshot)
        s->stats.rays_closest += (uint64_t)next_live + (uint64_t)c[kTailClosest]; // traced by iteration it + 1 / by k_finish
        if (!dead)
            live_bound = next_live;
        if (next_live == 0)
            dead = true;
    }
    int stamp() // a timestamp on the stream between two launches
    {
        if (prof_used == w->prof_ev.size())
        {
            cudaEvent_t e;
            TRT_CUDA(cudaEventCreate(&e));
            w->prof_ev.push_back(e);
        }
        TRT_CUDA(cudaEventRecord(w->prof_ev[prof_used++], stream));
        return TRT_OK;
    }
    void walk(int qsel, int what, long long rays_bound, int publish_it)
    {
        const WfBuffers &b = w->buf;
        const long long full_walk = (long long)s->sm_count * std::max(1, w->blocks_walk), full_plain = (long long)s->sm_count * 16;
        // grids follow the queue: its length two iterations ago bounds it (queues only shrink)
        const long long need = std::max(1ll, (rays_bound + kBlock - 1) / kBlock);
        int32_t *snap = (publish_it >= 0 && (what & 4)) ? w->d_ring + slotOf(publish_it) : nullptr;
        const int32_t seq = seq0 + 1 + publish_it;
        if (mode == 1)
            launchPdl(k_walk<1, false>, (unsigned)std::min(full_plain, need), kBlock, stream, pdl, s->view, b, qsel, what, snap, seq);
        else if (mode == 2)
            launchPdl(k_walk<2, false>, (unsigned)std::min(full_plain, need), kBlock, stream, pdl, s->view, b, qsel, what, snap, seq);
        else if (s->shadow_stop)
            launchPdl(k_walk<0, true>, (unsigned)std::min(full_walk, need), kBlock, stream, pdl, s->view, b, qsel, what, snap, seq);
        else
            launchPdl(k_walk<0, false>, (unsigned)std::min(full_walk, need), kBlock, stream, pdl, s->view, b, qsel, what, snap, seq);
        s->stats.kernel_launches++;
    }
    int launchDepth()
    {
        int rc;
        // the shadow rays of this walk come from the previous depth's vertices: at most n_lights per path of a
        // queue that was no longer than the bound either
        const long long sh_bound = depth > 0 ? live_bound * nl : 0;
        const bool walk_closest = depth > 0 || first_walk;
        if (profile)
        {
            if ((rc = stamp()))
                return rc;
            if (walk_closest)
                walk(q, 1, live_bound, -1);
            if ((rc = stamp()))
                return rc;
            if (depth > 0)
                walk(q, 2 | 4, sh_bound, depth - 1);
            if ((rc = stamp()))
                return rc;
        }
        else if (depth > 0)
            walk(q, 1 | 2 | 4, live_bound + sh_bound, depth - 1);
        else if (walk_closest)
            walk(q, 1 | 4, live_bound, -1);
        if (live_bound <= tail_paths)
        {
            // few paths left: this depth's vertices and everything after them in one launch
            const long long grid = std::min((long long)s->sm_count * 32, std::max(1ll, (live_bound + kShadeBlock - 1) / kShadeBlock));
            launchPdl(k_finish, (unsigned)grid, kShadeBlock, stream, pdl, s->view, w->buf, q, depth, max_depth, sample0, seed, npix, pixel0,
                      w->d_ring + slotOf(depth), seq0 + 1 + depth);
            s->stats.kernel_launches++;
            ++depth;
            w->seq = seq0 + depth;
            final_walk = true;
            TRT_CUDA(cudaGetLastError());
            return TRT_OK;
        }
        // k_shade: at most the resident CTAs (grid-stride loop inside): a second, partly filled wave of CTAs would cost a
        // whole extra pass of ~40 us warp iterations
        const long long full_shade = (long long)s->sm_count * std::max(1, w->blocks_shade);
        const long long shade_grid = std::min(full_shade, std::max(1ll, (live_bound + kShadeBlock - 1) / kShadeBlock));
        launchPdl(k_shade, (unsigned)shade_grid, kShadeBlock, stream, pdl, s->view, w->buf, q, depth, max_depth, sample0, seed, npix, pixel0);
        s->stats.kernel_launches++;
        if (profile && (rc = stamp()))
            return rc;
        q ^= 1;
        ++depth;
        TRT_CUDA(cudaGetLastError());
        return TRT_OK;
    }
};

static inline void hostPause()
{
#if defined(__x86_64__)
    __builtin_ia32_pause(); // one host thread per GPU polls under trt_render_multi: be a quiet spinner
#endif
}

// drives one loop to its end (the single-lane users: profile renders, trt_shade)
static int runToEnd(DepthLoop &loop)
{
    for (unsigned idle = 0; !loop.finished();)
    {
        const int r = loop.step();
        if (r < 0)
            return r;
        if (r > 0)
        {
            idle = 0;
            continue;
        }
        hostPause();
        if ((++idle & 0x3ff) == 0)
        {
            const int rc = loop.checkStalled();
            if (rc)
                return rc;
        }
    }
    return loop.finishProfile();
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
    const int nl1 = std::max(1, s->view.n_lights);
    long long max_paths = 0;
    const long long target = batchTarget(s, p.batch_paths, max_paths);
    const int total_samples = p.sample_end - p.sample_begin;
    if (total_samples <= 0)
        return TRT_OK;
    const int spb = (int)std::max(1ll, std::min((long long)total_samples, target / npix));
    if (npix * spb > max_paths)
    {
        setLastError("batch too large for 29-bit ray tokens (paths x lights must stay below 2^29)");
        return TRT_ERR_LIMIT;
    }
    int rc = ensureWavefront(s, (int)(npix * spb));
    if (rc)
        return rc;
    Wavefront *w = s->wf;
    const WfBuffers &b = w->buf;
    const bool pdl = usePdl() && !(p.flags & TRT_RENDER_PROFILE);
    double prof_ms[3] = {0, 0, 0};
    TRT_CUDA(cudaEventRecord(w->ev0, stream));
    for (int s0 = p.sample_begin; s0 < p.sample_end; s0 += spb)
    {
        const int ns = std::min(spb, p.sample_end - s0);
        const int n_paths = (int)(npix * ns);
        k_raygen<<<(unsigned)((n_paths + kBlock - 1) / kBlock), kBlock, 0, stream>>>(s->view, b, n_paths, s0, p.seed);
        s->stats.kernel_launches++;
        s->stats.paths += (uint64_t)n_paths;
        s->stats.rays_closest += (uint64_t)n_paths; // iteration 0 traces every path's camera ray
        DepthLoop loop;
        loop.begin(s, w, stream, n_paths, (int)npix, 0, s0, p.seed, p.max_depth, p.flags, true, prof_ms);
        if ((rc = runToEnd(loop)))
            return rc;
        launchPdl(k_deposit, (unsigned)((npix + 255) / 256), 256u, stream, pdl, b, d_accum, (int)npix, ns, nl1);
        s->stats.kernel_launches++;
        TRT_CUDA(cudaGetLastError());
    }
    TRT_CUDA(cudaEventRecord(w->ev1, stream));
    TRT_CUDA(cudaStreamSynchronize(stream));
    float ms = 0;
    TRT_CUDA(cudaEventElapsedTime(&ms, w->ev0, w->ev1));
    s->stats.last_render_ms = ms;
    s->stats.ms_trace = prof_ms[0], s->stats.ms_shadow = prof_ms[1], s->stats.ms_shade = prof_ms[2];
    s->stats.ms_accumulate = 0.0; // folded into k_shade / k_deposit (settleVertex)
    return TRT_OK;
}

// shade(hit, wi) for n already-traced rays (device pointers): radiance3[i] = what the reference's recursive shade()
// returns for the record of ray i (pathTracing.cpp:3-102), wi = -direction.  Stream of ray i: pixel = first + i.
int shadeBatch(trt_scene *s, const float *d_rays6, const int32_t *d_id, const float *d_t, size_t n, const trt_shade_params &p,
               float *d_radiance3, cudaStream_t stream)
{
    if (s->view.n_lights > kMaxLights)
    {
        setLastError("more than 32 lights");
        return TRT_ERR_LIMIT;
    }
    if (n == 0)
        return TRT_OK;
    const int nl1 = std::max(1, s->view.n_lights);
    long long max_paths = 0;
    const long long target = std::min<long long>(batchTarget(s, 0, max_paths), (long long)n);
    int rc = ensureWavefront(s, (int)target);
    if (rc)
        return rc;
    Wavefront *w = s->wf;
    const WfBuffers &b = w->buf;
    if (n > 0x7fffffffull)
    {
        setLastError("trt_shade: more than 2^31 - 1 rays in one call");
        return TRT_ERR_LIMIT;
    }
    for (size_t off = 0; off < n; off += (size_t)target)
    {
        const int m = (int)std::min<size_t>((size_t)target, n - off);
        k_inject<<<(unsigned)((m + kBlock - 1) / kBlock), kBlock, 0, stream>>>(b, d_rays6 + off * 6, d_id + off, d_t + off, m);
        s->stats.kernel_launches++;
        s->stats.paths += (uint64_t)m;
        // stream of ray i: pixel key = i = chunk offset + slot (npix = INT_MAX keeps slot % npix = slot, slot / npix = 0)
        DepthLoop loop;
        loop.begin(s, w, stream, m, 0x7fffffff, (int)off, p.sample, p.seed, p.max_depth, p.flags & ~TRT_RENDER_PROFILE, false, nullptr);
        if ((rc = runToEnd(loop)))
            return rc;
        k_collect<<<(unsigned)((m + kBlock - 1) / kBlock), kBlock, 0, stream>>>(b, d_radiance3 + off * 3, m, nl1);
        s->stats.kernel_launches++;
        TRT_CUDA(cudaGetLastError());
    }
    TRT_CUDA(cudaStreamSynchronize(stream));
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
