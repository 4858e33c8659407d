/* trt.h — C ABI of the B200-native hot path of TinyRayTracing (libtrt_b200.so).
 *
 * The reference has no plugin / FFI interface: its hot path is reached by plain C++ calls
 *   main.cpp:76      buildBVH(scene.triangles, 0, n-1, 8)
 *   main.cpp:95-101  Camera::getRay -> traverseBVH -> shade            (bvh.h:28-32, pathtracing.h:14-17)
 *   main.cpp:79-113  the sample / pixel loop that accumulates into `double image[W*H*3]`
 * This header is the boundary that replaces the loop body: the host keeps the reference's Scene / Camera /
 * Material / BVH classes and loaders (csrc/host), converts them ONCE into the POD arrays below after buildBVH,
 * and every traversal / shading step then runs in hand-written sm_100a kernels.
 *
 * Conventions: every call returns 0 on success or a negative trt_status; trt_last_error() gives the text.
 * No call exits the process (C++ exceptions are caught at the boundary and returned as codes), none falls back to
 * the CPU: without an sm_100 device trt_scene_create fails.
 * Calls on one trt_scene must be serialised by the caller (the reference's main is single-threaded outside
 * its OpenMP loop, which this library replaces).  Plain pointers and sizes only — no C++ / torch types.
 */
#ifndef TRT_H
#define TRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRT_VERSION 201

typedef enum trt_status {
    TRT_OK = 0,
    TRT_ERR_INVALID = -1,   /* bad argument / inconsistent description            */
    TRT_ERR_NO_DEVICE = -2, /* no sm_100 GPU: there is deliberately no CPU path   */
    TRT_ERR_CUDA = -3,      /* a CUDA runtime call failed (text in trt_last_error) */
    TRT_ERR_LIMIT = -4,     /* a documented capacity limit was exceeded           */
    TRT_ERR_NCCL = -5,      /* libnccl.so.2 could not be loaded, or an NCCL call failed (trt_render_multi) */
    TRT_ERR_IO = -6         /* a checkpoint file could not be written / read / does not match the scene    */
} trt_status;

/* reference INF (bvh.h:5): distance reported for a miss */
#define TRT_INF 114514.0f

/* Material record = the fields of reference `Material` (material.h:11-33) that shade()/nextRay() read. */
typedef struct trt_material {
    float Kd[3], Ks[3], Tr[3];
    float Ns, Ni;
    float radiance[3];   /* scene.cpp:52 */
    int32_t is_emissive; /* scene.cpp:51 */
    int32_t texture;     /* index into textures, -1 when map_Kd == "" (material.h:20) */
    double area;         /* total light area, scene.cpp:202 */
} trt_material;

/* One <light> of the XML, in XML order (scene.cpp:23-54).  Its triangles are the reference's
 * materials[mtl].triangles (scene.cpp:204): OBJ order, `cum_area` = running sum (scene.cpp:201-203). */
typedef struct trt_light {
    int32_t material;  /* index into materials */
    int32_t first_tri; /* offset into light_v / light_vn / light_cum_area */
    int32_t n_tris;
    int32_t _pad;
} trt_light;

typedef struct trt_texture {
    int32_t rows, cols;  /* cv::Mat rows / cols (material.cpp:10) */
    const uint8_t *bgr;  /* rows*cols*3, OpenCV BGR byte order as cv::imread returns it */
} trt_texture;

/* Scene description: POD view of the reference's Scene AFTER buildBVH (post-build triangle order).
 * Triangle identity = position in Scene::triangles after buildBVH (the reference has no id field,
 * triangle.h:25 is commented out). */
typedef struct trt_scene_desc {
    int32_t n_tris;
    const float *v;      /* n_tris*9 : v[0].xyz v[1].xyz v[2].xyz           (triangle.h:17) */
    const float *vn;     /* n_tris*9 : vertex normals                        (triangle.h:18) */
    const float *vt;     /* n_tris*6 : vertex uvs                            (triangle.h:19) */
    const float *normal; /* n_tris*3 : face normal EXACTLY as scene.cpp:196 computed it     */
    const int32_t *mtl;  /* n_tris   : material index                                         */

    /* reference BVH topology, pre-order array of the pointer tree buildBVH returns (bvh.cpp:16-144) */
    int32_t n_nodes;
    const float *node_box;    /* n_nodes*6 : AA.xyz BB.xyz (bvh.h:21)                          */
    const int32_t *node_link; /* n_nodes*4 : left, right (node indices, -1 = NULL), index, num */

    int32_t n_materials;
    const trt_material *materials;
    int32_t n_lights;
    const trt_light *lights;
    int32_t n_light_tris;
    const float *light_v;         /* n_light_tris*9 */
    const float *light_vn;        /* n_light_tris*9 */
    const double *light_cum_area; /* n_light_tris   */
    int32_t n_textures;
    const trt_texture *textures;

    /* camera as computed by Camera::setCamera on the host (camera.cpp:3-17) */
    float eye[3], lower_left_corner[3], horizontal[3], vertical[3];
    int32_t width, height;
} trt_scene_desc;

typedef struct trt_scene trt_scene;

/* ---- lifetime ------------------------------------------------------------------------------------ */

/* Number of sm_100 devices among the CUDA devices of this process (0 when none / no driver).  `device` arguments
 * below are plain CUDA ordinals (cudaSetDevice numbering, after CUDA_VISIBLE_DEVICES): on a box that mixes GPU
 * generations the usable ordinals need not be 0 .. count-1; trt_scene_create says so for an ordinal that is not sm_100. */
int trt_device_count(void);

/* Copies the description to `device` and builds the GPU acceleration layout from the reference topology.
 * The caller keeps ownership of every input pointer; nothing is referenced after the call returns.
 * Replaces: nothing in the reference (there the Scene is used in place); called once after main.cpp:76. */
int trt_scene_create(const trt_scene_desc *desc, int device, trt_scene **out);
void trt_scene_destroy(trt_scene *scene);

/* trt_scene_create with the acceleration layouts kept in a file (since v201): when `layout_cache_path` holds the layouts
 * of exactly this description — the file is keyed by a hash of every array and count the layouts are built from and of
 * the environment switches that shape them, and carries a checksum — they are read instead of built (the build is most
 * of trt_scene_create's host time on large scenes: 17 s of it at 10 M triangles).  Otherwise they are built as usual and
 * the file is (re)written; a missing, stale, foreign, truncated or corrupt file is never an error, and neither is a file
 * that cannot be written.  *from_cache (may be NULL) tells which happened.  The scene is the same either way.
 * Replaces: nothing in the reference (it rebuilds its tree on every run, main.cpp:76). */
int trt_scene_create_cached(const trt_scene_desc *desc, int device, const char *layout_cache_path, int32_t *from_cache,
                            trt_scene **out);

/* A copy of `scene` on another device, made from the device-resident arrays (device-to-device copies; the layouts
 * are not rebuilt on the host): how the scene is replicated for trt_render_multi ("BVH build: host, once,
 * broadcast by plain cudaMemcpy to each device").  The copy is independent of the original afterwards. */
int trt_scene_replicate(const trt_scene *scene, int device, trt_scene **out);

/* Page-locked host memory for ray / result buffers: the blocking entry points DMA straight from / to such
 * buffers; pageable buffers are staged through internal pinned chunks (one extra host copy). */
void *trt_host_alloc(size_t bytes);
void trt_host_free(void *p);

/* ---- closest hit: replaces traverseBVH (bvh.cpp:146-175) for a batch of rays --------------------- */

#define TRT_TRACE_DEVICE_PTRS 1u /* rays / outputs are device pointers on the scene's device          */
#define TRT_TRACE_EXHAUSTIVE 2u  /* walk the reference topology with the reference's visiting rule
                                    (no ordering, no pruning: bvh.cpp:156-174) — validation mode     */
#define TRT_TRACE_REFTOPO 4u     /* ordered + pruned walk of the reference binary topology          */
#define TRT_TRACE_PLAIN 8u       /* fast layout, one thread per ray without the persistent ray pool  */
#define TRT_TRACE_POOLED 16u     /* persistent walker with warp-pooled inside tests (experimental)   */
#define TRT_TRACE_PERSISTENT 32u /* the warp-persistent walker even where the library would pick the plain kernel
                                    (scenes of a few nodes); results are identical in every mode                   */

/* rays6: n*(origin.xyz, direction.xyz) float32.  tri_id: post-build triangle index of the reference's
 * winner (tie rule of bvh.cpp:168-172,219), -1 on miss.  t: HitRecord::distance, TRT_INF on miss.
 * Either output may be NULL.  With host pointers the call stages through pinned memory in chunks and
 * overlaps copies with kernels; it returns after the results are in the output arrays. */
int trt_trace_closest(trt_scene *scene, const float *rays6, size_t n, int32_t *tri_id, float *t, uint32_t flags);

/* Asynchronous device-pointer form on a caller-provided CUDA stream (cudaStream_t passed as void*).  Launches of one
 * scene may be in flight on several streams at once (each takes its own ray-pool cursor from a ring of 64); the CALLS
 * themselves must still be serialised by the caller like every call on a trt_scene. */
int trt_trace_closest_async(trt_scene *scene, const float *d_rays6, size_t n, int32_t *d_tri_id, float *d_t,
                            uint32_t flags, void *stream);

/* The same for one host batch spread over several GPUs of the box: scenes[i] are replicas of one scene on different
 * devices (trt_scene_replicate); GPU i traces rays [i*n/k, (i+1)*n/k) — the batch shards by ray index, there is no
 * exchange step — each through its own chunked copy / kernel pipeline, driven by one host thread per GPU.  Host
 * pointers only (TRT_TRACE_DEVICE_PTRS is refused).  Results are those of trt_trace_closest on any one of the scenes. */
int trt_trace_closest_multi(trt_scene *const *scenes, int32_t k, const float *rays6, size_t n, int32_t *tri_id, float *t,
                            uint32_t flags);

/* Hit attributes the reference stores in HitRecord (bvh.h:7-15) for already-traced rays: hit point
 * S + d*t (bvh.cpp:191) and shading normal pn (bvh.cpp:223-224, least-squares barycentrics of
 * triangle.cpp:12-29 in double).  Host pointers. Either output may be NULL. */
int trt_hit_attributes(trt_scene *scene, const float *rays6, const int32_t *tri_id, const float *t, size_t n,
                       float *hitpoint3, float *pn3);

/* ---- shade: replaces PathTracing::shade (pathtracing.h:14, pathTracing.cpp:3-102) for a batch of hits ---------- */

typedef struct trt_shade_params {
    uint64_t seed;     /* Philox key; the stream of ray i is (seed; pixel = i, sample, bounce, slot)              */
    int32_t sample;    /* sample index of the streams (lets a caller shade the same hits with fresh numbers)     */
    int32_t max_depth; /* 0 = reference behaviour: Russian roulette only                                         */
    uint32_t flags;    /* TRT_RENDER_REFTOPO / TRT_RENDER_PLAIN                                                  */
    uint32_t _pad;
} trt_shade_params;

/* radiance3[i] = shade(hit_i, -direction_i): what the reference's recursive shade() returns for the HitRecord of ray i
 * — emitted radiance on a light, else next-event estimation over the lights (XML order) plus the Russian-roulette
 * bounce chain (nextRay / Sample / RR) — run as one wavefront that starts from the caller's hits instead of camera
 * rays.  tri_id / t as trt_trace_closest returned them for rays6 (tri_id < 0: no hit, radiance 0 as main.cpp:99-101).
 * Host pointers.  Not divided by anything. */
int trt_shade(trt_scene *scene, const float *rays6, const int32_t *tri_id, const float *t, size_t n,
              const trt_shade_params *params, float *radiance3);

/* Work done by the default traversal on a host ray batch, summed over rays: out4 = {wide nodes visited, child
 * box tests, reference leaves scanned, triangles in those leaves}.  Reporting only (rays that take the strict
 * walk are not counted). */
int trt_trace_counters(trt_scene *scene, const float *rays6, size_t n, uint64_t out4[4]);

/* ---- render: replaces the loop body main.cpp:79-113 (getRay, traverseBVH, shade, accumulate) ----- */

typedef struct trt_render_params {
    int32_t spp;          /* SAMPLE (main.cpp:12,55): the image is divided by this                       */
    int32_t sample_begin; /* this call renders samples [sample_begin, sample_end) of every pixel        */
    int32_t sample_end;   /*   (sample-range sharding across GPUs; 0,spp for the whole job)              */
    int32_t max_depth;    /* 0 = reference behaviour: unbounded, Russian roulette only (pathtracing.h:12) */
    uint64_t seed;        /* Philox4x32-10 key; streams are keyed (pixel, sample, bounce, slot)           */
    int32_t batch_paths;  /* paths in flight per wavefront batch, 0 = default                             */
    uint32_t flags;       /* TRT_RENDER_*                                                                 */
} trt_render_params;

#define TRT_RENDER_REFTOPO 1u /* trace with the reference-topology kernel instead of the fast layout */
#define TRT_RENDER_PLAIN 2u   /* fast layout, plain thread-per-ray traversal instead of the persistent walker */
#define TRT_RENDER_PROFILE 4u /* time each kernel class with CUDA events (trt_stats.ms_*); the image is unchanged.  The
                                 closest-hit and the shadow rays of a depth are then walked in two launches instead of one */
#define TRT_RENDER_PEER_REDUCE 8u /* trt_render_multi: sum the per-GPU buffers with this library's own kernel over NVLink
                                     peer memory (fused with the resolve) instead of ncclReduce */

/* Renders the sample range and writes the reference's image buffer: double[H*W*3], row-major RGB, rows
 * top to bottom, already divided by spp (main.cpp:74,101-108) — what imshow (main.cpp:19-42) consumes. */
int trt_render(trt_scene *scene, const trt_render_params *params, double *image_rgb);

/* Multi-GPU building block: ADDS the per-pixel radiance sums (not divided by spp) of the sample range
 * into a device buffer double[H*W*3] on `stream`; ranks then sum their buffers with one NCCL reduce and
 * call trt_resolve.  d_accum must be zero-initialised by the caller before the first call. */
int trt_render_accumulate(trt_scene *scene, const trt_render_params *params, double *d_accum, void *stream);

/* image = accum / spp (main.cpp:101); optional 8-bit gamma-2.2 pack as imshow (main.cpp:30-38).
 * d_accum device pointer; image_rgb / rgb8 host pointers, either may be NULL. */
int trt_resolve(trt_scene *scene, const double *d_accum, int32_t spp, double *image_rgb, uint8_t *rgb8,
                void *stream);

/* ---- multi-GPU render inside the library: replaces the reference's own fan-out, the OpenMP loop over samples of
 * main.cpp:79-81.  scenes[i] are n scenes created from the SAME description on n different devices of one box (the
 * scene is replicated; trt_scene_create once per device).  One host thread per GPU renders samples
 * [sample_begin + i*k/n ..) of every pixel into that GPU's accumulation buffer (trt_render_accumulate), the buffers
 * are summed onto scenes[0]'s device — ncclReduce(sum, float64, W*H*3) over NVLink on communicators from
 * ncclCommInitAll (libnccl.so.2 is loaded on first use), or, with TRT_RENDER_PEER_REDUCE in params->flags, by one
 * kernel of this library that reads the peers' buffers directly and resolves in the same pass — and scenes[0] resolves.
 * image_rgb / rgb8 as trt_resolve (host pointers, either may be NULL).  n = 1 is trt_render.  With the peer reduce
 * two scenes may share a device (NCCL refuses duplicate devices).  trt_stats of scenes[0]: last_render_ms covers all
 * GPUs' renders, the reduce and the resolve; ray counts stay per scene. */
int trt_render_multi(trt_scene *const *scenes, int32_t n, const trt_render_params *params, double *image_rgb,
                     uint8_t *rgb8);

/* ---- accumulation checkpoints (the reference writes its image once, main.cpp:114) --------------------
 * A device accumulation buffer (what trt_render_accumulate adds into) created zeroed / released by the library, for
 * hosts that do not link the CUDA runtime themselves. */
double *trt_accum_create(trt_scene *scene);
void trt_accum_destroy(trt_scene *scene, double *d_accum);
/* Writes / reads the buffer and the progress it stands for to a file: samples [0, samples_done) of `spp` rendered with
 * `seed` and `max_depth`.  A render resumed from a checkpoint — trt_accum_load, then trt_render_accumulate of
 * [samples_done, spp) — gives the image of an uninterrupted one bit for bit (every pixel-sample has its own Philox
 * stream, sums are double).  The file carries the frame size, a format version and a checksum: trt_accum_load fails
 * with TRT_ERR_IO on a truncated / corrupted file or one made for another frame size. */
int trt_accum_save(trt_scene *scene, const double *d_accum, int32_t samples_done, int32_t spp, uint64_t seed,
                   int32_t max_depth, const char *path);
int trt_accum_load(trt_scene *scene, const char *path, double *d_accum, int32_t *samples_done, int32_t *spp,
                   uint64_t *seed, int32_t *max_depth);

/* ---- introspection ------------------------------------------------------------------------------- */

typedef struct trt_stats {
    uint64_t rays_closest; /* closest-hit rays traced (primary + bounce) since creation / last reset   */
    uint64_t rays_shadow;  /* NEE "shadow" rays (closest hit + material compare, pathTracing.cpp:51-58) */
    uint64_t paths;        /* pixel-samples started                                                     */
    uint64_t kernel_launches;
    double last_render_ms; /* device time of the last trt_render / trt_render_accumulate (CUDA events)  */
    double last_trace_ms;  /* device time of the traversal kernel(s) of the last trt_trace_closest      */
    int32_t accel_nodes;   /* nodes of the GPU layout                                                   */
    int32_t accel_leaves;  /* reference leaves (bvh.cpp:43-48) kept as scan units                       */
    int32_t ref_depth;     /* depth of the reference tree                                               */
    int32_t device;
    int32_t accel_slivers; /* triangles boxed with the larger sliver pad (fast layout)                    */
    int32_t accel_needles; /* triangles left under their reference leaf's box (fast layout)               */
    /* device time per kernel class of the last render made with TRT_RENDER_PROFILE (CUDA events between the launches;
       the role the clock() prints of main.cpp:60-61,116-117 play in the reference), else 0 */
    double ms_trace, ms_shade, ms_shadow, ms_accumulate; /* ms_accumulate: always 0 since v200 (folded into the shade pass) */
    /* rays that took the reference's own exhaustive walk instead of the fast layout: non-finite rays and origins beyond
       8 x scene scale from the coordinate origin.  Exact results, but orders of magnitude slower per
       ray — a large number here explains a slow trace.  Counted on the device, read by trt_get_stats. */
    uint64_t rays_strict;
} trt_stats;

/* ---- host-only: build the GPU layouts for `desc` and verify their invariants (no device needed) ------------------
   What the exactness argument needs of the fast layout and a kernel cannot check: every triangle that
   interactTriangle (bvh.cpp:177-209) could accept is in exactly one leaf, carries the reference leaf (bvh.cpp:43-48)
   it belongs to, and lies — with the pad — inside every box on its path from the root; the tree fits the per-thread
   stack.  violations == 0 on success; returns TRT_ERR_INVALID with the first violation in trt_last_error otherwise. */
typedef struct trt_layout_report {
    int32_t n_tris;        /* triangles of the scene                                                       */
    int32_t n_fast_tris;   /* triangles in the fast layout                                                 */
    int32_t n_dropped;     /* left out: NaN face normal / non-finite vertex (can never be hit)              */
    int32_t use_wide;      /* 0: scene keeps the reference-topology kernels (single leaf / too deep / empty) */
    int32_t wide_nodes, wide_depth, ref_leaves, ref_depth, slivers, needles;
    int32_t violations;
    double sah_wide;       /* expected child-box tests per random ray through the root box (surface-area heuristic) */
    double sah_inner;      /* its part from child boxes that are inner nodes = expected node visits below the root   */
    double sah_leaf;       /* its part from child boxes that are leaves = expected leaf scans                        */
} trt_layout_report;
int trt_layout_check(const trt_scene_desc *desc, trt_layout_report *report);

/* ---- host-only: the GPU layouts as plain arrays (inspection and test tooling, no device needed) ------------------
   tests/ walks these arrays on the CPU with the traversal rules of DESIGN.md §3 (oracle/layout_walk.cpp, test
   infrastructure) to check the layout's DATA against the oracle, and tools count node visits per ray class with it.
   Nothing in the product path reads a trt_layout. */
typedef struct trt_layout_view {
    int32_t n_wide_nodes;         /* 0 when the scene is a single scan unit or has no fast layout              */
    int32_t wide_root;            /* node 0, a leaf link (< 0), or 0x7fffffff = no fast layout                 */
    int32_t n_fast_tris, n_ref_leaves, n_ref_inner;
    int32_t check_leaf_box;       /* 0 only when the whole scene is one reference leaf (bvh.cpp:151-154)       */
    float strict_origin_limit;    /* rays starting farther out take the exhaustive reference walk              */
    uint32_t miss_key;            /* tie key of the miss state                                    */
    const float *wide_nodes;      /* n_wide_nodes * 32 words: lox[4] loy[4] loz[4] hix[4] hiy[4] hiz[4] link[4] pad[4] */
    const float *fast_geom;       /* n_fast_tris * 12: N.xyz p1.xyz p2.xyz p3.xyz                              */
    const uint32_t *fast_key;     /* n_fast_tris: tie key, higher wins at equal t                              */
    const int32_t *fast_orig;     /* n_fast_tris: post-build triangle index                                    */
    const int32_t *fast_leaf;     /* n_fast_tris: reference leaf ordinal                                       */
    const float *ref_leaf_box;    /* n_ref_leaves * 8: AA.xyz - BB.xyz -                                       */
    const int32_t *ref_leaf_parent; /* n_ref_leaves: (parent inner index << 1) | right-child bit, -1 = root    */
    const float *ref_nodes;       /* n_ref_inner * 16 words: left box, right box (6 + 6 floats), left link, right
                                     link, parent token, unused                                                */
} trt_layout_view;
typedef struct trt_layout trt_layout;
int trt_layout_build(const trt_scene_desc *desc, trt_layout **out, trt_layout_view *view);
/* the same through the layout cache of trt_scene_create_cached (host only: what the cache tests use) */
int trt_layout_build_cached(const trt_scene_desc *desc, const char *layout_cache_path, int32_t *from_cache, trt_layout **out,
                            trt_layout_view *view);
void trt_layout_free(trt_layout *layout);

int trt_get_stats(trt_scene *scene, trt_stats *out);
int trt_reset_stats(trt_scene *scene);
const char *trt_last_error(void);
int trt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TRT_H */
