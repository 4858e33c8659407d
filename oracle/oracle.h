/* oracle.h — TEST INFRASTRUCTURE ONLY (oracle tier B).
 *
 * Deterministic CPU restatement of the reference's hot path, used as the parity checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg — never by the product path.
 * Pinned against tier A (the unmodified reference objects, oracle/_ref/libref.so) by tests/test_oracle_vs_ref.py:
 * identical post-build triangle order / node boxes, identical closest-hit distance + triangle on ray batches,
 * statistical image agreement.  The reference ships no golden vectors and no tests of its own.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_material {
    float Kd[3], Ks[3], Tr[3];
    float Ns, Ni;
    int32_t texture; /* -1: none */
} orc_material;

typedef struct orc_texture {
    int32_t rows, cols;
    const uint8_t *bgr;
} orc_texture;

typedef struct orc_scene orc_scene;

/* Triangles in OBJ order (pre-build), materials, lights in XML order (material index + radiance), camera
 * parameters as the XML gives them.  Derived fields as scene.cpp:196-205, camera as camera.cpp:3-17,
 * BVH as bvh.cpp:16-144 with leaf_num. */
orc_scene *orc_scene_create(int32_t n_tris, const float *v9, const float *vn9, const float *vt6, const int32_t *mtl,
                            int32_t n_materials, const orc_material *materials, int32_t n_lights,
                            const int32_t *light_mtl, const float *light_radiance3, int32_t n_textures,
                            const orc_texture *textures, const float *eye, const float *lookat, const float *up,
                            float fovy, int32_t width, int32_t height, int32_t leaf_num);
void orc_scene_destroy(orc_scene *s);

int32_t orc_num_nodes(orc_scene *s);
/* post-build order: perm[i] = input index of the triangle now at position i */
void orc_get_order(orc_scene *s, int32_t *perm);
/* derived per-triangle fields in post-build order (any may be NULL) */
void orc_get_derived(orc_scene *s, float *normal3, float *center3, double *cum_area, int32_t *emissive);
/* pre-order nodes: boxes[6n], links[4n] = left, right, index, num */
void orc_get_nodes(orc_scene *s, float *boxes, int32_t *links);
void orc_get_camera(orc_scene *s, float *out12); /* eye, llc, horizontal, vertical */
void orc_bvh_stats(orc_scene *s, int32_t *nodes, int32_t *leaves, int32_t *maxdepth);

/* traverseBVH (bvh.cpp:146-175) per ray, exhaustive recursion with the reference's merge rule.
 * id = post-build triangle index (-1 miss), t = distance (114514 on miss); pn / hitp optional. */
void orc_trace(orc_scene *s, const float *rays6, int64_t n, float *t, int32_t *id, float *pn3, float *hitp3,
               int32_t threads);

/* Work counters. mode 0: the reference's exhaustive walk; mode 1: ordered, distance-pruned binary walk of the
 * same topology with the A.4 tie key (the accounting basis of the roofline). Sums over rays.
 * Also returns ids of mode 1 in id_out (optional) so the pruned walk can be checked against orc_trace. */
void orc_trace_counts(orc_scene *s, const float *rays6, int64_t n, int32_t mode, uint64_t *box_tests,
                      uint64_t *tri_tests, int32_t *id_out, int32_t threads);

/* The render loop main.cpp:79-113 with shade()/nextRay()/Sample()/RR() of pathTracing.cpp:3-209, random
 * numbers from Philox4x32-10 keyed (seed; pixel, sample, depth, slot).  image: double[H*W*3], ADDS
 * color/spp per sample (zero it first).  max_depth 0 = unbounded (reference).  ray_counts[0] += closest-hit
 * rays (primary + bounce), ray_counts[1] += shadow rays. */
void orc_render(orc_scene *s, int32_t spp, int32_t sample_begin, int32_t sample_end, int32_t max_depth,
                uint64_t seed, double *image, uint64_t *ray_counts, int32_t threads);

/* The same loop for a chosen set of pixels (row-major pixel indices of the W x H frame): rgb_out[3*q..] += the
 * pixel's colour.  The Philox stream is keyed by the pixel index, so these are exactly the pixels orc_render would
 * produce — how the 1280x720 / 1920x1080 configs are checked without rendering two million pixels on the CPU. */
void orc_render_pixels(orc_scene *s, int32_t spp, int32_t sample_begin, int32_t sample_end, int32_t max_depth,
                       uint64_t seed, const int32_t *pixels, int32_t n_pixels, double *rgb_out, int32_t threads);

/* shade(hit, wi) (pathTracing.cpp:3-102, pathtracing.h:14) for a batch of already-traced rays: ray i hit post-build
 * triangle tri_id[i] at distance t[i] (miss: tri_id < 0 -> radiance 0, main.cpp:99-101); wi = -direction.  Stream of
 * ray i: (seed; pixel = i, sample).  radiance3 = the returned colour, not divided by anything. */
void orc_shade_batch(orc_scene *s, const float *rays6, const int32_t *tri_id, const float *t, int64_t n, uint64_t seed,
                     int32_t sample, int32_t max_depth, float *radiance3, int32_t threads);

/* One Philox4x32-10 block (known-answer tests) and the uniform double derived from a slot. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_uniform(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t depth, uint32_t slot);

/* Primary ray of (pixel row i, column j, sample k): origin + direction (main.cpp:88-95, camera.cpp:19-28). */
void orc_primary_ray(orc_scene *s, int32_t i, int32_t j, int32_t k, uint64_t seed, float *ray6);

int32_t orc_max_threads(void);
#ifdef __cplusplus
}
#endif
#endif
