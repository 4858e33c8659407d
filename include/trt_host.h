/* trt_host.h — C view of the HOST side that the drop-in surface keeps (loaders + buildBVH), for bindings
 * that are not C++ (the ctypes test / bench harness).  C++ callers use csrc/host/tinyrt.h directly.
 *   loaders   scene.cpp:3-213 (readxml, readobj, readmtl; order main.cpp:66-69)
 *   buildBVH  bvh.cpp:16-144  (leaf_num = 8 at main.cpp:76)
 * Nothing here touches the GPU. */
#ifndef TRT_HOST_H
#define TRT_HOST_H
#include "trt.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct trt_host_scene trt_host_scene;

/* readxml -> readobj -> readmtl -> buildBVH(leaf_num). */
int trt_host_scene_load(const char *xml_path, const char *obj_path, const char *mtl_path, const char *basedir,
                        int leaf_num, trt_host_scene **out);

/* Same pipeline for geometry that does not come from files (synthetic meshes): triangles in "OBJ order",
 * derived fields (normal, center, cumulative light area) computed as scene.cpp:196-205 does, then buildBVH.
 * materials: n_materials records (is_emissive / radiance / area are derived from `light_materials`).
 * camera: eye, lookat, up (3 floats each), fovy degrees, width, height -> Camera::setCamera. */
int trt_host_scene_from_arrays(int32_t n_tris, const float *v9, const float *vn9, const float *vt6,
                               const int32_t *mtl, int32_t n_materials, const trt_material *materials,
                               int32_t n_lights, const int32_t *light_materials, const float *light_radiance3,
                               const float *eye, const float *lookat, const float *up, double fovy, int32_t width,
                               int32_t height, int leaf_num, trt_host_scene **out);

/* Binary scene cache: everything trt_host_scene_desc / _faces / _material_name return — the parsed
 * OBJ / MTL / XML, the decoded textures and the buildBVH topology (scene.cpp:3-213, bvh.cpp:16-144 never run again) — in
 * one file with a format version and a checksum.  trt_host_scene_load_cache gives a scene that is bit-identical in
 * every array to the one that was saved; a truncated, corrupted or foreign file is an error, never a partial scene. */
int trt_host_scene_save(trt_host_scene *s, const char *path);
int trt_host_scene_load_cache(const char *path, trt_host_scene **out);

/* POD view (valid until trt_host_scene_free); feed it to trt_scene_create. */
const trt_scene_desc *trt_host_scene_desc(trt_host_scene *s);
/* post-build triangle index -> ordinal of the OBJ `f` statement (or input index for from_arrays) */
const int32_t *trt_host_scene_faces(trt_host_scene *s);
const char *trt_host_scene_material_name(trt_host_scene *s, int i);
double trt_host_scene_build_seconds(trt_host_scene *s);
void trt_host_scene_free(trt_host_scene *s);

/* Baseline JPEG -> 8-bit BGR rows x cols x 3, bit-identical to cv::imread (what material.cpp:6 calls) on such
 * files.  Call with bgr_out = NULL to get the size first. */
int trt_decode_jpeg(const char *path, int32_t *rows, int32_t *cols, uint8_t *bgr_out, size_t capacity);

/* Uncompressed 8-bit RGB(A) PNG, the role svpng.inc:77-108 plays in the reference (main.cpp:40). */
int trt_write_png(const char *path, int32_t w, int32_t h, const uint8_t *rgb, int alpha);

/* The LINEAR image (the double[W*H*3] buffer of main.cpp:74, before imshow's gamma and 8-bit quantisation) as a
 * 32-bit float PFM file ("PF", little-endian, rows bottom-to-top as the format prescribes): the reference can only
 * write the quantised PNG, which is useless for comparing renders numerically. */
int trt_write_pfm(const char *path, int32_t w, int32_t h, const double *rgb_linear);

#ifdef __cplusplus
}
#endif
#endif
