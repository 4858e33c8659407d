// trt_bridge.h — declarations of integration/trt_bridge.cpp for the patched main.cpp.
#pragma once
#include "scene.h"
#include "bvh.h"
#include "trt.h"

#include <string>
#include <vector>

struct TrtBridge {
    std::vector<float> v, vn, vt, nrm, box, lv, lvn; std::vector<int32_t> mtl, link; std::vector<double> lcum;
    std::vector<trt_material> mats; std::vector<trt_light> lights; std::vector<trt_texture> tex;
    std::vector<std::string> names; trt_scene_desc d{}; trt_scene *scene = nullptr;
};
// Converts the reference's Scene + BVH (after buildBVH, main.cpp:76) and uploads it to `device`; exits with the
// library's message on failure, in the reference's own error style (scene.cpp:7-11).
TrtBridge *trt_bridge_create(Scene &sc, BVHNode *root, int device);
