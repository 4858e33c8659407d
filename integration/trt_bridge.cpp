// trt_bridge.cpp — the translation unit a maintainer ADDS to the reference (RayTracingOnCPU.vcxproj) to put its inner
// loop on the B200 path: converts the reference's own Scene / BVHNode (scene.h, bvh.h) into the POD description of
// include/trt.h once, after buildBVH (main.cpp:76).  Compiled against the UNMODIFIED reference headers by
// `make -C oracle bridge` (with integration/main.patch applied to a copy of the reference's main.cpp) and exercised by
// tests/test_gpu_bridge.py; INTEGRATION.md §2 walks through it.
#include "trt_bridge.h"
#include <functional>


TrtBridge *trt_bridge_create(Scene &sc, BVHNode *root, int device) {
    auto *b = new TrtBridge();
    auto id = [&](const std::string &n) {                       // material name -> index
        for (size_t i = 0; i < b->names.size(); i++) if (b->names[i] == n) return (int)i;
        b->names.push_back(n); return (int)b->names.size() - 1; };
    for (auto &l : sc.lights) id(l.mtl_name);
    for (auto &t : sc.triangles) {                               // post-build order = triangle identity
        for (int k = 0; k < 3; k++) {
            b->v.insert(b->v.end(), {t.v[k].x, t.v[k].y, t.v[k].z});
            b->vn.insert(b->vn.end(), {t.vn[k].x, t.vn[k].y, t.vn[k].z});
            b->vt.insert(b->vt.end(), {t.vt[k].x, t.vt[k].y}); }
        b->nrm.insert(b->nrm.end(), {t.normal.x, t.normal.y, t.normal.z});   // exactly scene.cpp:196's value
        b->mtl.push_back(id(t.mtl_name)); }
    std::function<int(BVHNode *)> flat = [&](BVHNode *n) {       // pre-order array of the pointer tree
        int me = (int)b->link.size() / 4;
        b->link.insert(b->link.end(), {-1, -1, n->index, n->num});
        b->box.insert(b->box.end(), {n->AA.x, n->AA.y, n->AA.z, n->BB.x, n->BB.y, n->BB.z});
        if (n->num == 0) { int l = flat(n->left);  b->link[me * 4 + 0] = l;
                           int r = flat(n->right); b->link[me * 4 + 1] = r; }
        return me; };
    if (root) flat(root);
    for (auto &name : b->names) {
        Material &m = sc.materials[name]; trt_material o{};
        o.Kd[0] = m.Kd.x; o.Kd[1] = m.Kd.y; o.Kd[2] = m.Kd.z;   o.Ks[0] = m.Ks.x; o.Ks[1] = m.Ks.y; o.Ks[2] = m.Ks.z;
        o.Tr[0] = m.Tr.x; o.Tr[1] = m.Tr.y; o.Tr[2] = m.Tr.z;   o.Ns = m.Ns; o.Ni = m.Ni;
        o.radiance[0] = m.radiance.x; o.radiance[1] = m.radiance.y; o.radiance[2] = m.radiance.z;
        o.is_emissive = m.is_emissive; o.area = m.area; o.texture = -1;
        if (m.map_Kd != "") { o.texture = (int)b->tex.size();    // cv::Mat: 8-bit BGR, continuous
                              b->tex.push_back({m.img.rows, m.img.cols, m.img.data}); }
        b->mats.push_back(o); }
    for (auto &l : sc.lights) {                                  // XML order; OBJ-order triangles, running area sums
        Material &m = sc.materials[l.mtl_name];
        b->lights.push_back({id(l.mtl_name), (int)b->lcum.size(), (int)m.triangles.size(), 0});
        for (auto &t : m.triangles) { for (int k = 0; k < 3; k++) {
                b->lv.insert(b->lv.end(), {t.v[k].x, t.v[k].y, t.v[k].z});
                b->lvn.insert(b->lvn.end(), {t.vn[k].x, t.vn[k].y, t.vn[k].z}); }
            b->lcum.push_back(t.area); } }
    trt_scene_desc &d = b->d;
    d.n_tris = (int)sc.triangles.size(); d.v = b->v.data(); d.vn = b->vn.data(); d.vt = b->vt.data();
    d.normal = b->nrm.data(); d.mtl = b->mtl.data();
    d.n_nodes = (int)b->link.size() / 4; d.node_box = b->box.data(); d.node_link = b->link.data();
    d.n_materials = (int)b->mats.size(); d.materials = b->mats.data();
    d.n_lights = (int)b->lights.size(); d.lights = b->lights.data();
    d.n_light_tris = (int)b->lcum.size(); d.light_v = b->lv.data(); d.light_vn = b->lvn.data(); d.light_cum_area = b->lcum.data();
    d.n_textures = (int)b->tex.size(); d.textures = b->tex.data();
    Camera &c = sc.camera;                                       // after Camera::setCamera (scene.cpp:24)
    float *dst[4] = {d.eye, d.lower_left_corner, d.horizontal, d.vertical};
    vec3 src[4] = {c.eye, c.lower_left_corner, c.horizontal, c.vertical};
    for (int k = 0; k < 4; k++) { dst[k][0] = src[k].x; dst[k][1] = src[k].y; dst[k][2] = src[k].z; }
    d.width = sc.img_width; d.height = sc.img_height;
    if (trt_scene_create(&d, device, &b->scene) != TRT_OK) { printf("%s\n", trt_last_error()); exit(1); }
    return b;
}
