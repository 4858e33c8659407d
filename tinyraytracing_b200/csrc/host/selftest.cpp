// Self-test of the C++ drop-in surface (csrc/host/tinyrt.h) on a GPU box: the same calls the reference's main()
// makes — readxml / readobj / readmtl, buildBVH, traverseBVH, the sample loop — against the GPU-backed versions.
// usage: trt_host_selftest <basedir> <mtl> <xml> <obj>     exit code 0 = all checks passed
#include "tinyrt.h"

#include <cmath>
#include <cstdio>
#include <cstring>

using namespace trt;

#define CHECK(cond)                                                                                                 \
    do                                                                                                              \
    {                                                                                                               \
        if (!(cond))                                                                                                \
        {                                                                                                           \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);                                           \
            return 1;                                                                                               \
        }                                                                                                           \
    } while (0)

int main(int argc, char **argv)
{
    if (argc != 5)
    {
        std::printf("usage: %s <basedir> <mtl> <xml> <obj>\n", argv[0]);
        return 2;
    }
    try
    {
        Scene scene;
        scene.readxml(argv[3]);
        scene.readobj(argv[4]);
        scene.readmtl(argv[2], argv[1]);
        CHECK(!scene.triangles.empty() && !scene.lights.empty());
        BVHNode *root = buildBVH(scene.triangles, 0, (int)scene.triangles.size() - 1, 8);
        CHECK(root != nullptr && nodeCountBVH(root) >= 1);
        DeviceScene dev(scene, root, 0);

        // centre pixel: single-ray form with the signature shape of bvh.h:32
        Ray centre = scene.camera.getRay(0.5f, 0.5f);
        HitRecord rec = traverseBVH(centre, dev);
        CHECK(rec.is_hit && rec.distance > 0.0005f && rec.distance < INF);
        CHECK(rec.triangle_index >= 0 && rec.triangle_index < (int)scene.triangles.size());
        CHECK(rec.triangle.mtl_name == scene.triangles[rec.triangle_index].mtl_name);
        // hitpoint = S + d * t in float (bvh.cpp:191), bit for bit
        const vec3 P = centre.startpoint + centre.direction * rec.distance;
        CHECK(std::memcmp(&P, &rec.hitpoint, sizeof P) == 0);
        CHECK(std::fabs(length(rec.pn) - 1.0f) < 1e-4f);
        // host-side barycentrics reproduce the hit point (least squares of triangle.cpp:12-29)
        const vec3 b = rec.triangle.findBaryCor(rec.hitpoint);
        const vec3 back = rec.triangle.v[0] * b.x + rec.triangle.v[1] * b.y + rec.triangle.v[2] * b.z;
        CHECK(length(back - rec.hitpoint) < 1e-3f * (1.0f + length(rec.hitpoint)));

        // batch form: a ray that starts far outside and flies away is a miss with the HitRecord defaults (bvh.h:9-10)
        std::vector<Ray> rays = {centre, Ray(vec3(1.0e5f, 2.0e5f, 3.0e5f), normalize(vec3(1.f, 2.f, 3.f)))};
        std::vector<HitRecord> out = traverseBVH(rays, dev);
        CHECK(out.size() == 2 && out[0].is_hit && out[0].triangle_index == rec.triangle_index);
        CHECK(!out[1].is_hit && out[1].distance == INF && out[1].triangle_index == -1);

        // the sample loop: deterministic for a seed, non-negative, not black
        const size_t n = (size_t)scene.img_width * scene.img_height * 3;
        std::vector<double> a(n), c(n);
        renderImage(dev, 2, a.data(), 5);
        renderImage(dev, 2, c.data(), 5);
        CHECK(std::memcmp(a.data(), c.data(), n * sizeof(double)) == 0);
        double sum = 0;
        for (double v : a)
        {
            CHECK(v >= 0.0 && std::isfinite(v));
            sum += v;
        }
        CHECK(sum > 0.0);
        freeBVH(root);
    }
    catch (const std::exception &e)
    {
        std::printf("FAILED with exception: %s\n", e.what());
        return 1;
    }
    std::printf("host selftest ok\n");
    return 0;
}
