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

        // PathTracing::shade (pathtracing.h:17): batch and single-record forms agree, emitters return their radiance,
        // a miss shades to black, the stream index changes the numbers of a non-emissive hit
        {
            std::vector<vec3> rad = shade(out, dev, 7, 0, 0);
            CHECK(rad.size() == 2 && rad[1].x == 0.f && rad[1].y == 0.f && rad[1].z == 0.f);
            CHECK(std::isfinite(rad[0].x) && rad[0].x >= 0.f && rad[0].y >= 0.f && rad[0].z >= 0.f);
            vec3 one = shade(out[0], -centre.direction, dev, 7, 0, 0);
            CHECK(std::memcmp(&one, &rad[0], sizeof one) == 0);
            if (out[0].triangle.is_emissive)
            {
                const vec3 r = scene.materials[out[0].triangle.mtl_name].radiance;
                CHECK(rad[0].x == r.x && rad[0].y == r.y && rad[0].z == r.z);
            }
        }

        // the fan-out over GPUs (main.cpp:79-81's role) with a replica of the scene on the same device and the library's
        // peer reduce: the frame of the single-GPU loop up to the order of two double partial sums
        {
            DeviceScene replica(dev, 0, true);
            std::vector<DeviceScene *> devs = {&dev, &replica};
            std::vector<double> m(n);
            renderImage(devs, 3, m.data(), 5, 0, TRT_RENDER_PEER_REDUCE);
            std::vector<double> s3(n);
            renderImage(dev, 3, s3.data(), 5);
            for (size_t i = 0; i < n; ++i)
                CHECK(std::fabs(m[i] - s3[i]) <= 1e-12 * std::fabs(s3[i]));
        }

        // checkpointed render: 5 samples in steps of 2, then "resumed" from the finished checkpoint (0 samples left): both
        // times the frame of the uninterrupted render, bit for bit; a checkpoint made with another seed does not apply
        {
            const std::string ckpt = std::string(argv[1]) + "/selftest.ckpt";
            std::remove(ckpt.c_str());
            std::vector<double> u(n), r1(n), r2(n);
            renderImage(dev, 5, u.data(), 9);
            // (the uninterrupted render deposits its one batch sample by sample: the same sequence of additions)
            CHECK(renderImageCheckpointed(dev, 5, r1.data(), ckpt, 2, 9) == 5);
            CHECK(std::memcmp(u.data(), r1.data(), n * sizeof(double)) == 0);
            CHECK(renderImageCheckpointed(dev, 5, r2.data(), ckpt, 2, 9) == 0);
            CHECK(std::memcmp(u.data(), r2.data(), n * sizeof(double)) == 0);
            CHECK(renderImageCheckpointed(dev, 5, r2.data(), ckpt, 2, 10) == 5); // another seed: the checkpoint does not apply
            std::remove(ckpt.c_str());
        }
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
