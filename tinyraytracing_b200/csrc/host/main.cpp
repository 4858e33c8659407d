// Drop-in driver: the reference's program (main.cpp:44-119) with the sample loop replaced by the GPU path.
// Same stdin protocol (basedir, .mtl, .xml, .obj, SPP — main.cpp:46-55), same load order (:66-69), same
// buildBVH call (:76), same gamma-2.2 8-bit PNG `<basedir>/image<SPP>.png` (:19-42).
// Optional overrides that default to reference behaviour (environment): TRT_SEED, TRT_MAX_DEPTH, TRT_DEVICE;
// TRT_DEVICES=0,1,... renders on several GPUs of the box (samples sharded, one reduce: trt_render_multi;
// TRT_PEER_REDUCE=1 selects the library's peer-memory reduce instead of NCCL); TRT_CHECKPOINT=<file> with
// TRT_CHECKPOINT_EVERY=<spp> (default 64) writes accumulation checkpoints and resumes from one (single GPU).
#include "tinyrt.h"
#include "png_writer.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <memory>

using namespace trt;

static void imshow(double *SRC, const std::string &basedir, const std::string &index, int img_width, int img_height)
{
    std::vector<unsigned char> image((size_t)img_width * img_height * 3);
    for (size_t i = 0; i < image.size(); ++i)
    {
        // main.cpp:34-36: (unsigned char)clamp(pow(v, 1.0f / 2.2f) * 255, 0.0, 255.0), glm::clamp = min(max(x,lo),hi)
        double v = std::pow(SRC[i], (double)(1.0f / 2.2f)) * 255;
        v = (v < 0.0) ? 0.0 : v;
        v = (255.0 < v) ? 255.0 : v;
        image[i] = (unsigned char)v;
    }
    std::string name = basedir + "/image" + index + ".png";
    FILE *fp = std::fopen(name.c_str(), "wb");
    if (!fp || !writePNG(fp, img_width, img_height, image.data(), 0))
    {
        std::cerr << "cannot write " << name << std::endl;
        std::exit(1);
    }
    std::fclose(fp);
    std::cout << "\nIamge output to " << name << std::endl;
}

int main()
{
    int SAMPLE = 256;
    std::string basedir, mtl_path, xml_path, obj_path;
    std::printf("Please input base directory of the scene:\n");
    std::cin >> basedir;
    std::printf("Please input .mtl file path of the scene:\n");
    std::cin >> mtl_path;
    std::printf("Please input .xml file path of the scene:\n");
    std::cin >> xml_path;
    std::printf("Please input .obj file path of the scene:\n");
    std::cin >> obj_path;
    std::printf("Please input SPP:\n");
    std::cin >> SAMPLE;

    auto start = std::chrono::steady_clock::now();
    try
    {
        Scene scene;
        scene.readxml(xml_path); // load order cannot be changed (main.cpp:66)
        scene.readobj(obj_path);
        scene.readmtl(mtl_path, basedir);
        std::printf("image info:\nwidth: %d height: %d\n", scene.img_width, scene.img_height);
        scene.camera.Print();

        std::vector<double> image((size_t)scene.img_width * scene.img_height * 3, 0.0);
        BVHNode *root = buildBVH(scene.triangles, 0, (int)scene.triangles.size() - 1, 8);
        std::printf("Build BVH down.\n");

        const char *e;
        uint64_t seed = (e = std::getenv("TRT_SEED")) ? std::strtoull(e, nullptr, 0) : 0;
        int max_depth = (e = std::getenv("TRT_MAX_DEPTH")) ? std::atoi(e) : 0;
        int device = (e = std::getenv("TRT_DEVICE")) ? std::atoi(e) : 0;
        std::vector<int> devices;
        if ((e = std::getenv("TRT_DEVICES")))
            for (const char *q = e; *q;)
            {
                char *end;
                const long v = std::strtol(q, &end, 10);
                if (end == q)
                    break;
                devices.push_back((int)v);
                q = (*end == ',') ? end + 1 : end;
            }
        if (devices.empty())
            devices.push_back(device);
        std::vector<std::unique_ptr<DeviceScene>> devs;
        for (int d : devices) // the scene is replicated on every GPU: built and uploaded once, then copied device to device
            devs.emplace_back(devs.empty() ? new DeviceScene(scene, root, d, std::getenv("TRT_LAYOUT_CACHE")) : new DeviceScene(*devs[0], d, true));
        DeviceScene &dev = *devs[0];
        const char *ckpt = std::getenv("TRT_CHECKPOINT");
        if (devs.size() > 1)
        {
            std::vector<DeviceScene *> ptrs;
            for (auto &d : devs)
                ptrs.push_back(d.get());
            const bool peer = (e = std::getenv("TRT_PEER_REDUCE")) && std::atoi(e) != 0;
            renderImage(ptrs, SAMPLE, image.data(), seed, max_depth, peer ? TRT_RENDER_PEER_REDUCE : 0u);
            std::printf("rendered on %zu GPUs (%s reduce)\n", devs.size(), peer ? "peer-memory" : "NCCL");
        }
        else if (ckpt && *ckpt)
        {
            const int every = (e = std::getenv("TRT_CHECKPOINT_EVERY")) ? std::atoi(e) : 64;
            const int fresh = renderImageCheckpointed(dev, SAMPLE, image.data(), ckpt, every, seed, max_depth);
            std::printf("checkpoint %s: %d of %d samples rendered by this run\n", ckpt, fresh, SAMPLE);
        }
        else
            renderImage(dev, SAMPLE, image.data(), seed, max_depth);

        trt_stats st;
        trt_get_stats(dev.handle(), &st);
        for (size_t i = 1; i < devs.size(); ++i) // ray counts are per scene: sum the replicas'
        {
            trt_stats o;
            trt_get_stats(devs[i]->handle(), &o);
            st.rays_closest += o.rays_closest, st.rays_shadow += o.rays_shadow;
        }
        std::printf("rays: %llu closest + %llu shadow, %.3f ms on device (%.1f Mrays/s)\n",
                    (unsigned long long)st.rays_closest, (unsigned long long)st.rays_shadow, st.last_render_ms,
                    (st.rays_closest + st.rays_shadow) / (st.last_render_ms * 1e3));
        imshow(image.data(), basedir, std::to_string(SAMPLE), scene.img_width, scene.img_height);
        freeBVH(root);
    }
    catch (const std::exception &e)
    {
        std::cout << e.what() << std::endl;
        return 1;
    }
    std::cerr << "\nDone.\n";
    std::cout << std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count() << std::endl;
    return 0;
}
