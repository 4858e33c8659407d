// Drop-in driver: the reference's program (main.cpp:44-119) with the sample loop replaced by the GPU path.
// Same stdin protocol (basedir, .mtl, .xml, .obj, SPP — main.cpp:46-55), same load order (:66-69), same
// buildBVH call (:76), same gamma-2.2 8-bit PNG `<basedir>/image<SPP>.png` (:19-42).
// Optional overrides that default to reference behaviour: TRT_SEED, TRT_MAX_DEPTH, TRT_DEVICE (environment).
#include "tinyrt.h"
#include "png_writer.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <iostream>

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
        DeviceScene dev(scene, root, device);
        renderImage(dev, SAMPLE, image.data(), seed, max_depth);

        trt_stats st;
        trt_get_stats(dev.handle(), &st);
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
