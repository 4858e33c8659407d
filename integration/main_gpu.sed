# Applied by `make -C oracle bridge` to a COPY of the reference's RayTracingOnCPU/main.cpp (sed -f): the program stays the
# reference's own — stdin protocol, loaders, buildBVH, imshow / svpng — except that the sample loop main.cpp:79-113
# (omp_set_num_threads ... closing brace of the `for k` loop) is replaced by one call into include/trt.h.
# A line-number script rather than a unified diff, so that no reference source is quoted in this repository;
# oracle/Makefile checks that lines 79 and 114 still are what this script assumes.
4a\
#include "trt_bridge.h"
79,113c\
    // B200 path (integration/trt_bridge.cpp): the sample loop runs on the GPU through include/trt.h\
    TrtBridge *gpu = trt_bridge_create(scene, root, getenv("TRT_DEVICE") ? atoi(getenv("TRT_DEVICE")) : 0);\
    trt_render_params p{};\
    p.spp = SAMPLE, p.sample_begin = 0, p.sample_end = SAMPLE, p.max_depth = 0;\
    p.seed = getenv("TRT_SEED") ? strtoull(getenv("TRT_SEED"), NULL, 0) : (uint64_t)time(NULL);\
    if (trt_render(gpu->scene, &p, image) != TRT_OK)\
    {\
        printf("%s\\n", trt_last_error());\
        exit(1);\
    }
