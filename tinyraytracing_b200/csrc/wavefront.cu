// placeholder until the wavefront integrator lands (next commit)
#include "scene_impl.h"
namespace trt
{
int renderAccumulate(trt_scene *, const trt_render_params &, double *, cudaStream_t)
{
    setLastError("render path not built yet");
    return TRT_ERR_INVALID;
}
int resolveImage(trt_scene *, const double *, int, double *, uint8_t *, cudaStream_t)
{
    setLastError("render path not built yet");
    return TRT_ERR_INVALID;
}
void destroyWavefront(trt_scene *) {}
} // namespace trt
