"""tinyraytracing_b200 — B200-native hot path of TinyRayTracing (BVH traversal + ray/triangle intersection
inside the path-tracing bounce loop) behind a C ABI (include/trt.h, libtrt_b200.so).

Python here is harness only (ctypes bindings for tests / bench / multi-GPU plumbing); the product is the
shared library and the C++ host mirror under csrc/.
"""
from .api import (HostScene, DeviceScene, TrtError, RenderParams, load_library, library_path,  # noqa: F401
                  TRACE_EXHAUSTIVE, TRACE_REFTOPO, TRACE_PLAIN, TRACE_POOLED, TRACE_PERSISTENT, TRACE_DEVICE_PTRS, INF,
                  RENDER_REFTOPO, RENDER_PLAIN, RENDER_PROFILE, RENDER_PEER_REDUCE, render_multi, trace_closest_multi)
