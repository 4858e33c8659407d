#!/usr/bin/env python3
"""Small driver for ncu captures: traces N rays of the config-2 population against one scene, a few launches
of the closest-hit kernel on device-resident buffers.  usage: profile_closest.py <scene> [n_rays] [flags] [launches]"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import tinyraytracing_b200 as trt  # noqa: E402
from tinyraytracing_b200 import scenes, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "staircase"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4 << 20
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 3
with tempfile.TemporaryDirectory() as tmp:
    f = scenes.materialize(name, tmp)
    host = trt.HostScene.load(f["xml"], f["obj"], f["mtl"], f["basedir"])
    dev = trt.DeviceScene(host, 0)

    def tracer(rays):
        ids, t = dev.trace_closest(rays, trt.TRACE_REFTOPO)
        hp, pn = dev.hit_attributes(rays, ids, t)
        return ids, hp, pn

    rays = workloads.fixed_ray_batch(n, host.camera(), host.root_box(), tracer)
    d_rays = torch.from_numpy(rays).cuda()
    d_id = torch.empty(n, dtype=torch.int32, device="cuda")
    d_t = torch.empty(n, dtype=torch.float32, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), flags, sp)
    e0.record()
    for _ in range(launches):
        dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), flags, sp)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / launches
    print("%s n=%d flags=%d: %.3f ms/launch, %.1f Mrays/s, hits %.3f" % (name, n, flags, ms, n / ms / 1e3,
                                                                         float((d_id >= 0).float().mean())))
    if flags == 0:
        print("  traversal work per ray:", dev.trace_counters(rays[:: max(1, n >> 20)]), {k: v for k, v in dev.stats().items() if k.startswith("accel")})
    dev.close()
