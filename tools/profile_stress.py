#!/usr/bin/env python3
"""BASELINE config 5 driver for timing / ncu: the procedurally tessellated stress mesh (workloads.stress_mesh),
N config-2 rays, a few launches of the default closest-hit kernel on device-resident buffers.
usage: profile_stress.py <nq> [n_rays] [launches]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import tinyraytracing_b200 as trt  # noqa: E402
from tinyraytracing_b200 import workloads  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 708
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4 << 20
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 3
settings = ["auto"]
m = workloads.stress_mesh(nq)
cam = m["camera"]
t0 = time.time()
host = trt.HostScene.from_arrays(m["v9"], m["mtl"], m["materials"], m["lights"], cam["eye"], cam["lookat"], cam["up"],
                                 cam["fovy"], 3840, 2160, vn9=m["vn9"])
print("nq %d: %d triangles, host build %.1f s" % (nq, host.n_tris, time.time() - t0))
rays = None
for setting in settings:
    t0 = time.time()
    dev = trt.DeviceScene(host, 0)
    t_dev = time.time() - t0
    if rays is None:
        def tracer(r):
            ids, t = dev.trace_closest(r)
            hp, pn = dev.hit_attributes(r, ids, t)
            return ids, hp, pn

        rays = workloads.fixed_ray_batch(n, host.camera(), host.root_box(), tracer)
        d_rays = torch.from_numpy(rays).cuda()
        d_id = torch.empty(n, dtype=torch.int32, device="cuda")
        d_t = torch.empty(n, dtype=torch.float32, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), 0, sp)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(launches):
        dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), 0, sp)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / launches
    print("  device create %.1f s, %.3f ms/launch, %.1f Mrays/s, checksum %d" % (
        t_dev, ms, n / ms / 1e3, int(d_id.to(torch.int64).sum().item())))
    if setting == settings[-1]:
        print("  traversal work per ray:", dev.trace_counters(rays[:: max(1, n >> 20)]))
    dev.close()
