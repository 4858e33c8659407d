#!/usr/bin/env python3
"""BASELINE config 5: procedurally tessellated stress mesh (displaced UV sphere in a diffuse box with a ceiling
light), Mrays/s vs triangle count / BVH depth.  For each size: host build time (reference topology),
closest-hit Mrays/s of the config-2 ray population (device-resident), agreement of the default traversal with the
exhaustive reference walk on a sample, and a 1-spp render at the requested resolution.
usage: bench_stress.py [--nq 71 224 708 2236] [--rays 4194304] [--width 3840 --height 2160] [--out profiles/x.json]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import tinyraytracing_b200 as trt  # noqa: E402
from tinyraytracing_b200 import workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nq", type=int, nargs="+", default=[71, 224, 708, 2236])
ap.add_argument("--rays", type=int, default=4 << 20)
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--spp", type=int, default=1)
ap.add_argument("--out", default=None)
a = ap.parse_args()
rows = []
for nq in a.nq:
    t0 = time.time()
    m = workloads.stress_mesh(nq)
    cam = m["camera"]
    host = trt.HostScene.from_arrays(m["v9"], m["mtl"], m["materials"], m["lights"], cam["eye"], cam["lookat"], cam["up"],
                                     cam["fovy"], a.width, a.height, vn9=m["vn9"])
    t_host = time.time() - t0
    t0 = time.time()
    dev = trt.DeviceScene(host, 0)
    t_dev = time.time() - t0
    st = dev.stats()

    def tracer(rays):
        ids, t = dev.trace_closest(rays)
        hp, pn = dev.hit_attributes(rays, ids, t)
        return ids, hp, pn

    n = a.rays
    rays = workloads.fixed_ray_batch(n, host.camera(), host.root_box(), tracer)
    d_rays = torch.from_numpy(rays).cuda()
    d_id = torch.empty(n, dtype=torch.int32, device="cuda")
    d_t = torch.empty(n, dtype=torch.float32, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    res = {}
    for label, flags in (("default", 0), ("reftopo", trt.TRACE_REFTOPO)):
        for _ in range(2):
            dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), flags, sp)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), flags, sp)
        e1.record()
        torch.cuda.synchronize()
        res[label] = 3 * n / (e0.elapsed_time(e1) * 1e-3) / 1e6
    ids = d_id.cpu().numpy()
    sel = np.random.default_rng(0).choice(n, min(n, 1 << 18), replace=False)
    ids_d, t_d = dev.trace_closest(rays[sel])
    ids_x, t_x = dev.trace_closest(rays[sel], trt.TRACE_EXHAUSTIVE)
    same = bool(np.array_equal(ids_d, ids_x) and np.array_equal(t_d.view(np.uint32), t_x.view(np.uint32)))
    work = dev.trace_counters(rays[:: max(1, n >> 19)])
    dev.render(1, seed=1)
    dev.reset_stats()
    dev.render(a.spp, seed=1)
    rs = dev.stats()
    row = dict(nq=nq, tris=int(host.n_tris), ref_nodes=int(host.desc.n_nodes), ref_depth=int(st["ref_depth"]),
               accel_nodes=int(st["accel_nodes"]), leaves=int(st["accel_leaves"]), host_load_build_s=round(t_host, 2),
               host_build_s=round(host.build_seconds, 2), device_create_s=round(t_dev, 2), closest_mrays_default=res["default"],
               closest_mrays_reftopo=res["reftopo"], default_equals_exhaustive=same, hit_fraction=float((ids >= 0).mean()),
               work_per_ray=work, render=dict(width=a.width, height=a.height, spp=a.spp, ms=rs["last_render_ms"],
                                              mrays=(rs["rays_closest"] + rs["rays_shadow"]) / rs["last_render_ms"] / 1e3,
                                              spp_per_s=a.spp / rs["last_render_ms"] * 1e3))
    print(json.dumps(row), flush=True)
    rows.append(row)
    dev.close()
    host.close()
if a.out:
    with open(a.out, "w") as f:
        json.dump(rows, f, indent=1)
