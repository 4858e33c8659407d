#!/usr/bin/env python3
"""BASELINE config 5 at 1 and N GPUs of one box, in ONE process (the library's own multi-GPU path): the procedurally
tessellated stress mesh is built once on the host, uploaded to GPU 0 and replicated device to device (trt_scene_replicate);
every GPU then traces its own config-2 ray batch (rays shard by index, no collective) and the 3840x2160 frame is
rendered with the samples sharded over the GPUs and one reduce (trt_render_multi).  Prints one JSON line.
usage: bench_stress_multi.py [--nq 2236] [--gpus 8] [--rays 4194304] [--spp 8]"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import tinyraytracing_b200 as trt  # noqa: E402
from tinyraytracing_b200 import workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nq", type=int, default=2236)
ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
ap.add_argument("--rays", type=int, default=4 << 20)
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
a = ap.parse_args()
t0 = time.time()
m = workloads.stress_mesh(a.nq)
cam = m["camera"]
host = trt.HostScene.from_arrays(m["v9"], m["mtl"], m["materials"], m["lights"], cam["eye"], cam["lookat"], cam["up"],
                                 cam["fovy"], a.width, a.height, vn9=m["vn9"])
t_host = time.time() - t0
t0 = time.time()
devs = [trt.DeviceScene(host, 0)]
t_create = time.time() - t0
t0 = time.time()
devs += [devs[0].replicate(i) for i in range(1, a.gpus)]
t_repl = time.time() - t0
st = devs[0].stats()


def tracer(rays):
    ids, t = devs[0].trace_closest(rays)
    hp, pn = devs[0].hit_attributes(rays, ids, t)
    return ids, hp, pn


n = a.rays
bufs = []
for i in range(a.gpus):
    rays = workloads.fixed_ray_batch(n, host.camera(), host.root_box(), tracer, seed=0x5EED0001 + i)
    with torch.cuda.device(i):
        bufs.append((torch.from_numpy(rays).cuda(i), torch.empty(n, dtype=torch.int32, device="cuda:%d" % i),
                     torch.empty(n, dtype=torch.float32, device="cuda:%d" % i), torch.cuda.Stream(device=i)))


def closest(gpus, launches=5):
    """launches x n rays on each of `gpus` GPUs at once; device time = max over GPUs (CUDA events per GPU)."""
    ms = [0.0] * gpus

    def run(i):
        d_r, d_i, d_t, s = bufs[i]
        with torch.cuda.device(i):
            for _ in range(2):
                devs[i].trace_closest_async(d_r.data_ptr(), n, d_i.data_ptr(), d_t.data_ptr(), 0, s.cuda_stream)
            s.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(launches):
                devs[i].trace_closest_async(d_r.data_ptr(), n, d_i.data_ptr(), d_t.data_ptr(), 0, s.cuda_stream)
            e1.record(s)
            s.synchronize()
            ms[i] = e0.elapsed_time(e1) / launches

    th = [threading.Thread(target=run, args=(i,)) for i in range(gpus)]
    [t.start() for t in th]
    [t.join() for t in th]
    return max(ms), gpus * n / (max(ms) * 1e-3) / 1e6


def render(gpus, flags=0):
    sub = devs[:gpus]
    trt.render_multi(sub, a.spp, seed=2, flags=flags)
    for d in sub:
        d.reset_stats()
    img = trt.render_multi(sub, a.spp, seed=1, flags=flags)
    ms = sub[0].stats()["last_render_ms"]
    rays = sum(d.stats()["rays_closest"] + d.stats()["rays_shadow"] for d in sub)
    g = np.clip(np.power(img, np.float64(np.float32(1.0) / np.float32(2.2))) * 255, 0, 255).astype(np.uint8)
    import hashlib

    return {"ms": ms, "spp_per_s": a.spp / (ms * 1e-3), "mrays_per_s": rays / (ms * 1e-3) / 1e6,
            "frame_rgb8_sha256": hashlib.sha256(g.tobytes()).hexdigest()[:16], "image_mean": float(img.mean())}


out = {"what": "BASELINE config 5: tessellated stress mesh, %d triangles, reference BVH depth %d" % (host.n_tris, st["ref_depth"]),
       "host_build_s": round(host.build_seconds, 2), "host_total_s": round(t_host, 1), "device_create_s": round(t_create, 1),
       "replicate_s_total": round(t_repl, 2), "gpus": a.gpus, "rays_per_gpu": n, "frame": "%dx%d, %d spp" % (a.width, a.height, a.spp)}
for g in sorted({1, a.gpus}):
    ms, mr = closest(g)
    out["closest_hit_%dgpu" % g] = {"ms_per_launch": ms, "mrays_per_s": mr}
    out["render_%dgpu_nccl" % g] = render(g, 0)
    if g > 1:
        out["render_%dgpu_peer" % g] = render(g, trt.RENDER_PEER_REDUCE)
print(json.dumps(out))
for d in devs:
    d.close()
