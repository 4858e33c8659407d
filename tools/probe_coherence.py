#!/usr/bin/env python3
"""How much does ray coherence buy? Camera rays in raster order (neighbouring lanes = neighbouring pixels) vs the
same rays shuffled, device-resident, default traversal."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import tinyraytracing_b200 as trt
from tinyraytracing_b200 import scenes

def rate(dev, rays, flags=0):
    n = len(rays)
    d_rays = torch.from_numpy(np.ascontiguousarray(rays)).cuda()
    d_id = torch.empty(n, dtype=torch.int32, device="cuda"); d_t = torch.empty(n, dtype=torch.float32, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), flags, sp)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), flags, sp)
    e1.record(); torch.cuda.synchronize()
    return 3 * n / e0.elapsed_time(e1) / 1e3

for name in ("back", "veach-mis", "staircase"):
    with tempfile.TemporaryDirectory() as tmp:
        f = scenes.materialize(name, tmp, width=1280, height=720)
        host = trt.HostScene.load(f["xml"], f["obj"], f["mtl"], f["basedir"])
        dev = trt.DeviceScene(host, 0)
        cam = host.camera(); W, H = cam["width"], cam["height"]
        rng = np.random.default_rng(0)
        jj, ii = np.meshgrid(np.arange(W), np.arange(H))
        reps = 4
        i = np.tile(ii.reshape(-1), reps); j = np.tile(jj.reshape(-1), reps); n = len(i)
        x = j / (W - 1.0) + (rng.random(n) - 0.5) / W; y = (H - i) / (H - 1.0) + (rng.random(n) - 0.5) / H
        d = cam["llc"][None] + x.astype(np.float32)[:, None] * cam["horizontal"][None] + y.astype(np.float32)[:, None] * cam["vertical"][None] - cam["eye"][None]
        d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        rays = np.concatenate([np.broadcast_to(cam["eye"][None], d.shape), d], 1).astype(np.float32)
        # tile order: 8x4 pixel tiles per warp
        ti = (ii // 4) * (W // 8) + (jj // 8); order_t = np.argsort((ti * 32 + (ii % 4) * 8 + (jj % 8)).reshape(-1), kind="stable")
        tiled = np.concatenate([rays[k * W * H:(k + 1) * W * H][order_t] for k in range(reps)])
        shuf = rays[rng.permutation(n)]
        ids, t = dev.trace_closest(rays[: W * H]); hp, pn = dev.hit_attributes(rays[: W * H], ids, t)
        print("%-10s primary raster %.0f  tiled8x4 %.0f  shuffled %.0f Mrays/s | plain kernel raster %.0f shuffled %.0f" % (
            name, rate(dev, rays), rate(dev, tiled), rate(dev, shuf), rate(dev, rays, trt.TRACE_PLAIN), rate(dev, shuf, trt.TRACE_PLAIN)))
        dev.close()
