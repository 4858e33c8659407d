#!/usr/bin/env python3
"""Small driver for ncu launch lists of the wavefront renderer. usage: profile_render.py <scene> <w> <h> <spp> [flags]"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tinyraytracing_b200 as trt  # noqa: E402
from tinyraytracing_b200 import scenes  # noqa: E402

name, w, h, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0
batch = int(sys.argv[6]) if len(sys.argv) > 6 else 0
with tempfile.TemporaryDirectory() as tmp:
    f = scenes.materialize(name, tmp, width=w, height=h)
    host = trt.HostScene.load(f["xml"], f["obj"], f["mtl"], f["basedir"])
    dev = trt.DeviceScene(host, 0)
    dev.render(spp if batch else 1, seed=1, flags=flags, batch_paths=batch)
    dev.reset_stats()
    img = dev.render(spp, seed=1, flags=flags, batch_paths=batch)
    st = dev.stats()
    rays = st["rays_closest"] + st["rays_shadow"]
    print("%s %dx%d %d spp: %.2f ms, %.2f spp/s, %.1f Mrays/s (%d closest + %d shadow), %d launches, mean %.4f" % (
        name, w, h, spp, st["last_render_ms"], spp / st["last_render_ms"] * 1e3, rays / st["last_render_ms"] / 1e3,
        st["rays_closest"], st["rays_shadow"], st["kernel_launches"], img.mean()))
    dev.close()
