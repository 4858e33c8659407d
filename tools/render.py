#!/usr/bin/env python3
"""Render a scene with the B200 path tracer and write the reference's gamma-2.2 8-bit PNG.

  python tools/render.py --scene veach-mis --spp 256 --out veach.png            # packed cg22 scene (scenes/*.npz)
  python tools/render.py --files basedir scene.mtl scene.xml scene.obj --spp 64  # the reference's four inputs
  torchrun --nproc-per-node 8 tools/render.py --scene staircase --width 1920 --height 1080 --spp 1024

Under torchrun every rank renders its sample range on its own GPU and rank 0 writes the image (one NCCL reduce)."""
import argparse
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import tinyraytracing_b200 as trt  # noqa: E402
from tinyraytracing_b200 import scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", choices=scenes.NAMES)
    ap.add_argument("--files", nargs=4, metavar=("BASEDIR", "MTL", "XML", "OBJ"))
    ap.add_argument("--width", type=int)
    ap.add_argument("--height", type=int)
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--max-depth", type=int, default=0, help="0 = reference behaviour (Russian roulette only)")
    ap.add_argument("--out", default="image.png")
    ap.add_argument("--pfm", default=None, help="also write the linear (pre-gamma) image as a 32-bit float PFM file")
    a = ap.parse_args()
    if bool(a.scene) == bool(a.files):
        ap.error("give exactly one of --scene / --files")
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    with tempfile.TemporaryDirectory() as tmp:
        if a.scene:
            f = scenes.materialize(a.scene, tmp, width=a.width, height=a.height)
            base, mtl, xml, obj = f["basedir"], f["mtl"], f["xml"], f["obj"]
        else:
            base, mtl, xml, obj = a.files
        host = trt.HostScene.load(xml, obj, mtl, base)
        dev = trt.DeviceScene(host, local)
        t0 = time.perf_counter()
        if world > 1:
            import torch
            import torch.distributed as dist

            from tinyraytracing_b200.distributed import render_on_gpus

            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            img = render_on_gpus(dev, a.spp, seed=a.seed, max_depth=a.max_depth)
            dist.destroy_process_group()
        else:
            img = dev.render(a.spp, seed=a.seed, max_depth=a.max_depth)
        dt = time.perf_counter() - t0
        if rank == 0:
            g = np.power(img, np.float64(np.float32(1.0) / np.float32(2.2))) * 255  # imshow, main.cpp:30-38
            rgb = np.ascontiguousarray(np.minimum(np.maximum(g, 0.0), 255.0).astype(np.uint8))
            trt.load_library().trt_write_png(a.out.encode(), rgb.shape[1], rgb.shape[0], rgb.ctypes.data, 0)
            if a.pfm:  # the linear buffer, for numerical comparison of renders
                lin = np.ascontiguousarray(img, np.float64)
                trt.load_library().trt_write_pfm(a.pfm.encode(), lin.shape[1], lin.shape[0], lin.ctypes.data)
            st = dev.stats()
            print("%dx%d, %d spp on %d GPU(s): %.2f s, wrote %s (this rank: %d closest + %d shadow rays)" % (
                rgb.shape[1], rgb.shape[0], a.spp, world, dt, a.out, st["rays_closest"], st["rays_shadow"]))
        dev.close()


if __name__ == "__main__":
    main()
