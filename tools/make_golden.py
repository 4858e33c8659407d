#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the UNMODIFIED reference (tier A, oracle/_ref/libref.so) in the build
container (needs /root/reference to have built oracle/_ref).  The reference ships no golden vectors or tests
(SURVEY §4), so these outputs of the reference itself are what pins the oracle and the GPU path:

  <scene>_closest.npz  4096 rays of the config-2 population -> the reference traverseBVH's distance bits,
                       canonical triangle index, hit point, shading normal; post-build order (as OBJ face
                       ordinals), node boxes / links, camera vectors, per-triangle face normals
  <scene>_render.npz   two independent 64-spp renders by the reference's own shade() at the reduced test
                       resolution (linear float32) — statistical image reference and its noise floor

usage: python tools/make_golden.py [--render-threads 1]
"""
import argparse
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refbridge  # noqa: E402
from conftest import SMALL_RES  # noqa: E402
from tinyraytracing_b200 import scenes, workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--render-threads", type=int, default=1)
ap.add_argument("--spp", type=int, default=64)
ap.add_argument("--skip-render", action="store_true")
ap.add_argument("--scene", default=None)
a = ap.parse_args()
out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)
if a.scene is None:
    # one fresh process per scene: shade()'s `static uniform_real_distribution u1(0, total_area)`
    # (pathTracing.cpp:38) is initialised by the first light the PROCESS ever samples
    import subprocess

    for name in scenes.NAMES:
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--scene", name, "--spp", str(a.spp),
                               "--render-threads", str(a.render_threads)] + (["--skip-render"] if a.skip_render else []))
    sys.exit(0)
for name in [a.scene]:
    with tempfile.TemporaryDirectory() as tmp:
        w, h = SMALL_RES[name]
        f = scenes.materialize(name, tmp, width=w, height=h)
        ref = refbridge.RefScene(f["xml"], f["obj"], f["mtl"], f["basedir"])
        pre = refbridge.RefScene(f["xml"], f["obj"], f["mtl"], f["basedir"], build=False)
        tris, pre_tris = ref.triangles(), pre.triangles()
        # post-build index -> OBJ face ordinal, by content (duplicates resolved to the first unused ordinal)
        key = lambda t, nm, i: t["v"][i].tobytes() + t["vn"][i].tobytes() + t["vt"][i].tobytes() + nm[t["mtl"][i]].encode()
        slots = {}
        for i in range(pre.n):
            slots.setdefault(key(pre_tris, pre.material_names(), i), []).append(i)
        face = np.array([slots[key(tris, ref.material_names(), i)].pop(0) for i in range(ref.n)], np.int32)
        boxes, links = ref.bvh_flatten()
        cam12 = ref.camera()
        cam = dict(eye=cam12[0:3], llc=cam12[3:6], horizontal=cam12[6:9], vertical=cam12[9:12], width=w, height=h)

        def tracer(rays):
            t, ids, pn, hp = ref.trace(rays, want_pn=True)
            return ids, hp, pn

        rays = workloads.fixed_ray_batch(4096, cam, (boxes[0, :3], boxes[0, 3:]), tracer, seed=20221018)
        t, ids, pn, hp = ref.trace(rays, want_pn=True)
        np.savez_compressed(os.path.join(out, name + "_closest.npz"), rays=rays, t=t, id=ids, pn=pn, hitpoint=hp, face=face,
                            canon=tris["canon"], normal=tris["normal"], node_box=boxes, node_link=links, camera=cam12,
                            bvh_stats=np.array(ref.bvh_stats(), np.int32))
        print(name, "closest: hits", int((ids >= 0).sum()), "of", len(rays), "bvh", ref.bvh_stats())
        if not a.skip_render:
            r1 = ref.render(a.spp, threads=a.render_threads, seed=1).astype(np.float32)
            r2 = ref.render(a.spp, threads=a.render_threads, seed=2).astype(np.float32)
            np.savez_compressed(os.path.join(out, name + "_render.npz"), run1=r1, run2=r2, spp=np.int32(a.spp))
            print(name, "render mean", r1.mean(), r2.mean(), "rmse between runs", float(np.sqrt(((r1 - r2) ** 2).mean())))
