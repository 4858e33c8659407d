#!/usr/bin/env python3
"""Counts A (box tests) and T (triangle tests) per ray of the ordered, distance-pruned binary walk of the
reference topology (B = 48 + 32 A + 48 T algorithmic bytes per ray) on the BASELINE config-2 ray
population, per ray class, with the oracle (tier B).  Writes profiles/algorithmic_work.json, which bench.py
reads for roofline.achieved.   usage: python tools/measure_algorithmic_work.py [n_rays]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oraclelib  # noqa: E402
from tinyraytracing_b200 import workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
out = {}
for name, res in (("back", (512, 512)), ("veach-mis", (1280, 720)), ("staircase", (1280, 720))):
    orc = oraclelib.OracleScene(oraclelib.parsed_scene(name, *res))
    cam12 = orc.camera()
    cam = dict(eye=cam12[0:3], llc=cam12[3:6], horizontal=cam12[6:9], vertical=cam12[9:12], width=res[0], height=res[1])
    boxes, _ = orc.nodes()

    def tracer(rays):
        ids, t, pn, hp = orc.trace(rays, want_pn=True)
        return ids, hp, pn

    rays = workloads.fixed_ray_batch(n, cam, (boxes[0, :3], boxes[0, 3:]), tracer)
    A, T, ids = orc.trace_counts(rays, 1)
    Ax, Tx, idx = orc.trace_counts(rays, 0)
    assert np.array_equal(ids, idx)
    rec = dict(A=A, T=T, bytes_per_ray=48 + 32 * A + 48 * T, exhaustive_A=Ax, exhaustive_T=Tx, rays=n,
               hit_fraction=float((ids >= 0).mean()), classes={})
    q = n // 4
    for cls, sl in (("camera", slice(0, q)), ("uniform_box", slice(q, n - q)), ("bounce", slice(n - q, n))):
        a, t, _ = orc.trace_counts(rays[sl], 1)
        rec["classes"][cls] = dict(A=a, T=t, bytes_per_ray=48 + 32 * a + 48 * t)
    out[name] = rec
    print(name, json.dumps(rec))
with open(os.path.join(ROOT, "profiles", "algorithmic_work.json"), "w") as f:
    json.dump(out, f, indent=1)
