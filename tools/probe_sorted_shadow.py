#!/usr/bin/env python3
"""Would reordering paths by hit-point cell pay? Depth-1 shadow rays (from the hit points of cosine bounce rays off the
primary hits, aimed at light sample points) traced in parent-pixel order vs sorted by a Morton code of the origin."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import tinyraytracing_b200 as trt
from tinyraytracing_b200 import scenes, workloads

def rate(dev, rays):
    n = len(rays)
    d_rays = torch.from_numpy(np.ascontiguousarray(rays)).cuda()
    d_id = torch.empty(n, dtype=torch.int32, device="cuda"); d_t = torch.empty(n, dtype=torch.float32, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), 0, sp)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), 0, sp)
    e1.record(); torch.cuda.synchronize()
    return 3 * n / e0.elapsed_time(e1) / 1e3

def morton(p, lo, hi, bits):
    q = np.clip(((p - lo) / (hi - lo) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    code = np.zeros(len(p), np.int64)
    for b in range(bits):
        for a in range(3):
            code |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return code

for name in ("veach-mis", "staircase"):
    with tempfile.TemporaryDirectory() as tmp:
        f = scenes.materialize(name, tmp, width=1280, height=720)
        host = trt.HostScene.load(f["xml"], f["obj"], f["mtl"], f["basedir"])
        dev = trt.DeviceScene(host, 0)
        cam = host.camera(); W, H = cam["width"], cam["height"]
        rng = np.random.default_rng(0)
        jj, ii = np.meshgrid(np.arange(W), np.arange(H)); i = np.tile(ii.reshape(-1), 2); j = np.tile(jj.reshape(-1), 2); n = len(i)
        x = j / (W - 1.0) + (rng.random(n) - 0.5) / W; y = (H - i) / (H - 1.0) + (rng.random(n) - 0.5) / H
        d = cam["llc"][None] + x.astype(np.float32)[:, None] * cam["horizontal"][None] + y.astype(np.float32)[:, None] * cam["vertical"][None] - cam["eye"][None]
        d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        prim = np.concatenate([np.broadcast_to(cam["eye"][None], d.shape), d], 1).astype(np.float32)
        ids, t = dev.trace_closest(prim); hp, pn = dev.hit_attributes(prim, ids, t)
        ok = (ids >= 0) & np.isfinite(pn).all(1)
        nrm = pn[ok].astype(np.float64); flip = (nrm * prim[ok, 3:]).sum(1) > 0; nrm[flip] *= -1
        bnc = workloads.bounce_rays(hp[ok], nrm, rng)
        ids1, t1 = dev.trace_closest(bnc); hp1, pn1 = dev.hit_attributes(bnc, ids1, t1)
        ok1 = ids1 >= 0
        P = hp1[ok1]
        ls, lv, lvn, cum = host.lights()
        lo, hi = host.root_box()
        out = []
        for li, l in enumerate(ls):
            if l["n_tris"] == 0: continue
            k = l["first_tri"] + rng.integers(0, min(l["n_tris"], 24), len(P))
            b = rng.dirichlet((1, 1, 1), len(P))[:, :, None]
            q = (lv[k].reshape(-1, 3, 3) * b).sum(1)
            dd = q - P; dd /= np.linalg.norm(dd, axis=1, keepdims=True)
            rays = np.concatenate([P, dd], 1).astype(np.float32)
            r0 = rate(dev, rays)
            res = [r0]
            for bits in (5, 7, 10):
                order = np.argsort(morton(P, lo, hi, bits), kind="stable")
                res.append(rate(dev, rays[order]))
            out.append("light %d: parent order %.0f | sorted 15-bit %.0f  21-bit %.0f  30-bit %.0f" % (li, *res))
        rb = [rate(dev, bnc)]
        order = np.argsort(morton(hp[ok], lo, hi, 10), kind="stable"); rb.append(rate(dev, bnc[order]))
        print(name, "(%d depth-1 vertices) bounce rays: pixel order %.0f, origin-sorted %.0f Mrays/s" % (len(P), *rb))
        for o in out: print("   ", o)
        dev.close()
