#!/usr/bin/env python3
"""A/B of library variants (tools/build_variant.sh): for back / veach-mis / staircase, closest-hit Grays/s of a 4 Mi
config-2 batch with a hash of ids + distances, render time (best of 3) with a hash of the image, and device time per
kernel class (TRT_RENDER_PROFILE).  One JSON line per library; equal hashes = bit-identical results.
usage: ab_variants.py <lib.so>      (environment variables such as TRT_REINSERT are honoured at scene creation)"""
import hashlib, json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import tinyraytracing_b200 as trt
from tinyraytracing_b200 import api, scenes, workloads
lib = os.path.abspath(sys.argv[1])
api.library_path = lambda: lib
out = {"lib": os.path.basename(lib)}
dn = os.open(os.devnull, os.O_WRONLY); sv = os.dup(1)
def load(name, w, h, tmp):
    f = scenes.materialize(name, os.path.join(tmp, name + "%d" % w), width=w, height=h)
    os.dup2(dn, 1)
    try:
        return trt.HostScene.load(f["xml"], f["obj"], f["mtl"], f["basedir"])
    finally:
        os.dup2(sv, 1)
with tempfile.TemporaryDirectory() as tmp:
    for name, (w, h, spp) in {"back": (512, 512, 16), "veach-mis": (1280, 720, 32), "staircase": (1280, 720, 16)}.items():
        host = load(name, w, h, tmp)
        dev = trt.DeviceScene(host, 0)
        def tracer(rays):
            ids, t = dev.trace_closest(rays)
            hp, pn = dev.hit_attributes(rays, ids, t)
            return ids, hp, pn
        n = 4 << 20
        rays = workloads.fixed_ray_batch(n, host.camera(), host.root_box(), tracer)
        d_rays = torch.from_numpy(rays).cuda()
        d_id = torch.empty(n, dtype=torch.int32, device="cuda"); d_t = torch.empty(n, dtype=torch.float32, device="cuda")
        sp = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), 0, sp)
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                dev.trace_closest_async(d_rays.data_ptr(), n, d_id.data_ptr(), d_t.data_ptr(), 0, sp)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 5)
        rec = {"closest_grays": n / best / 1e6, "ids_sha": hashlib.sha256(d_id.cpu().numpy().tobytes() + d_t.cpu().numpy().tobytes()).hexdigest()[:12]}
        frame = dev.pinned_image() if hasattr(dev, "pinned_image") else None
        kw = {"out": frame} if frame is not None else {}
        dev.render(spp, seed=1, **kw)
        ms = []
        for rep in range(3):
            dev.reset_stats()
            img = dev.render(spp, seed=1, **kw)
            ms.append(dev.stats()["last_render_ms"])
        rec["render_ms"] = min(ms); rec["spp_per_s"] = spp / min(ms) * 1e3
        rec["img_sha"] = hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest()[:12]
        rec["img_mean"] = float(img.mean())
        try:
            dev.render(spp, seed=1, flags=4, **kw)
            st = dev.stats()
            rec["prof"] = {k: round(st[k], 2) for k in ("ms_trace", "ms_shade", "ms_shadow", "ms_accumulate")}
        except Exception as e:
            rec["prof"] = str(e)
        out[name] = rec
        dev.close(); host.close()
print(json.dumps(out))
