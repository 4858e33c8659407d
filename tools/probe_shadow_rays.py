import os, sys, tempfile
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import tinyraytracing_b200 as trt
from tinyraytracing_b200 import scenes, workloads
for name in ("veach-mis", "staircase"):
    with tempfile.TemporaryDirectory() as tmp:
        f = scenes.materialize(name, tmp)
        host = trt.HostScene.load(f["xml"], f["obj"], f["mtl"], f["basedir"])
        dev = trt.DeviceScene(host, 0)
        rng = np.random.default_rng(1)
        n = 1 << 20
        cam = workloads.camera_rays(host.camera(), n, rng)
        ids, t = dev.trace_closest(cam)
        hp, pn = dev.hit_attributes(cam, ids, t)
        ok = ids >= 0
        hp = hp[ok]
        ls, lv, lvn, cum = host.lights()
        print(name, "camera rays:", dev.trace_counters(cam))
        for li, l in enumerate(ls):
            k = l["first_tri"] + rng.integers(0, max(1, min(l["n_tris"], 24)), len(hp))
            b = rng.random((len(hp), 3)); b /= b.sum(1, keepdims=True)
            tri = lv[k].reshape(-1, 3, 3)
            q = (tri * b[:, :, None]).sum(1)
            d = q - hp; r = np.linalg.norm(d, axis=1, keepdims=True); d /= r
            rays = np.concatenate([hp, d], 1).astype(np.float32)
            ids2, t2 = dev.trace_closest(rays)
            c = dev.trace_counters(rays)
            d_rays = torch.from_numpy(rays).cuda(); m=len(rays)
            d_id = torch.empty(m, dtype=torch.int32, device="cuda"); d_t = torch.empty(m, dtype=torch.float32, device="cuda")
            sp = torch.cuda.current_stream().cuda_stream
            dev.trace_closest_async(d_rays.data_ptr(), m, d_id.data_ptr(), d_t.data_ptr(), 0, sp)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3): dev.trace_closest_async(d_rays.data_ptr(), m, d_id.data_ptr(), d_t.data_ptr(), 0, sp)
            e1.record(); torch.cuda.synchronize()
            print("  light %d (%d tris): %s  mean r %.2f mean t %.2f hitfrac %.2f  %.0f Mrays/s" % (li, l["n_tris"], {k_: round(v, 1) for k_, v in c.items()}, r.mean(), t2[ids2>=0].mean(), (ids2>=0).mean(), 3*m/e0.elapsed_time(e1)/1e3))
        dev.close()
