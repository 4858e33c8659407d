"""ctypes view of oracle/liboracle.so (tier B: deterministic CPU restatement) + scene-pack helpers.
Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg."""
import ctypes as C
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBORACLE = os.path.join(ROOT, "oracle", "liboracle.so")


class OrcMaterial(C.Structure):
    _fields_ = [("Kd", C.c_float * 3), ("Ks", C.c_float * 3), ("Tr", C.c_float * 3), ("Ns", C.c_float),
                ("Ni", C.c_float), ("texture", C.c_int32)]


class OrcTexture(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("bgr", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIBORACLE)
        vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
        L.orc_scene_create.restype = vp
        L.orc_scene_create.argtypes = [i32, vp, vp, vp, vp, i32, C.POINTER(OrcMaterial), i32, vp, vp, i32,
                                       C.POINTER(OrcTexture), vp, vp, vp, C.c_float, i32, i32, i32]
        L.orc_scene_destroy.argtypes = [vp]
        L.orc_num_nodes.argtypes = [vp]
        L.orc_get_order.argtypes = [vp, vp]
        L.orc_get_derived.argtypes = [vp, vp, vp, vp, vp]
        L.orc_get_nodes.argtypes = [vp, vp, vp]
        L.orc_get_camera.argtypes = [vp, vp]
        L.orc_bvh_stats.argtypes = [vp, vp, vp, vp]
        L.orc_trace.argtypes = [vp, vp, i64, vp, vp, vp, vp, i32]
        L.orc_trace_counts.argtypes = [vp, vp, i64, i32, vp, vp, vp, i32]
        L.orc_render.argtypes = [vp, i32, i32, i32, i32, u64, vp, vp, i32]
        L.orc_render_pixels.argtypes = [vp, i32, i32, i32, i32, u64, vp, i32, vp, i32]
        L.orc_shade_batch.argtypes = [vp, vp, vp, vp, i64, u64, i32, i32, vp, i32]
        L.orc_philox4x32_10.argtypes = [vp, vp, vp]
        L.orc_uniform.restype = C.c_double
        L.orc_uniform.argtypes = [u64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_primary_ray.argtypes = [vp, i32, i32, i32, u64, vp]
        L.orc_walk_layout.argtypes = [vp, vp, i64, vp, vp, vp, i32]
        L.orc_walk_layout.restype = None
        _lib = L
    return _lib


def available():
    return os.path.exists(LIBORACLE)


def parsed_scene(name, width=None, height=None):
    """Parsed (pre-build, OBJ-order) scene arrays straight from scenes/<name>.npz (tinyraytracing_b200.scenes.parsed)."""
    from tinyraytracing_b200 import scenes

    return scenes.parsed(name, width, height)


class OracleScene:
    def __init__(self, ps, leaf_num=8):
        """ps: dict as returned by parsed_scene() (or hand-made with the same keys)."""
        L = lib()
        self.ps = ps
        n = len(ps["v"])
        self.n = n
        v = np.ascontiguousarray(ps["v"], np.float32)
        vn = np.ascontiguousarray(ps["vn"], np.float32)
        vt = np.ascontiguousarray(ps["vt"], np.float32)
        mtl = np.ascontiguousarray(ps["mtl"], np.int32)
        M = (OrcMaterial * len(ps["materials"]))()
        for i, m in enumerate(ps["materials"]):
            M[i].Kd[:] = [float(x) for x in m["Kd"]]
            M[i].Ks[:] = [float(x) for x in m["Ks"]]
            M[i].Tr[:] = [float(x) for x in m["Tr"]]
            M[i].Ns, M[i].Ni, M[i].texture = float(m["Ns"]), float(m["Ni"]), int(m.get("texture", -1))
        lm = np.array([l[0] for l in ps["lights"]], np.int32)
        lr = np.array([l[1] for l in ps["lights"]], np.float32).reshape(-1, 3)
        self._tex = [np.ascontiguousarray(t, np.uint8) for t in ps.get("textures", [])]
        T = (OrcTexture * max(1, len(self._tex)))()
        for i, t in enumerate(self._tex):
            T[i].rows, T[i].cols, T[i].bgr = t.shape[0], t.shape[1], t.ctypes.data
        e, la, u = (np.ascontiguousarray(ps[k], np.float32) for k in ("eye", "lookat", "up"))
        self.h = L.orc_scene_create(n, v.ctypes.data, vn.ctypes.data, vt.ctypes.data, mtl.ctypes.data, len(M), M,
                                    len(lm), lm.ctypes.data, lr.ctypes.data, len(self._tex), T, e.ctypes.data,
                                    la.ctypes.data, u.ctypes.data, float(ps["fovy"]), ps["width"], ps["height"], leaf_num)
        self.width, self.height = ps["width"], ps["height"]

    def order(self):
        p = np.zeros(self.n, np.int32)
        lib().orc_get_order(self.h, p.ctypes.data)
        return p

    def derived(self):
        n = self.n
        normal, center = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32)
        area, em = np.zeros(n, np.float64), np.zeros(n, np.int32)
        lib().orc_get_derived(self.h, normal.ctypes.data, center.ctypes.data, area.ctypes.data, em.ctypes.data)
        return dict(normal=normal, center=center, area=area, emissive=em)

    def nodes(self):
        n = lib().orc_num_nodes(self.h)
        b, l = np.zeros((n, 6), np.float32), np.zeros((n, 4), np.int32)
        lib().orc_get_nodes(self.h, b.ctypes.data, l.ctypes.data)
        return b, l

    def camera(self):
        o = np.zeros(12, np.float32)
        lib().orc_get_camera(self.h, o.ctypes.data)
        return o

    def bvh_stats(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        lib().orc_bvh_stats(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def trace(self, rays, threads=0, want_pn=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = len(rays)
        t, ids = np.zeros(n, np.float32), np.zeros(n, np.int32)
        pn = np.zeros((n, 3), np.float32) if want_pn else None
        hp = np.zeros((n, 3), np.float32) if want_pn else None
        lib().orc_trace(self.h, rays.ctypes.data, n, t.ctypes.data, ids.ctypes.data,
                        pn.ctypes.data if want_pn else None, hp.ctypes.data if want_pn else None, threads)
        return (ids, t, pn, hp) if want_pn else (ids, t)

    def trace_counts(self, rays, mode, threads=0):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = len(rays)
        b, t = C.c_uint64(), C.c_uint64()
        ids = np.zeros(n, np.int32)
        lib().orc_trace_counts(self.h, rays.ctypes.data, n, mode, C.byref(b), C.byref(t), ids.ctypes.data, threads)
        return b.value / max(n, 1), t.value / max(n, 1), ids

    def render(self, spp, seed=0, max_depth=0, sample_begin=0, sample_end=None, threads=0):
        img = np.zeros((self.height, self.width, 3), np.float64)
        counts = (C.c_uint64 * 2)(0, 0)
        lib().orc_render(self.h, spp, sample_begin, spp if sample_end is None else sample_end, max_depth, seed,
                         img.ctypes.data, counts, threads)
        return img, (counts[0], counts[1])

    def render_pixels(self, pixels, spp, seed=0, max_depth=0, sample_begin=0, sample_end=None, threads=0):
        """The colours orc_render would give the row-major pixel indices `pixels` of the W x H frame: (n, 3) float64."""
        pixels = np.ascontiguousarray(pixels, np.int32)
        out = np.zeros((len(pixels), 3), np.float64)
        lib().orc_render_pixels(self.h, spp, sample_begin, spp if sample_end is None else sample_end, max_depth, seed,
                                pixels.ctypes.data, len(pixels), out.ctypes.data, threads)
        return out

    def shade_batch(self, rays, ids, t, seed=0, sample=0, max_depth=0, threads=0):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        ids, t = np.ascontiguousarray(ids, np.int32), np.ascontiguousarray(t, np.float32)
        out = np.zeros((len(rays), 3), np.float32)
        lib().orc_shade_batch(self.h, rays.ctypes.data, ids.ctypes.data, t.ctypes.data, len(rays), seed, sample,
                              max_depth, out.ctypes.data, threads)
        return out

    def primary_ray(self, i, j, k, seed=0):
        r = np.zeros(6, np.float32)
        lib().orc_primary_ray(self.h, i, j, k, seed, r.ctypes.data)
        return r

    def __del__(self):
        try:
            lib().orc_scene_destroy(self.h)
        except Exception:
            pass


def philox(ctr, key):
    c, k, o = np.array(ctr, np.uint32), np.array(key, np.uint32), np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
    return o


def uniform(seed, pixel, sample, depth, slot):
    return lib().orc_uniform(seed, pixel, sample, depth, slot)


def walk_layout(host_scene, rays, threads=0):
    """CPU walk (oracle/layout_walk.cpp) of the product's fast layout for `host_scene` (trt_layout_build): returns
    (ids, t, work) with ids = post-build triangle index, -1 miss, -2 = ray class the fast layout does not serve, and
    work = per-ray averages of wide nodes visited / child boxes tested / leaves scanned / triangles tested."""
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
    n = len(rays)
    handle, view = host_scene.layout_arrays()
    try:
        ids, t = np.zeros(n, np.int32), np.zeros(n, np.float32)
        cnt = (C.c_uint64 * 4)(0, 0, 0, 0)
        lib().orc_walk_layout(C.addressof(view), rays.ctypes.data, n, ids.ctypes.data, t.ctypes.data, C.addressof(cnt), threads)
    finally:
        host_scene.free_layout(handle)
    served = max(int((ids != -2).sum()), 1)
    work = dict(zip(("nodes", "boxes", "leaves", "tris"), (c / served for c in cnt)))
    return ids, t, work
