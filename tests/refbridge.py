"""ctypes view of oracle/_ref/libref.so (tier A: the UNMODIFIED reference objects). Test infrastructure only."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBREF = os.path.join(ROOT, "oracle", "_ref", "libref.so")
REF_SCENES = "/root/reference/RayTracingOnCPU/example-scenes-cg22"

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def available():
    return os.path.exists(LIBREF)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIBREF)
        L.ref_scene_load.restype = C.c_void_p
        L.ref_scene_load.argtypes = [C.c_char_p] * 4 + [C.c_int]
        L.ref_num_triangles.argtypes = [C.c_void_p]
        L.ref_num_materials.argtypes = [C.c_void_p]
        L.ref_material_name.restype = C.c_char_p
        L.ref_material_name.argtypes = [C.c_void_p, C.c_int]
        L.ref_get_triangles.argtypes = [C.c_void_p] + [C.c_void_p] * 9
        L.ref_get_material.argtypes = [C.c_void_p, C.c_int, _f32p, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
        L.ref_get_camera.argtypes = [C.c_void_p, _f32p]
        L.ref_image_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_bvh_stats.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 3
        L.ref_bvh_flatten.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_trace.restype = C.c_long
        L.ref_trace.argtypes = [C.c_void_p, _f32p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_render.argtypes = [C.c_void_p, C.c_int, _f64p, C.c_int, C.c_uint]
        _lib = L
    return _lib


class RefScene:
    """The reference's Scene + BVH, loaded and built by the reference's own code."""

    def __init__(self, xml, obj, mtl, basedir, build=True):
        self.h = lib().ref_scene_load(xml.encode(), obj.encode(), mtl.encode(), basedir.encode(), int(build))
        self.n = lib().ref_num_triangles(self.h)

    def triangles(self):
        n = self.n
        out = dict(v=np.zeros((n, 9), np.float32), vn=np.zeros((n, 9), np.float32), vt=np.zeros((n, 6), np.float32),
                   normal=np.zeros((n, 3), np.float32), center=np.zeros((n, 3), np.float32),
                   area=np.zeros(n, np.float64), emissive=np.zeros(n, np.int32), mtl=np.zeros(n, np.int32),
                   canon=np.zeros(n, np.int32))
        lib().ref_get_triangles(self.h, *[out[k].ctypes.data for k in
                                         ("v", "vn", "vt", "normal", "center", "area", "emissive", "mtl", "canon")])
        return out

    def material_names(self):
        return [lib().ref_material_name(self.h, i).decode() for i in range(lib().ref_num_materials(self.h))]

    def material(self, i):
        o = np.zeros(16, np.float32)
        a, n = C.c_double(), C.c_int32()
        lib().ref_get_material(self.h, i, o, C.byref(a), C.byref(n))
        return o, a.value, n.value

    def camera(self):
        o = np.zeros(12, np.float32)
        lib().ref_get_camera(self.h, o)
        return o

    def image_size(self):
        w, h = C.c_int(), C.c_int()
        lib().ref_image_size(self.h, C.byref(w), C.byref(h))
        return w.value, h.value

    def bvh_stats(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        lib().ref_bvh_stats(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def bvh_flatten(self):
        n = lib().ref_bvh_flatten(self.h, None, None, 0)
        boxes, links = np.zeros((n, 6), np.float32), np.zeros((n, 4), np.int32)
        lib().ref_bvh_flatten(self.h, boxes.ctypes.data, links.ctypes.data, n)
        return boxes, links

    def trace(self, rays, threads=0, want_pn=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = len(rays)
        t, ids = np.zeros(n, np.float32), np.zeros(n, np.int32)
        pn = np.zeros((n, 3), np.float32) if want_pn else None
        hp = np.zeros((n, 3), np.float32) if want_pn else None
        lib().ref_trace(self.h, rays, n, t.ctypes.data, ids.ctypes.data, pn.ctypes.data if want_pn else None,
                        hp.ctypes.data if want_pn else None, threads)
        return (t, ids, pn, hp) if want_pn else (t, ids)

    def render(self, spp, threads=0, seed=1):
        w, h = self.image_size()
        img = np.zeros((h, w, 3), np.float64)
        lib().ref_render(self.h, spp, img, threads, seed)
        return img
