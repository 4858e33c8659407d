"""ctypes bindings of include/trt.h + include/trt_host.h.

There is no fallback of any kind: if libtrt_b200.so is missing, or no sm_100 GPU is present when a
DeviceScene is created, the call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
INF = np.float32(114514.0)
TRACE_DEVICE_PTRS, TRACE_EXHAUSTIVE, TRACE_REFTOPO, TRACE_PLAIN, TRACE_POOLED, TRACE_PERSISTENT = 1, 2, 4, 8, 16, 32
RENDER_REFTOPO = 1
RENDER_PLAIN = 2
RENDER_PROFILE = 4
RENDER_PEER_REDUCE = 8


class TrtError(RuntimeError):
    pass


class Material(C.Structure):
    _fields_ = [("Kd", C.c_float * 3), ("Ks", C.c_float * 3), ("Tr", C.c_float * 3), ("Ns", C.c_float),
                ("Ni", C.c_float), ("radiance", C.c_float * 3), ("is_emissive", C.c_int32), ("texture", C.c_int32),
                ("area", C.c_double)]


class Light(C.Structure):
    _fields_ = [("material", C.c_int32), ("first_tri", C.c_int32), ("n_tris", C.c_int32), ("_pad", C.c_int32)]


class Texture(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("bgr", C.POINTER(C.c_uint8))]


class SceneDesc(C.Structure):
    _fields_ = [("n_tris", C.c_int32), ("v", C.POINTER(C.c_float)), ("vn", C.POINTER(C.c_float)),
                ("vt", C.POINTER(C.c_float)), ("normal", C.POINTER(C.c_float)), ("mtl", C.POINTER(C.c_int32)),
                ("n_nodes", C.c_int32), ("node_box", C.POINTER(C.c_float)), ("node_link", C.POINTER(C.c_int32)),
                ("n_materials", C.c_int32), ("materials", C.POINTER(Material)),
                ("n_lights", C.c_int32), ("lights", C.POINTER(Light)),
                ("n_light_tris", C.c_int32), ("light_v", C.POINTER(C.c_float)), ("light_vn", C.POINTER(C.c_float)),
                ("light_cum_area", C.POINTER(C.c_double)),
                ("n_textures", C.c_int32), ("textures", C.POINTER(Texture)),
                ("eye", C.c_float * 3), ("lower_left_corner", C.c_float * 3), ("horizontal", C.c_float * 3),
                ("vertical", C.c_float * 3), ("width", C.c_int32), ("height", C.c_int32)]


class RenderParams(C.Structure):
    _fields_ = [("spp", C.c_int32), ("sample_begin", C.c_int32), ("sample_end", C.c_int32),
                ("max_depth", C.c_int32), ("seed", C.c_uint64), ("batch_paths", C.c_int32), ("flags", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64), ("paths", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("last_render_ms", C.c_double), ("last_trace_ms", C.c_double),
                ("accel_nodes", C.c_int32), ("accel_leaves", C.c_int32), ("ref_depth", C.c_int32),
                ("device", C.c_int32), ("accel_slivers", C.c_int32), ("accel_needles", C.c_int32),
                ("ms_trace", C.c_double), ("ms_shade", C.c_double), ("ms_shadow", C.c_double),
                ("ms_accumulate", C.c_double), ("rays_strict", C.c_uint64)]


class ShadeParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("sample", C.c_int32), ("max_depth", C.c_int32), ("flags", C.c_uint32),
                ("_pad", C.c_uint32)]


class LayoutReport(C.Structure):
    _fields_ = [("n_tris", C.c_int32), ("n_fast_tris", C.c_int32), ("n_dropped", C.c_int32), ("use_wide", C.c_int32),
                ("wide_nodes", C.c_int32), ("wide_depth", C.c_int32), ("ref_leaves", C.c_int32), ("ref_depth", C.c_int32),
                ("slivers", C.c_int32), ("needles", C.c_int32), ("violations", C.c_int32), ("sah_wide", C.c_double),
                ("sah_inner", C.c_double), ("sah_leaf", C.c_double)]


class LayoutView(C.Structure):
    _fields_ = [("n_wide_nodes", C.c_int32), ("wide_root", C.c_int32), ("n_fast_tris", C.c_int32), ("n_ref_leaves", C.c_int32),
                ("n_ref_inner", C.c_int32), ("check_leaf_box", C.c_int32), ("strict_origin_limit", C.c_float),
                ("miss_key", C.c_uint32), ("wide_nodes", C.c_void_p), ("fast_geom", C.c_void_p), ("fast_key", C.c_void_p),
                ("fast_orig", C.c_void_p), ("fast_leaf", C.c_void_p), ("ref_leaf_box", C.c_void_p),
                ("ref_leaf_parent", C.c_void_p), ("ref_nodes", C.c_void_p)]


# every symbol include/trt.h and include/trt_host.h declare (tests check the library exports all of them)
EXPORTS = ["trt_device_count", "trt_scene_create", "trt_scene_create_cached", "trt_layout_build_cached", "trt_scene_replicate", "trt_scene_destroy", "trt_host_alloc", "trt_host_free",
           "trt_trace_closest", "trt_trace_closest_multi", "trt_trace_closest_async", "trt_trace_counters", "trt_hit_attributes", "trt_render",
           "trt_render_accumulate", "trt_resolve", "trt_render_multi", "trt_shade", "trt_accum_create", "trt_accum_destroy",
           "trt_accum_save", "trt_accum_load", "trt_layout_check", "trt_layout_build", "trt_layout_free", "trt_get_stats", "trt_reset_stats", "trt_last_error",
           "trt_version", "trt_host_scene_load", "trt_host_scene_from_arrays", "trt_host_scene_desc",
           "trt_host_scene_faces", "trt_host_scene_material_name", "trt_host_scene_build_seconds",
           "trt_host_scene_free", "trt_host_scene_save", "trt_host_scene_load_cache", "trt_decode_jpeg", "trt_write_png", "trt_write_pfm"]


def library_path():
    return os.path.join(_HERE, "libtrt_b200.so")


_lib = None


def load_library():
    """Load libtrt_b200.so (built in-tree by __graft_entry__.build()). Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise TrtError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no Python/CPU fallback)" % path)
    L = C.CDLL(path)
    vp, cp, i32, u32, sz = C.c_void_p, C.c_char_p, C.c_int32, C.c_uint32, C.c_size_t
    L.trt_last_error.restype = cp
    L.trt_scene_create.argtypes = [C.POINTER(SceneDesc), C.c_int, C.POINTER(vp)]
    L.trt_scene_replicate.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.trt_scene_destroy.argtypes = [vp]
    L.trt_scene_destroy.restype = None
    L.trt_host_alloc.restype = vp
    L.trt_host_alloc.argtypes = [sz]
    L.trt_host_free.argtypes = [vp]
    L.trt_host_free.restype = None
    L.trt_trace_closest.argtypes = [vp, vp, sz, vp, vp, u32]
    L.trt_trace_closest_async.argtypes = [vp, vp, sz, vp, vp, u32, vp]
    L.trt_trace_closest_multi.argtypes = [C.POINTER(vp), i32, vp, sz, vp, vp, u32]
    L.trt_hit_attributes.argtypes = [vp, vp, vp, vp, sz, vp, vp]
    L.trt_trace_counters.argtypes = [vp, vp, sz, vp]
    L.trt_render.argtypes = [vp, C.POINTER(RenderParams), vp]
    L.trt_layout_check.argtypes = [C.POINTER(SceneDesc), C.POINTER(LayoutReport)]
    L.trt_layout_build.argtypes = [C.POINTER(SceneDesc), C.POINTER(vp), C.POINTER(LayoutView)]
    L.trt_layout_build_cached.argtypes = [C.POINTER(SceneDesc), C.c_char_p, C.POINTER(C.c_int32), C.POINTER(vp), C.POINTER(LayoutView)]
    L.trt_scene_create_cached.argtypes = [C.POINTER(SceneDesc), C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(vp)]
    L.trt_layout_free.argtypes = [vp]
    L.trt_layout_free.restype = None
    L.trt_render_accumulate.argtypes = [vp, C.POINTER(RenderParams), vp, vp]
    L.trt_resolve.argtypes = [vp, vp, i32, vp, vp, vp]
    L.trt_render_multi.argtypes = [C.POINTER(vp), i32, C.POINTER(RenderParams), vp, vp]
    L.trt_shade.argtypes = [vp, vp, vp, vp, sz, C.POINTER(ShadeParams), vp]
    L.trt_accum_create.restype = vp
    L.trt_accum_create.argtypes = [vp]
    L.trt_accum_destroy.restype = None
    L.trt_accum_destroy.argtypes = [vp, vp]
    L.trt_accum_save.argtypes = [vp, vp, i32, i32, C.c_uint64, i32, cp]
    L.trt_accum_load.argtypes = [vp, cp, vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_uint64), C.POINTER(i32)]
    L.trt_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.trt_reset_stats.argtypes = [vp]
    L.trt_host_scene_load.argtypes = [cp, cp, cp, cp, C.c_int, C.POINTER(vp)]
    L.trt_host_scene_from_arrays.argtypes = [i32, vp, vp, vp, vp, i32, C.POINTER(Material), i32, vp, vp, vp, vp, vp,
                                             C.c_double, i32, i32, C.c_int, C.POINTER(vp)]
    L.trt_host_scene_desc.restype = C.POINTER(SceneDesc)
    L.trt_host_scene_desc.argtypes = [vp]
    L.trt_host_scene_faces.restype = C.POINTER(C.c_int32)
    L.trt_host_scene_faces.argtypes = [vp]
    L.trt_host_scene_material_name.restype = cp
    L.trt_host_scene_material_name.argtypes = [vp, C.c_int]
    L.trt_host_scene_build_seconds.restype = C.c_double
    L.trt_host_scene_build_seconds.argtypes = [vp]
    L.trt_host_scene_free.argtypes = [vp]
    L.trt_host_scene_save.argtypes = [vp, cp]
    L.trt_host_scene_load_cache.argtypes = [cp, C.POINTER(vp)]
    L.trt_host_scene_free.restype = None
    L.trt_write_png.argtypes = [cp, i32, i32, vp, C.c_int]
    L.trt_write_pfm.argtypes = [cp, i32, i32, vp]
    L.trt_decode_jpeg.argtypes = [cp, C.POINTER(i32), C.POINTER(i32), vp, sz]
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise TrtError("%s failed (%d): %s" % (what, rc, load_library().trt_last_error().decode()))


def _np(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return np.zeros(shape, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype).reshape(shape).copy()


def decode_jpeg(path):
    """Baseline JPEG -> (rows, cols, 3) uint8 BGR with the library's own decoder (csrc/host/jpeg_decoder.cpp)."""
    L = load_library()
    r, c = C.c_int32(), C.c_int32()
    _check(L.trt_decode_jpeg(path.encode(), C.byref(r), C.byref(c), None, 0), "trt_decode_jpeg")
    out = np.empty((r.value, c.value, 3), np.uint8)
    _check(L.trt_decode_jpeg(path.encode(), C.byref(r), C.byref(c), out.ctypes.data, out.size), "trt_decode_jpeg")
    return out


class HostScene:
    """The host side the drop-in surface keeps: loaders + buildBVH (csrc/host), no GPU involved."""

    def __init__(self, handle):
        self.h = handle
        self.lib = load_library()
        self.desc = self.lib.trt_host_scene_desc(self.h).contents

    def save(self, path):
        """Binary scene cache (trt_host_scene_save): parsed scene + decoded textures + buildBVH topology in one file."""
        _check(self.lib.trt_host_scene_save(self.h, path.encode()), "trt_host_scene_save")

    @classmethod
    def load_cache(cls, path):
        h = C.c_void_p()
        _check(load_library().trt_host_scene_load_cache(path.encode(), C.byref(h)), "trt_host_scene_load_cache")
        return cls(h)

    def layout_arrays(self, layout_cache=None):
        """(handle, LayoutView): the GPU layouts of this scene as plain host arrays (trt_layout_build); release the
        handle with free_layout().  Inspection / test tooling — the device path never reads it.
        With layout_cache=<path> the layouts go through the layout cache (trt_layout_build_cached) and the call returns
        (handle, view, from_cache)."""
        h, view = C.c_void_p(), LayoutView()
        if layout_cache is None:
            _check(self.lib.trt_layout_build(C.byref(self.desc), C.byref(h), C.byref(view)), "trt_layout_build")
            return h, view
        hit = C.c_int32(0)
        _check(self.lib.trt_layout_build_cached(C.byref(self.desc), os.fsencode(layout_cache), C.byref(hit), C.byref(h), C.byref(view)),
               "trt_layout_build_cached")
        return h, view, bool(hit.value)

    def free_layout(self, handle):
        self.lib.trt_layout_free(handle)

    def layout_check(self):
        """Builds the GPU layouts on the host and verifies their invariants (trt_layout_check, no device needed)."""
        rep = LayoutReport()
        _check(self.lib.trt_layout_check(C.byref(self.desc), C.byref(rep)), "trt_layout_check")
        return {k: getattr(rep, k) for k, _ in LayoutReport._fields_}

    @classmethod
    def load(cls, xml, obj, mtl, basedir, leaf_num=8):
        h = C.c_void_p()
        _check(load_library().trt_host_scene_load(xml.encode(), obj.encode(), mtl.encode(), basedir.encode(),
                                                  leaf_num, C.byref(h)), "trt_host_scene_load")
        return cls(h)

    @classmethod
    def from_arrays(cls, v9, mtl, materials, lights, eye, lookat, up, fovy, width, height, vn9=None, vt6=None,
                    leaf_num=8):
        """materials: list of dict(Kd,Ks,Tr,Ns,Ni); lights: list of (material index, radiance3)."""
        v9 = np.ascontiguousarray(v9, np.float32).reshape(-1, 9)
        n = len(v9)
        mtl = np.ascontiguousarray(mtl, np.int32)
        vn9 = None if vn9 is None else np.ascontiguousarray(vn9, np.float32)
        vt6 = None if vt6 is None else np.ascontiguousarray(vt6, np.float32)
        M = (Material * len(materials))()
        for i, m in enumerate(materials):
            M[i].Kd[:] = m.get("Kd", (0, 0, 0))
            M[i].Ks[:] = m.get("Ks", (0, 0, 0))
            M[i].Tr[:] = m.get("Tr", (0, 0, 0))
            M[i].Ns, M[i].Ni, M[i].texture = m.get("Ns", 1.0), m.get("Ni", 1.0), -1
        lm = np.array([l[0] for l in lights], np.int32)
        lr = np.array([l[1] for l in lights], np.float32).reshape(-1, 3)
        e, la, u = (np.array(x, np.float32) for x in (eye, lookat, up))
        h = C.c_void_p()
        p = lambda a: None if a is None else a.ctypes.data
        _check(load_library().trt_host_scene_from_arrays(n, p(v9), p(vn9), p(vt6), p(mtl), len(materials), M,
                                                         len(lights), p(lm), p(lr), p(e), p(la), p(u), float(fovy),
                                                         width, height, leaf_num, C.byref(h)),
               "trt_host_scene_from_arrays")
        return cls(h)

    # ---- numpy views (copies) of the POD description
    @property
    def n_tris(self):
        return self.desc.n_tris

    def triangles(self):
        d, n = self.desc, self.desc.n_tris
        return dict(v=_np(d.v, (n, 9), np.float32), vn=_np(d.vn, (n, 9), np.float32), vt=_np(d.vt, (n, 6), np.float32),
                    normal=_np(d.normal, (n, 3), np.float32), mtl=_np(d.mtl, (n,), np.int32),
                    face=_np(self.lib.trt_host_scene_faces(self.h), (n,), np.int32))

    def nodes(self):
        d = self.desc
        return _np(d.node_box, (d.n_nodes, 6), np.float32), _np(d.node_link, (d.n_nodes, 4), np.int32)

    def material_names(self):
        return [self.lib.trt_host_scene_material_name(self.h, i).decode() for i in range(self.desc.n_materials)]

    def materials(self):
        out = []
        for i in range(self.desc.n_materials):
            m = self.desc.materials[i]
            out.append(dict(Kd=tuple(m.Kd), Ks=tuple(m.Ks), Tr=tuple(m.Tr), Ns=m.Ns, Ni=m.Ni,
                            radiance=tuple(m.radiance), is_emissive=m.is_emissive, texture=m.texture, area=m.area))
        return out

    def lights(self):
        d = self.desc
        ls = [dict(material=d.lights[i].material, first_tri=d.lights[i].first_tri, n_tris=d.lights[i].n_tris)
              for i in range(d.n_lights)]
        return ls, _np(d.light_v, (d.n_light_tris, 9), np.float32), _np(d.light_vn, (d.n_light_tris, 9), np.float32), \
            _np(d.light_cum_area, (d.n_light_tris,), np.float64)

    def textures(self):
        out = []
        for i in range(self.desc.n_textures):
            t = self.desc.textures[i]
            out.append(_np(t.bgr, (t.rows, t.cols, 3), np.uint8))
        return out

    def camera(self):
        d = self.desc
        return dict(eye=np.array(d.eye[:], np.float32), llc=np.array(d.lower_left_corner[:], np.float32),
                    horizontal=np.array(d.horizontal[:], np.float32), vertical=np.array(d.vertical[:], np.float32),
                    width=d.width, height=d.height)

    def root_box(self):
        boxes, _ = self.nodes()
        return boxes[0, :3].copy(), boxes[0, 3:].copy()

    @property
    def build_seconds(self):
        return self.lib.trt_host_scene_build_seconds(self.h)

    def close(self):
        if self.h:
            self.lib.trt_host_scene_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def trace_closest_multi(devs, rays, flags=0, out_id=None, out_t=None):
    """trt_trace_closest_multi: one host batch sharded by ray index over the replicas `devs` (one per GPU)."""
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
    n = len(rays)
    ids = np.empty(n, np.int32) if out_id is None else out_id
    t = np.empty(n, np.float32) if out_t is None else out_t
    handles = (C.c_void_p * len(devs))(*[d.h for d in devs])
    _check(devs[0].lib.trt_trace_closest_multi(handles, len(devs), rays.ctypes.data, n, ids.ctypes.data, t.ctypes.data, flags),
           "trt_trace_closest_multi")
    return ids, t


def render_multi(devs, spp, seed=0, max_depth=0, flags=0, batch_paths=0, want_rgb8=False, out=None):
    """trt_render_multi: the frame rendered by all of `devs` (DeviceScene replicas of one scene on different GPUs) inside
    the library — one host thread per GPU over sample ranges, one reduce onto devs[0]'s GPU (ncclReduce, or the
    library's own peer-memory kernel with RENDER_PEER_REDUCE), resolve there."""
    d0 = devs[0]
    img = d0._image_out(out)
    rgb = np.empty((d0.height, d0.width, 3), np.uint8) if want_rgb8 else None
    handles = (C.c_void_p * len(devs))(*[d.h for d in devs])
    p = d0.params(spp, 0, spp, max_depth, seed, batch_paths, flags)
    _check(d0.lib.trt_render_multi(handles, len(devs), C.byref(p), img.ctypes.data, rgb.ctypes.data if want_rgb8 else None),
           "trt_render_multi")
    return (img, rgb) if want_rgb8 else img


class DeviceScene:
    """Device-resident scene: the GPU hot path (closest hit, render). Needs an sm_100 GPU."""

    def __init__(self, host_scene, device=0, _replica_of=None, layout_cache=None):
        self.lib = load_library()
        self.host = host_scene
        self.h = C.c_void_p()
        self.layout_from_cache = False  # layout_cache=<path>: trt_scene_create_cached (layouts read from / written to the file)
        if _replica_of is None and layout_cache is not None:
            hit = C.c_int32(0)
            _check(self.lib.trt_scene_create_cached(C.byref(host_scene.desc), device, os.fsencode(layout_cache), C.byref(hit),
                                                    C.byref(self.h)), "trt_scene_create_cached")
            self.layout_from_cache = bool(hit.value)
        elif _replica_of is None:
            _check(self.lib.trt_scene_create(C.byref(host_scene.desc), device, C.byref(self.h)), "trt_scene_create")
        else:
            _check(self.lib.trt_scene_replicate(_replica_of.h, device, C.byref(self.h)), "trt_scene_replicate")
        self.width, self.height = host_scene.desc.width, host_scene.desc.height

    def replicate(self, device):
        """A copy of this scene on another GPU (trt_scene_replicate: device-to-device, no host rebuild of the layouts)."""
        return DeviceScene(self.host, device, _replica_of=self)

    def trace_closest(self, rays, flags=0, out_id=None, out_t=None):
        """rays: (n,6) float32 host array. Returns (tri_id int32[n], t float32[n])."""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = len(rays)
        ids = np.empty(n, np.int32) if out_id is None else out_id
        t = np.empty(n, np.float32) if out_t is None else out_t
        _check(self.lib.trt_trace_closest(self.h, rays.ctypes.data, n, ids.ctypes.data, t.ctypes.data, flags),
               "trt_trace_closest")
        return ids, t

    def trace_closest_ptr(self, rays_ptr, n, id_ptr, t_ptr, flags=0):
        """Raw-pointer blocking form (host pinned / pageable pointers, or device pointers with TRACE_DEVICE_PTRS)."""
        _check(self.lib.trt_trace_closest(self.h, rays_ptr, n, id_ptr, t_ptr, flags), "trt_trace_closest")

    def trace_closest_async(self, d_rays_ptr, n, d_id_ptr, d_t_ptr, flags=0, stream=0):
        _check(self.lib.trt_trace_closest_async(self.h, d_rays_ptr, n, d_id_ptr, d_t_ptr, flags, stream),
               "trt_trace_closest_async")

    def trace_counters(self, rays):
        """Per-ray averages of the default traversal's work: wide nodes, box tests, leaves, leaf triangles."""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        out = np.zeros(4, np.uint64)
        _check(self.lib.trt_trace_counters(self.h, rays.ctypes.data, len(rays), out.ctypes.data), "trt_trace_counters")
        return dict(zip(("nodes", "boxes", "leaves", "tris"), (out / max(len(rays), 1)).tolist()))

    def hit_attributes(self, rays, ids, t):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        n = len(rays)
        ids, t = np.ascontiguousarray(ids, np.int32), np.ascontiguousarray(t, np.float32)
        hp, pn = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        _check(self.lib.trt_hit_attributes(self.h, rays.ctypes.data, ids.ctypes.data, t.ctypes.data, n,
                                           hp.ctypes.data, pn.ctypes.data), "trt_hit_attributes")
        return hp, pn

    def params(self, spp, sample_begin=0, sample_end=None, max_depth=0, seed=0, batch_paths=0, flags=0):
        return RenderParams(spp, sample_begin, spp if sample_end is None else sample_end, max_depth, seed,
                            batch_paths, flags)

    def pinned_image(self):
        """A float64 (H, W, 3) array in page-locked host memory (trt_host_alloc), owned by this scene and reused:
        pass it as `out=` to render / resolve so that the device-to-host copy of the frame is one DMA into memory
        that is already mapped (a fresh pageable array costs a staged copy plus a page fault per 4 kB).  The array is
        valid until close()."""
        if getattr(self, "_pinned_img", None) is None:
            n = self.height * self.width * 3
            ptr = self.lib.trt_host_alloc(n * 8)
            if not ptr:
                raise TrtError("trt_host_alloc failed: " + self.lib.trt_last_error().decode())
            self._pinned_ptr = ptr
            self._pinned_img = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), (n,)).reshape(
                self.height, self.width, 3)
        return self._pinned_img

    def _image_out(self, out):
        if out is None:
            return np.empty((self.height, self.width, 3), np.float64)
        if out.dtype != np.float64 or out.shape != (self.height, self.width, 3) or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape (H, W, 3)")
        return out

    def render(self, spp, seed=0, max_depth=0, sample_begin=0, sample_end=None, batch_paths=0, flags=0, out=None):
        """The reference's image buffer: float64 (H, W, 3), divided by spp."""
        img = self._image_out(out)
        p = self.params(spp, sample_begin, sample_end, max_depth, seed, batch_paths, flags)
        _check(self.lib.trt_render(self.h, C.byref(p), img.ctypes.data), "trt_render")
        return img

    def render_accumulate(self, params, d_accum_ptr, stream=0):
        _check(self.lib.trt_render_accumulate(self.h, C.byref(params), d_accum_ptr, stream), "trt_render_accumulate")

    def resolve(self, d_accum_ptr, spp, want_rgb8=False, stream=0, out=None):
        img = self._image_out(out)
        rgb = np.empty((self.height, self.width, 3), np.uint8) if want_rgb8 else None
        _check(self.lib.trt_resolve(self.h, d_accum_ptr, spp, img.ctypes.data,
                                    rgb.ctypes.data if want_rgb8 else None, stream), "trt_resolve")
        return (img, rgb) if want_rgb8 else img

    def shade(self, rays, ids, t, seed=0, sample=0, max_depth=0, flags=0):
        """PathTracing::shade for a batch of traced rays (trt_shade): (n, 3) float32 radiance, wi = -direction."""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        ids, t = np.ascontiguousarray(ids, np.int32), np.ascontiguousarray(t, np.float32)
        out = np.empty((len(rays), 3), np.float32)
        p = ShadeParams(seed, sample, max_depth, flags, 0)
        _check(self.lib.trt_shade(self.h, rays.ctypes.data, ids.ctypes.data, t.ctypes.data, len(rays), C.byref(p),
                                  out.ctypes.data), "trt_shade")
        return out

    # ---- accumulation buffers and checkpoints (trt_accum_*)
    def accum_create(self):
        p = self.lib.trt_accum_create(self.h)
        if not p:
            raise TrtError("trt_accum_create failed: " + self.lib.trt_last_error().decode())
        return p

    def accum_destroy(self, d_accum):
        self.lib.trt_accum_destroy(self.h, d_accum)

    def accum_save(self, d_accum, samples_done, spp, path, seed=0, max_depth=0):
        _check(self.lib.trt_accum_save(self.h, d_accum, samples_done, spp, seed, max_depth, path.encode()), "trt_accum_save")

    def accum_load(self, path, d_accum):
        """-> dict(samples_done, spp, seed, max_depth) of the checkpoint now in d_accum."""
        a, b, c, d = C.c_int32(), C.c_int32(), C.c_uint64(), C.c_int32()
        _check(self.lib.trt_accum_load(self.h, path.encode(), d_accum, C.byref(a), C.byref(b), C.byref(c), C.byref(d)),
               "trt_accum_load")
        return dict(samples_done=a.value, spp=b.value, seed=c.value, max_depth=d.value)

    def stats(self):
        s = Stats()
        _check(self.lib.trt_get_stats(self.h, C.byref(s)), "trt_get_stats")
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def reset_stats(self):
        _check(self.lib.trt_reset_stats(self.h), "trt_reset_stats")

    def close(self):
        if self.h:
            self.lib.trt_scene_destroy(self.h)
            self.h = None
            if getattr(self, "_pinned_img", None) is not None:
                self._pinned_img = None
                self.lib.trt_host_free(self._pinned_ptr)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
