"""Parity of the CUDA closest-hit path (through the C ABI) against the oracle: bit-exact triangle ids and
distance bits on the BASELINE config-2 ray population, in every traversal mode, plus edge cases."""
import numpy as np
import pytest

import tinyraytracing_b200 as trt
from conftest import SCENES, make_rays

pytestmark = pytest.mark.gpu

MODES = {"default": 0, "plain": trt.TRACE_PLAIN, "persistent": trt.TRACE_PERSISTENT, "pooled": trt.TRACE_POOLED, "reftopo": trt.TRACE_REFTOPO, "exhaustive": trt.TRACE_EXHAUSTIVE}


@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("mode", list(MODES))
def test_closest_hit_bit_exact(name, mode, host_scenes, oracle_scenes, device_scenes):
    rays = make_rays(host_scenes[name], oracle_scenes[name], 1 << 20, seed=0x5EED0001)
    oid, ot = oracle_scenes[name].trace(rays)
    ids, t = device_scenes[name].trace_closest(rays, MODES[mode])
    bad = np.flatnonzero(ids != oid)
    assert len(bad) == 0, "%d id mismatches, first ray %d gpu %d oracle %d" % (len(bad), bad[0], ids[bad[0]], oid[bad[0]])
    assert np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    assert (ids >= 0).sum() > len(rays) // 4  # the batch really hits geometry


@pytest.mark.parametrize("name", SCENES)
def test_gpu_reproduces_reference_golden(name, device_scenes):
    """Closest hits of the UNMODIFIED reference (tests/golden, tools/make_golden.py): distance bits, triangle
    identity (by content: identical triangles are one identity in the reference's HitRecord), hit point, pn."""
    import os

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + "_closest.npz"))
    dev = device_scenes[name]
    for flags in MODES.values():
        ids, t = dev.trace_closest(g["rays"], flags)
        assert np.array_equal(t.view(np.uint32), g["t"].view(np.uint32))
        assert np.array_equal(np.where(ids >= 0, g["canon"][np.maximum(ids, 0)], -1), g["id"])
    hp, pn = dev.hit_attributes(g["rays"], ids, t)
    hit = ids >= 0
    assert np.array_equal(hp[hit].view(np.uint32), g["hitpoint"][hit].view(np.uint32))
    ok = np.isfinite(g["pn"][hit]).all(axis=1)
    assert np.abs(pn[hit][ok] - g["pn"][hit][ok]).max() <= 1e-6


@pytest.mark.parametrize("name", SCENES)
def test_hit_attributes(name, host_scenes, oracle_scenes, device_scenes):
    rays = make_rays(host_scenes[name], oracle_scenes[name], 1 << 16, seed=11)
    oid, ot, opn, ohp = oracle_scenes[name].trace(rays, want_pn=True)
    ids, t = device_scenes[name].trace_closest(rays)
    hp, pn = device_scenes[name].hit_attributes(rays, ids, t)
    hit = ids >= 0
    assert np.array_equal(hp[hit].view(np.uint32), ohp[hit].view(np.uint32))  # S + d*t in float: bit-exact
    ok = np.isfinite(opn[hit]).all(axis=1)
    # pn goes through a double least-squares solve (Eigen QR in the reference, closed form on the device): tolerance 1e-6
    assert np.abs(pn[hit][ok] - opn[hit][ok]).max() <= 1e-6
    assert (~hit).sum() == 0 or (np.all(hp[~hit] == 0) and np.all(pn[~hit] == 0))  # HitRecord defaults on a miss


def test_edge_cases(host_scenes, oracle_scenes, device_scenes):
    name = "veach-mis"
    dev, orc, host = device_scenes[name], oracle_scenes[name], host_scenes[name]
    # empty batch
    ids, t = dev.trace_closest(np.zeros((0, 6), np.float32))
    assert len(ids) == 0 and len(t) == 0
    # single ray, and a ragged size that is not a multiple of the block / chunk
    for n in (1, 33, 1000003 % 4099):
        rays = make_rays(host, orc, max(n, 4), seed=n)[:n]
        ids, t = dev.trace_closest(rays)
        oid, ot = orc.trace(rays)
        assert np.array_equal(ids, oid) and np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    # axis-parallel directions (zero components -> +-inf reciprocals, inf*0 = NaN slabs), NaN / zero directions,
    # origins exactly on box faces and on surfaces
    lo, hi = host.root_box()
    rng = np.random.default_rng(5)
    o = rng.uniform(lo, hi, (6000, 3)).astype(np.float32)
    d = np.zeros((6000, 3), np.float32)
    d[np.arange(6000), rng.integers(0, 3, 6000)] = rng.choice([-1.0, 1.0], 6000)
    d[:500, 1] = -0.0
    o[1000:1500, 0] = lo[0]  # on the root box face
    o[1500:2000, 1] = hi[1]
    special = np.concatenate([o, d], 1)
    special[2000:2100, 3:] = 0.0       # zero direction
    special[2100:2200, 3:] = np.nan    # NaN direction (normalize of a zero vector in the reference)
    special[2200:2300, :3] = np.inf    # garbage origin
    for flags in MODES.values():
        ids, t = dev.trace_closest(special, flags)
        oid, ot = orc.trace(special)
        assert np.array_equal(ids, oid)
        assert np.array_equal(t.view(np.uint32), ot.view(np.uint32))


def test_pinned_and_device_pointer_entry_points(host_scenes, oracle_scenes, device_scenes):
    import ctypes as C

    name = "back"
    dev, orc, host = device_scenes[name], oracle_scenes[name], host_scenes[name]
    n = (1 << 21) + 12345  # more than one staging chunk, ragged tail
    rays = make_rays(host, orc, n, seed=3)
    oid, ot = orc.trace(rays)
    lib = trt.load_library()
    # pinned host buffers: the library DMAs straight from / to them
    pr, pi, pt = lib.trt_host_alloc(n * 24), lib.trt_host_alloc(n * 4), lib.trt_host_alloc(n * 4)
    assert pr and pi and pt
    C.memmove(pr, rays.ctypes.data, n * 24)
    dev.trace_closest_ptr(pr, n, pi, pt)
    ids = np.ctypeslib.as_array(C.cast(pi, C.POINTER(C.c_int32)), (n,)).copy()
    t = np.ctypeslib.as_array(C.cast(pt, C.POINTER(C.c_float)), (n,)).copy()
    for p in (pr, pi, pt):
        lib.trt_host_free(p)
    assert np.array_equal(ids, oid) and np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    # pageable path gives the same
    ids2, t2 = dev.trace_closest(rays)
    assert np.array_equal(ids2, oid) and np.array_equal(t2.view(np.uint32), ot.view(np.uint32))
    # device pointers through torch (plumbing only)
    import torch

    dr = torch.from_numpy(rays).cuda()
    di = torch.empty(n, dtype=torch.int32, device="cuda")
    dt = torch.empty(n, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    dev.trace_closest_ptr(dr.data_ptr(), n, di.data_ptr(), dt.data_ptr(), trt.TRACE_DEVICE_PTRS)
    assert np.array_equal(di.cpu().numpy(), oid) and np.array_equal(dt.cpu().numpy().view(np.uint32), ot.view(np.uint32))


@pytest.mark.parametrize("name", SCENES)
def test_full_size_batch_properties(name, host_scenes, oracle_scenes, device_scenes):
    """BASELINE config 2 at its full size (16 Mi rays): the pruned walk equals the exhaustive reference walk
    (order / pruning independence), misses carry (-1, INF), hits lie in [5e-4, INF], and a bounded sample is
    checked against the oracle."""
    n = 16 << 20
    rays = make_rays(host_scenes[name], oracle_scenes[name], n, seed=0x5EED0001)
    dev = device_scenes[name]
    ids, t = dev.trace_closest(rays)
    ids_x, t_x = dev.trace_closest(rays, trt.TRACE_EXHAUSTIVE)
    assert np.array_equal(ids, ids_x) and np.array_equal(t.view(np.uint32), t_x.view(np.uint32))
    miss = ids < 0
    assert np.all(t[miss] == trt.INF) and np.all(t[~miss] >= np.float32(0.0005)) and np.all(t[~miss] <= trt.INF)
    assert ids.max() < host_scenes[name].n_tris
    sel = np.random.default_rng(1).choice(n, 1 << 19, replace=False)
    oid, ot = oracle_scenes[name].trace(rays[sel])
    assert np.array_equal(ids[sel], oid) and np.array_equal(t[sel].view(np.uint32), ot.view(np.uint32))


def test_synthetic_mesh_from_arrays():
    """BASELINE config 5 entry path (trt_host_scene_from_arrays) at a size the oracle handles: 10k-triangle stress
    mesh, closest hits bit-exact against the oracle in the default and the exhaustive mode, render parity."""
    import oraclelib
    from tinyraytracing_b200 import workloads

    m = workloads.stress_mesh(71)
    cam = m["camera"]
    host = trt.HostScene.from_arrays(m["v9"], m["mtl"], m["materials"], m["lights"], cam["eye"], cam["lookat"], cam["up"],
                                     cam["fovy"], 96, 54, vn9=m["vn9"])
    ps = dict(v=m["v9"], vn=m["vn9"], vt=np.zeros((len(m["v9"]), 6), np.float32), mtl=m["mtl"],
              materials=[dict(m_, name=str(i)) for i, m_ in enumerate(m["materials"])], lights=m["lights"], textures=[],
              eye=np.array(cam["eye"], np.float32), lookat=np.array(cam["lookat"], np.float32),
              up=np.array(cam["up"], np.float32), fovy=np.float32(cam["fovy"]), width=96, height=54)
    orc = oraclelib.OracleScene(ps)
    dev = trt.DeviceScene(host, 0)
    rays = make_rays(host, orc, 1 << 19, seed=5)
    oid, ot = orc.trace(rays)
    for flags in MODES.values():
        ids, t = dev.trace_closest(rays, flags)
        assert np.array_equal(ids, oid) and np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    img = dev.render(4, seed=2)
    ref, _ = orc.render(4, seed=2)
    assert np.sqrt(((img - ref) ** 2).mean()) <= 1e-3 * ref.mean()
    dev.close()


@pytest.mark.parametrize("n_tris", (0, 1, 8, 9))
def test_degenerate_topologies(n_tris):
    """Empty scene (NULL root, bvh.cpp:148), a root that is itself a leaf (scanned WITHOUT a box test,
    bvh.cpp:151-154), and the smallest tree with an inner node; no lights."""
    import oraclelib

    rng = np.random.default_rng(n_tris)
    v = rng.uniform(-1, 1, (n_tris, 9)).astype(np.float32)
    mtl = np.zeros(n_tris, np.int32)
    mats = [dict(Kd=(0.5, 0.5, 0.5), Ks=(0, 0, 0), Tr=(1, 1, 1), Ns=1.0, Ni=1.0)]
    cam = dict(eye=(0.0, 0.0, -4.0), lookat=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fovy=40.0)
    host = trt.HostScene.from_arrays(v, mtl, mats, [], cam["eye"], cam["lookat"], cam["up"], cam["fovy"], 16, 16)
    dev = trt.DeviceScene(host, 0)
    n = 20000
    o = rng.uniform(-2, 2, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d], 1).astype(np.float32)
    if n_tris:
        # aim half of the rays at triangle interiors so that hits are plentiful
        tri = v[rng.integers(0, n_tris, n // 2)].reshape(-1, 3, 3)
        b = rng.dirichlet((1, 1, 1), n // 2)[:, :, None]
        tgt = (tri * b).sum(1)
        dd = tgt - rays[: n // 2, :3]
        rays[: n // 2, 3:] = dd / np.linalg.norm(dd, axis=1, keepdims=True)
        ps = dict(v=v, vn=np.zeros((n_tris, 9), np.float32), vt=np.zeros((n_tris, 6), np.float32), mtl=mtl,
                  materials=[dict(mats[0], name="m")], lights=[], textures=[], eye=np.array(cam["eye"], np.float32),
                  lookat=np.array(cam["lookat"], np.float32), up=np.array(cam["up"], np.float32), fovy=np.float32(40.0),
                  width=16, height=16)
        orc = oraclelib.OracleScene(ps)
        oid, ot = orc.trace(rays)
        assert (oid >= 0).sum() > n // 8
    else:
        oid, ot = np.full(n, -1, np.int32), np.full(n, trt.INF, np.float32)
    for flags in MODES.values():
        ids, t = dev.trace_closest(rays, flags)
        assert np.array_equal(ids, oid) and np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    img = dev.render(2, seed=1)  # no lights: everything is black, nothing hangs
    assert img.shape == (16, 16, 3) and np.all(img == 0)
    dev.close()


def test_scene_without_wide_layout(host_scenes, oracle_scenes, monkeypatch):
    """Scenes whose fast layout would be deeper than its stack keep the reference-topology kernels (every mode,
    the persistent kernel included, must then walk the reference tree).  Forced here with TRT_WIDE_SOURCE=off."""
    monkeypatch.setenv("TRT_WIDE_SOURCE", "off")
    dev = trt.DeviceScene(host_scenes["veach-mis"], 0)
    monkeypatch.delenv("TRT_WIDE_SOURCE")
    assert dev.stats()["accel_nodes"] == 390  # inner nodes of the reference tree (781 nodes, 391 leaves)
    rays = make_rays(host_scenes["veach-mis"], oracle_scenes["veach-mis"], 1 << 18, seed=21)
    oid, ot = oracle_scenes["veach-mis"].trace(rays)
    for flags in MODES.values():
        ids, t = dev.trace_closest(rays, flags)
        assert np.array_equal(ids, oid) and np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    ref, _ = oracle_scenes["veach-mis"].render(2, seed=8)
    img = dev.render(2, seed=8)
    assert np.sqrt(((img - ref) ** 2).mean()) <= 1e-3 * ref.mean()
    dev.close()


@pytest.mark.parametrize("scale", (1e-3, 1.0, 1e3, 1e5))
def test_scene_scale_and_far_origins(scale):
    """The fast layout's culling boxes are padded relative to the scene scale and rays from far origins take the
    strict walk: the same mesh at four magnitudes, with origins inside the scene, on its surfaces (bounce rays) and
    up to 100 scene sizes away, must still match the oracle bit for bit in every mode."""
    import oraclelib
    from tinyraytracing_b200 import workloads

    m = workloads.stress_mesh(40)
    v = (m["v9"].astype(np.float64) * scale).astype(np.float32)
    cam = {k: (tuple(np.float32(x) * np.float32(scale) for x in val) if k != "fovy" else val) for k, val in m["camera"].items()}
    host = trt.HostScene.from_arrays(v, m["mtl"], m["materials"], m["lights"], cam["eye"], cam["lookat"], cam["up"],
                                     cam["fovy"], 64, 36, vn9=m["vn9"])
    ps = dict(v=v, vn=m["vn9"], vt=np.zeros((len(v), 6), np.float32), mtl=m["mtl"],
              materials=[dict(m_, name=str(i)) for i, m_ in enumerate(m["materials"])], lights=m["lights"], textures=[],
              eye=np.array(cam["eye"], np.float32), lookat=np.array(cam["lookat"], np.float32),
              up=np.array(cam["up"], np.float32), fovy=np.float32(cam["fovy"]), width=64, height=36)
    orc = oraclelib.OracleScene(ps)
    dev = trt.DeviceScene(host, 0)
    rays = make_rays(host, orc, 1 << 18, seed=9)
    # far origins aimed back at the scene: 2 .. 100 scene sizes away
    rng = np.random.default_rng(3)
    lo, hi = host.root_box()
    centre, size = 0.5 * (lo + hi), float(np.max(hi - lo))
    n = 1 << 16
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    dist = size * 10 ** rng.uniform(0.3, 2.0, (n, 1))
    target = centre + rng.uniform(-0.4, 0.4, (n, 3)) * size
    o = target - d * dist
    far = np.concatenate([o, d], 1).astype(np.float32)
    rays = np.concatenate([rays, far])
    oid, ot = orc.trace(rays)
    if scale <= 1e3:  # at 1e5 every far hit lies beyond INF = 114514 and is a miss by bvh.cpp:219 — also worth testing
        assert (oid[-n:] >= 0).mean() > 0.3  # the far rays really hit the scene
    for name, flags in MODES.items():
        ids, t = dev.trace_closest(rays, flags)
        bad = np.flatnonzero(ids != oid)
        assert len(bad) == 0, (name, scale, len(bad), rays[bad[0]], ids[bad[0]], oid[bad[0]])
        assert np.array_equal(t.view(np.uint32), ot.view(np.uint32)), (name, scale)
    dev.close()


@pytest.mark.parametrize("name", SCENES)
def test_zero_direction_components_with_origins_on_box_planes(name, host_scenes, oracle_scenes, device_scenes):
    """Rays with one or two zero / denormal direction components (1/d overflows: the centre column of an axis-aligned
    camera) walk the fast layout with the reference's whole root-to-leaf box path as the acceptance gate.  The hard
    case is an origin coordinate EXACTLY on a reference box plane: (b - S) * inf = NaN inside interactAABB
    (bvh.cpp:231-245), whose glm min / max semantics decide whether the reference descends.  Every traversal mode must
    reproduce the oracle bit for bit."""
    dev, orc, host = device_scenes[name], oracle_scenes[name], host_scenes[name]
    boxes, links = host.nodes()
    lo, hi = host.root_box()
    rng = np.random.default_rng(17)
    n = 40000
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    axis = rng.integers(0, 3, n)
    d[np.arange(n), axis] = rng.choice(np.array([0.0, -0.0, 1e-42, -1e-42], np.float32), n)  # zero and denormal
    two = rng.random(n) < 0.25
    axis2 = (axis + 1 + rng.integers(0, 2, n)) % 3
    d[two, axis2[two]] = 0.0
    nz = np.abs(d).max(axis=1) > 0
    d[nz] /= np.linalg.norm(d[nz].astype(np.float64), axis=1, keepdims=True).astype(np.float32)
    d[np.arange(n), axis] = np.where(np.abs(d[np.arange(n), axis]) < 1e-30, d[np.arange(n), axis], 0.0)
    # half of the origins: the zeroed axis' coordinate is a plane of a reference node box, bit for bit
    pick = rng.integers(0, len(boxes), n)
    side = rng.integers(0, 2, n)
    plane = boxes[pick, axis + 3 * side]
    on_plane = (rng.random(n) < 0.5) & np.isfinite(plane)
    o[on_plane, axis[on_plane]] = plane[on_plane]
    rays = np.concatenate([o, d], 1).astype(np.float32)
    oid, ot = orc.trace(rays)
    assert (oid >= 0).sum() > n // 20
    for mode, flags in MODES.items():
        ids, t = dev.trace_closest(rays, flags)
        assert np.array_equal(ids, oid), (mode, int((ids != oid).sum()))
        assert np.array_equal(t.view(np.uint32), ot.view(np.uint32)), mode
