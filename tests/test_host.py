"""Host side kept by the drop-in surface (loaders + buildBVH, csrc/host) against the reference-generated golden
fixtures and against the oracle's independent restatement.  CPU only."""
import os

import numpy as np
import pytest

import oraclelib
import tinyraytracing_b200 as trt
from conftest import SCENES, SMALL_RES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("name", SCENES)
def test_host_build_matches_reference_golden(name, host_scenes):
    g = np.load(os.path.join(GOLD, name + "_closest.npz"))
    h = host_scenes[name]
    tr = h.triangles()
    boxes, links = h.nodes()
    # post-build order == the reference's std::sort sequence (compared by content: identical triangles are
    # interchangeable, the reference itself cannot tell them apart)
    ps = oraclelib.parsed_scene(name)
    for k in ("v", "vn", "vt"):
        assert np.array_equal(bits(ps[k][tr["face"]]), bits(ps[k][g["face"]])), k
    assert np.array_equal(ps["mtl"][tr["face"]], ps["mtl"][g["face"]])
    assert np.array_equal(bits(tr["normal"]), bits(g["normal"]))
    assert np.array_equal(bits(boxes), bits(g["node_box"])) and np.array_equal(links, g["node_link"])
    cam = h.camera()
    assert np.array_equal(bits(np.concatenate([cam["eye"], cam["llc"], cam["horizontal"], cam["vertical"]])), bits(g["camera"]))


@pytest.mark.parametrize("name", SCENES)
def test_host_build_matches_oracle(name, host_scenes, oracle_scenes):
    h, o = host_scenes[name], oracle_scenes[name]
    tr = h.triangles()
    assert np.array_equal(tr["face"], o.order())
    d = o.derived()
    assert np.array_equal(bits(tr["normal"]), bits(d["normal"]))
    ob, ol = o.nodes()
    hb, hl = h.nodes()
    assert np.array_equal(bits(hb), bits(ob)) and np.array_equal(hl, ol)
    assert np.array_equal(bits(np.concatenate(list(h.camera()[k] for k in ("eye", "llc", "horizontal", "vertical")))), bits(o.camera()))
    # lights: cumulative areas in OBJ order
    ls, lv, lvn, cum = h.lights()
    ps = o.ps
    names = h.material_names()
    for l, (mi, rad) in zip(ls, ps["lights"]):
        assert names[l["material"]] == ps["materials"][mi]["name"]
        mat = h.materials()[l["material"]]
        assert mat["is_emissive"] == 1 and np.allclose(mat["radiance"], rad, rtol=0, atol=0)
        sel = np.flatnonzero(ps["mtl"] == mi)
        assert l["n_tris"] == len(sel)
        assert np.array_equal(bits(lv[l["first_tri"]:l["first_tri"] + l["n_tris"]]), bits(ps["v"][sel]))
        seg = cum[l["first_tri"]:l["first_tri"] + l["n_tris"]]
        assert np.all(np.diff(seg) >= 0) and seg[-1] == mat["area"]
        assert np.array_equal(seg, d["area"][np.isin(tr["face"], sel)][np.argsort(tr["face"][np.isin(tr["face"], sel)])])


def test_textures_loaded(host_scenes):
    tex = host_scenes["staircase"].textures()
    assert sorted(t.shape for t in tex) == sorted([(894, 894, 3), (512, 512, 3), (1200, 1600, 3)])
    mats = {n: m for n, m in zip(host_scenes["staircase"].material_names(), host_scenes["staircase"].materials())}
    assert mats["Wood"]["texture"] >= 0 and mats["Glass"]["texture"] == -1 and abs(mats["Glass"]["Ni"] - 1.5) < 1e-7


def write(path, text):
    with open(path, "w") as f:
        f.write(text)


def test_loader_quirks(tmp_path):
    """Behaviours of the reference's hand-rolled loaders that scenes rely on."""
    d = str(tmp_path)
    write(d + "/q.xml", '<?xml version="1.0"?>\n<!-- c -->\n<camera type="perspective" width="32" height="16" fovy="40.5">\n'
          ' <eye x="0" y="1" z="-5"/> <lookat x="0" y="1" z="0"/> <up x="0" y="1" z="0"/>\n</camera>\n'
          '<light mtlname="L" radiance="1.5,\n    2.5,\n  3.5"/>\n<other/>\n<light mtlname="M" radiance="4, 5"/>\n')
    # vn before the first vt -> slots are read v/vn/vt (isvnvt); a quad: only the first three corners are used
    write(d + "/q.obj", "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 -1\nvn 0 0 1\nvt 0.25 0.75\nvt 0.5 0.5\n"
          "usemtl A\nf 1/1/2 2/1/2 3/1/2 4/1/2\nusemtl L\nf 1/2/1 3/2/1 4/2/1\nf 1/2/1 2/2/1 3/2/1\n")
    write(d + "/q.mtl", "newmtl A\nKd 0.1 0.2 0.3\nKs 0.4 0.5 0.6\nKt 0.9 0.9 0.9\nNs 12\nNi 1.25\nillum 2\nnewmtl L\nKd 0 0 0\nTr 0.7 0.8 0.9\n")
    h = trt.HostScene.load(d + "/q.xml", d + "/q.obj", d + "/q.mtl", d, leaf_num=8)
    assert h.n_tris == 3 and (h.desc.width, h.desc.height) == (32, 16)
    names = h.material_names()
    mats = dict(zip(names, h.materials()))
    assert mats["L"]["radiance"] == (1.5, 2.5, 3.5) and mats["L"]["is_emissive"] == 1
    assert mats["M"]["radiance"] == (4.0, 0.0, 5.0)  # two fields: x, then the LAST field lands in z (scene.cpp:31-49)
    assert mats["A"]["Tr"] == (0.0, 0.0, 0.0)  # `Kt` is ignored, only `Tr` is parsed (scene.cpp:90-94)
    assert mats["L"]["Tr"] == pytest.approx((0.7, 0.8, 0.9)) and mats["A"]["Ns"] == 12 and mats["A"]["Ni"] == 1.25
    tr = h.triangles()
    order = list(tr["face"])
    a = order.index(0)
    assert np.array_equal(tr["vn"][a].reshape(3, 3), np.tile([0, 0, -1], (3, 1)))  # slot 2 = vn index 1
    assert np.array_equal(tr["vt"][a].reshape(3, 2), np.tile([0.5, 0.5], (3, 1)))  # slot 3 = vt index 2
    assert np.array_equal(tr["v"][a].reshape(3, 3), [[0, 0, 0], [1, 0, 0], [1, 1, 0]])
    # cumulative light areas: two right triangles of area 0.5
    ls, lv, lvn, cum = h.lights()
    assert [names[l["material"]] for l in ls] == ["L", "M"] and ls[0]["n_tris"] == 2 and ls[1]["n_tris"] == 0
    assert cum == pytest.approx([0.5, 1.0]) and mats["L"]["area"] == pytest.approx(1.0)
    # standard order (vt before vn): slots are v/vt/vn
    write(d + "/s.obj", "v 0 0 0\nv 1 0 0\nv 1 1 0\nvt 0.25 0.75\nvt 0.5 0.5\nvn 0 0 -1\nvn 0 0 1\nusemtl A\nf 1/2/1 2/2/1 3/2/1\n")
    h2 = trt.HostScene.load(d + "/q.xml", d + "/s.obj", d + "/q.mtl", d)
    t2 = h2.triangles()
    assert np.array_equal(t2["vt"][0].reshape(3, 2), np.tile([0.5, 0.5], (3, 1))) and np.array_equal(t2["vn"][0][:3], [0, 0, -1])


def test_loader_errors_are_reported(tmp_path):
    d = str(tmp_path)
    write(d + "/q.xml", '<camera width="8" height="8" fovy="40"><eye x="0" y="0" z="0"/><lookat x="0" y="0" z="1"/><up x="0" y="1" z="0"/></camera>')
    write(d + "/bad.obj", "v 0 0 0\nf 1/1/1 2/1/1 3/1/1\n")
    write(d + "/q.mtl", "newmtl A\n")
    with pytest.raises(trt.TrtError):
        trt.HostScene.load(d + "/q.xml", d + "/bad.obj", d + "/q.mtl", d)
    with pytest.raises(trt.TrtError):
        trt.HostScene.load(d + "/q.xml", d + "/missing.obj", d + "/q.mtl", d)


@pytest.mark.parametrize("nq", (9, 40, 224))  # 224: 100 k triangles, the upper subtrees are built by forked tasks
def test_from_arrays_matches_oracle_build(nq):
    """Synthetic meshes (BASELINE config 5) enter through trt_host_scene_from_arrays: same derived fields and the
    same topology as the oracle's restatement of scene.cpp:196-205 + bvh.cpp:16-144."""
    from tinyraytracing_b200 import workloads

    m = workloads.stress_mesh(nq)
    cam = m["camera"]
    h = trt.HostScene.from_arrays(m["v9"], m["mtl"], m["materials"], m["lights"], cam["eye"], cam["lookat"], cam["up"],
                                  cam["fovy"], 64, 36, vn9=m["vn9"])
    ps = dict(v=m["v9"], vn=m["vn9"], vt=np.zeros((len(m["v9"]), 6), np.float32), mtl=m["mtl"],
              materials=[dict(m_, name=str(i)) for i, m_ in enumerate(m["materials"])], lights=m["lights"], textures=[],
              eye=np.array(cam["eye"], np.float32), lookat=np.array(cam["lookat"], np.float32),
              up=np.array(cam["up"], np.float32), fovy=np.float32(cam["fovy"]), width=64, height=36)
    o = oraclelib.OracleScene(ps)
    assert np.array_equal(h.triangles()["face"], o.order())
    hb, hl = h.nodes()
    ob, ol = o.nodes()
    assert np.array_equal(bits(hb), bits(ob)) and np.array_equal(hl, ol)
    assert np.array_equal(bits(h.triangles()["normal"]), bits(o.derived()["normal"]))


def test_png_writer_roundtrip(tmp_path):
    import cv2

    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    p = str(tmp_path / "x.png")
    assert trt.load_library().trt_write_png(p.encode(), 53, 37, img.ctypes.data, 0) == 0
    back = cv2.imread(p, cv2.IMREAD_COLOR)
    assert np.array_equal(back[:, :, ::-1], img)
    big = rng.integers(0, 256, (300, 300, 3), dtype=np.uint8)  # > 64 KiB: several stored deflate blocks
    assert trt.load_library().trt_write_png(p.encode(), 300, 300, big.ctypes.data, 0) == 0
    assert np.array_equal(cv2.imread(p, cv2.IMREAD_COLOR)[:, :, ::-1], big)


# ---- fast-layout invariants, checked on the host (trt_layout_check) -----------------------------------------------
def _random_soup(n_tris, seed, scale=1.0, degenerate=0):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, (n_tris, 1, 3)) * scale
    v = (c + rng.normal(size=(n_tris, 3, 3)) * 0.05 * scale).astype(np.float32)
    for k in range(min(degenerate, n_tris)):
        v[k, 2] = v[k, 1]  # zero-area triangle: scene.cpp:196 gives it a NaN face normal
    mats = [dict(Kd=(0.5, 0.5, 0.5), Ks=(0, 0, 0), Tr=(1, 1, 1), Ns=1.0, Ni=1.0)]
    return trt.HostScene.from_arrays(v.reshape(-1, 9), np.zeros(n_tris, np.int32), mats, [], (0.0, 0.0, -4.0 * scale),
                                     (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 40.0, 16, 16)


@pytest.mark.parametrize("name", SCENES)
def test_fast_layout_invariants_on_the_cg22_scenes(name, host_scenes):
    """Every triangle interactTriangle could accept is in exactly one leaf of the fast layout, tagged with its
    reference leaf (bvh.cpp:43-48), inside every box on its path with the pad; the tree fits the traversal stack."""
    rep = host_scenes[name].layout_check()
    assert rep["violations"] == 0 and rep["use_wide"] == 1
    assert rep["n_fast_tris"] + rep["n_dropped"] == rep["n_tris"]
    # staircase: 2 432 zero-area triangles (NaN normal) can never be hit and are left out
    assert rep["n_dropped"] == {"back": 0, "veach-mis": 0, "staircase": 2432}[name]
    assert 3 * rep["wide_depth"] + 2 <= 96


@pytest.mark.parametrize("n_tris", (0, 1, 2, 3, 8, 9, 17, 1000))
def test_fast_layout_invariants_on_small_and_degenerate_inputs(n_tris):
    rep = _random_soup(n_tris, seed=n_tris, degenerate=n_tris // 5).layout_check()
    assert rep["violations"] == 0
    assert rep["n_tris"] == n_tris and rep["n_fast_tris"] + rep["n_dropped"] == n_tris or rep["use_wide"] == 0
    if n_tris == 0:
        assert rep["use_wide"] == 0 and rep["wide_nodes"] == 0


@pytest.mark.parametrize("scale", (1e-3, 1.0, 1e3, 1e5))
def test_fast_layout_pad_follows_the_scene_scale(scale):
    """The pad is 256 ulp(scene scale): the checker recomputes it from the reference leaf boxes."""
    rep = _random_soup(500, seed=7, scale=scale).layout_check()
    assert rep["violations"] == 0 and rep["n_fast_tris"] == 500


def test_fast_layout_invariants_on_the_stress_mesh():
    from tinyraytracing_b200 import workloads

    m = workloads.stress_mesh(71)
    cam = m["camera"]
    h = trt.HostScene.from_arrays(m["v9"], m["mtl"], m["materials"], m["lights"], cam["eye"], cam["lookat"], cam["up"],
                                  cam["fovy"], 64, 36, vn9=m["vn9"])
    rep = h.layout_check()
    assert rep["violations"] == 0 and rep["n_fast_tris"] == rep["n_tris"] and rep["wide_nodes"] > 1000


def test_leaf_atomic_layout_invariants(host_scenes, monkeypatch):
    """TRT_WIDE_SOURCE=leaves (the first design, still selectable): reference leaves are the scan units."""
    monkeypatch.setenv("TRT_WIDE_SOURCE", "leaves")
    rep = host_scenes["veach-mis"].layout_check()
    assert rep["violations"] == 0 and rep["n_fast_tris"] == rep["n_tris"]
    monkeypatch.setenv("TRT_WIDE_SOURCE", "off")
    assert host_scenes["veach-mis"].layout_check()["use_wide"] == 0


def test_tree_optimisation_is_deterministic_and_never_worse(host_scenes, monkeypatch):
    """The insertion-based optimisation of the fast layout's tree keeps the binned-SAH tree when that collapses
    cheaper, has no randomness, and leaves the invariants intact (TRT_REINSERT=0 switches it off)."""
    for name in SCENES:
        monkeypatch.delenv("TRT_REINSERT", raising=False)
        a, b = host_scenes[name].layout_check(), host_scenes[name].layout_check()
        assert a == b and a["violations"] == 0
        monkeypatch.setenv("TRT_REINSERT", "0")
        off = host_scenes[name].layout_check()
        assert off["violations"] == 0 and off["n_fast_tris"] == a["n_fast_tris"]
        assert a["sah_wide"] <= off["sah_wide"] * (1 + 1e-12)
    monkeypatch.delenv("TRT_REINSERT", raising=False)


# ---- binary scene cache (trt_host_scene_save / trt_host_scene_load_cache) ----------------------------
def _same_scene(a, b):
    ta, tb = a.triangles(), b.triangles()
    assert all(np.array_equal(ta[k].view(np.uint8), tb[k].view(np.uint8)) for k in ta)
    (ba, la), (bb, lb) = a.nodes(), b.nodes()
    assert np.array_equal(bits(ba), bits(bb)) and np.array_equal(la, lb)
    assert a.material_names() == b.material_names() and a.materials() == b.materials()
    for x, y in zip(a.lights(), b.lights()):
        if isinstance(x, np.ndarray):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
        else:
            assert x == y
    ca, cb = a.camera(), b.camera()
    assert all(np.array_equal(np.asarray(ca[k]), np.asarray(cb[k])) for k in ca)
    assert a.layout_check() == b.layout_check()


@pytest.mark.parametrize("name", SCENES)
def test_scene_cache_roundtrip_is_bit_identical(name, host_scenes, tmp_path):
    p = str(tmp_path / (name + ".trtscn"))
    host_scenes[name].save(p)
    back = trt.HostScene.load_cache(p)
    _same_scene(host_scenes[name], back)
    ta, tb = host_scenes[name].textures(), back.textures()
    assert len(ta) == len(tb) == (3 if name == "staircase" else 0)  # decoded textures travel with the cache
    assert all(np.array_equal(x, y) for x, y in zip(ta, tb))
    back.close()


def test_scene_cache_rejects_damaged_files(host_scenes, tmp_path):
    p = str(tmp_path / "v.trtscn")
    host_scenes["veach-mis"].save(p)
    raw = open(p, "rb").read()
    cases = {"truncated": raw[: len(raw) // 2], "one flipped bit": raw[:1000] + bytes([raw[1000] ^ 1]) + raw[1001:],
             "foreign file": b"P6 1 1 255 " + bytes(64), "empty": b"", "wrong version": raw[:8] + bytes([9]) + raw[9:]}
    for what, data in cases.items():
        q = str(tmp_path / "bad.trtscn")
        open(q, "wb").write(data)
        with pytest.raises(trt.TrtError):
            trt.HostScene.load_cache(q)
    with pytest.raises(trt.TrtError):
        trt.HostScene.load_cache(str(tmp_path / "missing.trtscn"))


def test_pfm_writer_keeps_the_linear_image(tmp_path):
    """trt_write_pfm: the double[W*H*3] buffer of main.cpp:74 as 32-bit floats, rows bottom-to-top (PFM convention)."""
    rng = np.random.default_rng(1)
    img = rng.uniform(0, 40, (23, 31, 3))  # linear radiance, well above 1
    p = str(tmp_path / "x.pfm")
    assert trt.load_library().trt_write_pfm(p.encode(), 31, 23, np.ascontiguousarray(img).ctypes.data) == 0
    raw = open(p, "rb").read()
    head, rest = raw.split(b"\n", 3)[:3], raw.split(b"\n", 3)[3]
    assert head == [b"PF", b"31 23", b"-1.0"]
    back = np.frombuffer(rest, "<f4").reshape(23, 31, 3)[::-1]
    assert np.array_equal(back, img.astype(np.float32))
    assert trt.load_library().trt_write_pfm(p.encode(), 0, 23, np.ascontiguousarray(img).ctypes.data) != 0


def test_optimal_collapse_keeps_the_invariants_and_never_costs_more(host_scenes, monkeypatch):
    """TRT_COLLAPSE=optimal (experimental): the dynamic-programming collapse must give a valid layout whose summed
    child-box area is at most the greedy collapse's."""
    for name in SCENES:
        monkeypatch.delenv("TRT_COLLAPSE", raising=False)
        greedy = host_scenes[name].layout_check()
        monkeypatch.setenv("TRT_COLLAPSE", "optimal")
        opt = host_scenes[name].layout_check()
        assert opt["violations"] == 0 and opt["n_fast_tris"] == greedy["n_fast_tris"]
        assert opt["sah_wide"] <= greedy["sah_wide"] * (1 + 1e-6)
        assert opt["wide_nodes"] <= greedy["wide_nodes"]
    monkeypatch.delenv("TRT_COLLAPSE", raising=False)


@pytest.mark.parametrize("seed", (1, 2))
def test_layout_with_interpenetrating_lights(seed):
    """Several lights whose triangles are scattered among the occluders and each other (workloads.light_soup): the layout's
    invariants hold — among them that every triangle of a light's material lies in that light's box, which the early stop
    of occluded light samples rests on — and the CPU walk of the layout returns the oracle's ids and distances."""
    from tinyraytracing_b200 import workloads

    m = workloads.light_soup(seed)
    cam = m["camera"]
    h = trt.HostScene.from_arrays(m["v9"], m["mtl"], m["materials"], m["lights"], cam["eye"], cam["lookat"], cam["up"],
                                  cam["fovy"], m["width"], m["height"], vn9=m["vn9"])
    rep = h.layout_check()
    assert rep["violations"] == 0 and rep["use_wide"] == 1
    ps = dict(v=m["v9"], vn=m["vn9"], vt=np.zeros((len(m["v9"]), 6), np.float32), mtl=m["mtl"],
              materials=[dict(m_, name=str(i)) for i, m_ in enumerate(m["materials"])], lights=m["lights"], textures=[],
              eye=np.array(cam["eye"], np.float32), lookat=np.array(cam["lookat"], np.float32),
              up=np.array(cam["up"], np.float32), fovy=np.float32(cam["fovy"]), width=m["width"], height=m["height"])
    o = oraclelib.OracleScene(ps)
    rng = np.random.Generator(np.random.Philox(key=seed))
    rays = np.concatenate([workloads.camera_rays(h.camera(), 3000, rng), workloads.box_rays(*h.root_box(), 6000, rng)])
    ids, t, _ = oraclelib.walk_layout(h, rays)
    oid, ot = o.trace(rays)
    served = ids != -2
    assert served.mean() > 0.99
    assert np.array_equal(ids[served], oid[served]) and np.array_equal(t[served].view(np.uint32), ot[served].view(np.uint32))


# ---- layout cache (trt_layout_build_cached / trt_scene_create_cached, csrc/layout_cache.cu) ----------------------------
def _view_arrays(view):
    """Every array of a trt_layout_view as bytes (+ its scalar fields): what 'the same layouts' means."""
    import ctypes as C

    def raw(ptr, n_bytes):
        return C.string_at(ptr, n_bytes) if n_bytes else b""

    nf, nl, ni, nw = view.n_fast_tris, view.n_ref_leaves, view.n_ref_inner, view.n_wide_nodes
    return (
        (nw, view.wide_root, nf, nl, ni, view.check_leaf_box, view.strict_origin_limit, view.miss_key),
        raw(view.wide_nodes, nw * 128), raw(view.fast_geom, nf * 48), raw(view.fast_key, nf * 4), raw(view.fast_orig, nf * 4),
        raw(view.fast_leaf, nf * 4), raw(view.ref_leaf_box, nl * 32), raw(view.ref_leaf_parent, nl * 4), raw(view.ref_nodes, ni * 64),
    )


@pytest.mark.parametrize("name", SCENES)
def test_layout_cache_round_trip(name, host_scenes, tmp_path):
    """Built once, read back: bit-identical layouts; the second call reads the file instead of building."""
    host = host_scenes[name]
    path = str(tmp_path / (name + ".layout"))
    h0, v0 = host.layout_arrays()
    h1, v1, hit1 = host.layout_arrays(layout_cache=path)  # no file yet: built and written
    h2, v2, hit2 = host.layout_arrays(layout_cache=path)  # read back
    try:
        assert (hit1, hit2) == (False, True) and os.path.getsize(path) > 0
        assert _view_arrays(v0) == _view_arrays(v1) == _view_arrays(v2)
    finally:
        for h in (h0, h1, h2):
            host.free_layout(h)


def test_layout_cache_rejects_what_is_not_its_own(host_scenes, tmp_path, monkeypatch):
    """A stale, foreign, truncated or corrupt file is never an error and never used: the layouts are rebuilt and the file is
    rewritten.  'Stale' includes another scene and another setting of a switch that shapes the layout."""
    a, b = host_scenes["veach-mis"], host_scenes["back"]
    path = str(tmp_path / "x.layout")

    def build(host):
        h, v, hit = host.layout_arrays(layout_cache=path)
        arrays = _view_arrays(v)
        host.free_layout(h)
        return arrays, hit

    ref_a, hit = build(a)
    assert not hit and build(a) == (ref_a, True)
    # another scene under the same path: rebuilt for that scene, file now holds scene b
    ref_b, hit = build(b)
    assert not hit and ref_b != ref_a and build(b) == (ref_b, True)
    assert build(a) == (ref_a, False)
    # a switch that shapes the layout is part of the key
    monkeypatch.setenv("TRT_FAST_LEAF", "4")
    leaf4, hit = build(a)
    assert not hit and leaf4 != ref_a
    monkeypatch.delenv("TRT_FAST_LEAF")
    assert build(a) == (ref_a, False) and build(a) == (ref_a, True)
    good = open(path, "rb").read()
    for damage in (good[: len(good) // 2], good + b"\0", good[:100] + bytes([good[100] ^ 1]) + good[101:], b"", b"TRTLAYOT" + b"\0" * 64):
        open(path, "wb").write(damage)
        assert build(a) == (ref_a, False), len(damage)
        assert open(path, "rb").read() == good  # rewritten
    # an unwritable cache path costs nothing but the next build
    h, v, hit = a.layout_arrays(layout_cache=str(tmp_path / "no_such_dir" / "x.layout"))
    assert not hit and _view_arrays(v) == ref_a
    a.free_layout(h)
