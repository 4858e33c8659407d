"""The oracle (tier B, oracle/oracle.cpp) against what pins it: Philox known-answer vectors and the golden
fixtures generated from the unmodified reference (tier A) by tools/make_golden.py.  CPU only."""
import os

import numpy as np
import pytest

import oraclelib
from conftest import SCENES, SMALL_RES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert list(oraclelib.philox([0, 0, 0, 0], [0, 0])) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(oraclelib.philox([0xffffffff] * 4, [0xffffffff] * 2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert list(oraclelib.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    # uniforms: 53-bit doubles in [0,1), two per block
    o = oraclelib.philox([5, 6, 7, 4], [0x9abcdef0, 0x12345678])
    seed = 0x123456789abcdef0
    assert oraclelib.uniform(seed, 5, 6, 7, 8) == float((int(o[0]) << 21) | (int(o[1]) >> 11)) * 2.0 ** -53
    assert oraclelib.uniform(seed, 5, 6, 7, 9) == float((int(o[2]) << 21) | (int(o[3]) >> 11)) * 2.0 ** -53
    u = [oraclelib.uniform(1, p, 0, 0, 2) for p in range(2000)]
    assert 0 <= min(u) and max(u) < 1 and abs(np.mean(u) - 0.5) < 0.03


@pytest.mark.parametrize("name", SCENES)
def test_oracle_reproduces_reference_closest_hits(name, oracle_scenes):
    g = np.load(os.path.join(GOLD, name + "_closest.npz"))
    o = oracle_scenes[name]
    ids, t, pn, hp = o.trace(g["rays"], want_pn=True)
    assert np.array_equal(bits(t), bits(g["t"]))  # distance: bit-exact
    canon = g["canon"]  # identical-content triangles are one identity in the reference's HitRecord
    assert np.array_equal(np.where(ids >= 0, canon[np.maximum(ids, 0)], -1), g["id"])
    hit = ids >= 0
    assert np.array_equal(bits(hp[hit]), bits(g["hitpoint"][hit]))
    ok = np.isfinite(g["pn"][hit]).all(axis=1)
    assert np.abs(pn[hit][ok] - g["pn"][hit][ok]).max() <= 1e-6
    assert tuple(g["bvh_stats"]) == o.bvh_stats()


@pytest.mark.parametrize("name", SCENES)
def test_pruned_walk_equals_exhaustive_walk(name, oracle_scenes, host_scenes):
    from conftest import make_rays

    o = oracle_scenes[name]
    rays = make_rays(host_scenes[name], o, 1 << 17, seed=99)
    ids, _ = o.trace(rays)
    A1, T1, ids_pruned = o.trace_counts(rays, 1)
    A0, T0, ids_exh = o.trace_counts(rays, 0)
    assert np.array_equal(ids_pruned, ids) and np.array_equal(ids_exh, ids)
    assert A1 <= A0 and T1 <= T0


def test_primary_ray_mapping(oracle_scenes):
    """main.cpp:88-95: row 0 maps to y = H/(H-1) > 1; camera.cpp:19-28 normalises."""
    o = oracle_scenes["back"]
    cam = o.camera()
    eye, llc, hor, ver = cam[0:3], cam[3:6], cam[6:9], cam[9:12]
    W, H = o.width, o.height
    for (i, j, k) in ((0, 0, 0), (H - 1, W - 1, 3), (H // 2, W // 3, 1)):
        r = o.primary_ray(i, j, k, seed=17)
        pix = i * W + j
        x = j / (W - 1.0) + (oraclelib.uniform(17, pix, k, 0, 0) - 0.5) / W
        y = (H - i) / (H - 1.0) + (oraclelib.uniform(17, pix, k, 0, 1) - 0.5) / H
        d = llc + np.float32(x) * hor + np.float32(y) * ver - eye
        d = d / np.linalg.norm(d)
        assert np.array_equal(r[:3], eye) and np.allclose(r[3:], d, atol=2e-7)


def test_render_is_deterministic_and_shardable(oracle_scenes):
    o = oracle_scenes["veach-mis"]
    a, ca = o.render(4, seed=3)
    b, cb = o.render(4, seed=3, threads=2)
    assert np.array_equal(a, b) and ca == cb
    lo, c1 = o.render(4, seed=3, sample_begin=0, sample_end=1)
    hi, c2 = o.render(4, seed=3, sample_begin=1, sample_end=4)
    assert np.allclose(lo + hi, a, rtol=1e-12, atol=0) and (c1[0] + c2[0], c1[1] + c2[1]) == ca
    c, _ = o.render(4, seed=4)
    assert not np.array_equal(a, c)


def robust_stats(img, clip):
    c = np.minimum(img, clip)
    return c.mean(axis=(0, 1)), c


@pytest.mark.parametrize("name", SCENES)
def test_oracle_render_agrees_with_reference_statistically(name, oracle_scenes):
    """The reference's own shade() is only statistically reproducible (racy time-seeded engines).
    Noise-floor test on radiance clipped at 4x the image mean (NEE 1/r^2 fireflies dominate raw RMSE):
    channel means within 2 % of the reference's and RMSE(oracle, A1) <= 1.25 * RMSE(A1, A2)."""
    g = np.load(os.path.join(GOLD, name + "_render.npz"))
    a1, a2, spp = g["run1"].astype(np.float64), g["run2"].astype(np.float64), int(g["spp"])
    img, _ = oracle_scenes[name].render(spp, seed=1234)
    clip = 4 * a1.mean()
    m1, c1 = robust_stats(a1, clip)
    m2, c2 = robust_stats(a2, clip)
    mo, co = robust_stats(img, clip)
    ref_mean = 0.5 * (m1 + m2)
    assert np.all(np.abs(mo - ref_mean) <= 0.02 * ref_mean), (mo, m1, m2)
    floor = np.sqrt(((c1 - c2) ** 2).mean())
    assert np.sqrt(((co - c1) ** 2).mean()) <= 1.25 * floor
    assert np.sqrt(((co - c2) ** 2).mean()) <= 1.25 * floor


# ---- the closed-form barycentric solve of csrc/barycentric.cuh, restated in numpy ---------------------------------
def closed_form_bary(V, P):
    """Least-squares solution of [v0 v1 v2; 1 1 1] b = [P; 1] (triangle.cpp:12-29) without a factorisation: in-plane
    2x2 Gram solve for q = P - v0 and for v0, plus the sum-to-one slack u along the normal (derivation in the header
    of tinyraytracing_b200/csrc/barycentric.cuh, which computes exactly this)."""
    v0 = V[:, 0].astype(np.float64)
    e1, e2, q = V[:, 1].astype(np.float64) - v0, V[:, 2].astype(np.float64) - v0, P.astype(np.float64) - v0
    dot = lambda a, b: (a * b).sum(1)  # noqa: E731
    g11, g12, g22 = dot(e1, e1), dot(e1, e2), dot(e2, e2)
    inv = 1.0 / (g11 * g22 - g12 * g12)
    a1, a2, c1, c2 = dot(e1, q), dot(e2, q), dot(e1, v0), dot(e2, v0)
    p1, p2 = (a1 * g22 - a2 * g12) * inv, (a2 * g11 - a1 * g12) * inv
    w1, w2 = (c1 * g22 - c2 * g12) * inv, (c2 * g11 - c1 * g12) * inv
    n = np.cross(e1, e2)
    vn, qn, nn = dot(v0, n), dot(q, n), dot(n, n)
    u = vn * qn / (vn * vn + nn)
    b1, b2 = p1 - u * w1, p2 - u * w2
    return np.stack([1.0 + u - b1 - b2, b1, b2], 1)


def exact_bary(Vi, Pi):
    """The same least-squares problem by Cramer's rule on the normal equations in exact rational arithmetic."""
    from fractions import Fraction as Fr

    A = [[Fr(float(Vi[j][i])) for j in range(3)] for i in range(3)] + [[Fr(1)] * 3]
    c = [Fr(float(x)) for x in Pi] + [Fr(1)]
    G = [[sum(A[k][i] * A[k][j] for k in range(4)) for j in range(3)] for i in range(3)]
    h = [sum(A[k][i] * c[k] for k in range(4)) for i in range(3)]

    def det3(M):
        return (M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0])
                + M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]))

    D, out = det3(G), []
    for i in range(3):
        M = [row[:] for row in G]
        for r in range(3):
            M[r][i] = h[r]
        out.append(float(det3(M) / D))
    return out


def test_closed_form_barycentrics_against_exact_arithmetic():
    """0.01-sized triangles 1000 units from the origin, hit points rounded to float (so they sit off the plane, which
    is what makes the sum-to-one row matter): the closed form agrees with exact arithmetic to 1e-13, where a double
    SVD / QR solve of the 4x3 system is only good to ~1e-9 there."""
    rng = np.random.default_rng(0)
    n = 200
    c = rng.uniform(-1, 1, (n, 1, 3)) * 1000
    V = (c + rng.normal(size=(n, 3, 3)) * 0.01).astype(np.float32)
    b = rng.dirichlet((1, 1, 1), n)
    P = (V * b[:, :, None]).sum(1).astype(np.float32)
    ex = np.array([exact_bary(V[i], P[i]) for i in range(n)])
    assert np.abs(closed_form_bary(V, P) - ex).max() <= 1e-13
    lst = np.array([np.linalg.lstsq(np.vstack([V[i].T.astype(np.float64), np.ones(3)]),
                                    np.concatenate([P[i].astype(np.float64), [1.0]]), rcond=None)[0] for i in range(n)])
    assert np.abs(lst - ex).max() > 1e-12  # the factorisation route is the less accurate one


@pytest.mark.parametrize("name", SCENES)
def test_closed_form_barycentrics_give_the_reference_pn(name, oracle_scenes):
    """pn = normalize(sum vn_k b_k) (bvh.cpp:223-224) from the closed form equals the pn the unmodified reference
    computed through Eigen's QR (golden fixture), to the 1e-6 the GPU tests use."""
    g = np.load(os.path.join(GOLD, name + "_closest.npz"))
    o = oracle_scenes[name]
    ids, t, pn, hp = o.trace(g["rays"], want_pn=True)
    hit = ids >= 0
    ps = oraclelib.parsed_scene(name, *SMALL_RES[name])
    order = o.order()
    V = ps["v"].reshape(-1, 3, 3)[order][ids[hit]]
    VN = ps["vn"].reshape(-1, 3, 3)[order][ids[hit]]
    b = closed_form_bary(V, hp[hit]).astype(np.float32)
    m = (VN[:, 0] * b[:, :1] + VN[:, 1] * b[:, 1:2]) + VN[:, 2] * b[:, 2:3]
    mine = m / np.sqrt((m * m).sum(1, keepdims=True))
    ok = np.isfinite(g["pn"][hit]).all(axis=1) & np.isfinite(mine).all(axis=1)
    assert ok.sum() > 1000
    assert np.abs(mine[ok] - g["pn"][hit][ok]).max() <= 1e-6


@pytest.mark.parametrize("name", SCENES)
def test_oracle_matches_the_live_reference_including_nan_slabs(name, oracle_scenes, host_scenes, scene_files):
    """Tier B against tier A LIVE (oracle/_ref/libref.so = the unmodified reference sources, built in the build
    container): the config-2 ray population plus the rays whose treatment rests on the reference's NaN behaviour —
    zero / denormal direction components (1/d = +-inf) with an origin coordinate exactly on a node box plane, so that
    (b - S) * inf = NaN inside interactAABB (bvh.cpp:231-245) and glm's min / max decide.  The GPU path is validated
    against the oracle on the same kind of rays (tests/test_gpu_closest.py); this test pins the oracle's side."""
    import refbridge
    from conftest import make_rays

    if not refbridge.available():
        pytest.skip("oracle/_ref/libref.so not built")
    f = scene_files[name]
    ref = refbridge.RefScene(f["xml"], f["obj"], f["mtl"], f["basedir"])
    orc, host = oracle_scenes[name], host_scenes[name]
    rays = make_rays(host, orc, 60000, seed=23)
    boxes, _ = host.nodes()
    lo, hi = host.root_box()
    rng = np.random.default_rng(29)
    n = 40000
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    axis = rng.integers(0, 3, n)
    d[np.arange(n), axis] = rng.choice(np.array([0.0, -0.0, 1e-42, -1e-42], np.float32), n)
    pick, side = rng.integers(0, len(boxes), n), rng.integers(0, 2, n)
    plane = boxes[pick, axis + 3 * side]
    on_plane = (rng.random(n) < 0.5) & np.isfinite(plane)
    o[on_plane, axis[on_plane]] = plane[on_plane]
    rays = np.concatenate([rays, np.concatenate([o, d], 1)]).astype(np.float32)
    rt, rid = ref.trace(rays)
    oid, ot = orc.trace(rays)
    assert np.array_equal(ot.view(np.uint32), rt.view(np.uint32))
    assert np.array_equal(oid >= 0, rid >= 0) and (oid >= 0).sum() > 20000
    # the reference identifies a triangle by content (geometrically identical triangles are one identity)
    v = host.triangles()["v"]
    hit = oid >= 0
    assert np.array_equal(v[oid[hit]], v[rid[hit]])


# ---- the product's fast layout walked on the CPU (oracle/layout_walk.cpp) against the exhaustive reference walk ------
def _special_rays(host, n, seed):
    """Zero / denormal direction components, half of the origins exactly on a reference node box plane."""
    boxes, _ = host.nodes()
    lo, hi = host.root_box()
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    axis = rng.integers(0, 3, n)
    d[np.arange(n), axis] = rng.choice(np.array([0.0, -0.0, 1e-42, -1e-42], np.float32), n)
    pick, side = rng.integers(0, len(boxes), n), rng.integers(0, 2, n)
    plane = boxes[pick, axis + 3 * side]
    on_plane = (rng.random(n) < 0.5) & np.isfinite(plane)
    o[on_plane, axis[on_plane]] = plane[on_plane]
    return np.concatenate([o, d], 1).astype(np.float32)


@pytest.mark.parametrize("variant", ("default", "no-reinsertion", "optimal-collapse", "leaf-atomic"))
@pytest.mark.parametrize("name", SCENES)
def test_fast_layout_walked_on_the_cpu_equals_the_reference_walk(name, variant, oracle_scenes, host_scenes, monkeypatch):
    """The DATA of the GPU layout (boxes, leaf tags, tie keys, parent links — what trt_scene_create uploads), walked
    with the rules of DESIGN.md §3 restated on the CPU, gives the reference's triangle ids and distance bits: 100 k
    config-2 rays plus 40 k rays with zero direction components / origins on box planes per scene, for the default
    builder and its selectable variants.  No GPU involved; the CUDA kernels are tested against the same oracle."""
    from conftest import make_rays

    env = {"default": {}, "no-reinsertion": {"TRT_REINSERT": "0"}, "optimal-collapse": {"TRT_COLLAPSE": "optimal"},
           "leaf-atomic": {"TRT_WIDE_SOURCE": "leaves"}}[variant]
    for k in ("TRT_REINSERT", "TRT_COLLAPSE", "TRT_WIDE_SOURCE"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    host, orc = host_scenes[name], oracle_scenes[name]
    rays = np.concatenate([make_rays(host, orc, 100000, seed=41), _special_rays(host, 40000, seed=43)])
    ids, t, work = oraclelib.walk_layout(host, rays)
    oid, ot = orc.trace(rays)
    served = ids != -2
    assert served.all()  # finite rays from inside the scene: the fast layout serves every one of them
    assert np.array_equal(ids, oid) and np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    assert 0 < work["nodes"] < 60 and (oid >= 0).sum() > 30000


def test_cpu_walk_counters_are_the_gpu_counters_definition(host_scenes, oracle_scenes):
    """The walker counts what trt_trace_counters counts (wide nodes visited, child boxes tested, leaves scanned,
    triangles tested), so builder variants can be compared on real ray populations without a GPU: the optimised tree
    must not visit more nodes than the plain binned-SAH one on the config-2 population of staircase."""
    import os
    from conftest import make_rays

    host, orc = host_scenes["staircase"], oracle_scenes["staircase"]
    rays = make_rays(host, orc, 60000, seed=47)
    os.environ.pop("TRT_REINSERT", None)
    _, _, opt = oraclelib.walk_layout(host, rays)
    os.environ["TRT_REINSERT"] = "0"
    try:
        _, _, plain = oraclelib.walk_layout(host, rays)
    finally:
        os.environ.pop("TRT_REINSERT", None)
    assert opt["nodes"] < plain["nodes"] and opt["boxes"] < plain["boxes"]
