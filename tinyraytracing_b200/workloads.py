"""Synthetic workloads of BASELINE.json: the fixed random-ray batch (config 2) and the procedural stress
mesh (config 5).  Pure numpy; no reference, no oracle."""
import numpy as np

RAY_SEED = 0x5EED0001


def _unit(v):
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def camera_rays(cam, n, rng):
    """Jittered pinhole rays with the reference's pixel mapping (main.cpp:88-95, camera.cpp:19-28), float32."""
    W, H = cam["width"], cam["height"]
    j = rng.integers(0, W, n)
    i = rng.integers(0, H, n)
    x = j / (W - 1.0) + (rng.random(n) - 0.5) / W
    y = (H - i) / (H - 1.0) + (rng.random(n) - 0.5) / H
    sx, sy = x.astype(np.float32)[:, None], y.astype(np.float32)[:, None]
    d = cam["llc"][None, :] + sx * cam["horizontal"][None, :] + sy * cam["vertical"][None, :] - cam["eye"][None, :]
    d = d.astype(np.float32)
    inv = (np.float32(1.0) / np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2], dtype=np.float32))
    d = d * inv[:, None]
    o = np.broadcast_to(cam["eye"][None, :], d.shape)
    return np.concatenate([o, d], 1).astype(np.float32)


def box_rays(lo, hi, n, rng):
    """Origins uniform in the root AABB, directions uniform on the sphere."""
    o = rng.uniform(lo, hi, (n, 3))
    d = _unit(rng.normal(size=(n, 3)))
    return np.concatenate([o, d], 1).astype(np.float32)


def bounce_rays(hitpoints, normals, rng):
    """Cosine-distributed rays leaving surface points (origin = hit point, no offset, as pathTracing.cpp:52,180)."""
    n = len(hitpoints)
    u1, u2 = rng.random(n), rng.random(n)
    r, phi = np.sqrt(u1), 2 * np.pi * u2
    lx, lz, ly = r * np.cos(phi), r * np.sin(phi), np.sqrt(1 - u1)
    nrm = normals.astype(np.float64)
    a = np.where(np.abs(nrm[:, :1]) > 0.9, np.array([[0.0, 1.0, 0.0]]), np.array([[1.0, 0.0, 0.0]]))
    t1 = _unit(np.cross(a, nrm))
    t2 = np.cross(nrm, t1)
    d = _unit(t1 * lx[:, None] + nrm * ly[:, None] + t2 * lz[:, None])
    return np.concatenate([hitpoints, d], 1).astype(np.float32)


def fixed_ray_batch(n, cam, root_box, tracer, seed=RAY_SEED):
    """BASELINE config 2 ray population: 25 % jittered camera rays, 50 % uniform-in-root-AABB
    origins with uniform directions, 25 % cosine bounce rays from the camera rays' hit points.
    tracer(rays) -> (tri_id, hitpoint, shading normal) supplies the surface points for the bounce class."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    n_cam = n // 4
    n_bounce = n // 4
    n_box = n - n_cam - n_bounce
    cam_r = camera_rays(cam, n_cam, rng)
    box_r = box_rays(root_box[0], root_box[1], n_box, rng)
    ids, hp, pn = tracer(cam_r)
    ok = np.flatnonzero((ids >= 0) & np.isfinite(pn).all(axis=1))
    if len(ok) == 0:
        bnc = box_rays(root_box[0], root_box[1], n_bounce, rng)
    else:
        pick = ok[rng.integers(0, len(ok), n_bounce)]
        nrm = pn[pick].astype(np.float64)
        # leave on the side the camera ray arrived from
        flip = np.sum(nrm * cam_r[pick, 3:6], axis=1) > 0
        nrm[flip] *= -1
        bnc = bounce_rays(hp[pick], nrm, rng)
    rays = np.concatenate([cam_r, box_r, bnc], 0)
    return np.ascontiguousarray(rays, np.float32)


def stress_mesh(nq, radius=3.0, center=(0.0, 3.0, 0.0), room=8.0):
    """BASELINE config 5: displaced UV sphere r = 1 + 0.15 sin 9θ sin 7φ + 0.05 sin(31θ + 3φ),
    nq x nq quads -> 2 nq² triangles, inside a 5-quad diffuse box with a 2-triangle ceiling light.
    Returns dict(v9, vn9, mtl, materials, lights, camera) for HostScene.from_arrays. All diffuse (Kd .7)."""
    th = np.linspace(0.0, np.pi, nq + 1)
    ph = np.linspace(0.0, 2 * np.pi, nq + 1)
    T, P = np.meshgrid(th, ph, indexing="ij")
    r = radius * (1 + 0.15 * np.sin(9 * T) * np.sin(7 * P) + 0.05 * np.sin(31 * T + 3 * P))
    pts = np.stack([r * np.sin(T) * np.cos(P), r * np.cos(T), r * np.sin(T) * np.sin(P)], -1) + np.array(center)
    nrm = _unit((pts - np.array(center)).reshape(-1, 3)).reshape(pts.shape)
    a, b, c, d = pts[:-1, :-1], pts[1:, :-1], pts[1:, 1:], pts[:-1, 1:]
    na, nb, nc, nd = nrm[:-1, :-1], nrm[1:, :-1], nrm[1:, 1:], nrm[:-1, 1:]
    v = np.concatenate([np.stack([a, b, c], 2).reshape(-1, 9), np.stack([a, c, d], 2).reshape(-1, 9)], 0)
    vn = np.concatenate([np.stack([na, nb, nc], 2).reshape(-1, 9), np.stack([na, nc, nd], 2).reshape(-1, 9)], 0)
    # interleave the two triangles of each quad (OBJ-like order)
    nqq = nq * nq
    order = np.arange(2 * nqq).reshape(2, nqq).T.reshape(-1)
    v, vn = v[order], vn[order]
    mtl = np.zeros(len(v), np.int32)
    R, Hh = room, 2 * room
    x0, x1, y0, y1, z0, z1 = -R, R, -0.5 * 0 - 1.0, Hh - 1.0, -R, R

    def quad(p0, p1, p2, p3, n):
        t = np.array([p0 + p1 + p2, p0 + p2 + p3], np.float64)
        return t, np.tile(np.array(n * 3, np.float64), (2, 1))

    walls = [quad([x0, y0, z0], [x1, y0, z0], [x1, y0, z1], [x0, y0, z1], [0, 1, 0]),   # floor
             quad([x0, y1, z0], [x0, y1, z1], [x1, y1, z1], [x1, y1, z0], [0, -1, 0]),  # ceiling
             quad([x0, y0, z1], [x1, y0, z1], [x1, y1, z1], [x0, y1, z1], [0, 0, -1]),  # back
             quad([x0, y0, z0], [x0, y0, z1], [x0, y1, z1], [x0, y1, z0], [1, 0, 0]),   # left
             quad([x1, y0, z0], [x1, y1, z0], [x1, y1, z1], [x1, y0, z1], [-1, 0, 0])]  # right
    lw = 0.35 * R
    light = quad([-lw, y1 - 0.01, -lw], [lw, y1 - 0.01, -lw], [lw, y1 - 0.01, lw], [-lw, y1 - 0.01, lw], [0, -1, 0])
    wv = np.concatenate([w[0] for w in walls], 0)
    wn = np.concatenate([w[1] for w in walls], 0)
    v = np.concatenate([v, wv, light[0]], 0).astype(np.float32)
    vn = np.concatenate([vn, wn, light[1]], 0).astype(np.float32)
    mtl = np.concatenate([mtl, np.full(len(wv), 1, np.int32), np.full(2, 2, np.int32)])
    diffuse = dict(Kd=(0.7, 0.7, 0.7), Ks=(0, 0, 0), Tr=(1, 1, 1), Ns=1.0, Ni=1.0)
    materials = [diffuse, diffuse, dict(Kd=(0, 0, 0), Ks=(0, 0, 0), Tr=(1, 1, 1), Ns=1.0, Ni=1.0)]
    camera = dict(eye=(0.0, 4.0, -R * 2.6), lookat=(0.0, 3.5, 0.0), up=(0.0, 1.0, 0.0), fovy=40.0)
    return dict(v9=v, vn9=vn, mtl=mtl, materials=materials, lights=[(2, (18.0, 18.0, 18.0))], camera=camera)


def light_soup(seed, n_occ=300, n_light_tris=24, n_lights=3, width=48, height=32):
    """A random triangle soup in which emitters and occluders interpenetrate: several lights whose triangles are
    scattered among (and behind, in front of, inside the boxes of) the occluders and of each other — the worst case for
    the shadow logic of the walk (search bound at the light point, early stop in front of the light's box)."""
    rng = np.random.default_rng(seed)

    def tris(n, spread, size):
        c = rng.uniform(-spread, spread, (n, 1, 3))
        return (c + rng.normal(size=(n, 3, 3)) * size).reshape(n, 9)

    v = [tris(n_occ, 1.0, 0.18)]
    mtl = [rng.integers(0, 2, n_occ)]
    for l in range(n_lights):
        v.append(tris(n_light_tris, 1.0, 0.12))
        mtl.append(np.full(n_light_tris, 2 + l))
    # a floor and a back wall so that most camera rays hit something
    v.append(np.array([[-3, -1.5, -3, 3, -1.5, -3, 3, -1.5, 3], [-3, -1.5, -3, 3, -1.5, 3, -3, -1.5, 3],
                       [-3, -1.5, 3, 3, -1.5, 3, 3, 3, 3], [-3, -1.5, 3, 3, 3, 3, -3, 3, 3]], np.float64))
    mtl.append(np.zeros(4, np.int64))
    v = np.concatenate(v).astype(np.float32)
    mtl = np.concatenate(mtl).astype(np.int32)
    e1, e2 = v[:, 3:6] - v[:, 0:3], v[:, 6:9] - v[:, 0:3]
    nrm = np.cross(e1, e2)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-20)
    vn = np.tile(nrm, (1, 3)).astype(np.float32)
    materials = [dict(Kd=(0.7, 0.6, 0.5), Ks=(0, 0, 0), Tr=(1, 1, 1), Ns=1.0, Ni=1.0),
                 dict(Kd=(0.2, 0.3, 0.4), Ks=(0.5, 0.5, 0.5), Tr=(1, 1, 1), Ns=60.0, Ni=1.0)]
    materials += [dict(Kd=(0, 0, 0), Ks=(0, 0, 0), Tr=(1, 1, 1), Ns=1.0, Ni=1.0) for _ in range(n_lights)]
    lights = [(2 + l, (20.0 + 5 * l, 18.0, 12.0 - 2 * l)) for l in range(n_lights)]
    cam = dict(eye=(0.0, 0.3, -4.5), lookat=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), fovy=45.0)
    return dict(v9=v, vn9=vn, mtl=mtl, materials=materials, lights=lights, camera=cam, width=width, height=height)
