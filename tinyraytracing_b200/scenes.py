"""Scene fixtures: write the packed example-scenes-cg22 inputs back out as OBJ / MTL / XML files.

The reference takes its inputs as files (stdin protocol, main.cpp:46-55; loaders scene.cpp:3-213) and so
does this framework's C++ host (csrc/host/scene.cpp).  /root/reference is not available on the GPU box, so
`scenes/<name>.npz` (made by tools/pack_scenes.py) carries the parsed content and this module regenerates
equivalent text files: same statement order (so the `isvnvt` slot-order quirk of scene.cpp:149-153 is
preserved), numbers printed with %.9g so every float32 round-trips exactly.

Textures: the reference decodes JPEG with cv::imread (material.cpp:6).  The C++ loader decodes baseline JPEG itself,
bit-identically (csrc/host/jpeg_decoder.cpp); `materialize` also writes the cv2-decoded side-car `<texture>.bgr`
("BGR8", int32 rows, int32 cols, bytes) that Material::readinMap falls back to for other formats, and that the
oracle's shim imread uses.
"""
import json
import os
import zlib

import numpy as np

SCENE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scenes")
NAMES = ("back", "veach-mis", "staircase")


def _g(x):
    return "%.9g" % float(x)


def write_bgr_sidecar(path, img):
    """img: HxWx3 uint8 BGR (cv2.imread layout) -> `path` side-car understood by Material::readinMap."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    with open(path, "wb") as f:
        f.write(b"BGR8")
        f.write(np.array([img.shape[0], img.shape[1]], np.int32).tobytes())
        f.write(img.tobytes())


def materialize(name, outdir, width=None, height=None):
    """Write <outdir>/<stem>.{obj,mtl,xml} (+ textures) for packed scene `name`.

    width/height override the XML resolution (the reference takes resolution from the XML, scene.cpp:13-14).
    Returns dict(basedir, obj, mtl, xml, width, height).
    """
    z = np.load(os.path.join(SCENE_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    stem = meta["stem"]
    os.makedirs(outdir, exist_ok=True)
    v, vn, vt, faces, fmtl = z["v"], z["vn"], z["vt"], z["faces"], z["face_mtl"]

    with open(os.path.join(outdir, stem + ".obj"), "w") as f:
        f.write("# regenerated from scenes/%s.npz (%s)\n" % (name, meta["source"]))
        for p in v:
            f.write("v %s %s %s\n" % (_g(p[0]), _g(p[1]), _g(p[2])))

        def emit_vt():
            for p in vt:
                f.write("vt %s %s\n" % (_g(p[0]), _g(p[1])))

        def emit_vn():
            for p in vn:
                f.write("vn %s %s %s\n" % (_g(p[0]), _g(p[1]), _g(p[2])))

        # scene.cpp:149-153: slots are read v/vn/vt unless a vt line is seen before any vn line
        if meta["isvnvt"]:
            emit_vn(), emit_vt()
        else:
            emit_vt(), emit_vn()
        cur = None
        names = meta["obj_mtl_names"]
        for tri, m in zip(faces, fmtl):
            if m != cur:
                cur = m
                if names[m] != "":
                    f.write("usemtl %s\n" % names[m])
            f.write("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % tuple(int(x) for x in tri.reshape(-1)))

    with open(os.path.join(outdir, stem + ".mtl"), "w") as f:
        for tok in meta["mtl"]:
            f.write(" ".join(tok) + "\n")

    xml = meta["xml"]
    cam = dict(xml["camera"])
    if width is not None:
        cam["width"] = str(int(width))
    if height is not None:
        cam["height"] = str(int(height))

    def attrs(d):
        return " ".join('%s="%s"' % kv for kv in d.items())

    with open(os.path.join(outdir, stem + ".xml"), "w") as f:
        f.write('<?xml version="1.0" encoding="utf-8"?>\n<camera %s>\n' % attrs(cam))
        for tag in ("eye", "lookat", "up"):
            f.write("\t<%s %s/>\n" % (tag, attrs(xml[tag])))
        f.write("</camera>\n")
        for l in xml["lights"]:
            f.write("<light %s/>\n" % attrs(l))

    for rel, info in meta["textures"].items():
        import cv2

        p = os.path.join(outdir, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        raw = z["jpeg:" + rel]
        raw.tofile(p)
        img = cv2.imdecode(raw, cv2.IMREAD_COLOR)
        if img is None or img.shape[0] != info["rows"] or img.shape[1] != info["cols"]:
            raise RuntimeError("texture decode failed: " + rel)
        if zlib.crc32(img.tobytes()) != info["crc32"]:
            # a different libjpeg build may differ in low bits; parity is then statistical only
            import warnings

            warnings.warn("decoded texture %s differs from the pack-time decode (crc32)" % rel)
        write_bgr_sidecar(p + ".bgr", img)

    return dict(basedir=outdir, obj=os.path.join(outdir, stem + ".obj"), mtl=os.path.join(outdir, stem + ".mtl"),
                xml=os.path.join(outdir, stem + ".xml"), width=int(cam["width"]), height=int(cam["height"]))


def parsed(name, width=None, height=None):
    """Parsed (pre-build, OBJ-order) scene arrays straight from scenes/<name>.npz — independent of both the
    product's C++ loader and the reference's: numbers are converted with numpy float32 (== stof)."""
    import cv2

    z = np.load(os.path.join(SCENE_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    faces = z["faces"]  # (n, 3 corners, 3 slots v/vt/vn or v/vn/vt)
    v, vn, vt = z["v"], z["vn"], z["vt"]
    s2, s3 = (2, 1) if not meta["isvnvt"] else (1, 2)  # column of vn, column of vt
    tri_v = v[faces[:, :, 0] - 1].reshape(-1, 9)
    tri_vn = vn[faces[:, :, s2] - 1].reshape(-1, 9)
    tri_vt = vt[faces[:, :, s3] - 1].reshape(-1, 6)
    # materials: names in order of first appearance: lights (xml), obj usemtl, mtl newmtl
    xml = meta["xml"]
    names = []
    for l in xml["lights"]:
        if l["mtlname"] not in names:
            names.append(l["mtlname"])
    for nme in meta["obj_mtl_names"]:
        if nme not in names:
            names.append(nme)
    mats = {}
    cur = ""
    tex_files = []
    for tok in meta["mtl"]:
        k = tok[0]
        if k == "newmtl":
            cur = tok[1]
        elif k in ("Kd", "Ks", "Tr"):
            mats.setdefault(cur, {})[k] = [np.float32(x) for x in tok[1:4]]
        elif k in ("Ns", "Ni"):
            mats.setdefault(cur, {})[k] = np.float32(tok[1])
        elif k == "map_Kd":
            mats.setdefault(cur, {})["map_Kd"] = tok[1]
            if tok[1] not in tex_files:
                tex_files.append(tok[1])
    for nme in mats:
        if nme not in names:
            names.append(nme)
    textures = [cv2.imdecode(z["jpeg:" + f], cv2.IMREAD_COLOR) for f in tex_files]
    materials = []
    for nme in names:
        m = mats.get(nme, {})
        materials.append(dict(name=nme, Kd=m.get("Kd", [0, 0, 0]), Ks=m.get("Ks", [0, 0, 0]), Tr=m.get("Tr", [0, 0, 0]),
                              Ns=m.get("Ns", 1.0), Ni=m.get("Ni", 1.0),
                              texture=tex_files.index(m["map_Kd"]) if "map_Kd" in m else -1))
    face_mtl = np.array([names.index(meta["obj_mtl_names"][i]) for i in z["face_mtl"]], np.int32)
    lights = []
    for l in xml["lights"]:
        rs = l["radiance"].split(",")
        lights.append((names.index(l["mtlname"]), [np.float32(x) for x in rs[:3]]))
    f3 = lambda d: np.array([np.float32(d["x"]), np.float32(d["y"]), np.float32(d["z"])], np.float32)
    cam = xml["camera"]
    return dict(name=name, v=tri_v, vn=tri_vn, vt=tri_vt, mtl=face_mtl, materials=materials, lights=lights,
                textures=textures, eye=f3(xml["eye"]), lookat=f3(xml["lookat"]), up=f3(xml["up"]),
                fovy=np.float32(cam["fovy"]), width=int(width or cam["width"]), height=int(height or cam["height"]))


def cornell_with_standin(nq=224, width=512, height=512):
    """BASELINE configs 1/2 name the cornell-box scene, whose mesh (cornell-box.obj, the dragon) is a missing blob in the
    reference checkout.  STAND-IN, clearly not the reference's geometry: the Cornell shell `test/back` (same camera,
    materials, light) plus a procedurally displaced sphere of 2*nq^2 triangles (nq = 224 -> 100 352) in material
    DiffuseWhite standing on the floor.  Returns the arguments of HostScene.from_arrays."""
    from . import workloads

    ps = parsed("back", width, height)
    m = workloads.stress_mesh(nq, radius=110.0, center=(278.0, 135.0, 300.0))
    n_sphere = 2 * nq * nq
    white = [i for i, mm in enumerate(ps["materials"]) if mm["name"].endswith("DiffuseWhite")][0]
    v = np.concatenate([ps["v"], m["v9"][:n_sphere]]).astype(np.float32)
    vn = np.concatenate([ps["vn"], m["vn9"][:n_sphere]]).astype(np.float32)
    vt = np.concatenate([ps["vt"], np.zeros((n_sphere, 6), np.float32)])
    mtl = np.concatenate([ps["mtl"], np.full(n_sphere, white, np.int32)]).astype(np.int32)
    return dict(v9=v, vn9=vn, vt6=vt, mtl=mtl, materials=ps["materials"], lights=ps["lights"], eye=ps["eye"],
                lookat=ps["lookat"], up=ps["up"], fovy=float(ps["fovy"]), width=width, height=height)
