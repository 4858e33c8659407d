#!/usr/bin/env python3
"""Pack the reference's example-scenes-cg22 fixtures into scenes/<name>.npz (run in the build container).

/root/reference does not exist on the GPU box, so the scene inputs travel as *derived* fixtures: the OBJ /
MTL / XML text is parsed into arrays + a small JSON record here, and tinyraytracing_b200/scenes.py writes
equivalent OBJ / MTL / XML files back out (same statement order, float32-exact numbers) for the C++ loader.
tests/test_scenes.py checks (in this container) that the reference's own loader reads the regenerated
files to bit-identical triangles.  JPEG textures are carried as their original bytes; the decoded BGR crc32
(cv2.imread, i.e. OpenCV's decoder as the reference's material.cpp:6 uses) is recorded so the box can
verify its decode.

usage: python tools/pack_scenes.py [--ref /root/reference/RayTracingOnCPU/example-scenes-cg22]
"""
import argparse
import json
import os
import re
import sys
import zlib

import numpy as np

SCENES = {
    # name: (dir, stem)
    "back": ("test", "back"),
    "veach-mis": ("veach-mis", "veach-mis"),
    "staircase": ("staircase", "staircase"),
}


def parse_obj(path):
    v, vn, vt, faces, fmtl, mtls = [], [], [], [], [], []
    isvnvt = True
    cur = ""
    with open(path) as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                v.append([np.float32(x) for x in tok[1:4]])
            elif tok[0] == "vn":
                vn.append([np.float32(x) for x in tok[1:4]])
            elif tok[0] == "vt":
                if not vn:
                    isvnvt = False
                vt.append([np.float32(x) for x in tok[1:3]])
            elif tok[0] == "usemtl":
                cur = tok[1]
            elif tok[0] == "f":
                # the reference reads only the first three corners (scene.cpp:162)
                corners = [[int(s) for s in c.split("/")] for c in tok[1:4]]
                assert all(len(c) == 3 for c in corners), "only v/x/y faces occur in cg22"
                faces.append(corners)
                if cur not in mtls:
                    mtls.append(cur)
                fmtl.append(mtls.index(cur))
    return dict(
        v=np.array(v, np.float32).reshape(-1, 3),
        vn=np.array(vn, np.float32).reshape(-1, 3),
        vt=np.array(vt, np.float32).reshape(-1, 2),
        faces=np.array(faces, np.int32).reshape(-1, 3, 3),
        face_mtl=np.array(fmtl, np.int32),
        obj_mtl_names=mtls,
        isvnvt=bool(isvnvt),
    )


def parse_mtl(path):
    lines = []
    with open(path) as f:
        for line in f:
            tok = line.split()
            if tok:
                lines.append(tok)
    return lines


def parse_xml(path):
    txt = open(path).read()
    cam = re.search(r"<camera\b([^>]*)>", txt, re.S)
    attrs = dict(re.findall(r'(\w+)\s*=\s*"([^"]*)"', cam.group(1), re.S))
    rec = {"camera": attrs}
    for tag in ("eye", "lookat", "up"):
        m = re.search(r"<%s\b([^>]*)/>" % tag, txt, re.S)
        rec[tag] = dict(re.findall(r'(\w+)\s*=\s*"([^"]*)"', m.group(1), re.S))
    rec["lights"] = [
        dict(re.findall(r'(\w+)\s*=\s*"([^"]*)"', m, re.S)) for m in re.findall(r"<light\b([^>]*)/>", txt, re.S)
    ]
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference/RayTracingOnCPU/example-scenes-cg22")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "scenes"))
    a = ap.parse_args()
    import cv2

    os.makedirs(a.out, exist_ok=True)
    for name, (d, stem) in SCENES.items():
        base = os.path.join(a.ref, d)
        obj = parse_obj(os.path.join(base, stem + ".obj"))
        mtl = parse_mtl(os.path.join(base, stem + ".mtl"))
        xml = parse_xml(os.path.join(base, stem + ".xml"))
        tex = {}
        for tok in mtl:
            if tok[0] == "map_Kd":
                p = os.path.join(base, tok[1])
                raw = np.fromfile(p, np.uint8)
                img = cv2.imread(p)
                tex[tok[1]] = dict(rows=int(img.shape[0]), cols=int(img.shape[1]),
                                   crc32=int(zlib.crc32(img.tobytes())))
                obj["jpeg:" + tok[1]] = raw
        meta = dict(name=name, stem=stem, mtl=mtl, xml=xml, textures=tex,
                    obj_mtl_names=obj.pop("obj_mtl_names"), isvnvt=obj.pop("isvnvt"),
                    source="example-scenes-cg22/%s/%s.{obj,mtl,xml}" % (d, stem))
        out = os.path.join(a.out, name + ".npz")
        np.savez_compressed(out, meta=np.array(json.dumps(meta)), **obj)
        print("%s: %d faces, %d v, %d vn, %d vt, %d textures -> %s (%d bytes)" % (
            name, len(obj["faces"]), len(obj["v"]), len(obj["vn"]), len(obj["vt"]), len(tex), out,
            os.path.getsize(out)))


if __name__ == "__main__":
    sys.exit(main())
